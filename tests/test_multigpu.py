"""Multi-GPU (x-slab, NCCL) parity: decomposed CUDA run == undecomposed oracle, bit for bit."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def ngpus():
    # ask the driver, not torch: importing torch after libludwig_b200 has bound the system libnccl can fail
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=60).stdout
        return sum(1 for line in out.splitlines() if line.startswith("GPU "))
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("periodic,reduced,peer,binary", [("111", 0, 1, 1), ("111", 0, 0, 1), ("011", 0, 1, 1), ("111", 1, 1, 1),
                                                         ("111", 0, 1, 0), ("111", 0, 0, 0)])
def test_slab_decomposition_matches_single_domain_oracle(periodic, reduced, peer, binary):
    """x-slabs on 2 (4) GPUs == the undecomposed oracle, bit for bit, with the planes of lb200_step travelling by
    NVLink peer stores from inside the kernels (peer = 1) or by NCCL send/recv (peer = 0); binary and single fluid."""
    n = ngpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(HERE, "multigpu_parity.py"),
           periodic, str(reduced), str(peer), str(binary)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0
    assert "MISMATCH" not in r.stdout and "OK" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("le,peer,fast", [(1, 0, 0), (2, 0, 0), (1, 1, 1), (1, 0, 1), (2, 1, 1)])
def test_lees_edwards_slabs_match_single_domain_oracle(le, peer, fast):
    """Lees-Edwards sheared binary fluid (BASELINE config 5) on x-slabs, `le` planes per GPU: lb200_step and the
    individual entry points over 2 (4) GPUs == the undecomposed Lees-Edwards oracle -- strict mode bit for bit
    (reference step structure, NCCL halo planes); fast mode within tolerance on the halo-free step with the planes
    of phi, u_x and f travelling by NVLink peer stores (peer = 1) or NCCL (peer = 0)."""
    n = ngpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29535", os.path.join(HERE, "multigpu_parity.py"),
           "111", "0", str(peer), "1", str(le), str(fast)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0
    assert "MISMATCH" not in r.stdout and "OK" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("peer", [1, 0])
def test_one_kernel_step_on_slabs_matches_single_domain_oracle(peer):
    """fast mode, no planes: the one-kernel step (LB200_KNOB_FUSED; TMA boxes of f read the x halo planes WITH their y / z
    rims, which the neighbour's kernel stores over NVLink, or which NCCL brings) over 2 (4) GPUs, mixed with the
    individual entry points, within tolerance of the undecomposed oracle"""
    n = ngpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29539", os.path.join(HERE, "multigpu_parity.py"),
           "111", "0", str(peer), "1", "0", "1"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0
    assert "MISMATCH" not in r.stdout and "OK" in r.stdout


@pytest.mark.gpu
def test_conserve_2_all_gather_on_slabs(monkeypatch):
    """cahn_hilliard_options_conserve 2 over 2 (4) GPUs: the one global reduction of the path (NCCL all-gather of one pair per
    GPU, added in rank order) -- initial sum and corrected steps against the undecomposed oracle"""
    n = ngpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    env = dict(os.environ, LB200_TEST_CONSERVE2="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(HERE, "multigpu_parity.py"),
           "111", "0", "0", "1"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0
    assert "MISMATCH" not in r.stdout and "OK" in r.stdout


@pytest.mark.gpu
def test_liquid_crystal_slabs_match_single_domain_oracle():
    """liquid crystal (Q tensor + Beris-Edwards) on x-slabs: lb200_step_lc and the individual entry points over
    2 (4) GPUs == the undecomposed liquid-crystal oracle, bit for bit (strict mode, NCCL x-planes of q, u, f)."""
    n = ngpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29537", os.path.join(HERE, "multigpu_parity.py"),
           "111", "0", "0", "1", "0", "0", "1"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0
    assert "MISMATCH" not in r.stdout and "OK" in r.stdout
