#!/usr/bin/env python3
"""bench.py -- MLUPS of the D3Q19 symmetric binary-fluid time step (BASELINE.json metric).

One "step" = one full Ludwig time step of the hot path (reference src/ludwig.c:528-860: phi halo,
27-pt gradient, stress-divergence force, Cahn-Hilliard update with 3rd-order advection, u halo,
pull-stream + MRT collision, distribution halo) on a 256^3 lattice PER GPU (x-slab decomposition,
weak scaling), FP64, synthetic spinodal initial state.

  python bench.py --gpus N --steps K --warmup W              our arm (CUDA, libludwig_b200.so)
  python bench.py --impl reference --gpus N --steps K ...    the reference's own CPU code on the
                                                             host cores (oracle/_ref, else the C port)

Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_ALG = {"step_binary": 496.0, "collide": 360.0, "force_ch": 96.0, "grad": 40.0}
BINARY = dict(a=-0.00625, b=0.00625, kappa=0.004, mobility=1.25)
ETA = 0.00625
ADV_ORDER = 3


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            if "hbm_gbs" in d:
                return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (started before the warm-up so
    that the tool's own start-up time does not eat a short timed region; only samples whose timestamp
    falls inside [mark_begin, mark_end] are reported, falling back to every sample taken under load)."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.path = None
        self.t0 = self.t1 = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
            # nvidia-smi needs ~1 s to produce its first line: wait for it so a short run is still sampled
            t_end = time.time() + 5.0
            while time.time() < t_end and os.path.getsize(self.path) == 0:
                time.sleep(0.02)
        except Exception:
            self.proc = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = []
        try:
            for line in open(self.path):
                t = [x.strip() for x in line.split(",")]
                if len(t) < 10:
                    continue
                try:
                    ts = datetime.datetime.strptime(t[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    rows.append((ts, float(t[2]), float(t[3]), float(t[4]), t[6:10]))
                except ValueError:
                    continue
            os.unlink(self.path)
        except Exception:
            pass
        inside = [r for r in rows if self.t0 is not None and self.t1 is not None and self.t0 <= r[0] <= self.t1 + 0.02]
        where = "timed region"
        if not inside:
            # very short timed region: every sample taken under load (warm-up included)
            inside = [r for r in rows if r[3] > 300.0] or rows
            where = "warm-up + timed region (under load)"
        if inside:
            sm = sorted(r[1] for r in inside)
            reasons = set()
            for r in inside:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(r[2] for r in inside), reasons=sorted(reasons),
                       samples=len(inside), window=where, power_w_max=max(r[3] for r in inside))
        return out


CV19 = [(0, 0, 0), (1, 1, 0), (1, 0, 1), (1, 0, 0), (1, 0, -1), (1, -1, 0), (0, 1, 1), (0, 1, 0), (0, 1, -1),
        (0, 0, 1), (0, 0, -1), (0, -1, 1), (0, -1, 0), (0, -1, -1), (-1, 1, 0), (-1, 0, 1), (-1, 0, 0),
        (-1, 0, -1), (-1, -1, 0)]


# ------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation (oracle/_ref) or the C port (oracle/), host cores
# ------------------------------------------------------------------------------------------------------

def host_threads():
    """CPUs this process may run on (the box's host cores); what the CPU arm's OpenMP team is set to"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def flowing_state(n, nhalo):
    """f = equilibrium(rho(x), u(x)) on the interior of an n^3 lattice (canonical allocated layout, halos zero):
    rho = 1 + 0.01 sin(2 pi x) cos(2 pi y), u = 0.01 (sin 2 pi y, sin 2 pi z, sin 2 pi x)"""
    import numpy as np
    na = n + 2 * nhalo
    f = np.zeros((19, na, na, na))
    c = np.arange(1, n + 1) / n
    x, y, z = np.meshgrid(c, c, c, indexing="ij")
    rho = 1.0 + 0.01 * np.sin(2 * np.pi * x) * np.cos(2 * np.pi * y)
    u = (0.01 * np.sin(2 * np.pi * y), 0.01 * np.sin(2 * np.pi * z), 0.01 * np.sin(2 * np.pi * x))
    uu = u[0] * u[0] + u[1] * u[1] + u[2] * u[2]
    for p, cv in enumerate(CV19):
        w = (12.0 if p == 0 else 2.0 if sum(abs(a) for a in cv) == 1 else 1.0) / 36.0
        cu = cv[0] * u[0] + cv[1] * u[1] + cv[2] * u[2]
        f[p, nhalo:-nhalo, nhalo:-nhalo, nhalo:-nhalo] = rho * w * (1.0 + 3.0 * cu + 4.5 * cu * cu - 1.5 * uu)
    return f.reshape(19, -1)


def cpu_steps_per_second(n, nsteps, warm=1, keep=False):
    """(seconds per step, kind, threads[, state]) for an n^3 binary-fluid lattice on the host CPU.  The OpenMP team is
    set through the library (omp_set_num_threads) to the host core count whatever OMP_NUM_THREADS says -- torchrun
    exports OMP_NUM_THREADS=1 -- and `threads` is what omp_get_max_threads() then reports.  keep: also return the
    initial (f, phi) and final (f, phi, u) arrays of the reference run (canonical layout) for bench.py's parity check."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    want = host_threads()
    os.environ["OMP_NUM_THREADS"] = str(want)
    import refharness
    if refharness.available(fast=True):
        threads = refharness.omp_threads(want, fast=True) or want
        sim = refharness.RefSim((n, n, n), nhalo=2, have_phi=1, adv_order=ADV_ORDER, eta_shear=ETA,
                                ghost_off=1, fast=True, **BINARY)
        sim.init_rest(1.0)
        sim.init_spinodal(8361235, 0.0, 0.05)
        state = None
        if keep:
            # for the parity check the sample starts from a flowing state (|u| = 0.01, rho = 1 +- 0.01) instead of rest, so
            # that the velocity is compared at a magnitude where a RELATIVE error means something (from rest |u| ~ 1e-6
            # after a few steps, i.e. round-off of the O(1) populations / |u| ~ 1e-10)
            f0 = flowing_state(n, 2)
            sim.set(refharness.REF_F, f0)
            state = {"f0": f0, "phi0": sim.get(refharness.REF_PHI)}
        sim.step(warm)
        t = sim.time_steps(nsteps)
        if keep:
            state.update(f=sim.get(refharness.REF_F), phi=sim.get(refharness.REF_PHI), u=sim.get(refharness.REF_U),
                         nsteps=warm + nsteps, n=n)
        sim.close()
        return (t / nsteps, "reference", threads, state) if keep else (t / nsteps, "reference", threads)
    import numpy as np
    from oracle import Oracle
    orc = Oracle((n, n, n), nhalo=2)
    rng = np.random.default_rng(8361235)
    f = orc.equilibrium(1.0, (0.0, 0.0, 0.0))
    phi = np.zeros((1, orc.nsites))
    orc.interior(phi)[0] = 0.05 * (rng.random((n, n, n)) - 0.5)
    z3 = lambda: np.zeros((3, orc.nsites))
    z1 = lambda: np.zeros((1, orc.nsites))
    u, rho, force, grad, delsq = z3(), z1(), z3(), z3(), z1()
    cp = orc.collide_param(0, 1.0, ETA)
    sp = orc.symm_param(adv_order=ADV_ORDER, **BINARY)
    if keep:
        f = flowing_state(n, 2)
    state = {"f0": f.copy(), "phi0": phi.copy()} if keep else None
    orc.step(cp, sp, 1, warm, f, phi, u, rho, force, grad, delsq)
    t0 = time.perf_counter()
    orc.step(cp, sp, 1, nsteps, f, phi, u, rho, force, grad, delsq)
    t = (time.perf_counter() - t0) / nsteps
    if keep:
        state.update(f=f, phi=phi, u=u, nsteps=warm + nsteps, n=n)
        return t, "port", 1, state
    return t, "port", 1


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    # bounded sample: the largest n^3 (<= 256) for which steps + warmup finish in ~150 s
    t64, kind, threads = cpu_steps_per_second(64, 2, warm=1)
    per_site = t64 / 64 ** 3
    total = args.steps + args.warmup
    n = 64
    for cand in (96, 128, 192, 256):
        if per_site * cand ** 3 * total <= 150.0:
            n = cand
    t0 = time.perf_counter()
    tstep, kind, threads = cpu_steps_per_second(n, args.steps, warm=max(args.warmup, 1))
    mlups = n ** 3 / tstep / 1e6
    line = {
        "impl": "reference", "metric": "MLUPS", "value": mlups, "unit": "MLUPS", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": tstep * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "D3Q19 symmetric binary fluid (spinodal), 27pt gradient, stress-divergence force, "
                               "Cahn-Hilliard advection order 3, MRT(M10) collision + propagation + halos; "
                               f"CPU sample lattice {n}^3 (target workload 256^3 per GPU)"},
        "cpu_baseline": {"value": mlups, "unit": "MLUPS", "cores": threads, "kind": kind,
                         "sample": f"{args.steps} full time steps of a {n}^3 lattice after {max(args.warmup, 1)} warm-up "
                                   f"steps, OpenMP team of {threads} threads (omp_get_max_threads)"},
        "e2e": {"value": mlups, "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------

def rel_err_interior(sim, a, b):
    """max |a - b| over the interior sites / max |b| (arrays in the canonical allocated layout)"""
    import numpy as np
    ai, bi = sim.interior(a), sim.interior(b)
    scale = float(np.abs(bi).max())
    d = float(np.abs(ai - bi).max())
    return d / scale if scale > 0 else d


def check_against_reference(lb, state, device, math):
    """The GPU library on the CPU leg's own lattice: same initial (f, phi), same parameters, same number of steps;
    returns max relative errors of f, phi, u against the arrays the reference's own code produced."""
    n = state["n"]
    with lb.Lb200((n, n, n), nhalo=2, have_phi=True, math=math, device=device) as sim:
        sim.put(lb.F, state["f0"]); sim.put(lb.PHI, state["phi0"])
        sim.step(lb.CollideParam.make(lb.RELAX_M10, 1.0, ETA), lb.SymmParam.make(adv_order=ADV_ORDER, **BINARY), state["nsteps"])
        out = {"f": rel_err_interior(sim, sim.get(lb.F), state["f"]),
               "phi": rel_err_interior(sim, sim.get(lb.PHI), state["phi"]),
               "u": rel_err_interior(sim, sim.get(lb.U), state["u"])}
    return out


def check_decomposition(lb, dist, torch, rank, local_rank, world, nsteps=8, nxl=32, ny=48, nz=64):
    """x-slab run on `world` GPUs (nxl planes each) against the SAME library on one GPU over the undecomposed lattice:
    strict mode must agree bit for bit, fast mode within 1e-12 (x-chunk boundaries of the phi sector fall on different
    planes in the two runs).  Returns {"multi_gpu_bit_exact": bool, "multi_gpu_max_rel_err_fast": float} on rank 0."""
    import numpy as np
    nh = 2
    nglob = (nxl * world, ny, nz)
    nall_g = tuple(m + 2 * nh for m in nglob)
    nall_l = (nxl + 2 * nh, ny + 2 * nh, nz + 2 * nh)
    rng = np.random.default_rng(424242)
    wv = np.array([12.0] + [2.0 if sum(abs(c) for c in cv) == 1 else 1.0 for cv in CV19[1:]]) / 36.0
    f = np.zeros((19,) + nall_g)
    f[:, nh:-nh, nh:-nh, nh:-nh] = wv[:, None, None, None] * (1.0 + 1e-3 * (rng.random((19,) + nglob) - 0.5))
    phi = np.zeros((1,) + nall_g)
    phi[:, nh:-nh, nh:-nh, nh:-nh] = 0.05 * (rng.random((1,) + nglob) - 0.5)

    def slab(a):
        out = np.zeros((a.shape[0],) + nall_l)
        out[:, nh:nh + nxl] = a[:, nh + rank * nxl:nh + (rank + 1) * nxl]
        return out.reshape(a.shape[0], -1)

    cp = lb.CollideParam.make(lb.RELAX_M10, 1.0, ETA, force=(1e-6, -2e-6, 5e-7))
    sp = lb.SymmParam.make(adv_order=ADV_ORDER, **BINARY)
    arrays = (("f", lb.F), ("phi", lb.PHI), ("u", lb.U))
    out = {}
    for math, key in ((lb.MATH_STRICT, "strict"), (lb.MATH_FAST, "fast")):
        sim = lb.Lb200((nxl, ny, nz), nhalo=nh, have_phi=True, math=math, device=local_rank, cart_size=world, cart_rank=rank)
        ids = [sim.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        sim.nccl_init(ids[0], world, rank)
        sim.put(lb.F, slab(f)); sim.put(lb.PHI, slab(phi))
        sim.step(cp, sp, nsteps)
        mine = {k: np.ascontiguousarray(sim.interior(sim.get(a))) for k, a in arrays}
        mode = sim.exchange_mode()
        sim.close()
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        if rank == 0:
            with lb.Lb200(nglob, nhalo=nh, have_phi=True, math=math, device=local_rank) as one:
                one.put(lb.F, f.reshape(19, -1)); one.put(lb.PHI, phi.reshape(1, -1))
                one.step(cp, sp, nsteps)
                single = {k: np.ascontiguousarray(one.interior(one.get(a))) for k, a in arrays}
            same, err = True, 0.0
            for k, _ in arrays:
                full = np.concatenate([g[k] for g in gathered], axis=1)
                same = same and bool(np.array_equal(full, single[k]))
                scale = float(np.abs(single[k]).max())
                err = max(err, float(np.abs(full - single[k]).max()) / (scale if scale > 0 else 1.0))
            if key == "strict":
                out["multi_gpu_bit_exact"] = same
            else:
                out["multi_gpu_max_rel_err_fast"] = err
                out["multi_gpu_bit_exact_fast"] = same
            out["multi_gpu_check"] = (f"{world} x-slabs of {nxl} planes ({nglob[0]}x{ny}x{nz} lattice), {nsteps} steps, exchange mode {mode}, "
                                      "against one GPU over the undecomposed lattice")
    return out


def numa_bind(torch, local_rank):
    """Pin this process to the CPUs of the NUMA node of its GPU (sysfs), so that host buffers allocated and first touched
    next are local to the GPU's PCIe root.  Returns what numa_unbind needs; (None, None) where the topology is not exposed."""
    try:
        old = os.sched_getaffinity(0)
        prop = torch.cuda.get_device_properties(local_rank)
        bdf = f"{prop.pci_domain_id:04x}:{prop.pci_bus_id:02x}:{prop.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        if node < 0:
            return (None, None)
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= old
        if not cpus:
            return (None, None)
        os.sched_setaffinity(0, cpus)
        return (old, node)
    except Exception:
        return (None, None)


def numa_unbind(state):
    if state and state[0] is not None:
        try:
            os.sched_setaffinity(0, state[0])
        except Exception:
            pass


def secondary_block(lb, torch, dist, rank, local_rank, world, steps=20, warmup=3):
    """The other BASELINE / SURVEY configurations, measured in the same run (device-resident, CUDA events on the library's
    stream, max over ranks): short, so that the default bench stays within minutes.  Each entry: lattice per GPU, ms per step,
    whole-job MLUPS, SURVEY's algorithmic bytes per site and the fraction of the measured HBM peak they amount to."""
    import numpy as np
    from ludwig_b200.initial import lc_twist_q
    peak, _ = measured_peaks()
    wv = np.array([12.0] + [2.0 if sum(abs(c) for c in cv) == 1 else 1.0 for cv in CV19[1:]]) / 36.0
    out = {}

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def run(name, nlocal, b_alg, note, kind="binary", math=None, decomposed=True, **kw):
        try:
            w = world if decomposed else 1
            if not decomposed and rank != 0:
                return
            sim = lb.Lb200(nlocal, nhalo=(1 if kind == "single" else 2), have_phi=(kind == "binary"), have_q=(kind == "lc"),
                           math=lb.MATH_FAST if math is None else math, device=local_rank, cart_size=w, cart_rank=(rank if decomposed else 0), **kw)
            if w > 1:
                ids = [sim.nccl_unique_id() if rank == 0 else None]
                dist.broadcast_object_list(ids, src=0)
                sim.nccl_init(ids[0], w, rank)
            f = np.empty((19, sim.nsites_lb))
            f[...] = wv[:, None]
            sim.put(lb.F, f)
            del f
            rng = np.random.default_rng(99 + rank)
            cp = lb.CollideParam.make(lb.RELAX_M10, 1.0, ETA if kind != "lc" else 0.1)
            if kind == "binary":
                phi = np.zeros((1, sim.nsites))
                sim.interior(phi)[...] = 0.05 * (rng.random((1,) + tuple(nlocal)) - 0.5)
                sim.put(lb.PHI, phi)
                sp = lb.SymmParam.make(adv_order=ADV_ORDER, **BINARY)
                step = lambda k: sim.step(cp, sp, k)
            elif kind == "lc":
                q = np.array(lc_twist_q(nlocal, 2, LC["q0"], 1.0 / 3.0, 2))
                sim.interior(q)[...] += 0.01 * (rng.random((5,) + tuple(nlocal)) - 0.5)
                sim.put(lb.Q, q)
                lc = lb.LcParam.make(adv_order=3, **LC)
                step = lambda k: sim.step_lc(cp, lc, k)
            else:
                step = lambda k: sim.step(cp, None, k)
            stream = torch.cuda.ExternalStream(sim.stream(), device=torch.device("cuda", local_rank))
            step(warmup)
            sim.sync()
            if dist is not None and decomposed:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            step(steps)
            e1.record(stream)
            sim.sync()
            ms = e0.elapsed_time(e1)
            if decomposed:
                ms = max_over_ranks(ms)
            sites = float(nlocal[0]) * nlocal[1] * nlocal[2]
            mlups = sites * w * steps / (ms * 1e-3) / 1e6
            out[name] = {"lattice_per_gpu": list(nlocal), "n_gpus": w, "ms_per_step": ms / steps, "mlups": mlups,
                         "algorithmic_bytes_per_site": b_alg, "frac_of_hbm_peak": mlups / w * 1e6 * b_alg / 1e9 / peak,
                         "exchange": sim.exchange_mode(), "note": note}
            sim.close()
        except Exception as exc:                 # a secondary line never sinks the headline
            out[name] = {"failed": str(exc)}

    n = 256
    run("binary_256_strict", (n, n, n), 496.0, "the headline workload in the bit-exact arithmetic mode (two kernels per step)", math=lb.MATH_STRICT)
    run("single_fluid_d3q19_256", (n, n, n), 360.0, "config 1 family at 256^3 per GPU: MRT(M10) pull-stream-collide, one kernel per step", kind="single")
    run("single_fluid_d3q19_64", (64, 64, 64), 360.0, "BASELINE config 1 (64^3, one GPU): launch-latency bound", kind="single", decomposed=False)
    run("liquid_crystal_128", (128, 128, 128), 672.0, "BASELINE config 4: Q tensor + Beris-Edwards + D3Q19, 128^3 per GPU (weak)", kind="lc")
    if (2 * n) % world == 0 and 8 % world == 0:
        run("binary_strong_512x256x256", (2 * n // world, n, n), 496.0, "SURVEY 8(d) config 3, strong scaling: 512x256x256 over all GPUs")
        run("lees_edwards_512x256x256", (2 * n // world, n, n), 496.0,
            "BASELINE config 5: sheared binary fluid, 512x256x256 over all GPUs, 8 Lees-Edwards planes, plane speed 0.05",
            le_nplanes=8, le_uy=0.05)
    return out


def run_ours(args, rank, local_rank, world):
    import numpy as np
    import torch
    import ludwig_b200 as lb

    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)

    n = args.size
    nlocal = (n, n, n)
    if args.strong:
        # strong scaling: a fixed 2n x n x n lattice cut into `world` x-slabs
        assert (2 * n) % world == 0, "--strong needs 2*size divisible by the number of GPUs"
        nlocal = (2 * n // world, n, n)
    local_sites = float(nlocal[0]) * nlocal[1] * nlocal[2]
    nhalo = 2
    # --le P: Lees-Edwards sheared binary fluid (BASELINE config 5), P planes per GPU, plane speed 0.05
    sim = lb.Lb200(nlocal, nhalo=nhalo, have_phi=True, math=lb.MATH_STRICT if args.strict else lb.MATH_FAST,
                   device=local_rank, cart_size=world, cart_rank=rank, le_nplanes=args.le * world, le_uy=0.05)
    if world > 1:
        ids = [sim.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        sim.nccl_init(ids[0], world, rank)

    if args.f32:
        sim.set_knob(lb.KNOB_F32, 1)
    if args.pipe >= 2 and world == 1:
        sim.set_knob(lb.KNOB_PIPE_SMS, args.pipe_sms)
        sim.set_knob(lb.KNOB_PIPE, args.pipe)
    ns = sim.nsites
    nall = sim.nall
    # pinned host state (the reference's host arrays: lb->f, phi->data, hydro->u, hydro->rho), allocated and first touched
    # on the NUMA node this GPU hangs off (8 ranks uploading 2.8 GB each through one socket's memory is what limits the
    # end-to-end figure at N = 8); the affinity is restored afterwards: the CPU baseline wants every core
    numa = numa_bind(torch, local_rank)
    h_f = torch.empty((19, sim.nsites_lb), dtype=torch.float64, pin_memory=True)
    h_phi = torch.empty((1, ns), dtype=torch.float64, pin_memory=True)
    h_u = torch.empty((3, ns), dtype=torch.float64, pin_memory=True)
    h_rho = torch.empty((1, ns), dtype=torch.float64, pin_memory=True)
    wv = np.array([12.0] + [2.0 if sum(abs(c) for c in cv) == 1 else 1.0 for cv in CV19[1:]]) / 36.0
    fv = h_f.numpy()
    for p in range(19):
        fv[p, :] = wv[p]                         # rho = 1, u = 0 equilibrium
    rng = np.random.default_rng(8361235 + rank)
    h_phi.numpy()[...] = 0.0
    pv = h_phi.numpy()[0, :sim.nsites_lb].reshape(nall)
    pv[nhalo:-nhalo, nhalo:-nhalo, nhalo:-nhalo] = 0.05 * (rng.random(nlocal) - 0.5)

    h_u.zero_(); h_rho.zero_()
    numa_unbind(numa)
    cp = lb.CollideParam.make(lb.RELAX_M10, 1.0, ETA)
    sp = lb.SymmParam.make(adv_order=ADV_ORDER, **BINARY)
    stream = torch.cuda.ExternalStream(sim.stream(), device=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        sim.sync()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    H2D, D2H = 1, 2

    def upload():
        sim.memcpy_async(lb.F, h_f.data_ptr(), H2D)
        sim.memcpy_async(lb.PHI, h_phi.data_ptr(), H2D)

    def download():
        sim.memcpy_async(lb.PHI, h_phi.data_ptr(), D2H)
        sim.memcpy_async(lb.U, h_u.data_ptr(), D2H)
        sim.memcpy_async(lb.RHO, h_rho.data_ptr(), D2H)

    # ---- device-resident timing: inputs in HBM before the timed region -------------------------------
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    upload()
    sim.sync()
    sim.step(cp, sp, args.warmup)
    barrier()
    clocks.mark_begin()
    l0 = sim.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    sim.step(cp, sp, args.steps)
    e1.record(stream)
    barrier()
    clocks.mark_end()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = sim.launch_count() - l0
    clk = clocks.stop() if rank == 0 else None

    phi_sum = float(np.nansum(sim.interior(sim.get(lb.PHI))))      # sanity: finite, conserved
    sites_total = local_sites * world
    mlups = sites_total * args.steps / (ms * 1e-3) / 1e6

    # ---- per-kernel device time (CUDA events on the launching stream), short separate pass -----------
    sim.profile(True)
    sim.step(cp, sp, min(args.steps, 20))
    sim.sync()
    prof = sim.profile_get()
    sim.profile(False)
    kernels = {k: {"ms_per_launch": (t / c if c else None), "launches": c} for k, (t, c) in prof.items()}
    peak, peak_src = measured_peaks()
    col_ms = kernels["collide"]["ms_per_launch"]
    # Bytes the kernels must move.  SURVEY 8(d): collide 360, gradient 40 + force/Cahn-Hilliard 96, step 496 B/site.  Inside a
    # multi-step call hydro->rho (8) and grad / delsq (32) are stored by the LAST step only (nobody reads them in between), so
    # the per-launch average over the profiled call of kp steps is used for the per-kernel rooflines (conservative: fewer
    # bytes for the same time); the whole-step figure stays SURVEY's 496 B/site, the definition of BASELINE's roofline MLUPS.
    kp = min(args.steps, 20)
    lazy = os.environ.get("LB200_LAZY_DIAG", "1") != "0" and not args.le
    b_col = 208.0 if args.f32 else B_ALG["collide"]            # f32 storage: 19 x 4 x 2 + 56
    b_step = B_ALG["step_binary"] - (B_ALG["collide"] - b_col)
    if lazy:
        b_col -= 8.0 * (kp - 1) / kp
    ach = b_col * local_sites / (col_ms * 1e-3) / 1e9 if col_ms else None
    ps_ms = kernels.get("phi_sector", {}).get("ms_per_launch")
    ps_alg = B_ALG["grad"] + B_ALG["force_ch"]            # SURVEY 8(d) sweeps A + B, done here in one kernel
    if lazy:
        ps_alg -= 32.0 * (kp - 1) / kp                     # grad + delsq stay in registers except on the last step
    ps_ach = ps_alg * local_sites / (ps_ms * 1e-3) / 1e9 if ps_ms else None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("collide_bytes_per_launch_256" if lazy else "earlier_capture_every_array_stored")
            if isinstance(traffic, dict):
                traffic = traffic.get("collide_bytes_per_launch_256")
            if nlocal != (256, 256, 256) or args.f32:
                traffic = None
        except Exception:
            traffic = None

    # One kernel per step (LB200_KNOB_FUSED, the default where it applies): the dominant kernel IS the step.  Its
    # algorithmic bytes are SURVEY 8(d)'s 496 B/site for the three sweeps it replaces (minus the rho / grad / delsq stores
    # of the intermediate steps); the bytes it really moves are 368 B/site (f 152 + 152, phi 8 + 8, u 24 + 24).
    fu_ms = kernels.get("step_fused", {}).get("ms_per_launch")
    fused = bool(fu_ms) and kernels["step_fused"]["launches"] >= kernels["collide"]["launches"]
    fu_alg = b_step - (40.0 * (kp - 1) / kp if lazy else 0.0)
    fu_ach = fu_alg * local_sites / (fu_ms * 1e-3) / 1e9 if fu_ms else None
    fu_traffic = None
    if fused and os.path.exists(tpath):
        try:
            fu_traffic = json.load(open(tpath)).get("step_fused_bytes_per_launch_256") if nlocal == (256, 256, 256) else None
        except Exception:
            fu_traffic = None

    # ---- end to end through the C-ABI with HOST buffers: H2D state, K steps, D2H observables ----------
    barrier()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    t0.record(stream)
    upload()
    sim.step(cp, sp, args.steps)
    download()
    t1.record(stream)
    sim.sync()
    wall = time.perf_counter() - w0
    barrier()
    e2e_ms = max_over_ranks(max(t0.elapsed_time(t1), wall * 1e3))
    e2e = sites_total * args.steps / (e2e_ms * 1e-3) / 1e6
    h2d = (19 * sim.nsites_lb + ns) * 8 / args.steps
    d2h = (1 + 3 + 1) * ns * 8 / args.steps

    exch_mode = sim.exchange_mode()
    pipe_st = sim.pipe_state()
    sim.close()
    del h_f, h_phi, h_u, h_rho

    # ---- CPU baseline (rank 0) and parity of what was just timed ---------------------------------------
    # check.max_rel_err_*: the library on the CPU leg's own lattice (same initial state, same steps) against the arrays
    # the reference's own code produced; at N > 1 also the decomposed run against one GPU on the undecomposed lattice.
    cpu = None
    check = {"phi_sum": phi_sum}
    if rank == 0 and not args.no_cpu:
        try:
            nsamp = args.cpu_size
            tstep, kind, threads, state = cpu_steps_per_second(nsamp, 3, warm=1, keep=True)
            cpu = {"value": nsamp ** 3 / tstep / 1e6, "unit": "MLUPS", "cores": threads, "kind": kind,
                   "sample": f"3 full time steps of a {nsamp}^3 lattice (same physics, same parameters) after 1 warm-up step, "
                             f"OpenMP team of {threads} threads"}
            err = check_against_reference(lb, state, local_rank, lb.MATH_STRICT if args.strict else lb.MATH_FAST)
            check.update({"max_rel_err_f": err["f"], "max_rel_err_phi": err["phi"], "max_rel_err_u": err["u"],
                          "against": f"the {'compiled reference' if kind == 'reference' else 'C port of the reference'} on {nsamp}^3, "
                                     f"{state['nsteps']} steps from the same initial state (the cpu_baseline leg's run)",
                          "tolerance": 1e-12, "ok": bool(max(err.values()) <= 1e-12)})
            del state
        except Exception as exc:          # the baseline is reported, never allowed to sink the bench
            cpu = {"value": None, "unit": "MLUPS", "cores": 0, "kind": "port", "sample": f"failed: {exc}"}
    if world > 1 and not args.no_check:
        try:
            check.update(check_decomposition(lb, dist, torch, rank, local_rank, world))
        except Exception as exc:
            check["multi_gpu_check"] = f"failed: {exc}"

    secondary = None
    if args.secondary:
        secondary = secondary_block(lb, torch, dist, rank, local_rank, world)

    if rank == 0:
        line = {
            "metric": "MLUPS", "value": mlups, "unit": "MLUPS", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if args.strong else "weak",
            "vs_baseline": None, "dtype": "f64 arithmetic, f32 storage of the distributions" if args.f32 else "f64", "data": "synthetic",
            "config": {"workload": f"D3Q19 symmetric binary fluid (spinodal decomposition), {nlocal[0]}x{nlocal[1]}x{nlocal[2]} per GPU, "
                                   "27pt phi gradient + stress-divergence force + Cahn-Hilliard (advection order 3) "
                                   "+ MRT(M10) pull-stream-collide; no halo sweeps (periodic images read in-kernel / stored by the "
                                   "producing kernel), x-planes over NVLink when sharded",
                       "lattice_per_gpu": list(nlocal), "decomposition": f"{world}_1_1 x-slabs",
                       "lees_edwards": (f"{args.le * world} planes, plane speed 0.05 (steady shear): reference-structured step with "
                                        "halo kernels + plane patches (not the halo-free path)" if args.le else "none"),
                       "math": "strict" if args.strict else "fast(fma)",
                       "kernels_per_step": ("1 (step_fused)" if fused else "2 (phi_sector + collide)"),
                       "x_plane_exchange": {0: "none (one GPU)", 1: "NCCL send/recv on a second stream",
                                            2: "NVLink peer stores from inside the kernels + flags"}[exch_mode],
                       "distribution_storage": ("f32 (float(f_p - w_p), FP64 arithmetic; LB200_KNOB_F32: error bound in "
                                                "tests/test_gpu_parity.py::test_f32_storage_error_bound)" if args.f32 else "f64"),
                       "slab_pipeline": (lambda st: {"slabs": args.pipe, "mode": {1: "green contexts", 2: "priority streams"}.get(st[0], "off"),
                                                     "sms_phi_sector": st[1][0], "sms_collide": st[1][1]})(pipe_st),
                       "diagnostic_stores": ("hydro->rho, grad, delsq stored by the last step of each lb200_step call only (LB200_LAZY_DIAG=1)"
                                             if lazy else "every step"),
                       "l2": "inputs (2.8 GB of lattice state per sweep) exceed the 126 MB L2; no flush needed",
                       "e2e_protocol": "pinned-host f+phi -> device, K steps, phi+u+rho -> pinned host "
                                       "(the reference's own lb_memcpy/field_memcpy usage, src/ludwig.c:501-506, 985)",
                       "host_buffers_numa_node": (numa[1] if numa else None)},
            "roofline": ({"bound": "hbm", "kernel": "step_fused_ws (27pt gradient + stress-divergence force + Cahn-Hilliard + pull-stream + MRT "
                                                    "collision in one sweep: warp-specialised phi-sector / collision / TMA-loader warps)",
                          "achieved": fu_ach, "peak": peak, "unit": "GB/s", "frac": (fu_ach / peak if fu_ach else None),
                          "traffic": fu_traffic, "peak_source": peak_src,
                          "algorithmic_bytes_per_site": fu_alg,
                          "real_bytes_per_site": 368.0,
                          "achieved_real_bytes": 368.0 * local_sites / (fu_ms * 1e-3) / 1e9,
                          "note": "algorithmic = SURVEY 8(d)'s three sweeps (496 B/site) this kernel replaces; it moves 368 B/site",
                          "whole_step": {"algorithmic_bytes_per_site": b_step,
                                         "achieved": mlups / world * 1e6 * b_step / 1e9,
                                         "frac": mlups / world * 1e6 * b_step / 1e9 / peak}} if fused else
                         {"bound": "hbm", "kernel": "collide_d3q19 (pull-stream + MRT collision)",
                          "achieved": ach, "peak": peak, "unit": "GB/s", "frac": (ach / peak if ach else None),
                          "traffic": traffic, "peak_source": peak_src,
                          "algorithmic_bytes_per_site": b_col,
                          "second_kernel": {"kernel": "phi_sector (27pt gradient + stress-divergence force + Cahn-Hilliard)",
                                            "algorithmic_bytes_per_site": ps_alg, "achieved": ps_ach,
                                            "frac": (ps_ach / peak if ps_ach else None),
                                            "note": "issue/latency-bound FP64 stencil, not HBM-bound (ncu: profiles/)"},
                          "whole_step": {"algorithmic_bytes_per_site": b_step,
                                         "achieved": mlups / world * 1e6 * b_step / 1e9,
                                         "frac": mlups / world * 1e6 * b_step / 1e9 / peak}}),
            "kernels": kernels,
            "cpu_baseline": cpu,
            "clocks": clk,
            "e2e": {"value": e2e, "unit": "MLUPS", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "check": check,
            "secondary": secondary,
        }
        print(json.dumps(line), flush=True)

    if dist is not None:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------
# secondary workload (BASELINE config 4): liquid-crystal blue phase / cholesteric, 128^3 per GPU
# ------------------------------------------------------------------------------------------------------

LC = dict(a0=0.01, q0=0.19635, gamma=3.0, kappa0=0.000648456, kappa1=0.000648456, xi=0.7, Gamma=0.5)   # pmpi08-chol-s01.inp
LC_B_ALG = {"stress": 40.0 + 72.0, "force_be": 72.0 + 40.0 + 24.0 + 24.0 + 40.0, "collide": 360.0}


def run_lc(args, rank, local_rank, world):
    """python bench.py --lc [--size 128]: Landau-de Gennes Q tensor + Beris-Edwards coupled to D3Q19,
    cholesteric twist initial state (tests/regression/d3q19/pmpi08-chol-s01.inp parameters, advection order 3)."""
    import numpy as np
    import torch
    import ludwig_b200 as lb
    from ludwig_b200.initial import lc_twist_q

    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    n = args.size
    nlocal = (n, n, n)
    sim = lb.Lb200(nlocal, nhalo=2, have_q=True, math=lb.MATH_STRICT if args.strict else lb.MATH_FAST, device=local_rank,
                   cart_size=world, cart_rank=rank)
    if world > 1:
        ids = [sim.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        sim.nccl_init(ids[0], world, rank)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        sim.sync()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    ns = sim.nsites
    h_f = torch.empty((19, ns), dtype=torch.float64, pin_memory=True)
    h_q = torch.empty((5, ns), dtype=torch.float64, pin_memory=True)
    h_u = torch.empty((3, ns), dtype=torch.float64, pin_memory=True)
    wv = np.array([12.0] + [2.0 if sum(abs(c) for c in cv) == 1 else 1.0 for cv in CV19[1:]]) / 36.0
    for p in range(19):
        h_f.numpy()[p, :] = wv[p]
    rng = np.random.default_rng(8361235 + rank)
    h_q.numpy()[...] = lc_twist_q(nlocal, 2, LC["q0"], 1.0 / 3.0, 2)
    sim.interior(h_q.numpy())[...] += 0.01 * (rng.random((5,) + nlocal) - 0.5)
    cp = lb.CollideParam.make(lb.RELAX_M10, 1.0, 0.1)
    lc = lb.LcParam.make(adv_order=3, **LC)
    stream = torch.cuda.ExternalStream(sim.stream(), device=torch.device("cuda", local_rank))
    H2D, D2H = 1, 2

    def upload():
        sim.memcpy_async(lb.F, h_f.data_ptr(), H2D)
        sim.memcpy_async(lb.Q, h_q.data_ptr(), H2D)

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    upload()
    sim.sync()
    sim.step_lc(cp, lc, args.warmup)
    barrier()
    clocks.mark_begin()
    l0 = sim.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    sim.step_lc(cp, lc, args.steps)
    e1.record(stream)
    barrier()
    clocks.mark_end()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = sim.launch_count() - l0
    clk = clocks.stop() if rank == 0 else None
    sites = float(n) ** 3
    mlups = sites * world * args.steps / (ms * 1e-3) / 1e6

    sim.profile(True)
    sim.step_lc(cp, lc, min(args.steps, 20))
    sim.sync()
    prof = sim.profile_get()
    sim.profile(False)
    kernels = {k: {"ms_per_launch": (t / c if c else None), "launches": c} for k, (t, c) in prof.items() if c}
    peak, peak_src = measured_peaks()

    def gbs(key, alg):
        t = kernels.get(key, {}).get("ms_per_launch")
        return alg * sites / (t * 1e-3) / 1e9 if t else None

    barrier()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    t0.record(stream)
    upload()
    sim.step_lc(cp, lc, args.steps)
    sim.memcpy_async(lb.Q, h_q.data_ptr(), D2H)
    sim.memcpy_async(lb.U, h_u.data_ptr(), D2H)
    t1.record(stream)
    sim.sync()
    wall = (time.perf_counter() - w0) * 1e3
    barrier()
    e2e_ms = max_over_ranks(max(t0.elapsed_time(t1), wall))
    q_sum = float(np.nansum(sim.interior(h_q.numpy())[0]))

    cpu = None
    if rank == 0 and not args.no_cpu:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import refharness
            nsamp = 64
            if refharness.available(fast=True):
                lc_threads = refharness.omp_threads(host_threads(), fast=True) or host_threads()
                ref = refharness.RefSim((nsamp,) * 3, nhalo=2, adv_order=3, eta_shear=0.1, fast=True, lc=LC)
                ref.init_rest(1.0); ref.lc_twist_init(2, 1.0 / 3.0)
                ref.step(1)
                t = ref.time_steps(3) / 3
                ref.close()
                cpu = {"value": nsamp ** 3 / t / 1e6, "unit": "MLUPS", "cores": lc_threads, "kind": "reference",
                       "sample": f"3 full liquid-crystal time steps of a {nsamp}^3 lattice after 1 warm-up step"}
        except Exception as exc:
            cpu = {"value": None, "unit": "MLUPS", "cores": 0, "kind": "reference", "sample": f"failed: {exc}"}

    b_step = LC_B_ALG["stress"] + LC_B_ALG["force_be"] + LC_B_ALG["collide"]
    if rank != 0:
        sim.close()
        dist.destroy_process_group()
        return
    line = {
        "metric": "MLUPS", "value": mlups, "unit": "MLUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"D3Q19 + liquid crystal (Landau-de Gennes Q tensor, Beris-Edwards, 7pt gradient, advection order 3), "
                               f"{n}^3 per GPU, cholesteric twist + noise (pmpi08-chol-s01.inp parameters); "
                               + ("halo-free steps" if world == 1 else f"{world}_1_1 x-slabs, halo kernels + NCCL x-planes of q, u, f"),
                   "lattice_per_gpu": list(nlocal), "math": "strict" if args.strict else "fast(fma)",
                   "l2": f"lattice state per step {b_step * sites / 1e9:.2f} GB vs 126 MB L2"},
        "roofline": {"bound": "hbm", "kernel": "collide_d3q19 (pull-stream + MRT collision)",
                     "achieved": gbs("collide", LC_B_ALG["collide"]), "peak": peak, "unit": "GB/s",
                     "frac": (gbs("collide", LC_B_ALG["collide"]) or 0.0) / peak, "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes_per_site": LC_B_ALG["collide"],
                     "lc_stress": {"algorithmic_bytes_per_site": LC_B_ALG["stress"], "achieved": gbs("lc_stress", LC_B_ALG["stress"])},
                     "lc_force_be": {"algorithmic_bytes_per_site": LC_B_ALG["force_be"], "achieved": gbs("lc_be", LC_B_ALG["force_be"])},
                     "whole_step": {"algorithmic_bytes_per_site": b_step, "achieved": mlups / world * 1e6 * b_step / 1e9,
                                    "frac": mlups / world * 1e6 * b_step / 1e9 / peak}},
        "kernels": kernels, "cpu_baseline": cpu, "clocks": clk,
        "e2e": {"value": sites * world * args.steps / (e2e_ms * 1e-3) / 1e6, "unit": "MLUPS",
                "h2d_bytes_per_step": (19 + 5) * ns * 8 / args.steps, "d2h_bytes_per_step": (5 + 3) * ns * 8 / args.steps},
        "gpu_launches": launches, "check": {"qxx_sum": q_sum},
    }
    print(json.dumps(line), flush=True)
    sim.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=256, help="lattice edge per GPU")
    ap.add_argument("--cpu-size", type=int, default=128, help="edge of the CPU-baseline sample lattice")
    ap.add_argument("--strict", action="store_true", help="bit-exact arithmetic mode (no FMA contraction)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-check", action="store_true", help="skip the decomposed-vs-single-GPU parity run at N > 1")
    ap.add_argument("--f32", action="store_true", help="FP32 storage of the distributions inside lb200_step (LB200_KNOB_F32); "
                    "a separate mode with a stated error bound, not the FP64 headline")
    ap.add_argument("--pipe", type=int, default=int(os.environ.get("LB200_PIPE", "0")),
                    help="slab pipeline of the single-GPU step: x-slabs (0: off; LB200_KNOB_PIPE)")
    ap.add_argument("--pipe-sms", type=int, default=int(os.environ.get("LB200_PIPE_SMS", "56")),
                    help="SMs of the phi-sector partition (LB200_KNOB_PIPE_SMS)")
    ap.add_argument("--lc", action="store_true", help="secondary workload: liquid crystal (BASELINE config 4), --size 128 unless given")
    ap.add_argument("--le", type=int, default=0, help="Lees-Edwards planes per GPU (0: none, the headline workload)")
    ap.add_argument("--no-secondary", dest="secondary", action="store_false",
                    help="skip the `secondary` block (strict mode, single fluid, liquid crystal, strong scaling, Lees-Edwards)")
    ap.add_argument("--strong", action="store_true",
                    help="strong scaling: a fixed (2*size) x size x size lattice over all GPUs (default: weak, size^3 per GPU)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if args.lc and not (world == 1 and args.gpus > 1):
        if "--size" not in sys.argv:
            args.size = 128
        run_lc(args, rank, local_rank, world)
        return
    if world == 1 and args.gpus > 1:
        # not under torchrun: launch ourselves one process per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
