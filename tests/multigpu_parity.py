#!/usr/bin/env python3
"""x-slab decomposition parity, one process per GPU (launched by torchrun from test_multigpu.py):
N ranks each advance their slab of a global binary-fluid lattice with libludwig_b200 (strict mode,
NCCL x-plane exchange); rank 0 gathers the interiors and compares them BIT FOR BIT with the CPU
oracle run on the undecomposed lattice (the reference's results are decomposition independent)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, ".."))

import ludwig_b200 as lb  # noqa: E402
from common import BINARY, ETA, close_fast, seeded_state  # noqa: E402
from oracle import Oracle  # noqa: E402


LC = dict(a0=0.01, q0=0.19635, gamma=3.0, kappa0=0.000648456, kappa1=0.0008, xi=0.7, Gamma=0.5)


def main_lc(rank, world, local):
    """liquid crystal on x-slabs (strict) == the undecomposed liquid-crystal oracle, bit for bit"""
    from ludwig_b200.initial import equilibrium_f, lc_twist_q
    nxl, ny, nz, nhalo, nsteps = 6, 10, 34, 2, 6
    n = (nxl * world, ny, nz)
    orc_g = Oracle(n, nhalo=nhalo)
    orc_l = Oracle((nxl, ny, nz), nhalo=nhalo)
    rng = np.random.default_rng(5)
    q = lc_twist_q(n, nhalo, LC["q0"], 1.0 / 3.0, 0)
    orc_g.interior(q)[...] += 0.02 * (rng.random(orc_g.interior(q).shape) - 0.5)
    f = equilibrium_f(n, nhalo)

    def slab(a):
        v = a.reshape((a.shape[0],) + orc_g.nall)
        out = np.zeros((v.shape[0],) + orc_l.nall)
        out[:, nhalo:nhalo + nxl] = v[:, nhalo + rank * nxl:nhalo + (rank + 1) * nxl]
        return out.reshape(v.shape[0], -1)

    sim = lb.Lb200((nxl, ny, nz), nhalo=nhalo, have_q=True, math=lb.MATH_STRICT, device=local, cart_size=world, cart_rank=rank)
    ids = [sim.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    sim.nccl_init(ids[0], world, rank)
    sim.put(lb.F, slab(f)); sim.put(lb.Q, slab(q))
    cp = lb.CollideParam.make(lb.RELAX_M10, 1.0, 0.1)
    pg = lb.LcParam.make(adv_order=3, **LC)
    sim.step_lc(cp, pg, 2); sim.step_lc_api(cp, pg, 2); sim.step_lc(cp, pg, nsteps - 4)
    mine = {k: np.ascontiguousarray(orc_l.interior(sim.get(a))) for k, a in (("f", lb.F), ("q", lb.Q), ("u", lb.U), ("force", lb.FORCE))}
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    ok = True
    if rank == 0:
        z = lambda k: np.zeros((k, orc_g.nsites))
        u, rho, force, qgrad, qdelsq = z(3), z(1), z(3), z(15), z(5)
        orc_g.lc_step(orc_g.collide_param(0, 1.0, 0.1), orc_g.lc_param(**LC), 3, nsteps, f, q, u, rho, force, qgrad, qdelsq)
        want = dict(f=f, q=q, u=u, force=force)
        for k in mine:
            full = np.concatenate([g[k] for g in gathered], axis=1)
            same = np.array_equal(full, orc_g.interior(want[k]))
            print(f"multigpu liquid-crystal parity world={world} {k}: {'OK' if same else 'MISMATCH'}", flush=True)
            ok = ok and same
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    sim.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nsteps = 6
    nxl, ny, nz = 6, 10, 34
    periodic = (1, 1, 1) if len(sys.argv) < 2 else tuple(int(c) for c in sys.argv[1])
    reduced = 0 if len(sys.argv) < 3 else int(sys.argv[2])
    peer = 1 if len(sys.argv) < 4 else int(sys.argv[3])
    binary = 1 if len(sys.argv) < 5 else int(sys.argv[4])
    le = 0 if len(sys.argv) < 6 else int(sys.argv[5])          # Lees-Edwards planes per slab (plane speed 0.05)
    fast = 0 if len(sys.argv) < 7 else int(sys.argv[6])        # 1: LB200_MATH_FAST, compared within tolerance
    if len(sys.argv) >= 8 and int(sys.argv[7]):
        return main_lc(rank, world, local)
    if le:
        nxl = 8*le
    nglobal = (nxl * world, ny, nz)
    nhalo = 2
    orc_g = Oracle(nglobal, nhalo=nhalo, periodic=periodic, le_nplanes=le*world, le_uy=0.05)
    st = seeded_state(orc_g, seed=21)
    fg = (1e-6, -2e-6, 5e-7)

    # my slab of the initial state (field arrays of a Lees-Edwards run carry buffer planes: zero here, rebuilt
    # from the interior every step)
    orc_l = Oracle((nxl, ny, nz), nhalo=nhalo, le_nplanes=le, le_uy=0.05)
    def slab(a):
        v = a.reshape((a.shape[0], -1) + orc_g.nall[1:])
        is_field = le and a.shape[1] == orc_g.nsites          # distributions never carry buffer planes
        nxa = orc_l.nall[0] + (orc_l.nxbuffer if is_field else 0)
        out = np.zeros((v.shape[0], nxa) + orc_l.nall[1:])
        x0 = rank * nxl
        out[:, nhalo:nhalo + nxl] = v[:, nhalo + x0:nhalo + x0 + nxl]
        return out.reshape(v.shape[0], -1)

    sim = lb.Lb200((nxl, ny, nz), nhalo=nhalo, periodic=periodic, have_phi=True, math=lb.MATH_FAST if fast else lb.MATH_STRICT, device=local,
                   halo_scheme=lb.HALO_REDUCED if reduced else lb.HALO_FULL, cart_size=world, cart_rank=rank,
                   le_nplanes=le*world, le_uy=0.05)
    ids = [sim.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    sim.nccl_init(ids[0], world, rank)
    sim.set_knob(lb.KNOB_PEER, peer)
    sim.put(lb.F, slab(st["f"])); sim.put(lb.PHI, slab(st["phi"]))
    cp = lb.CollideParam.make(lb.RELAX_M10, 1.0, ETA, force=fg)
    conserve2 = bool(os.environ.get("LB200_TEST_CONSERVE2"))
    sp = lb.SymmParam.make(adv_order=3, conserve=2 if conserve2 else 0, **BINARY)
    sum0 = 0.0
    if conserve2:
        # cahn_hilliard_options_conserve 2: the initial sum over ALL slabs (collective), offset so that the correction is visible
        sum0 = sim.phi_conserve_sum() + 1.0e-3
        assert abs(sum0 - 1.0e-3 - orc_g.phi_sum_time0(st["phi"])) <= 4*np.spacing(abs(sum0))
        sim.phi_init_sum_set(sum0)
    # every path over the decomposed lattice, and the hand-overs between them: halo-free lb200_step (only the
    # planes the kernels read cross NVLink), the individual entry points (full halo swaps), lb200_step again
    if not binary:
        sp = None
    sim.step(cp, sp, 2)
    mode = sim.exchange_mode()
    sim.step_api(cp, sp, 2)
    sim.step(cp, sp, nsteps - 4)
    fully_periodic = all(periodic)
    if rank == 0:
        print(f"exchange mode of lb200_step: {mode} (0 single, 1 NCCL, 2 peer stores); requested peer={peer}", flush=True)
    if fully_periodic and peer and os.environ.get("LB200_ALLOW_NO_PEER") is None:
        assert mode == 2, "peer-store exchange was requested but could not be set up"
    mine = {k: np.ascontiguousarray(orc_l.interior(sim.get(a))) for k, a in
            (("f", lb.F), ("phi", lb.PHI), ("u", lb.U), ("rho", lb.RHO), ("force", lb.FORCE))}
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    ok = True
    if rank == 0:
        if le:
            orc_g.le_step(orc_g.collide_param(0, 1.0, ETA, force=fg), orc_g.symm_param(adv_order=3, **BINARY), 0, nsteps,
                          st["f"], st["phi"], st["u"], st["rho"], st["force"], st["grad"], st["delsq"])
        elif binary:
            orc_g.step(orc_g.collide_param(0, 1.0, ETA, force=fg),
                       orc_g.symm_param(adv_order=3, conserve=2 if conserve2 else 0, phi_init_sum=sum0, **BINARY), 1, nsteps,
                       st["f"], st["phi"], st["u"], st["rho"], st["force"], st["grad"], st["delsq"], halo_reduced=reduced)
        else:
            orc_g.step(orc_g.collide_param(0, 1.0, ETA, force=fg), None, 0, nsteps,
                       st["f"], None, st["u"], st["rho"], st["force"], None, None, halo_reduced=reduced)
        for k in mine:
            if not binary and k in ("phi", "force"):
                continue
            full = np.concatenate([g[k] for g in gathered], axis=1)
            same = close_fast(full, orc_g.interior(st[k])) if (fast or conserve2) else np.array_equal(full, orc_g.interior(st[k]))
            print(f"multigpu parity world={world} periodic={periodic} reduced={reduced} {k}: {'OK' if same else 'MISMATCH'}", flush=True)
            ok = ok and same
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    sim.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
