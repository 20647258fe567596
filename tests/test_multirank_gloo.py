"""Multi-rank host logic on CPU (gloo, world_size 2 and 3): the x-slab exchange plan exported by the library
(lb200_slab_plan: neighbour ranks, plane offsets, staging layout -- the same function the CUDA exchange uses)
drives a CPU emulation of the decomposed time step: every rank advances its slab with the oracle's local
operators, the boundary planes travel through torch.distributed (gloo) exactly as they travel through NCCL
on the GPUs, and the gathered result must equal the undecomposed oracle run bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, ".."))

from common import BINARY, ETA, seeded_state  # noqa: E402


def exchange(a, ncomp, depth, nlocal, nhalo, periodic, world, rank):
    """x-halo planes of canonical array a (ncomp, nsites) <- neighbours' boundary planes, via the plan."""
    import ludwig_b200 as lb
    p = lb.slab_plan(nlocal, nhalo, periodic, world, rank, ncomp, depth)
    nall = tuple(n + 2 * nhalo for n in nlocal)
    flat = a.reshape(ncomp, -1)
    assert flat.shape[1] == p.nsites

    def planes(off):
        return np.ascontiguousarray(flat[:, off:off + p.chunk])

    send_hi, send_lo = torch.from_numpy(planes(p.off_hi)), torch.from_numpy(planes(p.off_lo))
    xlo, xhi = torch.empty_like(send_hi), torch.empty_like(send_lo)
    reqs = []
    if p.has_hi:
        reqs.append(dist.isend(send_hi, p.right, tag=1))
        reqs.append(dist.irecv(xhi, p.right, tag=2))
    if p.has_lo:
        reqs.append(dist.isend(send_lo, p.left, tag=2))
        reqs.append(dist.irecv(xlo, p.left, tag=1))
    for r in reqs:
        r.wait()
    # staging layout [comp][depth][y][z]; the halo-shell kernel reads interior (j,k) of the staged planes and
    # wraps y/z itself: emulate by wrapping the rims of each staged plane from its own interior
    h = nhalo
    for buf, off, present in ((xlo, p.halo_lo, p.has_lo), (xhi, p.halo_hi, p.has_hi)):
        if not present:
            flat[:, off:off + p.chunk] = 0.0          # absent neighbour: zeros arrive (reference quirk)
            continue
        st = buf.numpy().reshape(ncomp, depth, nall[1], nall[2])
        core = st[:, :, h:-h, h:-h]
        wrapped = np.pad(core, ((0, 0), (0, 0), (h, h), (h, h)), mode="wrap") if periodic[1] and periodic[2] else None
        assert wrapped is not None
        flat[:, off:off + p.chunk] = wrapped.reshape(ncomp, -1)


def worker(rank, world, port, periodic, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import Oracle
    nxl, ny, nz, nhalo, nsteps = 4, 5, 6, 2, 4
    nglobal = (nxl * world, ny, nz)
    og = Oracle(nglobal, nhalo=nhalo, periodic=periodic)
    st = seeded_state(og, seed=33)
    # local slab: x is never locally periodic (its images live on the neighbours)
    ol = Oracle((nxl, ny, nz), nhalo=nhalo, periodic=(0, periodic[1], periodic[2]))

    def slab(a):
        v = a.reshape((-1,) + og.nall)
        out = np.zeros((v.shape[0],) + ol.nall)
        out[:, nhalo:nhalo + nxl] = v[:, nhalo + rank * nxl:nhalo + (rank + 1) * nxl]
        return out.reshape(v.shape[0], -1)

    f, phi = slab(st["f"]), slab(st["phi"])
    z = lambda k: np.zeros((k, ol.nsites))
    u, rho, force, grad, delsq, strs, flux = z(3), z(1), z(3), z(3), z(1), z(9), z(4)
    fp = f.copy()
    cp = ol.collide_param(0, 1.0, ETA, force=(1e-6, -2e-6, 5e-7))
    sp = ol.symm_param(adv_order=3, **BINARY)
    ex = lambda a, d: exchange(a, a.shape[0], d, (nxl, ny, nz), nhalo, periodic, world, rank)
    for _ in range(nsteps):
        force[...] = 0.0
        ol.field_halo(phi); ex(phi, nhalo)
        ol.grad_27pt(phi, grad, delsq)
        ol.stress_symm(sp, phi, grad, delsq, strs)
        ol.force_divergence(strs, force)
        ol.field_halo(u); ex(u, nhalo)
        ol.advection(3, u, phi, flux); ol.flux_mu(sp, phi, delsq, flux); ol.flux_mu_ext(sp, flux)
        ol.phi_update(flux, phi)
        u[...] = 0.0
        ol.collide(cp, f, force, rho, u)
        ol.lb_halo(f); ex(f, 1)
        ol.propagation(f, fp)
        f, fp = fp, f
    mine = {k: np.ascontiguousarray(ol.interior(a)) for k, a in (("f", f), ("phi", phi), ("u", u), ("rho", rho))}
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        og.step(og.collide_param(0, 1.0, ETA, force=(1e-6, -2e-6, 5e-7)), og.symm_param(adv_order=3, **BINARY), 1, nsteps,
                st["f"], st["phi"], st["u"], st["rho"], st["force"], st["grad"], st["delsq"])
        ok = all(np.array_equal(np.concatenate([g[k] for g in gathered], axis=1), og.interior(st[k])) for k in mine)
        ret.put(ok)
    dist.destroy_process_group()


def exchange_le(a, ncomp, depth, ol, world, rank):
    """as exchange(), for arrays that carry Lees-Edwards buffer planes after the high x halo (the plan's component
    stride counts them; the exchanged planes are the same)"""
    import ludwig_b200 as lb
    o = lb.Options()
    o.nlocal[:] = ol.nlocal
    o.nhalo, o.nvel, o.ndist, o.halo_scheme = ol.nhalo, 19, 1, lb.HALO_FULL
    o.periodic[:] = (1, 1, 1)
    o.cart_size, o.cart_rank = world, rank
    o.le_nplanes = ol.le_nplanes * world
    p = lb.SlabPlan()
    assert lb.load_library().lb200_slab_plan(__import__("ctypes").byref(o), ncomp, depth, __import__("ctypes").byref(p)) == 0
    flat = a.reshape(ncomp, -1)
    assert flat.shape[1] == p.nsites, (flat.shape, p.nsites)
    planes = lambda off: np.ascontiguousarray(flat[:, off:off + p.chunk])
    send_hi, send_lo = torch.from_numpy(planes(p.off_hi)), torch.from_numpy(planes(p.off_lo))
    xlo, xhi = torch.empty_like(send_hi), torch.empty_like(send_lo)
    reqs = [dist.isend(send_hi, p.right, tag=1), dist.irecv(xhi, p.right, tag=2),
            dist.isend(send_lo, p.left, tag=2), dist.irecv(xlo, p.left, tag=1)]
    for r in reqs:
        r.wait()
    h = ol.nhalo
    for buf, off in ((xlo, p.halo_lo), (xhi, p.halo_hi)):
        st = buf.numpy().reshape(ncomp, depth, ol.nall[1], ol.nall[2])
        flat[:, off:off + p.chunk] = np.pad(st[:, :, h:-h, h:-h], ((0, 0), (0, 0), (h, h), (h, h)), mode="wrap").reshape(ncomp, -1)


def worker_le(rank, world, port, ret):
    """Lees-Edwards sheared binary fluid on x-slabs, one plane per rank: every plane operation is local to a rank (planes
    are interior to the slabs, the displacement is along y), only the ordinary x-planes travel"""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import Oracle
    nxl, ny, nz, nhalo, nsteps, uy, eta = 8, 6, 5, 2, 5, 0.05, 0.1
    fe = dict(a=-0.0625, b=0.0625, kappa=0.04, mobility=0.15)
    og = Oracle((nxl * world, ny, nz), nhalo=nhalo, le_nplanes=world, le_uy=uy)
    ol = Oracle((nxl, ny, nz), nhalo=nhalo, periodic=(0, 1, 1), le_nplanes=1, le_uy=uy)
    rng = np.random.default_rng(8)
    fg = np.zeros((19, og.nsites_lb)); og.le_init_shear_profile(1.0, eta, fg)
    phig = np.zeros((1, og.nsites)); og.interior(phig)[...] = 0.05 * (rng.random(og.nlocal) - 0.5)

    def slab(a, field):
        v = a.reshape((a.shape[0], -1) + og.nall[1:])
        out = np.zeros((v.shape[0], ol.nall[0] + (ol.nxbuffer if field else 0)) + ol.nall[1:])
        out[:, nhalo:nhalo + nxl] = v[:, nhalo + rank * nxl:nhalo + (rank + 1) * nxl]
        return out.reshape(v.shape[0], -1)

    f, phi = slab(fg, False), slab(phig, True)
    z = lambda k: np.zeros((k, ol.nsites))
    u, rho, force, grad, delsq, flux = z(3), z(1), z(3), z(3), z(1), z(4)
    fp = f.copy()
    cp, sp = ol.collide_param(0, 1.0, eta), ol.symm_param(adv_order=3, **fe)
    exf = lambda a, d: exchange(a, a.shape[0], d, (nxl, ny, nz), nhalo, (1, 1, 1), world, rank)     # distributions: no buffers
    exl = lambda a, d: exchange_le(a, a.shape[0], d, ol, world, rank)
    for n in range(nsteps):
        tstep = float(n + 1); time = tstep - 1.0
        force[...] = 0.0
        ol.field_halo(phi); exl(phi, nhalo)
        ol.le_field(time, phi); ol.grad_27pt(phi, grad, delsq); ol.le_grad_buffer(phi, grad, delsq)
        ol.le_phi_force(sp, phi, grad, delsq, force)
        ol.field_halo(u); exl(u, nhalo)
        ol.le_hydro(time, u)
        ol.advection(3, u, phi, flux); ol.flux_mu(sp, phi, delsq, flux); ol.flux_mu_ext(sp, flux)
        ol.le_fix_fluxes(time, flux); ol.phi_update(flux, phi)
        u[...] = 0.0
        ol.collide(cp, f, force, rho, u)
        ol.le_lb_bc(tstep, f)
        ol.lb_halo(f); exf(f, 1)
        ol.propagation(f, fp)
        f, fp = fp, f
    mine = {k: np.ascontiguousarray(ol.interior(a)) for k, a in (("f", f), ("phi", phi), ("u", u))}
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        zg = lambda k: np.zeros((k, og.nsites))
        ug = zg(3)
        og.le_step(og.collide_param(0, 1.0, eta), og.symm_param(adv_order=3, **fe), 0, nsteps, fg, phig, ug, zg(1), zg(3), zg(3), zg(1))
        want = dict(f=fg, phi=phig, u=ug)
        ret.put(all(np.array_equal(np.concatenate([g[k] for g in gathered], axis=1), og.interior(want[k])) for k in mine))
    dist.destroy_process_group()


def worker_lc(rank, world, port, ret):
    """liquid crystal on x-slabs: the x-planes of q (5 components, depth nhalo), u and f travel; everything else is local"""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import Oracle
    from ludwig_b200.initial import equilibrium_f, lc_twist_q
    nxl, ny, nz, nhalo, nsteps, eta = 4, 5, 6, 2, 5, 0.1
    lc = dict(a0=0.01, q0=0.19635, gamma=3.0, kappa0=0.000648456, kappa1=0.0008, xi=0.7, Gamma=0.5)
    n = (nxl * world, ny, nz)
    og = Oracle(n, nhalo=nhalo)
    ol = Oracle((nxl, ny, nz), nhalo=nhalo, periodic=(0, 1, 1))
    rng = np.random.default_rng(9)
    qg = lc_twist_q(n, nhalo, lc["q0"], 1.0 / 3.0, 0)
    og.interior(qg)[...] += 0.02 * (rng.random(og.interior(qg).shape) - 0.5)
    fg = equilibrium_f(n, nhalo)

    def slab(a):
        v = a.reshape((a.shape[0],) + og.nall)
        out = np.zeros((v.shape[0],) + ol.nall)
        out[:, nhalo:nhalo + nxl] = v[:, nhalo + rank * nxl:nhalo + (rank + 1) * nxl]
        return out.reshape(v.shape[0], -1)

    f, q = slab(fg), slab(qg)
    z = lambda k: np.zeros((k, ol.nsites))
    u, rho, force, qgrad, qdelsq, strs, h, flux = z(3), z(1), z(3), z(15), z(5), z(9), z(5), z(20)
    fp = f.copy()
    cp, p = ol.collide_param(0, 1.0, eta), ol.lc_param(**lc)
    ex = lambda a, d: exchange(a, a.shape[0], d, (nxl, ny, nz), nhalo, (1, 1, 1), world, rank)
    for _ in range(nsteps):
        force[...] = 0.0
        ol.field_halo(q); ex(q, nhalo)
        ol.grad_7pt(q, qgrad, qdelsq)
        ol.lc_stress(p, q, qgrad, qdelsq, strs)
        ol.force_divergence(strs, force)
        ol.field_halo(u); ex(u, nhalo)
        ol.advection_nf(3, u, q, flux)
        ol.lc_mol_field(p, q, qgrad, qdelsq, h)
        ol.beris_edw_update(p, u, h, flux, q)
        u[...] = 0.0
        ol.collide(cp, f, force, rho, u)
        ol.lb_halo(f); ex(f, 1)
        ol.propagation(f, fp)
        f, fp = fp, f
    mine = {k: np.ascontiguousarray(ol.interior(a)) for k, a in (("f", f), ("q", q), ("u", u))}
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        zg = lambda k: np.zeros((k, og.nsites))
        ug = zg(3)
        og.lc_step(og.collide_param(0, 1.0, eta), og.lc_param(**lc), 3, nsteps, fg, qg, ug, zg(1), zg(3), zg(15), zg(5))
        want = dict(f=fg, q=qg, u=ug)
        ret.put(all(np.array_equal(np.concatenate([g[k] for g in gathered], axis=1), og.interior(want[k])) for k in mine))
    dist.destroy_process_group()


@pytest.mark.parametrize("target,world", [(worker_le, 2), (worker_lc, 2), (worker_lc, 3)])
def test_le_and_lc_decomposed_steps(target, world):
    """Lees-Edwards and liquid-crystal time steps on x-slabs (gloo, CPU): local oracle operators + the library's slab
    plan == the undecomposed oracle, bit for bit"""
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29700 + world * 11 + (3 if target is worker_le else 0)
    procs = [ctx.Process(target=target, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert ret.get(timeout=10) is True


@pytest.mark.parametrize("world,periodic", [(2, (1, 1, 1)), (3, (1, 1, 1)), (2, (0, 1, 1))])
def test_slab_plan_drives_a_bit_exact_decomposed_step(world, periodic):
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29600 + world * 7 + periodic[0]
    procs = [ctx.Process(target=worker, args=(r, world, port, periodic, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert ret.get(timeout=10) is True


def test_slab_plan_values():
    import ludwig_b200 as lb
    p = lb.slab_plan((256, 256, 256), 2, (1, 1, 1), 8, 0, 19, 1)
    assert (p.left, p.right, p.has_lo, p.has_hi) == (7, 1, 1, 1)
    assert p.chunk == 260 * 260 and p.count == 19 * 260 * 260 and p.nsites == 260 ** 3
    assert p.off_lo == 2 * 260 * 260 and p.off_hi == 257 * 260 * 260
    assert p.halo_lo == 1 * 260 * 260 and p.halo_hi == 258 * 260 * 260
    q = lb.slab_plan((256, 256, 256), 2, (0, 1, 1), 8, 7, 3, 2)
    assert (q.has_lo, q.has_hi) == (1, 0) and q.chunk == 2 * 260 * 260 and q.off_hi == 256 * 260 * 260
    with pytest.raises(lb.Lb200Error):
        lb.slab_plan((8, 8, 8), 1, (1, 1, 1), 2, 0, 1, 2)      # depth > nhalo


# ---- the halo-free time step of lb200_step on x-slabs: only what lb200_step_plan lists crosses the slab boundary ----

def step_exchange(a, what, nlocal, nhalo, world, rank, nvel=19):
    """Boundary planes of canonical array a -> the neighbours' halo planes, exactly the components and planes of
    lb200_step_plan (the peer-store kernels and the NCCL path of lb200_step move precisely these)."""
    import ludwig_b200 as lb
    p = lb.step_plan(nlocal, nhalo, world, rank, what, nvel=nvel)
    flat = a.reshape(a.shape[0], -1)
    assert flat.shape[1] == p.nsites
    up, down = list(p.comp_up[:p.ncomp_up]), list(p.comp_down[:p.ncomp_down])
    assert p.dst_down - p.src_down == p.peer_shift and p.src_up - p.dst_up == p.peer_shift
    send_up = torch.from_numpy(np.ascontiguousarray(flat[up, p.src_up:p.src_up + p.chunk]))
    send_dn = torch.from_numpy(np.ascontiguousarray(flat[down, p.src_down:p.src_down + p.chunk]))
    recv_lo, recv_hi = torch.empty_like(send_up), torch.empty_like(send_dn)
    reqs = [dist.isend(send_up, p.right, tag=11), dist.isend(send_dn, p.left, tag=12),
            dist.irecv(recv_lo, p.left, tag=11), dist.irecv(recv_hi, p.right, tag=12)]
    for r in reqs:
        r.wait()
    nall = tuple(n + 2 * nhalo for n in nlocal)
    h = nhalo
    for buf, comps, off in ((recv_lo, up, p.dst_up), (recv_hi, down, p.dst_down)):
        st = buf.numpy().reshape(len(comps), p.depth, nall[1], nall[2])
        # the kernels read the y/z images of the received planes from their interior: emulate by wrapping the rims
        wrapped = np.pad(st[:, :, h:-h, h:-h], ((0, 0), (0, 0), (h, h), (h, h)), mode="wrap")
        flat[comps, off:off + p.chunk] = wrapped.reshape(len(comps), -1)


def worker_halo_free(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ludwig_b200 as lb
    from oracle import Oracle
    nxl, ny, nz, nhalo, nsteps = 4, 5, 6, 2, 4
    nglobal = (nxl * world, ny, nz)
    og = Oracle(nglobal, nhalo=nhalo)
    st = seeded_state(og, seed=34)
    ol = Oracle((nxl, ny, nz), nhalo=nhalo, periodic=(0, 1, 1))

    def slab(a):
        v = a.reshape((-1,) + og.nall)
        out = np.zeros((v.shape[0],) + ol.nall)
        out[:, nhalo:nhalo + nxl] = v[:, nhalo + rank * nxl:nhalo + (rank + 1) * nxl]
        return out.reshape(v.shape[0], -1)

    f, phi = slab(st["f"]), slab(st["phi"])
    z = lambda k: np.zeros((k, ol.nsites))
    u, rho, force, grad, delsq, strs, flux = z(3), z(1), z(3), z(3), z(1), z(9), z(4)
    fp = f.copy()
    fg = (1e-6, -2e-6, 5e-7)
    cp = ol.collide_param(0, 1.0, ETA, force=fg)
    sp = ol.symm_param(adv_order=3, **BINARY)
    n3 = (nxl, ny, nz)
    for _ in range(nsteps):
        force[...] = 0.0
        ol.field_halo(phi)                                   # y/z images (in-kernel wrap on the GPU); x: nothing
        step_exchange(phi, lb.STEP_PHI, n3, nhalo, world, rank)
        ol.grad_27pt(phi, grad, delsq)
        ol.stress_symm(sp, phi, grad, delsq, strs)
        ol.force_divergence(strs, force)
        ol.field_halo(u)
        step_exchange(u, lb.STEP_UX, n3, nhalo, world, rank)  # u_x only, one plane
        ol.advection(3, u, phi, flux); ol.flux_mu(sp, phi, delsq, flux); ol.flux_mu_ext(sp, flux)
        ol.phi_update(flux, phi)
        u[...] = 0.0
        ol.collide(cp, f, force, rho, u)
        ol.lb_halo(f)
        step_exchange(f, lb.STEP_F, n3, nhalo, world, rank)   # 5 of 19 populations per direction
        ol.propagation(f, fp)
        f, fp = fp, f
    mine = {k: np.ascontiguousarray(ol.interior(a)) for k, a in (("f", f), ("phi", phi), ("u", u), ("rho", rho))}
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        og.step(og.collide_param(0, 1.0, ETA, force=fg), og.symm_param(adv_order=3, **BINARY), 1, nsteps,
                st["f"], st["phi"], st["u"], st["rho"], st["force"], st["grad"], st["delsq"])
        ok = all(np.array_equal(np.concatenate([g[k] for g in gathered], axis=1), og.interior(st[k])) for k in mine)
        ret.put(ok)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_free_step_plan_is_sufficient_and_bit_exact(world):
    """Only 2 planes of phi, 1 plane of u_x and the populations with c_x = +-1 cross the slab boundary per step and
    direction (lb200_step_plan); the decomposed run still equals the undecomposed oracle bit for bit."""
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29700 + world * 7
    procs = [ctx.Process(target=worker_halo_free, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert ret.get(timeout=10) is True


def test_step_plan_values():
    import ludwig_b200 as lb
    p = lb.step_plan((256, 256, 256), 2, 8, 3, lb.STEP_F)
    assert (p.left, p.right, p.depth) == (2, 4, 1)
    assert list(p.comp_up[:p.ncomp_up]) == [1, 2, 3, 4, 5] and list(p.comp_down[:p.ncomp_down]) == [14, 15, 16, 17, 18]
    xs = 260 * 260
    assert (p.src_up, p.dst_up, p.src_down, p.dst_down, p.peer_shift) == (257 * xs, 1 * xs, 2 * xs, 258 * xs, 256 * xs)
    q = lb.step_plan((64, 256, 256), 2, 8, 0, lb.STEP_PHI)
    assert (q.left, q.right, q.depth, q.chunk) == (7, 1, 2, 2 * xs) and (q.src_up, q.dst_up) == (64 * xs, 0)
    r = lb.step_plan((8, 8, 8), 1, 2, 1, lb.STEP_F, nvel=27)
    assert r.ncomp_up == 9 and r.ncomp_down == 9
