"""Host-side initial conditions in the canonical layout, restating the reference's own initialisers so
that a run started here is bit-identical to one started by Ludwig:

* spinodal_phi: `phi_initialisation spinodal` -- phi = phi0 + amp (r - 1/2) with r from the per-site
  lattice RNG (reference src/field_phi_init.c:484-518, src/noise.c:278-345, 437-453, src/noise.h:96-110).
* equilibrium_f: lb_1st_moment_equilib_set (reference src/lb_data.c:809-834), e.g. the rest state
  written by lb_init_rest_f (:659-680).
"""
import numpy as np

CV19 = np.array([(0, 0, 0), (1, 1, 0), (1, 0, 1), (1, 0, 0), (1, 0, -1), (1, -1, 0), (0, 1, 1), (0, 1, 0),
                 (0, 1, -1), (0, 0, 1), (0, 0, -1), (0, -1, 1), (0, -1, 0), (0, -1, -1), (-1, 1, 0),
                 (-1, 0, 1), (-1, 0, 0), (-1, 0, -1), (-1, -1, 0)], dtype=np.int64)
WV19 = np.array([12.0 / 36.0] + [(2.0 / 36.0 if abs(c).sum() == 1 else 1.0 / 36.0) for c in CV19[1:]])


def _ns_uniform(s):
    """One draw of the reference's lattice RNG on uint32 state arrays s[0..3] (modified in place)."""
    u32 = np.uint32
    with np.errstate(over="ignore"):
        s[0] = u32(69069) * s[0] + u32(1234567)
        b = s[1] ^ (s[1] << u32(17))
        b ^= (b >> u32(13))
        s[1] = b ^ (b << u32(5))
        s[2] = u32(36969) * (s[2] & u32(0xffff)) + (s[2] >> u32(16))
        s[3] = u32(18000) * (s[3] & u32(0xffff)) + (s[3] >> u32(16))
        b = (s[2] << u32(16)) + s[3]
        return s[1] + (s[0] ^ b)


def spinodal_phi(nlocal, nhalo, seed, phi0=0.0, amp=0.05, noffset=(0, 0, 0), out=None):
    """phi on the allocated lattice (1, nsites); interior set, halo zero."""
    nall = tuple(n + 2 * nhalo for n in nlocal)
    if out is None:
        out = np.zeros((1,) + nall)
    else:
        out = out.reshape((1,) + nall)
        out[...] = 0.0
    ig, jg, kg = np.meshgrid(*(np.arange(1, n + 1, dtype=np.uint32) + np.uint32(o)
                               for n, o in zip(nlocal, noffset)), indexing="ij")
    with np.errstate(over="ignore"):
        sl = [np.uint32(seed) + ig, np.uint32(12953) + jg, np.uint32(712357) + kg,
              np.full(ig.shape, 22383979, dtype=np.uint32)]
        st = [_ns_uniform(sl) for _ in range(4)]
        st = [np.array(x, dtype=np.uint32) for x in st]
        iu = _ns_uniform(st)
    ran = (1.0 / 4294967295) * iu.astype(np.float64)
    out[0, nhalo:nhalo + nlocal[0], nhalo:nhalo + nlocal[1], nhalo:nhalo + nlocal[2]] = phi0 + amp * (ran - 0.5)
    return out.reshape(1, -1)


def equilibrium_f(nlocal, nhalo, rho=1.0, u=(0.0, 0.0, 0.0), out=None):
    """D3Q19 f_p = rho w_p (1 + 3 u.c + 4.5 (cc - I/3):uu) on the interior, zero in the halo."""
    nall = tuple(n + 2 * nhalo for n in nlocal)
    if out is None:
        out = np.zeros((19,) + nall)
    else:
        out = out.reshape((19,) + nall)
        out[...] = 0.0
    cs2 = 1.0 / 3.0
    rcs2 = 1.0 / cs2
    for p in range(19):
        udotc = 0.0
        sdotq = 0.0
        for ia in range(3):
            udotc = udotc + u[ia] * float(CV19[p, ia])
            for ib in range(3):
                dab = 1.0 if ia == ib else 0.0
                sdotq = sdotq + (float(CV19[p, ia] * CV19[p, ib]) - cs2 * dab) * u[ia] * u[ib]
        out[p, nhalo:nhalo + nlocal[0], nhalo:nhalo + nlocal[1], nhalo:nhalo + nlocal[2]] = \
            rho * WV19[p] * (1.0 + rcs2 * udotc + 0.5 * rcs2 * rcs2 * sdotq)
    return out.reshape(19, -1)


def lc_uniaxial_q(nlocal, nhalo, director, amplitude):
    """Q_ab = A/2 (3 n_a n_b - d_ab) on the interior for a director field n (3, Nx, Ny, Nz) (reference
    fe_lc_q_uniaxial, src/blue_phase.c:1406-1418); compressed (XX, XY, XZ, YY, YZ), shape (5, nsites)."""
    nall = tuple(n + 2 * nhalo for n in nlocal)
    out = np.zeros((5,) + nall)
    n = np.broadcast_to(np.asarray(director, dtype=np.float64), (3,) + tuple(nlocal))
    sl = (slice(nhalo, nhalo + nlocal[0]), slice(nhalo, nhalo + nlocal[1]), slice(nhalo, nhalo + nlocal[2]))
    for c, (a, b) in enumerate(((0, 0), (0, 1), (0, 2), (1, 1), (1, 2))):
        d = 1.0 if a == b else 0.0
        out[(c,) + sl] = 0.5 * amplitude * (3.0 * n[a] * n[b] - d)
    return out.reshape(5, -1)


def lc_twist_q(nlocal, nhalo, q0, amplitude, axis=2, noffset=(0, 0, 0)):
    """`lc_q_initialisation twist` (= cholesteric along z): n = (cos q0 z, sin q0 z, 0) (reference
    blue_phase_twist_init, src/blue_phase_init.c:763-823; x: n = (0, cos q0 x, sin q0 x); y: n = (cos q0 y, 0, -sin q0 y))."""
    import math
    n = np.zeros((3,) + tuple(nlocal))
    coord = np.arange(1, nlocal[axis] + 1, dtype=np.float64) + noffset[axis]
    c = np.array([math.cos(q0 * x) for x in coord])
    s = np.array([math.sin(q0 * x) for x in coord])
    shape = [1, 1, 1]
    shape[axis] = -1
    c, s = c.reshape(shape), s.reshape(shape)
    if axis == 2:
        n[0], n[1] = c, s
    elif axis == 0:
        n[1], n[2] = c, s
    else:
        n[0], n[2] = c, -s
    return lc_uniaxial_q(nlocal, nhalo, n, amplitude)


def lc_nematic_q(nlocal, nhalo, director, amplitude):
    """`lc_q_initialisation nematic`: uniform uniaxial nematic along the (normalised) director
    (reference blue_phase_nematic_init, src/blue_phase_init.c:836-880)."""
    import math
    d = [float(x) for x in director]
    norm = math.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])
    n = np.array([x / norm for x in d]).reshape(3, 1, 1, 1)
    return lc_uniaxial_q(nlocal, nhalo, n, amplitude)
