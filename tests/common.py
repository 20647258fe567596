"""Shared helpers for the parity tests: seeded states in the canonical layout."""
import numpy as np

BINARY = dict(a=-0.00625, b=0.00625, kappa=0.004, mobility=1.25)   # serial-spin-fd1.inp parameters
ETA = 0.00625


def seeded_state(orc, seed=7, binary=True, u_amp=0.01, phi_amp=0.05):
    """Non-trivial interior state: f = equilibrium(rho(x), u(x)) + small non-equilibrium part,
    phi = noise.  Halos are zero (they must be filled by the halo operators under test)."""
    rng = np.random.default_rng(seed)
    n = orc.nlocal
    ns = orc.nsites
    x, y, z = np.meshgrid(*(np.arange(1, m + 1) / m for m in n), indexing="ij")
    rho = 1.0 + 0.01 * np.sin(2 * np.pi * x) * np.cos(2 * np.pi * y)
    ux = u_amp * np.sin(2 * np.pi * y)
    uy = u_amp * np.sin(2 * np.pi * z)
    uz = u_amp * np.sin(2 * np.pi * x)
    f = np.zeros((orc.nvel, orc.nsites_lb))      # distributions never carry the LE buffer planes
    fi = orc.interior(f)
    for p in range(orc.nvel):
        cu = orc.cv[p, 0] * ux + orc.cv[p, 1] * uy + orc.cv[p, 2] * uz
        uu = ux * ux + uy * uy + uz * uz
        fi[p] = rho * orc.wv[p] * (1.0 + 3.0 * cu + 4.5 * cu * cu - 1.5 * uu) \
            * (1.0 + 1e-3 * (rng.random(n) - 0.5))
    st = dict(f=f, u=np.zeros((3, ns)), rho=np.zeros((1, ns)), force=np.zeros((3, ns)))
    if binary:
        phi = np.zeros((1, ns))
        orc.interior(phi)[0] = phi_amp * (rng.random(n) - 0.5)
        st.update(phi=phi, grad=np.zeros((3, ns)), delsq=np.zeros((1, ns)))
    return st


def rel_err(a, b):
    d = np.abs(a - b).max()
    s = np.abs(b).max()
    return d / s if s > 0 else d


def close_fast(a, b, rtol=1e-12, atol=1e-14):
    """Tolerance of the FMA-contracted build: 1e-12 relative to the field's magnitude (north_star), with an
    absolute floor of 1e-14 lattice units (velocities are differences of O(1) populations, so their
    rounding error does not scale with |u|; the reference's own regression diff is absolute 1e-12)."""
    return np.abs(a - b).max() <= rtol * np.abs(b).max() + atol
