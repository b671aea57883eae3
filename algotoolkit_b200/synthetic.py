"""Deterministic synthetic inputs of SURVEY.md 8(d) (numpy, host side; these are INPUT generators, not solvers).

u01(s) is the splitmix64 finaliser mapped to [0,1).  RHS-R is the seeded random right-hand side used for every
Poisson throughput / parity run: f = 2*u01(seed + idx) - 1 inside, 0 on the Dirichlet shell (g = 0), which is what
Poisson3DSolver::setRHS followed by setDirichletBC leaves in b (reference src/PDEs/FDM/Poisson.hpp:103-151).
"""
from __future__ import annotations

import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def u01_array(seed: int, start: int, count: int, out: np.ndarray | None = None) -> np.ndarray:
    """u01(seed + i) for i in [start, start+count)."""
    with np.errstate(over="ignore"):
        s = np.arange(start, start + count, dtype=np.uint64)
        s += np.uint64(seed & 0xFFFFFFFFFFFFFFFF)
        s += np.uint64(0x9E3779B97F4A7C15)
        s = (s ^ (s >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        s = (s ^ (s >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        s ^= s >> np.uint64(31)
        s >>= np.uint64(11)
    r = s.astype(np.float64)
    r *= 1.0 / 9007199254740992.0
    if out is not None:
        out[...] = r
        return out
    return r


def poisson_coeffs(nx, ny, nz, lx=1.0, ly=1.0, lz=1.0):
    """cx, cy, cz exactly as Poisson.hpp:36,53-56 computes them (h = l/(n-1), c = 1/(h*h))."""
    hx, hy, hz = lx / (nx - 1), ly / (ny - 1), lz / (nz - 1)
    return 1.0 / (hx * hx), 1.0 / (hy * hy), 1.0 / (hz * hz)


def poisson_rhs_r(nx, ny, nz, k_begin=0, k_end=None, seed=12345, out: np.ndarray | None = None) -> np.ndarray:
    """RHS-R for the z-slab [k_begin, k_end) of an nx*ny*nz grid (whole grid by default), plane by plane."""
    k_end = nz if k_end is None else k_end
    plane = nx * ny
    n = plane * (k_end - k_begin)
    b = out if out is not None else np.empty(n, np.float64)
    assert b.size == n
    for k in range(k_begin, k_end):
        v = b[(k - k_begin) * plane:(k - k_begin + 1) * plane]
        if k == 0 or k == nz - 1:
            v[:] = 0.0
            continue
        u01_array(seed, k * plane, plane, out=v)
        v *= 2.0
        v -= 1.0
        g = v.reshape(ny, nx)
        g[0, :] = 0.0
        g[-1, :] = 0.0
        g[:, 0] = 0.0
        g[:, -1] = 0.0
    return b


def stencil27_coefficients():
    """BASELINE config 5 operator: 40 on the diagonal, -1 + 0.3*(di + 0.5*dj + 0.25*dk) off it."""
    return np.array([40.0 if (di == 0 and dj == 0 and dk == 0) else -1.0 + 0.3 * (di + 0.5 * dj + 0.25 * dk)
                     for dk in (-1, 0, 1) for dj in (-1, 0, 1) for di in (-1, 0, 1)])
