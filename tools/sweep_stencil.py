"""GPU sweep of the stencil-kernel launch geometry (tile shape x planes per chunk) on the CG loop.
Usage (on the GPU box): python tools/sweep_stencil.py [grid] [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from algotoolkit_b200 import capi, synthetic

g = int(sys.argv[1]) if len(sys.argv) > 1 else 512
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 30
tiles = [int(t) for t in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 1, 2, 3]
kcs = [int(t) for t in sys.argv[4].split(",")] if len(sys.argv) > 4 else [4, 8, 16, 32, 64]
os.environ["ATK_CG_FUSED"] = "0"  # the three-kernel CG: K1 is the plain apply (+ p.Ap)
ctx = capi.Context(0)
cx, cy, cz = synthetic.poisson_coeffs(g, g, g)
n = g ** 3
b = synthetic.poisson_rhs_r(g, g, g)
db = ctx.vector(b)
dx, dr, dp = ctx.empty(n), ctx.empty(n), ctx.empty(n)
print(f"grid {g}^3 rows {n}  iters {iters}")
for tile in tiles:
    for kc in kcs:
        os.environ["ATK_STENCIL_TILE"] = str(tile)
        os.environ["ATK_STENCIL_KC"] = str(kc)
        op = ctx.operator_stencil7(g, g, g, cx, cy, cz)
        for timed in (False, True):
            dx.zero(); dr.copy_from(db); dp.copy_from(db)
            res = ctx.cg_solve(op, db, dx, dr, dp, iters, time_kernels=timed)
            if not timed:
                ms_it = res["device_ms"] / iters
        k = [v / iters for v in res["kernel_ms"]]
        gbs = [16 * n / (k[0] * 1e6), 48 * n / (k[1] * 1e6), 24 * n / (k[2] * 1e6)]
        print(f"tile {tile} kc {kc:3d}: graph {ms_it:.3f} ms/it ({1000 / ms_it:.1f} it/s, {88 * n / ms_it / 1e6:.0f} GB/s alg) | "
              f"K1 {k[0]:.3f} ms {gbs[0]:.0f} GB/s  K2 {k[1]:.3f} ms {gbs[1]:.0f} GB/s  K3 {k[2]:.3f} ms {gbs[2]:.0f} GB/s", flush=True)
        op.destroy()
