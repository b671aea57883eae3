"""GPU sweep of the fused CG step (KA) launch geometry.  python tools/sweep_fused.py [grid] [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from algotoolkit_b200 import capi, synthetic

g = int(sys.argv[1]) if len(sys.argv) > 1 else 512
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 30
tiles = [int(t) for t in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 1, 2, 3, 4, 5]
kcs = [int(t) for t in sys.argv[4].split(",")] if len(sys.argv) > 4 else [8, 16, 32, 64, 128]
ctx = capi.Context(0)
cx, cy, cz = synthetic.poisson_coeffs(g, g, g)
n = g ** 3
b = synthetic.poisson_rhs_r(g, g, g)
db = ctx.vector(b)
dx, dr, dp = ctx.empty(n), ctx.empty(n), ctx.empty(n)
names = {0: "64x8", 1: "128x4", 2: "64x16", 3: "128x8", 4: "64x4", 5: "128x2"}
print(f"grid {g}^3 rows {n} iters {iters}; tile = points per CTA plane (2 per thread)")
for tile in tiles:
    for kc in kcs:
        os.environ["ATK_FUSED_TILE"] = str(tile)
        os.environ["ATK_FUSED_KC"] = str(kc)
        op = ctx.operator_stencil7(g, g, g, cx, cy, cz)
        for timed in (False, True):
            dx.zero(); dr.copy_from(db); dp.copy_from(db)
            res = ctx.cg_solve(op, db, dx, dr, dp, iters, time_kernels=timed)
            if not timed:
                ms_it = res["device_ms"] / iters
        assert res["fused"]
        k = [v / iters for v in res["kernel_ms"]]
        print(f"tile {tile} ({names[tile]:>6}) kc {kc:3d}: graph {ms_it:.3f} ms/it ({1000 / ms_it:.1f} it/s; 72N: {72 * n / ms_it / 1e6:.0f} GB/s, "
              f"88N-equivalent: {88 * n / ms_it / 1e6:.0f} GB/s) | KA {k[0]:.3f} ms {48 * n / (k[0] * 1e6):.0f} GB/s  "
              f"KB {k[1]:.3f} ms {24 * n / (k[1] * 1e6):.0f} GB/s", flush=True)
        op.destroy()
