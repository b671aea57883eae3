"""Short CG run for ncu: python tools/profile_cg.py [grid] [iters] [fused 0|1]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
g = int(sys.argv[1]) if len(sys.argv) > 1 else 512
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 6
if len(sys.argv) > 3:
    os.environ["ATK_CG_FUSED"] = sys.argv[3]
from algotoolkit_b200 import capi, synthetic

ctx = capi.Context(0)
cx, cy, cz = synthetic.poisson_coeffs(g, g, g)
n = g ** 3
db = ctx.vector(synthetic.poisson_rhs_r(g, g, g))
dx, dr, dp = ctx.empty(n), ctx.empty(n), ctx.empty(n)
op = ctx.operator_stencil7(g, g, g, cx, cy, cz)
dx.zero(); dr.copy_from(db); dp.copy_from(db)
res = ctx.cg_solve(op, db, dx, dr, dp, iters, time_kernels=True)  # kernel-by-kernel: ncu sees plain launches
print("iters", res["iters"], "fused", res["fused"], "kernel_ms/iter", [v / iters for v in res["kernel_ms"]])
