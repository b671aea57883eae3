"""Plain 7-point apply at 512^3: K1 inside the three-kernel CG (CUDA events) vs back-to-back launches (wall clock),
LDGSTS ring vs TMA ring.  python tools/ab_plain_apply.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from algotoolkit_b200 import capi, synthetic
os.environ["ATK_CG_FUSED"] = "0"
g = 512; iters = 20
ctx = capi.Context(0)
cx, cy, cz = synthetic.poisson_coeffs(g, g, g)
n = g ** 3
b = synthetic.poisson_rhs_r(g, g, g)
db = ctx.vector(b)
dx, dr, dp = ctx.empty(n), ctx.empty(n), ctx.empty(n)
op = ctx.operator_stencil7(g, g, g, cx, cy, cz)
for rep in range(3):
    dx.zero(); dr.copy_from(db); dp.copy_from(db)
    res = ctx.cg_solve(op, db, dx, dr, dp, iters, time_kernels=True)
    k = [v / iters for v in res["kernel_ms"]]
    y = ctx.empty(n)
    for _ in range(3): op.apply(db, y)
    ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(20): op.apply(db, y)
    ctx.synchronize()
    t = (time.perf_counter() - t0) / 20
    print(f"rep {rep}: K1 (apply + dot, events) {k[0]:.4f} ms   plain apply (20 back-to-back, wall) {t*1e3:.4f} ms", flush=True)


for tma in ("0", "1", "0", "1"):
    os.environ["ATK_STENCIL_TMA"] = tma
    op2 = ctx.operator_stencil7(g, g, g, cx, cy, cz)
    y2 = ctx.empty(n)
    for _ in range(3): op2.apply(db, y2)
    ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(20): op2.apply(db, y2)
    ctx.synchronize()
    t = (time.perf_counter() - t0) / 20
    print(f"tma={tma}: plain apply {t*1e3:.4f} ms  ({16 * n / t / 1e9:.0f} GB/s)", flush=True)
    op2.destroy()
    y2.free()
