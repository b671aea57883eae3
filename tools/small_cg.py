"""CG to the |pAp| exit on small grids: on-chip cooperative loop vs the two-kernel graph loop.  python tools/small_cg.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from algotoolkit_b200 import capi, synthetic

ctx = capi.Context(0)
for g in (16, 32, 48, 64, 80):
    cx, cy, cz = synthetic.poisson_coeffs(g, g, g)
    b = synthetic.poisson_rhs_r(g, g, g)
    row = {}
    for mode in ("1", "0"):
        os.environ["ATK_CG_ONCHIP"] = mode
        op = ctx.operator_stencil7(g, g, g, cx, cy, cz)
        for _ in range(2):
            res = ctx.cg_solve_host(op, b, 100000)
        t0 = time.perf_counter()
        res = ctx.cg_solve_host(op, b, 100000)
        wall = time.perf_counter() - t0
        row[mode] = (res["iters"], res["device_ms"], wall * 1e3, float(np.linalg.norm(res["x"])), res["kernel_launches"])
        op.destroy()
    a, c = row["1"], row["0"]
    print(f"{g}^3: on-chip {a[0]} it {a[1]:.3f} ms ({a[1]/a[0]*1e3:.2f} us/it, e2e {a[2]:.2f} ms, {a[4]} launches) | graph loop {c[0]} it {c[1]:.3f} ms "
          f"({c[1]/c[0]*1e3:.2f} us/it, e2e {c[2]:.2f} ms) | |x| rel diff {abs(a[3]-c[3])/c[3]:.1e}", flush=True)
