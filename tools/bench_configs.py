"""One-GPU measurements of BASELINE configs 1, 2, 3 and 5 (config 4 is bench.py).  python tools/bench_configs.py"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from algotoolkit_b200 import capi, mmio, synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ctx = capi.Context(0)
out = {}

# ---- config 1: GMRES(300), tol 1e-6, bcsstk08 as the reference loads it, 100 restarts (the README run)
A = mmio.read_matrix_market_csc(os.path.join(ROOT, "tests", "golden", "data", "bcsstk08.mtx"))
op = ctx.operator_from_csc(A["n_rows"], A["n_cols"], A["col_ptr"], A["row_idx"], A["vals"])
b = op.apply_numpy(synthetic.u01_array(777, 0, A["n_cols"]))
for mode, name, restarts in ((capi.REDUCE_SEQUENTIAL, "sequential", 100), (capi.REDUCE_TREE, "tree", 100)):
    ctx.set_reduction(mode)
    t0 = time.perf_counter()
    res = ctx.gmres_solve_host(op, b, np.zeros(A["n_rows"]), restarts, 300, 1e-6)
    wall = time.perf_counter() - t0
    out[f"config1_gmres300_bcsstk08_{name}"] = {"restarts": res["restarts"], "status": res["status"], "wall_s": wall,
                                                "device_s": res["device_ms"] * 1e-3, "s_per_restart": res["device_ms"] * 1e-3 / res["restarts"],
                                                "arnoldi_steps_per_s": res["arnoldi_steps"] / (res["device_ms"] * 1e-3),
                                                "resid_first": res["resid"][:3].tolist(), "resid_last": float(res["resid"][-1])}
ctx.set_reduction(capi.REDUCE_TREE)

# ---- configs 2 and 3: Poisson CG at 32^3 (README example + RHS-R to the exit), 64^3, 256^3
for g in (32, 64, 256):
    cx, cy, cz = synthetic.poisson_coeffs(g, g, g)
    st = ctx.operator_stencil7(g, g, g, cx, cy, cz)
    bb = synthetic.poisson_rhs_r(g, g, g)
    n = g ** 3
    db = ctx.vector(bb)
    x, r, p = ctx.empty(n), ctx.empty(n), ctx.empty(n)
    for rep in range(2):
        x.zero(); r.copy_from(db); p.copy_from(db)
        res = ctx.cg_solve(st, db, x, r, p, 100000)
    out[f"poisson{g}_cg_to_exit_matrix_free"] = {"iterations": res["iters"], "stopped_on_pAp": res["stopped_on_pAp"], "device_ms": res["device_ms"],
                                                 "it_per_s": res["iters"] / (res["device_ms"] * 1e-3), "x_norm": ctx.nrm2(x)}
    t0 = time.perf_counter()
    e2e = ctx.cg_solve_host(st, bb, 100000, want_rp=False)
    out[f"poisson{g}_cg_to_exit_matrix_free"]["e2e_wall_ms"] = (time.perf_counter() - t0) * 1e3
    if g <= 256:
        asm = ctx.operator_poisson_assembled(g, g, g, cx, cy, cz, capi.FORMAT_SELL)
        for rep in range(2):
            x.zero(); r.copy_from(db); p.copy_from(db)
            res = ctx.cg_solve(asm, db, x, r, p, 100000)
        out[f"poisson{g}_cg_to_exit_assembled_sell"] = {"iterations": res["iters"], "device_ms": res["device_ms"],
                                                        "it_per_s": res["iters"] / (res["device_ms"] * 1e-3), "x_norm": ctx.nrm2(x)}
        asm.destroy()
    for v in (db, x, r, p):
        v.free()
    st.destroy()

# ---- config 5 on ONE gpu: GMRES(300), 27-point, 256^3
g = 256
n = g ** 3
op5 = ctx.operator_stencil27(g, g, g, synthetic.stencil27_coefficients())
x_ext = synthetic.u01_array(777, 0, n)
b5 = op5.apply_numpy(x_ext)
for name, ortho in (("batched", capi.GMRES_MGS_BATCHED), ("mgs", capi.GMRES_MGS)):
    res = ctx.gmres_solve_host(op5, b5, np.zeros(n), 1, 300, 1e-6, ortho=ortho)
    out[f"config5_gmres300_op27_256_1gpu_{name}"] = {"s_per_restart": res["device_ms"] * 1e-3, "arnoldi_steps_per_s": res["arnoldi_steps"] / (res["device_ms"] * 1e-3),
                                                     "resid": res["resid"].tolist(), "kernel_launches": res["kernel_launches"],
                                                     "projection_GBs": (16.0 if name == "batched" else 40.0) * n * 300 * 299 / 2 / (res["device_ms"] * 1e-3) / 1e9}
print(json.dumps(out, indent=1))
