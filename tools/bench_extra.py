"""Secondary measurements (BASELINE configs 1, 3, 5 on one GPU): SpMV GB/s, CG matrix-free vs assembled, GMRES.
    python tools/bench_extra.py [quick]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from algotoolkit_b200 import capi, synthetic

quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
ctx = capi.Context(0)
PEAK = 6556.5
out = {}


def time_apply(op, n, reps=20):
    x = ctx.vector(np.random.default_rng(0).standard_normal(n))
    y = ctx.empty(op.n_rows_local)
    for _ in range(3):
        op.apply(x, y)
    ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        op.apply(x, y)
    ctx.synchronize()
    return (time.perf_counter() - t0) / reps


# ---- config 3: 256^3, SpMV alone and CG, matrix-free vs assembled
g = 128 if quick else 256
n = g ** 3
cx, cy, cz = synthetic.poisson_coeffs(g, g, g)
b = synthetic.poisson_rhs_r(g, g, g)
t0 = time.perf_counter()
ops = {"stencil7": ctx.operator_stencil7(g, g, g, cx, cy, cz)}
t1 = time.perf_counter()
ops["sell"] = ctx.operator_poisson_assembled(g, g, g, cx, cy, cz, capi.FORMAT_SELL)
ctx.synchronize()
t2 = time.perf_counter()
ops["csr_stream"] = ctx.operator_poisson_assembled(g, g, g, cx, cy, cz, capi.FORMAT_CSR_VECTOR)
ctx.synchronize()
t3 = time.perf_counter()
nnz = ops["sell"].nnz
out["assemble_seconds"] = {"sell_device_assembly_and_conversion": t2 - t1, "csr_device_assembly_and_conversion": t3 - t2}
for name, op in ops.items():
    t = time_apply(op, n)
    alg = 16.0 * n if name == "stencil7" else 12.0 * nnz + 4.0 * (n + 1) + 16.0 * n
    out[f"spmv_{name}_{g}"] = {"ms": t * 1e3, "algorithmic_GB": alg / 1e9, "GBs": alg / t / 1e9, "frac_of_measured_peak": alg / t / 1e9 / PEAK}
db = ctx.vector(b)
for name, op in ops.items():
    x, r, p = ctx.empty(n), ctx.empty(n), ctx.empty(n)
    for rep in range(2):
        x.zero(); r.copy_from(db); p.copy_from(db)
        res = ctx.cg_solve(op, db, x, r, p, 200)
    alg = (72.0 if res["fused"] else 88.0) * n + (0 if name == "stencil7" else 12.0 * nnz + 4.0 * (n + 1))
    out[f"cg_{name}_{g}"] = {"iters": res["iters"], "it_per_s": res["iters"] / (res["device_ms"] * 1e-3), "fused": res["fused"],
                             "GBs": alg * res["iters"] / (res["device_ms"] * 1e-3) / 1e9}
for op in ops.values():
    op.destroy()

# ---- config 1: GMRES(300) on bcsstk08
from oracle.pyoracle import Port  # only to read the matrix file and to build b (inputs), and as the CPU timing baseline

P = Port()
A = P.read_matrix_market(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/golden/data/bcsstk08.mtx"))
bb = P.spmv(A, P.u01_array(777, A.n_cols))
op = ctx.operator_from_csc(A.n_rows, A.n_cols, A.col_ptr, A.row_idx, A.vals)
for mode, name in ((capi.REDUCE_SEQUENTIAL, "sequential"), (capi.REDUCE_TREE, "tree")):
    ctx.set_reduction(mode)
    restarts = 1 if quick else 3
    res = ctx.gmres_solve_host(op, bb, np.zeros(A.n_rows), restarts, 300, 1e-6)
    out[f"gmres300_bcsstk08_{name}"] = {"restarts": res["restarts"], "s_per_restart": res["device_ms"] * 1e-3 / res["restarts"],
                                        "arnoldi_steps_per_s": res["arnoldi_steps"] / (res["device_ms"] * 1e-3),
                                        "kernel_launches": res["kernel_launches"], "resid": res["resid"].tolist()}
ctx.set_reduction(capi.REDUCE_TREE)
t0 = time.perf_counter()
P.gmres(A, bb, np.zeros(A.n_rows), 1, 300, 1e-6)
out["gmres300_bcsstk08_cpu_port_s_per_restart"] = time.perf_counter() - t0

# ---- config 5 operator on one GPU: 27-point, K = 50
g5 = 64 if quick else 128
n5 = g5 ** 3
op5 = ctx.operator_stencil27_assembled(g5, g5, g5, synthetic.stencil27_coefficients())
x_ext = synthetic.u01_array(777, 0, n5)
b5 = op5.apply_numpy(x_ext)
K = 50
res = ctx.gmres_solve_host(op5, b5, np.zeros(n5), 1, K, 1e-6)
mgs_bytes = 40.0 * n5 * K * (K - 1) / 2 + 24.0 * n5 * K + K * (12.0 * op5.nnz + 20.0 * n5)
out[f"gmres{K}_op27_{g5}"] = {"s_per_restart": res["device_ms"] * 1e-3, "arnoldi_steps_per_s": res["arnoldi_steps"] / (res["device_ms"] * 1e-3),
                              "algorithmic_GBs": mgs_bytes / (res["device_ms"] * 1e-3) / 1e9, "resid": res["resid"].tolist(),
                              "x_err_vs_ext": float(np.linalg.norm(res["x"] - x_ext) / np.linalg.norm(x_ext))}
for name, ortho in (("batched", capi.GMRES_MGS_BATCHED),):
    res = ctx.gmres_solve_host(op5, b5, np.zeros(n5), 1, K, 1e-6, ortho=ortho)
    out[f"gmres{K}_op27_{g5}_{name}"] = {"s_per_restart": res["device_ms"] * 1e-3, "arnoldi_steps_per_s": res["arnoldi_steps"] / (res["device_ms"] * 1e-3),
                                          "resid": res["resid"].tolist()}
os.environ["ATK_GMRES_NO_PERSISTENT"] = "1"
res = ctx.gmres_solve_host(op5, b5, np.zeros(n5), 1, K, 1e-6)
out[f"gmres{K}_op27_{g5}_per_projection"] = {"s_per_restart": res["device_ms"] * 1e-3, "resid": res["resid"].tolist()}
del os.environ["ATK_GMRES_NO_PERSISTENT"]
# a size whose slices do not fit on chip (persistent kernel not applicable): 256^3 = 16.7 M rows, K = 40
if not quick:
    g6 = 256
    n6 = g6 ** 3
    op6 = ctx.operator_stencil27(g6, g6, g6, synthetic.stencil27_coefficients())
    b6 = op6.apply_numpy(synthetic.u01_array(777, 0, n6))
    K6 = 40
    for name, ortho in (("per_projection", capi.GMRES_MGS), ("batched", capi.GMRES_MGS_BATCHED)):
        res = ctx.gmres_solve_host(op6, b6, np.zeros(n6), 1, K6, 1e-6, ortho=ortho)
        byts = (40.0 if ortho == capi.GMRES_MGS else 16.0) * n6 * K6 * (K6 - 1) / 2
        out[f"gmres{K6}_op27_{g6}_{name}"] = {"s_per_restart": res["device_ms"] * 1e-3, "projection_GBs": byts / (res["device_ms"] * 1e-3) / 1e9,
                                              "resid": res["resid"].tolist()}
print(json.dumps(out, indent=1))
