"""ILU preconditioner timings on one GPU ("next" row f4): elimination, one preconditioner application in both sweep
modes, and a preconditioned GMRES solve, on the symmetric-expanded bcsstk08 (n = 1074) and a random n = 4000 system.
    python tools/bench_ilu.py"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from algotoolkit_b200 import capi
from algotoolkit_b200.mmio import read_matrix_market_csc
from algotoolkit_b200.synthetic import u01_array

ctx = capi.Context(0)
out = {}


def run(name, n, col_ptr, row_idx, vals, gmres_K):
    t0 = time.perf_counter()
    pc = ctx.ilu_compute(n, n, col_ptr, row_idx, vals)
    ctx.synchronize()
    t_first = time.perf_counter() - t0
    pc.destroy()
    t0 = time.perf_counter()
    pc = ctx.ilu_compute(n, n, col_ptr, row_idx, vals)
    ctx.synchronize()
    t_compute = time.perf_counter() - t0
    b = ctx.vector(u01_array(99, 0, n) * 2.0 - 1.0)
    x = ctx.empty(n)
    e = {"n": n, "nnz": int(len(vals)), "compute_s_first_call": t_first, "compute_s": t_compute}
    for mode, key in ((capi.REDUCE_TREE, "tree"), (capi.REDUCE_SEQUENTIAL, "sequential")):
        ctx.set_reduction(mode)
        for _ in range(2):
            capi._check(ctx.L.atk_ilu_solve(ctx.h, pc.h, n, b.ptr, x.ptr))
        ctx.synchronize()
        reps = 10
        t0 = time.perf_counter()
        for _ in range(reps):
            capi._check(ctx.L.atk_ilu_solve(ctx.h, pc.h, n, b.ptr, x.ptr))
        ctx.synchronize()
        e[f"solve_ms_{key}"] = (time.perf_counter() - t0) / reps * 1e3
    ctx.set_reduction(capi.REDUCE_TREE)
    op = ctx.operator_from_csc(n, n, col_ptr, row_idx, vals)
    x_ext = u01_array(777, 0, n)
    rhs = op.apply_numpy(x_ext)
    g = ctx.gmres_solve_host(op, rhs, np.zeros(n), 3, gmres_K, 1e-10, precond=pc)
    e["gmres_precond"] = {"K": gmres_K, "restarts": g["restarts"], "arnoldi_steps": g["arnoldi_steps"], "device_ms": g["device_ms"],
                          "resid": [float(v) for v in g["resid"]], "x_err_vs_ext": float(np.linalg.norm(g["x"] - x_ext) / np.linalg.norm(x_ext))}
    pc.destroy()
    out[name] = e


d = read_matrix_market_csc(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "data", "bcsstk08.mtx"), expand_symmetric=True)
run("bcsstk08_symmetric", d["n_rows"], d["col_ptr"], d["row_idx"], d["vals"], 30)
n = 4000
rng = np.random.default_rng(4)
import scipy.sparse as sp

A = sp.random(n, n, density=0.002, random_state=rng, format="csc", data_rvs=rng.standard_normal)
A = (A + sp.diags(np.asarray(abs(A).sum(axis=1)).ravel() + 1.0)).tocsc()
A.sort_indices()
run("random_4000", n, A.indptr, A.indices, A.data, 30)
print(json.dumps(out, indent=1))
