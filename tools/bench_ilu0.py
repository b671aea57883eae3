"""Sparse ILU(0) timings on one GPU: python tools/bench_ilu0.py [grid]
Factorisation and one application (both triangular sweeps) on the 27-point operator of BASELINE config 5 and on the
assembled Poisson matrix, and GMRES(30) with / without the preconditioner."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from algotoolkit_b200 import capi, synthetic

g = int(sys.argv[1]) if len(sys.argv) > 1 else 96
ctx = capi.Context(0)
for label, make in (("op27", lambda: ctx.operator_stencil27_assembled(g, g, g, synthetic.stencil27_coefficients(), capi.FORMAT_CSR_VECTOR)),
                    ("poisson", lambda: ctx.operator_poisson_assembled(g, g, g, *synthetic.poisson_coeffs(g, g, g), capi.FORMAT_CSR_VECTOR))):
    op = make()
    n = op.n_rows_local
    rp, ci, vv = op.export_csr()
    # CSR -> CSC on the host for the C ABI (input format of the reference's containers)
    order = np.lexsort((np.repeat(np.arange(n), np.diff(rp)), ci))
    rows = np.repeat(np.arange(n), np.diff(rp))[order].astype(np.int32)
    cp = np.zeros(n + 1, np.int64)
    np.add.at(cp, ci + 1, 1)
    cp = np.cumsum(cp).astype(np.int32)
    vals = vv[order]
    ctx.synchronize()
    t0 = time.perf_counter()
    pc = ctx.ilu0_compute(n, n, cp, rows, vals)
    ctx.synchronize()
    t_fact = time.perf_counter() - t0
    b = ctx.vector(np.random.default_rng(0).standard_normal(n))
    x = ctx.empty(n)
    L = ctx.L
    for _ in range(3):
        capi._check(L.atk_ilu_solve(ctx.h, pc.h, n, b.ptr, x.ptr))
    ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        capi._check(L.atk_ilu_solve(ctx.h, pc.h, n, b.ptr, x.ptr))
    ctx.synchronize()
    t_solve = (time.perf_counter() - t0) / 20
    x_ext = synthetic.u01_array(777, 0, n)
    rhs = op.apply_numpy(x_ext)
    res = {}
    if label == "op27":
        for name, p in (("plain", None), ("ilu0", pc)):
            r = ctx.gmres_solve_host(op, rhs, np.zeros(n), 3, 30, 1e-8, precond=p)
            res[name] = (r["device_ms"], [float(v) for v in r["resid"]])
    print(f"{label} {g}^3: n = {n}, nnz = {op.nnz}: ILU(0) factorisation {t_fact * 1e3:.1f} ms (incl. CSC->CSR and level analysis), "
          f"application {t_solve * 1e3:.3f} ms ({(12.0 * op.nnz + 24.0 * n) / t_solve / 1e9:.0f} GB/s on 12 B/nnz + 24 B/row)", flush=True)
    for name, (ms, lad) in res.items():
        print(f"   GMRES(30) x3 restarts, {name}: {ms:.1f} ms, residuals {['%.3e' % v for v in lad]}")
    pc.destroy()
    op.destroy()
