"""CSR-format SpMV variants on the assembled 256^3 Poisson and 128^3 27-point matrices: python tools/csr_probe.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from algotoolkit_b200 import capi, synthetic

ctx = capi.Context(0)
peak = 6556.5
for label, make in (("poisson256", lambda fmt: ctx.operator_poisson_assembled(256, 256, 256, *synthetic.poisson_coeffs(256, 256, 256), fmt)),
                    ("op27_128", lambda fmt: ctx.operator_stencil27_assembled(128, 128, 128, synthetic.stencil27_coefficients(), fmt))):
    for variant, env in (("sell", {}), ("csr_warp_stream", {}), ("csr_stream_v1", {"ATK_CSR_STREAM_V1": "1"})):
        for k in ("ATK_CSR_STREAM_V1",):
            os.environ.pop(k, None)
        os.environ.update(env)
        op = make(capi.FORMAT_SELL if variant == "sell" else capi.FORMAT_CSR_VECTOR)
        x = ctx.vector(np.random.default_rng(0).standard_normal(op.n_cols))
        y = ctx.empty(op.n_rows_local)
        for _ in range(3):
            op.apply(x, y)
        ctx.synchronize()
        t0 = time.perf_counter()
        for _ in range(20):
            op.apply(x, y)
        ctx.synchronize()
        dt = (time.perf_counter() - t0) / 20
        alg = 12.0 * op.nnz + 4.0 * (op.n_rows_local + 1) + 16.0 * op.n_rows_local
        print(f"{label:12s} {variant:16s} {dt * 1e3:8.4f} ms  {alg / dt / 1e9:8.1f} GB/s  {alg / dt / 1e9 / peak:6.3f} of measured peak", flush=True)
        x.free(); y.free(); op.destroy()
