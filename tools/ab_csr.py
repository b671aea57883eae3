"""CSR-format SpMV on the assembled Poisson matrix: CSR-stream vs CSR-vector vs SELL.  python tools/ab_csr.py [g]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from algotoolkit_b200 import capi, synthetic

g = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ctx = capi.Context(0)
n = g ** 3
cx, cy, cz = synthetic.poisson_coeffs(g, g, g)
x = ctx.vector(np.random.default_rng(0).standard_normal(n))
y = ctx.empty(n)
for name, fmt, env in (("sell", capi.FORMAT_SELL, None), ("csr-stream", capi.FORMAT_CSR_VECTOR, None),
                       ("csr-vector", capi.FORMAT_CSR_VECTOR, "1"), ("csr-stream", capi.FORMAT_CSR_VECTOR, None)):
    if env:
        os.environ["ATK_CSR_NO_STREAM"] = env
    else:
        os.environ.pop("ATK_CSR_NO_STREAM", None)
    op = ctx.operator_poisson_assembled(g, g, g, cx, cy, cz, fmt)
    alg = 12.0 * op.nnz + 4.0 * (n + 1) + 16.0 * n
    for _ in range(3): op.apply(x, y)
    ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(20): op.apply(x, y)
    ctx.synchronize()
    t = (time.perf_counter() - t0) / 20
    print(f"{name:11s} g {g}: {t*1e3:.4f} ms  {alg / t / 1e9:.0f} GB/s", flush=True)
    op.destroy()
