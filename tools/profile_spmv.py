"""SpMV launches for ncu: python tools/profile_spmv.py [grid]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from algotoolkit_b200 import capi, synthetic

g = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ctx = capi.Context(0)
cx, cy, cz = synthetic.poisson_coeffs(g, g, g)
n = g ** 3
x = ctx.vector(np.random.default_rng(0).standard_normal(n))
y = ctx.empty(n)
for name, op in (("stencil", ctx.operator_stencil7(g, g, g, cx, cy, cz)),
                 ("sell", ctx.operator_poisson_assembled(g, g, g, cx, cy, cz, capi.FORMAT_SELL)),
                 ("csrv", ctx.operator_poisson_assembled(g, g, g, cx, cy, cz, capi.FORMAT_CSR_VECTOR))):
    for _ in range(4):
        op.apply(x, y)
    ctx.synchronize()
    print(name, "ok")
