import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from algotoolkit_b200 import capi, synthetic
ctx = capi.Context(0)
for g in (256, 128):
    n = g ** 3
    cx, cy, cz = synthetic.poisson_coeffs(g, g, g)
    x = ctx.vector(np.random.default_rng(0).standard_normal(n)); y = ctx.empty(n)
    for kc in (64, 32, 16, 8):
        os.environ["ATK_STENCIL_KC"] = str(kc)
        op = ctx.operator_stencil7(g, g, g, cx, cy, cz)
        for _ in range(5): op.apply(x, y)
        ctx.synchronize(); t0 = time.perf_counter()
        for _ in range(50): op.apply(x, y)
        ctx.synchronize(); t = (time.perf_counter() - t0) / 50
        print(f"g {g} kc {kc}: {t*1e3:.4f} ms {16*n/t/1e9:.0f} GB/s", flush=True)
        op.destroy()
    x.free(); y.free()
