"""Matrix-free 27-point apply: time per launch for chunk lengths.  python tools/sweep_stencil27.py [g]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from algotoolkit_b200 import capi, synthetic

g = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ctx = capi.Context(0)
n = g ** 3
x = ctx.vector(synthetic.u01_array(777, 0, n))
y = ctx.empty(n)
for kc, simple in ((64, "1"), (8, ""), (16, ""), (32, ""), (64, ""), (128, "")):
    os.environ["ATK_STENCIL27_KC"] = str(kc)
    if simple:
        os.environ["ATK_STENCIL27_SIMPLE"] = "1"
    else:
        os.environ.pop("ATK_STENCIL27_SIMPLE", None)
    op = ctx.operator_stencil27(g, g, g, synthetic.stencil27_coefficients())
    for _ in range(3): op.apply(x, y)
    ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(20): op.apply(x, y)
    ctx.synchronize()
    t = (time.perf_counter() - t0) / 20
    print(f"g {g} {'simple kernel' if simple else 'smem kernel kc %3d' % kc}: {t*1e3:.4f} ms  {16 * n / t / 1e9:.0f} GB/s (16 B/row)", flush=True)
    op.destroy()
