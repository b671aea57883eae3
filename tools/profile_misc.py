"""Secondary device loops for an ncu launch list: GradientDescent, CG on the assembled SELL operator (three-kernel form),
Arnoldi, operator assembly / conversion.  python tools/profile_misc.py [g]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from algotoolkit_b200 import capi, synthetic

g = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ctx = capi.Context(0)
n = g ** 3
cx, cy, cz = synthetic.poisson_coeffs(g, g, g)
b = synthetic.poisson_rhs_r(g, g, g)
st = ctx.operator_stencil7(g, g, g, cx, cy, cz)
sell = ctx.operator_poisson_assembled(g, g, g, cx, cy, cz, capi.FORMAT_SELL)
r = capi.gd_solve_host(ctx, st, b, np.zeros(n), 20, 1e-30)
print("gd iters", r["iters"], "ms", r["device_ms"])
r = ctx.cg_solve_host(sell, b, 20)
print("cg sell iters", r["iters"], "ms", r["device_ms"])
q0 = b / np.linalg.norm(b)
a = capi.arnoldi_host(ctx, st, q0, 10, 1e-12)
print("arnoldi ok", a["breakdown"])
