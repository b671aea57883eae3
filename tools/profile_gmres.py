"""GMRES(K) on the matrix-free 27-point operator at g^3 for an ncu launch list: python tools/profile_gmres.py [g] [K] [batched|mgs]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from algotoolkit_b200 import capi, synthetic

g = int(sys.argv[1]) if len(sys.argv) > 1 else 256
K = int(sys.argv[2]) if len(sys.argv) > 2 else 40
mode = sys.argv[3] if len(sys.argv) > 3 else "batched"
ctx = capi.Context(0)
n = g ** 3
op = ctx.operator_stencil27(g, g, g, synthetic.stencil27_coefficients())
b = op.apply_numpy(synthetic.u01_array(777, 0, n))
res = ctx.gmres_solve_host(op, b, np.zeros(n), 1, K, 1e-6, ortho=capi.GMRES_MGS_BATCHED if mode == "batched" else capi.GMRES_MGS)
print(mode, "g", g, "K", K, "device_ms", res["device_ms"], "launches", res["kernel_launches"], "resid", res["resid"])
