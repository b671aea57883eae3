"""Persistent Arnoldi-MGS step at the per-GPU size of BASELINE config 5 on 8 GPUs (2.1 M rows), on ONE GPU:
    python tools/mgs_probe.py [grid=128] [K=300]
GMRES(K), one restart, exact MGS (persistent step kernel, w in shared memory) and batched MGS; seconds and microseconds per projection."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from algotoolkit_b200 import capi, synthetic

g = int(sys.argv[1]) if len(sys.argv) > 1 else 128
K = int(sys.argv[2]) if len(sys.argv) > 2 else 300
ctx = capi.Context(0)
op = ctx.operator_stencil27(g, g, g, synthetic.stencil27_coefficients())
n = op.n_rows_local
b = op.apply_numpy(synthetic.u01_array(777, 0, n))
for name, ortho in (("exact_mgs", capi.GMRES_MGS), ("batched_mgs", capi.GMRES_MGS_BATCHED)):
    for rep in range(2):
        r = ctx.gmres_solve_host(op, b, np.zeros(n), 1, K, 1e-6, ortho=ortho)
    print(f"{g}^3 K={K} {name}: {r['device_ms'] * 1e-3:.4f} s per restart, {r['device_ms'] * 1e3 / (K * (K - 1) / 2):.2f} us per projection, "
          f"launches {r['kernel_launches']}, residual {r['resid'][-1]:.12e}")
