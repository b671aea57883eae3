"""Each process runs an INDEPENDENT single-GPU CG on one 512x512x(512/N) slab-sized grid (no communication at all):
separates slab-size / per-GPU effects from synchronisation cost in the multi-GPU numbers.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... tools/slab_probe.py [nz] [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from algotoolkit_b200 import capi, synthetic

nz = int(sys.argv[1]) if len(sys.argv) > 1 else 64
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 300
lr = int(os.environ.get("LOCAL_RANK", 0))
ctx = capi.Context(lr)
g = 512
cx, cy, cz = synthetic.poisson_coeffs(g, g, 512)
op = ctx.operator_stencil7(g, g, nz, cx, cy, cz)
n = g * g * nz
b = synthetic.poisson_rhs_r(g, g, nz)
db = ctx.vector(b)
x, r, p = ctx.empty(n), ctx.empty(n), ctx.empty(n)
out = []
for timed in (False, False, True):
    x.zero(); r.copy_from(db); p.copy_from(db)
    res = ctx.cg_solve(op, db, x, r, p, iters, time_kernels=timed)
    out.append(res)
ms_it = out[1]["device_ms"] / iters
k = [v / iters for v in out[2]["kernel_ms"]]
print(f"gpu {lr}: slab 512x512x{nz}: {ms_it * 1e3:.1f} us/iteration ({1000 / ms_it:.0f} it/s)  KA {k[0] * 1e3:.1f} us ({48 * n / k[0] / 1e6:.0f} GB/s)  "
      f"KB {k[1] * 1e3:.1f} us ({24 * n / k[1] / 1e6:.0f} GB/s)  ideal at 1-GPU rates: {1580.0 * nz / 512:.1f} us", flush=True)
