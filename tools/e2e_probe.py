"""Wall time of repeated host-buffer solves (one GPU): python tools/e2e_probe.py [grid] [iters]"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from algotoolkit_b200 import capi, synthetic

g = int(sys.argv[1]) if len(sys.argv) > 1 else 512
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 100
ctx = capi.Context(0)
cx, cy, cz = synthetic.poisson_coeffs(g, g, g)
op = ctx.operator_stencil7(g, g, g, cx, cy, cz)
n = g ** 3
L = ctx.L
hb, hx = C.c_void_p(), C.c_void_p()
capi._check(L.atk_host_alloc(n * 8, C.byref(hb)))
capi._check(L.atk_host_alloc(n * 8, C.byref(hx)))
b = np.ctypeslib.as_array(C.cast(hb, C.POINTER(C.c_double)), shape=(n,))
synthetic.poisson_rhs_r(g, g, g, out=b)
for rep in range(4):
    info = capi.CgInfo()
    ctx.synchronize()
    t0 = time.perf_counter()
    capi._check(L.atk_cg_solve_host(ctx.h, op.h, hb, hx, None, None, 1, iters, C.byref(info)))
    ctx.synchronize()
    t = time.perf_counter() - t0
    print(f"call {rep}: wall {t * 1e3:.1f} ms, device loop {info.device_ms:.1f} ms, overhead {t * 1e3 - info.device_ms:.1f} ms, "
          f"{iters / t:.1f} it/s e2e")
# pageable numpy buffers for comparison
bp, xp = np.array(b), np.zeros(n)
for rep in range(2):
    t0 = time.perf_counter()
    res = ctx.cg_solve_host(op, bp, iters, want_rp=False)
    t = time.perf_counter() - t0
    print(f"pageable call {rep}: wall {t * 1e3:.1f} ms, device loop {res['device_ms']:.1f} ms")
