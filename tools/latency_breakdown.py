"""Where a cross-GPU reduction of the persistent CG loop spends its time (profile build, one process per GPU):
    python -m algotoolkit_b200.build --profile
    ATK_LIB=profile tools/run_dist.sh N tools/latency_breakdown.py [grid] [iters]

CTA 0 of every rank stamps %globaltimer at five stations of every in-kernel reduction: (0) its own arrival, (1) all CTAs
of the rank have arrived, (2) the rank's sum is added up and stored into every rank's packed words, (3) the words of ALL
ranks have been seen locally, (4) the value is broadcast inside the CTA and the next phase starts.  Printed per rank:
the mean length of each segment for the p.Ap reduction (after phase A) and the r.r reduction (after phase B), and the
mean compute time of the two phases (station 4 of one reduction to station 0 of the next)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from algotoolkit_b200 import capi, synthetic

g = int(sys.argv[1]) if len(sys.argv) > 1 else 512
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 100
world, rank, lr = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
uid = None
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    blob = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        blob = torch.tensor(list(capi.Context.nccl_unique_id()), dtype=torch.uint8, device="cuda")
    dist.broadcast(blob, 0)
    uid = bytes(blob.cpu().tolist())
os.environ["ATK_CG_PERSISTENT"] = "1"
os.environ["ATK_CG_ONCHIP"] = "0"
ctx = capi.Context(lr, rank, world, uid)
cx, cy, cz = synthetic.poisson_coeffs(g, g, g)
op = ctx.operator_stencil7(g, g, g, cx, cy, cz)
plane = g * g
k0 = op.row_begin // plane
b = synthetic.poisson_rhs_r(g, g, g, k0, k0 + op.n_rows_local // plane)
db = ctx.vector(b)
x, r, p = ctx.empty(b.size), ctx.empty(b.size), ctx.empty(b.size)


def run():
    x.zero()
    r.copy_from(db)
    p.copy_from(db)
    cap = 5 * (2 * iters + 2)
    stamps = np.full(cap, -1.0)
    info = capi.CgInfo()
    info.time_kernels = 1
    info.rs_trace = stamps.ctypes.data_as(C.POINTER(C.c_double))
    info.trace_capacity = cap
    capi._check(ctx.L.atk_cg_solve(ctx.h, op.h, db.ptr, x.ptr, r.ptr, p.ptr, iters, C.byref(info)))
    assert info.persistent, "the persistent loop did not run"
    return stamps.reshape(-1, 5), info


run()  # warm-up
st, info = run()
# reductions: 0, 1 = start-up sums (r.r, b.b); then per iteration: p.Ap (after phase A), r.r (after phase B)
done = min(iters, info.iterations)  # the |pAp| < 1e-12 exit may end the loop early on small grids
st = st[2:2 + 2 * done]
A, B = st[0::2], st[1::2]
seg = lambda R: [float(np.mean(R[:, k + 1] - R[:, k])) * 1e-3 for k in range(4)]
phaseA = float(np.mean(A[1:, 0] - B[:-1, 4])) * 1e-3   # station 4 of the previous r.r reduction -> station 0 of this p.Ap reduction
phaseB = float(np.mean(B[:, 0] - A[:, 4])) * 1e-3
line = (f"rank {rank}/{world} grid {g}^3 ({info.device_ms / max(done, 1) * 1e3:.1f} us/iteration over {done}): "
        f"phase A compute {phaseA:7.1f} us | p.Ap reduction: arrive-spread {seg(A)[0]:5.1f}  sum+publish {seg(A)[1]:5.1f}  remote-wait {seg(A)[2]:5.1f}  resume {seg(A)[3]:4.1f} us || "
        f"phase B compute {phaseB:7.1f} us | r.r reduction: arrive-spread {seg(B)[0]:5.1f}  sum+publish {seg(B)[1]:5.1f}  remote-wait {seg(B)[2]:5.1f}  resume {seg(B)[3]:4.1f} us")
if world > 1:
    lines = [None] * world
    dist.all_gather_object(lines, line)
    if rank == 0:
        print("\n".join(lines))
    dist.barrier()
    dist.destroy_process_group()
else:
    print(line)
