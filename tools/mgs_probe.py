"""Persistent Arnoldi-MGS step: GMRES(K), one restart, on the 27-point operator at grid^3, exact MGS (persistent step kernel)
and batched MGS; seconds per restart and microseconds per projection (operator applies included).
    python tools/mgs_probe.py [grid=128] [K=300] [reps=2] [exact|batched]   one GPU; 128^3 = the per-GPU slice of BASELINE config 5 on 8 GPUs
    tools/run_dist.sh 8 tools/mgs_probe.py 256 300             BASELINE config 5 itself, z-slabs over 8 GPUs"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from algotoolkit_b200 import capi, synthetic

g = int(sys.argv[1]) if len(sys.argv) > 1 else 128
K = int(sys.argv[2]) if len(sys.argv) > 2 else 300
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2          # the last repetition is reported
only = sys.argv[4] if len(sys.argv) > 4 else ""              # "exact" or "batched": that mode only
world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
if world > 1:
    import torch
    import torch.distributed as dist

    lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    blob = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        blob = torch.tensor(list(capi.Context.nccl_unique_id()), dtype=torch.uint8, device="cuda")
    dist.broadcast(blob, 0)
    ctx = capi.Context(lr, rank, world, bytes(blob.cpu().tolist()))
else:
    ctx = capi.Context(0)
op = ctx.operator_stencil27(g, g, g, synthetic.stencil27_coefficients())
n = op.n_rows_local
b = op.apply_numpy(synthetic.u01_array(777, op.row_begin, n))
for name, ortho in (("exact_mgs", capi.GMRES_MGS), ("batched_mgs", capi.GMRES_MGS_BATCHED)):
    if only and not name.startswith(only):
        continue
    for rep in range(reps):
        r = ctx.gmres_solve_host(op, b, np.zeros(n), 1, K, 1e-6, ortho=ortho)
    ms = r["device_ms"]
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        print(f"{g}^3 K={K} world {world} {name}: {ms * 1e-3:.4f} s per restart, {ms * 1e3 / (K * (K - 1) / 2):.2f} us per projection, "
              f"launches {r['kernel_launches']}, residual {r['resid'][-1]:.12e}", flush=True)
if world > 1:
    dist.destroy_process_group()
