"""BASELINE config 5: GMRES(300) on the synthetic 27-point operator, 256^3 = 16.7 M rows, sharded over the ranks.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_config5.py [grid] [K] [restarts] [modes]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from algotoolkit_b200 import capi, synthetic

g = int(sys.argv[1]) if len(sys.argv) > 1 else 256
K = int(sys.argv[2]) if len(sys.argv) > 2 else 300
restarts = int(sys.argv[3]) if len(sys.argv) > 3 else 1
modes = sys.argv[4].split(",") if len(sys.argv) > 4 else ["batched", "mgs"]
world, rank, lr = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
uid = None
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    blob = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        blob = torch.tensor(list(capi.Context.nccl_unique_id()), dtype=torch.uint8, device="cuda")
    dist.broadcast(blob, 0)
    uid = bytes(blob.cpu().tolist())
ctx = capi.Context(lr, rank, world, uid)
op = ctx.operator_stencil27(g, g, g, synthetic.stencil27_coefficients())
n_loc, r0 = op.n_rows_local, op.row_begin
x_ext = synthetic.u01_array(777, r0, n_loc)          # x_ext[i] = u01(777 + i), this rank's rows
b = op.apply_numpy(x_ext)                            # b = A x_ext (halo exchange inside)
out = {"grid": g, "rows": g ** 3, "K": K, "restarts": restarts, "n_gpus": world}
for mode in modes:
    ortho = capi.GMRES_MGS_BATCHED if mode == "batched" else capi.GMRES_MGS
    res = ctx.gmres_solve_host(op, b, np.zeros(n_loc), restarts, K, 1e-6, ortho=ortho)
    ms = res["device_ms"]
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    # true residual of the returned x, and distance to x_ext (global norms through the library)
    dx = ctx.vector(res["x"])
    ax = ctx.empty(n_loc)
    op.apply(dx, ax)
    rr = ctx.sub(ctx.vector(b), ax)
    err = ctx.sub(dx, ctx.vector(x_ext))
    bytes_mgs = (16.0 if mode == "batched" else 40.0) * g ** 3 * K * (K - 1) / 2 * res["restarts"]
    out[mode] = {"s_per_restart": ms * 1e-3 / max(res["restarts"], 1), "arnoldi_steps_per_s": res["arnoldi_steps"] / (ms * 1e-3),
                 "resid": res["resid"].tolist(), "true_residual": ctx.nrm2(rr), "x_err_vs_ext": ctx.nrm2(err) / ctx.nrm2(ctx.vector(x_ext)),
                 "projection_GBs_total": bytes_mgs / (ms * 1e-3) / 1e9, "kernel_launches": res["kernel_launches"]}
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
