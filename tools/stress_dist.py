"""Stress test of the peer-memory protocols (one process per GPU):
    ATK_EPOCH_SEED=0xFFFFFE00 [ATK_LIB=checked] tools/run_dist.sh N tools/stress_dist.py [solves]

Runs `solves` (default 1000) CG solves of mixed grid sizes and iteration counts on ONE context - persistent one-launch loop,
CUDA-graph peer loop and plain distributed applies interleaved - with the sequence-number base seeded just below 2^32, so
that the 32-bit tags of the packed {value | sequence} words wrap around during the run.  Every result is compared with
the first result obtained for the same (operator, iterations, mode) - bit for bit, since the reductions are deterministic -
and the first results are compared with the single-process oracle on rank 0 (oracle/ is the checker only)."""
import hashlib
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from algotoolkit_b200 import capi, synthetic

solves = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
world, rank, lr = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
blob = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    blob = torch.tensor(list(capi.Context.nccl_unique_id()), dtype=torch.uint8, device="cuda")
dist.broadcast(blob, 0)
ctx = capi.Context(lr, rank, world, bytes(blob.cpu().tolist()))

grids = [(16, 18, 8 * world + 3), (64, 36, 4 * world), (34, 20, 2 * world + 1), (128, 64, 6 * world + 2)]
ops, rhs = [], []
for (nx, ny, nz) in grids:
    cx, cy, cz = synthetic.poisson_coeffs(nx, ny, nz)
    op = ctx.operator_stencil7(nx, ny, nz, cx, cy, cz)
    plane = nx * ny
    k0 = op.row_begin // plane
    ops.append((op, (nx, ny, nz, cx, cy, cz)))
    rhs.append(synthetic.poisson_rhs_r(nx, ny, nz, k0, k0 + op.n_rows_local // plane, seed=4242))


def gather(local):
    t = torch.from_numpy(np.ascontiguousarray(local)).cuda()
    sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([t.numel()], dtype=torch.int64, device="cuda"))
    outs = [torch.zeros(int(s.item()), dtype=torch.float64, device="cuda") for s in sizes]
    dist.all_gather(outs, t)
    return np.concatenate([o.cpu().numpy() for o in outs])


rng = np.random.default_rng(2026)  # same stream on every rank
seen, first_x = {}, {}
bad = 0
t0 = time.time()
for s in range(solves):
    g = int(rng.integers(len(grids)))
    iters = int(rng.integers(1, 31))
    mode = ("persistent", "graph", "apply")[int(rng.integers(3))]
    op, geo = ops[g]
    if mode == "apply":
        y = op.apply_numpy(rhs[g])
        key, val = (g, 0, mode), y
    else:
        os.environ["ATK_CG_PERSISTENT"] = "1" if mode == "persistent" else "0"
        res = ctx.cg_solve_host(op, rhs[g], iters, want_rp=False)
        if res["iters"] != iters or res["persistent"] != (mode == "persistent") or not res["peer"]:
            bad += 1
        key, val = (g, iters, mode), res["x"]
    h = hashlib.sha256(val.tobytes()).hexdigest()
    if key not in seen:
        seen[key] = h
        first_x[key] = gather(val)
    elif seen[key] != h:
        bad += 1
        print(f"rank {rank}: solve {s} {key}: result differs from the first run of the same case", flush=True)
t = torch.tensor([bad], dtype=torch.int64, device="cuda")
dist.all_reduce(t)
bad = int(t.item())
ok = bad == 0
if rank == 0:
    from oracle.pyoracle import Port

    P = Port()
    worst = 0.0
    for (g, iters, mode), x in first_x.items():
        nx, ny, nz, cx, cy, cz = ops[g][1]
        b = synthetic.poisson_rhs_r(nx, ny, nz, seed=4242)
        if mode == "apply":
            want = P.poisson_apply(nx, ny, nz, cx, cy, cz, b)
            e = 0.0 if np.array_equal(x, want) else 1.0
        else:
            want = P.cg(None, b, iters, stencil=(nx, ny, nz, cx, cy, cz))["x"]
            e = float(np.linalg.norm(x - want) / max(np.linalg.norm(want), 1e-300))
        worst = max(worst, e)
    ok = ok and worst < 1e-10
    print(f"world {world}: {solves} solves / applies over {len(grids)} grids in {time.time() - t0:.1f} s, epoch seed "
          f"{os.environ.get('ATK_EPOCH_SEED', '(none)')}, lib {os.environ.get('ATK_LIB', 'release')}: {len(first_x)} distinct cases, "
          f"mismatches {bad}, worst relative error vs the oracle {worst:.2e}")
    print("STRESS", "PASS" if ok else "FAIL")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
