"""Multi-GPU parity check, one process per GPU:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py [grid] [iters]
Every rank owns a z-slab; rank 0 gathers x and compares with the single-process oracle (oracle/ is the checker only)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from algotoolkit_b200 import capi, synthetic

g = int(sys.argv[1]) if len(sys.argv) > 1 else 48
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 60
world, rank, lr = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
blob = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    blob = torch.tensor(list(capi.Context.nccl_unique_id()), dtype=torch.uint8, device="cuda")
dist.broadcast(blob, 0)
ctx = capi.Context(lr, rank, world, bytes(blob.cpu().tolist()))
nx, ny, nz = g, g + 2, g + 5  # uneven on purpose: nz not a multiple of the rank count
cx, cy, cz = synthetic.poisson_coeffs(nx, ny, nz)
op = ctx.operator_stencil7(nx, ny, nz, cx, cy, cz)
plane = nx * ny
k0, k1 = op.row_begin // plane, (op.row_begin + op.n_rows_local) // plane
b_loc = synthetic.poisson_rhs_r(nx, ny, nz, k0, k1)


def gather(local):
    t = torch.from_numpy(np.ascontiguousarray(local)).cuda()
    sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([t.numel()], dtype=torch.int64, device="cuda"))
    outs = [torch.zeros(int(s.item()), dtype=torch.float64, device="cuda") for s in sizes]
    dist.all_gather(outs, t)
    return np.concatenate([o.cpu().numpy() for o in outs])


ok = True
results = {}
# "1": fused, persistent one-launch loop (peer memory); "graph": fused two-kernel CUDA-graph loop (peer memory);
# "0": generic three-kernel loop over NCCL
for fused in ("1", "graph", "0"):
    os.environ["ATK_CG_FUSED"] = "0" if fused == "0" else "1"
    os.environ["ATK_CG_PERSISTENT"] = "0" if fused == "graph" else "1"
    res = ctx.cg_solve_host(op, b_loc, iters, trace=True)
    results[fused] = (gather(res["x"]), gather(res["r"]), gather(res["p"]), res)
os.environ["ATK_CG_PERSISTENT"] = "1"
# a second fused solve on the same context (sequence-number epochs must advance), continuing from the first state
os.environ["ATK_CG_FUSED"] = "1"
x1, r1, p1, _ = [results["1"][q] for q in range(4)]
sl = slice(op.row_begin, op.row_begin + op.n_rows_local)
res2 = ctx.cg_solve_host(op, b_loc, 7, state=(x1[sl], r1[sl], p1[sl]))
x_cont = gather(res2["x"])
# latency floor of one iteration on this (small) grid, warm, per communication mode
lat = {}
for mode in ("persistent", "peer", "nccl"):
    os.environ["ATK_CG_FUSED"] = "1"
    os.environ["ATK_CG_FORCE_NCCL"] = "1" if mode == "nccl" else "0"
    os.environ["ATK_CG_PERSISTENT"] = "1" if mode == "persistent" else "0"
    for rep in range(2):
        rr = ctx.cg_solve_host(op, b_loc, 200)
    lat[mode] = (rr["device_ms"] / max(rr["iters"], 1) * 1e3, rr["peer"], rr["iters"])
os.environ["ATK_CG_FORCE_NCCL"] = "0"
os.environ["ATK_CG_PERSISTENT"] = "1"
# plain operator apply with halo exchange
xin = np.random.default_rng(3).standard_normal(nx * ny * nz)
y_loc = op.apply_numpy(xin[op.row_begin: op.row_begin + op.n_rows_local])
y_all = gather(y_loc)
d = ctx.dot(ctx.vector(b_loc), ctx.vector(b_loc))
# 27-point operator of BASELINE config 5, matrix-free and sharded: apply + a short GMRES
coef = synthetic.stencil27_coefficients()
op27 = ctx.operator_stencil27(nx, ny, nz, coef)
y27 = gather(op27.apply_numpy(xin[op27.row_begin: op27.row_begin + op27.n_rows_local]))
x_ext = synthetic.u01_array(777, 0, nx * ny * nz)
b27_loc = op27.apply_numpy(x_ext[op27.row_begin: op27.row_begin + op27.n_rows_local])
g27 = ctx.gmres_solve_host(op27, b27_loc, np.zeros(op27.n_rows_local), 2, 20, 1e-6)
x27 = gather(g27["x"])
g27b = ctx.gmres_solve_host(op27, b27_loc, np.zeros(op27.n_rows_local), 2, 20, 1e-6, ortho=capi.GMRES_MGS_BATCHED)
x27b = gather(g27b["x"])
# the per-projection path (what slices too large for the chip run on): fused update+products launch + all-gather per
# projection, and the dot / all-gather / axmy form it replaces
os.environ["ATK_GMRES_NO_PERSISTENT"] = "1"
g27p = ctx.gmres_solve_host(op27, b27_loc, np.zeros(op27.n_rows_local), 2, 20, 1e-6)
os.environ["ATK_GMRES_NO_FUSED_DOT"] = "1"
g27u = ctx.gmres_solve_host(op27, b27_loc, np.zeros(op27.n_rows_local), 2, 20, 1e-6)
del os.environ["ATK_GMRES_NO_FUSED_DOT"], os.environ["ATK_GMRES_NO_PERSISTENT"]
x27p, x27u = gather(g27p["x"]), gather(g27u["x"])
# the same at slices of 4428 rows per CTA (128 x 128 x 40*world): the Arnoldi step runs as the cp.async-ring kernel (w in
# registers); checked against the shared-memory step kernel and the batched form (no CPU oracle at 0.65 M rows per rank)
opr = ctx.operator_stencil27(128, 128, 40 * world, coef)
br_loc = opr.apply_numpy(synthetic.u01_array(778, opr.row_begin, opr.n_rows_local))
g_ring = ctx.gmres_solve_host(opr, br_loc, np.zeros(opr.n_rows_local), 2, 20, 1e-6)
os.environ["ATK_MGS_NO_RING"] = "1"
g_smem = ctx.gmres_solve_host(opr, br_loc, np.zeros(opr.n_rows_local), 2, 20, 1e-6)
del os.environ["ATK_MGS_NO_RING"]
g_bat = ctx.gmres_solve_host(opr, br_loc, np.zeros(opr.n_rows_local), 2, 20, 1e-6, ortho=capi.GMRES_MGS_BATCHED)
ring_ok = g_ring["kernel_launches"] < 2 * (20 * 3 + 20)
for other in (g_smem, g_bat):
    ring_ok = ring_ok and np.allclose(g_ring["resid"], other["resid"], rtol=1e-10, atol=0.0)
    ring_ok = ring_ok and np.linalg.norm(g_ring["x"] - other["x"]) <= 1e-10 * np.linalg.norm(other["x"])
ring_flags = gather(np.array([1.0 if ring_ok else 0.0]))
ring_ms = (g_ring["device_ms"], g_smem["device_ms"], g_bat["device_ms"])
# assembled operators partitioned by row blocks (SELL on the local rows, neighbour ranges exchanged per apply)
opa = ctx.operator_poisson_assembled(nx, ny, nz, cx, cy, cz, capi.FORMAT_SELL)
assert (opa.row_begin, opa.n_rows_local) == (op.row_begin, op.n_rows_local)
ya = gather(opa.apply_numpy(xin[opa.row_begin: opa.row_begin + opa.n_rows_local]))
os.environ["ATK_CG_FUSED"] = "0"
resa = ctx.cg_solve_host(opa, b_loc, iters)
xa = gather(resa["x"])
op27a = ctx.operator_stencil27_assembled(nx, ny, nz, coef, capi.FORMAT_SELL)
y27a = gather(op27a.apply_numpy(xin[op27a.row_begin: op27a.row_begin + op27a.n_rows_local]))
# a general banded matrix through atk_operator_from_csc (every rank passes the whole matrix, keeps its row block)
nb = 40 * world + 7
rngb = np.random.default_rng(8)
Db = np.zeros((nb, nb))
for dgl in range(-9, 6):
    idx = np.arange(max(0, -dgl), min(nb, nb - dgl))
    Db[idx, idx + dgl] = rngb.standard_normal(idx.size) * (rngb.random(idx.size) < 0.7)
rb, cb = np.nonzero(Db)
order = np.lexsort((rb, cb))
rb, cb = rb[order], cb[order]
cpb = np.zeros(nb + 1, np.int32)
np.add.at(cpb, cb + 1, 1)
cpb = np.cumsum(cpb).astype(np.int32)
opb = ctx.operator_from_csc(nb, nb, cpb, rb.astype(np.int32), Db[rb, cb], capi.FORMAT_SELL)
xb = rngb.standard_normal(nb)
yb = gather(opb.apply_numpy(xb[opb.row_begin: opb.row_begin + opb.n_rows_local]))
# a matrix with ARBITRARY sparsity (every rank's rows reach columns of every other rank; an empty row; SPD so that CG
# makes sense): through the whole-matrix entry and through the row-slice entry, peer stores and NCCL
ng = 64 * world + 11
rngg = np.random.default_rng(17)
Dg = np.where(rngg.random((ng, ng)) < 0.06, rngg.standard_normal((ng, ng)), 0.0)
Dg = Dg + Dg.T
Dg[np.arange(ng), np.arange(ng)] = np.abs(Dg).sum(1) + 1.0
Dg[5, :] = 0.0
Dg[:, 5] = 0.0
Dg[5, 5] = 2.0
cg_, rg_ = np.nonzero(Dg.T)  # column-major order = CSC order
cpg = np.zeros(ng + 1, np.int32)
np.add.at(cpg, cg_ + 1, 1)
cpg = np.cumsum(cpg).astype(np.int32)
vg = Dg[rg_, cg_]
xg = rngg.standard_normal(ng)
bg = rngg.standard_normal(ng)
opg = ctx.operator_from_csc(ng, ng, cpg, rg_.astype(np.int32), vg, capi.FORMAT_SELL)
g0, g1 = opg.row_begin, opg.row_begin + opg.n_rows_local
yg = gather(opg.apply_numpy(xg[g0:g1]))
os.environ["ATK_COMM"] = "nccl"
yg_nccl = gather(opg.apply_numpy(xg[g0:g1]))
os.environ.pop("ATK_COMM")
yg2 = gather(opg.apply_numpy(xg[g0:g1]))  # and back to peer stores: the sequence numbers of the two modes are independent
os.environ["ATK_APPLY_SPLIT"] = "1"        # three-piece form: push kernel on the communication stream, wait kernel, pieces
yg3 = gather(opg.apply_numpy(xg[g0:g1]))
y_split = gather(op.apply_numpy(xin[op.row_begin: op.row_begin + op.n_rows_local]))
os.environ.pop("ATK_APPLY_SPLIT")
yg4 = gather(opg.apply_numpy(xg[g0:g1]))
os.environ["ATK_CG_FUSED"] = "0"
resg = ctx.cg_solve_host(opg, bg[g0:g1], 8)
xgs = gather(resg["x"])
rpg, cig, vvg = opg.export_csr()
# row slice: uneven blocks chosen by the caller, only the rank's own entries are passed
cuts = [0] + [int(ng * (q + 1) / world) - (3 if q % 2 == 0 and q + 1 < world else 0) for q in range(world)]
s0, s1 = cuts[rank], cuts[rank + 1]
msk = (rg_ >= s0) & (rg_ < s1)
cps = np.zeros(ng + 1, np.int32)
np.add.at(cps, cg_[msk] + 1, 1)
cps = np.cumsum(cps).astype(np.int32)
ops_ = ctx.operator_from_csc_rows(ng, ng, s0, s1, cps, rg_[msk].astype(np.int32), vg[msk], capi.FORMAT_SELL)
assert (ops_.row_begin, ops_.n_rows_local) == (s0, s1 - s0)
ys = gather(ops_.apply_numpy(xg[s0:s1]))
if rank == 0:
    from oracle.pyoracle import Port

    P = Port()
    b = synthetic.poisson_rhs_r(nx, ny, nz)
    ora = P.cg(None, b, iters, stencil=(nx, ny, nz, cx, cy, cz), trace=True)
    y_ref = P.poisson_apply(nx, ny, nz, cx, cy, cz, xin)
    print(f"world {world} grid {nx}x{ny}x{nz} iters {iters}")
    print("apply with halo exchange bit-exact:", np.array_equal(y_all, y_ref))
    for mode, (us, peer, its) in lat.items():
        print(f"iteration latency floor ({mode}): {us:.1f} us/iteration over {its} iterations (peer={peer})")
    ok &= np.array_equal(y_all, y_ref)
    print("global dot rel err:", abs(d - P.dot(b, b)) / P.dot(b, b))
    ok &= abs(d - P.dot(b, b)) <= 1e-13 * P.dot(b, b)
    A27 = P.op27_csc(nx, ny, nz)
    print("27-point apply with halo exchange bit-exact:", np.array_equal(y27, P.spmv(A27, xin)))
    ok &= np.array_equal(y27, P.spmv(A27, xin))
    o27 = P.gmres(A27, P.spmv(A27, x_ext), np.zeros(nx * ny * nz), 2, 20, 1e-6)
    e27 = np.linalg.norm(x27 - o27["x"]) / np.linalg.norm(o27["x"])
    print("distributed GMRES(20) x2 restarts on the 27-point operator: resid", g27["resid"].tolist(), "oracle", o27["resid"].tolist(),
          f"relerr x {e27:.2e}")
    ok &= e27 < 1e-10 and np.allclose(g27["resid"], o27["resid"], rtol=1e-10)
    e27b = np.linalg.norm(x27b - o27["x"]) / np.linalg.norm(o27["x"])
    print(f"  batched MGS: relerr x {e27b:.2e}  ms {g27b['device_ms']:.2f} vs {g27['device_ms']:.2f} (per-projection path)")
    ok &= e27b < 1e-10 and np.allclose(g27b["resid"], o27["resid"], rtol=1e-10)
    e27p = np.linalg.norm(x27p - o27["x"]) / np.linalg.norm(o27["x"])
    e27u = np.linalg.norm(x27u - o27["x"]) / np.linalg.norm(o27["x"])
    print(f"  per-projection path: fused launch relerr x {e27p:.2e} ({g27p['kernel_launches']} launches), dot + axmy {e27u:.2e} "
          f"({g27u['kernel_launches']} launches)")
    ok &= e27p < 1e-10 and e27u < 1e-10 and np.allclose(g27p["resid"], o27["resid"], rtol=1e-10) and g27p["kernel_launches"] < g27u["kernel_launches"]
    print(f"GMRES(20) x2 at 128x128x{40 * world}, ring step kernel vs shared-memory step kernel vs batched: agree on every rank "
          f"{bool(ring_flags.min() == 1.0)}  ms {ring_ms[0]:.2f} / {ring_ms[1]:.2f} / {ring_ms[2]:.2f}  resid {g_ring['resid'].tolist()}")
    ok &= bool(ring_flags.min() == 1.0)
    for fused, (x, r, p, res) in results.items():
        ex = np.linalg.norm(x - ora["x"]) / np.linalg.norm(ora["x"])
        er = np.linalg.norm(r - ora["r"]) / np.linalg.norm(ora["r"])
        ep = np.linalg.norm(p - ora["p"]) / np.linalg.norm(ora["p"])
        print(f"fused={fused}: iters {res['iters']} (oracle {ora['iters']}) fusedflag {res['fused']} peer {res['peer']} persistent {res['persistent']}  relerr x {ex:.2e} r {er:.2e} p {ep:.2e} "
              f"launches {res['kernel_launches']} ms {res['device_ms']:.2f}")
        ok &= res["iters"] == ora["iters"] and ex < 1e-10 and er < 1e-8 and ep < 1e-8
        ok &= res["persistent"] == (fused == "1")
    print("assembled Poisson (SELL, row blocks) apply bit-exact:", np.array_equal(ya, y_ref),
          " 27-point assembled:", np.array_equal(y27a, P.spmv(A27, xin)))
    ok &= np.array_equal(ya, y_ref) and np.array_equal(y27a, P.spmv(A27, xin))
    exa = np.linalg.norm(xa - ora["x"]) / np.linalg.norm(ora["x"])
    print(f"CG on the distributed assembled operator: iters {resa['iters']} (oracle {ora['iters']}) relerr x {exa:.2e}")
    ok &= resa["iters"] == ora["iters"] and exa < 1e-10
    Ab = P.csc_from_triplets(nb, nb, rb, cb, Db[rb, cb])
    print("banded matrix via from_csc, row blocks, apply bit-exact:", np.array_equal(yb, P.spmv(Ab, xb)))
    ok &= np.array_equal(yb, P.spmv(Ab, xb))
    Ag = P.csc_from_triplets(ng, ng, rg_, cg_, vg)
    yg_ref = P.spmv(Ag, xg)
    print("general sparsity, whole-matrix entry, apply bit-exact (peer stores / NCCL / peer again):", np.array_equal(yg, yg_ref),
          np.array_equal(yg_nccl, yg_ref), np.array_equal(yg2, yg_ref), " row-slice entry:", np.array_equal(ys, yg_ref))
    ok &= np.array_equal(yg, yg_ref) and np.array_equal(yg_nccl, yg_ref) and np.array_equal(yg2, yg_ref) and np.array_equal(ys, yg_ref)
    print("three-piece apply (ATK_APPLY_SPLIT=1) bit-exact: general", np.array_equal(yg3, yg_ref), "stencil", np.array_equal(y_split, y_ref),
          " single launch again:", np.array_equal(yg4, yg_ref))
    ok &= np.array_equal(yg3, yg_ref) and np.array_equal(y_split, y_ref) and np.array_equal(yg4, yg_ref)
    orag = P.cg(Ag, bg, 8)
    exg = np.linalg.norm(xgs - orag["x"]) / np.linalg.norm(orag["x"])
    print(f"CG on the general distributed operator: iters {resg['iters']} (oracle {orag['iters']}) relerr x {exg:.2e}")
    ok &= resg["iters"] == orag["iters"] and exg < 1e-10
    # exported CSR of rank 0's block: global column numbers again
    want_rows = Dg[g0:g1]
    got = np.zeros_like(want_rows)
    for i in range(g1 - g0):
        got[i, cig[rpg[i]:rpg[i + 1]]] = vvg[rpg[i]:rpg[i + 1]]
    print("export_csr of a distributed block returns global columns:", np.array_equal(got, want_rows))
    ok &= np.array_equal(got, want_rows)
    ora2 = P.cg(None, b, 7, state=(ora["x"], ora["r"], ora["p"]), stencil=(nx, ny, nz, cx, cy, cz))
    ec = np.linalg.norm(x_cont - ora2["x"]) / np.linalg.norm(ora2["x"])
    print(f"continued solve (+7 iterations): iters {res2['iters']} peer {res2['peer']} relerr x {ec:.2e}")
    ok &= res2["iters"] == 7 and ec < 1e-10
    print("DIST_CHECK", "PASS" if ok else "FAIL")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
