"""world_size-2 (and 3) gloo runs on CPU of the HOST-SIDE logic of the distributed path: the z-slab partition the
library uses (atk_slab_partition), the per-rank RHS slabs bench.py generates, the halo-plane protocol (first plane
down, last plane up), the rank-ordered reduction of per-rank partial sums, and the row-block protocol of distributed
assembled operators (halo widths, renumbering, neighbour-range exchange).  The arithmetic is a numpy model of what the
CUDA kernels do per slab / row block; the result is compared with the single-process oracle / the global product."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _slab_apply(x_ext, nx, ny, nz, k0, k1, cx, cy, cz):
    """Poisson operator on planes [k0,k1) given x on planes [k0-1,k1+1) (ghosts included), same tap order as the kernel."""
    cc = -2.0 * (cx + cy + cz)
    X = x_ext.reshape(k1 - k0 + 2, ny, nx)
    c = X[1:-1]
    y = c.copy()  # shell rows: identity
    kg = np.arange(k0, k1)
    ik = np.nonzero((kg > 0) & (kg < nz - 1))[0]
    if ik.size and ny > 2 and nx > 2:
        a, b_ = ik[0], ik[-1] + 1
        I = (slice(a, b_), slice(1, -1), slice(1, -1))
        acc = np.zeros_like(c[I])
        acc = acc + cz * X[a:b_, 1:-1, 1:-1]
        acc = acc + cy * c[a:b_, :-2, 1:-1]
        acc = acc + cx * c[a:b_, 1:-1, :-2]
        acc = acc + cc * c[I]
        acc = acc + cx * c[a:b_, 1:-1, 2:]
        acc = acc + cy * c[a:b_, 2:, 1:-1]
        acc = acc + cz * X[a + 2:b_ + 2, 1:-1, 1:-1]
        y[I] = acc
    return y.reshape(-1)


def _worker(rank, world, port, dims, iters, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from algotoolkit_b200 import capi, synthetic

    nx, ny, nz = dims
    k0, k1 = capi.slab_partition(nz, rank, world)
    plane = nx * ny
    cx, cy, cz = synthetic.poisson_coeffs(nx, ny, nz)
    b = synthetic.poisson_rhs_r(nx, ny, nz, k0, k1)
    n = b.size

    def exchange(v):  # first plane down, last plane up (what exchange_halo() does with ncclSend/ncclRecv)
        lo, hi = np.zeros(plane), np.zeros(plane)
        reqs = []
        if rank > 0:
            reqs.append(dist.isend(torch.from_numpy(v[:plane].copy()), rank - 1))
            tlo = torch.zeros(plane, dtype=torch.float64)
            reqs.append(dist.irecv(tlo, rank - 1))
        if rank < world - 1:
            reqs.append(dist.isend(torch.from_numpy(v[-plane:].copy()), rank + 1))
            thi = torch.zeros(plane, dtype=torch.float64)
            reqs.append(dist.irecv(thi, rank + 1))
        for q in reqs:
            q.wait()
        if rank > 0:
            lo = tlo.numpy()
        if rank < world - 1:
            hi = thi.numpy()
        return np.concatenate([lo, v, hi])

    def gdot(u, v):  # all-gather of per-rank partials, summed in rank order: identical bits on every rank
        parts = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(parts, torch.tensor([float(np.dot(u, v))], dtype=torch.float64))
        s = 0.0
        for t in parts:
            s = s + float(t.item())
        return s

    x, r, p = np.zeros(n), b.copy(), b.copy()
    rs_old = gdot(r, r)
    it = 0
    for it in range(iters):
        Ap = _slab_apply(exchange(p), nx, ny, nz, k0, k1, cx, cy, cz)
        pAp = gdot(p, Ap)
        if abs(pAp) < 1e-12:
            break
        alpha = rs_old / pAp
        x = x + p * alpha
        r = r - Ap * alpha
        rs_new = gdot(r, r)
        p = r + p * (rs_new / rs_old)
        rs_old = rs_new
    else:
        it = iters
    np.save(os.path.join(out_dir, f"x_{rank}.npy"), x)
    np.save(os.path.join(out_dir, f"meta_{rank}.npy"), np.array([k0, k1, it, rs_old]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,dims", [(2, (10, 9, 11)), (3, (8, 8, 10))])
def test_slab_cg_matches_single_process_oracle(tmp_path, port, world, dims):
    iters = 25
    mp.spawn(_worker, args=(world, _free_port(), dims, iters, str(tmp_path)), nprocs=world, join=True)
    nx, ny, nz = dims
    metas = [np.load(tmp_path / f"meta_{r}.npy") for r in range(world)]
    # the slabs tile [0, nz) without gaps or overlap
    assert metas[0][0] == 0 and metas[-1][1] == nz
    assert all(metas[r][1] == metas[r + 1][0] for r in range(world - 1))
    sizes = [int(m[1] - m[0]) for m in metas]
    assert max(sizes) - min(sizes) <= 1
    x = np.concatenate([np.load(tmp_path / f"x_{r}.npy") for r in range(world)])
    cx, cy, cz, _ = port.poisson_coeffs(nx, ny, nz)
    b = port.poisson_rhs(nx, ny, nz, rhs_kind=1, seed=12345)
    ora = port.cg(None, b, iters, stencil=(nx, ny, nz, cx, cy, cz))
    assert all(int(m[2]) == ora["iters"] for m in metas)
    assert np.linalg.norm(x - ora["x"]) <= 1e-12 * np.linalg.norm(ora["x"])
    assert len({float(m[3]) for m in metas}) == 1  # every rank derived the same r.r, bit for bit


def test_slab_partition_properties():
    from algotoolkit_b200 import capi

    for nz in (8, 9, 512, 515):
        for world in (1, 2, 3, 4, 8):
            ranges = [capi.slab_partition(nz, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == nz
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1 and sorted(sizes, reverse=True) == sizes
    with pytest.raises(ValueError):
        capi.slab_partition(2, 0, 4)


# ------------------------------------------------------------------------------------------------ row-block operators
def _rowblock_matrix(n, seed, banded):
    rng = np.random.default_rng(seed)  # the same matrix and vector on every rank
    D = np.zeros((n, n))
    if banded:
        for dgl in range(-6, 5):
            idx = np.arange(max(0, -dgl), min(n, n - dgl))
            D[idx, idx + dgl] = rng.standard_normal(idx.size) * (rng.random(idx.size) < 0.8)
    else:  # arbitrary sparsity: every rank's rows reach columns of every other rank
        D = np.where(rng.random((n, n)) < 0.12, rng.standard_normal((n, n)), 0.0)
        D[np.arange(n), np.arange(n)] = 3.0
        D[n // 3] = 0.0  # an empty row
    return D, rng.standard_normal(n)


def _rowblock_worker(rank, world, port, n, seed, banded, out_dir):
    """Host-side model of the general row-block protocol of distributed assembled operators (atk_spmv.cu, PeerGather in
    atk_context.cu): slice the rows; the external columns in ascending order; local numbering [own block | external];
    all-gather of the need matrix; the requesters' column lists travel to the owners; per apply every owner gathers
    what each requester needs and sends it; multiply in the stored (ascending global column) order."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import scipy.sparse as sp

    from algotoolkit_b200 import capi

    D, x = _rowblock_matrix(n, seed, banded)
    A = sp.csr_matrix(D)
    A.sort_indices()
    blocks = [capi.slab_partition(n, q, world) for q in range(world)]
    r0, r1 = blocks[rank]
    n_loc = r1 - r0
    loc = A[r0:r1]
    cols = loc.indices
    ext = np.unique(cols[(cols < r0) | (cols >= r1)])                      # ascending external columns
    recv_count = [int(((ext >= a) & (ext < b)).sum()) for a, b in blocks]  # contiguous runs of `ext`, by owner
    recv_off = np.concatenate([[0], np.cumsum(recv_count)])
    assert recv_count[rank] == 0
    rows = [torch.zeros(world, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(rows, torch.tensor(recv_count, dtype=torch.int64))
    need = np.array([t.tolist() for t in rows])                            # need[p][q]: p receives from q
    send_count = need[:, rank]
    # the column lists go to their owners (ncclSend / ncclRecv of ints in the library)
    reqs, send_idx = [], {}
    for q in range(world):
        if q == rank:
            continue
        if recv_count[q]:
            reqs.append(dist.isend(torch.from_numpy(ext[recv_off[q]:recv_off[q + 1]].astype(np.int64)), q))
        if send_count[q]:
            send_idx[q] = torch.zeros(int(send_count[q]), dtype=torch.int64)
            reqs.append(dist.irecv(send_idx[q], q))
    for r in reqs:
        r.wait()
    # ---- one apply: owners gather and send, requesters receive into [external] at the owner's offset
    x_loc = x[r0:r1].copy()
    x_ext = torch.zeros(ext.size, dtype=torch.float64)
    reqs = []
    for q in range(world):
        if q == rank:
            continue
        if send_count[q]:
            reqs.append(dist.isend(torch.from_numpy(x_loc[send_idx[q].numpy() - r0].copy()), q))
        if recv_count[q]:
            reqs.append(dist.irecv(x_ext[recv_off[q]:recv_off[q + 1]], q))
    for r in reqs:
        r.wait()
    x_all = np.concatenate([x_loc, x_ext.numpy()])
    local_col = np.where((cols >= r0) & (cols < r1), cols - r0, n_loc + np.searchsorted(ext, cols))
    y = np.zeros(n_loc)
    for i in range(n_loc):  # entries stay in ascending GLOBAL column order, whatever their local number
        acc = 0.0
        for k in range(loc.indptr[i], loc.indptr[i + 1]):
            acc = acc + loc.data[k] * x_all[local_col[k]]
        y[i] = acc
    np.save(os.path.join(out_dir, f"y_{rank}.npy"), y)
    np.save(os.path.join(out_dir, f"need_{rank}.npy"), need)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n,banded", [(2, 41, True), (3, 64, True), (2, 37, False), (3, 50, False)])
def test_rowblock_spmv_protocol_matches_global_product(tmp_path, world, n, banded):
    import scipy.sparse as sp

    seed = 31
    mp.spawn(_rowblock_worker, args=(world, _free_port(), n, seed, banded, str(tmp_path)), nprocs=world, join=True)
    D, x = _rowblock_matrix(n, seed, banded)
    A = sp.csr_matrix(D)
    A.sort_indices()
    want = np.zeros(n)
    for i in range(n):
        acc = 0.0
        for k in range(A.indptr[i], A.indptr[i + 1]):
            acc = acc + A.data[k] * x[A.indices[k]]
        want[i] = acc
    y = np.concatenate([np.load(tmp_path / f"y_{r}.npy") for r in range(world)])
    assert np.array_equal(y, want)  # same per-row sequence of operations -> same bits
    needs = [np.load(tmp_path / f"need_{r}.npy") for r in range(world)]
    assert all(np.array_equal(needs[0], m) for m in needs)  # every rank derived the same need matrix
    if not banded:
        assert (needs[0] + np.eye(world, dtype=int) > 0).all()  # truly all-to-all
