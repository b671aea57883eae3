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
def _rowblock_worker(rank, world, port, n, seed, out_dir):
    """Host-side model of the row-block protocol of distributed assembled operators (atk_spmv.cu): slice the rows, find
    how far the columns reach into the adjacent blocks, renumber into [received-lo | own | received-hi], exchange the two
    neighbour ranges, multiply.  Sends / receives mirror exchange_ranges()."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import scipy.sparse as sp

    from algotoolkit_b200 import capi

    rng = np.random.default_rng(seed)  # the same matrix and vector on every rank
    D = np.zeros((n, n))
    for dgl in range(-6, 5):
        idx = np.arange(max(0, -dgl), min(n, n - dgl))
        D[idx, idx + dgl] = rng.standard_normal(idx.size) * (rng.random(idx.size) < 0.8)
    A = sp.csr_matrix(D)
    x = rng.standard_normal(n)
    r0, r1 = capi.slab_partition(n, rank, world)
    n_loc = r1 - r0
    loc = A[r0:r1]
    cols = loc.indices
    h_lo = max(0, r0 - int(cols.min())) if cols.size else 0
    h_hi = max(0, int(cols.max()) - (r1 - 1)) if cols.size else 0
    need = torch.tensor([h_lo, h_hi, n_loc], dtype=torch.int64)
    allneed = [torch.zeros(3, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allneed, need)
    allneed = [t.tolist() for t in allneed]
    for q in range(world):  # nobody reaches past an adjacent block
        assert (allneed[q][0] == 0 if q == 0 else allneed[q][0] <= allneed[q - 1][2])
        assert (allneed[q][1] == 0 if q == world - 1 else allneed[q][1] <= allneed[q + 1][2])
    send_prev = allneed[rank - 1][1] if rank > 0 else 0
    send_next = allneed[rank + 1][0] if rank + 1 < world else 0
    x_loc = x[r0:r1].copy()
    halo_lo, halo_hi = torch.zeros(h_lo, dtype=torch.float64), torch.zeros(h_hi, dtype=torch.float64)
    reqs = []
    if rank > 0:
        if send_prev:
            reqs.append(dist.isend(torch.from_numpy(x_loc[:send_prev].copy()), rank - 1))
        if h_lo:
            reqs.append(dist.irecv(halo_lo, rank - 1))
    if rank + 1 < world:
        if send_next:
            reqs.append(dist.isend(torch.from_numpy(x_loc[n_loc - send_next:].copy()), rank + 1))
        if h_hi:
            reqs.append(dist.irecv(halo_hi, rank + 1))
    for q in reqs:
        q.wait()
    x_ext = np.concatenate([halo_lo.numpy(), x_loc, halo_hi.numpy()])
    shift = r0 - h_lo
    loc_ext = sp.csr_matrix((loc.data, loc.indices - shift, loc.indptr), shape=(n_loc, h_lo + n_loc + h_hi))
    y = np.zeros(n_loc)
    for i in range(n_loc):  # ascending-column sums from 0.0, as the SELL kernel adds them
        acc = 0.0
        for k in range(loc_ext.indptr[i], loc_ext.indptr[i + 1]):
            acc = acc + loc_ext.data[k] * x_ext[loc_ext.indices[k]]
        y[i] = acc
    np.save(os.path.join(out_dir, f"y_{rank}.npy"), y)
    np.save(os.path.join(out_dir, f"rb_{rank}.npy"), np.array([r0, r1, h_lo, h_hi, send_prev, send_next]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 41), (3, 64)])
def test_rowblock_spmv_protocol_matches_global_product(tmp_path, world, n):
    import scipy.sparse as sp

    seed = 31
    mp.spawn(_rowblock_worker, args=(world, _free_port(), n, seed, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(seed)
    D = np.zeros((n, n))
    for dgl in range(-6, 5):
        idx = np.arange(max(0, -dgl), min(n, n - dgl))
        D[idx, idx + dgl] = rng.standard_normal(idx.size) * (rng.random(idx.size) < 0.8)
    A = sp.csr_matrix(D)
    x = rng.standard_normal(n)
    want = np.zeros(n)
    for i in range(n):
        acc = 0.0
        for k in range(A.indptr[i], A.indptr[i + 1]):
            acc = acc + A.data[k] * x[A.indices[k]]
        want[i] = acc
    y = np.concatenate([np.load(tmp_path / f"y_{r}.npy") for r in range(world)])
    assert np.array_equal(y, want)  # same per-row sequence of operations -> same bits
    meta = [np.load(tmp_path / f"rb_{r}.npy") for r in range(world)]
    assert meta[0][0] == 0 and meta[-1][1] == n and meta[0][2] == 0 and meta[-1][3] == 0
    for r in range(world - 1):
        assert meta[r][1] == meta[r + 1][0]
        assert meta[r][5] == meta[r + 1][2] and meta[r + 1][4] == meta[r][3]  # what one sends is what the other needs
