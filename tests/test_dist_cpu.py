"""world_size-2 (and 3) gloo runs on CPU of the HOST-SIDE logic of the distributed path: the z-slab partition the
library uses (atk_slab_partition), the per-rank RHS slabs bench.py generates, the halo-plane protocol (first plane
down, last plane up) and the rank-ordered reduction of per-rank partial sums.  The arithmetic is a numpy model of
what the CUDA kernels do per slab; the result is compared with the single-process oracle."""
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
