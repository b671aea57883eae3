"""Parity of the CUDA path (through the C ABI) against the oracle.  Run on the B200 box: pytest -m gpu.

Bar (BASELINE.json north_star): integer/index work bit-exact; SpMV and the elementwise vector ops bit-exact (they
have no reduction-order freedom); CG/GMRES iteration counts within +-2 and solution / residual within 1e-10
relative with tree reductions, and bit-exact in ATK_REDUCE_SEQUENTIAL mode.
"""
import hashlib

import numpy as np
import pytest

from algotoolkit_b200 import capi
from tests.conftest import mm_path

pytestmark = pytest.mark.gpu

REL = 1e-10  # north_star tolerance for solutions / residuals (relative, fp64)


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.float64).tobytes()).hexdigest()


def relerr(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


# ------------------------------------------------------------------------------------------------ VectorObj
@pytest.mark.parametrize("n", [0, 1, 2, 3, 31, 32, 33, 1000, 1074, 4097, 1 << 20, (1 << 20) + 1])
def test_vector_ops_bit_exact(ctx, port, n):
    rng = np.random.default_rng(n + 1)
    a, b = rng.standard_normal(n), rng.standard_normal(n)
    da, db = ctx.vector(a), ctx.vector(b)
    s = 0.37251
    assert np.array_equal(ctx.scale(da, s).numpy(), port.vec_scale(a, s))
    assert np.array_equal(ctx.div(da, s).numpy(), port.vec_div(a, s))
    assert np.array_equal(ctx.add(da, db).numpy(), port.vec_add(a, b))
    assert np.array_equal(ctx.sub(da, db).numpy(), port.vec_sub(a, b))
    assert np.array_equal(ctx.axpy(da, db, s).numpy(), port.vec_add(a, port.vec_scale(b, s)))
    assert np.array_equal(ctx.axmy(da, db, s).numpy(), port.vec_sub(a, port.vec_scale(b, s)))


def test_vector_div_by_zero_is_runtime_error(ctx):
    with pytest.raises(RuntimeError, match="Division by zero"):
        ctx.div(ctx.vector(np.ones(4)), 0.0)


@pytest.mark.parametrize("n", [0, 1, 5, 1024, 1025, 1074, 100003, 1 << 21])
def test_dot_and_norm(ctx, port, n):
    rng = np.random.default_rng(n + 7)
    a, b = rng.standard_normal(n), rng.standard_normal(n)
    da, db = ctx.vector(a), ctx.vector(b)
    # sequential mode: bit-identical to std::inner_product
    ctx.set_reduction(capi.REDUCE_SEQUENTIAL)
    try:
        if n <= 100003:
            assert ctx.dot(da, db) == port.dot(a, b)
            assert ctx.nrm2(da) == port.nrm2(a)
    finally:
        ctx.set_reduction(capi.REDUCE_TREE)
    # tree mode: deterministic, within a few ulp * sqrt(n) of the sequential sum
    d1, d2 = ctx.dot(da, db), ctx.dot(da, db)
    assert d1 == d2
    scale = float(np.abs(a * b).sum()) + 1e-300
    assert abs(d1 - port.dot(a, b)) <= 1e-13 * scale
    assert abs(ctx.nrm2(da) - port.nrm2(a)) <= 1e-13 * (port.nrm2(a) + 1e-300)


def test_dot_dimension_mismatch(ctx):
    with pytest.raises(ValueError):
        ctx.dot(ctx.vector(np.ones(3)), ctx.vector(np.ones(4)))


# ------------------------------------------------------------------------------------------------ SpMV + conversion
def _check_csr_against_csc(op, A):
    rp, ci, vv = op.export_csr()
    S = A.to_scipy().tocsr()
    S.sort_indices()
    assert np.array_equal(rp, S.indptr)
    assert np.array_equal(ci, S.indices)
    assert np.array_equal(vv, S.data)


@pytest.mark.parametrize("fmt", [capi.FORMAT_SELL, capi.FORMAT_CSR_VECTOR])
def test_spmv_kat_3x3(ctx, golden, port, fmt):
    k = golden["kat"]["spmv_3x3"]  # reference tests/SparseMatrixCSCTest.cpp:47-64
    A = port.csc_from_triplets(3, 3, k["rows"], k["cols"], k["vals"])
    op = ctx.operator_from_csc(3, 3, A.col_ptr, A.row_idx, A.vals, fmt)
    assert op.apply_numpy(k["x"]).tolist() == k["expect"] == k["y"]


def test_spmv_dimension_mismatch(ctx, port):
    A = port.csc_from_triplets(3, 3, [0], [0], [1.0])
    op = ctx.operator_from_csc(3, 3, A.col_ptr, A.row_idx, A.vals)
    with pytest.raises(ValueError, match="Dimension mismatch"):
        op.apply_numpy(np.ones(4))


def test_csc_out_of_range_row(ctx):
    with pytest.raises(IndexError):
        ctx.operator_from_csc(2, 2, np.array([0, 1, 2], np.int32), np.array([0, 5], np.int32), np.array([1.0, 2.0]))


@pytest.mark.parametrize("col_ptr", [[1, 1, 2], [0, 2, 1], [0, 1, 3], [0, 1, 1], [0, -1, 2]])
def test_csc_malformed_col_ptr_is_rejected_before_any_scatter(ctx, col_ptr):
    """first entry 0, non-decreasing, last entry nnz - anything else must not reach the scatter kernels"""
    with pytest.raises(ValueError):
        ctx.operator_from_csc(2, 2, np.array(col_ptr, np.int32), np.array([0, 1], np.int32), np.array([1.0, 2.0]))


@pytest.mark.parametrize("name", ["bcsstk01", "bcsstk08"])
def test_spmv_matrix_market_bit_exact(ctx, port, golden, name):
    A = port.read_matrix_market(mm_path(name))
    g = golden["gmres_mm"][name]
    assert (A.n_rows, A.nnz) == (g["n"], g["nnz"])
    x = port.u01_array(777, A.n_cols)
    y_ref = port.spmv(A, x)
    assert digest(y_ref) == g["b_sha256"]  # the oracle itself is pinned to the reference
    op = ctx.operator_from_csc(A.n_rows, A.n_cols, A.col_ptr, A.row_idx, A.vals, capi.FORMAT_SELL)
    _check_csr_against_csc(op, A)
    assert np.array_equal(op.apply_numpy(x), y_ref)  # SELL: bit-exact
    # x with exact zeros (the reference skips those columns)
    xz = x.copy()
    xz[::3] = 0.0
    assert np.array_equal(op.apply_numpy(xz), port.spmv(A, xz))
    opv = ctx.operator_from_csc(A.n_rows, A.n_cols, A.col_ptr, A.row_idx, A.vals, capi.FORMAT_CSR_VECTOR)
    yv = opv.apply_numpy(x)
    assert np.array_equal(yv, opv.apply_numpy(x))  # deterministic
    assert relerr(yv, y_ref) < 1e-14


def test_spmv_random_rectangular_and_empty_rows(ctx, port):
    rng = np.random.default_rng(5)
    n_rows, n_cols = 257, 131
    dense = rng.standard_normal((n_rows, n_cols)) * (rng.random((n_rows, n_cols)) < 0.08)
    dense[10, :] = rng.standard_normal(n_cols)  # a long row (> 32 entries)
    dense[20:40, :] = 0.0                       # empty rows
    r, c = np.nonzero(dense)
    A = port.csc_from_triplets(n_rows, n_cols, r, c, dense[r, c])
    x = rng.standard_normal(n_cols)
    for fmt in (capi.FORMAT_SELL, capi.FORMAT_CSR_VECTOR):
        op = ctx.operator_from_csc(n_rows, n_cols, A.col_ptr, A.row_idx, A.vals, fmt)
        _check_csr_against_csc(op, A)
        y = op.apply_numpy(x)
        if fmt == capi.FORMAT_SELL:
            assert np.array_equal(y, port.spmv(A, x))
        else:
            assert relerr(y, port.spmv(A, x)) < 1e-14
    # empty matrix
    op = ctx.operator_from_csc(4, 4, np.zeros(5, np.int32), np.zeros(0, np.int32), np.zeros(0))
    assert np.array_equal(op.apply_numpy(np.ones(4)), np.zeros(4))


@pytest.mark.parametrize("dims", [(5, 6, 7), (8, 8, 8), (32, 32, 32), (33, 17, 9), (64, 48, 40), (2, 2, 2), (3, 3, 3), (1, 4, 4)])
def test_poisson_operators_bit_exact(ctx, port, dims):
    nx, ny, nz = dims
    lx, ly, lz = 1.0, 2.0, 0.5
    if nx > 1:
        cx, cy, cz, cc = port.poisson_coeffs(nx, ny, nz, lx, ly, lz)
        A = port.poisson_csc(nx, ny, nz, lx, ly, lz)
    else:
        cx, cy, cz = 3.0, 5.0, 7.0
        A = None
    rng = np.random.default_rng(nx * 1000 + ny)
    x = rng.standard_normal(nx * ny * nz)
    x[::11] = 0.0
    y_ref = port.poisson_apply(nx, ny, nz, cx, cy, cz, x)
    if A is not None:
        assert np.array_equal(y_ref, port.spmv(A, x))
    st = ctx.operator_stencil7(nx, ny, nz, cx, cy, cz)
    assert np.array_equal(st.apply_numpy(x), y_ref)  # matrix-free stencil: bit-exact
    if A is not None:
        asm = ctx.operator_poisson_assembled(nx, ny, nz, cx, cy, cz, capi.FORMAT_SELL)
        assert asm.nnz == A.nnz == st.nnz
        _check_csr_against_csc(asm, A)  # device-side assembly + CSC->CSR conversion
        assert np.array_equal(asm.apply_numpy(x), y_ref)


def test_op27_assembled_bit_exact(ctx, port, golden):
    n = 8
    A = port.op27_csc(n, n, n)
    assert digest(A.vals) == golden["gmres_op27"]["8_K30"]["vals_sha"]
    coef = np.array([40.0 if (di == 0 and dj == 0 and dk == 0) else -1.0 + 0.3 * (di + 0.5 * dj + 0.25 * dk)
                     for dk in (-1, 0, 1) for dj in (-1, 0, 1) for di in (-1, 0, 1)])
    op = ctx.operator_stencil27_assembled(n, n, n, coef)
    _check_csr_against_csc(op, A)
    x = port.u01_array(777, n ** 3)
    assert np.array_equal(op.apply_numpy(x), port.spmv(A, x))


@pytest.mark.parametrize("dims", [(8, 8, 8), (7, 9, 5), (1, 3, 4), (33, 6, 2), (64, 12, 9), (66, 5, 70), (130, 7, 3), (2, 2, 2),
                                  (4, 1, 1), (96, 40, 33)])
def test_op27_matrix_free_bit_exact(ctx, port, dims):
    from algotoolkit_b200 import synthetic

    nx, ny, nz = dims
    coef = synthetic.stencil27_coefficients()
    A = port.op27_csc(nx, ny, nz)
    op = ctx.operator_stencil27(nx, ny, nz, coef)
    assert op.nnz == A.nnz
    x = np.random.default_rng(nx).standard_normal(nx * ny * nz)
    x[::5] = 0.0
    assert np.array_equal(op.apply_numpy(x), port.spmv(A, x))
    # sparse coefficient table (the 7-point open-boundary Laplacian used by Poisson3DSolver without Dirichlet sides)
    c7 = np.zeros(27)
    c7[13], c7[12], c7[14], c7[10], c7[16], c7[4], c7[22] = -6.0, 1.0, 1.0, 2.0, 2.0, 3.0, 3.0
    asm = ctx.operator_stencil27_assembled(nx, ny, nz, c7)
    mf = ctx.operator_stencil27(nx, ny, nz, c7)
    assert asm.nnz == mf.nnz
    assert np.array_equal(asm.apply_numpy(x), mf.apply_numpy(x))


# ------------------------------------------------------------------------------------------------ ConjugateGrad
@pytest.fixture(params=["onchip", "persistent", "graph"])
def cg_loop(request, monkeypatch):
    """The CG loop of the matrix-free operator runs three ways: one cooperative launch with the vectors on chip (small
    grids, default there), one persistent cooperative launch that streams the vectors (what large grids and the z-slabs
    of a distributed run use), or the fused two-kernel iteration replayed from a CUDA graph (fallback).  Tests taking
    this fixture cover all three."""
    monkeypatch.setenv("ATK_CG_ONCHIP", "1" if request.param == "onchip" else "0")
    monkeypatch.setenv("ATK_CG_PERSISTENT", "0" if request.param == "graph" else "1")
    return request.param


def test_cg_kats(ctx, port, golden):
    k = golden["kat"]["cg_tridiag3"]  # tests/ConjugateGradient_test.cpp:67-76
    A = port.csc_from_triplets(3, 3, k["rows"], k["cols"], k["vals"])
    op = ctx.operator_from_csc(3, 3, A.col_ptr, A.row_idx, A.vals)
    res = ctx.cg_solve_host(op, k["b"], k["max_iter"])
    assert np.allclose(res["x"], k["expect"], atol=k["tol"])
    assert relerr(res["x"], k["x"]) < REL and abs(res["iters"] - k["iters"]) <= 2
    one = ctx.cg_solve_host(op, k["b"], 1)  # :92-100 single iteration
    k1 = golden["kat"]["cg_tridiag3_1iter"]
    assert one["iters"] == 1 and relerr(one["x"], k1["x"]) < REL and relerr(one["p"], k1["p"]) < REL
    zero = ctx.cg_solve_host(op, np.zeros(3), 1000)  # :79-89 zero rhs
    assert zero["zero_rhs"] and zero["iters"] == 0 and np.array_equal(zero["x"], np.zeros(3))
    kd = golden["kat"]["cg_diag4"]  # tests/debuglogger.cpp
    Ad = port.csc_from_triplets(4, 4, [0, 1, 2, 3], [0, 1, 2, 3], [4.0, 3.0, 2.0, 1.0])
    opd = ctx.operator_from_csc(4, 4, Ad.col_ptr, Ad.row_idx, Ad.vals)
    rd = ctx.cg_solve_host(opd, kd["b"], kd["max_iter"])
    assert np.allclose(rd["x"], kd["expect"], atol=kd["tol"]) and relerr(rd["x"], kd["x"]) < REL
    k2 = golden["kat"]["cg_2x2_notebook"]  # code/test.ipynb cell 2
    A2 = port.csc_from_triplets(2, 2, [0, 0, 1, 1], [0, 1, 0, 1], [4.0, 1.0, 1.0, 3.0])
    op2 = ctx.operator_from_csc(2, 2, A2.col_ptr, A2.row_idx, A2.vals)
    r2 = ctx.cg_solve_host(op2, [1.0, 2.0], 100)
    assert relerr(r2["x"], k2["expect"]) < REL


@pytest.mark.parametrize("n", [8, 16, 32])
def test_poisson_readme_case(ctx, port, golden, n, cg_loop):
    """README / code/PDE.ipynb: RHS is an eigenvector of the discrete operator -> exactly 1 CG iteration."""
    g = golden["poisson_readme"][str(n)]
    b = port.poisson_rhs(n, n, n, rhs_kind=0)
    assert digest(b) == g["b_sha256"]
    cx, cy, cz, _ = port.poisson_coeffs(n, n, n)
    op = ctx.operator_stencil7(n, n, n, cx, cy, cz)
    res = ctx.cg_solve_host(op, b, 2000)
    assert res["iters"] == g["iters"] == 1 and res["stopped_on_pAp"]
    grid = np.arange(n) / (n - 1)
    an = np.sin(np.pi * grid)[None, None, :] * np.sin(np.pi * grid)[None, :, None] * np.sin(np.pi * grid)[:, None, None]
    err = np.abs(res["x"].reshape(n, n, n) - an)
    assert abs(err.max() - g["max_err"]) < 1e-12 and abs(np.sqrt((err ** 2).mean()) - g["l2_err"]) < 1e-12
    if n == 32:
        assert f"{err.max():.6e}" == golden["poisson_readme"]["notebook_32"]["max_err"]
        assert f"{np.sqrt((err ** 2).mean()):.6e}" == golden["poisson_readme"]["notebook_32"]["l2_err"]


@pytest.mark.parametrize("n,kind", [(12, "stencil"), (32, "stencil"), (32, "sell"), (32, "csrv"), (64, "stencil")])
def test_poisson_cg_to_exit(ctx, port, golden, n, kind, cg_loop):
    """RHS-R run until |p.Ap| < 1e-12 fires: iteration count +-2, x / r within 1e-10 relative of the reference."""
    g = golden["poisson_rhsr"][str(n)]
    b = port.poisson_rhs(n, n, n, rhs_kind=1, seed=12345)
    assert digest(b) == g["b_sha256"]
    cx, cy, cz, _ = port.poisson_coeffs(n, n, n)
    if kind == "stencil":
        op = ctx.operator_stencil7(n, n, n, cx, cy, cz)
    else:
        op = ctx.operator_poisson_assembled(n, n, n, cx, cy, cz, capi.FORMAT_SELL if kind == "sell" else capi.FORMAT_CSR_VECTOR)
    res = ctx.cg_solve_host(op, b, 100000, trace=True)
    assert res["stopped_on_pAp"] and abs(res["iters"] - g["iters"]) <= 2
    ora = port.cg(None, b, 100000, stencil=(n, n, n, cx, cy, cz), trace=True)
    assert ora["iters"] == g["iters"] and digest(ora["x"]) == g["x_sha256"]
    assert relerr(res["x"], ora["x"]) < REL
    assert abs(np.linalg.norm(res["x"]) - g["x_norm"]) < REL * g["x_norm"]
    # true residual of the device solution
    true_r = b - port.poisson_apply(n, n, n, cx, cy, cz, res["x"])
    assert abs(np.linalg.norm(true_r) - np.linalg.norm(b - port.poisson_apply(n, n, n, cx, cy, cz, ora["x"]))) < REL * g["b_norm"]
    # per-iteration scalars track the reference while they are well above rounding level
    m = min(res["iters"], ora["iters"], 60)
    assert np.allclose(res["rs_trace"][:m], ora["rs_trace"][:m], rtol=1e-9)
    assert np.allclose(res["pAp_trace"][:m], ora["pAp_trace"][:m], rtol=1e-9)


@pytest.mark.parametrize("dims", [(13, 11, 9), (6, 30, 7), (34, 5, 12)])
def test_poisson_cg_ragged_grids(ctx, port, dims, cg_loop):
    """Odd / non-tile-multiple grids: the fused step falls back to its L1 variant for odd nx, partial tiles everywhere."""
    nx, ny, nz = dims
    cx, cy, cz, _ = port.poisson_coeffs(nx, ny, nz, 1.0, 2.0, 0.5)
    b = port.poisson_rhs(nx, ny, nz, 1.0, 2.0, 0.5, rhs_kind=1, seed=7)
    ora = port.cg(None, b, 45, stencil=(nx, ny, nz, cx, cy, cz))
    for op in (ctx.operator_stencil7(nx, ny, nz, cx, cy, cz), ctx.operator_poisson_assembled(nx, ny, nz, cx, cy, cz)):
        res = ctx.cg_solve_host(op, b, 45)
        assert res["iters"] == ora["iters"] and relerr(res["x"], ora["x"]) < REL and relerr(res["p"], ora["p"]) < 1e-8


def test_cg_sequential_mode_is_bit_exact(ctx, port, golden):
    n = 12
    g = golden["poisson_rhsr"][str(n)]
    b = port.poisson_rhs(n, n, n, rhs_kind=1, seed=12345)
    cx, cy, cz, _ = port.poisson_coeffs(n, n, n)
    ctx.set_reduction(capi.REDUCE_SEQUENTIAL)
    try:
        for op in (ctx.operator_stencil7(n, n, n, cx, cy, cz), ctx.operator_poisson_assembled(n, n, n, cx, cy, cz)):
            res = ctx.cg_solve_host(op, b, 100000)
            assert res["iters"] == g["iters"]
            assert digest(res["x"]) == g["x_sha256"] and digest(res["r"]) == g["r_sha256"] and digest(res["p"]) == g["p_sha256"]
    finally:
        ctx.set_reduction(capi.REDUCE_TREE)


def test_cg_fixed_work_and_continuation(ctx, port, golden, cg_loop):
    """maxIter iterations exactly (no tolerance test in the reference); a second solve continues from x, r, p."""
    n = 32
    b = port.poisson_rhs(n, n, n, rhs_kind=1, seed=12345)
    cx, cy, cz, _ = port.poisson_coeffs(n, n, n)
    op = ctx.operator_stencil7(n, n, n, cx, cy, cz)
    r50 = ctx.cg_solve_host(op, b, 50)
    assert r50["iters"] == 50 and not r50["stopped_on_pAp"]
    assert abs(np.linalg.norm(r50["x"]) - golden["poisson_rhsr"]["32_fixed50"]["x_norm"]) < REL * golden["poisson_rhsr"]["32_fixed50"]["x_norm"]
    a = ctx.cg_solve_host(op, b, 20)
    bb = ctx.cg_solve_host(op, b, 30, state=(a["x"], a["r"], a["p"]))
    ora = port.cg(None, b, 20, stencil=(n, n, n, cx, cy, cz))
    ora2 = port.cg(None, b, 30, state=(ora["x"], ora["r"], ora["p"]), stencil=(n, n, n, cx, cy, cz))
    assert relerr(bb["x"], ora2["x"]) < REL and relerr(bb["x"], r50["x"]) < REL


@pytest.mark.parametrize("dims,split", [((96, 70, 40), None), ((64, 8, 130), None), ((130, 36, 24), "3"), ((32, 4, 9), None)])
def test_cg_persistent_loop(ctx, port, dims, split, monkeypatch):
    """The one-launch persistent loop on grids with partial tiles, several z-chunks per tile and more items than CTAs:
    it must be the path taken (info.persistent), agree with the oracle per iteration (traces), honour the exit, the
    zero right-hand side and the continuation semantics, and report phase times."""
    monkeypatch.setenv("ATK_CG_ONCHIP", "0")
    monkeypatch.setenv("ATK_CG_PERSISTENT", "1")
    if split:
        monkeypatch.setenv("ATK_PERSIST_SPLIT", split)
    nx, ny, nz = dims
    cx, cy, cz, _ = port.poisson_coeffs(nx, ny, nz, 1.0, 2.0, 0.5)
    b = port.poisson_rhs(nx, ny, nz, 1.0, 2.0, 0.5, rhs_kind=1, seed=99)
    op = ctx.operator_stencil7(nx, ny, nz, cx, cy, cz)
    ora = port.cg(None, b, 37, stencil=(nx, ny, nz, cx, cy, cz), trace=True)
    res = ctx.cg_solve_host(op, b, 37, trace=True)
    assert res["persistent"] and res["fused"] and res["iters"] == ora["iters"] == 37 and res["kernel_launches"] == 1
    assert relerr(res["x"], ora["x"]) < REL and relerr(res["r"], ora["r"]) < 1e-8 and relerr(res["p"], ora["p"]) < 1e-8
    assert np.allclose(res["pAp_trace"], ora["pAp_trace"], rtol=1e-9) and np.allclose(res["rs_trace"], ora["rs_trace"], rtol=1e-9)
    # continuation: 15 + 22 iterations == 37 iterations
    a = ctx.cg_solve_host(op, b, 15)
    c = ctx.cg_solve_host(op, b, 22, state=(a["x"], a["r"], a["p"]))
    assert c["persistent"] and relerr(c["x"], ora["x"]) < REL and relerr(c["p"], ora["p"]) < 1e-8
    # run to the |pAp| < 1e-12 exit
    full = ctx.cg_solve_host(op, b, 100000)
    ofull = port.cg(None, b, 100000, stencil=(nx, ny, nz, cx, cy, cz))
    assert full["stopped_on_pAp"] and abs(full["iters"] - ofull["iters"]) <= 2 and relerr(full["x"], ofull["x"]) < REL
    # zero right-hand side: immediate return, x untouched (ConjugateGradient.hpp:38-42)
    z = ctx.cg_solve_host(op, np.zeros_like(b), 10)
    assert z["zero_rhs"] and z["iters"] == 0 and not z["x"].any()
    # phase stamps (time_kernels): both phases must have been timed
    db = ctx.vector(b)
    x, r, p = ctx.vector(np.zeros_like(b)), ctx.vector(b), ctx.vector(b)
    t = ctx.cg_solve(op, db, x, r, p, 12, time_kernels=True)
    assert t["persistent"] and t["kernel_ms"][0] > 0 and t["kernel_ms"][1] > 0
    assert t["kernel_ms"][0] + t["kernel_ms"][1] <= t["device_ms"] * 1.05
    o12 = port.cg(None, b, 12, stencil=(nx, ny, nz, cx, cy, cz))
    assert relerr(x.numpy(), o12["x"]) < REL
    op.destroy()


def test_cg_resident_device_api(ctx, port, cg_loop):
    n = 16
    b = port.poisson_rhs(n, n, n, rhs_kind=1, seed=12345)
    cx, cy, cz, _ = port.poisson_coeffs(n, n, n)
    op = ctx.operator_stencil7(n, n, n, cx, cy, cz)
    db = ctx.vector(b)
    x, r, p = ctx.vector(np.zeros_like(b)), ctx.vector(b), ctx.vector(b)
    res = ctx.cg_solve(op, db, x, r, p, 40)
    ora = port.cg(None, b, 40, stencil=(n, n, n, cx, cy, cz))
    assert res["iters"] == 40 and res["kernel_launches"] > 0
    assert relerr(x.numpy(), ora["x"]) < REL and relerr(r.numpy(), ora["r"]) < 1e-8


# ------------------------------------------------------------------------------------------------ GMRES
def _gmres_events_to_text(events):
    """The host class prints these exact lines (GMRES.hpp:48,50,83,105,108,114); tests compare with the reference log."""
    out = []
    for ev, v in events:
        if ev == 0:
            out.append("Initial residual norm: %g" % v)
        elif ev == 1:
            out.append("Initial guess is good enough.")
        elif ev == 2:
            out.append("KrylovUpdate : %d" % int(v))
        elif ev == 3:
            out.append("Residual norm after restart: %g " % v)
        elif ev == 4:
            out.append("Converged after restart at iteration %d" % int(v))
        elif ev == 5:
            out.append("Reached maximum iterations without convergence.")
    return out


def test_gmres_kats(ctx, port, golden):
    k = golden["kat"]["cg_tridiag3"]
    A = port.csc_from_triplets(3, 3, k["rows"], k["cols"], k["vals"])
    op = ctx.operator_from_csc(3, 3, A.col_ptr, A.row_idx, A.vals)
    for K in (1, 2, 3):  # tests/GMRES_test.cpp:43-53,82-92
        g = golden["kat"][f"gmres_tridiag3_K{K}"]
        res = ctx.gmres_solve_host(op, [1.0, 5.0, 0.0], np.zeros(3), 10000, K, 1e-12)
        assert res["status"] == g["status"] == 1
        assert np.allclose(res["x"], g["expect"], atol=g["tol"]) and relerr(res["x"], g["x"]) < 1e-9
        assert abs(len(res["resid"]) - g["resid"]["n"]) <= 2
        assert np.linalg.norm(np.array([1.0, 5.0, 0.0]) - port.spmv(A, res["x"])) < 1e-6
    with pytest.raises(ValueError):  # :77-80 wrong-size x
        ctx.gmres_solve_host(op, [1.0, 5.0, 0.0], np.zeros(4), 10, 3, 1e-12)
    with pytest.raises(ValueError):  # Krylov dim > n
        ctx.gmres_solve_host(op, [1.0, 5.0, 0.0], np.zeros(3), 10, 4, 1e-12)
    z = ctx.gmres_solve_host(op, np.zeros(3), np.zeros(3), 10000, 3, 1e-12)  # :55-65 zero rhs
    assert z["status"] == 2 and np.array_equal(z["x"], np.zeros(3))


def test_gmres_2x2_notebook(ctx, port, golden):
    g = golden["kat"]["gmres_2x2_notebook"]  # code/test.ipynb cell 6
    A2 = port.csc_from_triplets(2, 2, [0, 0, 1, 1], [0, 1, 0, 1], [4.0, 1.0, 1.0, 3.0])
    op = ctx.operator_from_csc(2, 2, A2.col_ptr, A2.row_idx, A2.vals)
    ctx.set_reduction(capi.REDUCE_SEQUENTIAL)
    try:
        res = ctx.gmres_solve_host(op, [1.0, 2.0], [0.0, 0.0], 100, 1, 1e-6)
    finally:
        ctx.set_reduction(capi.REDUCE_TREE)
    assert res["x"].tolist() == g["x"] == g["expect"]
    text = _gmres_events_to_text(res["events"])
    assert text[:6] == [s.rstrip("\n") for s in g["printed_log_head"]]
    assert text[-1] == "Converged after restart at iteration %d" % g["expect_converged_at"]


@pytest.mark.parametrize("name,key", [("bcsstk01", "K48_it3"), ("bcsstk08", "K30_it2"), ("bcsstk08", "K100_it1"),
                                      ("bcsstk08", "K300_it6")])
def test_gmres_matrix_market_sequential_bit_exact(ctx, port, golden, name, key):
    """Config 1 (GMRES(300) on bcsstk08 as the reference loads it: lower triangle only).  The trajectory is chaotic
    in the reference itself; with reference-ordered reductions the device reproduces it bit for bit."""
    g = golden["gmres_mm"][name]
    run = g["runs"][key]
    A = port.read_matrix_market(mm_path(name))
    b = port.spmv(A, port.u01_array(777, A.n_cols))
    op = ctx.operator_from_csc(A.n_rows, A.n_cols, A.col_ptr, A.row_idx, A.vals, capi.FORMAT_SELL)
    ctx.set_reduction(capi.REDUCE_SEQUENTIAL)
    try:
        res = ctx.gmres_solve_host(op, b, np.zeros(A.n_rows), run["max_iter"], run["K"], 1e-6)
    finally:
        ctx.set_reduction(capi.REDUCE_TREE)
    assert res["resid"].tolist() == run["resid"]["all"]
    assert digest(res["x"]) == run["x_sha256"] and res["status"] == run["status"]


@pytest.mark.parametrize("K", [5, 10, 20])
def test_gmres_bcsstk08_tree_mode_small_K(ctx, port, golden, K):
    """With tree reductions parity holds where the reference itself is insensitive to summation order (K <= 20)."""
    run = golden["gmres_mm"]["bcsstk08"]["runs"][f"K{K}_it1"]
    A = port.read_matrix_market(mm_path("bcsstk08"))
    b = port.spmv(A, port.u01_array(777, A.n_cols))
    op = ctx.operator_from_csc(A.n_rows, A.n_cols, A.col_ptr, A.row_idx, A.vals)
    res = ctx.gmres_solve_host(op, b, np.zeros(A.n_rows), 1, K, 1e-6)
    assert np.allclose(res["resid"], run["resid"]["all"], rtol=REL)
    assert abs(np.linalg.norm(res["x"]) - run["x_norm"]) <= 1e-9 * run["x_norm"]


@pytest.mark.parametrize("key", ["8_K30", "12_K60"])
def test_gmres_op27_tree_mode(ctx, port, golden, key):
    """Config 5 operator at CPU-feasible size: nonsymmetric, diagonally dominant -> tree reductions keep 1e-10."""
    g = golden["gmres_op27"][key]
    n = g["n"]
    A = port.op27_csc(n, n, n)
    b = port.spmv(A, port.u01_array(777, n ** 3))
    assert digest(b) == g["b_sha256"]
    coef = np.array([40.0 if (di == 0 and dj == 0 and dk == 0) else -1.0 + 0.3 * (di + 0.5 * dj + 0.25 * dk)
                     for dk in (-1, 0, 1) for dj in (-1, 0, 1) for di in (-1, 0, 1)])
    ora = port.gmres(A, b, np.zeros(n ** 3), g["max_iter"], g["K"], 1e-6)
    assert digest(ora["x"]) == g["x_sha256"]
    for op in (ctx.operator_stencil27_assembled(n, n, n, coef), ctx.operator_stencil27(n, n, n, coef)):
        res = ctx.gmres_solve_host(op, b, np.zeros(n ** 3), g["max_iter"], g["K"], 1e-6)
        assert np.allclose(res["resid"], g["resid"]["all"], rtol=REL)
        assert relerr(res["x"], ora["x"]) < REL
        # batched (low-synchronisation) MGS: two passes over V and two reductions per step, same projection
        res = ctx.gmres_solve_host(op, b, np.zeros(n ** 3), g["max_iter"], g["K"], 1e-6, ortho=capi.GMRES_MGS_BATCHED)
        assert np.allclose(res["resid"], g["resid"]["all"], rtol=REL)
        assert relerr(res["x"], ora["x"]) < REL


def test_gmres_batched_mgs_k300(ctx, port):
    """K = 300 (the BASELINE Krylov dimension) on the config-5 operator at 14^3: the batched projection tracks the
    reference's sequential loop to 1e-10 over two restarts (numpy prototype: 2e-14)."""
    from algotoolkit_b200 import synthetic

    n = 14
    A = port.op27_csc(n, n, n)
    b = port.spmv(A, port.u01_array(777, n ** 3))
    ora = port.gmres(A, b, np.zeros(n ** 3), 2, 300, 1e-6)
    op = ctx.operator_stencil27(n, n, n, synthetic.stencil27_coefficients())
    for ortho in (capi.GMRES_MGS, capi.GMRES_MGS_BATCHED):
        res = ctx.gmres_solve_host(op, b, np.zeros(n ** 3), 2, 300, 1e-6, ortho=ortho)
        assert np.allclose(res["resid"], ora["resid"], rtol=REL), ortho
        assert relerr(res["x"], ora["x"]) < REL, ortho


# ------------------------------------------------------------------------------------------------ full-size properties
def test_full_size_properties_512(ctx):
    """BASELINE's headline size (512^3, 134 M rows) is beyond what the CPU oracle can solve in a test, so the device path
    is checked there through size-independent properties: the recurrence residual equals the true residual b - A x, the
    operator is linear, exactly K iterations run, and two runs give identical bits (deterministic reductions)."""
    from algotoolkit_b200 import synthetic

    g = 512
    n = g ** 3
    cx, cy, cz = synthetic.poisson_coeffs(g, g, g)
    op = ctx.operator_stencil7(g, g, g, cx, cy, cz)
    b = synthetic.poisson_rhs_r(g, g, g)
    db = ctx.vector(b)
    x, r, p = ctx.empty(n), ctx.empty(n), ctx.empty(n)
    runs = []
    for _ in range(2):
        x.zero(); r.copy_from(db); p.copy_from(db)
        res = ctx.cg_solve(op, db, x, r, p, 40)
        assert res["iters"] == 40 and res["fused"] and not res["stopped_on_pAp"]
        runs.append((ctx.nrm2(x), ctx.nrm2(r), res["rs"]))
    assert runs[0] == runs[1]                                   # bit-reproducible run to run
    assert abs(runs[0][1] ** 2 - runs[0][2]) <= 1e-12 * runs[0][2]  # state.rs is r.r of the returned r
    ax = ctx.empty(n)
    op.apply(x, ax)
    true_r = ctx.sub(db, ax)                                    # b - A x
    diff = ctx.sub(true_r, r)
    bn = ctx.nrm2(db)
    assert ctx.nrm2(diff) <= 1e-12 * bn, (ctx.nrm2(diff), bn)   # recurrence residual == true residual
    assert ctx.nrm2(r) < 0.9 * bn                               # and the iteration made progress
    # linearity of the operator: A(2x + p) == 2 A x + A p up to rounding
    comb = ctx.axpy(p, x, 2.0)                                  # p + x*2
    a1 = ctx.empty(n)
    op.apply(comb, a1)
    ap = ctx.empty(n)
    op.apply(p, ap)
    a2 = ctx.axpy(ap, ax, 2.0)
    assert ctx.nrm2(ctx.sub(a1, a2)) <= 1e-13 * ctx.nrm2(a1)
    for v in (db, x, r, p, ax, true_r, diff, comb, a1, ap, a2):
        v.free()


def test_assembled_equals_matrix_free_256(ctx):
    """256^3 (BASELINE config 3): the device-assembled SELL operator and the matrix-free stencil give identical bits,
    and CG on either stops after the same number of iterations with solutions within 1e-10."""
    from algotoolkit_b200 import synthetic

    g = 256
    n = g ** 3
    cx, cy, cz = synthetic.poisson_coeffs(g, g, g)
    st = ctx.operator_stencil7(g, g, g, cx, cy, cz)
    sell = ctx.operator_poisson_assembled(g, g, g, cx, cy, cz, capi.FORMAT_SELL)
    assert sell.nnz == st.nnz == 115099600
    v = ctx.vector(np.random.default_rng(1).standard_normal(n))
    y1, y2 = ctx.empty(n), ctx.empty(n)
    st.apply(v, y1)
    sell.apply(v, y2)
    assert ctx.nrm2(ctx.sub(y1, y2)) == 0.0                      # bit-identical
    b = ctx.vector(synthetic.poisson_rhs_r(g, g, g))
    sols = []
    for op in (st, sell):
        x, r, p = ctx.empty(n), ctx.empty(n), ctx.empty(n)
        x.zero(); r.copy_from(b); p.copy_from(b)
        res = ctx.cg_solve(op, b, x, r, p, 60)
        assert res["iters"] == 60
        sols.append(x)
    assert ctx.nrm2(ctx.sub(sols[0], sols[1])) <= 1e-10 * ctx.nrm2(sols[0])


# ------------------------------------------------------------------------------------------------ "next" rows
def test_gradient_descent(ctx, port):
    A = port.csc_from_triplets(3, 3, [0, 1, 0, 1, 2, 1, 2], [0, 0, 1, 1, 1, 2, 2], [4.0, -1.0, -1.0, 4.0, -1.0, -1.0, 4.0])
    op = ctx.operator_from_csc(3, 3, A.col_ptr, A.row_idx, A.vals)
    o = port.gd(A, [1.0, 5.0, 0.0], np.zeros(3), 1000, 1e-6)
    g = capi.gd_solve_host(ctx, op, [1.0, 5.0, 0.0], np.zeros(3), 1000, 1e-6)   # tests/itersolver_test.cpp:5-39
    assert g["iters"] == o["iters"] == 15 and np.allclose(g["x"], [0.625, 1.5, 0.375], atol=1e-5)
    assert relerr(g["x"], o["x"]) < REL
    n = 20
    cx, cy, cz, _ = port.poisson_coeffs(n, n, n)
    b = port.poisson_rhs(n, n, n, rhs_kind=1, seed=12345)
    Ap = port.poisson_csc(n, n, n)
    st = ctx.operator_stencil7(n, n, n, cx, cy, cz)
    for max_iter, tol in ((60, 1e-9), (200, 1e-3), (0, 1e-6)):
        o = port.gd(Ap, b, np.zeros(n ** 3), max_iter, tol)
        g = capi.gd_solve_host(ctx, st, b, np.zeros(n ** 3), max_iter, tol)
        assert abs(g["iters"] - o["iters"]) <= 2 and relerr(g["x"], o["x"]) < 1e-9 if o["iters"] else np.array_equal(g["x"], o["x"])
    ctx.set_reduction(capi.REDUCE_SEQUENTIAL)
    try:
        o = port.gd(Ap, b, np.zeros(n ** 3), 25, 1e-12)
        g = capi.gd_solve_host(ctx, st, b, np.zeros(n ** 3), 25, 1e-12)
        assert g["iters"] == o["iters"] == 25 and np.array_equal(g["x"], o["x"])   # reference-ordered sums: bit-exact
    finally:
        ctx.set_reduction(capi.REDUCE_TREE)


def test_arnoldi(ctx, port):
    rng = np.random.default_rng(0)
    n = 300
    D = rng.standard_normal((n, n)) * (rng.random((n, n)) < 0.05) + np.diag(5 + rng.random(n))
    r, c = np.nonzero(D)
    A = port.csc_from_triplets(n, n, r, c, D[r, c])
    op = ctx.operator_from_csc(n, n, A.col_ptr, A.row_idx, A.vals)
    q0 = rng.standard_normal(n)
    q0 /= np.linalg.norm(q0)
    o = port.arnoldi(A, q0, 12, 1e-10)
    g = capi.arnoldi_host(ctx, op, q0, 12, 1e-10)
    assert g["breakdown"] == -1 and np.abs(g["Q"] - o["Q"]).max() < 1e-10 and np.abs(g["H"] - o["H"]).max() < 1e-10
    assert np.abs(g["Q"] @ g["Q"].T - np.eye(12)).max() < 1e-10
    ctx.set_reduction(capi.REDUCE_SEQUENTIAL)
    try:
        g = capi.arnoldi_host(ctx, op, q0, 12, 1e-10)
        assert np.array_equal(g["Q"], o["Q"]) and np.array_equal(g["H"], o["H"])   # bit-exact in reference-ordered mode
    finally:
        ctx.set_reduction(capi.REDUCE_TREE)
    A2 = port.csc_from_triplets(4, 4, [0, 1, 2, 3], [0, 1, 2, 3], [1.0, 2.0, 3.0, 4.0])
    op2 = ctx.operator_from_csc(4, 4, A2.col_ptr, A2.row_idx, A2.vals)
    g2, o2 = capi.arnoldi_host(ctx, op2, [1.0, 0, 0, 0], 3, 1e-8), port.arnoldi(A2, [1.0, 0, 0, 0], 3, 1e-8)
    assert g2["breakdown"] == o2["breakdown"] == 1 and np.array_equal(g2["Q"], o2["Q"]) and np.array_equal(g2["H"], o2["H"])


# ------------------------------------------------------------------------------------------------ 8(f) rank 3
@pytest.mark.parametrize("name,iters", [("bcsstk01", 60), ("bcsstk08", 150)])
def test_cg_on_symmetric_expanded_matrix_market(ctx, port, name, iters):
    """Opt-in symmetric expansion -> the SPD matrix the file describes -> device SELL -> CG; the reference-ordered
    reductions reproduce the oracle's CG on the same CSC arrays bit for bit."""
    from algotoolkit_b200.mmio import read_matrix_market_csc
    from oracle.pyoracle import CscMatrix

    d = read_matrix_market_csc(mm_path(name), expand_symmetric=True)
    n = d["n_rows"]
    A = CscMatrix(n, n, d["col_ptr"], d["row_idx"], d["vals"])
    x_ext = port.u01_array(777, n)
    b = port.spmv(A, x_ext)
    op = ctx.operator_from_csc(n, n, d["col_ptr"], d["row_idx"], d["vals"])
    assert digest(op.apply_numpy(x_ext)) == digest(b)
    ora = port.cg(A, b, iters)
    ctx.set_reduction(capi.REDUCE_SEQUENTIAL)
    try:
        res = ctx.cg_solve_host(op, b, iters)
    finally:
        ctx.set_reduction(capi.REDUCE_TREE)
    assert res["iters"] == ora["iters"]
    assert digest(res["x"]) == digest(ora["x"]) and digest(res["r"]) == digest(ora["r"])
    tree = ctx.cg_solve_host(op, b, 10)  # a few iterations: rounding-order noise has not been amplified yet
    ora10 = port.cg(A, b, 10)
    assert tree["iters"] == ora10["iters"] and relerr(tree["x"], ora10["x"]) < 1e-8


# ------------------------------------------------------------------------------------------------ alternative kernels
@pytest.mark.parametrize("env", [{"ATK_STENCIL_TMA": "1"}, {"ATK_FUSED_SMEM": "0"}, {"ATK_STENCIL27_SIMPLE": "1"},
                                 {"ATK_STENCIL_TILE": "0", "ATK_STENCIL_KC": "16"}, {"ATK_STENCIL27_KC": "8"}])
def test_alternative_stencil_kernels_stay_bit_exact(ctx, port, env, monkeypatch):
    """The selectable kernel variants (TMA ring, L1-served fallbacks, other tile / chunk shapes) against the oracle."""
    from algotoolkit_b200 import synthetic

    for key, val in env.items():
        monkeypatch.setenv(key, val)
    nx, ny, nz = 66, 13, 21
    cx, cy, cz, _ = port.poisson_coeffs(nx, ny, nz)
    x = np.random.default_rng(3).standard_normal(nx * ny * nz)
    op = ctx.operator_stencil7(nx, ny, nz, cx, cy, cz)
    assert np.array_equal(op.apply_numpy(x), port.poisson_apply(nx, ny, nz, cx, cy, cz, x))
    b = port.poisson_rhs(nx, ny, nz, rhs_kind=1, seed=12345)
    res = ctx.cg_solve_host(op, b, 40)
    ora = port.cg(None, b, 40, stencil=(nx, ny, nz, cx, cy, cz))
    assert res["iters"] == ora["iters"] and relerr(res["x"], ora["x"]) < REL
    A27 = port.op27_csc(nx, ny, nz)
    op27 = ctx.operator_stencil27(nx, ny, nz, synthetic.stencil27_coefficients())
    assert np.array_equal(op27.apply_numpy(x), port.spmv(A27, x))


@pytest.mark.parametrize("dims", [(66, 13, 21), (128, 40, 70), (64, 4, 9)])
def test_fused_step_through_tensor_map_tma_is_bit_identical(ctx, port, dims, monkeypatch):
    """KA with its r / p plane tiles fetched by cp.async.bulk.tensor (3-D tensor maps, zero fill outside the grid) must
    give the same bits as the LDGSTS kernel: same arithmetic, different data path.  CUDA-graph loop, several z-chunks."""
    import hashlib

    nx, ny, nz = dims
    cx, cy, cz, _ = port.poisson_coeffs(nx, ny, nz)
    b = port.poisson_rhs(nx, ny, nz, rhs_kind=1, seed=5)
    monkeypatch.setenv("ATK_CG_ONCHIP", "0")
    monkeypatch.setenv("ATK_CG_PERSISTENT", "0")
    monkeypatch.setenv("ATK_FUSED_KC", "16")
    out = {}
    for tma in ("0", "1"):
        monkeypatch.setenv("ATK_FUSED_TMA", tma)
        op = ctx.operator_stencil7(nx, ny, nz, cx, cy, cz)
        res = ctx.cg_solve_host(op, b, 33)
        assert res["fused"] and not res["persistent"] and res["iters"] == 33
        out[tma] = [hashlib.sha256(res[k].tobytes()).hexdigest() for k in ("x", "r", "p")]
        if tma == "1":
            ora = port.cg(None, b, 33, stencil=(nx, ny, nz, cx, cy, cz))
            assert relerr(res["x"], ora["x"]) < REL
        op.destroy()
    assert out["0"] == out["1"]


# ------------------------------------------------------------------------------------------------ config 3 at full size
@pytest.mark.parametrize("n", [128, 256])
def test_poisson_cg_to_exit_full_size_vs_reference(ctx, n):
    """BASELINE config 3: CG to the |pAp| < 1e-12 exit against fixtures computed by the unmodified reference at the same
    size (tests/golden/make_golden_large.py): iteration count within +-2, solution within 1e-10 (norm and 512 sampled
    entries); at 128^3 the reference-ordered mode must reproduce the reference's x bit for bit."""
    import json
    import os

    from algotoolkit_b200 import synthetic
    from tests.conftest import GOLDEN

    path = os.path.join(GOLDEN, "golden_large.json")
    if not os.path.exists(path):
        pytest.skip("tests/golden/golden_large.json not generated")
    with open(path) as f:
        gl = json.load(f)
    if str(n) not in gl:
        pytest.skip(f"no {n}^3 fixture")
    g = gl[str(n)]
    cx, cy, cz = synthetic.poisson_coeffs(n, n, n)
    b = synthetic.poisson_rhs_r(n, n, n)
    assert abs(np.linalg.norm(b) - g["b_norm"]) <= 1e-13 * g["b_norm"]
    op = ctx.operator_stencil7(n, n, n, cx, cy, cz)
    res = ctx.cg_solve_host(op, b, 100000)
    assert res["stopped_on_pAp"] and abs(res["iters"] - g["iterations"]) <= 2
    x = res["x"]
    idx = np.array(g["sample_index"])
    want = np.array(g["sample_x"])
    assert abs(np.linalg.norm(x) - g["x_norm"]) <= REL * g["x_norm"]
    assert np.linalg.norm(x[idx] - want) <= REL * np.linalg.norm(want)
    if n <= 128:
        ctx.set_reduction(capi.REDUCE_SEQUENTIAL)
        try:
            seq = ctx.cg_solve_host(op, b, 100000)
        finally:
            ctx.set_reduction(capi.REDUCE_TREE)
        assert seq["iters"] == g["iterations"] and digest(seq["x"]) == g["x_sha256"]


@pytest.mark.parametrize("key", ["gmres_op27_24_K40", "gmres_op27_32_K60", "gmres_op27_64_K300"])
def test_gmres_op27_larger_systems_vs_reference(ctx, port, key):
    """BASELINE config 5 operator at sizes the reference's shipped sparse-H GMRES can still do (fixtures from
    make_golden_large.py gmres): per-projection / persistent MGS and the batched MGS within 1e-10 of the reference's
    residual ladder and solution; reference-ordered sums bit-exact at the smallest size."""
    import json
    import os

    from algotoolkit_b200 import synthetic
    from tests.conftest import GOLDEN

    path = os.path.join(GOLDEN, "golden_large.json")
    if not os.path.exists(path) or key not in json.load(open(path)):
        pytest.skip(f"no fixture {key}")
    g = json.load(open(path))[key]
    n, K, iters = g["n"], g["K"], g["max_iter"]
    rows = n ** 3
    op = ctx.operator_stencil27(n, n, n, synthetic.stencil27_coefficients())
    assert op.nnz == g["nnz"]
    b = op.apply_numpy(synthetic.u01_array(777, 0, rows))
    assert digest(b) == g["b_sha256"]  # the matrix-free apply is bit-identical to the reference's CSC matvec
    idx, want = np.array(g["sample_index"]), np.array(g["sample_x"])
    for ortho in (capi.GMRES_MGS, capi.GMRES_MGS_BATCHED):
        res = ctx.gmres_solve_host(op, b, np.zeros(rows), iters, K, 1e-6, ortho=ortho)
        assert res["status"] == g["status"] and len(res["resid"]) == len(g["resid"])
        assert np.allclose(res["resid"], g["resid"], rtol=REL, atol=0.0)
        assert abs(np.linalg.norm(res["x"]) - g["x_norm"]) <= REL * g["x_norm"]
        assert np.linalg.norm(res["x"][idx] - want) <= REL * np.linalg.norm(want)
    if rows <= 20000:
        ctx.set_reduction(capi.REDUCE_SEQUENTIAL)
        try:
            seq = ctx.gmres_solve_host(op, b, np.zeros(rows), iters, K, 1e-6)
        finally:
            ctx.set_reduction(capi.REDUCE_TREE)
        assert digest(seq["x"]) == g["x_sha256"] and list(seq["resid"]) == g["resid"]


@pytest.mark.parametrize("dims,ring", [((96, 96, 96), True), ((96, 96, 96), False), ((95, 97, 96), True), ((95, 97, 95), False),
                                       ((128, 128, 127), True)])
def test_gmres_persistent_step_with_w_in_shared_memory(ctx, dims, ring, monkeypatch):
    """Slices above 4096 rows per CTA: (ring) w in registers and the columns in a two-deep cp.async ring in shared memory
    (slices up to 14 336 rows, even column stride - 128x128x127 is 14 060 rows per CTA, next to the limit), or (no ring:
    ATK_MGS_NO_RING, or an odd column stride as in 95x97x95) w in shared memory with the update of projection i-1 fused
    into the products of projection i, 16-byte accesses when the stride is even and scalar otherwise.  Must agree with the
    per-projection kernels and with the batched form to 1e-10."""
    from algotoolkit_b200 import synthetic

    if not ring:
        monkeypatch.setenv("ATK_MGS_NO_RING", "1")
    nx, ny, nz = dims
    rows = nx * ny * nz
    op = ctx.operator_stencil27(nx, ny, nz, synthetic.stencil27_coefficients())
    b = op.apply_numpy(synthetic.u01_array(777, 0, rows))
    pers = ctx.gmres_solve_host(op, b, np.zeros(rows), 2, 24, 1e-6)
    assert pers["kernel_launches"] < 2 * (24 * 3 + 20)  # one step kernel per Arnoldi step, not one pair per projection
    monkeypatch.setenv("ATK_GMRES_NO_PERSISTENT", "1")
    plain = ctx.gmres_solve_host(op, b, np.zeros(rows), 2, 24, 1e-6)
    bat = ctx.gmres_solve_host(op, b, np.zeros(rows), 2, 24, 1e-6, ortho=capi.GMRES_MGS_BATCHED)
    for other in (plain, bat):
        assert np.allclose(pers["resid"], other["resid"], rtol=REL, atol=0.0)
        assert relerr(pers["x"], other["x"]) < REL
    op.destroy()


@pytest.mark.parametrize("dims", [(48, 48, 48), (41, 43, 45)])
def test_gmres_fused_update_and_products_launch_is_bit_identical(ctx, dims, monkeypatch):
    """Per-projection path on one GPU (what config 5 runs on at 16.7 M rows): one launch per projection applies the update of
    projection i-1 and forms the products of projection i (mgs_axmy_dot_kernel, 32 B/row) instead of a dot and an axmy launch
    (40 B/row).  Same thread layout and summation order as the dot kernel: with an even column stride (16-byte accesses in
    both forms) the residual ladder and x must not change by one bit.  With an odd stride every other column is misaligned,
    the dot kernel alternates between its 16-byte and its scalar layout while the fused launch (two columns at once) always
    takes the scalar one: equal to 1e-10."""
    from algotoolkit_b200 import synthetic

    nx, ny, nz = dims
    rows = nx * ny * nz
    op = ctx.operator_stencil27(nx, ny, nz, synthetic.stencil27_coefficients())
    b = op.apply_numpy(synthetic.u01_array(777, 0, rows))
    monkeypatch.setenv("ATK_GMRES_NO_PERSISTENT", "1")
    fused = ctx.gmres_solve_host(op, b, np.zeros(rows), 2, 30, 1e-6)
    monkeypatch.setenv("ATK_GMRES_NO_FUSED_DOT", "1")
    plain = ctx.gmres_solve_host(op, b, np.zeros(rows), 2, 30, 1e-6)
    assert fused["kernel_launches"] < plain["kernel_launches"] - 2 * 300  # 30*29/2 projections per restart: one launch each instead of two
    if rows % 2 == 0:
        assert list(fused["resid"]) == list(plain["resid"])
        assert digest(fused["x"]) == digest(plain["x"])
    else:
        assert np.allclose(fused["resid"], plain["resid"], rtol=REL, atol=0.0)
        assert relerr(fused["x"], plain["x"]) < REL
    op.destroy()


# ------------------------------------------------------------------------------------------------ CSR format kernels
def test_csr_stream_is_bit_exact_and_vector_kernel_still_covered(ctx, port, monkeypatch):
    """CSR format: short rows -> CSR-stream (per-row ascending sums: bit-identical to the CSC matvec, empty rows, rows
    straddling the block boundaries, exact-zero x entries); long rows or ATK_CSR_NO_STREAM=1 -> CSR-vector (lane tree)."""
    rng = np.random.default_rng(17)
    n_rows, n_cols = 5000, 4100
    lens = rng.integers(0, 40, n_rows)
    lens[100:400] = 0  # a run of empty rows inside a block
    lens[-50:] = 0     # trailing empty rows
    rows = np.repeat(np.arange(n_rows), lens)
    cols = np.concatenate([np.sort(rng.choice(n_cols, k, replace=False)) for k in lens]) if lens.sum() else np.zeros(0, int)
    vals = rng.standard_normal(rows.size)
    A = port.csc_from_triplets(n_rows, n_cols, rows, cols, vals)
    x = rng.standard_normal(n_cols)
    x[::7] = 0.0
    want = port.spmv(A, x)
    op = ctx.operator_from_csc(n_rows, n_cols, A.col_ptr, A.row_idx, A.vals, capi.FORMAT_CSR_VECTOR)
    assert np.array_equal(op.apply_numpy(x), want)  # warp-private staging of the matrix entries (default)
    monkeypatch.setenv("ATK_CSR_STREAM_V1", "1")
    assert np.array_equal(op.apply_numpy(x), want)  # entry-parallel form: products staged in shared memory
    monkeypatch.delenv("ATK_CSR_STREAM_V1")
    # nnz not a multiple of four and a last block that ends inside the final 16-byte chunk
    for extra in (1, 2, 3):
        rr = np.concatenate([rows, np.full(extra, n_rows - 1)])
        cc = np.concatenate([cols, np.arange(extra)])
        vv = np.concatenate([vals, rng.standard_normal(extra)])
        A3 = port.csc_from_triplets(n_rows, n_cols, rr, cc, vv)
        op3 = ctx.operator_from_csc(n_rows, n_cols, A3.col_ptr, A3.row_idx, A3.vals, capi.FORMAT_CSR_VECTOR)
        assert np.array_equal(op3.apply_numpy(x), port.spmv(A3, x))
        op3.destroy()
    monkeypatch.setenv("ATK_CSR_NO_STREAM", "1")
    opv = ctx.operator_from_csc(n_rows, n_cols, A.col_ptr, A.row_idx, A.vals, capi.FORMAT_CSR_VECTOR)
    assert relerr(opv.apply_numpy(x), want) < 1e-14
    monkeypatch.delenv("ATK_CSR_NO_STREAM")
    # one row longer than half the product buffer: the operator falls back to CSR-vector by itself
    n2 = 3000
    rows2 = np.concatenate([np.zeros(n2, int), np.arange(1, 50)])
    cols2 = np.concatenate([np.arange(n2), rng.integers(0, n2, 49)])
    A2 = port.csc_from_triplets(50, n2, rows2, cols2, rng.standard_normal(rows2.size))
    x2 = rng.standard_normal(n2)
    op2 = ctx.operator_from_csc(50, n2, A2.col_ptr, A2.row_idx, A2.vals, capi.FORMAT_CSR_VECTOR)
    assert relerr(op2.apply_numpy(x2), port.spmv(A2, x2)) < 1e-13
    # CG through the CSR format on the assembled Poisson matrix: same iterations and solution as the oracle
    n = 16
    cx, cy, cz, _ = port.poisson_coeffs(n, n, n)
    b = port.poisson_rhs(n, n, n, rhs_kind=1, seed=12345)
    opp = ctx.operator_poisson_assembled(n, n, n, cx, cy, cz, capi.FORMAT_CSR_VECTOR)
    res = ctx.cg_solve_host(opp, b, 100000)
    ora = port.cg(None, b, 100000, stencil=(n, n, n, cx, cy, cz))
    assert res["iters"] == ora["iters"] and relerr(res["x"], ora["x"]) < REL
