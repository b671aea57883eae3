"""The oracle (oracle/krylov_oracle.c) pinned against (a) the committed golden vectors produced by the unmodified
reference and (b) the unmodified reference itself when oracle/_ref is available.  CPU only."""
import hashlib

import numpy as np
import pytest

from tests.conftest import mm_path


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.float64).tobytes()).hexdigest()


# ------------------------------------------------------------------------------------------------ vs golden vectors
def test_kat_spmv_cg(port, golden):
    k = golden["kat"]["spmv_3x3"]
    A = port.csc_from_triplets(3, 3, k["rows"], k["cols"], k["vals"])
    assert port.spmv(A, k["x"]).tolist() == k["y"] == k["expect"]
    assert port.csc_get(A, 2, 1) == 5.0 and port.csc_get(A, 0, 0) == 0.0
    with pytest.raises(IndexError):
        port.csc_get(A, 3, 0)
    k = golden["kat"]["cg_tridiag3"]
    A = port.csc_from_triplets(3, 3, k["rows"], k["cols"], k["vals"])
    res = port.cg(A, k["b"], k["max_iter"])
    assert res["x"].tolist() == k["x"] and res["iters"] == k["iters"]
    assert np.allclose(res["x"], k["expect"], atol=k["tol"])
    one = port.cg(A, k["b"], 1)
    k1 = golden["kat"]["cg_tridiag3_1iter"]
    assert one["x"].tolist() == k1["x"] and one["r"].tolist() == k1["r"] and one["p"].tolist() == k1["p"]
    assert port.cg(A, [0.0, 0.0, 0.0], 1000)["x"].tolist() == golden["kat"]["cg_zero_rhs"]["x"]
    kd = golden["kat"]["cg_diag4"]
    Ad = port.csc_from_triplets(4, 4, [0, 1, 2, 3], [0, 1, 2, 3], [4.0, 3.0, 2.0, 1.0])
    rd = port.cg(Ad, kd["b"], kd["max_iter"])
    assert rd["x"].tolist() == kd["x"] and np.allclose(rd["x"], kd["expect"], atol=kd["tol"])
    k2 = golden["kat"]["cg_2x2_notebook"]
    A2 = port.csc_from_triplets(2, 2, [0, 0, 1, 1], [0, 1, 0, 1], [4.0, 1.0, 1.0, 3.0])
    assert np.allclose(port.cg(A2, [1.0, 2.0], 100)["x"], k2["expect"], rtol=1e-14)


def test_kat_gmres(port, golden):
    k = golden["kat"]["cg_tridiag3"]
    A = port.csc_from_triplets(3, 3, k["rows"], k["cols"], k["vals"])
    for K in (1, 2, 3):
        g = golden["kat"][f"gmres_tridiag3_K{K}"]
        res = port.gmres(A, [1.0, 5.0, 0.0], np.zeros(3), 10000, K, 1e-12)
        assert res["x"].tolist() == g["x"] and res["status"] == g["status"] and len(res["resid"]) == g["resid"]["n"]
        head = g["resid"].get("all", g["resid"].get("head"))
        assert res["resid"][: len(head)].tolist() == head
    g = golden["kat"]["gmres_2x2_notebook"]
    A2 = port.csc_from_triplets(2, 2, [0, 0, 1, 1], [0, 1, 0, 1], [4.0, 1.0, 1.0, 3.0])
    res = port.gmres(A2, [1.0, 2.0], [0.0, 0.0], 100, 1, 1e-6)
    assert res["x"].tolist() == g["x"] == g["expect"]
    assert ["%g" % v for v in res["resid"][:5]] == g["expect_ladder_6sig"]  # the notebook's printed ladder
    assert len(res["resid"]) - 2 == g["expect_converged_at"] and res["status"] == 1
    with pytest.raises(ValueError):
        port.gmres(A2, [1.0, 2.0], [0.0, 0.0, 0.0], 10, 1, 1e-6)
    with pytest.raises(ValueError):
        port.gmres(A2, [1.0, 2.0], [0.0, 0.0], 10, 3, 1e-6)


@pytest.mark.parametrize("n", [8, 16, 32])
def test_poisson_readme_golden(port, golden, n):
    g = golden["poisson_readme"][str(n)]
    b = port.poisson_rhs(n, n, n, rhs_kind=0)
    assert digest(b) == g["b_sha256"]
    cx, cy, cz, _ = port.poisson_coeffs(n, n, n)
    res = port.cg(None, b, 2000, stencil=(n, n, n, cx, cy, cz))
    assert res["iters"] == g["iters"] == 1 and res["last_pAp"] == g["exit_pAp"] and digest(res["x"]) == g["u_sha256"]
    grid = np.arange(n) / (n - 1)
    an = np.sin(np.pi * grid)[None, None, :] * np.sin(np.pi * grid)[None, :, None] * np.sin(np.pi * grid)[:, None, None]
    err = np.abs(res["x"].reshape(n, n, n) - an)
    assert float(err.max()) == g["max_err"]
    if n == 32:  # the numbers stored in code/PDE.ipynb
        assert f"{err.max():.6e}" == golden["poisson_readme"]["notebook_32"]["max_err"]
        assert f"{np.sqrt((err ** 2).mean()):.6e}" == golden["poisson_readme"]["notebook_32"]["l2_err"]


@pytest.mark.parametrize("n", [12, 32, 64])
def test_poisson_rhsr_golden(port, golden, n):
    g = golden["poisson_rhsr"][str(n)]
    b = port.poisson_rhs(n, n, n, rhs_kind=1, seed=12345)
    assert digest(b) == g["b_sha256"]
    cx, cy, cz, _ = port.poisson_coeffs(n, n, n)
    res = port.cg(None, b, 100000, stencil=(n, n, n, cx, cy, cz), trace=True)
    assert res["iters"] == g["iters"] and res["last_pAp"] == g["exit_pAp"]
    assert digest(res["x"]) == g["x_sha256"] and digest(res["r"]) == g["r_sha256"] and digest(res["p"]) == g["p_sha256"]
    it = g["iters"]
    assert res["rs_trace"][:it][::20].tolist() == g["rs_trace_every20"]
    if n <= 32:  # assembled CSC path gives the same bits as the matrix-free restatement
        A = port.poisson_csc(n, n, n)
        res2 = port.cg(A, b, 100000)
        assert digest(res2["x"]) == g["x_sha256"]
    if n == 32:
        r50 = port.cg(None, b, 50, stencil=(n, n, n, cx, cy, cz))
        assert digest(r50["x"]) == golden["poisson_rhsr"]["32_fixed50"]["x_sha256"]


@pytest.mark.parametrize("name", ["bcsstk01", "bcsstk08"])
def test_gmres_matrix_market_golden(port, golden, name):
    g = golden["gmres_mm"][name]
    A = port.read_matrix_market(mm_path(name))
    assert (A.n_rows, A.nnz) == (g["n"], g["nnz"])
    assert hashlib.sha256(A.col_ptr.tobytes()).hexdigest() == g["col_ptr_sha"]
    assert hashlib.sha256(A.row_idx.tobytes()).hexdigest() == g["row_idx_sha"] and digest(A.vals) == g["vals_sha"]
    b = port.spmv(A, port.u01_array(777, A.n_cols))
    assert digest(b) == g["b_sha256"]
    for key, run in g["runs"].items():
        if run["K"] > 100 and run["max_iter"] > 2:
            continue  # the 6-restart K=300 ladder is checked on the GPU path; keep the CPU suite short
        res = port.gmres(A, b, np.zeros(A.n_rows), run["max_iter"], run["K"], 1e-6)
        assert res["resid"].tolist() == run["resid"]["all"], key
        assert digest(res["x"]) == run["x_sha256"] and res["status"] == run["status"]


def test_gmres_bcsstk08_k300_first_restarts(port, golden):
    run = golden["gmres_mm"]["bcsstk08"]["runs"]["K300_it6"]
    A = port.read_matrix_market(mm_path("bcsstk08"))
    b = port.spmv(A, port.u01_array(777, A.n_cols))
    res = port.gmres(A, b, np.zeros(A.n_rows), 2, 300, 1e-6)
    assert res["resid"].tolist() == run["resid"]["all"][:3]
    assert "%.6e" % res["resid"][1] == "9.978768e+07"  # residual after the first restart


@pytest.mark.parametrize("key", ["8_K30", "12_K60"])
def test_gmres_op27_golden(port, golden, key):
    g = golden["gmres_op27"][key]
    A = port.op27_csc(g["n"], g["n"], g["n"])
    assert A.nnz == g["nnz"] and digest(A.vals) == g["vals_sha"]
    b = port.spmv(A, port.u01_array(777, A.n_cols))
    assert digest(b) == g["b_sha256"]
    res = port.gmres(A, b, np.zeros(A.n_rows), g["max_iter"], g["K"], 1e-6)
    assert res["resid"].tolist() == g["resid"]["all"] and digest(res["x"]) == g["x_sha256"]


# ------------------------------------------------------------------------------------------------ vs the live reference
def test_port_vs_reference_l0(port, ref):
    rng = np.random.default_rng(1)
    for n in (0, 1, 7, 1000):
        a, b = rng.standard_normal(n), rng.standard_normal(n)
        assert port.dot(a, b) == ref.dot(a, b) and port.nrm2(a) == ref.nrm2(a)
        assert np.array_equal(port.vec_scale(a, 0.3), ref.vec_scale(a, 0.3))
        assert np.array_equal(port.vec_div(a, 0.3), ref.vec_div(a, 0.3))
        assert np.array_equal(port.vec_add(a, b), ref.vec_add(a, b))
        assert np.array_equal(port.vec_sub(a, b), ref.vec_sub(a, b))
    with pytest.raises(RuntimeError):
        ref.vec_div(np.ones(3), 0.0)
    with pytest.raises(RuntimeError):
        port.vec_div(np.ones(3), 0.0)


def test_port_vs_reference_matrices(port, ref):
    rng = np.random.default_rng(2)
    n_rows, n_cols = 40, 33
    mask = rng.random((n_rows, n_cols)) < 0.15
    r, c = np.nonzero(mask)
    v = rng.standard_normal(r.size)
    v[::9] = 0.0  # explicit zeros are dropped by addValue
    perm = rng.permutation(r.size)
    A = port.csc_from_triplets(n_rows, n_cols, r[perm], c[perm], v[perm])
    M = ref.mat_from_triplets(n_rows, n_cols, r[perm], c[perm], v[perm])
    Ar = M.export()
    assert np.array_equal(A.col_ptr, Ar.col_ptr) and np.array_equal(A.row_idx, Ar.row_idx) and np.array_equal(A.vals, Ar.vals)
    x = rng.standard_normal(n_cols)
    x[::4] = 0.0
    assert np.array_equal(port.spmv(A, x), M.spmv(x))
    for (i, j) in [(0, 0), (5, 7), (39, 32)]:
        assert port.csc_get(A, i, j) == M.get(i, j)
    with pytest.raises(IndexError):
        M.get(40, 0)
    with pytest.raises(ValueError):
        M.spmv(np.ones(n_cols + 1))
    with pytest.raises(IndexError):
        ref.mat_from_triplets(3, 3, [3], [0], [1.0])
    for name in ("bcsstk01", "bcsstk08"):
        P_, R_ = port.read_matrix_market(mm_path(name)), ref.mat_read_mm(mm_path(name)).export()
        assert np.array_equal(P_.col_ptr, R_.col_ptr) and np.array_equal(P_.row_idx, R_.row_idx) and np.array_equal(P_.vals, R_.vals)


@pytest.mark.parametrize("dims", [(4, 5, 6), (9, 7, 8)])
def test_port_vs_reference_poisson(port, ref, dims):
    nx, ny, nz = dims
    A, Ar = port.poisson_csc(nx, ny, nz, 1.0, 2.0, 0.5), ref.mat_poisson(nx, ny, nz, 1.0, 2.0, 0.5).export()
    assert np.array_equal(A.col_ptr, Ar.col_ptr) and np.array_equal(A.row_idx, Ar.row_idx) and np.array_equal(A.vals, Ar.vals)
    for kind in (0, 1):
        assert np.array_equal(port.poisson_rhs(nx, ny, nz, 1.0, 2.0, 0.5, kind, 99), ref.poisson_rhs(nx, ny, nz, 1.0, 2.0, 0.5, kind, 99))
    res = ref.poisson_solve(nx, ny, nz, 1.0, 2.0, 0.5, rhs_kind=1, seed=99, max_iter=25)
    cx, cy, cz, _ = port.poisson_coeffs(nx, ny, nz, 1.0, 2.0, 0.5)
    mine = port.cg(None, res["b"], 25, stencil=(nx, ny, nz, cx, cy, cz))
    assert np.array_equal(mine["x"], res["u"])
    tr = ref.cg_trace(ref.mat_from_csc(A), res["b"], 25)
    assert np.array_equal(tr["x"], res["u"]) and tr["iters"] == mine["iters"]  # the instrumented restatement == solve()


def test_port_vs_reference_gmres_random(port, ref):
    rng = np.random.default_rng(4)
    n = 60
    dense = rng.standard_normal((n, n)) * (rng.random((n, n)) < 0.2) + np.diag(8.0 + rng.random(n))
    r, c = np.nonzero(dense)
    A = port.csc_from_triplets(n, n, r, c, dense[r, c])
    M = ref.mat_from_csc(A)
    b = rng.standard_normal(n)
    for K, iters, dense_h in ((5, 4, False), (17, 3, True), (60, 2, True)):
        a = port.gmres(A, b, np.zeros(n), iters, K, 1e-9)
        g = ref.gmres(M, b, np.zeros(n), iters, K, 1e-9, dense_h=dense_h)
        assert a["resid"].tolist() == g["resid"].tolist() and np.array_equal(a["x"], g["x"]) and a["status"] == g["status"]


def test_python_matrix_market_reader_matches_reference_semantics(port):
    """algotoolkit_b200/mmio.py (input handling for Python drivers) gives the same CSC arrays as the reference reader."""
    from algotoolkit_b200 import mmio

    for name in ("bcsstk01", "bcsstk08"):
        a, b = mmio.read_matrix_market_csc(mm_path(name)), port.read_matrix_market(mm_path(name))
        assert (a["n_rows"], a["n_cols"]) == (b.n_rows, b.n_cols)
        assert np.array_equal(a["col_ptr"], b.col_ptr) and np.array_equal(a["row_idx"], b.row_idx)
        assert np.array_equal(a["vals"], b.vals)


def test_next_rows_gd_and_arnoldi_vs_reference(port, ref):
    """GradientDescent (IterSolver.hpp:23-44) and Krylov::Arnoldi (KrylovSubspace.hpp:11-38)."""
    A = port.csc_from_triplets(3, 3, [0, 1, 0, 1, 2, 1, 2], [0, 0, 1, 1, 1, 2, 2], [4.0, -1.0, -1.0, 4.0, -1.0, -1.0, 4.0])
    M = ref.mat_from_csc(A)
    g = port.gd(A, [1.0, 5.0, 0.0], np.zeros(3), 1000, 1e-6)  # tests/itersolver_test.cpp:5-39
    assert np.array_equal(g["x"], ref.gd(M, [1.0, 5.0, 0.0], np.zeros(3), 1000, 1e-6)["x"])
    assert np.allclose(g["x"], [0.625, 1.5, 0.375], atol=1e-5) and g["iters"] == 15
    rng = np.random.default_rng(0)
    n = 40
    D = rng.standard_normal((n, n)) * (rng.random((n, n)) < 0.2) + np.diag(5 + rng.random(n))
    D = (D + D.T) / 2
    r, c = np.nonzero(D)
    A = port.csc_from_triplets(n, n, r, c, D[r, c])
    M = ref.mat_from_csc(A)
    b = rng.standard_normal(n)
    for max_iter, tol in ((500, 1e-9), (7, 1e-12), (0, 1e-6)):
        assert np.array_equal(port.gd(A, b, np.zeros(n), max_iter, tol)["x"], ref.gd(M, b, np.zeros(n), max_iter, tol)["x"])
    q0 = rng.standard_normal(n)
    q0 /= np.linalg.norm(q0)
    a, ar = port.arnoldi(A, q0, 8, 1e-10), ref.arnoldi(M, q0, 8, 1e-10)
    assert np.array_equal(a["Q"], ar["Q"]) and np.array_equal(a["H"], ar["H"]) and a["breakdown"] == -1
    assert np.abs(a["Q"] @ a["Q"].T - np.eye(8)).max() < 1e-10      # tests/KrylovSubspace_test.cpp: orthonormal basis
    assert np.abs(D @ a["Q"][:7].T - a["Q"].T @ a["H"]).max() < 1e-12  # A Q = Q H
    A2 = port.csc_from_triplets(4, 4, [0, 1, 2, 3], [0, 1, 2, 3], [1.0, 2.0, 3.0, 4.0])
    a2, r2 = port.arnoldi(A2, [1.0, 0, 0, 0], 3, 1e-8), ref.arnoldi(ref.mat_from_csc(A2), [1.0, 0, 0, 0], 3, 1e-8)
    assert a2["breakdown"] == 1 and r2["log"] == "Breakdown: Norm below tolerance at iteration 1\n"
    assert np.array_equal(a2["Q"], r2["Q"]) and np.array_equal(a2["H"], r2["H"])


def test_port_matches_full_size_reference_fixture_128(port):
    """Config 3 at 128^3: the fixture was produced by the unmodified reference (make_golden_large.py); the C port must
    reproduce its iteration count and solution bit for bit (18 s of one core)."""
    import hashlib
    import json
    import os

    from tests.conftest import GOLDEN

    path = os.path.join(GOLDEN, "golden_large.json")
    if not os.path.exists(path):
        pytest.skip("golden_large.json not generated")
    g = json.load(open(path))["128"]
    n = 128
    cx, cy, cz, _ = port.poisson_coeffs(n, n, n)
    b = port.poisson_rhs(n, n, n, rhs_kind=1, seed=12345)
    res = port.cg(None, b, 100000, stencil=(n, n, n, cx, cy, cz))
    assert res["iters"] == g["iterations"]
    assert hashlib.sha256(np.ascontiguousarray(res["x"]).tobytes()).hexdigest() == g["x_sha256"]
    gm = json.load(open(path))["gmres_op27_32_K60"]
    A = port.op27_csc(32, 32, 32)
    bb = port.spmv(A, port.u01_array(777, A.n_cols))
    gp = port.gmres(A, bb, np.zeros(A.n_rows), gm["max_iter"], gm["K"], 1e-6)
    assert list(gp["resid"]) == gm["resid"]
    assert hashlib.sha256(np.ascontiguousarray(gp["x"]).tobytes()).hexdigest() == gm["x_sha256"]
