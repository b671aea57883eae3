"""ILU-preconditioned GMRES branch: ILUPreconditioner::compute / solve and
GMRES::enablePreconditioner().  The fixtures in tests/golden/golden_ilu.json come from the unmodified reference
(tests/golden/make_golden_ilu.py).  CPU tests pin the C restatement; `gpu` tests compare the device path through the
C ABI: factors bit-exact always, sweeps / GMRES bit-exact with reference-ordered sums and within 1e-10 otherwise."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

from tests.conftest import GOLDEN, ROOT, mm_path

REL = 1e-10


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.float64).tobytes()).hexdigest()


def relerr(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))


@pytest.fixture(scope="module")
def gilu():
    with open(os.path.join(GOLDEN, "golden_ilu.json")) as f:
        return json.load(f)


def case_matrix(port, name, e):
    from algotoolkit_b200.mmio import read_matrix_market_csc
    from oracle.pyoracle import CscMatrix

    if "rows" in e:
        return port.csc_from_triplets(e["n"], e["n"], e["rows"], e["cols"], e["vals"])
    if name.startswith("bcsstk01"):
        d = read_matrix_market_csc(mm_path("bcsstk01"), expand_symmetric=name.endswith("symmetric"))
        return CscMatrix(d["n_rows"], d["n_cols"], d["col_ptr"], d["row_idx"], d["vals"])
    if name == "poisson_4":
        return port.poisson_csc(4, 4, 4)
    if name == "op27_5":
        return port.op27_csc(5, 5, 5)
    raise KeyError(name)


CASES = ["ilu_test_3x3", "bcsstk01_as_loaded", "bcsstk01_symmetric", "random_10", "random_37", "random_80", "drop_rules_12",
         "poisson_4", "op27_5"]


# ------------------------------------------------------------------------------------------------ oracle (CPU)
@pytest.mark.parametrize("name", CASES)
def test_oracle_restatement_matches_reference_golden(port, gilu, name):
    e = gilu["cases"][name]
    A = case_matrix(port, name, e)
    n = e["n"]
    LU = port.ilu_compute(A)
    Lf, Uf = port.ilu_factors(LU)
    assert digest(Lf) == e["L_sha256"] and digest(Uf) == e["U_sha256"]
    assert np.count_nonzero(Lf) == e["L_nnz"] and np.count_nonzero(Uf) == e["U_nnz"]
    b = port.u01_array(99, n) * 2.0 - 1.0
    assert digest(port.ilu_solve(LU, b)) == e["solve_sha256"]
    rhs = port.spmv(A, port.u01_array(777, n))
    g = port.gmres_precond(A, rhs, np.zeros(n), e["max_iter"], e["K"], e["tol"])
    assert digest(g["x"]) == e["gmres_x_sha256"] and g["status"] == e["gmres_status"] and g["rc"] == e["gmres_rc"]
    assert list(g["resid"]) == e["gmres_resid"]


def test_oracle_ilu_kat_and_errors(port, gilu):
    e = gilu["cases"]["ilu_test_3x3"]  # tests/ILU_test.cpp:81-105: A x = b with x_true = (1, 2, 3)
    A = case_matrix(port, "ilu_test_3x3", e)
    LU = port.ilu_compute(A)
    x = port.ilu_solve(LU, port.spmv(A, [1.0, 2.0, 3.0]))
    assert np.allclose(x, [1.0, 2.0, 3.0], atol=1e-10)
    Lf, Uf = port.ilu_factors(LU)
    assert np.array_equal(np.diag(Lf), np.ones(3)) and np.array_equal(np.triu(Lf, 1), np.zeros((3, 3)))  # :108-139
    assert np.array_equal(np.tril(Uf, -1), np.zeros((3, 3)))
    assert gilu["singular_3x3"] == "Matrix is numerically singular"
    with pytest.raises(RuntimeError):  # :64-76
        port.ilu_compute(port.csc_from_triplets(3, 3, [0], [1], [1.0]))
    with pytest.raises(ValueError):  # :57-62
        port.ilu_compute(port.csc_from_triplets(3, 4, [0], [1], [1.0]))


def test_oracle_vs_reference_live(port, ref):
    """Fresh random systems against the real reference (skipped where oracle/_ref cannot be built)."""
    rng = np.random.default_rng(11)
    for n in (6, 23, 41):
        D = np.where(rng.random((n, n)) < 0.2, rng.standard_normal((n, n)), 0.0)
        D[np.arange(n), np.arange(n)] = np.abs(D).sum(1) + 0.5
        c, r = np.nonzero(D.T)
        A = port.csc_from_triplets(n, n, r, c, D[r, c])
        M = ref.mat_from_csc(A)
        f = ref.ilu(M)
        LU = port.ilu_compute(A)
        Lf, Uf = port.ilu_factors(LU)
        assert np.array_equal(Lf, f["L"]) and np.array_equal(Uf, f["U"])
        b = rng.standard_normal(n)
        assert np.array_equal(port.ilu_solve(LU, b), f["solve"](b))
        rhs = port.spmv(A, port.u01_array(5, n))
        gp = port.gmres_precond(A, rhs, np.zeros(n), 2, min(n, 8), 1e-12)
        gr = ref.gmres_precond(M, rhs, np.zeros(n), 2, min(n, 8), 1e-12)
        assert np.array_equal(gp["x"], gr["x"]) and gp["status"] == gr["status"]
        f["free"]()


# ------------------------------------------------------------------------------------------------ ILU(0) (CPU)
def _no_fill_matrices(port, rng):
    """Matrices whose full elimination creates no fill-in: there ILU(0) must be the reference's ILU bit for bit."""
    out = []
    n = 57  # tridiagonal
    rows = np.concatenate([np.arange(n), np.arange(1, n), np.arange(n - 1)])
    cols = np.concatenate([np.arange(n), np.arange(n - 1), np.arange(1, n)])
    vals = np.concatenate([4.0 + rng.random(n), -1.0 - rng.random(n - 1), -1.0 - rng.random(n - 1)])
    out.append(("tridiagonal", port.csc_from_triplets(n, n, rows, cols, vals)))
    m, k = 9, 4  # block diagonal with dense blocks
    D = np.zeros((m * k, m * k))
    for q in range(m):
        B = rng.standard_normal((k, k))
        B[np.arange(k), np.arange(k)] = np.abs(B).sum(1) + 1.0
        D[q * k:(q + 1) * k, q * k:(q + 1) * k] = B
    c, r = np.nonzero(D.T)
    out.append(("block_diagonal", port.csc_from_triplets(m * k, m * k, r, c, D[r, c])))
    n = 40  # "arrow pointing down-right": dense last row and column, diagonal elsewhere
    D = np.diag(3.0 + rng.random(n))
    D[-1, :] = rng.standard_normal(n)
    D[:, -1] = rng.standard_normal(n)
    D[-1, -1] = 50.0
    c, r = np.nonzero(D.T)
    out.append(("arrow", port.csc_from_triplets(n, n, r, c, D[r, c])))
    return out


def test_ilu0_restatement_equals_reference_where_no_fill_occurs(port, ref):
    """The reference has no sparse ILU(0); what pins the restatement is that on matrices without fill-in it must give
    the reference's own factors, solves and preconditioned GMRES bit for bit."""
    rng = np.random.default_rng(3)
    for name, A in _no_fill_matrices(port, rng):
        n = A.n_rows
        M = ref.mat_from_csc(A)
        f = ref.ilu(M)
        LU0 = port.ilu0_compute(A)
        assert np.array_equal(LU0, port.ilu_compute(A)), name
        L0, U0 = port.ilu_factors(LU0)
        assert np.array_equal(L0, f["L"]) and np.array_equal(U0, f["U"]), name
        b = rng.standard_normal(n)
        assert np.array_equal(port.ilu_solve(LU0, b), f["solve"](b)), name
        rhs = port.spmv(A, port.u01_array(5, n))
        g0 = port.gmres_precond_ilu0(A, rhs, np.zeros(n), 2, 6, 1e-12)
        gr = ref.gmres_precond(M, rhs, np.zeros(n), 2, 6, 1e-12)
        assert np.array_equal(g0["x"], gr["x"]) and g0["status"] == gr["status"], name
        f["free"]()


def test_ilu0_restatement_drops_fill(port):
    """With fill-in the two factorisations differ, ILU(0) keeps A's pattern, and L*U reproduces A on that pattern."""
    A = port.poisson_csc(5, 5, 5)
    LU0, LU = port.ilu0_compute(A), port.ilu_compute(A)
    assert not np.array_equal(LU0, LU)
    dense = A.to_scipy().toarray()
    assert not (LU0[dense == 0.0] != 0.0).any()
    L0, U0 = port.ilu_factors(LU0)
    prod = L0 @ U0
    assert np.allclose(prod[dense != 0.0], dense[dense != 0.0], rtol=1e-12, atol=1e-12)
    with pytest.raises(RuntimeError, match="numerically singular"):
        port.ilu0_compute(port.csc_from_triplets(3, 3, [0, 1, 2, 2], [0, 1, 0, 1], [1.0, 1.0, 1.0, 1.0]))  # no (2,2) entry


# ------------------------------------------------------------------------------------------------ device (GPU)
def _ilu0_cases(port):
    rng = np.random.default_rng(8)
    cases = _no_fill_matrices(port, rng)
    cases.append(("poisson_6", port.poisson_csc(6, 6, 6)))
    cases.append(("op27_5", port.op27_csc(5, 5, 5)))
    n = 300
    D = np.where(rng.random((n, n)) < 0.03, rng.standard_normal((n, n)), 0.0)
    D[np.arange(n), np.arange(n)] = np.abs(D).sum(1) + 1.0
    D[17, :17] = 1e-13  # multipliers below the 1e-12 threshold: left in place, their row update skipped
    c, r = np.nonzero(D.T)
    cases.append(("random_300", port.csc_from_triplets(n, n, r, c, D[r, c])))
    return cases


@pytest.mark.gpu
def test_device_ilu0_factors_solves_and_gmres_are_bit_exact(ctx, port):
    """Level-scheduled sparse ILU(0) on the device against the dense restatement with a pattern mask: factors, solves
    (every row sums in ascending column order, whatever the reduction mode) and - with reference-ordered dot products -
    the preconditioned GMRES are bit-identical."""
    from algotoolkit_b200 import capi

    rng = np.random.default_rng(1)
    for name, A in _ilu0_cases(port):
        n = A.n_rows
        pc = ctx.ilu0_compute(n, n, A.col_ptr, A.row_idx, A.vals)
        LU0 = port.ilu0_compute(A)
        Lw, Uw = port.ilu_factors(LU0)
        Lf, Uf = pc.factors()
        assert np.array_equal(Lf, Lw) and np.array_equal(Uf, Uw), name
        b = rng.standard_normal(n)
        assert np.array_equal(pc.solve_host(b), port.ilu_solve(LU0, b)), name
        rhs = port.spmv(A, port.u01_array(777, n))
        op = ctx.operator_from_csc(n, n, A.col_ptr, A.row_idx, A.vals)
        K = min(n, 10)
        want = port.gmres_precond_ilu0(A, rhs, np.zeros(n), 3, K, 1e-10)
        tree = ctx.gmres_solve_host(op, rhs, np.zeros(n), 3, K, 1e-10, precond=pc)
        assert tree["status"] == want["status"] and np.allclose(tree["resid"], want["resid"], rtol=1e-8, atol=1e-12), name
        ctx.set_reduction(capi.REDUCE_SEQUENTIAL)
        try:
            seq = ctx.gmres_solve_host(op, rhs, np.zeros(n), 3, K, 1e-10, precond=pc)
        finally:
            ctx.set_reduction(capi.REDUCE_TREE)
        assert np.array_equal(seq["x"], want["x"]) and list(seq["resid"]) == list(want["resid"]), name
        op.destroy()
        pc.destroy()


@pytest.mark.gpu
def test_device_ilu0_errors_and_large_system(ctx, port):
    with pytest.raises(RuntimeError, match="numerically singular"):  # a missing diagonal entry is a zero pivot
        ctx.ilu0_compute(3, 3, *(lambda A: (A.col_ptr, A.row_idx, A.vals))(port.csc_from_triplets(3, 3, [0, 1, 2, 2], [0, 1, 0, 1], [1.0] * 4)))
    with pytest.raises(ValueError, match="square"):
        ctx.ilu0_compute(3, 4, np.zeros(5, np.int32), np.zeros(0, np.int32), np.zeros(0))
    # a system the dense elimination cannot hold (n = 262144 > 13500): ILU(0)-preconditioned GMRES on the 27-point operator
    from algotoolkit_b200 import synthetic

    g = 64
    A = port.op27_csc(g, g, g)
    n = A.n_rows
    with pytest.raises(ValueError):
        ctx.ilu_compute(n, n, A.col_ptr, A.row_idx, A.vals)
    pc = ctx.ilu0_compute(n, n, A.col_ptr, A.row_idx, A.vals)
    op = ctx.operator_from_csc(n, n, A.col_ptr, A.row_idx, A.vals)
    x_ext = synthetic.u01_array(777, 0, n)
    b = op.apply_numpy(x_ext)
    plain = ctx.gmres_solve_host(op, b, np.zeros(n), 1, 20, 1e-6)
    prec = ctx.gmres_solve_host(op, b, np.zeros(n), 1, 20, 1e-6, precond=pc)
    true_res = np.linalg.norm(b - op.apply_numpy(prec["x"]))
    assert true_res < 0.05 * np.linalg.norm(b - op.apply_numpy(plain["x"]))  # the preconditioner earns its keep
    assert relerr(prec["x"], x_ext) < 0.25 * relerr(plain["x"], x_ext)  # (the reference's GMRES converges slowly by design)
    # M^-1 applied to A*z gives back roughly z, exactly L*U*w = A*z on the level-scheduled sweeps' own terms
    z = np.random.default_rng(2).standard_normal(n)
    w = pc.solve_host(op.apply_numpy(z))
    assert relerr(w, z) < 0.2
    op.destroy()
    pc.destroy()


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_device_ilu_and_preconditioned_gmres(ctx, port, gilu, name):
    from algotoolkit_b200 import capi

    e = gilu["cases"][name]
    A = case_matrix(port, name, e)
    n = e["n"]
    pc = ctx.ilu_compute(n, n, A.col_ptr, A.row_idx, A.vals)
    try:
        Lf, Uf = pc.factors()
        assert digest(Lf) == e["L_sha256"] and digest(Uf) == e["U_sha256"]  # elimination: bit-exact in every mode
        b = port.u01_array(99, n) * 2.0 - 1.0
        LU = port.ilu_compute(A)
        want = port.ilu_solve(LU, b)
        assert relerr(pc.solve_host(b), want) < REL  # column-wavefront back substitution: rounding only
        rhs = port.spmv(A, port.u01_array(777, n))
        op = ctx.operator_from_csc(n, n, A.col_ptr, A.row_idx, A.vals)
        tree = ctx.gmres_solve_host(op, rhs, np.zeros(n), e["max_iter"], e["K"], e["tol"], precond=pc)
        assert tree["status"] == e["gmres_status"] and len(tree["resid"]) == len(e["gmres_resid"])
        assert np.allclose(tree["resid"], e["gmres_resid"], rtol=1e-9, atol=0.0)
        ctx.set_reduction(capi.REDUCE_SEQUENTIAL)
        try:
            assert digest(pc.solve_host(b)) == e["solve_sha256"]  # reference-ordered sums: bit-exact
            seq = ctx.gmres_solve_host(op, rhs, np.zeros(n), e["max_iter"], e["K"], e["tol"], precond=pc)
        finally:
            ctx.set_reduction(capi.REDUCE_TREE)
        assert digest(seq["x"]) == e["gmres_x_sha256"] and list(seq["resid"]) == e["gmres_resid"]
        assert seq["status"] == e["gmres_status"]
        # the events carry what the reference prints (KrylovUpdate lines included)
        kry = [int(v) for ev, v in seq["events"] if ev == capi.GMRES_KRYLOV_UPDATE]
        assert kry == [int(line.split(":")[1]) for line in e["gmres_log"].splitlines() if line.startswith("KrylovUpdate")]
    finally:
        pc.destroy()


@pytest.mark.gpu
def test_device_ilu_errors(ctx, port):
    with pytest.raises(RuntimeError, match="numerically singular"):  # tests/ILU_test.cpp:64-76
        ctx.ilu_compute(3, 3, *(lambda A: (A.col_ptr, A.row_idx, A.vals))(port.csc_from_triplets(3, 3, [0], [1], [1.0])))
    with pytest.raises(ValueError, match="square"):  # :57-62
        ctx.ilu_compute(3, 4, np.zeros(5, np.int32), np.zeros(0, np.int32), np.zeros(0))
    A = port.csc_from_triplets(3, 3, [0, 1, 2], [0, 1, 2], [2.0, 4.0, 8.0])
    pc = ctx.ilu_compute(3, 3, A.col_ptr, A.row_idx, A.vals)
    with pytest.raises(ValueError, match="Vector size"):  # :148-153
        pc.solve_host(np.ones(4))
    assert np.array_equal(pc.solve_host([2.0, 4.0, 8.0]), np.ones(3))
    op = ctx.operator_from_csc(3, 3, A.col_ptr, A.row_idx, A.vals)
    g = ctx.gmres_solve_host(op, [2.0, 4.0, 8.0], np.zeros(3), 5, 3, 1e-12, precond=pc)
    want = port.gmres_precond(A, [2.0, 4.0, 8.0], np.zeros(3), 5, 3, 1e-12)  # the i<j projection quirk: slow, not wrong
    assert relerr(g["x"], want["x"]) < REL and g["status"] == want["status"]
    pc.destroy()


@pytest.mark.gpu
def test_device_ilu_larger_dense_system(ctx, port):
    """n = 600 (several rows per thread in the sweeps, hundreds of elimination steps) against the C restatement."""
    n = 600
    rng = np.random.default_rng(21)
    D = np.where(rng.random((n, n)) < 0.02, rng.standard_normal((n, n)), 0.0)
    D[np.arange(n), np.arange(n)] = np.abs(D).sum(1) + 1.0
    c, r = np.nonzero(D.T)
    A = port.csc_from_triplets(n, n, r, c, D[r, c])
    from algotoolkit_b200 import capi

    pc = ctx.ilu_compute(n, n, A.col_ptr, A.row_idx, A.vals)
    LU = port.ilu_compute(A)
    Lw, Uw = port.ilu_factors(LU)
    Lf, Uf = pc.factors()
    assert np.array_equal(Lf, Lw) and np.array_equal(Uf, Uw)
    b = rng.standard_normal(n)
    want = port.ilu_solve(LU, b)
    assert relerr(pc.solve_host(b), want) < REL
    ctx.set_reduction(capi.REDUCE_SEQUENTIAL)
    try:
        assert np.array_equal(pc.solve_host(b), want)
    finally:
        ctx.set_reduction(capi.REDUCE_TREE)
    pc.destroy()


@pytest.mark.gpu
def test_two_live_preconditioners_of_different_sizes(ctx):
    """The dynamic shared-memory limit of the sweep kernel is per function, not per preconditioner: a small factorisation
    built AFTER a large one (n = 3400 needs 54 KB, above the 48 KB default) must not break solves with the large one."""
    rng = np.random.default_rng(5)

    def tridiag(n):
        rows = np.concatenate([np.arange(n), np.arange(1, n), np.arange(n - 1)])
        cols = np.concatenate([np.arange(n), np.arange(n - 1), np.arange(1, n)])
        vals = np.concatenate([np.full(n, 4.0), np.full(n - 1, -1.0), np.full(n - 1, -1.0)])
        order = np.lexsort((rows, cols))
        cp = np.zeros(n + 1, np.int32)
        np.add.at(cp, cols + 1, 1)
        return np.cumsum(cp).astype(np.int32), rows[order].astype(np.int32), vals[order]

    big_n, small_n = 3400, 40
    big = ctx.ilu_compute(big_n, big_n, *tridiag(big_n))
    small = ctx.ilu_compute(small_n, small_n, *tridiag(small_n))
    for pc, n in ((big, big_n), (small, small_n), (big, big_n)):
        b = rng.standard_normal(n)
        x = pc.solve_host(b)
        Ax = 4.0 * x
        Ax[1:] -= x[:-1]
        Ax[:-1] -= x[1:]
        assert relerr(Ax, b) < 1e-12  # the complete LU of a tridiagonal matrix: the sweeps solve the system
    big.destroy()
    small.destroy()


@pytest.mark.gpu
def test_fastsolver_gmres_enable_preconditioner(gilu, port, capsys):
    sys.path.insert(0, os.path.join(ROOT, "algotoolkit_b200"))
    import fastsolver as fs

    e = gilu["cases"]["random_37"]
    n = e["n"]
    A = fs.SparseMatrix(n, n)
    for r, c, v in zip(e["rows"], e["cols"], e["vals"]):
        A.addValue(r, c, v)
    A.finalize()
    Ac = case_matrix(port, "random_37", e)
    rhs = port.spmv(Ac, port.u01_array(777, n))
    fs.set_reduction("sequential")
    try:
        g = fs.GMRES()
        g.enablePreconditioner()
        x = fs.Vector(n)
        g.solve(A, fs.Vector(rhs), x, e["max_iter"], e["K"], e["tol"])
    finally:
        fs.set_reduction("tree")
    assert digest(x.to_numpy()) == e["gmres_x_sha256"]
    out = capsys.readouterr().out
    assert out.count("KrylovUpdate : 1") == 3 and "Reached maximum iterations without convergence." in out


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 64, 65, 97, 1056, 1100])
def test_device_ilu_sweep_block_boundaries(ctx, port, n):
    """Sizes around the 32-column blocks of the sweeps and the 1024-thread row ownership, dense factors."""
    from algotoolkit_b200 import capi

    rng = np.random.default_rng(n)
    dens = 1.0 if n <= 100 else 0.01
    D = np.where(rng.random((n, n)) < dens, rng.standard_normal((n, n)), 0.0)
    D[np.arange(n), np.arange(n)] = np.abs(D).sum(1) + 1.0
    c, r = np.nonzero(D.T)
    A = port.csc_from_triplets(n, n, r, c, D[r, c])
    pc = ctx.ilu_compute(n, n, A.col_ptr, A.row_idx, A.vals)
    LU = port.ilu_compute(A)
    Lw, Uw = port.ilu_factors(LU)
    Lf, Uf = pc.factors()
    assert np.array_equal(Lf, Lw) and np.array_equal(Uf, Uw)
    b = rng.standard_normal(n)
    want = port.ilu_solve(LU, b)
    assert relerr(pc.solve_host(b), want) < REL
    ctx.set_reduction(capi.REDUCE_SEQUENTIAL)
    try:
        assert np.array_equal(pc.solve_host(b), want)
    finally:
        ctx.set_reduction(capi.REDUCE_TREE)
    pc.destroy()
