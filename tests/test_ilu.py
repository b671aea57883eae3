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


# ------------------------------------------------------------------------------------------------ device (GPU)
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
