"""Golden vectors for the ILU-preconditioned GMRES branch, produced by the UNMODIFIED reference
(oracle/_ref/libatk_ref.so: ILUPreconditioner<double, SparseMatrixCSC<double>> and GMRES::enablePreconditioner()).

Run in the build container (needs /root/reference):   python tests/golden/make_golden_ilu.py
Writes tests/golden/golden_ilu.json.  Small random matrices are stored in the file itself (triplets), so the fixtures do
not depend on a random generator's stream."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from algotoolkit_b200.mmio import read_matrix_market_csc  # noqa: E402
from oracle.pyoracle import CscMatrix, Port, Ref  # noqa: E402


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.float64).tobytes()).hexdigest()


def main():
    P, R = Port(), Ref()
    cases = {}

    def add(name, A, K, max_iter, tol, store_triplets=False):
        n = A.n_rows
        M = R.mat_from_csc(A)
        f = R.ilu(M)
        b = P.u01_array(99, n) * 2.0 - 1.0
        xs = f["solve"](b)
        x_ext = P.u01_array(777, n)
        rhs = M.spmv(x_ext)
        g = R.gmres_precond(M, rhs, np.zeros(n), max_iter, K, tol)
        e = dict(n=n, K=K, max_iter=max_iter, tol=tol, L_sha256=digest(f["L"]), U_sha256=digest(f["U"]),
                 L_nnz=int(np.count_nonzero(f["L"])), U_nnz=int(np.count_nonzero(f["U"])), solve_sha256=digest(xs),
                 solve_x=[float(v) for v in xs[:8]], gmres_x_sha256=digest(g["x"]), gmres_resid=[float(v) for v in g["resid"]],
                 gmres_status=int(g["status"]), gmres_rc=int(g["rc"]), gmres_log=g["log"][:2000])
        if store_triplets:
            S = A.to_scipy().tocoo()
            e["rows"], e["cols"], e["vals"] = S.row.tolist(), S.col.tolist(), [float(v) for v in S.data]
        # the C restatement must agree bit for bit before anything is written
        LU = P.ilu_compute(A)
        Lp, Up = P.ilu_factors(LU)
        assert np.array_equal(Lp, f["L"]) and np.array_equal(Up, f["U"]), name
        assert np.array_equal(P.ilu_solve(LU, b), xs), name
        gp = P.gmres_precond(A, rhs, np.zeros(n), max_iter, K, tol)
        assert np.array_equal(gp["x"], g["x"]) and gp["status"] == g["status"], name
        f["free"]()
        cases[name] = e

    # tests/ILU_test.cpp:11-18 createTestMatrix, :81-105 SolveSystem (x_true = 1,2,3)
    A3 = P.csc_from_triplets(3, 3, [0, 0, 1, 1, 1, 2, 2], [0, 1, 0, 1, 2, 1, 2], [4.0, -1.0, -1.0, 4.0, -1.0, -1.0, 4.0])
    add("ilu_test_3x3", A3, 3, 3, 1e-10, store_triplets=True)
    for tag, sym in (("bcsstk01_as_loaded", False), ("bcsstk01_symmetric", True)):
        d = read_matrix_market_csc(os.path.join(HERE, "data", "bcsstk01.mtx"), expand_symmetric=sym)
        add(tag, CscMatrix(d["n_rows"], d["n_cols"], d["col_ptr"], d["row_idx"], d["vals"]), 20, 3, 1e-10)
    rng = np.random.default_rng(3)
    for n in (10, 37, 80):
        mask = rng.random((n, n)) < 0.15
        D = np.where(mask, rng.standard_normal((n, n)), 0.0)
        D[np.arange(n), np.arange(n)] = np.abs(D).sum(1) + 1.0
        c, r = np.nonzero(D.T)
        add(f"random_{n}", P.csc_from_triplets(n, n, r, c, D[r, c]), min(n, 20), 3, 1e-10, store_triplets=True)
    # entries straddling the 1e-12 drop rules (old value kept when an update is dropped; tiny multipliers skipped)
    n = 12
    D = np.where(rng.random((n, n)) < 0.5, rng.standard_normal((n, n)) * 10.0 ** rng.integers(-14, 2, (n, n)), 0.0)
    D[np.arange(n), np.arange(n)] = 1.0 + rng.random(n)
    c, r = np.nonzero(D.T)
    add("drop_rules_12", P.csc_from_triplets(n, n, r, c, D[r, c]), 12, 3, 1e-10, store_triplets=True)
    add("poisson_4", P.poisson_csc(4, 4, 4), 30, 3, 1e-10)
    add("op27_5", P.op27_csc(5, 5, 5), 25, 3, 1e-10)

    # error behaviour: tests/ILU_test.cpp:58-76
    sing = P.csc_from_triplets(3, 3, [0], [1], [1.0])
    try:
        R.ilu(R.mat_from_csc(sing))
        singular = "no error"
    except RuntimeError as e:
        singular = str(e)
    out = dict(cases=cases, singular_3x3=singular)
    path = os.path.join(HERE, "golden_ilu.json")
    with open(path, "w") as fjson:
        json.dump(out, fjson, indent=1)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
