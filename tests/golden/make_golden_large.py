"""Full-size parity fixtures for BASELINE config 3 (Poisson CG to the |pAp| < 1e-12 exit, RHS-R seed 12345) from the
UNMODIFIED reference (oracle/_ref): 128^3 and 256^3.  The reference does not report its iteration count; the C port
does, and the fixture is only written when the port's solution is bit-identical to the reference's.

    python tests/golden/make_golden_large.py [sizes...]      (256^3: ~7 GB and ~25 minutes of one core)
    python tests/golden/make_golden_large.py gmres           (BASELINE config 5 operator: 24^3 K=40, 32^3 K=60, 64^3 K=300)
Writes tests/golden/golden_large.json: iterations, ||x||, sha256(x), ||b - A x||, and 512 sampled entries of x; for the
GMRES cases the residual ladder printed by the reference (shipped sparse-H instantiation), ||x||, sha256(x), samples."""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from algotoolkit_b200.synthetic import u01_array  # noqa: E402
from oracle.pyoracle import Port, Ref  # noqa: E402


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.float64).tobytes()).hexdigest()


def gmres_cases(P, R, out, path):
    for n, K, iters in ((24, 40, 2), (32, 60, 2), (64, 300, 1)):
        A = P.op27_csc(n, n, n)
        M = R.mat_from_csc(A)
        x_ext = P.u01_array(777, A.n_cols)
        b = M.spmv(x_ext)
        t0 = time.time()
        gp = P.gmres(A, b, np.zeros(A.n_rows), iters, K, 1e-6)
        t1 = time.time()
        gr = R.gmres(M, b, np.zeros(A.n_rows), iters, K, 1e-6, dense_h=False)  # the instantiation the reference ships
        t2 = time.time()
        assert np.array_equal(gp["x"], gr["x"]) and list(gp["resid"]) == list(gr["resid"]), "port and reference differ"
        x = gr["x"]
        idx = (u01_array(4242, 0, 512) * A.n_rows).astype(np.int64)
        out[f"gmres_op27_{n}_K{K}"] = dict(n=n, K=K, max_iter=iters, rows=A.n_rows, nnz=A.nnz, b_sha256=digest(b),
                                           resid=[float(v) for v in gr["resid"]], status=int(gr["status"]),
                                           x_norm=float(np.linalg.norm(x)), x_sha256=digest(x),
                                           sample_index=[int(i) for i in idx], sample_x=[float(v) for v in x[idx]],
                                           x_err_vs_ext=float(np.linalg.norm(x - x_ext) / np.linalg.norm(x_ext)),
                                           port_seconds=t1 - t0, reference_seconds=t2 - t1)
        print("gmres", n, K, out[f"gmres_op27_{n}_K{K}"]["resid"], "port %.0f s, reference %.0f s" % (t1 - t0, t2 - t1), flush=True)
        with open(path, "w") as f:
            json.dump(out, f, indent=1)


def main():
    P, R = Port(), Ref()
    path = os.path.join(HERE, "golden_large.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    if sys.argv[1:] == ["gmres"]:
        return gmres_cases(P, R, out, path)
    sizes = [int(a) for a in sys.argv[1:]] or [128, 256]
    for n in sizes:
        t0 = time.time()
        cx, cy, cz, _ = P.poisson_coeffs(n, n, n)
        b = P.poisson_rhs(n, n, n, rhs_kind=1, seed=12345)
        port = P.cg(None, b, 100000, stencil=(n, n, n, cx, cy, cz))
        t1 = time.time()
        ref = R.poisson_solve(n, n, n, rhs_kind=1, seed=12345, max_iter=100000)
        t2 = time.time()
        assert np.array_equal(ref["b"], b), "RHS differs from the reference's"
        assert np.array_equal(ref["u"], port["x"]), "port and reference solutions differ"
        x = port["x"]
        idx = (u01_array(4242, 0, 512) * (n ** 3)).astype(np.int64)
        res = b - P.poisson_apply(n, n, n, cx, cy, cz, x)
        out[str(n)] = dict(n=n, iterations=int(port["iters"]), last_pAp=float(port["last_pAp"]), x_norm=float(np.linalg.norm(x)),
                           x_max=float(np.abs(x).max()), x_sha256=digest(x), b_norm=float(np.linalg.norm(b)),
                           true_residual=float(np.linalg.norm(res)), sample_index=[int(i) for i in idx],
                           sample_x=[float(v) for v in x[idx]], port_seconds=t1 - t0, reference_seconds=t2 - t1,
                           reference_solve_seconds=ref["seconds_solve"], reference_build_seconds=ref["seconds_build"])
        print(n, out[str(n)]["iterations"], out[str(n)]["x_norm"], "port %.0f s, reference %.0f s" % (t1 - t0, t2 - t1), flush=True)
        with open(path, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
