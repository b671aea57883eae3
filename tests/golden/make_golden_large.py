"""Full-size parity fixtures for BASELINE config 3 (Poisson CG to the |pAp| < 1e-12 exit, RHS-R seed 12345) from the
UNMODIFIED reference (oracle/_ref): 128^3 and 256^3.  The reference does not report its iteration count; the C port
does, and the fixture is only written when the port's solution is bit-identical to the reference's.

    python tests/golden/make_golden_large.py [sizes...]      (256^3: ~7 GB and ~15 minutes of one core)
Writes tests/golden/golden_large.json: iterations, ||x||, sha256(x), ||b - A x||, and 512 sampled entries of x."""
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


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [128, 256]
    P, R = Port(), Ref()
    path = os.path.join(HERE, "golden_large.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
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
