"""ctypes front end for the two CPU checkers.  TEST INFRASTRUCTURE ONLY.

  * ``Port``  -- oracle/liboracle.so, the plain-C restatement (krylov_oracle.c)
  * ``Ref``   -- oracle/_ref/libatk_ref.so, the UNMODIFIED reference headers behind
                 oracle/ref_driver.cpp (prebuilt in the build container; it travels to the
                 GPU box with the repo snapshot because /root/reference does not exist there)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module.  The product (algotoolkit_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def build(force: bool = False) -> None:
    """Compile liboracle.so (always) and _ref/libatk_ref.so (when /root/reference exists)."""
    port = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("krylov_oracle.c", "krylov_oracle.h")]
    if force or not os.path.exists(port) or any(os.path.getmtime(s) > os.path.getmtime(port) for s in srcs):
        subprocess.run(["make", "-C", _HERE, "liboracle.so"], check=True, capture_output=True)
    ref = os.path.join(_HERE, "_ref", "libatk_ref.so")
    drv = os.path.join(_HERE, "ref_driver.cpp")
    if os.path.isdir("/root/reference/src") and (
        force or not os.path.exists(ref) or os.path.getmtime(drv) > os.path.getmtime(ref)
    ):
        subprocess.run(["make", "-C", _HERE, "ref"], check=True, capture_output=True)


class CscMatrix:
    """Finished CSC arrays (what SparseMatrixCSC::finalize leaves)."""

    def __init__(self, n_rows, n_cols, col_ptr, row_idx, vals):
        self.n_rows, self.n_cols = int(n_rows), int(n_cols)
        self.col_ptr, self.row_idx, self.vals = _i32(col_ptr), _i32(row_idx), _f64(vals)

    @property
    def nnz(self):
        return int(self.vals.shape[0])

    def to_scipy(self):
        import scipy.sparse as sp

        return sp.csc_matrix((self.vals, self.row_idx, self.col_ptr), shape=(self.n_rows, self.n_cols))


# =============================================================================== port
class Port:
    def __init__(self):
        build()
        L = C.CDLL(os.path.join(_HERE, "liboracle.so"))
        self.L = L
        L.orc_u01.restype = C.c_double
        L.orc_u01.argtypes = [C.c_uint64]
        L.orc_dot.restype = C.c_double
        L.orc_dot.argtypes = [C.c_long, _dp, _dp]
        L.orc_nrm2.restype = C.c_double
        L.orc_nrm2.argtypes = [C.c_long, _dp]
        L.orc_vec_scale.argtypes = [C.c_long, _dp, C.c_double, _dp]
        L.orc_vec_div.argtypes = [C.c_long, _dp, C.c_double, _dp]
        L.orc_vec_add.argtypes = [C.c_long, _dp, _dp, _dp]
        L.orc_vec_sub.argtypes = [C.c_long, _dp, _dp, _dp]
        L.orc_csc_from_triplets.restype = C.c_long
        L.orc_csc_from_triplets.argtypes = [C.c_int, C.c_int, C.c_long, _ip, _ip, _dp, _ip, _ip, _dp]
        L.orc_csc_get.argtypes = [C.c_int, C.c_int, _ip, _ip, _dp, C.c_int, C.c_int, _dp]
        L.orc_spmv_csc.argtypes = [C.c_int, C.c_int, _ip, _ip, _dp, _dp, _dp]
        L.orc_read_matrix_market.restype = C.c_long
        L.orc_read_matrix_market.argtypes = [C.c_char_p, _ip, _ip, C.POINTER(C.c_long), _ip, _ip, _dp]
        L.orc_poisson_coeffs.argtypes = [C.c_int] * 3 + [C.c_double] * 3 + [_dp] * 4
        L.orc_poisson_nnz.restype = C.c_long
        L.orc_poisson_nnz.argtypes = [C.c_int] * 3
        L.orc_poisson_csc.argtypes = [C.c_int] * 3 + [C.c_double] * 3 + [_ip, _ip, _dp]
        L.orc_poisson_apply.argtypes = [C.c_int] * 3 + [C.c_double] * 3 + [_dp, _dp]
        L.orc_poisson_rhs.argtypes = [C.c_int] * 3 + [C.c_double] * 3 + [C.c_int, C.c_uint64, _dp]
        L.orc_op27_nnz.restype = C.c_long
        L.orc_op27_nnz.argtypes = [C.c_int] * 3
        L.orc_op27_csc.argtypes = [C.c_int] * 3 + [_ip, _ip, _dp]
        L.orc_cg_solve.argtypes = [C.c_int, _ip, _ip, _dp, C.c_void_p, _dp, C.c_int, _dp, _dp, _dp, _ip, _dp, _dp,
                                   _dp, C.c_int]
        L.orc_gmres_solve.argtypes = [C.c_int, _ip, _ip, _dp, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_double, _dp,
                                      C.c_int, _ip, _ip, _ip]
        L.orc_gd_solve.argtypes = [C.c_int, _ip, _ip, _dp, _dp, _dp, C.c_int, C.c_double, _ip, _dp, _dp]
        L.orc_ilu_compute.argtypes = [C.c_int, _ip, _ip, _dp, _dp]
        L.orc_ilu_solve.argtypes = [C.c_int, _dp, _dp, _dp]
        L.orc_ilu_factors.argtypes = [C.c_int, _dp, _dp, _dp]
        L.orc_gmres_solve_precond.argtypes = L.orc_gmres_solve.argtypes
        L.orc_ilu0_compute.argtypes = L.orc_ilu_compute.argtypes
        L.orc_gmres_solve_precond_ilu0.argtypes = L.orc_gmres_solve.argtypes
        L.orc_arnoldi.argtypes = [C.c_int, _ip, _ip, _dp, _dp, C.c_int, _dp, C.c_double, _ip]

    # -- vectors
    def u01(self, s):
        return self.L.orc_u01(C.c_uint64(int(s) & 0xFFFFFFFFFFFFFFFF))

    def u01_array(self, seed, n):
        # vectorised splitmix64 finaliser, identical to orc_u01(seed + i)
        s = (np.arange(n, dtype=np.uint64) + np.uint64(seed)) + np.uint64(0x9E3779B97F4A7C15)
        s = (s ^ (s >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        s = (s ^ (s >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        s ^= s >> np.uint64(31)
        return (s >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)

    def dot(self, a, b):
        a, b = _f64(a), _f64(b)
        return self.L.orc_dot(a.size, _d(a), _d(b))

    def nrm2(self, a):
        a = _f64(a)
        return self.L.orc_nrm2(a.size, _d(a))

    def vec_scale(self, a, s):
        a = _f64(a)
        o = np.empty_like(a)
        self.L.orc_vec_scale(a.size, _d(a), s, _d(o))
        return o

    def vec_div(self, a, s):
        a = _f64(a)
        o = np.empty_like(a)
        if self.L.orc_vec_div(a.size, _d(a), s, _d(o)) != 0:
            raise RuntimeError("Division by zero.")
        return o

    def vec_add(self, a, b):
        a, b = _f64(a), _f64(b)
        o = np.empty_like(a)
        self.L.orc_vec_add(a.size, _d(a), _d(b), _d(o))
        return o

    def vec_sub(self, a, b):
        a, b = _f64(a), _f64(b)
        o = np.empty_like(a)
        self.L.orc_vec_sub(a.size, _d(a), _d(b), _d(o))
        return o

    # -- matrices
    def csc_from_triplets(self, n_rows, n_cols, rows, cols, vals):
        rows, cols, vals = _i32(rows), _i32(cols), _f64(vals)
        nt = rows.size
        cp = np.zeros(n_cols + 1, np.int32)
        ri = np.zeros(max(nt, 1), np.int32)
        vv = np.zeros(max(nt, 1), np.float64)
        nnz = self.L.orc_csc_from_triplets(n_rows, n_cols, nt, _i(rows), _i(cols), _d(vals), _i(cp), _i(ri), _d(vv))
        if nnz < 0:
            raise IndexError("Index out of range")
        return CscMatrix(n_rows, n_cols, cp, ri[:nnz].copy(), vv[:nnz].copy())

    def csc_get(self, A, i, j):
        out = C.c_double(0.0)
        if self.L.orc_csc_get(A.n_rows, A.n_cols, _i(A.col_ptr), _i(A.row_idx), _d(A.vals), i, j, C.byref(out)) != 0:
            raise IndexError("Index out of range")
        return out.value

    def spmv(self, A, x):
        x = _f64(x)
        if x.size != A.n_cols:
            raise ValueError("Dimension mismatch")
        y = np.empty(A.n_rows, np.float64)
        self.L.orc_spmv_csc(A.n_rows, A.n_cols, _i(A.col_ptr), _i(A.row_idx), _d(A.vals), _d(x), _d(y))
        return y

    def read_matrix_market(self, path):
        nr, nc, ne = C.c_int(), C.c_int(), C.c_long()
        rc = self.L.orc_read_matrix_market(path.encode(), C.byref(nr), C.byref(nc), C.byref(ne), None, None, None)
        if rc < 0:
            raise FileNotFoundError(path)
        cp = np.zeros(nc.value + 1, np.int32)
        ri = np.zeros(max(ne.value, 1), np.int32)
        vv = np.zeros(max(ne.value, 1), np.float64)
        nnz = self.L.orc_read_matrix_market(path.encode(), C.byref(nr), C.byref(nc), C.byref(ne), _i(cp), _i(ri), _d(vv))
        if nnz < 0:
            raise IndexError("Index out of range")
        return CscMatrix(nr.value, nc.value, cp, ri[:nnz].copy(), vv[:nnz].copy())

    # -- Poisson
    def poisson_coeffs(self, nx, ny, nz, lx=1.0, ly=1.0, lz=1.0):
        c = (C.c_double * 4)()
        self.L.orc_poisson_coeffs(nx, ny, nz, lx, ly, lz, *[C.cast(C.byref(c, 8 * t), _dp) for t in range(4)])
        return tuple(c)

    def poisson_csc(self, nx, ny, nz, lx=1.0, ly=1.0, lz=1.0):
        n = nx * ny * nz
        nnz = self.L.orc_poisson_nnz(nx, ny, nz)
        cp = np.zeros(n + 1, np.int32)
        ri = np.zeros(nnz, np.int32)
        vv = np.zeros(nnz, np.float64)
        self.L.orc_poisson_csc(nx, ny, nz, lx, ly, lz, _i(cp), _i(ri), _d(vv))
        return CscMatrix(n, n, cp, ri, vv)

    def poisson_apply(self, nx, ny, nz, cx, cy, cz, x):
        x = _f64(x)
        y = np.empty_like(x)
        self.L.orc_poisson_apply(nx, ny, nz, cx, cy, cz, _d(x), _d(y))
        return y

    def poisson_rhs(self, nx, ny, nz, lx=1.0, ly=1.0, lz=1.0, rhs_kind=1, seed=12345):
        b = np.empty(nx * ny * nz, np.float64)
        self.L.orc_poisson_rhs(nx, ny, nz, lx, ly, lz, rhs_kind, C.c_uint64(seed), _d(b))
        return b

    def op27_csc(self, nx, ny, nz):
        n = nx * ny * nz
        nnz = self.L.orc_op27_nnz(nx, ny, nz)
        cp = np.zeros(n + 1, np.int32)
        ri = np.zeros(nnz, np.int32)
        vv = np.zeros(nnz, np.float64)
        self.L.orc_op27_csc(nx, ny, nz, _i(cp), _i(ri), _d(vv))
        return CscMatrix(n, n, cp, ri, vv)

    # -- solvers
    def cg(self, A, b, max_iter, state=None, stencil=None, trace=False):
        """A: CscMatrix or None (then stencil=(nx,ny,nz,cx,cy,cz)).  Returns dict."""
        b = _f64(b)
        n = b.size
        if state is None:
            x, r, p = np.zeros(n), b.copy(), b.copy()
        else:
            x, r, p = (_f64(s).copy() for s in state)
        iters, last = C.c_int(0), C.c_double(0.0)
        cap = max_iter if trace else 0
        tp = np.full(max(cap, 1), np.nan)
        tr = np.full(max(cap, 1), np.nan)

        class St(C.Structure):
            _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("cx", C.c_double), ("cy", C.c_double),
                        ("cz", C.c_double)]

        if A is not None:
            self.L.orc_cg_solve(n, _i(A.col_ptr), _i(A.row_idx), _d(A.vals), None, _d(b), max_iter, _d(x), _d(r),
                                _d(p), C.byref(iters), C.byref(last), _d(tp), _d(tr), cap)
        else:
            st = St(*stencil)
            self.L.orc_cg_solve(n, None, None, None, C.cast(C.byref(st), C.c_void_p), _d(b), max_iter, _d(x), _d(r),
                                _d(p), C.byref(iters), C.byref(last), _d(tp), _d(tr), cap)
        return dict(x=x, r=r, p=p, iters=iters.value, last_pAp=last.value, pAp_trace=tp[:cap], rs_trace=tr[:cap])

    def gmres(self, A, b, x0, max_iter, krylov_dim, tol):
        b, x = _f64(b), _f64(x0).copy()
        cap = max_iter + 2
        resid = np.zeros(cap)
        nres, status, kry = C.c_int(0), C.c_int(0), C.c_int(0)
        rc = self.L.orc_gmres_solve(A.n_rows, _i(A.col_ptr), _i(A.row_idx), _d(A.vals), _d(b), _d(x), x.size, max_iter,
                                    krylov_dim, tol, _d(resid), cap, C.byref(nres), C.byref(status), C.byref(kry))
        if rc == 1:
            raise ValueError("invalid argument")
        return dict(x=x, resid=resid[: nres.value].copy(), status=status.value, rc=rc, kry_update=kry.value)

    # -- ILU preconditioner (dense restatement) and the preconditioned GMRES branch
    def ilu_compute(self, A):
        """Eliminated matrix as an (n, n) array indexed [row, col]; RuntimeError when the reference would throw."""
        n = A.n_rows
        if A.n_rows != A.n_cols:
            raise ValueError("Matrix must be square for ILU factorization")
        LU = np.zeros((n, n))  # column-major for the C side: LU[c] is column c
        rc = self.L.orc_ilu_compute(n, _i(A.col_ptr), _i(A.row_idx), _d(A.vals), _d(LU))
        if rc == 2:
            raise RuntimeError("Matrix is numerically singular")
        return LU.T.copy()

    def ilu0_compute(self, A):
        """ILU(0): the reference's elimination restricted to the stored positions of A (no fill-in); same layout as ilu_compute."""
        n = A.n_rows
        if A.n_rows != A.n_cols:
            raise ValueError("Matrix must be square for ILU factorization")
        LU = np.zeros((n, n))
        rc = self.L.orc_ilu0_compute(n, _i(A.col_ptr), _i(A.row_idx), _d(A.vals), _d(LU))
        if rc == 2:
            raise RuntimeError("Matrix is numerically singular")
        return LU.T.copy()

    def gmres_precond_ilu0(self, A, b, x0, max_iter, krylov_dim, tol):
        b, x = _f64(b), _f64(x0).copy()
        cap = max_iter + 2
        resid = np.zeros(cap)
        nres, status, kry = C.c_int(0), C.c_int(0), C.c_int(0)
        rc = self.L.orc_gmres_solve_precond_ilu0(A.n_rows, _i(A.col_ptr), _i(A.row_idx), _d(A.vals), _d(b), _d(x), x.size,
                                                 max_iter, krylov_dim, tol, _d(resid), cap, C.byref(nres), C.byref(status),
                                                 C.byref(kry))
        if rc == 1:
            raise ValueError("invalid argument")
        return dict(x=x, resid=resid[: nres.value].copy(), status=status.value, rc=rc, kry_update=kry.value)

    def ilu_solve(self, LU, b):
        LUc = np.ascontiguousarray(np.asarray(LU, np.float64).T)
        b = _f64(b)
        x = np.zeros_like(b)
        self.L.orc_ilu_solve(b.size, _d(LUc), _d(b), _d(x))
        return x

    def ilu_factors(self, LU):
        LUc = np.ascontiguousarray(np.asarray(LU, np.float64).T)
        n = LUc.shape[0]
        Lf, Uf = np.zeros((n, n)), np.zeros((n, n))
        self.L.orc_ilu_factors(n, _d(LUc), _d(Lf), _d(Uf))
        return Lf.T.copy(), Uf.T.copy()

    def gmres_precond(self, A, b, x0, max_iter, krylov_dim, tol):
        b, x = _f64(b), _f64(x0).copy()
        cap = max_iter + 2
        resid = np.zeros(cap)
        nres, status, kry = C.c_int(0), C.c_int(0), C.c_int(0)
        rc = self.L.orc_gmres_solve_precond(A.n_rows, _i(A.col_ptr), _i(A.row_idx), _d(A.vals), _d(b), _d(x), x.size,
                                            max_iter, krylov_dim, tol, _d(resid), cap, C.byref(nres), C.byref(status),
                                            C.byref(kry))
        if rc == 1:
            raise ValueError("invalid argument")
        return dict(x=x, resid=resid[: nres.value].copy(), status=status.value, rc=rc, kry_update=kry.value)

    def gd(self, A, b, x0, max_iter, tol):
        b, x = _f64(b), _f64(x0).copy()
        it, rn, cn = C.c_int(), C.c_double(), C.c_double()
        self.L.orc_gd_solve(A.n_rows, _i(A.col_ptr), _i(A.row_idx), _d(A.vals), _d(b), _d(x), max_iter, tol, C.byref(it),
                            C.byref(rn), C.byref(cn))
        return dict(x=x, iters=it.value, residual_norm=rn.value, change_norm=cn.value)

    def arnoldi(self, A, q0, m, tol):
        """Q as (m, n) array (row i = Q[i]), H as (m, m-1) array."""
        n = A.n_rows
        Q = np.zeros((m, n))
        Q[0] = _f64(q0)
        H = np.zeros((m - 1, m))  # column-major m x (m-1): column c at H[c]
        bd = C.c_int()
        self.L.orc_arnoldi(n, _i(A.col_ptr), _i(A.row_idx), _d(A.vals), _d(Q), m, _d(H), tol, C.byref(bd))
        return dict(Q=Q, H=H.T.copy(), breakdown=bd.value)


# =============================================================================== reference
def ref_available() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libatk_ref.so"))


class Ref:
    """The unmodified reference (oracle/_ref/libatk_ref.so)."""

    def __init__(self):
        build()
        L = C.CDLL(os.path.join(_HERE, "_ref", "libatk_ref.so"))
        self.L = L
        L.ref_u01.restype = C.c_double
        L.ref_u01.argtypes = [C.c_uint64]
        L.ref_dot.restype = C.c_double
        L.ref_dot.argtypes = [C.c_int, _dp, _dp]
        L.ref_nrm2.restype = C.c_double
        L.ref_nrm2.argtypes = [C.c_int, _dp]
        L.ref_vec_scale.argtypes = [C.c_int, _dp, C.c_double, _dp]
        L.ref_vec_div.argtypes = [C.c_int, _dp, C.c_double, _dp]
        L.ref_vec_add.argtypes = [C.c_int, _dp, _dp, _dp]
        L.ref_vec_sub.argtypes = [C.c_int, _dp, _dp, _dp]
        for f in ("ref_mat_from_triplets", "ref_mat_from_csc", "ref_mat_read_mm", "ref_mat_poisson"):
            getattr(L, f).restype = C.c_void_p
        L.ref_mat_from_triplets.argtypes = [C.c_int, C.c_int, C.c_long, _ip, _ip, _dp]
        L.ref_mat_from_csc.argtypes = [C.c_int, C.c_int, C.c_long, _ip, _ip, _dp]
        L.ref_mat_read_mm.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
        L.ref_mat_poisson.argtypes = [C.c_int] * 3 + [C.c_double] * 3
        L.ref_mat_free.argtypes = [C.c_void_p]
        L.ref_mat_dims.argtypes = [C.c_void_p, _ip, _ip, C.POINTER(C.c_long)]
        L.ref_mat_export.argtypes = [C.c_void_p, _ip, _ip, _dp]
        L.ref_mat_get.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp]
        L.ref_mat_spmv.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
        L.ref_cg_solve.argtypes = [C.c_void_p, _dp, C.c_int, C.c_double, _dp, _dp, _dp]
        L.ref_cg_trace.argtypes = [C.c_void_p, _dp, C.c_int, _dp, _dp, _dp, _ip, _dp, _dp, _dp, C.c_int]
        L.ref_gmres_solve.argtypes = [C.c_void_p, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int,
                                      C.c_char_p, C.c_int]
        L.ref_poisson_solve.argtypes = [C.c_int] * 3 + [C.c_double] * 3 + [C.c_int, C.c_uint64, C.c_int, C.c_double,
                                                                           _dp, _dp, _dp, _dp]
        L.ref_poisson_rhs.argtypes = [C.c_int] * 3 + [C.c_double] * 3 + [C.c_int, C.c_uint64, _dp]
        L.ref_gd_solve.argtypes = [C.c_void_p, _dp, _dp, C.c_int, C.c_double]
        L.ref_arnoldi.argtypes = [C.c_void_p, _dp, C.c_int, _dp, C.c_double, C.c_char_p, C.c_int]

    # -- vectors
    def dot(self, a, b):
        a, b = _f64(a), _f64(b)
        return self.L.ref_dot(a.size, _d(a), _d(b))

    def nrm2(self, a):
        a = _f64(a)
        return self.L.ref_nrm2(a.size, _d(a))

    def _ew(self, fn, a, *rest):
        a = _f64(a)
        o = np.empty_like(a)
        args = [_d(_f64(r)) if isinstance(r, np.ndarray) else r for r in rest]
        rc = fn(a.size, _d(a), *args, _d(o))
        return o, rc

    def vec_scale(self, a, s):
        return self._ew(self.L.ref_vec_scale, a, s)[0]

    def vec_div(self, a, s):
        o, rc = self._ew(self.L.ref_vec_div, a, s)
        if rc != 0:
            raise RuntimeError("Division by zero.")
        return o

    def vec_add(self, a, b):
        return self._ew(self.L.ref_vec_add, a, _f64(b))[0]

    def vec_sub(self, a, b):
        return self._ew(self.L.ref_vec_sub, a, _f64(b))[0]

    # -- matrices (handles)
    class Mat:
        def __init__(self, ref, h):
            if not h:
                raise IndexError("reference matrix construction failed (out of range)")
            self.ref, self.h = ref, C.c_void_p(h)
            nr, nc, nnz = C.c_int(), C.c_int(), C.c_long()
            ref.L.ref_mat_dims(self.h, C.byref(nr), C.byref(nc), C.byref(nnz))
            self.n_rows, self.n_cols, self.nnz = nr.value, nc.value, nnz.value

        def __del__(self):
            try:
                self.ref.L.ref_mat_free(self.h)
            except Exception:
                pass

        def export(self):
            cp = np.zeros(self.n_cols + 1, np.int32)
            ri = np.zeros(max(self.nnz, 1), np.int32)
            vv = np.zeros(max(self.nnz, 1), np.float64)
            self.ref.L.ref_mat_export(self.h, _i(cp), _i(ri), _d(vv))
            return CscMatrix(self.n_rows, self.n_cols, cp, ri[: self.nnz].copy(), vv[: self.nnz].copy())

        def get(self, i, j):
            out = C.c_double()
            if self.ref.L.ref_mat_get(self.h, i, j, C.byref(out)) != 0:
                raise IndexError("Index out of range")
            return out.value

        def spmv(self, x):
            x = _f64(x)
            y = np.empty(self.n_rows)
            if self.ref.L.ref_mat_spmv(self.h, x.size, _d(x), _d(y)) != 0:
                raise ValueError("Dimension mismatch")
            return y

    def mat_from_triplets(self, n_rows, n_cols, rows, cols, vals):
        rows, cols, vals = _i32(rows), _i32(cols), _f64(vals)
        return Ref.Mat(self, self.L.ref_mat_from_triplets(n_rows, n_cols, rows.size, _i(rows), _i(cols), _d(vals)))

    def mat_from_csc(self, A: CscMatrix):
        return Ref.Mat(self, self.L.ref_mat_from_csc(A.n_rows, A.n_cols, A.nnz, _i(A.col_ptr), _i(A.row_idx), _d(A.vals)))

    def mat_read_mm(self, path):
        log = C.create_string_buffer(4096)
        m = Ref.Mat(self, self.L.ref_mat_read_mm(path.encode(), log, 4096))
        m.log = log.value.decode()
        return m

    def mat_poisson(self, nx, ny, nz, lx=1.0, ly=1.0, lz=1.0):
        return Ref.Mat(self, self.L.ref_mat_poisson(nx, ny, nz, lx, ly, lz))

    # -- solvers
    def cg_solve(self, M, b, max_iter, tol=1e-6):
        b = _f64(b)
        x, r, p = np.empty_like(b), np.empty_like(b), np.empty_like(b)
        self.L.ref_cg_solve(M.h, _d(b), max_iter, tol, _d(x), _d(r), _d(p))
        return dict(x=x, r=r, p=p)

    def cg_trace(self, M, b, max_iter):
        b = _f64(b)
        x, r, p = np.empty_like(b), np.empty_like(b), np.empty_like(b)
        iters, last = C.c_int(), C.c_double()
        tp, tr = np.full(max(max_iter, 1), np.nan), np.full(max(max_iter, 1), np.nan)
        self.L.ref_cg_trace(M.h, _d(b), max_iter, _d(x), _d(r), _d(p), C.byref(iters), C.byref(last), _d(tp), _d(tr),
                            max_iter)
        return dict(x=x, r=r, p=p, iters=iters.value, last_pAp=last.value, pAp_trace=tp[:max_iter],
                    rs_trace=tr[:max_iter])

    def gmres(self, M, b, x0, max_iter, krylov_dim, tol, dense_h=True, precision=17):
        b, x = _f64(b), _f64(x0).copy()
        cap = 1 << 20
        log = C.create_string_buffer(cap)
        rc = self.L.ref_gmres_solve(M.h, _d(b), _d(x), x.size, max_iter, krylov_dim, tol, int(dense_h), precision,
                                    log, cap)
        text = log.value.decode()
        if rc == 1:
            raise ValueError("invalid argument")
        return dict(x=x, rc=rc, log=text, resid=parse_gmres_log(text)[0], status=parse_gmres_log(text)[1])

    def gmres_precond(self, M, b, x0, max_iter, krylov_dim, tol, precision=17):
        b, x = _f64(b), _f64(x0).copy()
        cap = 1 << 20
        log = C.create_string_buffer(cap)
        self.L.ref_gmres_solve_precond.argtypes = [C.c_void_p, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int,
                                                   C.c_char_p, C.c_int]
        rc = self.L.ref_gmres_solve_precond(M.h, _d(b), _d(x), x.size, max_iter, krylov_dim, tol, precision, log, cap)
        text = log.value.decode()
        if rc == 1:
            raise ValueError("invalid argument")
        return dict(x=x, rc=rc, log=text, resid=parse_gmres_log(text)[0], status=parse_gmres_log(text)[1])

    def ilu(self, M):
        """ILUPreconditioner<double>::compute on M; returns dict(L=, U= as dense [row, col] arrays, solve=callable)."""
        L = self.L
        L.ref_ilu_compute.restype = C.c_void_p
        L.ref_ilu_compute.argtypes = [C.c_void_p, _ip]
        L.ref_ilu_factor.restype = C.c_void_p
        L.ref_ilu_factor.argtypes = [C.c_void_p, C.c_int]
        L.ref_ilu_solve.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
        L.ref_ilu_free.argtypes = [C.c_void_p]
        rc = C.c_int(0)
        h = L.ref_ilu_compute(M.h, C.byref(rc))
        if rc.value == 1:
            raise ValueError("Matrix must be square for ILU factorization")
        if rc.value == 2:
            raise RuntimeError("Matrix is numerically singular")
        out = {}
        for which, key in ((0, "L"), (1, "U")):
            F = Ref.Mat(self, L.ref_ilu_factor(h, which))
            out[key] = F.export().to_scipy().toarray()

        def solve(b, _h=h):
            b = _f64(b)
            x = np.zeros_like(b)
            r = L.ref_ilu_solve(_h, b.size, _d(b), _d(x))
            if r == 1:
                raise ValueError("Vector size does not match matrix size")
            return x

        out["solve"] = solve
        out["free"] = lambda _h=h: L.ref_ilu_free(_h)
        return out

    def gd(self, M, b, x0, max_iter, tol):
        b, x = _f64(b), _f64(x0).copy()
        self.L.ref_gd_solve(M.h, _d(b), _d(x), max_iter, tol)
        return dict(x=x)

    def arnoldi(self, M, q0, m, tol):
        n = M.n_rows
        Q = np.zeros((m, n))
        Q[0] = _f64(q0)
        H = np.zeros((m - 1, m))
        log = C.create_string_buffer(4096)
        self.L.ref_arnoldi(M.h, _d(Q), m, _d(H), tol, log, 4096)
        return dict(Q=Q, H=H.T.copy(), log=log.value.decode())

    def poisson_solve(self, nx, ny, nz, lx=1.0, ly=1.0, lz=1.0, rhs_kind=0, seed=12345, max_iter=1000, tol=1e-6):
        n = nx * ny * nz
        b, u = np.empty(n), np.empty(n)
        tb, ts = C.c_double(), C.c_double()
        self.L.ref_poisson_solve(nx, ny, nz, lx, ly, lz, rhs_kind, C.c_uint64(seed), max_iter, tol, _d(b), _d(u),
                                 C.byref(tb), C.byref(ts))
        return dict(b=b, u=u, seconds_build=tb.value, seconds_solve=ts.value)

    def poisson_rhs(self, nx, ny, nz, lx=1.0, ly=1.0, lz=1.0, rhs_kind=1, seed=12345):
        b = np.empty(nx * ny * nz)
        self.L.ref_poisson_rhs(nx, ny, nz, lx, ly, lz, rhs_kind, C.c_uint64(seed), _d(b))
        return b


def parse_gmres_log(text: str):
    """Residual ladder + status from the text GMRES::solve prints (GMRES.hpp:48,50,105,108,114).
    status: 0 reached max iterations, 1 converged after restart, 2 initial guess good enough."""
    resid, status = [], 0
    for line in text.splitlines():
        if line.startswith("Initial residual norm: "):
            resid.append(float(line.split(": ")[1]))
        elif line.startswith("Residual norm after restart: "):
            resid.append(float(line.split(": ")[1].strip()))
        elif line.startswith("Converged after restart"):
            status = 1
        elif line.startswith("Initial guess is good enough"):
            status = 2
    return np.array(resid), status
