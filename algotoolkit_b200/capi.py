"""ctypes binding of libatk_cuda.so (include/atk_cuda.h) for tests, bench.py and Python-side drivers.

This is plumbing: it owns no algorithm.  Every compute call goes to the CUDA library; if the library is missing or
no CUDA device is present the call raises - there is no CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# ATK_LIB=checked loads the build with device-side assertions (python -m algotoolkit_b200.build --checked)
# ATK_LIB=profile loads the build with reduction-station stamps (python -m algotoolkit_b200.build --profile)
LIB_PATH = os.path.join(_HERE, {"checked": "libatk_cuda_checked.so", "profile": "libatk_cuda_profile.so"}.get(
    os.environ.get("ATK_LIB", ""), "libatk_cuda.so"))
if os.environ.get("ATK_LIB_PATH"):  # A/B runs against another build of the same ABI
    LIB_PATH = os.environ["ATK_LIB_PATH"]

ATK_OK, ATK_ERR_INVALID_ARGUMENT, ATK_ERR_RUNTIME, ATK_ERR_OUT_OF_RANGE, ATK_ERR_CUDA, ATK_ERR_NCCL, ATK_ERR_NOMEM = range(7)
REDUCE_TREE, REDUCE_SEQUENTIAL = 0, 1
FORMAT_SELL, FORMAT_CSR_VECTOR = 0, 1
GMRES_MGS, GMRES_MGS_BATCHED = 0, 1
(GMRES_INITIAL_RESIDUAL, GMRES_INITIAL_GOOD, GMRES_KRYLOV_UPDATE, GMRES_RESTART_RESIDUAL, GMRES_CONVERGED,
 GMRES_MAX_ITER) = range(6)  # atk_gmres_event

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_vp = C.c_void_p


class AtkError(Exception):
    def __init__(self, code, msg):
        super().__init__(f"[atk status {code}] {msg}")
        self.code = code


class AtkCudaError(AtkError, RuntimeError):
    pass


def _raise(code, msg):
    # same mapping as the reference's pybind module: invalid_argument -> ValueError, runtime_error -> RuntimeError,
    # out_of_range -> IndexError
    if code == ATK_ERR_INVALID_ARGUMENT:
        raise ValueError(msg)
    if code == ATK_ERR_RUNTIME:
        raise RuntimeError(msg)
    if code == ATK_ERR_OUT_OF_RANGE:
        raise IndexError(msg)
    if code == ATK_ERR_NOMEM:
        raise MemoryError(msg)
    raise AtkCudaError(code, msg)


class CgInfo(C.Structure):
    _fields_ = [("iterations", C.c_int), ("stopped_on_pAp", C.c_int), ("zero_rhs", C.c_int), ("last_pAp", C.c_double),
                ("rs", C.c_double), ("device_ms", C.c_float), ("kernel_launches", C.c_int), ("pAp_trace", _dp),
                ("rs_trace", _dp), ("trace_capacity", C.c_int), ("time_kernels", C.c_int), ("kernel_ms", C.c_float * 3), ("fused", C.c_int), ("peer", C.c_int),
                ("persistent", C.c_int), ("h2d_ms", C.c_float), ("solve_ms", C.c_float), ("d2h_ms", C.c_float)]


class GmresInfo(C.Structure):
    _fields_ = [("restarts", C.c_int), ("status", C.c_int), ("arnoldi_steps", C.c_int), ("initial_residual", C.c_double),
                ("final_residual", C.c_double), ("device_ms", C.c_float), ("kernel_launches", C.c_int),
                ("residual_trace", _dp), ("trace_capacity", C.c_int), ("trace_count", C.c_int)]


class GdInfo(C.Structure):
    _fields_ = [("iterations", C.c_int), ("residual_norm", C.c_double), ("change_norm", C.c_double), ("device_ms", C.c_float),
                ("kernel_launches", C.c_int)]


GMRES_CALLBACK = C.CFUNCTYPE(None, C.c_int, C.c_double, C.c_void_p)

# every symbol include/atk_cuda.h declares (tests check the .so exports all of them)
EXPORTS = [
    "atk_abi_version", "atk_last_error", "atk_device_count", "atk_context_create", "atk_nccl_unique_id",
    "atk_context_create_distributed", "atk_context_destroy", "atk_context_rank", "atk_context_set_reduction",
    "atk_context_synchronize", "atk_context_launch_count", "atk_malloc", "atk_free", "atk_memcpy_h2d", "atk_memcpy_d2h",
    "atk_memcpy_d2d", "atk_memset_zero", "atk_host_alloc", "atk_host_free", "atk_host_numa_node", "atk_dot", "atk_nrm2", "atk_vec_scale",
    "atk_vec_div", "atk_vec_add", "atk_vec_sub", "atk_vec_axpy", "atk_vec_axmy", "atk_operator_from_csc", "atk_operator_from_csc_rows",
    "atk_operator_stencil7", "atk_operator_poisson_assembled", "atk_operator_stencil27", "atk_operator_stencil27_assembled", "atk_slab_partition",
    "atk_operator_destroy",
    "atk_operator_shape", "atk_operator_row_range", "atk_operator_export_csr", "atk_operator_apply", "atk_cg_solve",
    "atk_cg_solve_host", "atk_gmres_solve", "atk_gmres_solve_host", "atk_gd_solve", "atk_gd_solve_host", "atk_arnoldi",
    "atk_arnoldi_host", "atk_ilu_compute", "atk_ilu0_compute", "atk_ilu_solve", "atk_ilu_solve_host", "atk_ilu_export",
    "atk_preconditioner_destroy", "atk_gmres_solve_precond", "atk_gmres_solve_precond_host",
]

_lib = None


def load():
    """Load libatk_cuda.so; raise if it has not been built (python -m algotoolkit_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -m algotoolkit_b200.build` "
                          "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    L.atk_last_error.restype = C.c_char_p
    i64 = C.c_int64
    sig = {
        "atk_device_count": [_ip],
        "atk_context_create": [C.c_int, C.POINTER(_vp)],
        "atk_nccl_unique_id": [_vp, C.c_size_t, C.POINTER(C.c_size_t)],
        "atk_context_create_distributed": [C.c_int, C.c_int, C.c_int, _vp, C.c_size_t, C.POINTER(_vp)],
        "atk_context_destroy": [_vp],
        "atk_context_rank": [_vp, _ip, _ip],
        "atk_context_set_reduction": [_vp, C.c_int],
        "atk_context_synchronize": [_vp],
        "atk_context_launch_count": [_vp, C.POINTER(i64)],
        "atk_malloc": [_vp, C.c_size_t, C.POINTER(_vp)],
        "atk_free": [_vp, _vp],
        "atk_memcpy_h2d": [_vp, _vp, _vp, C.c_size_t],
        "atk_memcpy_d2h": [_vp, _vp, _vp, C.c_size_t],
        "atk_memcpy_d2d": [_vp, _vp, _vp, C.c_size_t],
        "atk_memset_zero": [_vp, _vp, C.c_size_t],
        "atk_host_alloc": [C.c_size_t, C.POINTER(_vp)],
        "atk_host_free": [_vp],
        "atk_host_numa_node": [_vp, _ip, _ip],
        "atk_dot": [_vp, i64, _vp, _vp, _dp],
        "atk_nrm2": [_vp, i64, _vp, _dp],
        "atk_vec_scale": [_vp, i64, _vp, C.c_double, _vp],
        "atk_vec_div": [_vp, i64, _vp, C.c_double, _vp],
        "atk_vec_add": [_vp, i64, _vp, _vp, _vp],
        "atk_vec_sub": [_vp, i64, _vp, _vp, _vp],
        "atk_vec_axpy": [_vp, i64, _vp, _vp, C.c_double, _vp],
        "atk_vec_axmy": [_vp, i64, _vp, _vp, C.c_double, _vp],
        "atk_operator_from_csc": [_vp, C.c_int, C.c_int, i64, _vp, _vp, _vp, C.c_int, C.c_int, C.POINTER(_vp)],
        "atk_operator_from_csc_rows": [_vp, C.c_int, C.c_int, C.c_int, C.c_int, i64, _vp, _vp, _vp, C.c_int, C.c_int, C.POINTER(_vp)],
        "atk_operator_stencil7": [_vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.POINTER(_vp)],
        "atk_operator_poisson_assembled": [_vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int,
                                           C.POINTER(_vp)],
        "atk_operator_stencil27": [_vp, C.c_int, C.c_int, C.c_int, _dp, C.POINTER(_vp)],
        "atk_operator_stencil27_assembled": [_vp, C.c_int, C.c_int, C.c_int, _dp, C.c_int, C.POINTER(_vp)],
        "atk_slab_partition": [C.c_int, C.c_int, C.c_int, _ip, _ip],
        "atk_operator_destroy": [_vp],
        "atk_operator_shape": [_vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)],
        "atk_operator_row_range": [_vp, C.POINTER(i64), C.POINTER(i64)],
        "atk_operator_export_csr": [_vp, _ip, _ip, _dp],
        "atk_operator_apply": [_vp, _vp, _vp, _vp],
        "atk_cg_solve": [_vp, _vp, _vp, _vp, _vp, _vp, C.c_int, C.POINTER(CgInfo)],
        "atk_cg_solve_host": [_vp, _vp, _vp, _vp, _vp, _vp, C.c_int, C.c_int, C.POINTER(CgInfo)],
        "atk_gmres_solve": [_vp, _vp, _vp, _vp, i64, C.c_int, C.c_int, C.c_double, C.c_int, GMRES_CALLBACK, _vp,
                            C.POINTER(GmresInfo)],
        "atk_gmres_solve_host": [_vp, _vp, _vp, _vp, i64, C.c_int, C.c_int, C.c_double, C.c_int, GMRES_CALLBACK, _vp,
                                 C.POINTER(GmresInfo)],
    }
    sig["atk_gd_solve"] = [_vp, _vp, _vp, _vp, C.c_int, C.c_double, C.POINTER(GdInfo)]
    sig["atk_gd_solve_host"] = sig["atk_gd_solve"]
    sig["atk_arnoldi"] = [_vp, _vp, _vp, i64, C.c_int, _dp, C.c_double, _ip]
    sig["atk_arnoldi_host"] = sig["atk_arnoldi"]
    sig["atk_ilu_compute"] = [_vp, i64, i64, i64, _vp, _vp, _vp, C.POINTER(_vp)]
    sig["atk_ilu0_compute"] = sig["atk_ilu_compute"]
    sig["atk_ilu_solve"] = [_vp, _vp, i64, _vp, _vp]
    sig["atk_ilu_solve_host"] = [_vp, _vp, i64, _vp, _vp]
    sig["atk_ilu_export"] = [_vp, _vp, _vp]
    sig["atk_preconditioner_destroy"] = [_vp]
    sig["atk_gmres_solve_precond"] = [_vp, _vp, _vp, _vp, _vp, i64, C.c_int, C.c_int, C.c_double, C.c_int, GMRES_CALLBACK, _vp,
                                      C.POINTER(GmresInfo)]
    sig["atk_gmres_solve_precond_host"] = sig["atk_gmres_solve_precond"]
    for name, args in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = C.c_int
    _lib = L
    return L


def _check(rc):
    if rc != ATK_OK:
        _raise(rc, load().atk_last_error().decode(errors="replace"))


def device_count() -> int:
    n = C.c_int(0)
    _check(load().atk_device_count(C.byref(n)))
    return n.value


def slab_partition(nz: int, rank: int, world: int):
    """Planes [k_begin, k_end) owned by `rank` (host arithmetic of the library, no device needed)."""
    a, b = C.c_int(), C.c_int()
    _check(load().atk_slab_partition(nz, rank, world, C.byref(a), C.byref(b)))
    return a.value, b.value


def _np_ptr(a):
    return a.ctypes.data_as(_vp)


class DeviceVector:
    """A caller-owned device array of doubles (atk_malloc)."""

    def __init__(self, ctx: "Context", n: int):
        self.ctx, self.n = ctx, int(n)
        p = _vp()
        _check(ctx.L.atk_malloc(ctx.h, max(self.n, 1) * 8, C.byref(p)))
        self.ptr = p
        ctx._adopt(self)

    @classmethod
    def from_numpy(cls, ctx, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        v = cls(ctx, a.size)
        if a.size:
            _check(ctx.L.atk_memcpy_h2d(ctx.h, v.ptr, _np_ptr(a), a.size * 8))
        return v

    def upload(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.size == self.n
        if a.size:
            _check(self.ctx.L.atk_memcpy_h2d(self.ctx.h, self.ptr, _np_ptr(a), a.size * 8))

    def zero(self):
        _check(self.ctx.L.atk_memset_zero(self.ctx.h, self.ptr, self.n * 8))

    def copy_from(self, other: "DeviceVector"):
        _check(self.ctx.L.atk_memcpy_d2d(self.ctx.h, self.ptr, other.ptr, self.n * 8))

    def numpy(self):
        out = np.empty(self.n, np.float64)
        if self.n:
            _check(self.ctx.L.atk_memcpy_d2h(self.ctx.h, _np_ptr(out), self.ptr, self.n * 8))
        return out

    def free(self):
        if self.ptr:
            if self.ctx.h:  # a destroyed context has already released its children
                self.ctx.L.atk_free(self.ctx.h, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Operator:
    def __init__(self, ctx, handle):
        self.ctx, self.h = ctx, handle
        ctx._adopt(self)
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        _check(ctx.L.atk_operator_shape(handle, C.byref(a), C.byref(b), C.byref(c)))
        self.n_rows_local, self.n_cols, self.nnz = a.value, b.value, c.value
        _check(ctx.L.atk_operator_row_range(handle, C.byref(a), C.byref(b)))
        self.row_begin = a.value

    def apply(self, x: DeviceVector, y: DeviceVector):
        _check(self.ctx.L.atk_operator_apply(self.ctx.h, self.h, x.ptr, y.ptr))

    def apply_numpy(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        if x.size != self.n_cols and self.ctx.world == 1:
            raise ValueError("Dimension mismatch")  # SparseObj.hpp:132-134
        dx = DeviceVector.from_numpy(self.ctx, x)
        dy = DeviceVector(self.ctx, self.n_rows_local)
        self.apply(dx, dy)
        return dy.numpy()

    def export_csr(self):
        rp = np.zeros(self.n_rows_local + 1, np.int32)
        ci = np.zeros(max(self.nnz, 1), np.int32)
        vv = np.zeros(max(self.nnz, 1), np.float64)
        _check(self.ctx.L.atk_operator_export_csr(self.h, rp.ctypes.data_as(_ip), ci.ctypes.data_as(_ip),
                                                  vv.ctypes.data_as(_dp)))
        return rp, ci[: self.nnz], vv[: self.nnz]

    def destroy(self):
        if self.h:
            if self.ctx.h:
                self.ctx.L.atk_operator_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class Context:
    def __init__(self, device: int = 0, rank: int = 0, world: int = 1, nccl_uid: bytes | None = None):
        self.L = load()
        h = _vp()
        if world == 1:
            _check(self.L.atk_context_create(device, C.byref(h)))
        else:
            buf = C.create_string_buffer(nccl_uid, len(nccl_uid))
            _check(self.L.atk_context_create_distributed(device, rank, world, buf, len(nccl_uid), C.byref(h)))
        self.h, self.rank, self.world, self.device = h, rank, world, device
        self._children = weakref.WeakSet()  # vectors, operators, preconditioners: released before the context goes

    def _adopt(self, child):
        self._children.add(child)
        return child

    @staticmethod
    def nccl_unique_id() -> bytes:
        L = load()
        buf = C.create_string_buffer(256)
        n = C.c_size_t(0)
        _check(L.atk_nccl_unique_id(buf, 256, C.byref(n)))
        return buf.raw[: n.value]

    def set_reduction(self, mode: int):
        _check(self.L.atk_context_set_reduction(self.h, mode))

    def synchronize(self):
        _check(self.L.atk_context_synchronize(self.h))

    def launch_count(self) -> int:
        n = C.c_int64(0)
        _check(self.L.atk_context_launch_count(self.h, C.byref(n)))
        return n.value

    def destroy(self):
        if self.h:
            for child in list(self._children):  # objects still referenced elsewhere (e.g. by a traceback) die here, safely
                (child.free if isinstance(child, DeviceVector) else child.destroy)()
            self.L.atk_context_destroy(self.h)
            self.h = None

    # ---- vectors
    def vector(self, a):
        return DeviceVector.from_numpy(self, a)

    def empty(self, n):
        return DeviceVector(self, n)

    def dot(self, x: DeviceVector, y: DeviceVector) -> float:
        if x.n != y.n:
            raise ValueError("Vector dimensions do not match for dot product.")  # VectorObj.hpp:106
        out = C.c_double()
        _check(self.L.atk_dot(self.h, x.n, x.ptr, y.ptr, C.byref(out)))
        return out.value

    def nrm2(self, x: DeviceVector) -> float:
        out = C.c_double()
        _check(self.L.atk_nrm2(self.h, x.n, x.ptr, C.byref(out)))
        return out.value

    def _ew(self, fn, x, *args):
        out = DeviceVector(self, x.n)
        _check(fn(self.h, x.n, x.ptr, *args, out.ptr))
        return out

    def scale(self, x, s):
        return self._ew(self.L.atk_vec_scale, x, float(s))

    def div(self, x, s):
        return self._ew(self.L.atk_vec_div, x, float(s))

    def add(self, x, y):
        if x.n != y.n:
            raise ValueError("Vector dimensions do not match.")
        return self._ew(self.L.atk_vec_add, x, y.ptr)

    def sub(self, x, y):
        if x.n != y.n:
            raise ValueError("Vector dimensions do not match.")
        return self._ew(self.L.atk_vec_sub, x, y.ptr)

    def axpy(self, x, y, a):
        return self._ew(self.L.atk_vec_axpy, x, y.ptr, float(a))

    def axmy(self, x, y, a):
        return self._ew(self.L.atk_vec_axmy, x, y.ptr, float(a))

    # ---- operators
    def operator_from_csc(self, n_rows, n_cols, col_ptr, row_idx, vals, fmt=FORMAT_SELL) -> Operator:
        col_ptr = np.ascontiguousarray(col_ptr, np.int32)
        row_idx = np.ascontiguousarray(row_idx, np.int32)
        vals = np.ascontiguousarray(vals, np.float64)
        h = _vp()
        _check(self.L.atk_operator_from_csc(self.h, n_rows, n_cols, vals.size, _np_ptr(col_ptr), _np_ptr(row_idx),
                                            _np_ptr(vals), 0, fmt, C.byref(h)))
        return Operator(self, h)

    def operator_from_csc_rows(self, n_rows, n_cols, row_begin, row_end, col_ptr, row_idx, vals, fmt=FORMAT_SELL) -> Operator:
        """This rank's row slice [row_begin, row_end) of a CSC matrix (global row numbers in row_idx)."""
        col_ptr = np.ascontiguousarray(col_ptr, np.int32)
        row_idx = np.ascontiguousarray(row_idx, np.int32)
        vals = np.ascontiguousarray(vals, np.float64)
        h = _vp()
        _check(self.L.atk_operator_from_csc_rows(self.h, n_rows, n_cols, row_begin, row_end, vals.size, _np_ptr(col_ptr),
                                                 _np_ptr(row_idx), _np_ptr(vals), 0, fmt, C.byref(h)))
        return Operator(self, h)

    def operator_stencil7(self, nx, ny, nz, cx, cy, cz) -> Operator:
        h = _vp()
        _check(self.L.atk_operator_stencil7(self.h, nx, ny, nz, cx, cy, cz, C.byref(h)))
        return Operator(self, h)

    def operator_poisson_assembled(self, nx, ny, nz, cx, cy, cz, fmt=FORMAT_SELL) -> Operator:
        h = _vp()
        _check(self.L.atk_operator_poisson_assembled(self.h, nx, ny, nz, cx, cy, cz, fmt, C.byref(h)))
        return Operator(self, h)

    def operator_stencil27(self, nx, ny, nz, coef) -> Operator:
        coef = np.ascontiguousarray(coef, np.float64)
        assert coef.size == 27
        h = _vp()
        _check(self.L.atk_operator_stencil27(self.h, nx, ny, nz, coef.ctypes.data_as(_dp), C.byref(h)))
        return Operator(self, h)

    def operator_stencil27_assembled(self, nx, ny, nz, coef, fmt=FORMAT_SELL) -> Operator:
        coef = np.ascontiguousarray(coef, np.float64)
        assert coef.size == 27
        h = _vp()
        _check(self.L.atk_operator_stencil27_assembled(self.h, nx, ny, nz, coef.ctypes.data_as(_dp), fmt, C.byref(h)))
        return Operator(self, h)

    # ---- solvers
    def cg_solve(self, op: Operator, b: DeviceVector, x: DeviceVector, r: DeviceVector, p: DeviceVector, max_iter: int,
                 trace: bool = False, time_kernels: bool = False):
        info = CgInfo()
        info.time_kernels = 1 if time_kernels else 0
        tp = tr = None
        if trace and max_iter > 0:
            tp, tr = np.full(max_iter, np.nan), np.full(max_iter, np.nan)
            info.pAp_trace, info.rs_trace = tp.ctypes.data_as(_dp), tr.ctypes.data_as(_dp)
            info.trace_capacity = max_iter
        _check(self.L.atk_cg_solve(self.h, op.h, b.ptr, x.ptr, r.ptr, p.ptr, max_iter, C.byref(info)))
        return _cg_result(info, tp, tr)

    def cg_solve_host(self, op: Operator, b, max_iter: int, state=None, want_rp: bool = True, trace: bool = False):
        """b: numpy; state None -> fresh solver (x=0, r=p=b), else (x, r, p) numpy arrays to continue from."""
        b = np.ascontiguousarray(b, np.float64)
        n = b.size
        if state is None:
            x = np.zeros(n)
            r = np.empty(n) if want_rp else None
            p = np.empty(n) if want_rp else None
            fresh = 1
        else:
            x, r, p = (np.ascontiguousarray(s, np.float64).copy() for s in state)
            fresh = 0
        info = CgInfo()
        tp = tr = None
        if trace and max_iter > 0:
            tp, tr = np.full(max_iter, np.nan), np.full(max_iter, np.nan)
            info.pAp_trace, info.rs_trace = tp.ctypes.data_as(_dp), tr.ctypes.data_as(_dp)
            info.trace_capacity = max_iter
        _check(self.L.atk_cg_solve_host(self.h, op.h, _np_ptr(b), _np_ptr(x), _np_ptr(r) if r is not None else None,
                                        _np_ptr(p) if p is not None else None, fresh, max_iter, C.byref(info)))
        out = _cg_result(info, tp, tr)
        out.update(x=x, r=r, p=p)
        return out

    def ilu_compute(self, n_rows, n_cols, col_ptr, row_idx, vals) -> "IluPreconditioner":
        """ILUPreconditioner::compute on finished CSC arrays (ValueError: not square, RuntimeError: singular)."""
        col_ptr = np.ascontiguousarray(col_ptr, np.int32)
        row_idx = np.ascontiguousarray(row_idx, np.int32)
        vals = np.ascontiguousarray(vals, np.float64)
        h = _vp()
        _check(self.L.atk_ilu_compute(self.h, n_rows, n_cols, vals.size, _np_ptr(col_ptr), _np_ptr(row_idx), _np_ptr(vals),
                                      C.byref(h)))
        return IluPreconditioner(self, h, n_rows)

    def ilu0_compute(self, n_rows, n_cols, col_ptr, row_idx, vals) -> "IluPreconditioner":
        """Sparse ILU(0) (no fill-in, level-scheduled) on finished CSC arrays; same handle type as ilu_compute."""
        col_ptr = np.ascontiguousarray(col_ptr, np.int32)
        row_idx = np.ascontiguousarray(row_idx, np.int32)
        vals = np.ascontiguousarray(vals, np.float64)
        h = _vp()
        _check(self.L.atk_ilu0_compute(self.h, n_rows, n_cols, vals.size, _np_ptr(col_ptr), _np_ptr(row_idx), _np_ptr(vals),
                                       C.byref(h)))
        return IluPreconditioner(self, h, n_rows)

    def gmres_solve_host(self, op: Operator, b, x0, max_iter: int, krylov_dim: int, tol: float, ortho=GMRES_MGS,
                         on_event=None, precond: "IluPreconditioner | None" = None):
        b = np.ascontiguousarray(b, np.float64)
        x = np.ascontiguousarray(x0, np.float64).copy()
        info = GmresInfo()
        cap = max_iter + 2
        trace = np.full(cap, np.nan)
        info.residual_trace, info.trace_capacity = trace.ctypes.data_as(_dp), cap
        events = []

        def _cb(ev, val, _user):
            events.append((ev, val))
            if on_event:
                on_event(ev, val)

        cb = GMRES_CALLBACK(_cb)
        rc = self.L.atk_gmres_solve_precond_host(self.h, op.h, precond.h if precond is not None else None, _np_ptr(b),
                                                 _np_ptr(x), x.size, max_iter, krylov_dim, tol, ortho, cb, None,
                                                 C.byref(info))
        err = None
        if rc == ATK_ERR_RUNTIME:
            err = self.L.atk_last_error().decode()
        else:
            _check(rc)
        return dict(x=x, rc=rc, error=err, resid=trace[: min(info.trace_count, cap)].copy(), status=info.status,
                    restarts=info.restarts, arnoldi_steps=info.arnoldi_steps, device_ms=info.device_ms,
                    kernel_launches=info.kernel_launches, events=events)


class IluPreconditioner:
    """Device-resident factors of ILUPreconditioner::compute (reference src/LinearAlgebra/Preconditioner/ILU.hpp)."""

    def __init__(self, ctx, handle, n):
        self.ctx, self.h, self.n = ctx, handle, n
        ctx._adopt(self)

    def solve_host(self, b):
        b = np.ascontiguousarray(b, np.float64)
        x = np.zeros_like(b)
        _check(self.ctx.L.atk_ilu_solve_host(self.ctx.h, self.h, b.size, _np_ptr(b), _np_ptr(x)))
        return x

    def factors(self):
        """(L, U) as dense arrays indexed [row, col], as getLFactor()/getUFactor() hold them."""
        Lf, Uf = np.zeros((self.n, self.n)), np.zeros((self.n, self.n))  # column-major on the C side
        _check(self.ctx.L.atk_ilu_export(self.h, _np_ptr(Lf), _np_ptr(Uf)))
        return Lf.T.copy(), Uf.T.copy()

    def destroy(self):
        if self.h:
            if self.ctx.h:
                self.ctx.L.atk_preconditioner_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def gd_solve_host(ctx: Context, op: Operator, b, x0, max_iter: int, tol: float):
    """GradientDescent::solve through host buffers."""
    b = np.ascontiguousarray(b, np.float64)
    x = np.ascontiguousarray(x0, np.float64).copy()
    info = GdInfo()
    _check(ctx.L.atk_gd_solve_host(ctx.h, op.h, _np_ptr(b), _np_ptr(x), max_iter, tol, C.byref(info)))
    return dict(x=x, iters=info.iterations, residual_norm=info.residual_norm, change_norm=info.change_norm,
                device_ms=info.device_ms, kernel_launches=info.kernel_launches)


def arnoldi_host(ctx: Context, op: Operator, q0, m: int, tol: float):
    """Krylov::Arnoldi through host buffers: returns Q as an (m, n) array (row i = Q[i]) and H as (m, m-1)."""
    n = op.n_rows_local
    Q = np.zeros((m, n))
    Q[0] = np.ascontiguousarray(q0, np.float64)
    H = np.zeros((m - 1, m))  # column-major m x (m-1)
    bd = C.c_int(-1)
    _check(ctx.L.atk_arnoldi_host(ctx.h, op.h, _np_ptr(Q), n, m, H.ctypes.data_as(_dp), tol, C.byref(bd)))
    return dict(Q=Q, H=H.T.copy(), breakdown=bd.value)


def _cg_result(info, tp, tr):
    return dict(iters=info.iterations, stopped_on_pAp=bool(info.stopped_on_pAp), zero_rhs=bool(info.zero_rhs),
                last_pAp=info.last_pAp, rs=info.rs, device_ms=info.device_ms, kernel_launches=info.kernel_launches,
                pAp_trace=tp, rs_trace=tr, kernel_ms=list(info.kernel_ms), fused=bool(info.fused), peer=bool(info.peer),
                persistent=bool(info.persistent), h2d_ms=info.h2d_ms, solve_ms=info.solve_ms, d2h_ms=info.d2h_ms)
