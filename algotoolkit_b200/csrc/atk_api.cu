// atk_api.cu -- the extern "C" surface declared in include/atk_cuda.h.  Every entry point converts C++
// exceptions into status codes (no exception ever crosses the boundary) and records the message for
// atk_last_error().
#include <cmath>
#include <cstring>

#include "atk_internal.cuh"

namespace atk {
atk_context_s* context_create(int device);
atk_context_s* context_create_distributed(int device, int rank, int world, const void* uid, size_t uid_bytes);
void nccl_unique_id(void* buf, size_t cap, size_t* bytes);
void context_destroy(atk_context_s* h);
}  // namespace atk

struct atk_preconditioner {
    atk::Preconditioner* pc = nullptr;
};

namespace {
thread_local std::string g_last_error;

template <typename F>
int guarded(F&& f) {
    try {
        f();
        return ATK_OK;
    } catch (const atk::Error& e) {
        g_last_error = e.msg;
        return e.code;
    } catch (const std::bad_alloc&) {
        g_last_error = "out of host memory";
        return ATK_ERR_NOMEM;
    } catch (const std::exception& e) {
        g_last_error = e.what();
        return ATK_ERR_RUNTIME;
    } catch (...) {
        g_last_error = "unknown error";
        return ATK_ERR_RUNTIME;
    }
}

atk::Context* C(atk_context_t h) {
    ATK_REQUIRE(h != nullptr, ATK_ERR_INVALID_ARGUMENT, "null context");
    ATK_CUDA_CHECK(cudaSetDevice(h->c.device));
    return &h->c;
}
atk::Operator* O(atk_operator_t h) {
    ATK_REQUIRE(h != nullptr && h->op != nullptr, ATK_ERR_INVALID_ARGUMENT, "null operator");
    return h->op;
}
atk::Preconditioner* P(atk_preconditioner_t h, atk::Context* c, bool nullable) {
    if (!h && nullable) return nullptr;
    ATK_REQUIRE(h != nullptr && h->pc != nullptr, ATK_ERR_INVALID_ARGUMENT, "ILU factorization not computed");  // ILU.hpp:85-87
    ATK_REQUIRE(h->pc->ctx == c, ATK_ERR_INVALID_ARGUMENT, "preconditioner belongs to another context");
    return h->pc;
}
}  // namespace

extern "C" {

int atk_abi_version(void) { return ATK_ABI_VERSION; }
const char* atk_last_error(void) { return g_last_error.c_str(); }

int atk_device_count(int* count) {
    return guarded([&] {
        ATK_REQUIRE(count, ATK_ERR_INVALID_ARGUMENT, "null pointer");
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess) {
            cudaGetLastError();
            n = 0;
        }
        *count = n;
    });
}

int atk_context_create(int device, atk_context_t* out) {
    return guarded([&] {
        ATK_REQUIRE(out, ATK_ERR_INVALID_ARGUMENT, "null pointer");
        *out = atk::context_create(device);
    });
}
int atk_nccl_unique_id(void* buf, size_t capacity, size_t* bytes) {
    return guarded([&] { atk::nccl_unique_id(buf, capacity, bytes); });
}
int atk_context_create_distributed(int device, int rank, int world, const void* nccl_uid, size_t uid_bytes,
                                   atk_context_t* out) {
    return guarded([&] {
        ATK_REQUIRE(out, ATK_ERR_INVALID_ARGUMENT, "null pointer");
        *out = atk::context_create_distributed(device, rank, world, nccl_uid, uid_bytes);
    });
}
int atk_context_destroy(atk_context_t ctx) {
    return guarded([&] { atk::context_destroy(ctx); });
}
int atk_context_rank(atk_context_t ctx, int* rank, int* world) {
    return guarded([&] {
        atk::Context* c = C(ctx);
        if (rank) *rank = c->rank;
        if (world) *world = c->world;
    });
}
int atk_context_set_reduction(atk_context_t ctx, atk_reduction mode) {
    return guarded([&] {
        ATK_REQUIRE(mode == ATK_REDUCE_TREE || mode == ATK_REDUCE_SEQUENTIAL, ATK_ERR_INVALID_ARGUMENT, "bad reduction mode");
        C(ctx)->reduction = mode;
    });
}
int atk_context_synchronize(atk_context_t ctx) {
    return guarded([&] {
        atk::Context* c = C(ctx);
        ATK_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        ATK_CUDA_CHECK(cudaStreamSynchronize(c->comm_stream));
    });
}
int atk_context_launch_count(atk_context_t ctx, int64_t* launches) {
    return guarded([&] {
        ATK_REQUIRE(launches, ATK_ERR_INVALID_ARGUMENT, "null pointer");
        *launches = C(ctx)->launches;
    });
}

// ------------------------------------------------------------------------------------------------ memory
int atk_malloc(atk_context_t ctx, size_t bytes, void** dptr) {
    return guarded([&] {
        C(ctx);
        ATK_REQUIRE(dptr, ATK_ERR_INVALID_ARGUMENT, "null pointer");
        cudaError_t e = cudaMalloc(dptr, bytes ? bytes : 8);
        if (e == cudaErrorMemoryAllocation) {
            cudaGetLastError();
            throw atk::Error(ATK_ERR_NOMEM, "out of device memory");
        }
        ATK_CUDA_CHECK(e);
    });
}
int atk_free(atk_context_t ctx, void* dptr) {
    return guarded([&] {
        C(ctx);
        ATK_CUDA_CHECK(cudaFree(dptr));
    });
}
int atk_memcpy_h2d(atk_context_t ctx, void* dst, const void* src, size_t bytes) {
    return guarded([&] {
        atk::Context* c = C(ctx);
        if (!bytes) return;
        ATK_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
        ATK_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    });
}
int atk_memcpy_d2h(atk_context_t ctx, void* dst, const void* src, size_t bytes) {
    return guarded([&] {
        atk::Context* c = C(ctx);
        if (!bytes) return;
        ATK_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
        ATK_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    });
}
int atk_memcpy_d2d(atk_context_t ctx, void* dst, const void* src, size_t bytes) {
    return guarded([&] {
        atk::Context* c = C(ctx);
        if (!bytes) return;
        ATK_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, c->stream));
    });
}
int atk_memset_zero(atk_context_t ctx, void* dptr, size_t bytes) {
    return guarded([&] {
        atk::Context* c = C(ctx);
        if (!bytes) return;
        ATK_CUDA_CHECK(cudaMemsetAsync(dptr, 0, bytes, c->stream));
    });
}
int atk_host_alloc(size_t bytes, void** hptr) {
    return guarded([&] {
        ATK_REQUIRE(hptr, ATK_ERR_INVALID_ARGUMENT, "null pointer");
        *hptr = atk::host_alloc_pinned(bytes ? bytes : 8);
    });
}
int atk_host_free(void* hptr) {
    return guarded([&] { atk::host_free_pinned(hptr); });
}
int atk_host_numa_node(const void* hptr, int* buffer_node, int* device_node) {
    return guarded([&] { atk::host_numa_query(hptr, buffer_node, device_node); });
}

// ------------------------------------------------------------------------------------------------ vectors
int atk_dot(atk_context_t ctx, int64_t n, const double* x, const double* y, double* result_host) {
    return guarded([&] {
        atk::Context* c = C(ctx);
        ATK_REQUIRE(n >= 0 && result_host && (n == 0 || (x && y)), ATK_ERR_INVALID_ARGUMENT, "bad argument to atk_dot");
        atk::launch_dot(c, n, x, y, atk::kSlotGeneric);
        *result_host = atk::read_slot(c, atk::kSlotGeneric);
    });
}
int atk_nrm2(atk_context_t ctx, int64_t n, const double* x, double* result_host) {
    return guarded([&] {
        atk::Context* c = C(ctx);
        ATK_REQUIRE(n >= 0 && result_host && (n == 0 || x), ATK_ERR_INVALID_ARGUMENT, "bad argument to atk_nrm2");
        atk::launch_dot(c, n, x, x, atk::kSlotGeneric);
        *result_host = std::sqrt(atk::read_slot(c, atk::kSlotGeneric));
    });
}
#define ATK_EW_CHECK() ATK_REQUIRE(n >= 0 && (n == 0 || (x && out)), ATK_ERR_INVALID_ARGUMENT, "bad vector argument")
int atk_vec_scale(atk_context_t ctx, int64_t n, const double* x, double s, double* out) {
    return guarded([&] { atk::Context* c = C(ctx); ATK_EW_CHECK(); atk::launch_scale(c, n, x, s, out); });
}
int atk_vec_div(atk_context_t ctx, int64_t n, const double* x, double s, double* out) {
    return guarded([&] {
        atk::Context* c = C(ctx);
        ATK_EW_CHECK();
        ATK_REQUIRE(s != 0.0, ATK_ERR_RUNTIME, "Division by zero.");  // VectorObj.hpp:76
        atk::launch_div(c, n, x, s, out);
    });
}
int atk_vec_add(atk_context_t ctx, int64_t n, const double* x, const double* y, double* out) {
    return guarded([&] { atk::Context* c = C(ctx); ATK_EW_CHECK(); atk::launch_add(c, n, x, y, out); });
}
int atk_vec_sub(atk_context_t ctx, int64_t n, const double* x, const double* y, double* out) {
    return guarded([&] { atk::Context* c = C(ctx); ATK_EW_CHECK(); atk::launch_sub(c, n, x, y, out); });
}
int atk_vec_axpy(atk_context_t ctx, int64_t n, const double* x, const double* y, double a, double* out) {
    return guarded([&] { atk::Context* c = C(ctx); ATK_EW_CHECK(); atk::launch_axpy(c, n, x, y, a, out); });
}
int atk_vec_axmy(atk_context_t ctx, int64_t n, const double* x, const double* y, double a, double* out) {
    return guarded([&] { atk::Context* c = C(ctx); ATK_EW_CHECK(); atk::launch_axmy(c, n, x, y, a, out); });
}

// ------------------------------------------------------------------------------------------------ operators
static int wrap_operator(atk_operator_t* out, atk::Operator* op) {
    *out = new atk_operator_s{op};
    return 0;
}
int atk_operator_from_csc(atk_context_t ctx, int n_rows, int n_cols, int64_t nnz, const int* col_ptr, const int* row_idx,
                          const double* vals, int on_device, atk_format format, atk_operator_t* out) {
    return guarded([&] {
        ATK_REQUIRE(out, ATK_ERR_INVALID_ARGUMENT, "null pointer");
        wrap_operator(out, atk::make_operator_from_csc(C(ctx), n_rows, n_cols, nnz, col_ptr, row_idx, vals, on_device != 0, format));
    });
}
int atk_operator_from_csc_rows(atk_context_t ctx, int n_rows, int n_cols, int row_begin, int row_end, int64_t nnz,
                               const int* col_ptr, const int* row_idx, const double* vals, int on_device, atk_format format,
                               atk_operator_t* out) {
    return guarded([&] {
        ATK_REQUIRE(out, ATK_ERR_INVALID_ARGUMENT, "null pointer");
        wrap_operator(out, atk::make_operator_from_csc_rows(C(ctx), n_rows, n_cols, row_begin, row_end, nnz, col_ptr, row_idx, vals,
                                                            on_device != 0, format));
    });
}
int atk_operator_stencil7(atk_context_t ctx, int nx, int ny, int nz, double cx, double cy, double cz, atk_operator_t* out) {
    return guarded([&] {
        ATK_REQUIRE(out, ATK_ERR_INVALID_ARGUMENT, "null pointer");
        wrap_operator(out, atk::make_stencil7(C(ctx), nx, ny, nz, cx, cy, cz));
    });
}
int atk_operator_poisson_assembled(atk_context_t ctx, int nx, int ny, int nz, double cx, double cy, double cz,
                                   atk_format format, atk_operator_t* out) {
    return guarded([&] {
        ATK_REQUIRE(out, ATK_ERR_INVALID_ARGUMENT, "null pointer");
        wrap_operator(out, atk::make_poisson_assembled(C(ctx), nx, ny, nz, cx, cy, cz, format));
    });
}
int atk_operator_stencil27(atk_context_t ctx, int nx, int ny, int nz, const double coef[27], atk_operator_t* out) {
    return guarded([&] {
        ATK_REQUIRE(out, ATK_ERR_INVALID_ARGUMENT, "null pointer");
        wrap_operator(out, atk::make_stencil27(C(ctx), nx, ny, nz, coef));
    });
}
int atk_operator_stencil27_assembled(atk_context_t ctx, int nx, int ny, int nz, const double coef[27], atk_format format,
                                     atk_operator_t* out) {
    return guarded([&] {
        ATK_REQUIRE(out, ATK_ERR_INVALID_ARGUMENT, "null pointer");
        wrap_operator(out, atk::make_stencil27_assembled(C(ctx), nx, ny, nz, coef, format));
    });
}
int atk_slab_partition(int nz, int rank, int world, int* k_begin, int* k_end) {
    return guarded([&] {
        ATK_REQUIRE(k_begin && k_end && world >= 1 && rank >= 0 && rank < world && nz >= world, ATK_ERR_INVALID_ARGUMENT,
                    "bad slab partition arguments");
        atk::slab_partition(nz, rank, world, k_begin, k_end);
    });
}
int atk_operator_destroy(atk_operator_t op) {
    return guarded([&] {
        if (!op) return;
        if (op->op) {
            cudaSetDevice(op->op->ctx->device);
            cudaStreamSynchronize(op->op->ctx->stream);
            delete op->op;
        }
        delete op;
    });
}
int atk_operator_shape(atk_operator_t op, int64_t* n_rows_local, int64_t* n_cols, int64_t* nnz) {
    return guarded([&] {
        atk::Operator* o = O(op);
        if (n_rows_local) *n_rows_local = o->n_rows_local;
        if (n_cols) *n_cols = o->n_cols;
        if (nnz) *nnz = o->nnz;
    });
}
int atk_operator_row_range(atk_operator_t op, int64_t* row_begin, int64_t* n_rows_local) {
    return guarded([&] {
        atk::Operator* o = O(op);
        if (row_begin) *row_begin = o->row_begin;
        if (n_rows_local) *n_rows_local = o->n_rows_local;
    });
}
int atk_operator_export_csr(atk_operator_t op, int* row_ptr, int* col_idx, double* vals) {
    return guarded([&] {
        atk::Operator* o = O(op);
        ATK_CUDA_CHECK(cudaSetDevice(o->ctx->device));
        o->export_csr(row_ptr, col_idx, vals);
    });
}
int atk_operator_apply(atk_context_t ctx, atk_operator_t op, const double* x, double* y) {
    return guarded([&] {
        atk::Context* c = C(ctx);
        atk::Operator* o = O(op);
        ATK_REQUIRE(o->ctx == c, ATK_ERR_INVALID_ARGUMENT, "operator belongs to another context");
        ATK_REQUIRE(x && y, ATK_ERR_INVALID_ARGUMENT, "null vector");
        o->apply(x, y, -1, nullptr);
    });
}

// ------------------------------------------------------------------------------------------------ CG
int atk_cg_solve(atk_context_t ctx, atk_operator_t op, const double* b, double* x, double* r, double* p, int max_iter,
                 atk_cg_info* info) {
    return guarded([&] {
        atk::Context* c = C(ctx);
        atk::Operator* o = O(op);
        ATK_REQUIRE(o->ctx == c, ATK_ERR_INVALID_ARGUMENT, "operator belongs to another context");
        atk::cg_solve(c, o, b, x, r, p, max_iter, info);
    });
}
int atk_cg_solve_host(atk_context_t ctx, atk_operator_t op, const double* b, double* x, double* r, double* p, int fresh,
                      int max_iter, atk_cg_info* info) {
    return guarded([&] {
        atk::Context* c = C(ctx);
        atk::Operator* o = O(op);
        ATK_REQUIRE(o->ctx == c, ATK_ERR_INVALID_ARGUMENT, "operator belongs to another context");
        ATK_REQUIRE(b && x && info, ATK_ERR_INVALID_ARGUMENT, "null argument");
        ATK_REQUIRE(fresh || (r && p), ATK_ERR_INVALID_ARGUMENT, "continuing a solve needs r and p");
        const size_t bytes = sizeof(double) * (size_t)o->n_rows_local;
        double* d[4] = {nullptr, nullptr, nullptr, nullptr};  // b x r p (cached workspace, recycled between calls)
        auto release = [&] { for (auto* q : d) if (q) c->ws_release(q); };
        try {
            for (auto& q : d) q = static_cast<double*>(c->ws_alloc(bytes ? bytes : 8));
            ATK_CUDA_CHECK(cudaEventRecord(c->ev_h0, c->stream));
            ATK_CUDA_CHECK(cudaMemcpyAsync(d[0], b, bytes, cudaMemcpyHostToDevice, c->stream));
            if (fresh) {  // what the ctor leaves: x = 0, r = b - A*0 = b, p = r  (ConjugateGradient.hpp:26-30)
                ATK_CUDA_CHECK(cudaMemsetAsync(d[1], 0, bytes, c->stream));
                ATK_CUDA_CHECK(cudaMemcpyAsync(d[2], d[0], bytes, cudaMemcpyDeviceToDevice, c->stream));
                ATK_CUDA_CHECK(cudaMemcpyAsync(d[3], d[0], bytes, cudaMemcpyDeviceToDevice, c->stream));
            } else {
                ATK_CUDA_CHECK(cudaMemcpyAsync(d[1], x, bytes, cudaMemcpyHostToDevice, c->stream));
                ATK_CUDA_CHECK(cudaMemcpyAsync(d[2], r, bytes, cudaMemcpyHostToDevice, c->stream));
                ATK_CUDA_CHECK(cudaMemcpyAsync(d[3], p, bytes, cudaMemcpyHostToDevice, c->stream));
            }
            ATK_CUDA_CHECK(cudaEventRecord(c->ev_h1, c->stream));
            atk::cg_solve(c, o, d[0], d[1], d[2], d[3], max_iter, info);
            ATK_CUDA_CHECK(cudaEventRecord(c->ev_h2, c->stream));
            ATK_CUDA_CHECK(cudaMemcpyAsync(x, d[1], bytes, cudaMemcpyDeviceToHost, c->stream));
            if (r) ATK_CUDA_CHECK(cudaMemcpyAsync(r, d[2], bytes, cudaMemcpyDeviceToHost, c->stream));
            if (p) ATK_CUDA_CHECK(cudaMemcpyAsync(p, d[3], bytes, cudaMemcpyDeviceToHost, c->stream));
            ATK_CUDA_CHECK(cudaEventRecord(c->ev_h3, c->stream));
            ATK_CUDA_CHECK(cudaStreamSynchronize(c->stream));
            // where the call's time went: copies up, device loop with its setup, copies down
            ATK_CUDA_CHECK(cudaEventElapsedTime(&info->h2d_ms, c->ev_h0, c->ev_h1));
            ATK_CUDA_CHECK(cudaEventElapsedTime(&info->solve_ms, c->ev_h1, c->ev_h2));
            ATK_CUDA_CHECK(cudaEventElapsedTime(&info->d2h_ms, c->ev_h2, c->ev_h3));
        } catch (...) {
            release();
            throw;
        }
        release();
    });
}

// ------------------------------------------------------------------------------------------------ GMRES
// ------------------------------------------------------------------------------------------------ ILU preconditioner
int atk_ilu_compute(atk_context_t ctx, int64_t n_rows, int64_t n_cols, int64_t nnz, const int32_t* col_ptr,
                    const int32_t* row_idx, const double* vals, atk_preconditioner_t* out) {
    return guarded([&] {
        atk::Context* c = C(ctx);
        ATK_REQUIRE(out, ATK_ERR_INVALID_ARGUMENT, "null pointer");
        *out = nullptr;
        atk::Preconditioner* pc = atk::make_ilu(c, n_rows, n_cols, nnz, col_ptr, row_idx, vals);
        *out = new atk_preconditioner{pc};
    });
}
int atk_ilu0_compute(atk_context_t ctx, int64_t n_rows, int64_t n_cols, int64_t nnz, const int32_t* col_ptr,
                     const int32_t* row_idx, const double* vals, atk_preconditioner_t* out) {
    return guarded([&] {
        atk::Context* c = C(ctx);
        ATK_REQUIRE(out, ATK_ERR_INVALID_ARGUMENT, "null pointer");
        *out = nullptr;
        atk::Preconditioner* pc = atk::make_ilu0(c, n_rows, n_cols, nnz, col_ptr, row_idx, vals);
        *out = new atk_preconditioner{pc};
    });
}
int atk_ilu_solve(atk_context_t ctx, atk_preconditioner_t pc, int64_t n, const double* b, double* x) {
    return guarded([&] {
        atk::Context* c = C(ctx);
        atk::Preconditioner* m = P(pc, c, false);
        ATK_REQUIRE(n == m->n, ATK_ERR_INVALID_ARGUMENT, "Vector size does not match matrix size");  // ILU.hpp:88-90
        ATK_REQUIRE((b && x) || n == 0, ATK_ERR_INVALID_ARGUMENT, "null pointer");
        m->solve(b, x, nullptr);
    });
}
int atk_ilu_solve_host(atk_context_t ctx, atk_preconditioner_t pc, int64_t n, const double* b, double* x) {
    return guarded([&] {
        atk::Context* c = C(ctx);
        atk::Preconditioner* m = P(pc, c, false);
        ATK_REQUIRE(n == m->n, ATK_ERR_INVALID_ARGUMENT, "Vector size does not match matrix size");
        ATK_REQUIRE((b && x) || n == 0, ATK_ERR_INVALID_ARGUMENT, "null pointer");
        const size_t bytes = sizeof(double) * (size_t)n;
        double *db = nullptr, *dx = nullptr;
        try {
            ATK_CUDA_CHECK(cudaMalloc(&db, bytes ? bytes : 8));
            ATK_CUDA_CHECK(cudaMalloc(&dx, bytes ? bytes : 8));
            ATK_CUDA_CHECK(cudaMemcpyAsync(db, b, bytes, cudaMemcpyHostToDevice, c->stream));
            m->solve(db, dx, nullptr);
            ATK_CUDA_CHECK(cudaMemcpyAsync(x, dx, bytes, cudaMemcpyDeviceToHost, c->stream));
            ATK_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        } catch (...) {
            cudaFree(db);
            cudaFree(dx);
            throw;
        }
        cudaFree(db);
        cudaFree(dx);
    });
}
int atk_ilu_export(atk_preconditioner_t pc, double* L, double* U) {
    return guarded([&] {
        ATK_REQUIRE(pc != nullptr && pc->pc != nullptr, ATK_ERR_INVALID_ARGUMENT, "ILU factorization not computed");
        ATK_CUDA_CHECK(cudaSetDevice(pc->pc->ctx->device));
        pc->pc->export_factors(L, U);
    });
}
int atk_preconditioner_destroy(atk_preconditioner_t pc) {
    return guarded([&] {
        if (!pc) return;
        if (pc->pc) {
            cudaSetDevice(pc->pc->ctx->device);
            cudaStreamSynchronize(pc->pc->ctx->stream);
            delete pc->pc;
        }
        delete pc;
    });
}

int atk_gmres_solve_precond(atk_context_t ctx, atk_operator_t op, atk_preconditioner_t pc, const double* b, double* x,
                            int64_t n_x, int max_iter, int krylov_dim, double tol, atk_gmres_ortho ortho,
                            atk_gmres_callback cb, void* user, atk_gmres_info* info) {
    return guarded([&] {
        atk::Context* c = C(ctx);
        atk::Operator* o = O(op);
        ATK_REQUIRE(o->ctx == c, ATK_ERR_INVALID_ARGUMENT, "operator belongs to another context");
        atk::gmres_solve(c, o, P(pc, c, true), b, x, n_x, max_iter, krylov_dim, tol, ortho, cb, user, info);
    });
}
int atk_gmres_solve(atk_context_t ctx, atk_operator_t op, const double* b, double* x, int64_t n_x, int max_iter,
                    int krylov_dim, double tol, atk_gmres_ortho ortho, atk_gmres_callback cb, void* user,
                    atk_gmres_info* info) {
    return atk_gmres_solve_precond(ctx, op, nullptr, b, x, n_x, max_iter, krylov_dim, tol, ortho, cb, user, info);
}
int atk_gmres_solve_host(atk_context_t ctx, atk_operator_t op, const double* b, double* x, int64_t n_x, int max_iter,
                         int krylov_dim, double tol, atk_gmres_ortho ortho, atk_gmres_callback cb, void* user,
                         atk_gmres_info* info) {
    return atk_gmres_solve_precond_host(ctx, op, nullptr, b, x, n_x, max_iter, krylov_dim, tol, ortho, cb, user, info);
}
int atk_gmres_solve_precond_host(atk_context_t ctx, atk_operator_t op, atk_preconditioner_t pc, const double* b, double* x,
                                 int64_t n_x, int max_iter, int krylov_dim, double tol, atk_gmres_ortho ortho,
                                 atk_gmres_callback cb, void* user, atk_gmres_info* info) {
    return guarded([&] {
        atk::Context* c = C(ctx);
        atk::Operator* o = O(op);
        atk::Preconditioner* m = P(pc, c, true);
        ATK_REQUIRE(o->ctx == c, ATK_ERR_INVALID_ARGUMENT, "operator belongs to another context");
        ATK_REQUIRE(b && x && info, ATK_ERR_INVALID_ARGUMENT, "null argument");
        // the size checks of GMRES.hpp:28-37 come before any work
        ATK_REQUIRE(n_x == o->n_rows_local, ATK_ERR_INVALID_ARGUMENT, "Initial guess vector must match system size");
        const size_t bytes = sizeof(double) * (size_t)o->n_rows_local;
        double *db = nullptr, *dx = nullptr;
        try {
            ATK_CUDA_CHECK(cudaMalloc(&db, bytes ? bytes : 8));
            ATK_CUDA_CHECK(cudaMalloc(&dx, bytes ? bytes : 8));
            ATK_CUDA_CHECK(cudaMemcpyAsync(db, b, bytes, cudaMemcpyHostToDevice, c->stream));
            ATK_CUDA_CHECK(cudaMemcpyAsync(dx, x, bytes, cudaMemcpyHostToDevice, c->stream));
            int rc = ATK_OK;
            std::string msg;
            try {
                atk::gmres_solve(c, o, m, db, dx, n_x, max_iter, krylov_dim, tol, ortho, cb, user, info);
            } catch (const atk::Error& e) {
                // a runtime_error mid-solve leaves x at its last value in the reference too: copy it back, then rethrow
                if (e.code != ATK_ERR_RUNTIME) throw;
                rc = e.code;
                msg = e.msg;
            }
            ATK_CUDA_CHECK(cudaMemcpyAsync(x, dx, bytes, cudaMemcpyDeviceToHost, c->stream));
            ATK_CUDA_CHECK(cudaStreamSynchronize(c->stream));
            if (rc != ATK_OK) throw atk::Error(rc, msg);
        } catch (...) {
            cudaFree(db);
            cudaFree(dx);
            throw;
        }
        cudaFree(db);
        cudaFree(dx);
    });
}

// ------------------------------------------------------------------------------------------------ "next" rows
int atk_gd_solve(atk_context_t ctx, atk_operator_t op, const double* b, double* x, int max_iter, double tol, atk_gd_info* info) {
    return guarded([&] {
        atk::Context* c = C(ctx);
        atk::Operator* o = O(op);
        ATK_REQUIRE(o->ctx == c, ATK_ERR_INVALID_ARGUMENT, "operator belongs to another context");
        atk::gd_solve(c, o, b, x, max_iter, tol, info);
    });
}
int atk_gd_solve_host(atk_context_t ctx, atk_operator_t op, const double* b, double* x, int max_iter, double tol,
                      atk_gd_info* info) {
    return guarded([&] {
        atk::Context* c = C(ctx);
        atk::Operator* o = O(op);
        ATK_REQUIRE(o->ctx == c && b && x && info, ATK_ERR_INVALID_ARGUMENT, "bad argument to atk_gd_solve_host");
        const size_t bytes = sizeof(double) * (size_t)o->n_rows_local;
        atk::WsBuffer db(c, bytes), dx(c, bytes);
        ATK_CUDA_CHECK(cudaMemcpyAsync(db.ptr, b, bytes, cudaMemcpyHostToDevice, c->stream));
        ATK_CUDA_CHECK(cudaMemcpyAsync(dx.ptr, x, bytes, cudaMemcpyHostToDevice, c->stream));
        atk::gd_solve(c, o, db.as<double>(), dx.as<double>(), max_iter, tol, info);
        ATK_CUDA_CHECK(cudaMemcpyAsync(x, dx.ptr, bytes, cudaMemcpyDeviceToHost, c->stream));
        ATK_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    });
}
int atk_arnoldi(atk_context_t ctx, atk_operator_t op, double* Q, int64_t ldq, int m, double* H, double tol, int* breakdown_at) {
    return guarded([&] {
        atk::Context* c = C(ctx);
        atk::Operator* o = O(op);
        ATK_REQUIRE(o->ctx == c, ATK_ERR_INVALID_ARGUMENT, "operator belongs to another context");
        atk::arnoldi(c, o, Q, ldq, m, H, tol, breakdown_at);
    });
}
int atk_arnoldi_host(atk_context_t ctx, atk_operator_t op, double* Q, int64_t ldq, int m, double* H, double tol,
                     int* breakdown_at) {
    return guarded([&] {
        atk::Context* c = C(ctx);
        atk::Operator* o = O(op);
        ATK_REQUIRE(o->ctx == c && Q && H && breakdown_at && m >= 1 && ldq >= o->n_rows_local, ATK_ERR_INVALID_ARGUMENT,
                    "bad argument to atk_arnoldi_host");
        const size_t bytes = sizeof(double) * (size_t)ldq * (size_t)m;
        atk::WsBuffer dq(c, bytes);
        ATK_CUDA_CHECK(cudaMemcpyAsync(dq.ptr, Q, bytes, cudaMemcpyHostToDevice, c->stream));
        atk::arnoldi(c, o, dq.as<double>(), ldq, m, H, tol, breakdown_at);
        ATK_CUDA_CHECK(cudaMemcpyAsync(Q, dq.ptr, bytes, cudaMemcpyDeviceToHost, c->stream));
        ATK_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    });
}

}  // extern "C"
