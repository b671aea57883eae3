// atk_ilu0.cu -- sparse ILU(0) with level-scheduled factorisation and triangular solves: the "next" step SURVEY 8(f)
// rank 4 names for the GMRES::enablePreconditioner() branch (GMRES.hpp:39-45,67-68,100-102).
//
// The reference has no sparse incomplete factorisation: its ILUPreconditioner (ILU.hpp:21-115) is a full elimination with
// fill-in, reproduced bit for bit by atk_ilu.cu on a dense n x n array - which limits it to n <= 13500.  This file is the
// same elimination RESTRICTED TO THE SPARSITY PATTERN OF A (no fill-in), for systems of any size:
//     for every stored (i,j), j < i, ascending j:  factor = A(i,j)/A(j,j); only when |factor| > 1e-12       (:48-51)
//         A(i,j) = factor; for every stored (i,k), k > j, with (j,k) stored:  update = A(i,k) - factor*A(j,k),
//         written only when |update| > 1e-12                                                                (:54-58)
//     |A(j,j)| < 1e-12 (or no stored diagonal) -> "Matrix is numerically singular"                           (:38-40)
//     L = unit diagonal + stored entries below it with |value| > 1e-12, U = the others above 1e-12            (:64-76)
//     solve: y_i = b_i - sum_{j<i} L(i,j) y_j ; x_i = (y_i - sum_{j>i} U(i,j) x_j) / U(i,i), sums from 0.0 in
//            ascending j                                                                                     (:94-109)
// Row i only needs rows j < i it has an entry in, so rows are grouped into LEVELS (level(i) = 1 + max level of the rows
// it depends on; the levels come from one sweep over the CSC columns on the host - index bookkeeping, no arithmetic):
//   factorisation  one launch per level, one thread per row, every update a merge of two sorted rows; per entry the
//                  updates arrive in ascending pivot order, exactly as in the column-by-column elimination, so the
//                  factors are bit-identical to the oracle's dense restatement with a pattern mask, and - on matrices
//                  whose full elimination creates no fill-in (tridiagonal ...) - to the reference itself;
//   solve          ONE cooperative launch for both sweeps: the CTAs walk the levels of L, then those of U, with a
//                  software grid barrier between levels; a row's sum takes its terms in ascending column order.
#include <algorithm>
#include <cmath>
#include <vector>

#include "atk_internal.cuh"

namespace atk {

namespace {

constexpr double kEps = 1e-12;  // ILU.hpp:33

__global__ void ilu0_diag_kernel(int n, const int* __restrict__ row_ptr, const int* __restrict__ col, int* __restrict__ diag,
                                 int* __restrict__ singular) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int lo = row_ptr[i], hi = row_ptr[i + 1], pos = -1;
    while (lo < hi) {  // columns are ascending inside a row
        const int mid = (lo + hi) >> 1;
        const int c = col[mid];
        if (c == i) { pos = mid; break; }
        if (c < i) lo = mid + 1;
        else hi = mid;
    }
    diag[i] = pos;
    if (pos < 0) atomicMin(singular, i);  // A(i,i) is 0.0: the reference would stop at step i
}

// rows[0 .. count): the rows of one level.  singular: smallest row whose pivot failed (n = none).
__global__ void __launch_bounds__(128) ilu0_factor_level_kernel(int count, const int* __restrict__ rows,
                                                                const int* __restrict__ row_ptr, const int* __restrict__ col,
                                                                double* __restrict__ val, const int* __restrict__ diag,
                                                                int* __restrict__ singular) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const int i = rows[t];
    const int r0 = row_ptr[i], r1 = row_ptr[i + 1], di = diag[i];
    if (di < 0) return;
    for (int p = r0; p < di; ++p) {
        const int j = col[p];
        const int dj = diag[j];
        if (dj < 0) continue;  // row j has no pivot: flagged by its own thread
        const double pivot = val[dj];  // row j belongs to an earlier level: final
        const double factor = __ddiv_rn(val[p], pivot);
        if (fabs(factor) > kEps) {
            val[p] = factor;
            int q = p + 1, s = dj + 1;
            const int s1 = row_ptr[j + 1];
            while (q < r1 && s < s1) {  // entries (i,k) and (j,k), k > j, both stored
                const int cq = col[q], cs = col[s];
                if (cq == cs) {
                    const double update = sub_rn(val[q], mul_rn(factor, val[s]));
                    if (fabs(update) > kEps) val[q] = update;
                    ++q;
                    ++s;
                } else if (cq < cs) {
                    ++q;
                } else {
                    ++s;
                }
            }
        }
    }
    if (fabs(val[di]) < kEps) atomicMin(singular, i);
}

struct Ilu0SolveArgs {
    int n;
    const int* row_ptr;
    const int* col;
    const double* val;
    const int* diag;
    const int* lrows;
    const int* lptr;
    int n_llev;
    const int* urows;
    const int* uptr;
    int n_ulev;
    const double* b;
    double* y;
    double* x;
    unsigned* bar;
    const int* skip_flag;
};
constexpr int kIlu0Block = 256;
__global__ void __launch_bounds__(kIlu0Block) ilu0_solve_kernel(Ilu0SolveArgs a) {
    if (a.skip_flag && *a.skip_flag) return;
    constexpr int U = 8;
    unsigned target = 0;
    const int gt = blockIdx.x * blockDim.x + threadIdx.x, gs = gridDim.x * blockDim.x;
    for (int lvl = 0; lvl < a.n_llev; ++lvl) {  // L y = b
        for (int t = a.lptr[lvl] + gt; t < a.lptr[lvl + 1]; t += gs) {
            const int i = a.lrows[t];
            double sum = 0.0;
            const int di = a.diag[i];
            const double bi = a.b[i];
            for (int p = a.row_ptr[i]; p < di; p += U) {  // U entries at a time: all their loads in flight before the ordered adds
                int c[U];
                double v[U], yv[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const bool in = p + u < di;
                    c[u] = in ? a.col[p + u] : i;
                    v[u] = in ? a.val[p + u] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < U; ++u) yv[u] = __ldcg(a.y + c[u]);
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (fabs(v[u]) > kEps) sum = add_rn(sum, mul_rn(v[u], yv[u]));
            }
            __stcg(a.y + i, sub_rn(bi, sum));
        }
        coop_grid_barrier(a.bar, target);
    }
    for (int lvl = 0; lvl < a.n_ulev; ++lvl) {  // U x = y
        for (int t = a.uptr[lvl] + gt; t < a.uptr[lvl + 1]; t += gs) {
            const int i = a.urows[t];
            double sum = 0.0;
            const int di = a.diag[i], r1 = a.row_ptr[i + 1];
            const double d = a.val[di];
            const double yi = __ldcg(a.y + i);
            for (int p = di + 1; p < r1; p += U) {
                int c[U];
                double v[U], xv[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const bool in = p + u < r1;
                    c[u] = in ? a.col[p + u] : i;
                    v[u] = in ? a.val[p + u] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < U; ++u) xv[u] = v[u] != 0.0 ? __ldcg(a.x + c[u]) : 0.0;  // (x[i] itself is not written yet)
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (fabs(v[u]) > kEps) sum = add_rn(sum, mul_rn(v[u], xv[u]));
            }
            __stcg(a.x + i, __ddiv_rn(sub_rn(yi, sum), fabs(d) > kEps ? d : 0.0));
        }
        coop_grid_barrier(a.bar, target);
    }
}

struct Ilu0Preconditioner : public Preconditioner {
    int* row_ptr = nullptr;
    int* col = nullptr;
    double* val = nullptr;
    int* diag = nullptr;
    int* lrows = nullptr;
    int* urows = nullptr;
    int* d_lptr = nullptr;
    int* d_uptr = nullptr;
    int n_llev = 0, n_ulev = 0;
    int64_t nnz = 0;
    double* y = nullptr;
    unsigned* bar = nullptr;
    int grid = 1;

    ~Ilu0Preconditioner() override {
        cudaFree(row_ptr); cudaFree(col); cudaFree(val); cudaFree(diag); cudaFree(lrows); cudaFree(urows);
        cudaFree(d_lptr); cudaFree(d_uptr); cudaFree(y); cudaFree(bar);
    }
    void solve(const double* b, double* x, const int* skip_flag) override {
        if (n == 0) return;
        ATK_NVTX("atk::ilu0_solve");
        ATK_CUDA_CHECK(cudaMemsetAsync(bar, 0, sizeof(unsigned), ctx->stream));
        Ilu0SolveArgs a{(int)n, row_ptr, col, val, diag, lrows, d_lptr, n_llev, urows, d_uptr, n_ulev, b, y, x, bar, skip_flag};
        void* kargs[] = {&a};
        ATK_CUDA_CHECK(cudaLaunchCooperativeKernel((void*)ilu0_solve_kernel, dim3((unsigned)grid), dim3(kIlu0Block), kargs, 0, ctx->stream));
        count_launch(ctx);
    }
    void export_factors(double* L, double* U) override {
        std::vector<int> rp((size_t)n + 1), ci((size_t)std::max<int64_t>(nnz, 1));
        std::vector<double> vv((size_t)std::max<int64_t>(nnz, 1));
        ATK_CUDA_CHECK(cudaMemcpy(rp.data(), row_ptr, sizeof(int) * ((size_t)n + 1), cudaMemcpyDeviceToHost));
        if (nnz > 0) {
            ATK_CUDA_CHECK(cudaMemcpy(ci.data(), col, sizeof(int) * (size_t)nnz, cudaMemcpyDeviceToHost));
            ATK_CUDA_CHECK(cudaMemcpy(vv.data(), val, sizeof(double) * (size_t)nnz, cudaMemcpyDeviceToHost));
        }
        const size_t nn = (size_t)n * (size_t)n;
        if (L) std::fill(L, L + nn, 0.0);
        if (U) std::fill(U, U + nn, 0.0);
        for (int64_t i = 0; i < n; ++i) {
            if (L) L[i + i * n] = 1.0;
            for (int p = rp[(size_t)i]; p < rp[(size_t)i + 1]; ++p) {
                const int64_t j = ci[(size_t)p];
                const double v = std::fabs(vv[(size_t)p]) > kEps ? vv[(size_t)p] : 0.0;  // ILU.hpp:67-76
                if (j < i) { if (L) L[i + j * n] = v; }
                else if (U) U[i + j * n] = v;
            }
        }
    }
};

// rows ordered by level (counting sort) + level pointers
void order_by_level(const std::vector<int>& lev, std::vector<int>& rows, std::vector<int>& ptr) {
    int n_lev = 0;
    for (int v : lev) n_lev = std::max(n_lev, v + 1);
    ptr.assign((size_t)n_lev + 1, 0);
    for (int v : lev) ptr[(size_t)v + 1]++;
    for (int l = 0; l < n_lev; ++l) ptr[(size_t)l + 1] += ptr[(size_t)l];
    rows.resize(lev.size());
    std::vector<int> fill(ptr.begin(), ptr.end() - 1);
    for (size_t i = 0; i < lev.size(); ++i) rows[(size_t)fill[(size_t)lev[i]]++] = (int)i;
}

template <typename T>
T* dalloc0(size_t n) {
    T* p = nullptr;
    ATK_CUDA_CHECK(cudaMalloc(&p, sizeof(T) * std::max<size_t>(n, 1)));
    return p;
}

}  // namespace

Preconditioner* make_ilu0(Context* ctx, int64_t n_rows, int64_t n_cols, int64_t nnz, const int32_t* col_ptr, const int32_t* row_idx,
                          const double* vals) {
    ATK_REQUIRE(ctx->world == 1, ATK_ERR_INVALID_ARGUMENT, "the ILU(0) preconditioner is built on single-GPU contexts only");
    ATK_NVTX("atk::ilu0_compute");
    ATK_REQUIRE(n_rows == n_cols, ATK_ERR_INVALID_ARGUMENT, "Matrix must be square for ILU factorization");  // ILU.hpp:24-26
    ATK_REQUIRE(n_rows >= 0 && n_rows < 0x7fffffff && nnz >= 0 && nnz < 0x7fffffff && col_ptr && (nnz == 0 || (row_idx && vals)),
                ATK_ERR_INVALID_ARGUMENT, "bad CSC arrays");
    const int n = (int)n_rows;
    auto* P = new Ilu0Preconditioner();
    Operator* tmp = nullptr;
    try {
        P->ctx = ctx;
        P->n = n;
        P->nnz = nnz;
        cudaStream_t st = ctx->stream;
        // CSC -> CSR (ascending columns) by the device conversion every assembled operator goes through (it also validates
        // col_ptr and the row indices); the factors take over a copy of those arrays
        tmp = make_operator_from_csc(ctx, n, n, nnz, col_ptr, row_idx, vals, false, ATK_FORMAT_CSR_VECTOR);
        const int* t_rp = nullptr;
        const int* t_ci = nullptr;
        const double* t_v = nullptr;
        ATK_REQUIRE(tmp->device_csr(&t_rp, &t_ci, &t_v), ATK_ERR_RUNTIME, "internal: no CSR arrays");
        P->row_ptr = dalloc0<int>((size_t)n + 1);
        P->col = dalloc0<int>((size_t)nnz);
        P->val = dalloc0<double>((size_t)nnz);
        P->diag = dalloc0<int>((size_t)n);
        P->y = dalloc0<double>((size_t)n);
        P->bar = dalloc0<unsigned>(1);
        ATK_CUDA_CHECK(cudaMemcpyAsync(P->row_ptr, t_rp, sizeof(int) * ((size_t)n + 1), cudaMemcpyDeviceToDevice, st));
        if (nnz > 0) {
            ATK_CUDA_CHECK(cudaMemcpyAsync(P->col, t_ci, sizeof(int) * (size_t)nnz, cudaMemcpyDeviceToDevice, st));
            ATK_CUDA_CHECK(cudaMemcpyAsync(P->val, t_v, sizeof(double) * (size_t)nnz, cudaMemcpyDeviceToDevice, st));
        }
        ATK_CUDA_CHECK(cudaStreamSynchronize(st));
        delete tmp;
        tmp = nullptr;
        // levels: row i waits for the rows j < i (L, and the factorisation itself) / j > i (U) it has an entry in
        std::vector<int> levL((size_t)n, 0), levU((size_t)n, 0);
        for (int j = 0; j < n; ++j)
            for (int p = col_ptr[j]; p < col_ptr[j + 1]; ++p) {
                const int i = row_idx[p];
                if (i > j) levL[(size_t)i] = std::max(levL[(size_t)i], levL[(size_t)j] + 1);
            }
        for (int j = n - 1; j >= 0; --j)
            for (int p = col_ptr[j]; p < col_ptr[j + 1]; ++p) {
                const int i = row_idx[p];
                if (i < j) levU[(size_t)i] = std::max(levU[(size_t)i], levU[(size_t)j] + 1);
            }
        std::vector<int> lrows, lptr, urows, uptr;
        order_by_level(levL, lrows, lptr);
        order_by_level(levU, urows, uptr);
        P->n_llev = (int)lptr.size() - 1;
        P->n_ulev = (int)uptr.size() - 1;
        P->lrows = dalloc0<int>((size_t)n);
        P->urows = dalloc0<int>((size_t)n);
        P->d_lptr = dalloc0<int>(lptr.size());
        P->d_uptr = dalloc0<int>(uptr.size());
        if (n > 0) {
            ATK_CUDA_CHECK(cudaMemcpyAsync(P->lrows, lrows.data(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, st));
            ATK_CUDA_CHECK(cudaMemcpyAsync(P->urows, urows.data(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, st));
        }
        ATK_CUDA_CHECK(cudaMemcpyAsync(P->d_lptr, lptr.data(), sizeof(int) * lptr.size(), cudaMemcpyHostToDevice, st));
        ATK_CUDA_CHECK(cudaMemcpyAsync(P->d_uptr, uptr.data(), sizeof(int) * uptr.size(), cudaMemcpyHostToDevice, st));
        // factorisation, level by level
        int* d_sing = dalloc0<int>(1);
        int h_sing = n;
        ATK_CUDA_CHECK(cudaMemcpyAsync(d_sing, &h_sing, sizeof(int), cudaMemcpyHostToDevice, st));
        if (n > 0) {
            ilu0_diag_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, P->row_ptr, P->col, P->diag, d_sing);
            count_launch(ctx);
            for (int l = 0; l < P->n_llev; ++l) {
                const int count = lptr[(size_t)l + 1] - lptr[(size_t)l];
                ilu0_factor_level_kernel<<<(count + 127) / 128, 128, 0, st>>>(count, P->lrows + lptr[(size_t)l], P->row_ptr, P->col, P->val,
                                                                             P->diag, d_sing);
            }
            count_launch(ctx, P->n_llev);
        }
        ATK_CUDA_CHECK(cudaGetLastError());
        ATK_CUDA_CHECK(cudaMemcpyAsync(&h_sing, d_sing, sizeof(int), cudaMemcpyDeviceToHost, st));
        ATK_CUDA_CHECK(cudaStreamSynchronize(st));
        cudaFree(d_sing);
        ATK_REQUIRE(h_sing >= n, ATK_ERR_RUNTIME, "Matrix is numerically singular");  // :38-40
        // the solve is one cooperative launch: every CTA must be resident
        int per_sm = 0, coop = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device);
        ATK_REQUIRE(coop, ATK_ERR_CUDA, "cooperative launch is not supported on this device");
        ATK_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ilu0_solve_kernel, kIlu0Block, 0));
        int widest = 1;
        for (size_t l = 0; l + 1 < lptr.size(); ++l) widest = std::max(widest, lptr[l + 1] - lptr[l]);
        for (size_t l = 0; l + 1 < uptr.size(); ++l) widest = std::max(widest, uptr[l + 1] - uptr[l]);
        P->grid = std::max(1, std::min(std::max(per_sm, 1) * ctx->sm_count, (widest + kIlu0Block - 1) / kIlu0Block));
    } catch (...) {
        cudaStreamSynchronize(ctx->stream);
        delete tmp;
        delete P;
        throw;
    }
    return P;
}

}  // namespace atk
