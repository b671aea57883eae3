// atk_ilu.cu -- ILUPreconditioner on the device (reference src/LinearAlgebra/Preconditioner/ILU.hpp:21-115), the
// preconditioner behind GMRES::enablePreconditioner() (GMRES.hpp:23-25,39-45,67-68,100-102).  "Next" row of the scope
// table.
//
// What the reference computes.  compute() is a FULL elimination without pivoting, carried out through SparseMatrixCSC
// lookups (fill-in is created wherever an update exceeds 1e-12), so the factors are dense in all but name and a dense
// n x n column-major array restates it exactly (an absent entry reads as 0.0):
//     step j:  |A(j,j)| < 1e-12 -> "Matrix is numerically singular"                                  (:38-40)
//              for i > j: factor = A(i,j) / A(j,j); only when |factor| > 1e-12                       (:48-50)
//                  A(i,j) = factor                                                                   (:51)
//                  for k > j: update = A(i,k) - factor*A(j,k); written only when |update| > 1e-12    (:54-58)
//                             (a smaller update leaves the OLD value in place)
//     L = unit diagonal + entries below the diagonal with |value| > 1e-12, U = the others above 1e-12 (:64-76)
// solve() is plain forward / back substitution with every sum taken from 0.0 in ascending column order (:94-109).
//
// Device formulation.
//   compute: step j is two launches - the multipliers of column j (one thread per row), then the rank-1 update of the
//            trailing block (one thread per entry, warps along a column so the loads coalesce).  Every entry does the
//            reference's two roundings (product, then difference): the factors are bit-identical.
//   solve:   one launch, one CTA, both sweeps, in blocks of 32 columns.  Forward substitution: when the y_j of a block are
//            final every row below adds L(i,j)*y_j to its running sum, j ascending - the reference's order, so the
//            forward sweep is always bit-exact.  The backward sweep has two forms: with ATK_REDUCE_TREE the same column wavefront
//            (x_j final -> rows above add U(i,j)*x_j), which accumulates in DESCENDING j and therefore differs from the
//            reference by rounding only; with ATK_REDUCE_SEQUENTIAL each row's products are formed in parallel and one
//            thread adds them in ascending j, the reference's chain - bit-exact, and inherently serial (n^2/2 dependent
//            adds).  Running sums and y live in shared memory: n <= 13500.
#include <algorithm>
#include <cmath>
#include <vector>

#include "atk_internal.cuh"

namespace atk {

namespace {

constexpr double kIluEps = 1e-12;  // ILU.hpp:33

__global__ void ilu_scatter_kernel(int64_t nnz, int64_t n, const int32_t* __restrict__ col_of, const int32_t* __restrict__ row_idx,
                                   const double* __restrict__ vals, double* __restrict__ A) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nnz) A[(int64_t)row_idx[k] + (int64_t)col_of[k] * n] = vals[k];
}

__global__ void ilu_expand_cols_kernel(int64_t n_cols, const int32_t* __restrict__ col_ptr, int32_t* __restrict__ col_of) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n_cols)
        for (int32_t k = col_ptr[c]; k < col_ptr[c + 1]; ++k) col_of[k] = (int32_t)c;
}

// multipliers of column j: f[i] = A(i,j)/A(j,j) when it passes the threshold (and A(i,j) takes it), else 0 = "skip row i"
__global__ void ilu_factor_kernel(int64_t n, int64_t j, double* __restrict__ A, double* __restrict__ f, int* __restrict__ singular) {
    if (*singular) return;
    const double pivot = A[j + j * n];
    if (fabs(pivot) < kIluEps) {  // every thread sees the same pivot; the flag stops the later steps
        if (blockIdx.x == 0 && threadIdx.x == 0) *singular = (int)j + 1;
        return;
    }
    const int64_t i = j + 1 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double factor = __ddiv_rn(A[i + j * n], pivot);
    const bool take = fabs(factor) > kIluEps;
    if (take) A[i + j * n] = factor;
    f[i] = take ? factor : 0.0;
}

// trailing block of step j: blockDim = (32 rows, 8 columns)
__global__ void __launch_bounds__(256) ilu_update_kernel(int64_t n, int64_t j, double* __restrict__ A, const double* __restrict__ f,
                                                         const int* __restrict__ singular) {
    if (*singular) return;
    const int64_t i = j + 1 + (int64_t)blockIdx.x * 32 + threadIdx.x;
    const int64_t k = j + 1 + (int64_t)blockIdx.y * 8 + threadIdx.y;
    if (i >= n || k >= n) return;
    const double factor = f[i];
    if (factor == 0.0) return;
    const double update = sub_rn(A[i + k * n], mul_rn(factor, A[j + k * n]));
    if (fabs(update) > kIluEps) A[i + k * n] = update;
}

// M = entries the reference extracts (|v| > 1e-12, else 0), column-major, and its transpose
__global__ void ilu_mask_kernel(int64_t n, const double* __restrict__ A, double* __restrict__ M, double* __restrict__ Mt) {
    const int64_t i = (int64_t)blockIdx.x * 32 + threadIdx.x;
    const int64_t k = (int64_t)blockIdx.y * 8 + threadIdx.y;
    if (i >= n || k >= n) return;
    const double v = A[i + k * n];
    const double m = fabs(v) > kIluEps ? v : 0.0;
    M[i + k * n] = m;
    Mt[k + i * n] = m;
}

// x = U^-1 L^-1 b.  One CTA; thread t owns rows t, t+T, ...; the sweeps go in blocks of 32 columns.
//
// Forward, block [jb, jb+32): warp 0 solves the 32 x 32 diagonal block (staged in shared memory) as a wavefront inside the
// warp - lane l owns row jb+l, y_j is broadcast by shuffle as soon as lane j-jb has subtracted its last term - then, after
// ONE block-wide barrier, every row below adds its 32 products L(i,j)*y_j, j ascending.  Each row's running sum therefore
// still receives its terms in ascending j from 0.0: the reference's order (ILU.hpp:94-100), bit-exact, with n/32 barriers
// instead of n.  The backward sweep mirrors it in tree mode (terms arrive in descending j: rounding-level difference to
// the reference); in reference-ordered mode each row's products are formed in parallel and ONE thread adds them in
// ascending j (ILU.hpp:103-109) - bit-exact, inherently serial.
constexpr int kIluB = 32;

__global__ void __launch_bounds__(1024) ilu_solve_kernel(int n, const double* __restrict__ M, const double* __restrict__ Mt,
                                                         const double* __restrict__ b, double* __restrict__ x, int exact_order,
                                                         const int* __restrict__ skip_flag) {
    if (skip_flag && *skip_flag) return;
    extern __shared__ __align__(16) double ilu_sm[];
    double* acc = ilu_sm;     // running sums (forward, then backward)
    double* yv = ilu_sm + n;  // b, then y, then x - entry by entry
    __shared__ double diag[kIluB][kIluB + 1];  // diagonal block of the current step, [row][col]
    const int T = blockDim.x, t = threadIdx.x;
    const int lane = t & 31;
    const int64_t N = n;
    for (int i = t; i < n; i += T) {
        acc[i] = 0.0;
        yv[i] = b[i];
    }
    // diagonal block of [jb, jb+32) -> shared memory (entries outside the matrix are never read)
    auto load_diag = [&](int jb) {
        for (int e = t; e < kIluB * kIluB; e += T) {
            const int r = e & (kIluB - 1), c = e >> 5;  // consecutive threads walk down a column: coalesced
            if (jb + r < n && jb + c < n) diag[r][c] = M[(jb + r) + (jb + c) * N];
        }
    };
    // rows [lo, hi) add the products of columns [jb, jb+w) in the order given by `descending`
    auto update_rows = [&](int lo, int hi, int jb, int w, bool descending) {
        int i0 = lo - (lo % T) + t;
        if (i0 < lo) i0 += T;
        for (int i = i0; i < hi; i += T) {
            double a = acc[i];
#pragma unroll 1
            for (int c0 = 0; c0 < w; c0 += 16) {
                double m[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const int s = descending ? w - 1 - (c0 + q) : c0 + q;
                    m[q] = (c0 + q < w) ? M[i + (jb + s) * N] : 0.0;
                }
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const int s = descending ? w - 1 - (c0 + q) : c0 + q;
                    if (c0 + q < w) a = add_rn(a, mul_rn(m[q], yv[jb + s]));
                }
            }
            acc[i] = a;
        }
    };
    load_diag(0);
    __syncthreads();
    // ---- forward: y_i = b_i - sum_{j<i} L(i,j) y_j
    for (int jb = 0; jb < n; jb += kIluB) {
        const int w = min(kIluB, n - jb);
        if (t < 32) {
            const int i = jb + lane;
            double a = lane < w ? acc[i] : 0.0;
            const double bi = lane < w ? yv[i] : 0.0;
            double y = 0.0;
            for (int s = 0; s < w; ++s) {
                if (lane == s) y = sub_rn(bi, a);
                const double ys = __shfl_sync(0xffffffffu, y, s);
                if (lane > s && lane < w) a = add_rn(a, mul_rn(diag[lane][s], ys));
            }
            if (lane < w) yv[i] = y;
        }
        __syncthreads();
        update_rows(jb + w, n, jb, w, false);
        if (jb + kIluB < n) load_diag(jb + kIluB);  // warp 0 has left the old block (barrier above)
        __syncthreads();
    }
    // ---- backward: x_i = (y_i - sum_{j>i} U(i,j) x_j) / U(i,i)
    if (!exact_order) {
        for (int i = t; i < n; i += T) acc[i] = 0.0;
        const int last = ((n - 1) / kIluB) * kIluB;
        if (n > 0) load_diag(last);
        __syncthreads();
        for (int jb = last; jb >= 0; jb -= kIluB) {
            const int w = min(kIluB, n - jb);
            if (t < 32) {
                const int i = jb + lane;
                double a = lane < w ? acc[i] : 0.0;
                const double yi = lane < w ? yv[i] : 0.0;
                const double d = lane < w ? diag[lane][lane] : 1.0;
                double xv = 0.0;
                for (int s = w - 1; s >= 0; --s) {
                    if (lane == s) xv = __ddiv_rn(sub_rn(yi, a), d);
                    const double xs = __shfl_sync(0xffffffffu, xv, s);
                    if (lane < s) a = add_rn(a, mul_rn(diag[lane][s], xs));
                }
                if (lane < w) yv[i] = xv;
            }
            __syncthreads();
            update_rows(0, jb, jb, w, true);
            if (jb > 0) load_diag(jb - kIluB);
            __syncthreads();
        }
    } else {
        for (int i = n - 1; i >= 0; --i) {
            const double* row = Mt + (int64_t)i * N;  // row i of U, contiguous
            for (int j = i + 1 + t; j < n; j += T) acc[j] = mul_rn(row[j], yv[j]);
            __syncthreads();
            if (t == 0) {
                double sum = 0.0;
#pragma unroll 8
                for (int j = i + 1; j < n; ++j) sum = add_rn(sum, acc[j]);
                yv[i] = __ddiv_rn(sub_rn(yv[i], sum), row[i]);
            }
            __syncthreads();
        }
    }
    __syncthreads();
    for (int i = t; i < n; i += T) x[i] = yv[i];
}

struct IluPreconditioner final : Preconditioner {
    double* d_m = nullptr;   // masked factors, column-major
    double* d_mt = nullptr;  // transpose
    int block = 32;
    size_t smem = 0;
    ~IluPreconditioner() override {
        cudaFree(d_m);
        cudaFree(d_mt);
    }
    void solve(const double* b, double* x, const int* skip_flag) override {
        if (n == 0) return;
        ilu_solve_kernel<<<1, block, smem, ctx->stream>>>((int)n, d_m, d_mt, b, x, ctx->reduction == ATK_REDUCE_SEQUENTIAL ? 1 : 0,
                                                          skip_flag);
        ctx->launches++;
        ATK_CUDA_CHECK(cudaGetLastError());
    }
    void export_factors(double* L, double* U) override {
        const size_t nn = (size_t)n * (size_t)n;
        std::vector<double> m(nn);
        ATK_CUDA_CHECK(cudaMemcpyAsync(m.data(), d_m, sizeof(double) * nn, cudaMemcpyDeviceToHost, ctx->stream));
        ATK_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        for (int64_t j = 0; j < n; ++j)
            for (int64_t i = 0; i < n; ++i) {
                const double v = m[(size_t)(i + j * n)];
                if (L) L[i + j * n] = i == j ? 1.0 : (i > j ? v : 0.0);
                if (U) U[i + j * n] = i <= j ? v : 0.0;
            }
    }
};

}  // namespace

Preconditioner* make_ilu(Context* ctx, int64_t n_rows, int64_t n_cols, int64_t nnz, const int32_t* col_ptr, const int32_t* row_idx,
                         const double* vals) {
    ATK_REQUIRE(ctx->world == 1, ATK_ERR_INVALID_ARGUMENT, "the ILU preconditioner is built on single-GPU contexts only");
    ATK_NVTX("atk::ilu_compute");
    ATK_REQUIRE(n_rows == n_cols, ATK_ERR_INVALID_ARGUMENT, "Matrix must be square for ILU factorization");  // :24-26
    ATK_REQUIRE(n_rows >= 0 && nnz >= 0 && (n_rows == 0 || (col_ptr && (nnz == 0 || (row_idx && vals)))), ATK_ERR_INVALID_ARGUMENT,
                "bad CSC arrays");
    const int64_t n = n_rows;
    ATK_REQUIRE(n <= 13500, ATK_ERR_INVALID_ARGUMENT,
                "ILU factorization: the reference's elimination is dense (n x n); this build keeps n <= 13500");
    cudaStream_t st = ctx->stream;
    auto pc = new IluPreconditioner();
    pc->ctx = ctx;
    pc->n = n;
    double *d_a = nullptr, *d_f = nullptr, *d_vals = nullptr;
    int32_t *d_cp = nullptr, *d_ri = nullptr, *d_colof = nullptr;
    int* d_sing = nullptr;
    auto cleanup = [&]() {
        cudaFree(d_a); cudaFree(d_f); cudaFree(d_vals); cudaFree(d_cp); cudaFree(d_ri); cudaFree(d_colof); cudaFree(d_sing);
    };
    try {
        const size_t nn = std::max<size_t>(1, (size_t)n * (size_t)n);
        ATK_CUDA_CHECK(cudaMalloc(&d_a, sizeof(double) * nn));
        ATK_CUDA_CHECK(cudaMalloc(&pc->d_m, sizeof(double) * nn));
        ATK_CUDA_CHECK(cudaMalloc(&pc->d_mt, sizeof(double) * nn));
        ATK_CUDA_CHECK(cudaMalloc(&d_f, sizeof(double) * (size_t)std::max<int64_t>(n, 1)));
        ATK_CUDA_CHECK(cudaMalloc(&d_sing, sizeof(int)));
        ATK_CUDA_CHECK(cudaMemsetAsync(d_a, 0, sizeof(double) * nn, st));
        ATK_CUDA_CHECK(cudaMemsetAsync(d_sing, 0, sizeof(int), st));
        if (nnz > 0) {
            // range checks on the host arrays first (SparseObj.hpp:107 out_of_range)
            ATK_REQUIRE(col_ptr[0] == 0 && col_ptr[n] == nnz, ATK_ERR_INVALID_ARGUMENT, "col_ptr does not span nnz");
            for (int64_t k = 0; k < nnz; ++k)
                ATK_REQUIRE(row_idx[k] >= 0 && row_idx[k] < n, ATK_ERR_OUT_OF_RANGE, "Index out of range");
            ATK_CUDA_CHECK(cudaMalloc(&d_cp, sizeof(int32_t) * (size_t)(n + 1)));
            ATK_CUDA_CHECK(cudaMalloc(&d_ri, sizeof(int32_t) * (size_t)nnz));
            ATK_CUDA_CHECK(cudaMalloc(&d_colof, sizeof(int32_t) * (size_t)nnz));
            ATK_CUDA_CHECK(cudaMalloc(&d_vals, sizeof(double) * (size_t)nnz));
            ATK_CUDA_CHECK(cudaMemcpyAsync(d_cp, col_ptr, sizeof(int32_t) * (size_t)(n + 1), cudaMemcpyHostToDevice, st));
            ATK_CUDA_CHECK(cudaMemcpyAsync(d_ri, row_idx, sizeof(int32_t) * (size_t)nnz, cudaMemcpyHostToDevice, st));
            ATK_CUDA_CHECK(cudaMemcpyAsync(d_vals, vals, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, st));
            ilu_expand_cols_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, d_cp, d_colof);
            ilu_scatter_kernel<<<(unsigned)((nnz + 255) / 256), 256, 0, st>>>(nnz, n, d_colof, d_ri, d_vals, d_a);
            ctx->launches += 2;
        }
        for (int64_t j = 0; j < n; ++j) {
            const int64_t rem = n - j - 1;
            // the singularity test of step j happens even when nothing is left to eliminate
            ilu_factor_kernel<<<(unsigned)std::max<int64_t>(1, (rem + 255) / 256), 256, 0, st>>>(n, j, d_a, d_f, d_sing);
            ctx->launches++;
            if (rem > 0) {
                ilu_update_kernel<<<dim3((unsigned)((rem + 31) / 32), (unsigned)((rem + 7) / 8)), dim3(32, 8), 0, st>>>(n, j, d_a, d_f,
                                                                                                                  d_sing);
                ctx->launches++;
            }
        }
        if (n > 0) {
            ilu_mask_kernel<<<dim3((unsigned)((n + 31) / 32), (unsigned)((n + 7) / 8)), dim3(32, 8), 0, st>>>(n, d_a, pc->d_m, pc->d_mt);
            ctx->launches++;
        }
        ATK_CUDA_CHECK(cudaGetLastError());
        int singular = 0;
        ATK_CUDA_CHECK(cudaMemcpyAsync(&singular, d_sing, sizeof(int), cudaMemcpyDeviceToHost, st));
        ATK_CUDA_CHECK(cudaStreamSynchronize(st));
        ATK_REQUIRE(singular == 0, ATK_ERR_RUNTIME, "Matrix is numerically singular");  // :38-40
        pc->block = (int)std::min<int64_t>(1024, std::max<int64_t>(32, (n + 31) / 32 * 32));
        pc->smem = sizeof(double) * 2 * (size_t)std::max<int64_t>(n, 1);
        // The attribute is per FUNCTION and process-wide, not per preconditioner: raise it to the device's opt-in maximum
        // once, so that building a small factorisation later can never lower the limit under a live, larger one.
        int max_optin = 0;
        ATK_CUDA_CHECK(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
        cudaFuncAttributes fa;
        ATK_CUDA_CHECK(cudaFuncGetAttributes(&fa, ilu_solve_kernel));
        ATK_REQUIRE(pc->smem + fa.sharedSizeBytes <= (size_t)max_optin, ATK_ERR_INVALID_ARGUMENT,
                    "ILU preconditioner: system too large for the single-CTA triangular sweeps");
        ATK_CUDA_CHECK(cudaFuncSetAttribute(ilu_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            max_optin - (int)fa.sharedSizeBytes));
    } catch (...) {
        cudaStreamSynchronize(st);
        cleanup();
        delete pc;
        throw;
    }
    cleanup();
    return pc;
}

}  // namespace atk
