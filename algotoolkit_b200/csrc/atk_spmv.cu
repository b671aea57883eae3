// atk_spmv.cu -- SparseMatrixCSC<double> on the device (reference src/Obj/SparseObj.hpp).
//
//   finished CSC arrays (col_ptr,row_indices,values after finalize(), :76-100)
//        |  count rows (atomics) -> exclusive scan -> scatter -> per-row sort by CSC position
//        v
//   CSR with each row's entries in ascending (column, CSC position) order
//        |  row lengths -> (optional sigma-window sort) -> slice widths -> scan -> fill
//        v
//   SELL-32-sigma: slices of 32 rows stored column-major, one thread per row
//
// Why SELL and one thread per row: the reference matvec (:131-152) adds val*x[col] into y[row] for col ascending,
// starting from 0.0, so the bit pattern of y[row] is fixed by a strictly sequential sum over the row.  A thread
// that walks its own row in order reproduces it exactly, and the column-major slice makes those walks coalesced.
// In the CSR format short rows go through the CSR-stream kernel (bit-exact, see there); the CSR-vector kernel (several
// lanes per row + shuffle tree) is kept as the throughput alternative for long,
// irregular rows; it is deterministic but sums in a different order.
//
// Algorithmic traffic per SpMV: 12*nnz + 4*(n+1) + 8*m + 8*n bytes.
#include <algorithm>

#include <string>

#include "atk_internal.cuh"

namespace atk {

namespace {

constexpr int kSliceRows = 32;

// ------------------------------------------------------------------------------------------ int32 exclusive scan
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__global__ void __launch_bounds__(kScanThreads) scan_tile_kernel(const int* __restrict__ in, int* __restrict__ out,
                                                                 int* __restrict__ tile_sums, int64_t n) {
    __shared__ int warp_sums[kScanThreads / 32];
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    int v[kScanItems];
    int local = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        local += v[k];
    }
    // inclusive scan of `local` across the block
    int incl = local;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    int warp_off = 0;
    for (int w = 0; w < warp; ++w) warp_off += warp_sums[w];
    int run = warp_off + incl - local;  // exclusive prefix of this thread inside the tile
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
    if (threadIdx.x == kScanThreads - 1) tile_sums[blockIdx.x] = warp_off + incl;
}

__global__ void __launch_bounds__(kScanThreads) scan_add_kernel(int* __restrict__ out, const int* __restrict__ tile_prefix,
                                                                int64_t n) {
    const int add = tile_prefix[blockIdx.x];
    const int64_t base = (int64_t)blockIdx.x * kScanTile;
    for (int k = threadIdx.x; k < kScanTile; k += kScanThreads)
        if (base + k < n) out[base + k] += add;
}

// out[i] = sum_{t<i} in[t], i in [0,n).  in and out may alias.
void exclusive_scan_i32(Context* ctx, const int* in, int* out, int64_t n) {
    if (n <= 0) return;
    const int64_t tiles = (n + kScanTile - 1) / kScanTile;
    int* tile_sums = nullptr;
    ATK_CUDA_CHECK(cudaMalloc(&tile_sums, sizeof(int) * (size_t)tiles));
    scan_tile_kernel<<<(unsigned)tiles, kScanThreads, 0, ctx->stream>>>(in, out, tile_sums, n);
    ATK_CUDA_CHECK(cudaGetLastError());
    count_launch(ctx);
    if (tiles > 1) {
        exclusive_scan_i32(ctx, tile_sums, tile_sums, tiles);
        scan_add_kernel<<<(unsigned)tiles, kScanThreads, 0, ctx->stream>>>(out, tile_sums, n);
        ATK_CUDA_CHECK(cudaGetLastError());
        count_launch(ctx);
    }
    ATK_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    ATK_CUDA_CHECK(cudaFree(tile_sums));
}

// ------------------------------------------------------------------------------------------ CSC -> CSR
__global__ void count_rows_kernel(int64_t nnz, const int* __restrict__ row_idx, int* __restrict__ row_cnt, int n_rows,
                                  int* __restrict__ bad) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < nnz; p += stride) {
        const int r = row_idx[p];
        if (r < 0 || r >= n_rows) { atomicOr(bad, 1); continue; }
        atomicAdd(row_cnt + r, 1);
    }
}

// col_ptr must be a valid CSC column pointer array before anything scatters through it: first entry 0, non-decreasing,
// last entry nnz (bad |= 2 otherwise)
__global__ void check_col_ptr_kernel(int n_cols, int64_t nnz, const int* __restrict__ col_ptr, int* __restrict__ bad) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c <= n_cols; c += stride) {
        const int v = col_ptr[c];
        bool ok = v >= 0 && (int64_t)v <= nnz;
        if (c == 0) ok = ok && v == 0;
        if (c == n_cols) ok = ok && (int64_t)v == nnz;
        if (c < n_cols) ok = ok && col_ptr[c + 1] >= v;
        if (!ok) atomicOr(bad, 2);
    }
}

// column index of every CSC position: 8 lanes per column
__global__ void expand_cols_kernel(int n_cols, const int* __restrict__ col_ptr, int* __restrict__ col_of_pos) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t c = g >> 3;
    const int sub = (int)(g & 7);
    if (c >= n_cols) return;
    const int e = col_ptr[c + 1];
    for (int k = col_ptr[c] + sub; k < e; k += 8) col_of_pos[k] = (int)c;
}

__global__ void scatter_pos_kernel(int64_t nnz, const int* __restrict__ row_idx, const int* __restrict__ row_ptr,
                                   int* __restrict__ fill, int* __restrict__ pos_of_slot) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < nnz; p += stride) {
        const int r = row_idx[p];
        const int s = atomicAdd(fill + r, 1);
        pos_of_slot[row_ptr[r] + s] = (int)p;
    }
}

// One warp per row: order the row's CSC positions ascending (= ascending column, ties in stored order) by rank
// counting, then gather column and value.  Rows of <= 32 entries are ranked with shuffles.
__global__ void __launch_bounds__(256) order_rows_kernel(int n_rows, const int* __restrict__ row_ptr,
                                                         const int* __restrict__ pos_of_slot,
                                                         const int* __restrict__ col_of_pos, const double* __restrict__ vals,
                                                         int* __restrict__ csr_col, double* __restrict__ csr_val) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = warp0; row < n_rows; row += nwarps) {
        const int s = row_ptr[row], e = row_ptr[row + 1], len = e - s;
        if (len <= 32) {
            const int key = (lane < len) ? pos_of_slot[s + lane] : 0x7fffffff;
            int rank = 0;
#pragma unroll
            for (int m = 0; m < 32; ++m) {
                const int other = __shfl_sync(0xffffffffu, key, m);
                rank += (other < key) ? 1 : 0;
            }
            if (lane < len) {
                csr_col[s + rank] = col_of_pos[key];
                csr_val[s + rank] = vals[key];
            }
        } else {
            for (int a = lane; a < len; a += 32) {
                const int key = pos_of_slot[s + a];
                int rank = 0;
                for (int b = 0; b < len; ++b) rank += (pos_of_slot[s + b] < key) ? 1 : 0;
                csr_col[s + rank] = col_of_pos[key];
                csr_val[s + rank] = vals[key];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------ CSR -> SELL-32-sigma
// rank of every row inside its sigma-window, longer rows first, ties by row index (a deterministic permutation)
constexpr int kSigma = 1024;
__global__ void __launch_bounds__(kSigma) sigma_rank_kernel(int n_rows, const int* __restrict__ row_ptr,
                                                            int* __restrict__ perm) {
    __shared__ int len[kSigma];
    const int base = blockIdx.x * kSigma;
    const int r = base + threadIdx.x;
    const int cnt = min(kSigma, n_rows - base);
    len[threadIdx.x] = (r < n_rows) ? row_ptr[r + 1] - row_ptr[r] : -1;
    __syncthreads();
    if (r >= n_rows) return;
    const int mine = len[threadIdx.x];
    int rank = 0;
    for (int t = 0; t < cnt; ++t) {
        const int o = len[t];
        rank += (o > mine || (o == mine && t < (int)threadIdx.x)) ? 1 : 0;
    }
    perm[base + rank] = r;
}

__global__ void slice_width_kernel(int n_rows, int n_slices, const int* __restrict__ row_ptr, const int* __restrict__ perm,
                                   int* __restrict__ width) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_slices) return;
    int w = 0;
    for (int l = 0; l < kSliceRows; ++l) {
        const int t = s * kSliceRows + l;
        if (t >= n_rows) break;
        const int r = perm ? perm[t] : t;
        w = max(w, row_ptr[r + 1] - row_ptr[r]);
    }
    width[s] = w;
}

__global__ void sell_fill_kernel(int n_rows, const int* __restrict__ row_ptr, const int* __restrict__ csr_col,
                                 const double* __restrict__ csr_val, const int* __restrict__ perm,
                                 const int* __restrict__ wscan, int* __restrict__ sell_col, double* __restrict__ sell_val) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t padded_rows = ((int64_t)n_rows + kSliceRows - 1) / kSliceRows * kSliceRows;
    if (t >= padded_rows) return;
    const int s = (int)(t / kSliceRows), lane = (int)(t % kSliceRows);
    const int w = wscan[s + 1] - wscan[s];
    const int64_t off = (int64_t)wscan[s] * kSliceRows + lane;
    int rs = 0, len = 0;
    if (t < n_rows) {
        const int r = perm ? perm[t] : (int)t;
        rs = row_ptr[r];
        len = row_ptr[r + 1] - rs;
    }
    for (int k = 0; k < w; ++k) {
        const bool real = k < len;
        sell_col[off + (int64_t)k * kSliceRows] = real ? csr_col[rs + k] : -1;  // -1: padding, never multiplied
        sell_val[off + (int64_t)k * kSliceRows] = real ? csr_val[rs + k] : 0.0;
    }
}

// ------------------------------------------------------------------------------------------ SpMV kernels
// SELL: one warp per slice (persistent, grid-stride over slices), one thread per row.
// acc starts at 0.0 and takes val*x[col] in ascending column order; a column whose x is exactly 0 is skipped,
// as the reference does (SparseObj.hpp:141).
// Row-block partitioned operator (distributed context): column c of the local matrix lives in one of three arrays -
// [0, h_lo) the tail of the previous rank's block (received), [h_lo, h_lo + n_mid) this rank's x, the rest the head of
// the next rank's block (received).
struct DistX {
    const double* lo;
    const double* hi;
    int h_lo;
    int n_mid;
};
// PUSH (distributed contexts in peer-memory mode): the whole apply in ONE launch.  The first `n_push` CTAs start by gathering
// the x entries the other ranks need and storing them straight into their receive buffers (NVLink), fence, and the last
// of them releases the apply's sequence number into every partner's flag word; every warp then walks the slices that
// need no external entry first and waits on the LOCAL flags only before its first slice that does.  The grid never
// exceeds the resident capacity, so the pushing CTAs of every rank are running before anybody waits.
struct SellPush {
    GatherTable tb;
    WaitTable wt;
    const int* send_idx = nullptr;
    unsigned* ticket = nullptr;
    unsigned long long seq = 0;
    int n_push = 0;
    int s_int_begin = 0, s_int_end = 0;
};
template <bool DOT, bool DIST, bool PUSH = false>
__global__ void __launch_bounds__(kReduceBlock) sell_spmv_kernel(int n_rows, int n_slices, const int* __restrict__ wscan,
                                                                 const int* __restrict__ sell_col,
                                                                 const double* __restrict__ sell_val,
                                                                 const int* __restrict__ perm, const double* __restrict__ x,
                                                                 double* __restrict__ y, double* partials, unsigned* ticket,
                                                                 double* dot_out, const int* skip_flag, DistX dx, int s_begin,
                                                                 int s_end, int partial_base, int is_final,
                                                                 SellPush sp = SellPush()) {
    if (skip_flag && *skip_flag) return;
    const int lane = threadIdx.x & 31;
    const int warp0 = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int nwarps = (int)(((int64_t)gridDim.x * blockDim.x) >> 5);
    double dacc = 0.0;
    constexpr int U = 8;  // entries of a row handled per batch: all 2U matrix loads are issued together, then all U gathers
    if constexpr (PUSH) {
        if ((int)blockIdx.x < sp.n_push) {
            const int total = sp.tb.begin[sp.tb.n_dst];
            for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += sp.n_push * blockDim.x) {
                int d = 0;
                while (t >= sp.tb.begin[d + 1]) ++d;
                sp.tb.dst[d][t - sp.tb.begin[d]] = x[sp.send_idx[t]];
            }
            __threadfence_system();
            if (last_block_done_among(sp.ticket, (unsigned)sp.n_push)) {
                if ((int)threadIdx.x < sp.tb.n_dst && sp.tb.flag[threadIdx.x]) {
                    __threadfence_system();
                    st_release_sys(sp.tb.flag[threadIdx.x], sp.seq);
                }
            }
        }
    }
    const int n_int = PUSH ? sp.s_int_end - sp.s_int_begin : 0;
    bool waited = false;
    for (int t = s_begin + warp0; t < s_end; t += nwarps) {
        int s = t;
        if constexpr (PUSH) {  // logical order: slices without external entries, then the lower ones, then the upper ones
            if (t < n_int) {
                s = sp.s_int_begin + t;
            } else {
                if (!waited) {
                    if (lane < sp.wt.n) {
                        unsigned spins = 0;
                        unsigned long long t0 = 0;
                        while (ld_acquire_sys(sp.wt.flag[lane]) < sp.seq) {
                            if ((++spins & 0xfffu) == 0) {
                                const unsigned long long now = global_timer_ns();
                                if (t0 == 0) t0 = now;
                                if (now - t0 > kPeerWaitNs) asm volatile("trap;");  // a partner died or fell out of step
                            }
                        }
                    }
                    __syncwarp();
                    waited = true;
                }
                const int u = t - n_int;
                s = u < sp.s_int_begin ? u : sp.s_int_end + (u - sp.s_int_begin);
            }
        }
        const int w0 = wscan[s], w = wscan[s + 1] - w0;
        const int64_t off = (int64_t)w0 * kSliceRows + lane;
        double acc = 0.0;
        for (int kb = 0; kb < w; kb += U) {
            int c[U];
            double v[U], xv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const bool in = kb + u < w;
                c[u] = in ? __ldg(sell_col + off + (int64_t)(kb + u) * kSliceRows) : -1;
                v[u] = in ? __ldg(sell_val + off + (int64_t)(kb + u) * kSliceRows) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {  // unconditional (padding reads the first local entry, then ignored)
                if constexpr (DIST) {
                    const int cc = c[u] < 0 ? dx.h_lo : c[u];
                    const double* base = cc < dx.h_lo ? dx.lo : (cc < dx.h_lo + dx.n_mid ? x - dx.h_lo : dx.hi - dx.h_lo - dx.n_mid);
                    xv[u] = base[cc];
                } else {
                    xv[u] = x[c[u] < 0 ? 0 : c[u]];
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                // ascending column order; padding and exact-zero x entries contribute nothing (SparseObj.hpp:141)
                const double t = add_rn(acc, mul_rn(v[u], xv[u]));
                acc = (c[u] >= 0 && xv[u] != 0.0) ? t : acc;
            }
        }
        const int tr = s * kSliceRows + lane;
        if (tr < n_rows) {
            const int r = perm ? perm[tr] : tr;
            y[r] = acc;
            if constexpr (DOT) dacc = add_rn(dacc, mul_rn(x[r], acc));
        }
    }
    if constexpr (DOT) {
        double total;
        if (grid_sum_last_block_chained<kReduceBlock>(dacc, partials, partial_base, is_final, ticket, &total)) {
            if (linear_thread_id() == 0) *dot_out = total;
        }
    }
}

// CSR-vector: LANES lanes cooperate on a row and a group of lanes works on R rows at a time, so that 2R matrix loads
// and R gathers are in flight per lane before anything is consumed; fixed shuffle tree per row (deterministic, but not
// the reference's summation order).
template <int LANES, bool DOT>
__global__ void __launch_bounds__(kReduceBlock) csr_vector_spmv_kernel(int n_rows, const int* __restrict__ row_ptr,
                                                                       const int* __restrict__ csr_col,
                                                                       const double* __restrict__ csr_val,
                                                                       const double* __restrict__ x, double* __restrict__ y,
                                                                       double* partials, unsigned* ticket, double* dot_out,
                                                                       const int* skip_flag) {
    if (skip_flag && *skip_flag) return;
    constexpr int R = 4;
    const int64_t gt = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int sub = (int)(gt % LANES);
    const int64_t group0 = gt / LANES, ngroups = ((int64_t)gridDim.x * blockDim.x) / LANES;
    const int64_t rows_per_pass = ngroups * R;
    const int64_t n_pass = ((int64_t)n_rows + rows_per_pass - 1) / rows_per_pass;  // same trip count for every lane
    double dacc = 0.0;
    for (int64_t pass = 0; pass < n_pass; ++pass) {
        const int64_t r0 = pass * rows_per_pass + group0 * R;
        double acc[R];
        int kb[R], ke[R];
        int maxlen = 0;
#pragma unroll
        for (int q = 0; q < R; ++q) {
            const bool valid = r0 + q < n_rows;
            kb[q] = valid ? row_ptr[r0 + q] : 0;
            ke[q] = valid ? row_ptr[r0 + q + 1] : 0;
            acc[q] = 0.0;
            maxlen = max(maxlen, ke[q] - kb[q]);
        }
        for (int o = sub; o < maxlen; o += LANES) {
            int c[R];
            double v[R], xv[R];
#pragma unroll
            for (int q = 0; q < R; ++q) {
                const bool in = kb[q] + o < ke[q];
                c[q] = in ? __ldg(csr_col + kb[q] + o) : -1;
                v[q] = in ? __ldg(csr_val + kb[q] + o) : 0.0;
            }
#pragma unroll
            for (int q = 0; q < R; ++q) xv[q] = x[c[q] < 0 ? 0 : c[q]];
#pragma unroll
            for (int q = 0; q < R; ++q) {
                const double t = add_rn(acc[q], mul_rn(v[q], xv[q]));
                acc[q] = (c[q] >= 0 && xv[q] != 0.0) ? t : acc[q];
            }
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < R; ++q) {
#pragma unroll
            for (int o = LANES / 2; o > 0; o >>= 1) acc[q] = add_rn(acc[q], __shfl_xor_sync(0xffffffffu, acc[q], o));
        }
        if (sub == 0) {
#pragma unroll
            for (int q = 0; q < R; ++q) {
                if (r0 + q < n_rows) {
                    y[r0 + q] = acc[q];
                    if constexpr (DOT) dacc = add_rn(dacc, mul_rn(x[r0 + q], acc[q]));
                }
            }
        }
    }
    if constexpr (DOT) {
        double total;
        if (grid_sum_last_block<kReduceBlock>(dacc, partials, ticket, &total)) {
            if (linear_thread_id() == 0) *dot_out = total;
        }
    }
}

// CSR-stream (the CSR-format kernel for short rows): a CTA takes a run of consecutive rows holding at most kStreamCap
// entries, streams their values / column indices in with fully coalesced loads (8 entries in flight per thread), gathers
// x, and leaves the rounded products in shared memory; then one thread per row adds its row's products from 0.0 in
// ascending column order.  Per row that is exactly the CSC matvec's sequence of operations (an exact-zero x entry adds
// +0.0, which leaves any running sum unchanged, as skipping it does): bit-identical to the SELL kernel and the reference,
// unlike the lane-tree of CSR-vector, with no padding and no format conversion beyond CSR.
// blk_row[b] = first row whose first entry is at or after b*T, T = kStreamCap - (longest row), so that block b's entries
// [row_ptr[blk_row[b]], row_ptr[blk_row[b+1]]) always fit.
constexpr int kStreamCap = 4096;   // products per CTA (32 KB of shared memory)
constexpr int kStreamBlock = 256;

__global__ void csr_stream_blocks_kernel(int n_rows, int n_blocks, int target, const int* __restrict__ row_ptr, int* __restrict__ blk_row) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > n_blocks) return;
    if (b == n_blocks) {
        blk_row[b] = n_rows;
        return;
    }
    const int64_t want = (int64_t)b * target;
    int lo = 0, hi = n_rows;  // first row r in [0, n_rows] with row_ptr[r] >= want
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (row_ptr[mid] >= want) hi = mid;
        else lo = mid + 1;
    }
    blk_row[b] = lo;
}
__global__ void max_row_len_kernel(int n_rows, const int* __restrict__ row_ptr, int* __restrict__ out) {
    int m = 0;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += gridDim.x * blockDim.x) m = max(m, row_ptr[r + 1] - row_ptr[r]);
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(out, m);
}

template <bool DOT>
__global__ void __launch_bounds__(kStreamBlock) csr_stream_spmv_kernel(int n_blocks, const int* __restrict__ blk_row,
                                                                       const int* __restrict__ row_ptr,
                                                                       const int* __restrict__ csr_col,
                                                                       const double* __restrict__ csr_val,
                                                                       const double* __restrict__ x, double* __restrict__ y,
                                                                       double* partials, unsigned* ticket, double* dot_out,
                                                                       const int* skip_flag) {
    if (skip_flag && *skip_flag) return;
    extern __shared__ __align__(16) double prod[];
    double dacc = 0.0;
    constexpr int U = 8;
    // block metadata one trip ahead, so that the chain blk_row -> row_ptr -> entries is not paid serially per block
    int b = blockIdx.x;
    int r0 = 0, r1 = 0, e0 = 0, e1 = 0;
    if (b < n_blocks) {
        r0 = blk_row[b];
        r1 = blk_row[b + 1];
        e0 = row_ptr[r0];
        e1 = row_ptr[r1];
    }
    while (b < n_blocks) {
        const int bn = b + gridDim.x;
        int nr0 = 0, nr1 = 0;
        if (bn < n_blocks) {
            nr0 = blk_row[bn];
            nr1 = blk_row[bn + 1];
        }
        // this thread's first row of the summation phase: its bounds are requested before the entries are
        int r = r0 + threadIdx.x;
        int k0 = 0, k1 = 0;
        if (r < r1) {
            k0 = row_ptr[r];
            k1 = row_ptr[r + 1];
        }
        for (int e = e0 + threadIdx.x; e < e1; e += kStreamBlock * U) {
            int c[U];
            double v[U], xv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int q = e + u * kStreamBlock;
                c[u] = q < e1 ? __ldg(csr_col + q) : -1;
                v[u] = q < e1 ? __ldg(csr_val + q) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) xv[u] = x[c[u] < 0 ? 0 : c[u]];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int q = e + u * kStreamBlock;
                if (q < e1) prod[q - e0] = xv[u] != 0.0 ? mul_rn(v[u], xv[u]) : 0.0;
            }
        }
        int ne0 = 0, ne1 = 0;
        if (bn < n_blocks) {
            ne0 = row_ptr[nr0];
            ne1 = row_ptr[nr1];
        }
        __syncthreads();
        while (r < r1) {
            const int rn = r + kStreamBlock;
            int nk0 = 0, nk1 = 0;
            if (rn < r1) {  // bounds of the next row before this row's dependent chain of adds
                nk0 = row_ptr[rn];
                nk1 = row_ptr[rn + 1];
            }
            double acc = 0.0;
            for (int k = k0 - e0; k < k1 - e0; ++k) acc = add_rn(acc, prod[k]);
            y[r] = acc;
            if constexpr (DOT) dacc = add_rn(dacc, mul_rn(x[r], acc));
            r = rn;
            k0 = nk0;
            k1 = nk1;
        }
        __syncthreads();  // prod[] free again
        b = bn;
        r0 = nr0;
        r1 = nr1;
        e0 = ne0;
        e1 = ne1;
    }
    if constexpr (DOT) {
        double total;
        if (grid_sum_last_block<kStreamBlock>(dacc, partials, ticket, &total)) {
            if (linear_thread_id() == 0) *dot_out = total;
        }
    }
}

// CSR-stream, second form (default for the CSR format): the MATRIX entries are staged in shared memory instead of the
// products, and the staging is WARP-private.  A warp owns 32
// consecutive rows: it copies their entries (values + column indices, one contiguous range of the CSR arrays) into its own
// slice of shared memory with 16-byte cp.async, waits for its own copies (__syncwarp, no block barrier anywhere), then
// every lane walks its row through shared memory (row starts an odd stride apart: no bank conflicts on stencil-like rows)
// and gathers x for the same tap position of 32 consecutive rows at a time - a few sectors per request.  Warps run ahead of
// one another freely, up to 64 per SM, so the DRAM latency of one warp's stream hides behind the gathers of the others.
// Products are added from 0.0 in ascending column order in registers: the CSC matvec's exact sequence per row.
// CAP = entries a warp can stage; a 32-row group with more entries is walked lane by lane straight from global memory.
__device__ __forceinline__ void cp_async16_g2s(void* smem, const void* gmem) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
constexpr int kWarpStreamWarps = 8;
template <int CAP, bool DOT>
__global__ void __launch_bounds__(kWarpStreamWarps * 32) csr_warp_stream_spmv_kernel(int n_rows, int64_t nnz,
                                                                                   const int* __restrict__ row_ptr,
                                                                                   const int* __restrict__ csr_col,
                                                                                   const double* __restrict__ csr_val,
                                                                                   const double* __restrict__ x, double* __restrict__ y,
                                                                                   double* partials, unsigned* ticket, double* dot_out,
                                                                                   const int* skip_flag) {
    if (skip_flag && *skip_flag) return;
    extern __shared__ __align__(16) double smem3[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* sv = smem3 + (size_t)warp * (CAP + 4);
    int* sc = reinterpret_cast<int*>(smem3 + (size_t)kWarpStreamWarps * (CAP + 4)) + (size_t)warp * (CAP + 4);
    const int n_groups = (n_rows + 31) >> 5;
    const int total_warps = gridDim.x * kWarpStreamWarps;
    const int64_t nnz4 = nnz & ~(int64_t)3;
    double dacc = 0.0;
    constexpr int U = 8;
    for (int grp = blockIdx.x * kWarpStreamWarps + warp; grp < n_groups; grp += total_warps) {
        const int r = (grp << 5) + lane;
        const int k0 = row_ptr[min(r, n_rows)], k1 = row_ptr[min(r + 1, n_rows)];
        const int e0 = __shfl_sync(0xffffffffu, k0, 0), e1 = __shfl_sync(0xffffffffu, k1, 31);
        double acc = 0.0;
        if (e1 - (e0 & ~3) <= CAP + 4 - 4) {
            const int e0a = e0 & ~3;
            const int n4 = (e1 - e0a + 3) >> 2;
            for (int q = lane; q < n4; q += 32) {
                const int64_t idx = (int64_t)e0a + 4 * q;
                if (idx + 4 <= nnz4) {
                    cp_async16_g2s(sc + 4 * q, csr_col + idx);
                    cp_async16_g2s(sv + 4 * q, csr_val + idx);
                    cp_async16_g2s(sv + 4 * q + 2, csr_val + idx + 2);
                } else {
                    for (int u = 0; u < 4; ++u)
                        if (idx + u < nnz) {
                            sc[4 * q + u] = csr_col[idx + u];
                            sv[4 * q + u] = csr_val[idx + u];
                        }
                }
            }
            asm volatile("cp.async.commit_group;\n" ::);
            asm volatile("cp.async.wait_group 0;\n" ::);
            __syncwarp();
            for (int kb = k0 - e0a; kb < k1 - e0a; kb += U) {
                int c[U];
                double v[U], xv[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const bool in = kb + u < k1 - e0a;
                    c[u] = in ? sc[kb + u] : -1;
                    v[u] = in ? sv[kb + u] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < U; ++u) xv[u] = x[c[u] < 0 ? 0 : c[u]];
#pragma unroll
                for (int u = 0; u < U; ++u) {  // ascending column order; exact-zero x entries contribute nothing (SparseObj.hpp:141)
                    const double t = add_rn(acc, mul_rn(v[u], xv[u]));
                    acc = (c[u] >= 0 && xv[u] != 0.0) ? t : acc;
                }
            }
            __syncwarp();  // the slice is restaged by the next trip
        } else {
            for (int k = k0; k < k1; ++k) {
                const double xv = x[csr_col[k]];
                const double t = add_rn(acc, mul_rn(csr_val[k], xv));
                acc = xv != 0.0 ? t : acc;
            }
        }
        if (r < n_rows) {
            y[r] = acc;
            if constexpr (DOT) dacc = add_rn(dacc, mul_rn(x[r], acc));
        }
    }
    if constexpr (DOT) {
        double total;
        if (grid_sum_last_block<kWarpStreamWarps * 32>(dacc, partials, ticket, &total)) {
            if (linear_thread_id() == 0) *dot_out = total;
        }
    }
}

// ------------------------------------------------------------------------------------------ the operator
struct AssembledMatrix : public Operator {
    int n_rows = 0;
    atk_format format = ATK_FORMAT_SELL;
    // CSR
    int* row_ptr = nullptr;
    int* csr_col = nullptr;
    double* csr_val = nullptr;
    // SELL
    int n_slices = 0;
    int* wscan = nullptr;  // [n_slices+1] exclusive scan of slice widths
    int* perm = nullptr;   // sorted position -> row, or null (identity)
    int* sell_col = nullptr;
    double* sell_val = nullptr;
    int64_t sell_elems = 0;
    int lanes = 8;
    // CSR-stream (CSR format, short rows)
    int* blk_row = nullptr;  // [stream_blocks + 1]
    int stream_blocks = 0;   // 0: rows too long for the stream kernel -> CSR-vector
    // row-block partition on a distributed context (SELL only), any sparsity: local columns = [n_rows own | n_ext external
    // entries in ascending global column order]; the external entries of an apply arrive through PeerGather
    bool dist = false;
    PeerGather pg;
    std::vector<int> ext_cols_host;      // global column of every external entry (for export)
    int push_per_sm_dot = 0, push_per_sm_plain = 0;  // resident CTAs per SM of the single-launch apply (evaluated once)
    int s_int_begin = 0, s_int_end = 0;  // slices whose rows touch no external entry: they run while the exchange is in flight

    ~AssembledMatrix() override {
        cudaFree(row_ptr); cudaFree(csr_col); cudaFree(csr_val);
        cudaFree(wscan); cudaFree(perm); cudaFree(sell_col); cudaFree(sell_val); cudaFree(blk_row);
        pg.release();
    }
    bool has_csr() const override { return true; }
    bool device_csr(const int** rp, const int** ci, const double** v) const override {
        if (dist) return false;  // distributed blocks number their columns locally
        *rp = row_ptr;
        *ci = csr_col;
        *v = csr_val;
        return true;
    }
    void export_csr(int* h_row_ptr, int* h_col, double* h_val) const override {
        ATK_CUDA_CHECK(cudaMemcpy(h_row_ptr, row_ptr, sizeof(int) * ((size_t)n_rows + 1), cudaMemcpyDeviceToHost));
        if (nnz > 0) {
            ATK_CUDA_CHECK(cudaMemcpy(h_col, csr_col, sizeof(int) * (size_t)nnz, cudaMemcpyDeviceToHost));
            ATK_CUDA_CHECK(cudaMemcpy(h_val, csr_val, sizeof(double) * (size_t)nnz, cudaMemcpyDeviceToHost));
            if (dist)  // back to global column indices
                for (int64_t k = 0; k < nnz; ++k)
                    h_col[k] = h_col[k] < n_rows ? h_col[k] + (int)row_begin : ext_cols_host[(size_t)(h_col[k] - n_rows)];
        }
    }

    void apply(const double* x, double* y, int dot_slot, const int* skip_flag) override {
        ATK_NVTX(format == ATK_FORMAT_SELL ? "atk::spmv_sell32" : "atk::spmv_csr");
        Context* c = ctx;
        double* dot_out = dot_slot >= 0 ? c->slot(dot_slot) + c->rank : nullptr;
        const int grid = c->sm_count * 8;
        if (format == ATK_FORMAT_SELL) {
            const bool dot = dot_slot >= 0;
            if (dist && !std::getenv("ATK_APPLY_SPLIT")) {
                // peer mode: ONE launch gathers and stores for the partners, computes, and waits where it must
                SellPush sp;
                const double* ext_in = nullptr;
                if (pg.begin_inline(&sp.tb, &sp.wt, &sp.seq, &ext_in)) {
                    sp.send_idx = pg.d_send_idx;
                    sp.ticket = pg.d_ticket;
                    sp.s_int_begin = s_int_begin;
                    sp.s_int_end = s_int_end;
                    int& per_sm = dot ? push_per_sm_dot : push_per_sm_plain;  // per operator (= per device)
                    if (per_sm == 0) {
                        if (dot) ATK_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sell_spmv_kernel<true, true, true>, kReduceBlock, 0));
                        else ATK_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sell_spmv_kernel<false, true, true>, kReduceBlock, 0));
                        per_sm = std::max(per_sm, 1);
                    }
                    const int cap_g = std::min(grid, per_sm * c->sm_count);  // every CTA resident: pushers run before anybody waits
                    const int g = (int)std::max<int64_t>(1, std::min<int64_t>(cap_g, ((int64_t)n_slices * 32 + kReduceBlock - 1) / kReduceBlock));
                    sp.n_push = std::max(1, std::min(g, (pg.n_send + 2 * kReduceBlock - 1) / (2 * kReduceBlock)));
                    const DistX dxi{nullptr, ext_in, 0, n_rows};
                    if (dot)
                        sell_spmv_kernel<true, true, true><<<g, kReduceBlock, 0, c->stream>>>(n_rows, n_slices, wscan, sell_col, sell_val, perm, x, y,
                                                                                              c->d_partials, c->d_tickets + 1, dot_out, skip_flag,
                                                                                              dxi, 0, n_slices, 0, 1, sp);
                    else
                        sell_spmv_kernel<false, true, true><<<g, kReduceBlock, 0, c->stream>>>(n_rows, n_slices, wscan, sell_col, sell_val, perm, x, y,
                                                                                               nullptr, nullptr, nullptr, skip_flag, dxi, 0,
                                                                                               n_slices, 0, 1, sp);
                    ATK_CUDA_CHECK(cudaGetLastError());
                    count_launch(c);
                    if (dot) allgather_slot(c, dot_slot, c->stream);
                    return;
                }
            }
            // distributed: the owners start sending the external entries of this apply (communication stream)
            const double* ext = dist ? pg.begin(x) : nullptr;
            const DistX dx{nullptr, ext, 0, n_rows};
            int pbase = 0;
            // (measured at 2 GPUs, 256^3: the persistent grid of the interior piece fills every SM, so the one-CTA transfer
            // kernel mostly runs after it - 0.173 ms per apply against 0.176 ms without the split; shrinking the grid to
            // leave room, or launching the piece as short CTAs, made the piece itself slower: 0.201 / 0.185 ms)
            auto piece = [&](int sb, int se, bool final_piece, bool beside_exchange = false) {
                if (se <= sb) return;
                (void)beside_exchange;
                const int cap_g = grid;
                const int g = (int)std::max<int64_t>(1, std::min<int64_t>(cap_g, ((int64_t)(se - sb) * 32 + kReduceBlock - 1) / kReduceBlock));
#define ATK_SELL(D, X)                                                                                                            \
    sell_spmv_kernel<D, X><<<g, kReduceBlock, 0, c->stream>>>(n_rows, n_slices, wscan, sell_col, sell_val, perm, x, y,           \
                                                              D ? c->d_partials : nullptr, D ? c->d_tickets + 1 : nullptr,      \
                                                              D ? dot_out : nullptr, skip_flag, dx, sb, se, pbase, final_piece ? 1 : 0)
                if (dist) { if (dot) ATK_SELL(true, true); else ATK_SELL(false, true); }
                else      { if (dot) ATK_SELL(true, false); else ATK_SELL(false, false); }
#undef ATK_SELL
                pbase += g;
                count_launch(c);
            };
            if (!dist) {
                piece(0, n_slices, true);
            } else {
                // the slices that need no external entry run while the exchange is in flight
                const bool has_lo = s_int_begin > 0, has_hi = s_int_end < n_slices;
                piece(s_int_begin, s_int_end, !has_lo && !has_hi, true);
                pg.wait();
                piece(0, s_int_begin, !has_hi);
                piece(s_int_end, n_slices, true);
            }
            ATK_CUDA_CHECK(cudaGetLastError());
            if (dot) allgather_slot(c, dot_slot, c->stream);
            return;
        } else if (stream_blocks > 0 && !std::getenv("ATK_CSR_STREAM_V1")) {
            // warp-private staging: CAP from the mean row length (a group of 32 rows must fit with ~15 % to spare)
            const double per_group = n_rows > 0 ? 32.0 * (double)nnz / n_rows * 1.15 : 0.0;
            const int cap_sel = per_group <= 256 ? 256 : per_group <= 512 ? 512 : per_group <= 1024 ? 1024 : 2048;
            const bool dot = dot_slot >= 0;
            auto go = [&](auto kern, int cap_v) {
                const size_t smem = (size_t)kWarpStreamWarps * (cap_v + 4) * (sizeof(double) + sizeof(int));
                ATK_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                int per_sm = 1;
                ATK_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kWarpStreamWarps * 32, smem));
                const int groups = (n_rows + 31) / 32;
                const int g = std::max(1, std::min((groups + kWarpStreamWarps - 1) / kWarpStreamWarps, c->sm_count * std::max(per_sm, 1)));
                kern<<<g, kWarpStreamWarps * 32, smem, c->stream>>>(n_rows, nnz, row_ptr, csr_col, csr_val, x, y,
                                                                    dot ? c->d_partials : nullptr, dot ? c->d_tickets + 1 : nullptr,
                                                                    dot ? dot_out : nullptr, skip_flag);
            };
#define ATK_WS(CAPV) \
    if (dot) go(csr_warp_stream_spmv_kernel<CAPV, true>, CAPV); \
    else go(csr_warp_stream_spmv_kernel<CAPV, false>, CAPV);
            switch (cap_sel) {
                case 256: ATK_WS(256) break;
                case 512: ATK_WS(512) break;
                case 1024: ATK_WS(1024) break;
                default: ATK_WS(2048) break;
            }
#undef ATK_WS
        } else if (stream_blocks > 0) {
            const int g = std::min(stream_blocks, c->sm_count * 32);
            constexpr size_t smem = sizeof(double) * kStreamCap;
            if (dot_slot >= 0)
                csr_stream_spmv_kernel<true><<<g, kStreamBlock, smem, c->stream>>>(stream_blocks, blk_row, row_ptr, csr_col, csr_val,
                                                                                 x, y, c->d_partials, c->d_tickets + 1, dot_out,
                                                                                 skip_flag);
            else
                csr_stream_spmv_kernel<false><<<g, kStreamBlock, smem, c->stream>>>(stream_blocks, blk_row, row_ptr, csr_col,
                                                                                  csr_val, x, y, nullptr, nullptr, nullptr,
                                                                                  skip_flag);
        } else {
            const int g = (int)std::max<int64_t>(1, std::min<int64_t>(grid, ((int64_t)n_rows * lanes + kReduceBlock - 1) / kReduceBlock));
#define ATK_CSRV(L)                                                                                                      \
    if (dot_slot >= 0)                                                                                                   \
        csr_vector_spmv_kernel<L, true><<<g, kReduceBlock, 0, c->stream>>>(n_rows, row_ptr, csr_col, csr_val, x, y,        \
                                                                           c->d_partials, c->d_tickets + 1, dot_out,      \
                                                                           skip_flag);                                    \
    else                                                                                                                 \
        csr_vector_spmv_kernel<L, false><<<g, kReduceBlock, 0, c->stream>>>(n_rows, row_ptr, csr_col, csr_val, x, y,       \
                                                                            nullptr, nullptr, nullptr, skip_flag);
            switch (lanes) {
                case 2: ATK_CSRV(2) break;
                case 4: ATK_CSRV(4) break;
                case 8: ATK_CSRV(8) break;
                case 16: ATK_CSRV(16) break;
                default: ATK_CSRV(32) break;
            }
#undef ATK_CSRV
        }
        ATK_CUDA_CHECK(cudaGetLastError());
        count_launch(c);
        if (dot_slot >= 0) allgather_slot(c, dot_slot, c->stream);
    }
};

template <typename T>
T* dalloc(size_t n) {
    T* p = nullptr;
    ATK_CUDA_CHECK(cudaMalloc(&p, sizeof(T) * std::max<size_t>(n, 1)));
    return p;
}

inline int grid_for(int64_t work, int block, int cap) {
    return (int)std::max<int64_t>(1, std::min<int64_t>((work + block - 1) / block, cap));
}

}  // namespace

// Takes DEVICE CSC arrays and builds the operator.  The CSC arrays are not retained.
// ---- row-block slicing of a global CSR (distributed contexts)
__global__ void slice_row_ptr_kernel(int n_loc, int r0, const int* __restrict__ g_row_ptr, int* __restrict__ row_ptr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n_loc) row_ptr[i] = g_row_ptr[r0 + i] - g_row_ptr[r0];
}
// flag[c] = 1 for every column outside [row_lo, row_hi) that a local entry references
__global__ void mark_ext_kernel(int64_t nnz, const int* __restrict__ col, int row_lo, int row_hi, int* __restrict__ flag) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += (int64_t)gridDim.x * blockDim.x) {
        const int c = col[k];
        if (c < row_lo || c >= row_hi) flag[c] = 1;
    }
}
__global__ void compact_ext_kernel(int n_cols, const int* __restrict__ flag, const int* __restrict__ pos, int* __restrict__ ext_cols) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n_cols && flag[c]) ext_cols[pos[c]] = c;
}
// global column -> local column: own block first, then the external entries in ascending global order
__global__ void renumber_cols_kernel(int64_t nnz, int row_lo, int row_hi, int n_loc, const int* __restrict__ pos, int* __restrict__ col) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += (int64_t)gridDim.x * blockDim.x) {
        const int c = col[k];
        col[k] = (c >= row_lo && c < row_hi) ? c - row_lo : n_loc + pos[c];
    }
}
// rows that reference an external entry: those in the lower half push out[0] up past them, those in the upper half pull
// out[1] down to them, so that rows [out[0], out[1]) reference none (out starts as {0, n_rows})
__global__ void ext_rows_kernel(int n_rows, const int* __restrict__ row_ptr, const int* __restrict__ col, int n_loc,
                                int* __restrict__ out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    bool ext = false;
    for (int k = row_ptr[r]; k < row_ptr[r + 1]; ++k) ext = ext || col[k] >= n_loc;
    if (!ext) return;
    if (r < n_rows / 2) atomicMax(out, r + 1);
    else atomicMin(out + 1, r);
}
__global__ void shift_rows_kernel(int64_t nnz, int shift, const int* __restrict__ in, int* __restrict__ out, int n_loc,
                                  int* __restrict__ bad) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += (int64_t)gridDim.x * blockDim.x) {
        const int r = in[k] - shift;
        if (r < 0 || r >= n_loc) atomicOr(bad, 1);
        out[k] = r;
    }
}

// row_lo / row_hi: the row block of this rank on a distributed context (row_hi < 0: balanced blocks of rows)
// rows_are_slice: the CSC arrays hold ONLY the entries of rows [row_lo, row_hi), their row indices already relative to
// row_lo (n_rows is then the global row count and the conversion works on row_hi - row_lo local rows).
static Operator* build_from_device_csc(Context* ctx, int n_rows, int n_cols, int64_t nnz, const int* d_col_ptr,
                                       const int* d_row_idx, const double* d_vals, atk_format format, int row_lo = 0,
                                       int row_hi = -1, bool rows_are_slice = false) {
    ATK_REQUIRE(nnz >= 0 && nnz < (int64_t)0x7fffffff, ATK_ERR_INVALID_ARGUMENT, "nnz must be below 2^31");
    const bool dist = ctx->world > 1;
    ATK_NVTX("atk::csc_to_device_format");
    const int n_rows_global = n_rows;
    if (dist) {
        ATK_REQUIRE(format == ATK_FORMAT_SELL, ATK_ERR_INVALID_ARGUMENT, "distributed assembled operators use the SELL format");
        ATK_REQUIRE(n_rows == n_cols, ATK_ERR_INVALID_ARGUMENT, "distributed assembled operators must be square");
        // every rank must own at least one row: a rank without rows would launch nothing and leave a stale partial in the
        // reduction slot that the all-gather then feeds into p.Ap on every rank
        ATK_REQUIRE(n_rows >= ctx->world, ATK_ERR_INVALID_ARGUMENT, "distributed assembled operator: fewer rows than ranks");
        if (row_hi < 0) slab_partition(n_rows, ctx->rank, ctx->world, &row_lo, &row_hi);
        ATK_REQUIRE(row_hi > row_lo && row_lo >= 0 && row_hi <= n_rows, ATK_ERR_INVALID_ARGUMENT,
                    "distributed assembled operator: empty or out-of-range row block");
        if (rows_are_slice) n_rows = row_hi - row_lo;  // the conversion below sees the local rows only
    } else {
        ATK_REQUIRE(!rows_are_slice || (row_lo == 0 && row_hi == n_rows), ATK_ERR_INVALID_ARGUMENT,
                    "a single-GPU context owns every row");
    }
    auto* A = new AssembledMatrix();
    try {
        A->ctx = ctx;
        A->n_rows = n_rows;
        A->n_rows_local = n_rows;
        A->n_cols = n_cols;
        A->nnz = nnz;
        A->format = format;
        cudaStream_t st = ctx->stream;
        const int cap = ctx->sm_count * 16;

        // ---- CSC -> CSR
        A->row_ptr = dalloc<int>((size_t)n_rows + 1);
        A->csr_col = dalloc<int>((size_t)nnz);
        A->csr_val = dalloc<double>((size_t)nnz);
        int* fill = dalloc<int>((size_t)n_rows + 1);
        int* col_of_pos = dalloc<int>((size_t)nnz);
        int* pos_of_slot = dalloc<int>((size_t)nnz);
        int* bad = dalloc<int>(1);
        ATK_CUDA_CHECK(cudaMemsetAsync(A->row_ptr, 0, sizeof(int) * ((size_t)n_rows + 1), st));
        ATK_CUDA_CHECK(cudaMemsetAsync(fill, 0, sizeof(int) * ((size_t)n_rows + 1), st));
        ATK_CUDA_CHECK(cudaMemsetAsync(bad, 0, sizeof(int), st));
        check_col_ptr_kernel<<<grid_for((int64_t)n_cols + 1, 256, cap), 256, 0, st>>>(n_cols, nnz, d_col_ptr, bad);
        count_launch(ctx);
        if (nnz > 0) {
            count_rows_kernel<<<grid_for(nnz, 256, cap), 256, 0, st>>>(nnz, d_row_idx, A->row_ptr, n_rows, bad);
            count_launch(ctx);
        }
        int h_bad = 0;
        ATK_CUDA_CHECK(cudaMemcpyAsync(&h_bad, bad, sizeof(int), cudaMemcpyDeviceToHost, st));
        ATK_CUDA_CHECK(cudaStreamSynchronize(st));
        if (h_bad) {
            cudaFree(fill); cudaFree(col_of_pos); cudaFree(pos_of_slot); cudaFree(bad);
            if (h_bad & 2)
                throw Error(ATK_ERR_INVALID_ARGUMENT, "col_ptr is not a CSC column pointer array (first 0, non-decreasing, last nnz)");
            throw Error(ATK_ERR_OUT_OF_RANGE, "Index out of range");
        }
        exclusive_scan_i32(ctx, A->row_ptr, A->row_ptr, (int64_t)n_rows + 1);
        if (nnz > 0) {
            expand_cols_kernel<<<(unsigned)(((int64_t)n_cols * 8 + 255) / 256), 256, 0, st>>>(n_cols, d_col_ptr, col_of_pos);
            scatter_pos_kernel<<<grid_for(nnz, 256, cap), 256, 0, st>>>(nnz, d_row_idx, A->row_ptr, fill, pos_of_slot);
            order_rows_kernel<<<grid_for((int64_t)n_rows * 32, 256, cap), 256, 0, st>>>(n_rows, A->row_ptr, pos_of_slot,
                                                                                      col_of_pos, d_vals, A->csr_col, A->csr_val);
            count_launch(ctx, 3);
        }
        ATK_CUDA_CHECK(cudaGetLastError());
        ATK_CUDA_CHECK(cudaStreamSynchronize(st));
        cudaFree(fill); cudaFree(col_of_pos); cudaFree(pos_of_slot); cudaFree(bad);

        // ---- distributed context: keep rows [row_lo, row_hi); local columns = [own block | external entries]
        if (dist) {
            const int n_loc = row_hi - row_lo;
            int64_t nnz_loc = nnz;
            if (!rows_are_slice) {  // every rank converted the whole matrix: cut this rank's rows out of it
                int e0 = 0, e1 = 0;
                ATK_CUDA_CHECK(cudaMemcpy(&e0, A->row_ptr + row_lo, sizeof(int), cudaMemcpyDeviceToHost));
                ATK_CUDA_CHECK(cudaMemcpy(&e1, A->row_ptr + row_hi, sizeof(int), cudaMemcpyDeviceToHost));
                nnz_loc = (int64_t)e1 - e0;
                int* rp = dalloc<int>((size_t)n_loc + 1);
                int* cl = dalloc<int>((size_t)nnz_loc);
                double* vl = dalloc<double>((size_t)nnz_loc);
                try {
                    slice_row_ptr_kernel<<<(n_loc + 1 + 255) / 256, 256, 0, st>>>(n_loc, row_lo, A->row_ptr, rp);
                    if (nnz_loc > 0) {
                        ATK_CUDA_CHECK(cudaMemcpyAsync(cl, A->csr_col + e0, sizeof(int) * (size_t)nnz_loc, cudaMemcpyDeviceToDevice, st));
                        ATK_CUDA_CHECK(cudaMemcpyAsync(vl, A->csr_val + e0, sizeof(double) * (size_t)nnz_loc, cudaMemcpyDeviceToDevice, st));
                    }
                    ATK_CUDA_CHECK(cudaStreamSynchronize(st));
                    count_launch(ctx);
                } catch (...) {
                    cudaFree(rp); cudaFree(cl); cudaFree(vl);
                    throw;
                }
                cudaFree(A->row_ptr); cudaFree(A->csr_col); cudaFree(A->csr_val);
                A->row_ptr = rp;
                A->csr_col = cl;
                A->csr_val = vl;
            }
            A->dist = true;
            // the row blocks of all ranks must tile [0, n) in rank order
            struct Blk { int lo, hi; } mine{row_lo, row_hi};
            std::vector<Blk> all((size_t)ctx->world);
            allgather_host_bytes(ctx, &mine, all.data(), sizeof(Blk));
            std::vector<std::pair<int, int>> blocks;
            for (int q = 0; q < ctx->world; ++q) {
                const bool ok = all[(size_t)q].lo == (q == 0 ? 0 : all[(size_t)q - 1].hi) &&
                                (q + 1 < ctx->world || all[(size_t)q].hi == n_rows_global);
                ATK_REQUIRE(ok, ATK_ERR_INVALID_ARGUMENT, "distributed assembled operator: the ranks' row blocks do not tile the matrix");
                blocks.emplace_back(all[(size_t)q].lo, all[(size_t)q].hi);
            }
            // external columns: flag, scan, compact (ascending), renumber
            int* flag = dalloc<int>((size_t)n_cols + 1);
            int* pos = dalloc<int>((size_t)n_cols + 1);
            int* ext_cols = nullptr;
            int n_ext = 0;
            try {
                ATK_CUDA_CHECK(cudaMemsetAsync(flag, 0, sizeof(int) * ((size_t)n_cols + 1), st));
                if (nnz_loc > 0) mark_ext_kernel<<<grid_for(nnz_loc, 256, cap), 256, 0, st>>>(nnz_loc, A->csr_col, row_lo, row_hi, flag);
                exclusive_scan_i32(ctx, flag, pos, (int64_t)n_cols + 1);
                ATK_CUDA_CHECK(cudaMemcpyAsync(&n_ext, pos + n_cols, sizeof(int), cudaMemcpyDeviceToHost, st));
                ATK_CUDA_CHECK(cudaStreamSynchronize(st));
                ext_cols = dalloc<int>((size_t)n_ext);
                if (n_ext > 0) compact_ext_kernel<<<(n_cols + 255) / 256, 256, 0, st>>>(n_cols, flag, pos, ext_cols);
                if (nnz_loc > 0)
                    renumber_cols_kernel<<<grid_for(nnz_loc, 256, cap), 256, 0, st>>>(nnz_loc, row_lo, row_hi, n_loc, pos, A->csr_col);
                count_launch(ctx, 3);
                A->ext_cols_host.resize((size_t)n_ext);
                if (n_ext > 0)
                    ATK_CUDA_CHECK(cudaMemcpyAsync(A->ext_cols_host.data(), ext_cols, sizeof(int) * (size_t)n_ext, cudaMemcpyDeviceToHost, st));
                ATK_CUDA_CHECK(cudaStreamSynchronize(st));
                A->pg.setup(ctx, ext_cols, n_ext, A->ext_cols_host, blocks);
            } catch (...) {
                cudaFree(flag); cudaFree(pos); cudaFree(ext_cols);
                throw;
            }
            cudaFree(flag); cudaFree(pos); cudaFree(ext_cols);
            {   // rows [a, b) reference no external entry: their slices run while the exchange is in flight
                int* d_hr = dalloc<int>(2);
                const int init_hr[2] = {0, n_loc};
                int h_hr[2] = {0, n_loc};
                ATK_CUDA_CHECK(cudaMemcpyAsync(d_hr, init_hr, sizeof(init_hr), cudaMemcpyHostToDevice, st));
                ext_rows_kernel<<<(n_loc + 255) / 256, 256, 0, st>>>(n_loc, A->row_ptr, A->csr_col, n_loc, d_hr);
                ATK_CUDA_CHECK(cudaMemcpyAsync(h_hr, d_hr, sizeof(h_hr), cudaMemcpyDeviceToHost, st));
                ATK_CUDA_CHECK(cudaStreamSynchronize(st));
                cudaFree(d_hr);
                count_launch(ctx);
                const int sa = (h_hr[0] + kSliceRows - 1) / kSliceRows, sb = h_hr[1] / kSliceRows;
                if (sa < sb) {
                    A->s_int_begin = sa;
                    A->s_int_end = sb;
                }
            }
            n_rows = n_loc;
            nnz = nnz_loc;
            A->n_rows = n_loc;
            A->n_rows_local = n_loc;
            A->row_begin = row_lo;
            A->nnz = nnz_loc;
        }

        // lanes per row for the CSR-vector kernel: next power of two >= mean row length, in [2,32]
        const double mean = n_rows > 0 ? (double)nnz / n_rows : 0.0;
        int lanes = 2;
        while (lanes < 32 && lanes < mean) lanes <<= 1;
        A->lanes = lanes;
        // CSR format: rows short enough (longest row <= half the product buffer, mean <= 64) go through the stream kernel
        if (format != ATK_FORMAT_SELL && n_rows > 0 && nnz > 0 && !std::getenv("ATK_CSR_NO_STREAM")) {
            int* d_max = dalloc<int>(1);
            int h_max = 0;
            ATK_CUDA_CHECK(cudaMemsetAsync(d_max, 0, sizeof(int), st));
            max_row_len_kernel<<<grid_for(n_rows, 256, cap), 256, 0, st>>>(n_rows, A->row_ptr, d_max);
            ATK_CUDA_CHECK(cudaMemcpyAsync(&h_max, d_max, sizeof(int), cudaMemcpyDeviceToHost, st));
            ATK_CUDA_CHECK(cudaStreamSynchronize(st));
            cudaFree(d_max);
            count_launch(ctx);
            if (h_max <= kStreamCap / 2 && mean <= 64.0) {
                const int target = kStreamCap - h_max;
                const int nb = (int)((nnz + target - 1) / target);
                A->blk_row = dalloc<int>((size_t)nb + 1);
                csr_stream_blocks_kernel<<<(nb + 1 + 255) / 256, 256, 0, st>>>(n_rows, nb, target, A->row_ptr, A->blk_row);
                count_launch(ctx);
                ATK_CUDA_CHECK(cudaGetLastError());
                ATK_CUDA_CHECK(cudaStreamSynchronize(st));
                A->stream_blocks = nb;
            }
        }

        // ---- CSR -> SELL-32-sigma
        if (format == ATK_FORMAT_SELL) {
            const int n_slices = (n_rows + kSliceRows - 1) / kSliceRows;
            A->n_slices = n_slices;
            A->wscan = dalloc<int>((size_t)n_slices + 1);
            ATK_CUDA_CHECK(cudaMemsetAsync(A->wscan, 0, sizeof(int) * ((size_t)n_slices + 1), st));
            auto total_width = [&](const int* perm) -> int64_t {
                if (n_slices == 0) return 0;
                slice_width_kernel<<<(n_slices + 255) / 256, 256, 0, st>>>(n_rows, n_slices, A->row_ptr, perm, A->wscan);
                count_launch(ctx);
                exclusive_scan_i32(ctx, A->wscan, A->wscan, (int64_t)n_slices + 1);
                int tot = 0;
                ATK_CUDA_CHECK(cudaMemcpy(&tot, A->wscan + n_slices, sizeof(int), cudaMemcpyDeviceToHost));
                return tot;
            };
            int64_t tw = total_width(nullptr);
            // padding above 25 %: sort rows by length inside windows of 1024 rows
            if (nnz > 0 && (double)tw * kSliceRows > 1.25 * (double)nnz && n_rows > kSliceRows) {
                A->perm = dalloc<int>((size_t)n_rows);
                A->s_int_begin = A->s_int_end = 0;  // permuted rows: slices no longer map to contiguous row ranges
                sigma_rank_kernel<<<(n_rows + kSigma - 1) / kSigma, kSigma, 0, st>>>(n_rows, A->row_ptr, A->perm);
                count_launch(ctx);
                ATK_CUDA_CHECK(cudaMemsetAsync(A->wscan, 0, sizeof(int) * ((size_t)n_slices + 1), st));
                tw = total_width(A->perm);
            }
            A->sell_elems = tw * kSliceRows;
            ATK_REQUIRE(A->sell_elems < (int64_t)0x7fffffff * 8, ATK_ERR_NOMEM, "SELL storage too large");
            A->sell_col = dalloc<int>((size_t)A->sell_elems);
            A->sell_val = dalloc<double>((size_t)A->sell_elems);
            if (n_slices > 0) {
                const int64_t padded = (int64_t)n_slices * kSliceRows;
                sell_fill_kernel<<<(unsigned)((padded + 255) / 256), 256, 0, st>>>(n_rows, A->row_ptr, A->csr_col, A->csr_val,
                                                                                  A->perm, A->wscan, A->sell_col, A->sell_val);
                count_launch(ctx);
            }
            ATK_CUDA_CHECK(cudaGetLastError());
            ATK_CUDA_CHECK(cudaStreamSynchronize(st));
        }
    } catch (...) {
        delete A;
        throw;
    }
    return A;
}

// A rank's ROW SLICE of a CSC matrix: only the entries of rows [row_begin, row_end) (global row numbers), col_ptr over all
// columns.  No rank ever holds the whole matrix.
Operator* make_operator_from_csc_rows(Context* ctx, int n_rows_global, int n_cols, int row_begin, int row_end, int64_t nnz,
                                      const int* col_ptr, const int* row_idx, const double* vals, bool on_device, atk_format format) {
    ATK_REQUIRE(n_rows_global >= 0 && n_cols >= 0 && nnz >= 0 && nnz < (int64_t)0x7fffffff, ATK_ERR_INVALID_ARGUMENT,
                "bad matrix dimensions");
    ATK_REQUIRE(col_ptr != nullptr, ATK_ERR_INVALID_ARGUMENT, "col_ptr is null");
    ATK_REQUIRE(row_begin >= 0 && row_end > row_begin && row_end <= n_rows_global, ATK_ERR_INVALID_ARGUMENT, "bad row block");
    const int n_loc = row_end - row_begin;
    int *d_cp = nullptr, *d_ri = nullptr, *d_rl = nullptr, *d_bad = nullptr;
    double* d_v = nullptr;
    Operator* op = nullptr;
    auto cleanup = [&] {
        if (!on_device) { cudaFree(d_cp); cudaFree(d_ri); cudaFree(d_v); }
        cudaFree(d_rl); cudaFree(d_bad);
    };
    try {
        cudaStream_t st = ctx->stream;
        if (on_device) {
            d_cp = const_cast<int*>(col_ptr); d_ri = const_cast<int*>(row_idx); d_v = const_cast<double*>(vals);
        } else {
            d_cp = dalloc<int>((size_t)n_cols + 1);
            d_ri = dalloc<int>((size_t)nnz);
            d_v = dalloc<double>((size_t)nnz);
            ATK_CUDA_CHECK(cudaMemcpyAsync(d_cp, col_ptr, sizeof(int) * ((size_t)n_cols + 1), cudaMemcpyHostToDevice, st));
            if (nnz > 0) {
                ATK_CUDA_CHECK(cudaMemcpyAsync(d_ri, row_idx, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice, st));
                ATK_CUDA_CHECK(cudaMemcpyAsync(d_v, vals, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, st));
            }
        }
        d_rl = dalloc<int>((size_t)nnz);
        d_bad = dalloc<int>(1);
        ATK_CUDA_CHECK(cudaMemsetAsync(d_bad, 0, sizeof(int), st));
        if (nnz > 0) {
            shift_rows_kernel<<<grid_for(nnz, 256, ctx->sm_count * 16), 256, 0, st>>>(nnz, row_begin, d_ri, d_rl, n_loc, d_bad);
            count_launch(ctx);
        }
        int h_bad = 0;
        ATK_CUDA_CHECK(cudaMemcpyAsync(&h_bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
        ATK_CUDA_CHECK(cudaStreamSynchronize(st));
        ATK_REQUIRE(!h_bad, ATK_ERR_OUT_OF_RANGE, "Index out of range");  // a row outside the rank's block
        if (ctx->world == 1) op = build_from_device_csc(ctx, n_rows_global, n_cols, nnz, d_cp, d_rl, d_v, format);
        else op = build_from_device_csc(ctx, n_rows_global, n_cols, nnz, d_cp, d_rl, d_v, format, row_begin, row_end, true);
    } catch (...) {
        cleanup();
        throw;
    }
    cleanup();
    return op;
}

Operator* make_operator_from_csc(Context* ctx, int n_rows, int n_cols, int64_t nnz, const int* col_ptr, const int* row_idx,
                                 const double* vals, bool on_device, atk_format format) {
    ATK_REQUIRE(n_rows >= 0 && n_cols >= 0 && nnz >= 0, ATK_ERR_INVALID_ARGUMENT, "negative matrix dimension");
    ATK_REQUIRE(col_ptr != nullptr, ATK_ERR_INVALID_ARGUMENT, "col_ptr is null");
    if (on_device) return build_from_device_csc(ctx, n_rows, n_cols, nnz, col_ptr, row_idx, vals, format);
    int* d_cp = dalloc<int>((size_t)n_cols + 1);
    int* d_ri = dalloc<int>((size_t)nnz);
    double* d_v = dalloc<double>((size_t)nnz);
    Operator* op = nullptr;
    try {
        ATK_CUDA_CHECK(cudaMemcpyAsync(d_cp, col_ptr, sizeof(int) * ((size_t)n_cols + 1), cudaMemcpyHostToDevice, ctx->stream));
        if (nnz > 0) {
            ATK_CUDA_CHECK(cudaMemcpyAsync(d_ri, row_idx, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice, ctx->stream));
            ATK_CUDA_CHECK(cudaMemcpyAsync(d_v, vals, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, ctx->stream));
        }
        op = build_from_device_csc(ctx, n_rows, n_cols, nnz, d_cp, d_ri, d_v, format);
    } catch (...) {
        cudaFree(d_cp); cudaFree(d_ri); cudaFree(d_v);
        throw;
    }
    cudaFree(d_cp); cudaFree(d_ri); cudaFree(d_v);
    return op;
}

// ------------------------------------------------------------------------------------------ device-side assembly
// The CSC arrays Poisson3DSolver::buildLaplacianMatrix + finalize() would leave (Poisson.hpp:49-95), generated
// column by column: column c=(i,j,k) holds its diagonal plus one entry from every INTERIOR row adjacent to it.
namespace {
__device__ __forceinline__ bool on_shell(int i, int j, int k, int nx, int ny, int nz) {
    return i == 0 || i == nx - 1 || j == 0 || j == ny - 1 || k == 0 || k == nz - 1;
}
template <bool FILL>
__global__ void poisson_csc_kernel(int nx, int ny, int nz, double cx, double cy, double cz, double cc, int* __restrict__ col_ptr,
                                   int* __restrict__ row_idx, double* __restrict__ vals, int64_t r_lo, int64_t r_hi) {
    const int64_t n = (int64_t)nx * ny * nz;
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const int i = (int)(c % nx), j = (int)((c / nx) % ny), k = (int)(c / ((int64_t)nx * ny));
    const int64_t sxy = (int64_t)nx * ny;
    int p = FILL ? col_ptr[c] : 0;
    auto emit = [&](int64_t row, double v) {
        if (row < r_lo || row >= r_hi) return;  // a distributed context assembles only the rows it owns
        if (FILL) { row_idx[p] = (int)row; vals[p] = v; }
        ++p;
    };
    if (k - 1 >= 0 && !on_shell(i, j, k - 1, nx, ny, nz)) emit(c - sxy, cz);
    if (j - 1 >= 0 && !on_shell(i, j - 1, k, nx, ny, nz)) emit(c - nx, cy);
    if (i - 1 >= 0 && !on_shell(i - 1, j, k, nx, ny, nz)) emit(c - 1, cx);
    emit(c, on_shell(i, j, k, nx, ny, nz) ? 1.0 : cc);
    if (i + 1 < nx && !on_shell(i + 1, j, k, nx, ny, nz)) emit(c + 1, cx);
    if (j + 1 < ny && !on_shell(i, j + 1, k, nx, ny, nz)) emit(c + nx, cy);
    if (k + 1 < nz && !on_shell(i, j, k + 1, nx, ny, nz)) emit(c + sxy, cz);
    if (!FILL) col_ptr[c] = p;
}

struct Coef27 { double v[27]; };
template <bool FILL>
__global__ void op27_csc_kernel(int nx, int ny, int nz, Coef27 coef, int* __restrict__ col_ptr, int* __restrict__ row_idx,
                                double* __restrict__ vals, int64_t r_lo, int64_t r_hi) {
    const int64_t n = (int64_t)nx * ny * nz;
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const int ic = (int)(c % nx), jc = (int)((c / nx) % ny), kc = (int)(c / ((int64_t)nx * ny));
    const int64_t sxy = (int64_t)nx * ny;
    int p = FILL ? col_ptr[c] : 0;
    for (int dk = 1; dk >= -1; --dk)
        for (int dj = 1; dj >= -1; --dj)
            for (int di = 1; di >= -1; --di) {
                const int ir = ic - di, jr = jc - dj, kr = kc - dk;
                if (ir < 0 || ir >= nx || jr < 0 || jr >= ny || kr < 0 || kr >= nz) continue;
                const double v = coef.v[(dk + 1) * 9 + (dj + 1) * 3 + (di + 1)];
                if (v == 0.0) continue;  // addValue never stores an explicit zero (SparseObj.hpp:71-73)
                const int64_t row = ir + (int64_t)jr * nx + (int64_t)kr * sxy;
                if (row < r_lo || row >= r_hi) continue;  // a distributed context assembles only the rows it owns
                if (FILL) {
                    row_idx[p] = (int)row;
                    vals[p] = v;
                }
                ++p;
            }
    if (!FILL) col_ptr[c] = p;
}
}  // namespace

// grid operators on a distributed context own whole z-planes, like the matrix-free stencils (same vector partition)
static int slab_rows_lo(Context* ctx, int nx, int ny, int nz) {
    int k0 = 0, k1 = 0;
    ATK_REQUIRE(nz >= ctx->world, ATK_ERR_INVALID_ARGUMENT, "fewer z-planes than ranks");
    slab_partition(nz, ctx->rank, ctx->world, &k0, &k1);
    return k0 * nx * ny;
}
static int slab_rows_hi(Context* ctx, int nx, int ny, int nz) {
    int k0 = 0, k1 = 0;
    slab_partition(nz, ctx->rank, ctx->world, &k0, &k1);
    return ctx->world > 1 ? k1 * nx * ny : -1;
}
template <typename CountK, typename FillK>
static Operator* assemble_on_device(Context* ctx, int64_t n, atk_format format, CountK count, FillK fill, int row_lo = 0,
                                    int row_hi = -1) {
    ATK_REQUIRE(n > 0 && n < (int64_t)0x7fffffff, ATK_ERR_INVALID_ARGUMENT, "grid too large for int32 indices");
    int* d_cp = dalloc<int>((size_t)n + 1);
    int* d_ri = nullptr;
    double* d_v = nullptr;
    Operator* op = nullptr;
    try {
        ATK_CUDA_CHECK(cudaMemsetAsync(d_cp, 0, sizeof(int) * ((size_t)n + 1), ctx->stream));
        count(d_cp);
        count_launch(ctx);
        exclusive_scan_i32(ctx, d_cp, d_cp, n + 1);
        int nnz = 0;
        ATK_CUDA_CHECK(cudaMemcpy(&nnz, d_cp + n, sizeof(int), cudaMemcpyDeviceToHost));
        ATK_REQUIRE(nnz >= 0, ATK_ERR_INVALID_ARGUMENT, "nnz overflows int32");
        d_ri = dalloc<int>((size_t)nnz);
        d_v = dalloc<double>((size_t)nnz);
        fill(d_cp, d_ri, d_v);
        count_launch(ctx);
        ATK_CUDA_CHECK(cudaGetLastError());
        op = build_from_device_csc(ctx, (int)n, (int)n, nnz, d_cp, d_ri, d_v, format, row_lo, row_hi);
    } catch (...) {
        cudaFree(d_cp); cudaFree(d_ri); cudaFree(d_v);
        throw;
    }
    cudaFree(d_cp); cudaFree(d_ri); cudaFree(d_v);
    return op;
}

Operator* make_poisson_assembled(Context* ctx, int nx, int ny, int nz, double cx, double cy, double cz, atk_format format) {
    ATK_REQUIRE(nx > 0 && ny > 0 && nz > 0, ATK_ERR_INVALID_ARGUMENT, "grid dimensions must be positive");
    const int64_t n = (int64_t)nx * ny * nz;
    const double cc = -2.0 * (cx + cy + cz);  // Poisson.hpp:57
    const unsigned grid = (unsigned)((n + 255) / 256);
    cudaStream_t st = ctx->stream;
    const int row_lo = slab_rows_lo(ctx, nx, ny, nz), row_hi = slab_rows_hi(ctx, nx, ny, nz);
    const int64_t r_lo = row_hi < 0 ? 0 : row_lo, r_hi = row_hi < 0 ? n : row_hi;
    return assemble_on_device(
        ctx, n, format,
        [&](int* cp) { poisson_csc_kernel<false><<<grid, 256, 0, st>>>(nx, ny, nz, cx, cy, cz, cc, cp, nullptr, nullptr, r_lo, r_hi); },
        [&](int* cp, int* ri, double* v) {
            poisson_csc_kernel<true><<<grid, 256, 0, st>>>(nx, ny, nz, cx, cy, cz, cc, cp, ri, v, r_lo, r_hi);
        },
        row_lo, row_hi);
}

Operator* make_stencil27_assembled(Context* ctx, int nx, int ny, int nz, const double* coef, atk_format format) {
    ATK_REQUIRE(nx > 0 && ny > 0 && nz > 0 && coef, ATK_ERR_INVALID_ARGUMENT, "bad 27-point operator arguments");
    const int64_t n = (int64_t)nx * ny * nz;
    Coef27 cf;
    for (int t = 0; t < 27; ++t) cf.v[t] = coef[t];
    const unsigned grid = (unsigned)((n + 255) / 256);
    cudaStream_t st = ctx->stream;
    const int row_lo = slab_rows_lo(ctx, nx, ny, nz), row_hi = slab_rows_hi(ctx, nx, ny, nz);
    const int64_t r_lo = row_hi < 0 ? 0 : row_lo, r_hi = row_hi < 0 ? n : row_hi;
    return assemble_on_device(
        ctx, n, format,
        [&](int* cp) { op27_csc_kernel<false><<<grid, 256, 0, st>>>(nx, ny, nz, cf, cp, nullptr, nullptr, r_lo, r_hi); },
        [&](int* cp, int* ri, double* v) { op27_csc_kernel<true><<<grid, 256, 0, st>>>(nx, ny, nz, cf, cp, ri, v, r_lo, r_hi); },
        row_lo, row_hi);
}

}  // namespace atk
