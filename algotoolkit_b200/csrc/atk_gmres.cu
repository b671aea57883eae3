// atk_gmres.cu -- GMRES::solve on the device (reference src/LinearAlgebra/Krylov/GMRES.hpp:27-171, the
// unpreconditioned branch), keeping the reference's behaviour:
//   * the Arnoldi projection runs over i < j only (:73) - w is never orthogonalised against V[j];
//   * H is allocated once per solve() and never cleared between restarts (:55), so from the second cycle on the
//     Givens step starts from the previous cycle's rho[j] left in H(j,j) (:89-94);
//   * no inner convergence test; `tol` is an absolute residual threshold tested once per restart (:107) and against
//     ||w|| inside the cycle (:81-85);
//   * x = x + V[i]*y[i] is added for i ascending with two roundings per term (:167-170).
//
// Device data: V is one (K+1) x n column-major array, w/r/t are work vectors; the (K+1)^2 Hessenberg, the Givens
// coefficients and e1 live on the HOST (O(K^2) scalar work per cycle, negligible bytes) and are computed with the
// reference's formulae in the reference's order.  The host needs ||w|| and the column of H after every Arnoldi step,
// so there is one small D2H per step; everything else stays on the device.
//
// Orthogonalisation modes:
//   MGS  (parity path)  per i: dot kernel -> h in a reduction slot; axmy kernel reads h from the slot: w = w - V[i]*h.
//        Algorithmic traffic per (i) pair 40 B/row.
//   MGS_BATCHED (throughput)  the same projection evaluated with TWO passes over V[0:j] and two reductions per step:
//        c = V[0:j]^T w and g = V[0:j]^T v_j in one pass (g extends the Gram matrix G = V^T V), the host solves the unit
//        lower-triangular system (I + strict_lower(G)) h = c, then w -= V[0:j] h in the second pass.
//        In exact arithmetic h_i = v_i.(w - sum_{m<i} h_m v_m) = c_i - sum_{m<i} G_im h_m, i.e. exactly the reference's
//        sequential loop - also for the reference's NON-orthonormal basis (it never projects against v_j, :73), which is
//        why plain classical Gram-Schmidt (h = c) is NOT equivalent and fails parity by orders of magnitude.
//        Traffic 16*n*j B per step instead of 40*n*j; rounding differs from the sequential loop at the 1e-14 level on
//        the BASELINE config-5 operator (tests/test_gpu_parity.py).
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "atk_internal.cuh"

namespace atk {

namespace {

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// w = w - v*h, h read from a reduction slot (sum over ranks in rank order); thread 0 also records h.
__global__ void __launch_bounds__(256) mgs_axmy_kernel(int64_t n, const double* __restrict__ v, double* __restrict__ w,
                                                       const double* __restrict__ slot, int world, double* __restrict__ h_out) {
    const double h = slot_sum(slot, world);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0) *h_out = h;
    for (int64_t i = t; i < n; i += stride) w[i] = sub_rn(w[i], mul_rn(v[i], h));
}

// Per-projection path, tree sums: the update of projection i-1 and the products of projection i in ONE pass,
//     w = w - v_prev * h_(i-1)   (h from the slot the previous launch left)  ;  partial sums of v_cur . w  (v_cur == nullptr: w . w)
// 32 B/row and one launch per projection instead of 40 B/row and two (dot, then axmy).  Thread layout, accumulation order
// and the grid-wide sum are those of dot_tree_kernel (atk_vector.cu), so with 16-byte-aligned columns (even n) h is
// bit-identical to the unfused path.
template <bool VEC2>
__global__ void __launch_bounds__(kReduceBlock) mgs_axmy_dot_kernel(int64_t n, const double* __restrict__ v_prev, const double* __restrict__ v_cur,
                                                                    double* __restrict__ w, const double* __restrict__ slot_prev, int world,
                                                                    double* __restrict__ h_out, double* partials, unsigned* ticket, double* out) {
    const double h = slot_sum(slot_prev, world);  // every rank's share of h_(i-1), added in rank order
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0) *h_out = h;  // H(i-1, j)
    double acc = 0.0;
    if constexpr (VEC2) {
        const int64_t n2 = n >> 1;
        const double2* p2 = reinterpret_cast<const double2*>(v_prev);
        const double2* c2 = reinterpret_cast<const double2*>(v_cur);
        double2* w2 = reinterpret_cast<double2*>(w);
        double acc1 = 0.0;
        for (int64_t i = t; i < n2; i += stride) {
            const double2 pv = p2[i];
            double2 wv = w2[i];
            wv.x = sub_rn(wv.x, mul_rn(pv.x, h));
            wv.y = sub_rn(wv.y, mul_rn(pv.y, h));
            w2[i] = wv;
            const double2 cv = v_cur ? c2[i] : wv;
            acc = add_rn(acc, mul_rn(cv.x, wv.x));
            acc1 = add_rn(acc1, mul_rn(cv.y, wv.y));
        }
        acc = add_rn(acc, acc1);
        if (t == 0 && (n & 1)) {
            const double wv = sub_rn(w[n - 1], mul_rn(v_prev[n - 1], h));
            w[n - 1] = wv;
            acc = add_rn(acc, mul_rn(v_cur ? v_cur[n - 1] : wv, wv));
        }
    } else {
        for (int64_t i = t; i < n; i += stride) {
            const double wv = sub_rn(w[i], mul_rn(v_prev[i], h));
            w[i] = wv;
            acc = add_rn(acc, mul_rn(v_cur ? v_cur[i] : wv, wv));
        }
    }
    double total;
    if (grid_sum_last_block<kReduceBlock>(acc, partials, ticket, &total)) {
        if (linear_thread_id() == 0) *out = total;
    }
}

// out = w / sqrt(slot)  (V[j+1] = w / H(j+1,j), :86) - skipped when the norm is zero (the host raises the error).
__global__ void __launch_bounds__(256) normalize_kernel(int64_t n, const double* __restrict__ w, double* __restrict__ out,
                                                        const double* __restrict__ slot, int world, double* __restrict__ norm_out,
                                                        double skip_below) {
    const double nrm = sqrt(slot_sum(slot, world));
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0) *norm_out = nrm;
    if (nrm == 0.0 || nrm < skip_below) return;  // the reference breaks before assigning the next basis vector
    for (int64_t i = t; i < n; i += stride) out[i] = __ddiv_rn(w[i], nrm);
}

// x[e] = (((x[e] + V0[e]*y0) + V1[e]*y1) + ...), i ascending: bit-identical to k successive x = x + V[i]*y[i].
__global__ void __launch_bounds__(256) update_solution_kernel(int64_t n, int k, const double* __restrict__ V, int64_t ldv,
                                                              const double* __restrict__ y, double* __restrict__ x) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
        double acc = x[e];
        for (int i = 0; i < k; ++i) acc = add_rn(acc, mul_rn(V[(int64_t)i * ldv + e], __ldg(y + i)));
        x[e] = acc;
    }
}

// ---- batched (low-synchronisation) MGS kernels -------------------------------------------------------------------------
// Pass 1: partial sums of V[col].w (col < ncols_w) and V[col].vj (col < ncols_g) for every column, one pass over V.
// A CTA owns kBdBlock*kBdRows consecutive rows, keeps its rows of w and vj in registers and walks the columns in tiles of
// kBdTile, so kBdTile*kBdRows column loads are in flight per thread and the block reduction is paid once per tile.
// part[(rhs*ncols_g + col) * gridDim.x + block]; a second kernel adds the per-CTA partials in CTA order.
constexpr int kBdBlock = 256;
constexpr int kBdRows = 4;
constexpr int kBdTile = 8;
__global__ void __launch_bounds__(kBdBlock) multi_dot2_kernel(int64_t n, int ncols_w, int ncols_g, const double* __restrict__ V,
                                                              int64_t ldv, const double* __restrict__ w,
                                                              const double* __restrict__ vj, double* __restrict__ part) {
    __shared__ double red[2 * kBdTile][kBdBlock / 32];
    const int64_t base = (int64_t)blockIdx.x * kBdBlock * kBdRows;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double wq[kBdRows], gq[kBdRows];
    bool in[kBdRows];
#pragma unroll
    for (int q = 0; q < kBdRows; ++q) {
        const int64_t e = base + (int64_t)q * kBdBlock + threadIdx.x;
        in[q] = e < n;
        wq[q] = in[q] ? w[e] : 0.0;
        gq[q] = in[q] ? vj[e] : 0.0;
    }
    for (int c0 = 0; c0 < ncols_g; c0 += kBdTile) {
        double vals[kBdTile][kBdRows];
#pragma unroll
        for (int t = 0; t < kBdTile; ++t) {
            const double* v = V + (int64_t)min(c0 + t, ncols_g - 1) * ldv + base + threadIdx.x;
#pragma unroll
            for (int q = 0; q < kBdRows; ++q) vals[t][q] = in[q] ? v[(int64_t)q * kBdBlock] : 0.0;
        }
        double acc[2 * kBdTile];
#pragma unroll
        for (int t = 0; t < kBdTile; ++t) {
            double aw = 0.0, ag = 0.0;
#pragma unroll
            for (int q = 0; q < kBdRows; ++q) {
                aw = add_rn(aw, mul_rn(vals[t][q], wq[q]));
                ag = add_rn(ag, mul_rn(vals[t][q], gq[q]));
            }
            acc[t] = aw;
            acc[kBdTile + t] = ag;
        }
#pragma unroll
        for (int a = 0; a < 2 * kBdTile; ++a) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[a] = add_rn(acc[a], __shfl_xor_sync(0xffffffffu, acc[a], o));
        }
        __syncthreads();  // red[] free again
        if (lane == 0) {
#pragma unroll
            for (int a = 0; a < 2 * kBdTile; ++a) red[a][warp] = acc[a];
        }
        __syncthreads();
        if (threadIdx.x < 2 * kBdTile) {
            const int a = threadIdx.x, t = a % kBdTile, rhs = a / kBdTile, col = c0 + t;
            if (col < ncols_g && (rhs == 1 || col < ncols_w)) {
                double sum = red[a][0];
#pragma unroll
                for (int ww = 1; ww < kBdBlock / 32; ++ww) sum = add_rn(sum, red[a][ww]);
                part[((int64_t)rhs * ncols_g + col) * gridDim.x + blockIdx.x] = sum;
            }
        }
    }
}
// out[rhs*ncols_g + col] = sum over CTAs (fixed order) of part[...]; one CTA per (rhs, col)
__global__ void __launch_bounds__(256) multi_dot2_final_kernel(int ncols_w, int ncols_g, int nblocks, const double* __restrict__ part,
                                                               double* __restrict__ out) {
    const int idx = blockIdx.x, rhs = idx / ncols_g, col = idx % ncols_g;
    if (rhs == 0 && col >= ncols_w) {
        if (threadIdx.x == 0) out[idx] = 0.0;
        return;
    }
    const double total = partials_sum<256>(part + (int64_t)idx * nblocks, (unsigned)nblocks);
    if (threadIdx.x == 0) out[idx] = total;
}
// Pass 2: w = w - V[0]*h0 - V[1]*h1 - ... (i ascending, product rounded first: the reference's order per element, :76)
// fused with the partial sums of w.w for the norm that follows (:79).
__global__ void __launch_bounds__(kReduceBlock) multi_axmy_norm_kernel(int64_t n, int j, const double* __restrict__ V, int64_t ldv,
                                                                       const double* __restrict__ h, double* __restrict__ w,
                                                                       double* partials, unsigned* ticket, double* norm2_out) {
    extern __shared__ double hs[];
    for (int i = threadIdx.x; i < j; i += blockDim.x) hs[i] = h[i];
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    double nacc = 0.0;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
        double acc = w[e];
        int i = 0;
        for (; i + 8 <= j; i += 8) {
            double v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = V[(int64_t)(i + u) * ldv + e];
#pragma unroll
            for (int u = 0; u < 8; ++u) acc = sub_rn(acc, mul_rn(v[u], hs[i + u]));
        }
        for (; i < j; ++i) acc = sub_rn(acc, mul_rn(V[(int64_t)i * ldv + e], hs[i]));
        w[e] = acc;
        nacc = add_rn(nacc, mul_rn(acc, acc));
    }
    double total;
    if (grid_sum_last_block<kReduceBlock>(nacc, partials, ticket, &total)) {
        if (threadIdx.x == 0) *norm2_out = total;
    }
}

// ---- persistent Arnoldi-MGS step ----------------------------------------------------------------------------------
// One cooperative launch performs the WHOLE projection loop of an Arnoldi step (GMRES.hpp:73-86):
//     for i < j: h = V[i].w ; w = w - V[i]*h        then  ||w|| ; V[j+1] = w / ||w||
// Each CTA owns a contiguous chunk of rows and keeps its slice of w ON CHIP for the whole step, so per projection only
// V[i] streams in (8 B/row from HBM) instead of the 40 B/row of separate dot and axmy kernels, and the 2j kernel launches
// of the step collapse into one.  Three variants by slice length (rows per CTA, 512 threads):
//   <= 4 096          arnoldi_mgs_kernel<.., true>    w and the current column in registers, next column requested early
//   <= 14 336         arnoldi_mgs_ring_kernel         w in registers, columns through a two-deep cp.async ring in smem
//   <= ~28 k          arnoldi_mgs_kernel<.., false>   w in shared memory, every column read twice (second read from L2,
//                                                     kept there by cache-hint policies)
// CTAs meet at one software grid barrier per projection (cooperative launch guarantees co-residency); every CTA then adds
// the per-CTA partials in CTA order.
// With a single CTA (small n, e.g. bcsstk08) the barrier degenerates to __syncthreads; that configuration also
// implements ATK_REDUCE_SEQUENTIAL (thread 0 adds the rounded products strictly in index order).
struct MgsArgs {
    int64_t n;
    int j;
    const double* V;
    int64_t ldv;
    const double* w;     // A*V[j]
    double* vnext;       // V[j+1]
    double* h_out;       // [j+1]: H(0..j-1, j) and ||w||
    double* partials;    // [2 * gridDim.x]
    unsigned* bar;       // zero-initialised counter of this step
    int64_t chunk;       // rows per CTA (multiple of 2)
    int sequential;
    double skip_below;   // V[j+1] is left untouched when ||w|| < skip_below (the reference breaks before assigning it)
    int* stop;           // optional {stopped, step}: set by the step whose norm ends the cycle; later steps return at once
    // distributed contexts in peer-memory mode: every dependent reduction is per-CTA partial -> arrival counter -> CTA 0
    // adds the partials in CTA order and stores the rank's sum into EVERY rank's packed {value half | sequence} words
    // (NVLink stores) -> every CTA of every rank polls its local words and adds the rank sums in rank order.  ~4 us per
    // projection instead of a kernel boundary, an NCCL all-gather and a second kernel.
    int world, rank;
    unsigned long long* ll;               // local packed words
    unsigned long long* const* peer_ll;   // every rank's packed words as mapped here
    unsigned long long seq_base;          // this launch uses sequence numbers seq_base + 1 .. seq_base + j + 1
    unsigned long long* stamps;           // profile build (ATK_REDUCE_PROFILE): [gridDim.x][4] + [gridDim.x][3] summed clocks, else nullptr
};
constexpr int kMgsLlSlot = 2;  // packed-word slots 2 and 3 (0 and 1 belong to the CG loops)

// Strictly ordered sum (single CTA, ATK_REDUCE_SEQUENTIAL): all threads write the rounded products of the whole slice
// into `stage`, one barrier, thread 0 adds them in index order from 0.0 (= std::inner_product), result broadcast.
template <typename F>
__device__ __forceinline__ double sequential_sum(int64_t len, double* stage, double* bcast, F product) {
    for (int64_t e = threadIdx.x; e < len; e += blockDim.x) stage[e] = product(e);
    __syncthreads();
    if (threadIdx.x == 0) {
        double acc = 0.0;
#pragma unroll 16
        for (int64_t k = 0; k < len; ++k) acc = add_rn(acc, stage[k]);
        *bcast = acc;
    }
    __syncthreads();
    return *bcast;
}

// h from a per-thread partial: fixed block tree, then (several CTAs) barrier + fixed-order sum of the CTA partials; on
// distributed contexts the rank sums cross NVLink as packed words (see MgsArgs).  Ends with a block-wide barrier.
// (Measured and rejected, profiles/r02_mgs_step.txt: every CTA leaving its sum as a packed {half | tag} word pair and every
// CTA polling all 148 pairs - no fence, no atomic, one L2 round trip less - is SLOWER at 2.1 M rows, with one thread per
// pair (10.4 vs 6.7 us per projection) and with one warp and 16-byte loads (5.46 vs 4.77); so is CTA 0 gathering the
// pairs and handing the total back as a pair (6.73 vs 6.74).)
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
struct MgsNoHook {
    __device__ __forceinline__ void operator()() const {}
};
// `after_arrive` runs on every thread once this CTA's sum has been handed over (before the wait for the others): the ring
// kernel parks there the cp.async requests of the arriving warp, so that its fence does not wait for column data.
template <int BLOCK, typename Hook = MgsNoHook>
__device__ __forceinline__ double mgs_finish_tree(const MgsArgs& a, double acc, int i, unsigned& target, double* rank_vals, long long* prof = nullptr,
                                                  Hook after_arrive = Hook()) {
#ifdef ATK_REDUCE_PROFILE
    long long pl = clock64();
    auto station = [&](int k) { if (prof) { const long long now = clock64(); prof[k] += now - pl; pl = now; } };
#else
    auto station = [](int) {};
#endif
    const unsigned G = gridDim.x;
    const double bs = block_sum<BLOCK>(acc);
    station(0);  // block tree (includes waiting for the CTA's slowest warp)
    if (G == 1 && a.world == 1) {
        after_arrive();
        return bs;
    }
    double* part = a.partials + (size_t)(i & 1) * G;
    if (a.world == 1) {
        target += G;
        if (threadIdx.x == 0) {
            __stcg(part + blockIdx.x, bs);
            __threadfence();
            atomicAdd(a.bar, 1u);
        }
        after_arrive();
        if (threadIdx.x == 0) {
            while (ld_acquire_gpu(a.bar) < target) {}
        }
        __syncthreads();
        station(1);  // store, fence, arrive, wait for the last CTA
        double s = 0.0;
        for (unsigned q = threadIdx.x; q < G; q += BLOCK) s = add_rn(s, __ldcg(part + q));
        const double tot = block_sum<BLOCK>(s);
        station(2);  // the partials and the second block tree
        return tot;
    }
    // ---- distributed: arrive, CTA 0 publishes the rank's sum to every rank, everyone polls its local words
    const int slot = kMgsLlSlot + (i & 1);
    const unsigned long long seq = a.seq_base + (unsigned long long)i + 1ull;
    target += G;
    if (threadIdx.x == 0) {
        __stcg(part + blockIdx.x, bs);
        __threadfence();
        atomicAdd(a.bar, 1u);
    }
    after_arrive();
    if (blockIdx.x == 0) {
        if (threadIdx.x == 0) {
            unsigned spins = 0;
            unsigned long long t0 = 0;
            while (*reinterpret_cast<volatile unsigned*>(a.bar) < target) {
                if ((++spins & 0xfffu) == 0) {
                    const unsigned long long now = global_timer_ns();
                    if (t0 == 0) t0 = now;
                    if (now - t0 > kPeerWaitNs) asm volatile("trap;");
                }
            }
            __threadfence();
        }
        __syncthreads();
        double s = 0.0;
        for (unsigned q = threadIdx.x; q < G; q += BLOCK) s = add_rn(s, __ldcg(part + q));
        const double total = block_sum<BLOCK>(s);
        if (threadIdx.x < 32 && (int)threadIdx.x < a.world) {
            const unsigned long long bits = (unsigned long long)__double_as_longlong(total);
            const unsigned long long tag = ll_tag(seq) << 32;
            unsigned long long* dst = ((int)threadIdx.x == a.rank ? a.ll : a.peer_ll[threadIdx.x]) + (size_t)(slot * kMaxRanks + a.rank) * 2;
            st_relaxed_sys(dst, (bits & 0xffffffffull) | tag);
            st_relaxed_sys(dst + 1, (bits >> 32) | tag);
        }
    }
    if (threadIdx.x < 32 && (int)threadIdx.x < a.world) {
        const unsigned long long* src = a.ll + (size_t)(slot * kMaxRanks + threadIdx.x) * 2;
        const unsigned long long tag = ll_tag(seq);
        unsigned long long w0, w1;
        unsigned spins = 0;
        unsigned long long t0 = 0;
        for (;;) {
            w0 = ld_relaxed_sys(src);
            w1 = ld_relaxed_sys(src + 1);
            if ((w0 >> 32) == tag && (w1 >> 32) == tag) break;
            if ((++spins & 0xfffu) == 0) {
                const unsigned long long now = global_timer_ns();
                if (t0 == 0) t0 = now;
                if (now - t0 > kPeerWaitNs) asm volatile("trap;");  // a rank died or fell out of step: fail loudly
            }
        }
        rank_vals[threadIdx.x] = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
    }
    __syncthreads();
    double h = rank_vals[0];
    for (int q = 1; q < a.world; ++q) h = add_rn(h, rank_vals[q]);
    __syncthreads();  // rank_vals is rewritten by the next projection
    return h;
}

constexpr int kMgsRegRows = 8;  // rows per thread the register variant can hold

// REG = true (slices of at most 8*BLOCK rows): the thread's rows of w and of the current V column live in registers,
// the next column is requested while the current one is being reduced, and nothing is re-read.
// REG = false: the slice of w lives in shared memory and V[i] is read twice, by pass i for the products and by pass
// i+1 for the update (from L2: see the cache-hint policies in the pass).
template <int BLOCK, bool REG>
__global__ void __launch_bounds__(BLOCK) arnoldi_mgs_kernel(MgsArgs a) {
    extern __shared__ __align__(16) double wsm[];  // REG: products [chunk] (sequential mode only); else w [chunk] (+ products)
    __shared__ double bcast;
    // Written only by the last instruction of an earlier step (every CTA of that step has passed its grid barriers by
    // then), so all CTAs of this launch read the same value.
    if (a.stop && *reinterpret_cast<volatile int*>(a.stop)) return;
    const int64_t c0 = (int64_t)blockIdx.x * a.chunk;
    const int64_t len = max((int64_t)0, min(a.chunk, a.n - c0));
    unsigned target = 0;
    __shared__ double rank_vals[kMaxRanks];
    auto finish_tree = [&](double acc, int i) -> double { return mgs_finish_tree<BLOCK>(a, acc, i, target, rank_vals); };

    if constexpr (REG) {
        double wr[kMgsRegRows], vc[kMgsRegRows], vn[kMgsRegRows];
#pragma unroll
        for (int q = 0; q < kMgsRegRows; ++q) {
            const int64_t e = threadIdx.x + (int64_t)q * BLOCK;
            wr[q] = e < len ? a.w[c0 + e] : 0.0;
            vc[q] = (e < len && a.j > 0) ? a.V[c0 + e] : 0.0;
            vn[q] = 0.0;
        }
        for (int i = 0; i <= a.j; ++i) {
            if (i + 1 < a.j) {  // request the next column before this one is reduced
                const double* v = a.V + (int64_t)(i + 1) * a.ldv + c0;
#pragma unroll
                for (int q = 0; q < kMgsRegRows; ++q) {
                    const int64_t e = threadIdx.x + (int64_t)q * BLOCK;
                    vn[q] = e < len ? v[e] : 0.0;
                }
            }
            double h;
            if (a.sequential) {
#pragma unroll
                for (int q = 0; q < kMgsRegRows; ++q) {
                    const int64_t e = threadIdx.x + (int64_t)q * BLOCK;
                    if (e < len) wsm[e] = (i < a.j) ? mul_rn(vc[q], wr[q]) : mul_rn(wr[q], wr[q]);
                }
                __syncthreads();
                if (threadIdx.x == 0) {
                    double acc = 0.0;
#pragma unroll 16
                    for (int64_t k = 0; k < len; ++k) acc = add_rn(acc, wsm[k]);
                    bcast = acc;
                }
                __syncthreads();
                h = bcast;
            } else {
                double acc = 0.0;
#pragma unroll
                for (int q = 0; q < kMgsRegRows; ++q)  // rows beyond len hold zeros: adding +0.0 products changes nothing
                    acc = add_rn(acc, (i < a.j) ? mul_rn(vc[q], wr[q]) : mul_rn(wr[q], wr[q]));
                h = finish_tree(acc, i);
            }
            if (i < a.j) {
                if (blockIdx.x == 0 && threadIdx.x == 0) a.h_out[i] = h;  // H(i,j)   (:75)
#pragma unroll
                for (int q = 0; q < kMgsRegRows; ++q) {
                    wr[q] = sub_rn(wr[q], mul_rn(vc[q], h));               // (:76)
                    vc[q] = vn[q];
                }
            } else {
                const double nrm = sqrt(h);                                // (:79)
                if (blockIdx.x == 0 && threadIdx.x == 0) {
                    a.h_out[a.j] = nrm;
                    if (a.stop && (nrm == 0.0 || nrm < a.skip_below)) { a.stop[1] = a.j; a.stop[0] = 1; }
                }
                if (nrm != 0.0 && !(nrm < a.skip_below)) {
#pragma unroll
                    for (int q = 0; q < kMgsRegRows; ++q) {
                        const int64_t e = threadIdx.x + (int64_t)q * BLOCK;
                        if (e < len) a.vnext[c0 + e] = __ddiv_rn(wr[q], nrm);  // (:86)
                    }
                }
            }
        }
    } else {
        double* stage = wsm + a.chunk;
        for (int64_t e = threadIdx.x; e < len; e += BLOCK) wsm[e] = a.w[c0 + e];
        __syncthreads();
        if (a.sequential) {  // single CTA, reference-ordered sums: separate dot and update passes
            for (int i = 0; i <= a.j; ++i) {
                const double* v = (i < a.j) ? a.V + (int64_t)i * a.ldv + c0 : nullptr;
                double h;
                if (i < a.j) h = sequential_sum(len, stage, &bcast, [&](int64_t e) { return mul_rn(v[e], wsm[e]); });
                else h = sequential_sum(len, stage, &bcast, [&](int64_t e) { return mul_rn(wsm[e], wsm[e]); });
                if (i < a.j) {
                    if (blockIdx.x == 0 && threadIdx.x == 0) a.h_out[i] = h;                                          // H(i,j)   (:75)
                    for (int64_t e = threadIdx.x; e < len; e += BLOCK) wsm[e] = sub_rn(wsm[e], mul_rn(v[e], h));      // (:76)
                    __syncthreads();
                } else {
                    const double nrm = sqrt(h);                                                                       // (:79)
                    if (blockIdx.x == 0 && threadIdx.x == 0) {
                        a.h_out[a.j] = nrm;
                        if (a.stop && (nrm == 0.0 || nrm < a.skip_below)) { a.stop[1] = a.j; a.stop[0] = 1; }
                    }
                    if (nrm != 0.0 && !(nrm < a.skip_below))
                        for (int64_t e = threadIdx.x; e < len; e += BLOCK) a.vnext[c0 + e] = __ddiv_rn(wsm[e], nrm);  // (:86)
                }
            }
        } else {
            // Tree sums: ONE pass per projection.  Pass i applies the update of projection i-1 (w -= V[i-1]*h_(i-1), the
            // column is still in L2) and forms the products with V[i] (streamed from HBM) on the updated values - the same
            // per-element arithmetic in the same order as separate passes, with every thread touching only its own entries
            // of w.  16-byte accesses when the columns allow it; four independent loads in flight per thread and array.
            const bool vec = ((a.ldv & 1) == 0) && ((c0 & 1) == 0) && ((reinterpret_cast<uintptr_t>(a.V) & 15u) == 0);
            double hprev = 0.0;
            for (int i = 0; i <= a.j; ++i) {
                const double* vp = (i > 0) ? a.V + (int64_t)(i - 1) * a.ldv + c0 : nullptr;
                const double* vc = (i < a.j) ? a.V + (int64_t)i * a.ldv + c0 : nullptr;
                double acc = 0.0;
                if (vec) {
                    const int64_t len2 = len >> 1;
                    double acc1 = 0.0;
                    double2* w2 = reinterpret_cast<double2*>(wsm);
                    auto one = [&](int64_t e, const double2 pv, const double2 cv) {
                        double2 wv = w2[e];
                        if (vp) {
                            wv.x = sub_rn(wv.x, mul_rn(pv.x, hprev));
                            wv.y = sub_rn(wv.y, mul_rn(pv.y, hprev));
                            w2[e] = wv;
                        }
                        acc = add_rn(acc, vc ? mul_rn(cv.x, wv.x) : mul_rn(wv.x, wv.x));
                        acc1 = add_rn(acc1, vc ? mul_rn(cv.y, wv.y) : mul_rn(wv.y, wv.y));
                    };
                    const double2 z2 = make_double2(0.0, 0.0);
                    // L2 priorities follow the reuse: column i is read again by the next pass (as the update column), so its
                    // lines are kept (evict_last, like the prefetch below that brings them in); column i-1 is read for the
                    // last time here (evict_first).  With the default / streaming hints three 33 MB column slices thrash
                    // the L2 at 4 M rows per GPU and the "second read from L2" comes from HBM.
                    unsigned long long pol_first, pol_last;
                    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
                    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
                    auto ldp = [&](int64_t e) -> double2 {
                        double2 v = z2;
                        if (vp) asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(reinterpret_cast<const double2*>(vp) + e), "l"(pol_first));
                        return v;
                    };
                    auto ldc = [&](int64_t e) -> double2 {
                        double2 v = z2;
                        if (vc) asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(reinterpret_cast<const double2*>(vc) + e), "l"(pol_last));
                        return v;
                    };
                    // eight entries per thread and trip, predicated: sixteen 16-byte loads in flight before anything is
                    // consumed, and the last trip is as wide as the first (a 14 k-row slice is one or two trips)
                    for (int64_t e = threadIdx.x; e < len2; e += 8 * BLOCK) {
                        double2 pp[8], qq[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const bool in = e + (int64_t)u * BLOCK < len2;
                            pp[u] = in ? ldp(e + u * BLOCK) : z2;
                            qq[u] = in ? ldc(e + u * BLOCK) : z2;
                        }
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            if (e + (int64_t)u * BLOCK < len2) one(e + u * BLOCK, pp[u], qq[u]);
                    }
                    acc = add_rn(acc, acc1);
                    if ((len & 1) && threadIdx.x == 0) {
                        const int64_t t = len - 1;
                        double wv = wsm[t];
                        if (vp) {
                            wv = sub_rn(wv, mul_rn(vp[t], hprev));
                            wsm[t] = wv;
                        }
                        acc = add_rn(acc, vc ? mul_rn(vc[t], wv) : mul_rn(wv, wv));
                    }
                } else {
                    for (int64_t e = threadIdx.x; e < len; e += BLOCK) {
                        double wv = wsm[e];
                        if (vp) {
                            wv = sub_rn(wv, mul_rn(vp[e], hprev));
                            wsm[e] = wv;
                        }
                        acc = add_rn(acc, vc ? mul_rn(vc[e], wv) : mul_rn(wv, wv));
                    }
                }
                if (i + 1 < a.j) {  // the next column's slice starts its way from HBM into L2 while the sum is in flight
                    const char* nx = reinterpret_cast<const char*>(a.V + (int64_t)(i + 1) * a.ldv + c0);
                    for (int64_t o = (int64_t)threadIdx.x * 128; o < len * 8; o += (int64_t)BLOCK * 128)
                        asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(nx + o));
                }
                const double h = finish_tree(acc, i);
                if (i < a.j) {
                    if (blockIdx.x == 0 && threadIdx.x == 0) a.h_out[i] = h;  // H(i,j)   (:75)
                    hprev = h;
                } else {
                    const double nrm = sqrt(h);  // (:79)
                    if (blockIdx.x == 0 && threadIdx.x == 0) {
                        a.h_out[a.j] = nrm;
                        if (a.stop && (nrm == 0.0 || nrm < a.skip_below)) { a.stop[1] = a.j; a.stop[0] = 1; }
                    }
                    if (nrm != 0.0 && !(nrm < a.skip_below))  // (finish_tree ended with a block-wide barrier: w is complete)
                        for (int64_t e = threadIdx.x; e < len; e += BLOCK) a.vnext[c0 + e] = __ddiv_rn(wsm[e], nrm);  // (:86)
                }
            }
        }
    }
}

// RING variant (slices of 8*BLOCK < rows <= 28*BLOCK whose TWO column slices fit in shared memory - the 14 170-row slices of
// a 2.1 M-row rank, i.e. BASELINE config 5 on 8 GPUs): the thread's rows of w live in REGISTERS for the whole step and the
// columns arrive through a two-deep shared-memory ring filled by cp.async (16 bytes per request, L2 only).  Column i+1 is
// requested by each thread, entry by entry, the moment pass i has read the entry it replaces (of column i-1), so it
// travels HBM -> shared memory while pass i and the reduction of projection i are in flight, and a pass reads shared
// memory only: no global load latency and no second (L2) read of a column on the critical path.  Every ring entry is written (cp.async) and read
// by the same thread, so the only block-wide barriers are those of the reductions.
constexpr int kMgsRingSlots = 14;  // 16-byte slots per thread
template <int BLOCK, int S>
__global__ void __launch_bounds__(BLOCK) arnoldi_mgs_ring_kernel(MgsArgs a) {
    extern __shared__ __align__(16) double wsm[];  // two column slices of a.chunk doubles
    __shared__ double rank_vals[kMaxRanks];
    if (a.stop && *reinterpret_cast<volatile int*>(a.stop)) return;
    const int64_t c0 = (int64_t)blockIdx.x * a.chunk;
    const int64_t len = max((int64_t)0, min(a.chunk, a.n - c0));  // even: the launcher takes this kernel for even n only
    unsigned target = 0;
    // Slot u of this thread is the 16-byte entry threadIdx.x + u*BLOCK of the slice.  Which slots exist is one bit mask,
    // and every address below is a per-thread base plus a compile-time multiple of BLOCK: the pass is bounded by
    // instruction issue (16 warps, 14 slots each), so nothing is recomputed per slot.
    const int len2 = (int)(len >> 1);
    unsigned full = 0;
#pragma unroll
    for (int u = 0; u < S; ++u) full |= ((int)threadIdx.x + u * BLOCK < len2) ? (1u << u) : 0u;
    const double2* const buf0 = reinterpret_cast<const double2*>(wsm) + threadIdx.x;            // columns 0, 2, 4, ...
    const double2* const buf1 = reinterpret_cast<const double2*>(wsm + a.chunk) + threadIdx.x;  // columns 1, 3, 5, ...
    const unsigned sb0 = (unsigned)__cvta_generic_to_shared(buf0), sb1 = (unsigned)__cvta_generic_to_shared(buf1);
    const double2* const vcol = reinterpret_cast<const double2*>(a.V + c0) + threadIdx.x;
    const int64_t ldv2 = a.ldv >> 1;  // column stride in 16-byte entries
    auto request = [&](int col) {
        const double2* src = vcol + (int64_t)col * ldv2;
        const unsigned dst = (col & 1) ? sb1 : sb0;
#pragma unroll
        for (int u = 0; u < S; ++u)
            if (full & (1u << u)) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + u * BLOCK * 16), "l"(src + u * BLOCK) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (a.j > 0) request(0);
    if (a.j > 1) request(1);
    double2 wr[S];
    {
        const double2* w2 = reinterpret_cast<const double2*>(a.w + c0) + threadIdx.x;
#pragma unroll
        for (int u = 0; u < S; ++u) wr[u] = (full & (1u << u)) ? w2[u * BLOCK] : make_double2(0.0, 0.0);
    }
    double hprev = 0.0;
#ifdef ATK_REDUCE_PROFILE
    long long pt_wait = 0, pt_pass = 0, pt_red = 0;  // thread 0's clocks: waiting for the column, the pass, the reduction
    long long pt_sta[3] = {0, 0, 0};                 // ... and inside the reduction: block tree / arrive + wait / partials + tree
#endif
    for (int i = 0; i <= a.j; ++i) {
#ifdef ATK_REDUCE_PROFILE
        const long long pc0 = clock64();
#endif
        if (i < a.j) {
            if (i == 0 && a.j > 1) asm volatile("cp.async.wait_group 1;" ::: "memory");
            else asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
#ifdef ATK_REDUCE_PROFILE
        const long long pc1 = clock64();
#endif
        const double2* const bp = (i & 1) ? buf0 : buf1;  // column i-1
        const double2* const bc = (i & 1) ? buf1 : buf0;  // column i
        // Column i+1 goes into the buffer column i-1 is leaving, and each slot is requested the moment this thread has read
        // it - not after the pass: the HBM fetch of a column (2.6 us of bandwidth plus latency) then has the pass AND the
        // reduction to hide behind (6.74 -> 5.49 us per projection).  A prefetch of column i+2 into L2 on top, after the pass
        // or at its start: slower (5.90 / 5.27 vs 5.03).
        const bool refill = i >= 1 && i + 1 < a.j;
        // ... except in the warp whose thread 0 hands over the CTA's sum: a fence waits for the thread's outstanding cp.async
        // too, i.e. for HBM, so that warp sends its requests once the sum is on its way (after_arrive below)
        const bool refill_now = refill && threadIdx.x >= 32;
        const double2* const nsrc = vcol + (int64_t)(i + 1) * ldv2;
        const unsigned ndst = (i & 1) ? sb0 : sb1;
        double acc = 0.0, acc1 = 0.0;
        if (i == 0 && a.j > 0) {                                           // products of projection 0   (:75)
#pragma unroll
            for (int u = 0; u < S; ++u)
                if (full & (1u << u)) {
                    const double2 cv = bc[u * BLOCK];
                    acc = add_rn(acc, mul_rn(cv.x, wr[u].x));
                    acc1 = add_rn(acc1, mul_rn(cv.y, wr[u].y));
                }
        } else if (i < a.j) {                                              // update of projection i-1 (:76), products of i (:75)
#pragma unroll
            for (int u = 0; u < S; ++u)
                if (full & (1u << u)) {
                    const double2 pv = bp[u * BLOCK];
                    const double2 cv = bc[u * BLOCK];
                    if (refill_now) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ndst + u * BLOCK * 16), "l"(nsrc + u * BLOCK) : "memory");
                    double2 wv = wr[u];
                    wv.x = sub_rn(wv.x, mul_rn(pv.x, hprev));
                    wv.y = sub_rn(wv.y, mul_rn(pv.y, hprev));
                    wr[u] = wv;
                    acc = add_rn(acc, mul_rn(cv.x, wv.x));
                    acc1 = add_rn(acc1, mul_rn(cv.y, wv.y));
                }
        } else {                                                           // last update (none when j == 0), ||w||^2   (:79)
#pragma unroll
            for (int u = 0; u < S; ++u)
                if (full & (1u << u)) {
                    double2 wv = wr[u];
                    if (i > 0) {
                        const double2 pv = bp[u * BLOCK];
                        wv.x = sub_rn(wv.x, mul_rn(pv.x, hprev));
                        wv.y = sub_rn(wv.y, mul_rn(pv.y, hprev));
                        wr[u] = wv;
                    }
                    acc = add_rn(acc, mul_rn(wv.x, wv.x));
                    acc1 = add_rn(acc1, mul_rn(wv.y, wv.y));
                }
        }
        acc = add_rn(acc, acc1);
        auto late_refill = [&]() {
            if (refill && !refill_now) {
#pragma unroll
                for (int u = 0; u < S; ++u)
                    if (full & (1u << u)) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ndst + u * BLOCK * 16), "l"(nsrc + u * BLOCK) : "memory");
            }
            if (refill) asm volatile("cp.async.commit_group;" ::: "memory");
        };
#ifdef ATK_REDUCE_PROFILE
        const long long pc2 = clock64();
        const double h = mgs_finish_tree<BLOCK>(a, acc, i, target, rank_vals, pt_sta, late_refill);
#else
        const double h = mgs_finish_tree<BLOCK>(a, acc, i, target, rank_vals, nullptr, late_refill);
#endif
#ifdef ATK_REDUCE_PROFILE
        pt_wait += pc1 - pc0;
        pt_pass += pc2 - pc1;
        pt_red += clock64() - pc2;
        if (i == a.j && threadIdx.x == 0 && a.stamps) {
            atomicAdd(a.stamps + 4 * blockIdx.x + 0, (unsigned long long)pt_wait);
            atomicAdd(a.stamps + 4 * blockIdx.x + 1, (unsigned long long)pt_pass);
            atomicAdd(a.stamps + 4 * blockIdx.x + 2, (unsigned long long)pt_red);
            atomicAdd(a.stamps + 4 * blockIdx.x + 3, (unsigned long long)(a.j + 1));
            for (int k = 0; k < 3; ++k) atomicAdd(a.stamps + 4 * gridDim.x + 3 * blockIdx.x + k, (unsigned long long)pt_sta[k]);
        }
#endif
        if (i < a.j) {
            if (blockIdx.x == 0 && threadIdx.x == 0) a.h_out[i] = h;  // H(i,j)
            hprev = h;
        } else {
            const double nrm = sqrt(h);
            if (blockIdx.x == 0 && threadIdx.x == 0) {
                a.h_out[a.j] = nrm;
                if (a.stop && (nrm == 0.0 || nrm < a.skip_below)) { a.stop[1] = a.j; a.stop[0] = 1; }
            }
            if (nrm != 0.0 && !(nrm < a.skip_below)) {
                double2* vn = reinterpret_cast<double2*>(a.vnext + c0) + threadIdx.x;
#pragma unroll
                for (int u = 0; u < S; ++u)
                    if (full & (1u << u)) vn[u * BLOCK] = make_double2(__ddiv_rn(wr[u].x, nrm), __ddiv_rn(wr[u].y, nrm));  // (:86)
            }
        }
    }
}

using MgsKernel = void (*)(MgsArgs);
inline MgsKernel mgs_kernel_for(int block, bool reg) {
    switch (block) {
        case 64: return reg ? arnoldi_mgs_kernel<64, true> : arnoldi_mgs_kernel<64, false>;
        case 128: return reg ? arnoldi_mgs_kernel<128, true> : arnoldi_mgs_kernel<128, false>;
        case 256: return reg ? arnoldi_mgs_kernel<256, true> : arnoldi_mgs_kernel<256, false>;
        default: return reg ? arnoldi_mgs_kernel<512, true> : arnoldi_mgs_kernel<512, false>;
    }
}

// Launch configuration of the persistent step kernel for slices of a system of n rows (single-GPU contexts), or
// usable == false when a slice does not fit on chip.
struct PersistentMgs {
    bool usable = false;
    int grid = 1, block = 512;
    MgsKernel kernel = nullptr;
    MgsKernel ring_kernel = nullptr;  // w in registers + cp.async column ring, when the slice allows it (else nullptr)
    int64_t chunk = 0;
    size_t smem = 0, ring_smem = 0;
    unsigned* d_bar = nullptr;
    double* d_part = nullptr;
    unsigned long long* d_stamps = nullptr;  // profile build only
    int n_bars = 0;
    ~PersistentMgs() {
#ifdef ATK_REDUCE_PROFILE
        if (d_stamps && ring_kernel) {  // thread 0 of every CTA: clocks per projection spent waiting for the column / in the pass / in the reduction
            std::vector<unsigned long long> h((size_t)grid * 7);
            if (cudaMemcpy(h.data(), d_stamps, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost) == cudaSuccess && h[3]) {
                const char* name[6] = {"column wait", "pass", "reduction", " block tree", " arrive+wait", " partials"};
                for (int k = 0; k < 6; ++k) {
                    double lo = 1e300, hi = 0, sum = 0;
                    for (int c = 0; c < grid; ++c) {
                        const double v = (double)(k < 3 ? h[(size_t)c * 4 + k] : h[(size_t)grid * 4 + (size_t)c * 3 + (k - 3)]) / (double)h[(size_t)c * 4 + 3];
                        lo = std::min(lo, v), hi = std::max(hi, v), sum += v;
                    }
                    std::fprintf(stderr, "[atk mgs ring] %-12s clocks per projection over %d CTAs: min %.0f mean %.0f max %.0f (CTA 0: %.0f)\n", name[k], grid, lo,
                                 sum / grid, hi, (double)(k < 3 ? h[k] : h[(size_t)grid * 4 + (k - 3)]) / (double)h[3]);
                }
            }
        }
#endif
        cudaFree(d_bar);
        cudaFree(d_part);
        cudaFree(d_stamps);
    }
    void setup(Context* ctx, int64_t n, int max_steps) {
        int mine = try_setup(ctx, n, max_steps) ? 1 : 0;
        if (ctx->world > 1) {  // every rank must take the same path: the kernels wait for each other
            std::vector<int> all((size_t)ctx->world);
            allgather_host_bytes(ctx, &mine, all.data(), sizeof(int));
            for (int v : all) mine = mine && v;
        }
        usable = mine != 0;
    }
    bool try_setup(Context* ctx, int64_t n, int max_steps) {
        if (n <= 0 || std::getenv("ATK_GMRES_NO_PERSISTENT")) return false;
        const bool seq = ctx->reduction == ATK_REDUCE_SEQUENTIAL;
        if (ctx->world > 1 && (!ctx->p2p || seq)) return false;  // distributed: peer-memory mode, tree sums only
        int64_t c = seq ? n : std::max<int64_t>(2048, (n + ctx->sm_count - 1) / ctx->sm_count);
        c = (c + 1) & ~(int64_t)1;
        const int64_t rows = std::min(c, n);
        block = rows <= 64 * 8 ? 64 : rows <= 128 * 8 ? 128 : rows <= 256 * 8 ? 256 : 512;
        const bool reg = rows <= (int64_t)block * kMgsRegRows;
        const size_t sm = sizeof(double) * (size_t)c * (reg ? (seq ? 1 : 0) : (seq ? 2 : 1));
        kernel = mgs_kernel_for(block, reg);
        int max_smem = 0;
        ATK_CUDA_CHECK(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
        if (sm + 1024 > (size_t)max_smem) return false;
        // per function and process-wide: always the device maximum, so that a later, smaller solve cannot lower the limit
        // under a live, larger one (two GMRES objects of different sizes)
        cudaFuncAttributes fa;
        ATK_CUDA_CHECK(cudaFuncGetAttributes(&fa, kernel));
        ATK_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            max_smem - (int)fa.sharedSizeBytes));
        int per_sm = 0;
        ATK_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, sm));
        const int g = (int)((n + c - 1) / c);
        if (per_sm < 1 || g > per_sm * ctx->sm_count) return false;
        int coop = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device);
        if (!coop) return false;
        if (!seq && !reg && block == 512 && rows <= (int64_t)2 * kMgsRingSlots * 512 && !std::getenv("ATK_MGS_NO_RING")) {
            MgsKernel rk = arnoldi_mgs_ring_kernel<512, kMgsRingSlots>;
            const size_t rsm = sizeof(double) * 2 * (size_t)c;
            cudaFuncAttributes rfa;
            ATK_CUDA_CHECK(cudaFuncGetAttributes(&rfa, rk));
            if (rsm + rfa.sharedSizeBytes <= (size_t)max_smem) {
                ATK_CUDA_CHECK(cudaFuncSetAttribute(rk, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem - (int)rfa.sharedSizeBytes));
                int rper = 0;
                ATK_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&rper, rk, 512, rsm));
                if (rper >= 1 && g <= rper * ctx->sm_count) {
                    ring_kernel = rk;
                    ring_smem = rsm;
                }
            }
        }
        grid = g;
        chunk = c;
        smem = sm;
        n_bars = max_steps + 1;
        ATK_CUDA_CHECK(cudaMalloc(&d_bar, sizeof(unsigned) * (size_t)n_bars));
        ATK_CUDA_CHECK(cudaMalloc(&d_part, sizeof(double) * 2 * (size_t)g));
#ifdef ATK_REDUCE_PROFILE
        ATK_CUDA_CHECK(cudaMalloc(&d_stamps, sizeof(unsigned long long) * 7 * (size_t)g));
        ATK_CUDA_CHECK(cudaMemset(d_stamps, 0, sizeof(unsigned long long) * 7 * (size_t)g));
#endif
        return true;
    }
    void reset_barriers(cudaStream_t st) { ATK_CUDA_CHECK(cudaMemsetAsync(d_bar, 0, sizeof(unsigned) * (size_t)n_bars, st)); }
    // projections of w against columns [0, ncols) of V, then the norm and V_next = w / norm; h_out[0..ncols] on the device
    void launch(Context* ctx, int64_t n, int ncols, int step, const double* V, int64_t ldv, const double* w, double* vnext,
                double* h_out, double skip_below, int* stop = nullptr) {
        MgsArgs ma{n, ncols, V, ldv, w, vnext, h_out, d_part, d_bar + step, chunk, ctx->reduction == ATK_REDUCE_SEQUENTIAL ? 1 : 0,
                   skip_below, stop, ctx->world, ctx->rank, ctx->d_ll, ctx->d_peer_ll, ctx->epoch, d_stamps};
        if (ctx->world > 1) ctx->epoch += (unsigned long long)ncols + 2ull;  // same advance on every rank
        void* kargs[] = {&ma};
        // the ring variant moves 16-byte pieces: every column slice, w and V_next must start on a 16-byte boundary
        const bool ring = ring_kernel && (n & 1) == 0 && (ldv & 1) == 0 && (chunk & 1) == 0 &&
                          ((reinterpret_cast<uintptr_t>(V) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(vnext)) & 15u) == 0;
        ATK_CUDA_CHECK(cudaLaunchCooperativeKernel((void*)(ring ? ring_kernel : kernel), dim3(grid), dim3(block), kargs,
                                                   ring ? ring_smem : smem, ctx->stream));
        count_launch(ctx);
    }
};

struct Givens {
    // reference GMRES.hpp:139-151
    static void generate(double dx, double dy, double& cs, double& sn, double& rho) {
        rho = std::sqrt(dx * dx + dy * dy);
        if (std::fabs(rho) < DBL_EPSILON) {
            cs = 1.0;
            sn = 0.0;
            rho = 0.0;
        } else {
            cs = dx / rho;
            sn = dy / rho;
        }
    }
};

}  // namespace

void gmres_solve(Context* ctx, Operator* op, Preconditioner* pc, const double* b, double* x, int64_t n_x, int max_iter,
                 int krylov_dim, double tol, atk_gmres_ortho ortho, atk_gmres_callback cb, void* user, atk_gmres_info* info) {
    ATK_REQUIRE(op && b && x && info, ATK_ERR_INVALID_ARGUMENT, "null argument to gmres_solve");
    ATK_NVTX("atk::gmres_solve");
    const int64_t n = op->n_rows_local;
    const int64_t n_global = op->n_cols;
    // GMRES.hpp:28-37
    ATK_REQUIRE(n_x == n, ATK_ERR_INVALID_ARGUMENT, "Initial guess vector must match system size");
    ATK_REQUIRE(krylov_dim > 0 && krylov_dim <= n_global, ATK_ERR_INVALID_ARGUMENT, "Invalid Krylov subspace dimension");
    ATK_REQUIRE(ortho == ATK_GMRES_MGS || ortho == ATK_GMRES_MGS_BATCHED, ATK_ERR_INVALID_ARGUMENT, "unknown orthogonalisation mode");
    ATK_REQUIRE(!pc || (pc->ctx == ctx && pc->n == n && ctx->world == 1), ATK_ERR_INVALID_ARGUMENT,
                "Vector size does not match matrix size");  // ILU.hpp:88-90
    const int K = krylov_dim;
    const int64_t ld = K + 1;
    cudaStream_t st = ctx->stream;
    const int64_t launches0 = ctx->launches;
    auto notify = [&](int ev, double v) { if (cb) cb(ev, v, user); };

    double *V = nullptr, *w = nullptr, *r = nullptr, *t = nullptr, *d_h = nullptr, *d_y = nullptr, *d_part = nullptr, *d_cg = nullptr;
    double* h_col = nullptr;  // pinned: H column + norm coming back each Arnoldi step
    auto cleanup = [&]() {
        cudaFree(V); cudaFree(w); cudaFree(r); cudaFree(t); cudaFree(d_h); cudaFree(d_y); cudaFree(d_part); cudaFree(d_cg);
        if (h_col) cudaFreeHost(h_col);
    };
    info->restarts = 0;
    info->status = 0;
    info->arnoldi_steps = 0;
    info->trace_count = 0;
    auto trace = [&](double v) {
        if (info->residual_trace && info->trace_count < info->trace_capacity) info->residual_trace[info->trace_count] = v;
        info->trace_count++;
    };
    try {
        const size_t nn = (size_t)std::max<int64_t>(n, 1);
        ATK_CUDA_CHECK(cudaMalloc(&V, sizeof(double) * nn * (size_t)ld));
        ATK_CUDA_CHECK(cudaMalloc(&w, sizeof(double) * nn));
        ATK_CUDA_CHECK(cudaMalloc(&r, sizeof(double) * nn));
        ATK_CUDA_CHECK(cudaMalloc(&t, sizeof(double) * nn));
        ATK_CUDA_CHECK(cudaMalloc(&d_h, sizeof(double) * (size_t)(ld + 1)));
        ATK_CUDA_CHECK(cudaMalloc(&d_y, sizeof(double) * (size_t)ld));
        ATK_CUDA_CHECK(cudaMallocHost(&h_col, sizeof(double) * (size_t)(ld + 1)));
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 8));
        // batched MGS: per-CTA partials of the two multi-dots, the reduced [c | g] vector (all ranks), Gram matrix on the host
        const bool batched = ortho == ATK_GMRES_MGS_BATCHED;
        // per-projection path with tree sums: one fused update+products launch per projection (mgs_axmy_dot_kernel)
        const bool fuse_dots = ctx->reduction != ATK_REDUCE_SEQUENTIAL && !std::getenv("ATK_GMRES_NO_FUSED_DOT");
        const int bd_blocks = (int)std::max<int64_t>(1, (n + kBdBlock * kBdRows - 1) / (kBdBlock * kBdRows));
        const int cg_len = 2 * (K + 1);
        double* h_cg = nullptr;  // pinned [world * cg_len]
        std::vector<double> G, hvec((size_t)K + 1, 0.0);
        if (batched) {
            ATK_CUDA_CHECK(cudaMalloc(&d_part, sizeof(double) * (size_t)bd_blocks * (size_t)cg_len));
            ATK_CUDA_CHECK(cudaMalloc(&d_cg, sizeof(double) * (size_t)cg_len * (size_t)ctx->world));
            ATK_CUDA_CHECK(cudaMallocHost(&h_cg, sizeof(double) * (size_t)cg_len * (size_t)ctx->world));
            G.assign((size_t)ld * ld, 0.0);
        }
        struct PinFree { double*& p; ~PinFree() { if (p) cudaFreeHost(p); } } pin_free{h_cg};

        // ---- persistent MGS step: usable when every CTA's slice of w fits on chip (single-GPU contexts).  The whole
        // restart cycle is then queued without a host round trip: step j writes column j of H to d_hcols, the step whose
        // norm ends the cycle raises d_stop and the remaining launches return at once; the host reads everything back
        // once per cycle and runs the Givens recurrences over the columns in order - the same arithmetic, later.
        PersistentMgs pm;
        if (ortho == ATK_GMRES_MGS) pm.setup(ctx, n, K);
        const bool persistent = pm.usable;
        double *d_hcols = nullptr, *h_hcols = nullptr;
        int *d_stop = nullptr, *h_stop = nullptr;
        struct DeferFree {
            double*& a; double*& b; int*& c; int*& d;
            ~DeferFree() { cudaFree(a); cudaFree(c); if (b) cudaFreeHost(b); if (d) cudaFreeHost(d); }
        } defer_free{d_hcols, h_hcols, d_stop, h_stop};
        if (persistent) {
            ATK_CUDA_CHECK(cudaMalloc(&d_hcols, sizeof(double) * (size_t)K * (size_t)ld));
            ATK_CUDA_CHECK(cudaMallocHost(&h_hcols, sizeof(double) * (size_t)K * (size_t)ld));
            ATK_CUDA_CHECK(cudaMalloc(&d_stop, sizeof(int) * 2));
            ATK_CUDA_CHECK(cudaMallocHost(&h_stop, sizeof(int) * 2));
        }

        std::vector<double> H((size_t)ld * ld, 0.0);  // :55 - allocated once, never cleared
        std::vector<double> cs(K, 1.0), sn(K, 0.0), rho(K, 0.0), e1(ld, 0.0), y(ld, 0.0);
        auto Hm = [&](int i, int j) -> double& { return H[(size_t)i + (size_t)j * ld]; };

        ATK_CUDA_CHECK(cudaEventRecord(ctx->ev_t0, st));
        auto residual = [&]() -> double {  // r = b - A*x ; ||r||   (:42,47 / :99,104)
            op->apply(x, t, -1, nullptr);
            if (pc) {  // r = M^-1 (b - A x)   (:43-45 / :100-102)
                launch_sub(ctx, n, b, t, w);
                pc->solve(w, r, nullptr);
            } else {
                launch_sub(ctx, n, b, t, r);
            }
            launch_dot(ctx, n, r, r, kSlotNorm);
            return std::sqrt(read_slot(ctx, kSlotNorm));
        };
        double beta = residual();
        info->initial_residual = beta;
        info->final_residual = beta;
        trace(beta);
        notify(ATK_GMRES_INITIAL_RESIDUAL, beta);
        if (beta < tol) {  // :49-52
            notify(ATK_GMRES_INITIAL_GOOD, 0.0);
            info->status = 2;
        } else {
            e1[0] = beta;
            bool finished = false;
            for (int iter = 0; iter < max_iter && !finished; ++iter) {
                ATK_REQUIRE(beta != 0.0, ATK_ERR_RUNTIME, "Division by zero.");  // V[0] = r / beta (:63)
                launch_div(ctx, n, r, beta, V);
                int kry = K;
                // Givens recurrences for column j of H (entries 0..j-1 = projections, entry j = ||w||); false = the cycle ends
                auto absorb_column = [&](int j, const double* col) -> bool {
                    info->arnoldi_steps++;
                    for (int i = 0; i < j; ++i) Hm(i, j) = col[i];  // :75
                    const double w_norm = col[j];
                    Hm(j + 1, j) = w_norm;  // :80
                    if (w_norm < tol) {     // :81-85
                        kry = j;
                        notify(ATK_GMRES_KRYLOV_UPDATE, (double)kry);
                        return false;
                    }
                    ATK_REQUIRE(w_norm != 0.0, ATK_ERR_RUNTIME, "Division by zero.");  // VectorObj.hpp:76 via :86
                    for (int i = 0; i < j; ++i) {  // :89-91 / :134-137
                        const double tmp = Hm(i, j);
                        Hm(i, j) = cs[i] * Hm(i, j) + sn[i] * Hm(i + 1, j);
                        Hm(i + 1, j) = -sn[i] * tmp + cs[i] * Hm(i + 1, j);
                    }
                    Givens::generate(Hm(j, j), Hm(j + 1, j), cs[j], sn[j], rho[j]);  // :92
                    Hm(j, j) = rho[j];      // :93
                    Hm(j + 1, j) = 0.0;     // :94
                    {                       // :95 / :128-132
                        const double tmp = e1[j];
                        e1[j] = cs[j] * tmp + sn[j] * e1[j + 1];
                        e1[j + 1] = -sn[j] * tmp + cs[j] * e1[j + 1];
                    }
                    return true;
                };
                if (persistent) {
                    pm.reset_barriers(st);
                    ATK_CUDA_CHECK(cudaMemsetAsync(d_stop, 0, sizeof(int) * 2, st));
                    for (int j = 0; j < K; ++j) {
                        if (pc) {  // w = M^-1 (A v_j)   (:67-68)
                            op->apply(V + (size_t)j * nn, t, -1, d_stop);
                            pc->solve(t, w, d_stop);
                        } else {
                            op->apply(V + (size_t)j * nn, w, -1, d_stop);  // :70
                        }
                        pm.launch(ctx, n, j, j, V, (int64_t)nn, w, V + (size_t)(j + 1) * nn, d_hcols + (size_t)j * ld, tol, d_stop);
                    }
                    ATK_CUDA_CHECK(cudaGetLastError());
                    ATK_CUDA_CHECK(cudaMemcpyAsync(h_stop, d_stop, sizeof(int) * 2, cudaMemcpyDeviceToHost, st));
                    ATK_CUDA_CHECK(cudaMemcpyAsync(h_hcols, d_hcols, sizeof(double) * (size_t)K * (size_t)ld, cudaMemcpyDeviceToHost, st));
                    ATK_CUDA_CHECK(cudaStreamSynchronize(st));
                    const int steps = h_stop[0] ? h_stop[1] + 1 : K;
                    for (int j = 0; j < steps; ++j)
                        if (!absorb_column(j, h_hcols + (size_t)j * ld)) break;
                } else {
                    for (int j = 0; j < K; ++j) {
                        const double* vj = V + (size_t)j * nn;
                        if (pc) {  // w = M^-1 (A v_j)   (:67-68)
                            op->apply(vj, t, -1, nullptr);
                            pc->solve(t, w, nullptr);
                        } else {
                            op->apply(vj, w, -1, nullptr);  // :70
                        }
                        bool norm_done = false;
                        if (ortho == ATK_GMRES_MGS && fuse_dots && j > 0) {  // :73-79, update i-1 fused with the products of i
                            launch_dot(ctx, n, V, w, kSlotH);
                            double* const slot_h = ctx->slot(kSlotH);
                            for (int i = 1; i <= j; ++i) {
                                const double* vp = V + (size_t)(i - 1) * nn;
                                const double* vc = i < j ? V + (size_t)i * nn : nullptr;  // last launch: w . w for the norm
                                const bool vec = aligned16(vp) && aligned16(w) && (!vc || aligned16(vc));
                                const int64_t work = vec ? (n + 1) / 2 : n;
                                const int fgrid = (int)std::min<int64_t>(std::max<int64_t>((work + kReduceBlock - 1) / kReduceBlock, 1),
                                                                         (int64_t)ctx->reduce_grid());
                                const int out_slot = i < j ? kSlotH : kSlotNorm;
                                double* out = ctx->slot(out_slot) + ctx->rank;
                                if (vec)
                                    mgs_axmy_dot_kernel<true><<<fgrid, kReduceBlock, 0, st>>>(n, vp, vc, w, slot_h, ctx->world, d_h + (i - 1),
                                                                                              ctx->d_partials, ctx->d_tickets + 0, out);
                                else
                                    mgs_axmy_dot_kernel<false><<<fgrid, kReduceBlock, 0, st>>>(n, vp, vc, w, slot_h, ctx->world, d_h + (i - 1),
                                                                                               ctx->d_partials, ctx->d_tickets + 0, out);
                                count_launch(ctx);
                                // distributed: the same all-gather as after a dot launch.  It runs in stream order on every rank,
                                // so a rank's slot array is rewritten only after its own launch i has read h_(i-1) from it.
                                allgather_slot(ctx, out_slot, st);
                            }
                            norm_done = true;
                        } else if (ortho == ATK_GMRES_MGS) {
                            for (int i = 0; i < j; ++i) {  // :73-77
                                const double* vi = V + (size_t)i * nn;
                                launch_dot(ctx, n, vi, w, kSlotH);
                                mgs_axmy_kernel<<<grid, 256, 0, st>>>(n, vi, w, ctx->slot(kSlotH), ctx->world, d_h + i);
                                count_launch(ctx);
                            }
                        } else if (j > 0) {
                            // ---- batched MGS: pass 1 (c = V[0:j]^T w, g = V[0:j+1]^T v_j), triangular solve, pass 2
                            const int ncw = j, ncg = j + 1, len = 2 * ncg;
                            multi_dot2_kernel<<<bd_blocks, kBdBlock, 0, st>>>(n, ncw, ncg, V, (int64_t)nn, w, vj, d_part);
                            multi_dot2_final_kernel<<<len, 256, 0, st>>>(ncw, ncg, bd_blocks, d_part, d_cg + (size_t)ctx->rank * len);
                            count_launch(ctx, 2);
                            allgather_doubles(ctx, d_cg, (size_t)len, st);
                            ATK_CUDA_CHECK(cudaMemcpyAsync(h_cg, d_cg, sizeof(double) * (size_t)len * ctx->world, cudaMemcpyDeviceToHost, st));
                            ATK_CUDA_CHECK(cudaStreamSynchronize(st));
                            auto Gm = [&](int a, int b) -> double& { return G[(size_t)a + (size_t)b * ld]; };
                            for (int i = 0; i < ncg; ++i) {  // rank-ordered sums: identical on every rank
                                double c = 0.0, g = 0.0;
                                for (int q = 0; q < ctx->world; ++q) {
                                    c = c + h_cg[(size_t)q * len + i];
                                    g = g + h_cg[(size_t)q * len + ncg + i];
                                }
                                hvec[(size_t)i] = c;  // c_i for i < j (entry j unused)
                                Gm(j, i) = g;         // G(j,i) = v_j . v_i
                                Gm(i, j) = g;
                            }
                            for (int i = 0; i < j; ++i) {  // (I + strict_lower(G)) h = c, forward substitution
                                double hi = hvec[(size_t)i];
                                for (int m = 0; m < i; ++m) hi = hi - Gm(i, m) * hvec[(size_t)m];
                                hvec[(size_t)i] = hi;
                            }
                            ATK_CUDA_CHECK(cudaMemcpyAsync(d_h, hvec.data(), sizeof(double) * (size_t)j, cudaMemcpyHostToDevice, st));
                            multi_axmy_norm_kernel<<<ctx->reduce_grid(), kReduceBlock, sizeof(double) * (size_t)j, st>>>(
                                n, j, V, (int64_t)nn, d_h, w, ctx->d_partials, ctx->d_tickets + 6, ctx->slot(kSlotNorm) + ctx->rank);
                            count_launch(ctx);
                            allgather_slot(ctx, kSlotNorm, st);
                        }
                        if (!(batched && j > 0) && !norm_done) launch_dot(ctx, n, w, w, kSlotNorm);  // :79 (the batched pass 2 already summed w.w)
                        normalize_kernel<<<grid, 256, 0, st>>>(n, w, V + (size_t)(j + 1) * nn, ctx->slot(kSlotNorm), ctx->world, d_h + j,
                                                               tol);  // :86 (skipped when the norm is 0 or below tol)
                        count_launch(ctx);
                        ATK_CUDA_CHECK(cudaGetLastError());
                        ATK_CUDA_CHECK(cudaMemcpyAsync(h_col, d_h, sizeof(double) * (size_t)(j + 1), cudaMemcpyDeviceToHost, st));
                        ATK_CUDA_CHECK(cudaStreamSynchronize(st));
                        if (!absorb_column(j, h_col)) break;
                    }
                }
                // updateSolution (:153-171)
                for (int i = kry - 1; i >= 0; --i) {
                    y[i] = e1[i];
                    for (int jj = i + 1; jj < kry; ++jj) y[i] = y[i] - Hm(i, jj) * y[jj];
                    ATK_REQUIRE(!(std::fabs(Hm(i, i)) < DBL_EPSILON), ATK_ERR_RUNTIME,
                                "GMRES updateSolution: Zero diagonal element encountered in H.");
                    y[i] = y[i] / Hm(i, i);
                }
                if (kry > 0) {
                    ATK_CUDA_CHECK(cudaMemcpyAsync(d_y, y.data(), sizeof(double) * (size_t)kry, cudaMemcpyHostToDevice, st));
                    update_solution_kernel<<<grid, 256, 0, st>>>(n, kry, V, (int64_t)nn, d_y, x);
                    count_launch(ctx);
                    ATK_CUDA_CHECK(cudaStreamSynchronize(st));  // y (pageable) must stay valid until the copy ran
                }
                beta = residual();  // :99-104
                info->restarts = iter + 1;
                info->final_residual = beta;
                trace(beta);
                notify(ATK_GMRES_RESTART_RESIDUAL, beta);
                if (beta < tol) {  // :107-110
                    notify(ATK_GMRES_CONVERGED, (double)iter);
                    info->status = 1;
                    finished = true;
                    break;
                }
                std::fill(e1.begin(), e1.end(), 0.0);  // :111-112
                e1[0] = beta;
            }
            if (!finished) notify(ATK_GMRES_MAX_ITER, 0.0);  // :114
        }
        ATK_CUDA_CHECK(cudaEventRecord(ctx->ev_t1, st));
        ATK_CUDA_CHECK(cudaStreamSynchronize(st));
        float ms = 0.f;
        ATK_CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->ev_t0, ctx->ev_t1));
        info->device_ms = ms;
        info->kernel_launches = (int)(ctx->launches - launches0);
    } catch (...) {
        cudaStreamSynchronize(st);
        cleanup();
        throw;
    }
    cleanup();
}

// ---------------------------------------------------------------------------------------------- Krylov::Arnoldi
// Reference src/LinearAlgebra/Krylov/KrylovSubspace.hpp:11-38 ("next" row): textbook MGS Arnoldi.
//     for i = 1..m-1:  Av = A*Q[i-1];  for j<i { h = Q[j].Av; H(j,i-1) = h; Av -= Q[j]*h }
//                      h = ||Av||; if h < tol { breakdown at i; stop }  H(i,i-1) = h;  Q[i] = Av/h
// Same kernels as the GMRES step (persistent on-chip kernel where the slices fit, per-projection kernels otherwise).
// Q: m columns (ldq apart) on the device, Q[0] given; H: host, m x (m-1) column-major - only the entries the reference
// writes are touched.  *breakdown = i of the breakdown, or -1.
void arnoldi(Context* ctx, Operator* op, double* Q, int64_t ldq, int m, double* H, double tol, int* breakdown) {
    ATK_REQUIRE(op && Q && H && breakdown && m >= 1, ATK_ERR_INVALID_ARGUMENT, "bad argument to arnoldi");
    ATK_NVTX("atk::arnoldi");
    const int64_t n = op->n_rows_local;
    ATK_REQUIRE(ldq >= n, ATK_ERR_INVALID_ARGUMENT, "Q columns overlap");
    cudaStream_t st = ctx->stream;
    *breakdown = -1;
    double *w = nullptr, *d_h = nullptr, *h_col = nullptr;
    auto cleanup = [&]() {
        cudaFree(w);
        cudaFree(d_h);
        if (h_col) cudaFreeHost(h_col);
    };
    try {
        ATK_CUDA_CHECK(cudaMalloc(&w, sizeof(double) * (size_t)std::max<int64_t>(n, 1)));
        ATK_CUDA_CHECK(cudaMalloc(&d_h, sizeof(double) * (size_t)(m + 1)));
        ATK_CUDA_CHECK(cudaMallocHost(&h_col, sizeof(double) * (size_t)(m + 1)));
        PersistentMgs pm;
        pm.setup(ctx, n, m);
        if (pm.usable) pm.reset_barriers(st);
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 8));
        for (int i = 1; i < m; ++i) {
            op->apply(Q + (size_t)(i - 1) * ldq, w, -1, nullptr);
            double* qi = Q + (size_t)i * ldq;
            if (pm.usable) {
                pm.launch(ctx, n, i, i, Q, ldq, w, qi, d_h, tol);
            } else {
                for (int j = 0; j < i; ++j) {
                    const double* qj = Q + (size_t)j * ldq;
                    launch_dot(ctx, n, qj, w, kSlotH);
                    mgs_axmy_kernel<<<grid, 256, 0, st>>>(n, qj, w, ctx->slot(kSlotH), ctx->world, d_h + j);
                    count_launch(ctx);
                }
                launch_dot(ctx, n, w, w, kSlotNorm);
                normalize_kernel<<<grid, 256, 0, st>>>(n, w, qi, ctx->slot(kSlotNorm), ctx->world, d_h + i, tol);
                count_launch(ctx);
            }
            ATK_CUDA_CHECK(cudaGetLastError());
            ATK_CUDA_CHECK(cudaMemcpyAsync(h_col, d_h, sizeof(double) * (size_t)(i + 1), cudaMemcpyDeviceToHost, st));
            ATK_CUDA_CHECK(cudaStreamSynchronize(st));
            for (int j = 0; j < i; ++j) H[(size_t)j + (size_t)(i - 1) * (size_t)m] = h_col[j];
            const double nrm = h_col[i];
            if (nrm < tol) {
                *breakdown = i;
                break;
            }
            H[(size_t)i + (size_t)(i - 1) * (size_t)m] = nrm;
            ATK_REQUIRE(nrm != 0.0, ATK_ERR_RUNTIME, "Division by zero.");
        }
    } catch (...) {
        cudaStreamSynchronize(st);
        cleanup();
        throw;
    }
    cleanup();
}

}  // namespace atk
