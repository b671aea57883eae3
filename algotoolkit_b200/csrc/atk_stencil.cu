// atk_stencil.cu -- matrix-free operator of Poisson3DSolver (reference src/PDEs/FDM/Poisson.hpp:49-95).
//
// The reference assembles a CSC matrix whose Dirichlet-shell rows are identity rows (:73) and whose interior rows
// hold 7 taps (:76-88), then multiplies through SparseObj.hpp:131-152, i.e. each y[row] is
//     ((((((0.0 + cz*x[k-1]) + cy*x[j-1]) + cx*x[i-1]) + cc*x[c]) + cx*x[i+1]) + cy*x[j+1]) + cz*x[k+1]
// with every product and sum rounded separately (ascending column order: idx = i + j*nx + k*nx*ny, :98-100).
// This kernel evaluates exactly that expression, so its output is bit-identical to the assembled matvec.  (The
// reference also skips columns whose x is exactly 0.0; adding the resulting +-0.0 products instead never changes a
// sum that started at +0.0, so the skip is not replicated here.)
//
// Data movement: the grid is cut into (2*TX) x TY tiles in the xy-plane and chunks of `kc` planes in z.  A thread
// owns two adjacent x points of one row and marches through its chunk keeping the k-1 / k / k+1 values in registers,
// so every x element is read once per chunk from HBM/L2 as a 16-byte load; the j+-1 and i+-1 neighbours come from L1
// (they are the centre values of the same CTA's other threads).  Algorithmic traffic: 16 B per grid point.
//
// Distributed contexts own a z-slab [k_begin,k_end); the planes k_begin-1 and k_end come from halo buffers that
// exchange_halo() fills on the communication stream while the interior planes are being computed.
#include <cuda.h>  // CUtensorMap (the encode function itself is fetched with cudaGetDriverEntryPoint: no libcuda link)

#include <algorithm>
#include <cstdlib>
#include <string>

#include "atk_internal.cuh"

namespace atk {

namespace {

struct StencilGeom {
    int nx, ny, nz;      // global grid
    int k_off;           // global index of local plane 0
    int nz_loc;          // local planes
    double cx, cy, cz, cc;
};

__device__ __forceinline__ double tap7(const StencilGeom& g, double km, double jm, double im, double c, double ip, double jp,
                                       double kp) {
    double acc = 0.0;
    acc = add_rn(acc, mul_rn(g.cz, km));
    acc = add_rn(acc, mul_rn(g.cy, jm));
    acc = add_rn(acc, mul_rn(g.cx, im));
    acc = add_rn(acc, mul_rn(g.cc, c));
    acc = add_rn(acc, mul_rn(g.cx, ip));
    acc = add_rn(acc, mul_rn(g.cy, jp));
    acc = add_rn(acc, mul_rn(g.cz, kp));
    return acc;
}
// identity row: 0.0 + 1.0*x
__device__ __forceinline__ double shell_row(double c) { return add_rn(0.0, mul_rn(1.0, c)); }

// Planes [k_lo, k_hi) of the local slab, cut into chunks of kc planes along blockIdx.z.
// VEC2: nx is even and x/y are 16-byte aligned, so the two points are moved as one double2.
template <int TX, int TY, bool VEC2, bool DOT>
__global__ void __launch_bounds__(TX* TY) stencil7_kernel(StencilGeom g, const double* __restrict__ x, double* __restrict__ y,
                                                           const double* __restrict__ halo_lo,
                                                           const double* __restrict__ halo_hi, int k_lo, int k_hi, int kc,
                                                           double* partials, int partial_base, unsigned* ticket,
                                                           double* dot_out, int dot_final, const int* skip_flag) {
    if (skip_flag && *skip_flag) return;
    const int i0 = (blockIdx.x * TX + threadIdx.x) * 2;
    const int j = blockIdx.y * TY + threadIdx.y;
    const int kb = k_lo + blockIdx.z * kc;
    const int ke = min(kb + kc, k_hi);
    const bool active = (i0 < g.nx) && (j < g.ny);
    const bool has1 = active && (i0 + 1 < g.nx);  // second point exists (odd nx)
    const int64_t sxy = (int64_t)g.nx * g.ny;
    const int64_t row = (int64_t)j * g.nx + i0;
    double dacc = 0.0;

    auto load2 = [&](const double* base, int64_t idx) -> double2 {
        if constexpr (VEC2) return *reinterpret_cast<const double2*>(base + idx);
        else return make_double2(base[idx], has1 ? base[idx + 1] : 0.0);
    };
    // plane k of the local slab, or a halo plane when k is outside it
    auto load_plane = [&](int k) -> double2 {
        if (k < 0) return load2(halo_lo, row);
        if (k >= g.nz_loc) return load2(halo_hi, row);
        return load2(x, (int64_t)k * sxy + row);
    };

    if (active && kb < ke) {
        const bool j_shell = (j == 0 || j == g.ny - 1);
        const bool a_shell_i = (i0 == 0 || i0 == g.nx - 1);
        const bool b_shell_i = (i0 + 1 == g.nx - 1);
        double2 prev = make_double2(0.0, 0.0), cur, next = make_double2(0.0, 0.0);
        if (g.k_off + kb - 1 >= 0) prev = load_plane(kb - 1);
        cur = load_plane(kb);
        for (int k = kb; k < ke; ++k) {
            const int kg = g.k_off + k;
            if (kg + 1 < g.nz) next = load_plane(k + 1);
            const int64_t idx = (int64_t)k * sxy + row;
            double2 out;
            if (j_shell || kg == 0 || kg == g.nz - 1) {
                out.x = shell_row(cur.x);
                out.y = shell_row(cur.y);
            } else {
                const double2 jm = load2(x, idx - g.nx);
                const double2 jp = load2(x, idx + g.nx);
                if (a_shell_i) {
                    out.x = shell_row(cur.x);
                } else {
                    const double left = x[idx - 1];
                    out.x = tap7(g, prev.x, jm.x, left, cur.x, cur.y, jp.x, next.x);
                }
                if (!has1) {
                    out.y = 0.0;
                } else if (b_shell_i) {
                    out.y = shell_row(cur.y);
                } else {
                    const double right = x[idx + 2];
                    out.y = tap7(g, prev.y, jm.y, cur.x, cur.y, right, jp.y, next.y);
                }
            }
            if constexpr (VEC2) {
                *reinterpret_cast<double2*>(y + idx) = out;
            } else {
                y[idx] = out.x;
                if (has1) y[idx + 1] = out.y;
            }
            if constexpr (DOT) {
                dacc = add_rn(dacc, mul_rn(cur.x, out.x));
                if (has1) dacc = add_rn(dacc, mul_rn(cur.y, out.y));
            }
            prev = cur;
            cur = next;
        }
    }
    if constexpr (DOT) {
        // partials of this launch go to [partial_base, partial_base + blocks); only the launch flagged dot_final
        // (the last one of an apply) elects a block to add up partials [0, partial_base + blocks).
        const double bs = block_sum<TX * TY>(dacc);
        __shared__ bool is_last;
        const unsigned nb = total_blocks();
        if (linear_thread_id() == 0) {
            partials[partial_base + linear_block_id()] = bs;
            is_last = false;
            if (dot_final) {
                __threadfence();
                const unsigned t = atomicInc(ticket, nb - 1);
                is_last = (t == nb - 1);
            }
        }
        __syncthreads();
        if (is_last) {
            __threadfence();
            double s = 0.0;
            const unsigned tot = partial_base + nb;
            for (unsigned q = linear_thread_id(); q < tot; q += TX * TY) s = add_rn(s, __ldcg(partials + q));
            const double total = block_sum<TX * TY>(s);
            if (linear_thread_id() == 0) *dot_out = total;
        }
    }
}


// ---------------------------------------------------------------------------------------------- fused CG step KA
// One pass that does what K3 + K1 + half of K2 did in the three-kernel formulation (ConjugateGradient.hpp:50-51,60,79):
//     p_new = r + p*beta          (beta = rs_new/rs_old, formed from the r.r slot; p_new = p on the first step)
//     x     = x + p*alpha         (the update of the PREVIOUS iteration, deferred so that p is read only once)
//     Ap    = A p_new ,  partial sums of p_new.Ap
// p_new at the six neighbours is recomputed from r and p there (one multiply-add each) instead of being re-read
// from a finished p_new array, so p is read once and written once: 48 B/row here + 24 B/row in KB = 72 B/row per
// iteration instead of 88.  p_new goes to the other buffer of a ping-pong pair because neighbouring CTAs still
// need the old p.  Element arithmetic and its order are those of the reference; only the kernel boundaries moved.
struct FusedArgs {
    const double* r;
    double* x;
    double* Ap;
    double* P0;
    double* P1;
    const double* r_halo_lo;
    const double* r_halo_hi;
    const double* p_halo_lo;
    const double* p_halo_hi;
    CgFusedState* st;
    const double* slot_rr;
    int world;
    double* rs_trace;
    // peer-memory mode (Context::p2p): no NCCL inside the iteration
    int p2p;
    int rank;
    const unsigned long long* ll;                // local packed value/sequence words
    unsigned long long* const* peer_ll;          // every rank's packed words, mapped here
    const double* my_p_lo[2];                    // local p halo planes, indexed by ping-pong parity
    const double* my_p_hi[2];
    double* nb_lo_p_hi[2];                       // lower neighbour's p_hi halo planes (null on rank 0)
    double* nb_hi_p_lo[2];                       // upper neighbour's p_lo halo planes (null on the last rank)
};

template <int TX, int TY, bool VEC2>
__global__ void __launch_bounds__(TX* TY) stencil7_cg_fused_kernel(StencilGeom g, FusedArgs a, int k_lo, int k_hi, int kc,
                                                                    double* partials, int partial_base, unsigned* ticket,
                                                                    double* dot_out, int is_final) {
    CgFusedState* st = a.st;
    if (st->done) return;
    const int first = st->first;
    const int cnt = st->ka_count;
    const double* __restrict__ ps = (cnt & 1) ? a.P1 : a.P0;
    double* __restrict__ pd = (cnt & 1) ? a.P0 : a.P1;
    const double* __restrict__ r = a.r;
    double rs_new = 0.0, beta = 0.0, alpha = 0.0;
    if (!first) {
        rs_new = slot_sum(a.slot_rr, a.world);
        beta = __ddiv_rn(rs_new, st->rs_old);
        alpha = st->alpha;
    }
    const int i0 = (blockIdx.x * TX + threadIdx.x) * 2;
    const int j = blockIdx.y * TY + threadIdx.y;
    const int kb = k_lo + blockIdx.z * kc;
    const int ke = min(kb + kc, k_hi);
    const bool active = (i0 < g.nx) && (j < g.ny);
    const bool has1 = active && (i0 + 1 < g.nx);
    const int64_t sxy = (int64_t)g.nx * g.ny;
    const int64_t row = (int64_t)j * g.nx + i0;
    double dacc = 0.0;

    auto load2 = [&](const double* base, int64_t idx) -> double2 {
        if constexpr (VEC2) return *reinterpret_cast<const double2*>(base + idx);
        else return make_double2(base[idx], has1 ? base[idx + 1] : 0.0);
    };
    auto store2 = [&](double* base, int64_t idx, double2 v) {
        if constexpr (VEC2) *reinterpret_cast<double2*>(base + idx) = v;
        else { base[idx] = v.x; if (has1) base[idx + 1] = v.y; }
    };
    auto plane_r = [&](int k) -> double2 {
        if (k < 0) return load2(a.r_halo_lo, row);
        if (k >= g.nz_loc) return load2(a.r_halo_hi, row);
        return load2(r, (int64_t)k * sxy + row);
    };
    auto plane_p = [&](int k) -> double2 {
        if (k < 0) return load2(a.p_halo_lo, row);
        if (k >= g.nz_loc) return load2(a.p_halo_hi, row);
        return load2(ps, (int64_t)k * sxy + row);
    };
    auto pnew = [&](double rv, double pv) -> double { return first ? pv : add_rn(rv, mul_rn(pv, beta)); };
    auto pnew2 = [&](double2 rv, double2 pv) -> double2 { return make_double2(pnew(rv.x, pv.x), pnew(rv.y, pv.y)); };

    if (active && kb < ke) {
        const bool j_shell = (j == 0 || j == g.ny - 1);
        const bool a_shell_i = (i0 == 0 || i0 == g.nx - 1);
        const bool b_shell_i = (i0 + 1 == g.nx - 1);
        const double2 zero2 = make_double2(0.0, 0.0);
        double2 pnA = zero2, pnB, pnC = zero2, poB;
        if (g.k_off + kb - 1 >= 0) pnA = pnew2(plane_r(kb - 1), plane_p(kb - 1));
        {
            const double2 rv = plane_r(kb);
            poB = plane_p(kb);
            pnB = pnew2(rv, poB);
        }
        double2 rN = zero2, pN = zero2;
        bool haveN = (g.k_off + kb + 1 < g.nz);
        if (haveN) { rN = plane_r(kb + 1); pN = plane_p(kb + 1); }
        for (int k = kb; k < ke; ++k) {
            const int kg = g.k_off + k;
            const int64_t idx = (int64_t)k * sxy + row;
            // plane k+2 (the "next" of the following iteration) is requested before anything of plane k is used
            double2 rNN = zero2, pNN = zero2;
            const bool haveNN = (k + 1 < ke) && (kg + 2 < g.nz);
            if (haveNN) { rNN = plane_r(k + 2); pNN = plane_p(k + 2); }
            double2 xv = zero2;
            if (!first) xv = load2(a.x, idx);
            if (haveN) pnC = pnew2(rN, pN);
            double2 out;
            if (j_shell || kg == 0 || kg == g.nz - 1) {
                out.x = shell_row(pnB.x);
                out.y = shell_row(pnB.y);
            } else {
                const double2 jm = pnew2(load2(r, idx - g.nx), load2(ps, idx - g.nx));
                const double2 jp = pnew2(load2(r, idx + g.nx), load2(ps, idx + g.nx));
                if (a_shell_i) {
                    out.x = shell_row(pnB.x);
                } else {
                    const double left = pnew(r[idx - 1], ps[idx - 1]);
                    out.x = tap7(g, pnA.x, jm.x, left, pnB.x, pnB.y, jp.x, pnC.x);
                }
                if (!has1) {
                    out.y = 0.0;
                } else if (b_shell_i) {
                    out.y = shell_row(pnB.y);
                } else {
                    const double right = pnew(r[idx + 2], ps[idx + 2]);
                    out.y = tap7(g, pnA.y, jm.y, pnB.x, pnB.y, right, jp.y, pnC.y);
                }
            }
            store2(a.Ap, idx, out);
            store2(pd, idx, pnB);
            if (!first) {
                xv.x = add_rn(xv.x, mul_rn(poB.x, alpha));  // x = x + p*alpha of the previous iteration (:60)
                xv.y = add_rn(xv.y, mul_rn(poB.y, alpha));
                store2(a.x, idx, xv);
            }
            dacc = add_rn(dacc, mul_rn(pnB.x, out.x));
            if (has1) dacc = add_rn(dacc, mul_rn(pnB.y, out.y));
            pnA = pnB;
            pnB = pnC;
            poB = pN;
            rN = rNN;
            pN = pNN;
            haveN = haveNN;
        }
    }
    const double bs = block_sum<TX * TY>(dacc);
    __shared__ bool is_last;
    const unsigned nb = total_blocks();
    if (linear_thread_id() == 0) {
        partials[partial_base + linear_block_id()] = bs;
        is_last = false;
        if (is_final) {
            __threadfence();
            const unsigned t = atomicInc(ticket, nb - 1);
            is_last = (t == nb - 1);
        }
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        const double total = partials_sum<TX * TY>(partials, partial_base + nb);
        if (linear_thread_id() == 0) {
            *dot_out = total;
            if (!first) {  // the p update of the previous iteration is now complete (:79-80)
                if (a.rs_trace && st->it < st->trace_cap) a.rs_trace[st->it] = rs_new;
                st->rs_old = rs_new;
                st->it = st->it + 1;
            }
            st->first = 0;
            st->ka_count = cnt + 1;
        }
    }
}

// ---------------------------------------------------------------------------------------------- KA, cp.async pipeline
// Used whenever nx is even and the vectors are 16-byte aligned.  ncu on the L1-based kernel above showed it
// latency-bound: 72 % of the stall samples were long-scoreboard waits on the eight neighbour loads each thread issues
// per plane (L1 hit rate 35 %), with occupancy capped at 31 % by the 96 registers the register prefetch costs.
// Here ALL global reads of r and p go through cp.async (LDGSTS) into a DEPTH-deep ring of shared-memory plane tiles,
// requested three planes ahead at no register cost; neighbours are read back from shared memory:
//   rt/pt[stage][jj][ii]: raw r / p at point (i_lo - 2 + ii, j_lo - 1 + jj) of one plane;
//   centre columns start at ii = 2 (16-byte aligned), ii = 1 and 2TX+2 are the left / right halo columns,
//   jj = 0 and TY+1 the halo rows (copied by the first / last thread row, halo columns by the first / last thread).
// p_new at a neighbour is re-formed from the two raw values (one multiply-add), exactly as the reference forms it.
// One __syncthreads per plane.
// LDGSTS moves a 32-byte global sector into shared memory in one piece only when the destination has the same offset
// inside its 32-byte segment as the source has inside its sector; otherwise each 16-byte half is fetched from L2 by a
// request of its own (measured on the plain apply at 512^3: L2->SM traffic 3.42 GB instead of 1.87 GB, 0.400 ms instead
// of 0.347 ms).  Tile rows keep their centre columns at tile column 2 (16 bytes into the row, row pitch a multiple of
// 32 bytes) while the grid columns they hold start on a sector boundary, so the ring of tiles starts 16 bytes past a
// 128-byte aligned base.  (The fused kernel had this alignment by accident of its static shared-memory size.)
constexpr int kRingShift = 2;  // doubles

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gmem));
}
// same, destination given as a shared-space byte address (saves the generic->shared conversion per copy)
__device__ __forceinline__ void cp_async16_s(unsigned sa, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async8_s(unsigned sa, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

template <int TX, int TY>
__global__ void __launch_bounds__(TX* TY) stencil7_cg_fused_smem_kernel(StencilGeom g, FusedArgs a, int k_lo, int k_hi, int kc,
                                                                         double* partials, int partial_base, unsigned* ticket,
                                                                         double* dot_out, int is_final) {
    constexpr int DEPTH = 4;       // ring stages: planes k-1 .. k+2 resident / in flight while plane k is computed
    constexpr int W = 2 * TX + 4;  // tile width in doubles
    extern __shared__ __align__(128) double smem_raw[];
    double* smem_dyn = smem_raw + kRingShift;
    double(*rt)[TY + 2][W] = reinterpret_cast<double(*)[TY + 2][W]>(smem_dyn);
    double(*pt)[TY + 2][W] = reinterpret_cast<double(*)[TY + 2][W]>(smem_dyn + (size_t)DEPTH * (TY + 2) * W);

    CgFusedState* st = a.st;
    if (st->done || st->peer_timeout) return;
    const int first = st->first;
    const int cnt = st->ka_count;
    const double* __restrict__ ps = (cnt & 1) ? a.P1 : a.P0;
    double* __restrict__ pd = (cnt & 1) ? a.P0 : a.P1;
    const double* __restrict__ r = a.r;
    double rs_new = 0.0, beta = 0.0, alpha = 0.0;
    const double* p_halo_lo = a.p_halo_lo;
    const double* p_halo_hi = a.p_halo_hi;
    double* nb_lo = nullptr;  // where this step's boundary planes of p_new go on the neighbours
    double* nb_hi = nullptr;
    if (a.p2p) {
        p_halo_lo = a.my_p_lo[cnt & 1];
        p_halo_hi = a.my_p_hi[cnt & 1];
        nb_lo = a.nb_lo_p_hi[(cnt + 1) & 1];
        nb_hi = a.nb_hi_p_lo[(cnt + 1) & 1];
    }
    if (!first) {
        // peer mode: every rank's r.r partial of the previous KB must have landed (this is also what orders the halo planes)
        rs_new = a.p2p ? peer_wait_sum(a.ll, a.world, kSlotRR, st->epoch + (unsigned long long)cnt, &st->peer_timeout)
                       : slot_sum(a.slot_rr, a.world);
        if (a.p2p && *reinterpret_cast<volatile int*>(&st->peer_timeout)) return;  // CTA-uniform: read after the wait's barrier
        beta = __ddiv_rn(rs_new, st->rs_old);
        alpha = st->alpha;
    }
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int i_lo = blockIdx.x * 2 * TX;
    const int j_lo = blockIdx.y * TY;
    const int i0 = i_lo + 2 * tx;
    const int j = j_lo + ty;
    const int kb = k_lo + blockIdx.z * kc;
    const int ke = min(kb + kc, k_hi);
    const bool active = (i0 < g.nx) && (j < g.ny);  // nx is even here: both points of the pair exist
    const int64_t sxy = (int64_t)g.nx * g.ny;
    const int64_t row = (int64_t)j * g.nx + i0;
    const bool do_jlo = active && ty == 0 && j_lo > 0;
    const bool do_jhi = active && ty == TY - 1 && (j + 1 < g.ny);
    const bool do_ilo = active && tx == 0 && i_lo > 0;
    const bool do_ihi = active && tx == TX - 1 && (i_lo + 2 * TX < g.nx);
    double dacc = 0.0;

    // first element of plane k of an array: the slab itself, or a halo plane outside it
    auto plane_ptr = [&](const double* arr, const double* lo, const double* hi, int k) -> const double* {
        if (k < 0) return lo;
        if (k >= g.nz_loc) return hi;
        return arr + (int64_t)k * sxy;
    };
    // global plane index valid?  (planes outside the global grid are never requested)
    auto issue_plane = [&](int k) {
        const int kg = g.k_off + k;
        if (kg >= 0 && kg < g.nz && active) {
            const int s = ((k % DEPTH) + DEPTH) % DEPTH;
            const double* rp = plane_ptr(r, a.r_halo_lo, a.r_halo_hi, k) + row;
            const double* pp = plane_ptr(ps, p_halo_lo, p_halo_hi, k) + row;
            cp_async16(&rt[s][ty + 1][2 * tx + 2], rp);
            cp_async16(&pt[s][ty + 1][2 * tx + 2], pp);
            if (do_jlo) {
                cp_async16(&rt[s][0][2 * tx + 2], rp - g.nx);
                cp_async16(&pt[s][0][2 * tx + 2], pp - g.nx);
            }
            if (do_jhi) {
                cp_async16(&rt[s][TY + 1][2 * tx + 2], rp + g.nx);
                cp_async16(&pt[s][TY + 1][2 * tx + 2], pp + g.nx);
            }
            if (do_ilo) {
                cp_async8(&rt[s][ty + 1][1], rp - 1);
                cp_async8(&pt[s][ty + 1][1], pp - 1);
            }
            if (do_ihi) {
                cp_async8(&rt[s][ty + 1][2 * TX + 2], rp + 2);
                cp_async8(&pt[s][ty + 1][2 * TX + 2], pp + 2);
            }
        }
        cp_async_commit();  // always commit, so that every thread counts the same number of groups
    };
    auto pnew = [&](double rv, double pv) -> double { return first ? pv : add_rn(rv, mul_rn(pv, beta)); };
    auto pnew2 = [&](double2 rv, double2 pv) -> double2 { return make_double2(pnew(rv.x, pv.x), pnew(rv.y, pv.y)); };
    auto centre_r = [&](int s) -> double2 { return *reinterpret_cast<const double2*>(&rt[s][ty + 1][2 * tx + 2]); };
    auto centre_p = [&](int s) -> double2 { return *reinterpret_cast<const double2*>(&pt[s][ty + 1][2 * tx + 2]); };
    auto stage_of = [&](int k) -> int { return ((k % DEPTH) + DEPTH) % DEPTH; };

    if (kb < ke) {
        const bool j_shell = (j == 0 || j == g.ny - 1);
        const bool a_shell_i = (i0 == 0 || i0 == g.nx - 1);
        const bool b_shell_i = (i0 + 1 == g.nx - 1);
        const double2 zero2 = make_double2(0.0, 0.0);
        // prologue: planes kb-1, kb, kb+1, kb+2 in flight (4 groups)
        issue_plane(kb - 1);
        issue_plane(kb);
        issue_plane(kb + 1);
        issue_plane(kb + 2);
        double2 xv = zero2;
        if (active && !first) xv = *reinterpret_cast<const double2*>(a.x + (int64_t)kb * sxy + row);
        cp_async_wait<2>();  // kb-1 and kb have landed
        __syncthreads();
        double2 pnA = zero2, pnB = zero2, pnC = zero2, poB = zero2;
        if (active) {
            if (g.k_off + kb - 1 >= 0) pnA = pnew2(centre_r(stage_of(kb - 1)), centre_p(stage_of(kb - 1)));
            poB = centre_p(stage_of(kb));
            pnB = pnew2(centre_r(stage_of(kb)), poB);
        }
        for (int k = kb; k < ke; ++k) {
            const int kg = g.k_off + k;
            const int64_t idx = (int64_t)k * sxy + row;
            cp_async_wait<1>();  // plane k+1 has landed (k+2 may still be in flight)
            __syncthreads();     // ... for every thread; also: everyone is done reading plane k-1's stage
            issue_plane(k + 3);  // into the stage plane k-1 occupied
            double2 xn = zero2;  // x of the next plane, requested one iteration early
            if (active && !first && k + 1 < ke) xn = *reinterpret_cast<const double2*>(a.x + idx + sxy);
            if (active) {
                const int s = stage_of(k), sn = stage_of(k + 1);
                double2 poC = zero2;
                if (kg + 1 < g.nz) {
                    poC = centre_p(sn);
                    pnC = pnew2(centre_r(sn), poC);
                }
                double2 out;
                if (j_shell || kg == 0 || kg == g.nz - 1) {
                    out.x = shell_row(pnB.x);
                    out.y = shell_row(pnB.y);
                } else {
                    const double2 jm = pnew2(*reinterpret_cast<const double2*>(&rt[s][ty][2 * tx + 2]),
                                             *reinterpret_cast<const double2*>(&pt[s][ty][2 * tx + 2]));
                    const double2 jp = pnew2(*reinterpret_cast<const double2*>(&rt[s][ty + 2][2 * tx + 2]),
                                             *reinterpret_cast<const double2*>(&pt[s][ty + 2][2 * tx + 2]));
                    if (a_shell_i) {
                        out.x = shell_row(pnB.x);
                    } else {
                        const double left = pnew(rt[s][ty + 1][2 * tx + 1], pt[s][ty + 1][2 * tx + 1]);
                        out.x = tap7(g, pnA.x, jm.x, left, pnB.x, pnB.y, jp.x, pnC.x);
                    }
                    if (b_shell_i) {
                        out.y = shell_row(pnB.y);
                    } else {
                        const double right = pnew(rt[s][ty + 1][2 * tx + 4], pt[s][ty + 1][2 * tx + 4]);
                        out.y = tap7(g, pnA.y, jm.y, pnB.x, pnB.y, right, jp.y, pnC.y);
                    }
                }
                *reinterpret_cast<double2*>(a.Ap + idx) = out;
                *reinterpret_cast<double2*>(pd + idx) = pnB;
                if (a.p2p) {  // boundary planes of p_new go straight into the neighbours' halo buffers over NVLink
                    if (k == 0 && nb_lo) *reinterpret_cast<double2*>(nb_lo + row) = pnB;
                    if (k == g.nz_loc - 1 && nb_hi) *reinterpret_cast<double2*>(nb_hi + row) = pnB;
                }
                dacc = add_rn(dacc, mul_rn(pnB.x, out.x));
                dacc = add_rn(dacc, mul_rn(pnB.y, out.y));
                if (!first) {
                    xv.x = add_rn(xv.x, mul_rn(poB.x, alpha));  // x = x + p*alpha of the previous iteration (:60)
                    xv.y = add_rn(xv.y, mul_rn(poB.y, alpha));
                    *reinterpret_cast<double2*>(a.x + idx) = xv;
                }
                pnA = pnB;
                pnB = pnC;
                poB = poC;
            }
            xv = xn;
        }
        cp_async_wait<0>();
    }
    const double bs = block_sum<TX * TY>(dacc);
    __shared__ bool is_last;
    const unsigned nb = total_blocks();
    if (a.p2p) {
        __threadfence_system();  // every thread orders its own peer stores at system scope ...
        __syncthreads();         // ... before thread 0 takes the ticket below
    }
    if (linear_thread_id() == 0) {
        partials[partial_base + linear_block_id()] = bs;
        is_last = false;
        if (is_final) {
            __threadfence();
            const unsigned t = atomicInc(ticket, nb - 1);
            is_last = (t == nb - 1);
        }
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        const double total = partials_sum<TX * TY>(partials, partial_base + nb);
        // system-scope release by the publishing threads themselves: the halo planes other CTAs stored (each behind its
        // own system fence and the gpu-scope ticket) are then formally ordered before the words the remote readers poll
        if (a.p2p) __threadfence_system();
        if (a.p2p && linear_thread_id() < 32)
            publish_partial_warp(a.peer_ll, a.world, a.rank, kSlotPAp, total, st->epoch + (unsigned long long)cnt + 1ull);
        if (linear_thread_id() == 0) {
            if (!a.p2p) *dot_out = total;
            if (!first) {  // the p update of the previous iteration is now complete (:79-80)
                if (a.rs_trace && st->it < st->trace_cap) a.rs_trace[st->it] = rs_new;
                st->rs_old = rs_new;
                st->it = st->it + 1;
            }
            st->first = 0;
            st->ka_count = cnt + 1;
        }
    }
}

// ---------------------------------------------------------------------------------------------- plain apply, cp.async
// y = A x (+ partial sums of x.y) with the same data path as KA: x planes stream through a 4-deep ring of
// shared-memory tiles filled by cp.async three planes ahead; k-1 / k / k+1 centre values live in registers.
// Used when nx is even and x, y are 16-byte aligned; the register-marching kernel above is the general fallback.
// PEER (distributed contexts in peer-memory mode): the WHOLE slab in one launch.  The CTAs whose chunk holds the slab's first
// / last plane are scheduled first; each stores its tile of that plane straight into the neighbour's halo buffer (NVLink),
// the last of them to finish releases the apply's sequence number into the neighbour's flag word, and each then waits on
// its LOCAL flag for the neighbour's plane before it marches.  No exchange kernel, no NCCL call, no second stream.
struct PeerApply {
    double* prev_dst = nullptr;            // previous rank's "hi" halo buffer of this apply's parity (null: no such rank)
    double* next_dst = nullptr;            // next rank's "lo" halo buffer
    unsigned long long* prev_flag = nullptr;
    unsigned long long* next_flag = nullptr;
    const unsigned long long* my_flags = nullptr;  // [0] set by the previous rank, [1] by the next rank
    unsigned* tickets = nullptr;           // [2] self-resetting arrival counters (lo side, hi side)
    unsigned long long seq = 0;
};
__device__ __forceinline__ void peer_flag_wait(const unsigned long long* flag, unsigned long long seq) {
    unsigned spins = 0;
    unsigned long long t0 = 0;
    while (ld_acquire_sys(flag) < seq) {
        if ((++spins & 0xfffu) == 0) {
            const unsigned long long now = global_timer_ns();
            if (t0 == 0) t0 = now;
            if (now - t0 > kPeerWaitNs) asm volatile("trap;");  // the neighbour died or fell out of step: fail loudly
        }
    }
}

template <int TX, int TY, bool DOT, bool PEER = false>
__global__ void __launch_bounds__(TX* TY) stencil7_smem_kernel(StencilGeom g, const double* __restrict__ x, double* __restrict__ y,
                                                                const double* __restrict__ halo_lo,
                                                                const double* __restrict__ halo_hi, int k_lo, int k_hi, int kc,
                                                                double* partials, int partial_base, unsigned* ticket,
                                                                double* dot_out, int dot_final, const int* skip_flag,
                                                                PeerApply pa = PeerApply()) {
    constexpr int DEPTH = 4;
    constexpr int W = 2 * TX + 4;
    constexpr int STAGE = (TY + 2) * W;  // doubles per ring stage
    constexpr int NT = TX * TY;
    constexpr int CPR = W / 2;               // 16-byte chunks per tile row: [left halo pair | TX centre pairs | right halo pair]
    constexpr int NCH = (TY + 2) * CPR;      // chunks per plane tile
    constexpr int Q = (NCH + NT - 1) / NT;   // chunks per thread
    extern __shared__ __align__(128) double smem_raw[];
    double* smem_dyn = smem_raw + kRingShift;  // see kRingShift: tile chunks must sit in shared memory as they sit in their sector
    if (skip_flag && *skip_flag) return;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int i_lo = blockIdx.x * 2 * TX;
    const int j_lo = blockIdx.y * TY;
    const int i0 = i_lo + 2 * tx;
    const int j = j_lo + ty;
    // PEER: blockIdx.z == 0 is a slice of push-only CTAs (one per xy tile): each stores its tile of the slab's first and last
    // plane into the neighbours, fences, and the last one to finish releases the sequence number into the neighbours' flag
    // words; they never wait.  The chunks follow with the two boundary chunks LAST (CTAs are dispatched in ascending block
    // order): by the time a boundary CTA spins on its local flag every push CTA of this launch has been dispatched, so a
    // waiting CTA can never keep a pushing one off the SMs, whatever the grid size.
    int zc = blockIdx.z;
    const int64_t sxy = (int64_t)g.nx * g.ny;
    const bool active = (i0 < g.nx) && (j < g.ny);
    if constexpr (PEER) {
        if (zc == 0) {
            const int64_t row = (int64_t)j * g.nx + i0;
            if (active) {
                if (pa.prev_dst) *reinterpret_cast<double2*>(pa.prev_dst + row) = *reinterpret_cast<const double2*>(x + row);
                if (pa.next_dst)
                    *reinterpret_cast<double2*>(pa.next_dst + row) = *reinterpret_cast<const double2*>(x + (int64_t)(g.nz_loc - 1) * sxy + row);
            }
            __threadfence_system();  // this thread's NVLink stores are performed ...
            __syncthreads();         // ... before thread 0 announces the tile
            if (tx == 0 && ty == 0) {
                const unsigned n_tiles = gridDim.x * gridDim.y;
                if (atomicInc(pa.tickets, n_tiles - 1) == n_tiles - 1) {
                    __threadfence_system();
                    if (pa.prev_flag) st_release_sys(pa.prev_flag, pa.seq);
                    if (pa.next_flag) st_release_sys(pa.next_flag, pa.seq);
                }
            }
            if constexpr (!DOT) return;
        }
        // chunks: the inner ones first, then the one with the first plane, then the one with the last plane
        const int nchunk = (int)gridDim.z - 1;
        const int q = zc - 1;
        zc = (zc == 0) ? nchunk /* no planes: kb >= ke below */ : (nchunk >= 2 ? (q < nchunk - 2 ? q + 1 : (q == nchunk - 2 ? 0 : nchunk - 1)) : q);
    }
    const int kb = k_lo + zc * kc;
    const int ke = min(kb + kc, k_hi);
    double dacc = 0.0;
    if constexpr (PEER) {
        if (kb < ke) {
            const bool need_lo = kb == 0 && g.k_off > 0, need_hi = ke == g.nz_loc && g.k_off + g.nz_loc < g.nz;
            if (need_lo || need_hi) {
                if (tx == 0 && ty == 0) {
                    if (need_lo) peer_flag_wait(pa.my_flags, pa.seq);
                    if (need_hi) peer_flag_wait(pa.my_flags + 1, pa.seq);
                }
                __syncthreads();
            }
        }
    }
    if (kb < ke) {
        // Planes kb-1 .. ke are read; those outside the global grid do not exist.  Plane k lives in ring stage
        // (k - kb + 1) mod DEPTH.  The steady-state loop is kept lean - this kernel was issue-bound (ncu: 76 % issue-active,
        // shared-memory pipe 70 %, ~160 instructions per plane and thread): the tile is copied as a flat list of 16-byte
        // chunks, Q per thread with addresses fixed before the loop (the halo columns come as the aligned pair that holds
        // them - same 32-byte sector), and the k-1 / k / k+1 registers rotate by renaming (three planes per trip).
        const int k_first = max(kb - 1, -g.k_off);
        const int k_last = min(ke, g.nz - 1 - g.k_off);
        const unsigned s0 = (unsigned)__cvta_generic_to_shared(smem_dyn);
        unsigned c_sm[Q];    // shared-memory byte address of the chunk in stage 0
        int64_t c_off[Q];    // element offset of the chunk inside a plane
        bool c_ok[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const int e = (ty * TX + tx) + q * NT;
            const int jj = e / CPR, c = e - jj * CPR;
            const int gj = j_lo - 1 + jj, gi = i_lo - 2 + 2 * c;
            const bool corner = (jj == 0 || jj == TY + 1) && (c == 0 || c == CPR - 1);
            c_ok[q] = e < NCH && !corner && gj >= 0 && gj < g.ny && gi >= 0 && gi < g.nx;
            c_sm[q] = s0 + 8u * (unsigned)(jj * W + 2 * c);
            c_off[q] = (int64_t)gj * g.nx + gi;
        }
        auto plane_of = [&](int k) -> const double* { return k < 0 ? halo_lo : (k >= g.nz_loc ? halo_hi : x + (int64_t)k * sxy); };
        auto issue_plane = [&](int k, const double* plane) {
            if (k >= k_first && k <= k_last) {
                const unsigned so = (unsigned)((k - kb + 1) & (DEPTH - 1)) * (unsigned)(STAGE * 8);
#pragma unroll
                for (int q = 0; q < Q; ++q)
                    if (c_ok[q]) cp_async16_s(c_sm[q] + so, plane + c_off[q]);
            }
            cp_async_commit();
        };
        const double* tile = smem_dyn + (ty + 1) * W + 2 * tx + 2;  // this thread's centre slot in stage 0
        auto ld2 = [](const double* q) -> double2 { return *reinterpret_cast<const double2*>(q); };
        const bool shA = (j == 0 || j == g.ny - 1) || (i0 == 0 || i0 == g.nx - 1);
        const bool shB = (j == 0 || j == g.ny - 1) || (i0 + 1 == g.nx - 1);
        issue_plane(kb - 1, plane_of(kb - 1));
        issue_plane(kb, plane_of(kb));
        issue_plane(kb + 1, plane_of(kb + 1));
        const double* plane = plane_of(kb + 2);
        issue_plane(kb + 2, plane);
        cp_async_wait<2>();
        __syncthreads();
        double2 v0 = make_double2(0.0, 0.0), v1 = v0, v2 = v0;
        if (active) {
            if (kb - 1 >= k_first) v0 = ld2(tile);  // stage 0
            v1 = ld2(tile + STAGE);                 // stage 1
        }
        double* yp = y + (int64_t)kb * sxy + (int64_t)j * g.nx + i0;
        unsigned m = 1;  // ring position of plane k
        int k = kb;
        // one plane: vA = k-1, vB = k (both in registers), vC receives k+1
        auto step = [&](const double2& vA, const double2& vB, double2& vC) {
            cp_async_wait<1>();
            __syncthreads();
            plane = (k + 3 == g.nz_loc) ? halo_hi : plane + sxy;
            issue_plane(k + 3, plane);
            if (active) {
                const double* st = tile + (m & (DEPTH - 1)) * STAGE;
                // plane k+1 may not exist (last global plane): its stage then holds stale data, and every row of this
                // plane takes the identity branch below
                vC = ld2(tile + ((m + 1) & (DEPTH - 1)) * STAGE);
                const double2 jm = ld2(st - W);
                const double2 jp = ld2(st + W);
                const double im = st[-1], ip = st[2];
                double2 out;
                out.x = tap7(g, vA.x, jm.x, im, vB.x, vB.y, jp.x, vC.x);
                out.y = tap7(g, vA.y, jm.y, vB.x, vB.y, ip, jp.y, vC.y);
                const int kg = g.k_off + k;
                const bool ksh = (kg == 0) || (kg == g.nz - 1);
                if (shA || shB || ksh) {  // Dirichlet shell: identity rows (the taps computed above are discarded)
                    if (shA || ksh) out.x = shell_row(vB.x);
                    if (shB || ksh) out.y = shell_row(vB.y);
                }
                *reinterpret_cast<double2*>(yp) = out;
                if constexpr (DOT) {
                    dacc = add_rn(dacc, mul_rn(vB.x, out.x));
                    dacc = add_rn(dacc, mul_rn(vB.y, out.y));
                }
            }
            ++k;
            ++m;
            yp += sxy;
        };
        while (k + 3 <= ke) {
            step(v0, v1, v2);
            step(v1, v2, v0);
            step(v2, v0, v1);
        }
        if (k < ke) {
            step(v0, v1, v2);
            if (k < ke) step(v1, v2, v0);
        }
        cp_async_wait<0>();
    }
    if constexpr (DOT) {
        const double bs = block_sum<TX * TY>(dacc);
        __shared__ bool is_last;
        const unsigned nb = total_blocks();
        if (linear_thread_id() == 0) {
            partials[partial_base + linear_block_id()] = bs;
            is_last = false;
            if (dot_final) {
                __threadfence();
                const unsigned t = atomicInc(ticket, nb - 1);
                is_last = (t == nb - 1);
            }
        }
        __syncthreads();
        if (is_last) {
            __threadfence();
            const double total = partials_sum<TX * TY>(partials, partial_base + nb);
            if (linear_thread_id() == 0) *dot_out = total;
        }
    }
}

// ---------------------------------------------------------------------------------------------- plain apply, TMA bulk copies
// Same ring of shared-memory plane tiles as stencil7_smem_kernel, but filled by the TMA unit: one elected thread arms an
// mbarrier with the byte count of a plane tile and issues one cp.async.bulk (global -> shared, completion on the mbarrier)
// per tile ROW (2TX+4 doubles incl. the left/right halo columns, clipped at the domain edge); all threads wait on the
// mbarrier's phase.  No thread spends issue slots or registers on address generation for loads.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "ATK_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra ATK_DONE_%=;\n"
        "bra ATK_WAIT_%=;\n"
        "ATK_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

template <int TX, int TY, bool DOT>
__global__ void __launch_bounds__(TX* TY) stencil7_tma_kernel(StencilGeom g, const double* __restrict__ x, double* __restrict__ y,
                                                               const double* __restrict__ halo_lo,
                                                               const double* __restrict__ halo_hi, int k_lo, int k_hi, int kc,
                                                               double* partials, int partial_base, unsigned* ticket,
                                                               double* dot_out, int dot_final, const int* skip_flag) {
    constexpr int DEPTH = 4;
    constexpr int W = 2 * TX + 4;
    extern __shared__ __align__(128) double smem_raw[];
    double(*xt)[TY + 2][W] = reinterpret_cast<double(*)[TY + 2][W]>(smem_raw + kRingShift);
    __shared__ __align__(8) unsigned long long full[DEPTH];
    if (skip_flag && *skip_flag) return;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int i_lo = blockIdx.x * 2 * TX;
    const int j_lo = blockIdx.y * TY;
    const int i0 = i_lo + 2 * tx;
    const int j = j_lo + ty;
    const int kb = k_lo + blockIdx.z * kc;
    const int ke = min(kb + kc, k_hi);
    const bool active = (i0 < g.nx) && (j < g.ny);
    const int64_t sxy = (int64_t)g.nx * g.ny;
    const int64_t row = (int64_t)j * g.nx + i0;
    const bool leader = (tx == 0 && ty == 0);
    double dacc = 0.0;
    if (leader) {
#pragma unroll
        for (int s = 0; s < DEPTH; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // tile geometry shared by every plane: columns [c_lo, c_hi) of the grid land at tile columns c_lo - (i_lo - 2) ...
    const int c_lo = max(i_lo - 2, 0), c_hi = min(i_lo + 2 * TX + 2, g.nx);
    const int jr_lo = max(j_lo - 1, 0), jr_hi = min(j_lo + TY + 1, g.ny);  // grid rows [jr_lo, jr_hi)
    const unsigned row_bytes = (unsigned)(c_hi - c_lo) * 8u;
    auto stage_of = [&](int k) -> int { return ((k % DEPTH) + DEPTH) % DEPTH; };
    auto plane_exists = [&](int k) -> bool { const int kg = g.k_off + k; return kg >= 0 && kg < g.nz; };
    auto issue_plane = [&](int k) {  // leader only
        if (!plane_exists(k)) return;
        const int s = stage_of(k);
        const double* src = (k < 0 ? halo_lo : (k >= g.nz_loc ? halo_hi : x + (int64_t)k * sxy));
        mbar_expect_tx(&full[s], row_bytes * (unsigned)(jr_hi - jr_lo));
        for (int jr = jr_lo; jr < jr_hi; ++jr)
            tma_bulk_g2s(&xt[s][jr - (j_lo - 1)][c_lo - (i_lo - 2)], src + (int64_t)jr * g.nx + c_lo, row_bytes, &full[s]);
    };
    unsigned phase_bits = 0;  // bit s = parity to wait for on stage s
    auto wait_plane = [&](int k) {
        if (!plane_exists(k)) return;
        const int s = stage_of(k);
        mbar_wait(&full[s], (phase_bits >> s) & 1u);
        phase_bits ^= 1u << s;
    };
    auto centre = [&](int s) -> double2 { return *reinterpret_cast<const double2*>(&xt[s][ty + 1][2 * tx + 2]); };
    if (kb < ke) {
        const bool j_shell = (j == 0 || j == g.ny - 1);
        const bool a_shell_i = (i0 == 0 || i0 == g.nx - 1);
        const bool b_shell_i = (i0 + 1 == g.nx - 1);
        const double2 zero2 = make_double2(0.0, 0.0);
        if (leader) {
            issue_plane(kb - 1);
            issue_plane(kb);
            issue_plane(kb + 1);
            issue_plane(kb + 2);
        }
        wait_plane(kb - 1);
        wait_plane(kb);
        double2 vA = zero2, vB = zero2, vC = zero2;
        if (active) {
            if (plane_exists(kb - 1)) vA = centre(stage_of(kb - 1));
            vB = centre(stage_of(kb));
        }
        for (int k = kb; k < ke; ++k) {
            const int kg = g.k_off + k;
            const int64_t idx = (int64_t)k * sxy + row;
            wait_plane(k + 1);
            __syncthreads();  // everyone is done reading plane k-1's stage: it may be refilled
            if (leader) issue_plane(k + 3);
            if (active) {
                const int s = stage_of(k);
                if (kg + 1 < g.nz) vC = centre(stage_of(k + 1));
                double2 out;
                if (j_shell || kg == 0 || kg == g.nz - 1) {
                    out.x = shell_row(vB.x);
                    out.y = shell_row(vB.y);
                } else {
                    const double2 jm = *reinterpret_cast<const double2*>(&xt[s][ty][2 * tx + 2]);
                    const double2 jp = *reinterpret_cast<const double2*>(&xt[s][ty + 2][2 * tx + 2]);
                    out.x = a_shell_i ? shell_row(vB.x) : tap7(g, vA.x, jm.x, xt[s][ty + 1][2 * tx + 1], vB.x, vB.y, jp.x, vC.x);
                    out.y = b_shell_i ? shell_row(vB.y) : tap7(g, vA.y, jm.y, vB.x, vB.y, xt[s][ty + 1][2 * tx + 4], jp.y, vC.y);
                }
                *reinterpret_cast<double2*>(y + idx) = out;
                if constexpr (DOT) {
                    dacc = add_rn(dacc, mul_rn(vB.x, out.x));
                    dacc = add_rn(dacc, mul_rn(vB.y, out.y));
                }
                vA = vB;
                vB = vC;
            }
        }
        // planes k+1.. that were requested beyond the chunk must have landed before the CTA may exit
        wait_plane(ke + 1);
        wait_plane(ke + 2);
    }
    if constexpr (DOT) {
        const double bs = block_sum<TX * TY>(dacc);
        __shared__ bool is_last;
        const unsigned nb = total_blocks();
        if (linear_thread_id() == 0) {
            partials[partial_base + linear_block_id()] = bs;
            is_last = false;
            if (dot_final) {
                __threadfence();
                const unsigned t = atomicInc(ticket, nb - 1);
                is_last = (t == nb - 1);
            }
        }
        __syncthreads();
        if (is_last) {
            __threadfence();
            const double total = partials_sum<TX * TY>(partials, partial_base + nb);
            if (linear_thread_id() == 0) *dot_out = total;
        }
    }
}

// ---------------------------------------------------------------------------------------------- KA through tensor-map TMA
// The same fused step as stencil7_cg_fused_smem_kernel with the r and p plane tiles fetched by the TMA unit: 3-D tensor maps
// over (nx, ny, nz) with a (2TX+4) x (TY+2) x 1 box, one cp.async.bulk.tensor.3d per array and plane issued by ONE elected
// thread at coordinates (i_lo-2, j_lo-1, k) - columns, rows and planes outside the grid are zero-filled by the hardware -
// completion on an mbarrier with expect_tx; nobody spends issue slots or registers on load addresses.  Single-GPU
// contexts (ATK_FUSED_TMA=1); arithmetic identical to the LDGSTS kernel, so the results are bit-identical.
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tm, int c0, int c1, int c2, unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

template <int TX, int TY>
__global__ void __launch_bounds__(TX* TY) stencil7_cg_fused_tma_kernel(const __grid_constant__ CUtensorMap tm_r,
                                                                        const __grid_constant__ CUtensorMap tm_p0,
                                                                        const __grid_constant__ CUtensorMap tm_p1, StencilGeom g,
                                                                        FusedArgs a, int k_lo, int k_hi, int kc, double* partials,
                                                                        int partial_base, unsigned* ticket, double* dot_out,
                                                                        int is_final) {
    constexpr int DEPTH = 4;
    constexpr int W = 2 * TX + 4;
    constexpr int TILE = (TY + 2) * W;                  // doubles per plane tile
    constexpr int STAGE = (TILE * 8 + 127) / 128 * 16;  // doubles per stage: TMA destinations are 128-byte aligned
    extern __shared__ __align__(128) double smem_raw[];
    double* ring_r = smem_raw;
    double* ring_p = smem_raw + DEPTH * STAGE;
    __shared__ __align__(8) unsigned long long full[DEPTH];

    CgFusedState* st = a.st;
    if (st->done) return;
    const int first = st->first;
    const int cnt = st->ka_count;
    const CUtensorMap* tm_p = (cnt & 1) ? &tm_p1 : &tm_p0;
    double* __restrict__ pd = (cnt & 1) ? a.P0 : a.P1;
    double rs_new = 0.0, beta = 0.0, alpha = 0.0;
    if (!first) {
        rs_new = slot_sum(a.slot_rr, a.world);
        beta = __ddiv_rn(rs_new, st->rs_old);
        alpha = st->alpha;
    }
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int i_lo = blockIdx.x * 2 * TX;
    const int j_lo = blockIdx.y * TY;
    const int i0 = i_lo + 2 * tx;
    const int j = j_lo + ty;
    const int kb = k_lo + blockIdx.z * kc;
    const int ke = min(kb + kc, k_hi);
    const bool active = (i0 < g.nx) && (j < g.ny);
    const int64_t sxy = (int64_t)g.nx * g.ny;
    const int64_t row = (int64_t)j * g.nx + i0;
    const bool leader = (tx == 0 && ty == 0);
    double dacc = 0.0;
    if (leader) {
#pragma unroll
        for (int s = 0; s < DEPTH; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto stage_of = [&](int k) -> int { return (k - kb + 1) & (DEPTH - 1); };
    auto issue_plane = [&](int k) {  // leader only; planes kb-1 .. ke (outside the grid: zero fill)
        if (k > ke) return;
        const int s = stage_of(k);
        mbar_expect_tx(&full[s], 2u * (unsigned)(TILE * 8));
        tma_load_3d(ring_r + s * STAGE, &tm_r, i_lo - 2, j_lo - 1, k, &full[s]);
        tma_load_3d(ring_p + s * STAGE, tm_p, i_lo - 2, j_lo - 1, k, &full[s]);
    };
    unsigned phase_bits = 0;
    auto wait_plane = [&](int k) {
        if (k > ke) return;
        const int s = stage_of(k);
        mbar_wait(&full[s], (phase_bits >> s) & 1u);
        phase_bits ^= 1u << s;
    };
    auto pnew = [&](double rv, double pv) -> double { return first ? pv : add_rn(rv, mul_rn(pv, beta)); };
    auto pnew2 = [&](double2 rv, double2 pv) -> double2 { return make_double2(pnew(rv.x, pv.x), pnew(rv.y, pv.y)); };
    auto ld2s = [](const double* q) -> double2 { return *reinterpret_cast<const double2*>(q); };
    const int centre = (ty + 1) * W + 2 * tx + 2;

    if (kb < ke) {
        const bool j_shell = (j == 0 || j == g.ny - 1);
        const bool a_shell_i = (i0 == 0 || i0 == g.nx - 1);
        const bool b_shell_i = (i0 + 1 == g.nx - 1);
        const double2 zero2 = make_double2(0.0, 0.0);
        if (leader) {
            issue_plane(kb - 1);
            issue_plane(kb);
            issue_plane(kb + 1);
            issue_plane(kb + 2);
        }
        double2 xv = zero2;
        if (active && !first) xv = *reinterpret_cast<const double2*>(a.x + (int64_t)kb * sxy + row);
        wait_plane(kb - 1);
        wait_plane(kb);
        double2 pnA = zero2, pnB = zero2, pnC = zero2, poB = zero2;
        if (active) {
            if (g.k_off + kb - 1 >= 0) pnA = pnew2(ld2s(ring_r + centre), ld2s(ring_p + centre));
            poB = ld2s(ring_p + STAGE + centre);
            pnB = pnew2(ld2s(ring_r + STAGE + centre), poB);
        }
        for (int k = kb; k < ke; ++k) {
            const int kg = g.k_off + k;
            const int64_t idx = (int64_t)k * sxy + row;
            wait_plane(k + 1);
            __syncthreads();  // everyone is done reading plane k-1's stage: it may be refilled
            if (leader) issue_plane(k + 3);
            double2 xn = zero2;
            if (active && !first && k + 1 < ke) xn = *reinterpret_cast<const double2*>(a.x + idx + sxy);
            if (active) {
                const int so = stage_of(k) * STAGE, sn = stage_of(k + 1) * STAGE;
                const double* rs = ring_r + so + centre;
                const double* pss = ring_p + so + centre;
                double2 poC = zero2;
                if (kg + 1 < g.nz) {
                    poC = ld2s(ring_p + sn + centre);
                    pnC = pnew2(ld2s(ring_r + sn + centre), poC);
                }
                double2 out;
                if (j_shell || kg == 0 || kg == g.nz - 1) {
                    out.x = shell_row(pnB.x);
                    out.y = shell_row(pnB.y);
                } else {
                    const double2 jm = pnew2(ld2s(rs - W), ld2s(pss - W));
                    const double2 jp = pnew2(ld2s(rs + W), ld2s(pss + W));
                    if (a_shell_i) {
                        out.x = shell_row(pnB.x);
                    } else {
                        const double left = pnew(rs[-1], pss[-1]);
                        out.x = tap7(g, pnA.x, jm.x, left, pnB.x, pnB.y, jp.x, pnC.x);
                    }
                    if (b_shell_i) {
                        out.y = shell_row(pnB.y);
                    } else {
                        const double right = pnew(rs[2], pss[2]);
                        out.y = tap7(g, pnA.y, jm.y, pnB.x, pnB.y, right, jp.y, pnC.y);
                    }
                }
                *reinterpret_cast<double2*>(a.Ap + idx) = out;
                *reinterpret_cast<double2*>(pd + idx) = pnB;
                dacc = add_rn(dacc, mul_rn(pnB.x, out.x));
                dacc = add_rn(dacc, mul_rn(pnB.y, out.y));
                if (!first) {
                    xv.x = add_rn(xv.x, mul_rn(poB.x, alpha));  // x = x + p*alpha of the previous iteration (:60)
                    xv.y = add_rn(xv.y, mul_rn(poB.y, alpha));
                    *reinterpret_cast<double2*>(a.x + idx) = xv;
                }
                pnA = pnB;
                pnB = pnC;
                poB = poC;
            }
            xv = xn;
        }
        // planes requested beyond the last one computed must have landed before the CTA may exit
        // (wait_plane(k+1) covered planes up to ke; nothing beyond ke was issued)
    }
    const double bs = block_sum<TX * TY>(dacc);
    __shared__ bool is_last;
    const unsigned nb = total_blocks();
    if (linear_thread_id() == 0) {
        partials[partial_base + linear_block_id()] = bs;
        is_last = false;
        if (is_final) {
            __threadfence();
            const unsigned t = atomicInc(ticket, nb - 1);
            is_last = (t == nb - 1);
        }
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        const double total = partials_sum<TX * TY>(partials, partial_base + nb);
        if (linear_thread_id() == 0) {
            *dot_out = total;
            if (!first) {
                if (a.rs_trace && st->it < st->trace_cap) a.rs_trace[st->it] = rs_new;
                st->rs_old = rs_new;
                st->it = st->it + 1;
            }
            st->first = 0;
            st->ka_count = cnt + 1;
        }
    }
}

// packs the boundary planes of r and of the current p (selected on the device by the ping-pong parity) for the
// halo exchange: send = [r_first | p_first | r_last | p_last]
__global__ void __launch_bounds__(256) pack_halo_kernel(const CgFusedState* st, const double* __restrict__ r, const double* P0,
                                                        const double* P1, int64_t plane, int64_t last_off,
                                                        double* __restrict__ send) {
    const double* ps = (st->ka_count & 1) ? P1 : P0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < plane; e += stride) {
        send[e] = r[e];
        send[plane + e] = ps[e];
        send[2 * plane + e] = r[last_off + e];
        send[3 * plane + e] = ps[last_off + e];
    }
}

// ---------------------------------------------------------------------------------------------- on-chip CG loop
// Small systems (up to ~1 M rows on one GPU) are bound by launch and reduction latency, not by HBM: two graph nodes per
// iteration cost 17-18 us at 32^3 - 64^3.  Here the WHOLE loop of ConjugateGradient.hpp:49-83 is one cooperative launch:
// every thread keeps its rows of x, r and p in registers for the entire solve, p is mirrored in a ping-pong pair of global
// buffers (L2-resident at these sizes) for the six neighbours, and the two dot products are grid-wide sums - per-CTA
// partial, software grid barrier, every CTA adds all partials in the same fixed order, so every thread derives identical
// alpha, beta and exit decisions.  Two grid barriers per iteration (p.Ap, r.r).  (Publishing the partials as tagged words
// that every CTA polls - the peer-memory protocol - was tried instead of barrier + read: faster with 2 CTAs, 1.5-2x slower
// from 64 CTAs up, the polls being quadratic in the CTA count.)
// Per-element arithmetic and its order are the reference's (same tap order, two roundings per multiply-add).
struct OnchipArgs {
    StencilGeom g;
    CgFusedState* st;
    double* x;
    double* r;
    double* p;       // user p: the buffer that holds p on entry and on exit
    double* p_alt;   // second buffer of the ping-pong pair
    double* part_a;  // [gridDim.x] partials of p.Ap
    double* part_b;  // [gridDim.x] partials of r.r
    unsigned* bar;   // zero at launch
    int max_iter;
    double* pAp_trace;
    double* rs_trace;
    int64_t n;
};
constexpr int kOnchipBlock = 256;
constexpr int kOnchipRows = 8;

__global__ void __launch_bounds__(kOnchipBlock) cg_onchip_kernel(OnchipArgs a) {
    CgFusedState* st = a.st;
    if (st->done) return;  // zero right-hand side: decided by the begin kernel, same for every CTA
    const StencilGeom g = a.g;
    const int64_t sxy = (int64_t)g.nx * g.ny;
    const int64_t c0 = (int64_t)blockIdx.x * kOnchipBlock * kOnchipRows;
    const unsigned G = gridDim.x;
    unsigned target = 0;
    double xr[kOnchipRows], rr[kOnchipRows], pr[kOnchipRows], ap[kOnchipRows];
    bool in[kOnchipRows], shell[kOnchipRows];
#pragma unroll
    for (int q = 0; q < kOnchipRows; ++q) {
        const int64_t e = c0 + (int64_t)q * kOnchipBlock + threadIdx.x;
        in[q] = e < a.n;
        xr[q] = in[q] ? a.x[e] : 0.0;
        rr[q] = in[q] ? a.r[e] : 0.0;
        pr[q] = in[q] ? a.p[e] : 0.0;
        const int64_t ee = in[q] ? e : 0;
        const int k = (int)(ee / sxy);
        const int rem = (int)(ee - (int64_t)k * sxy);
        const int j = rem / g.nx, i = rem - j * g.nx;
        shell[q] = i == 0 || i == g.nx - 1 || j == 0 || j == g.ny - 1 || k == 0 || k == g.nz - 1;
    }
    // Mirrors for the neighbours: r in the caller's r array, p in a ping-pong pair.  Iteration i forms p_i = r_i + beta_i *
    // p_(i-1) for its own rows AND, from the mirrors of r_i and p_(i-1), for the six neighbours of each row (one
    // multiply-add each, the very expression the owner evaluates), so the new p needs no barrier of its own: what it is
    // built from was complete before the r.r barrier of the previous iteration.  Two grid barriers per iteration.
    double* pm_prev = a.p;     // holds p_(i-1) of every row (iteration 0: the incoming p itself)
    double* pm_cur = a.p_alt;  // receives this iteration's own p_i
    double rs_old = st->rs_old;
    double beta = 0.0;
    int it = 0;
    bool stopped = false;
    double last_pAp = 0.0;
    while (it < a.max_iter) {
        const bool first = it == 0;
        double acc = 0.0;
#pragma unroll
        for (int q = 0; q < kOnchipRows; ++q) {
            const int64_t e = c0 + (int64_t)q * kOnchipBlock + threadIdx.x;
            if (!first) {  // p = r + p*beta of the previous iteration (:79)
                pr[q] = add_rn(rr[q], mul_rn(pr[q], beta));
                if (in[q]) __stcg(pm_cur + e, pr[q]);
            }
            double v = 0.0;
            if (in[q]) {
                if (shell[q]) {
                    v = shell_row(pr[q]);
                } else {
                    auto nb = [&](int64_t o) -> double {
                        const double pv = __ldcg(pm_prev + e + o);
                        return first ? pv : add_rn(__ldcg(a.r + e + o), mul_rn(pv, beta));
                    };
                    v = tap7(g, nb(-sxy), nb(-(int64_t)g.nx), nb(-1), pr[q], nb(1), nb(g.nx), nb(sxy));  // Ap = A p (:50)
                }
            }
            ap[q] = v;
            acc = add_rn(acc, mul_rn(pr[q], v));
        }
        const double bs = block_sum<kOnchipBlock>(acc);
        if (threadIdx.x == 0) __stcg(a.part_a + blockIdx.x, bs);
        coop_grid_barrier(a.bar, target);
        const double pAp = partials_sum<kOnchipBlock>(a.part_a, G);  // :51
        last_pAp = pAp;
        if (blockIdx.x == 0 && threadIdx.x == 0 && a.pAp_trace && it < st->trace_cap) a.pAp_trace[it] = pAp;
        if (!first) {  // the mirror written above is complete: it is the "previous p" of the next iteration
            double* tmp = pm_prev;
            pm_prev = pm_cur;
            pm_cur = tmp;
        }
        if (fabs(pAp) < 1e-12) {  // :53-57 - x and r stay as they are; p keeps the value formed for this iteration
            stopped = true;
            break;
        }
        const double alpha = __ddiv_rn(rs_old, pAp);  // :58
        double acc2 = 0.0;
#pragma unroll
        for (int q = 0; q < kOnchipRows; ++q) {
            const int64_t e = c0 + (int64_t)q * kOnchipBlock + threadIdx.x;
            xr[q] = add_rn(xr[q], mul_rn(pr[q], alpha));  // :60
            rr[q] = sub_rn(rr[q], mul_rn(ap[q], alpha));  // :61
            if (in[q]) __stcg(a.r + e, rr[q]);
            acc2 = add_rn(acc2, mul_rn(rr[q], rr[q]));
        }
        const double bs2 = block_sum<kOnchipBlock>(acc2);
        if (threadIdx.x == 0) __stcg(a.part_b + blockIdx.x, bs2);
        coop_grid_barrier(a.bar, target);
        const double rs_new = partials_sum<kOnchipBlock>(a.part_b, G);  // :63
        beta = __ddiv_rn(rs_new, rs_old);
        if (blockIdx.x == 0 && threadIdx.x == 0 && a.rs_trace && it < st->trace_cap) a.rs_trace[it] = rs_new;
        rs_old = rs_new;  // :80
        ++it;
    }
#pragma unroll
    for (int q = 0; q < kOnchipRows; ++q) {
        const int64_t e = c0 + (int64_t)q * kOnchipBlock + threadIdx.x;
        if (in[q]) {
            // a completed iteration still owes its p update (:79); after a |pAp| stop p is already what it must be
            if (!stopped && it > 0) pr[q] = add_rn(rr[q], mul_rn(pr[q], beta));
            a.x[e] = xr[q];
            a.r[e] = rr[q];
            a.p[e] = pr[q];
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->it = it;
        st->rs_old = rs_old;
        st->last_pAp = last_pAp;
        st->stopped_on_pAp = stopped ? 1 : 0;
        st->first = 0;
        st->done = stopped ? 1 : 0;
    }
}
// ---------------------------------------------------------------------------------------------- persistent CG loop
// The whole of ConjugateGradient.hpp:36-84 as ONE cooperative launch per rank (any grid size, one GPU or a z-slab of a
// distributed run in peer-memory mode).  Per iteration the kernel runs the two passes of the fused formulation above -
// phase A = KA (p update, deferred x update, operator apply, p.Ap), phase B = KB (r update, r.r) - but without kernel
// boundaries, last-CTA elections or NCCL calls: a reduction is
//     per-CTA partial -> arrive counter -> CTA 0 adds the partials in CTA order -> stores the rank's sum into EVERY rank's
//     packed {value half | sequence} words (one lane per peer, NVLink stores; the local rank included) ->
//     every CTA of every rank polls its LOCAL words and adds the rank sums in rank order.
// That single wait is both the grid barrier and the cross-GPU all-reduce, and it orders the halo planes the neighbours
// stored during the phase (every CTA's thread 0 issues a system-scope fence between its CTA's stores and its arrival).
// Work is split statically: item = (xy tile, z chunk), item i belongs to CTA i mod gridDim.x in every iteration, so x is
// only ever touched by one thread; everything another CTA (or GPU) may have written since the previous phase is read
// through L2 (cp.async.cg / ld.global.cg) - L1 is not coherent inside one launch.
// The r.r / b.b start-up sums, the initial halo planes and the final pending x / p updates (KF) are part of the launch.
struct PersistArgs {
    StencilGeom g;
    CgFusedState* st;
    const double* b;
    double* x;
    double* r;
    double* Ap;
    double* P0;  // the caller's p: holds p on entry and on exit
    double* P1;
    int max_iter;
    double* pAp_trace;
    double* rs_trace;
    int trace_cap;
    int tiles_x, tiles_y, kc, n_chunks, n_items;
    double* partials;             // [gridDim.x]
    unsigned long long* arrive;   // zero at launch
    unsigned long long* ll;       // local packed words
    unsigned long long* const* peer_ll;  // every rank's packed words as mapped here (null when world == 1)
    int world, rank;
    unsigned long long epoch;
    const double* r_halo_lo;
    const double* r_halo_hi;
    const double* my_p_lo[2];
    const double* my_p_hi[2];
    double* nb_lo_r_hi;
    double* nb_hi_r_lo;
    double* nb_lo_p_hi[2];
    double* nb_hi_p_lo[2];
    unsigned long long* stamps;   // optional globaltimer stamps: [0] after the start-up sums, then one per reduction
    int stamp_cap;
};

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
    return *reinterpret_cast<const volatile unsigned long long*>(p);
}
__device__ __forceinline__ double2 ldcg2(const double* p) { return __ldcg(reinterpret_cast<const double2*>(p)); }

template <int TX, int TY>
__global__ void __launch_bounds__(TX* TY, 7) cg_persistent_kernel(PersistArgs a) {
    constexpr int DEPTH = 4;
    constexpr int W = 2 * TX + 4;
    constexpr int STAGE = (TY + 2) * W;
    constexpr int NT = TX * TY;
    constexpr int CPR = W / 2;
    constexpr int NCH = (TY + 2) * CPR;
    constexpr int Q = (NCH + NT - 1) / NT;
    extern __shared__ __align__(128) double smem_raw[];
    double* ring_r = smem_raw + kRingShift;  // see kRingShift
    double* ring_p = ring_r + DEPTH * STAGE;
    __shared__ double rank_vals[kMaxRanks];
    __shared__ int s_abort;

    const StencilGeom g = a.g;
    CgFusedState* st = a.st;
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TX + tx;
    const unsigned G = gridDim.x, bid = blockIdx.x;
    const int64_t sxy = (int64_t)g.nx * g.ny;
    const int64_t n2 = (sxy * g.nz_loc) >> 1;  // nx is even: the slab is a whole number of double2
    const int64_t plane2 = sxy >> 1;
    const bool dist = a.world > 1;
    unsigned long long phase = 0;
    int n_stamps = 0;
    bool aborted = false;
    if (tid == 0) s_abort = 0;

    auto stamp = [&]() {
#ifndef ATK_REDUCE_PROFILE
        if (a.stamps && bid == 0 && tid == 0 && n_stamps < a.stamp_cap) a.stamps[n_stamps] = global_timer_ns();
#endif
        ++n_stamps;
    };
    // grid-wide and cross-GPU sum of one value per thread; every thread of every rank returns the same bits
#ifdef ATK_REDUCE_PROFILE
    // profile build only (python -m algotoolkit_b200.build --profile): CTA 0 stamps the five stations of every reduction -
    // own arrival, all CTAs arrived, rank sum published, all ranks' words seen, resume - at stamps[1 + 5*(phase-1) + k]
    auto rstamp = [&](int k) {
        if (a.stamps && bid == 0 && tid == 0) {
            const long long idx = 1 + 5 * ((long long)phase - 1) + k;
            if (idx < a.stamp_cap) a.stamps[idx] = global_timer_ns();
        }
    };
#else
    auto rstamp = [&](int) {};
#endif
    auto reduce = [&](double v) -> double {
        const double bs = block_sum<NT>(v);
        ++phase;
        rstamp(0);
        const int slot = (int)(phase & 1ull);
        const unsigned long long seq = a.epoch + phase;
        if (tid == 0) {
            __stcg(a.partials + bid, bs);
            if (dist) __threadfence_system();  // this CTA's halo stores (all threads: ordered by the barrier in block_sum) first
            else __threadfence();
            atomicAdd(a.arrive, 1ull);
        }
        if (bid == 0) {
            if (tid == 0) {
                const unsigned long long want = phase * (unsigned long long)G;
                unsigned spins = 0;
                unsigned long long t0 = 0;
                while (ld_volatile_u64(a.arrive) < want) {
                    if ((++spins & 0xfffu) == 0) {
                        const unsigned long long now = global_timer_ns();
                        if (t0 == 0) t0 = now;
                        if (now - t0 > kPeerWaitNs || *reinterpret_cast<volatile int*>(&st->peer_timeout)) {
                            st->peer_timeout = 1;
                            break;
                        }
                    }
                }
                __threadfence();
            }
            __syncthreads();
            rstamp(1);
            const double total = partials_sum<NT>(a.partials, G);
            if (tid < 32 && tid < a.world) {
                if (dist) __threadfence_system();
                const unsigned long long bits = (unsigned long long)__double_as_longlong(total);
                const unsigned long long tag = ll_tag(seq) << 32;
                unsigned long long* dst = (tid == a.rank ? a.ll : a.peer_ll[tid]) + (size_t)(slot * kMaxRanks + a.rank) * 2;
                st_relaxed_sys(dst, (bits & 0xffffffffull) | tag);
                st_relaxed_sys(dst + 1, (bits >> 32) | tag);
            }
            rstamp(2);
        }
        // every CTA: wait for all ranks' words of this phase (local memory), lane q polls rank q
        if (tid < 32) {
            if (tid < a.world) {
                const unsigned long long* src = a.ll + (size_t)(slot * kMaxRanks + tid) * 2;
                const unsigned long long tag = ll_tag(seq);
                unsigned long long w0, w1;
                unsigned spins = 0;
                unsigned long long t0 = 0;
                bool ok = true;
                for (;;) {  // acquire loads: the halo planes the neighbours stored before they published are visible after it
                    w0 = ld_acquire_sys(src);
                    w1 = ld_acquire_sys(src + 1);
                    if ((w0 >> 32) == tag && (w1 >> 32) == tag) break;
                    // a word two or more phases AHEAD of this rank would mean a peer overwrote a value nobody here has read
                    ATK_DEVICE_ASSERT((w0 >> 32) == 0ull || (((unsigned)((w0 >> 32) - tag)) & 0x7fffffffu) == 0u ||
                                      (((unsigned)((w0 >> 32) - tag)) & 0x7fffffffu) >= 0x40000000u);
                    if ((++spins & 0xfffu) == 0) {
                        const unsigned long long now = global_timer_ns();
                        if (t0 == 0) t0 = now;
                        if (now - t0 > kPeerWaitNs || *reinterpret_cast<volatile int*>(&st->peer_timeout)) {
                            st->peer_timeout = 1;
                            ok = false;
                            break;
                        }
                    }
                }
                if (!ok) s_abort = 1;
                rank_vals[tid] = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
            }
        }
        __syncthreads();
        rstamp(3);
        double s = rank_vals[0];
        for (int q = 1; q < a.world; ++q) s = add_rn(s, rank_vals[q]);
        aborted = s_abort != 0;
        __syncthreads();  // rank_vals is rewritten by the next reduction
        rstamp(4);
        return s;
    };

    // ---- start-up: rs_old = r.r, ||b||^2 (:37-38); the boundary planes of r and p go to the neighbours' halo planes
    const int64_t t_lin = (int64_t)bid * NT + tid;
    const int64_t stride = (int64_t)G * NT;
    double rr0 = 0.0, bb0 = 0.0;
    {
        double r1 = 0.0, b1 = 0.0;
        for (int64_t i = t_lin; i < n2; i += stride) {
            const double2 rv = ldcg2(a.r + 2 * i), bv = ldcg2(a.b + 2 * i);
            rr0 = add_rn(rr0, mul_rn(rv.x, rv.x));
            r1 = add_rn(r1, mul_rn(rv.y, rv.y));
            bb0 = add_rn(bb0, mul_rn(bv.x, bv.x));
            b1 = add_rn(b1, mul_rn(bv.y, bv.y));
            if (dist) {
                if (i < plane2) {
                    if (a.nb_lo_r_hi) reinterpret_cast<double2*>(a.nb_lo_r_hi)[i] = rv;
                    if (a.nb_lo_p_hi[0]) reinterpret_cast<double2*>(a.nb_lo_p_hi[0])[i] = ldcg2(a.P0 + 2 * i);
                }
                if (i >= n2 - plane2) {
                    const int64_t h = i - (n2 - plane2);
                    if (a.nb_hi_r_lo) reinterpret_cast<double2*>(a.nb_hi_r_lo)[h] = rv;
                    if (a.nb_hi_p_lo[0]) reinterpret_cast<double2*>(a.nb_hi_p_lo[0])[h] = ldcg2(a.P0 + 2 * i);
                }
            }
        }
        rr0 = add_rn(rr0, r1);
        bb0 = add_rn(bb0, b1);
    }
    double rs_old = reduce(rr0);
    const double bb = reduce(bb0);
    stamp();
    const bool zero_rhs = sqrt(bb) < 1e-12;  // :38-42
    double alpha = 0.0, beta = 0.0, last_pAp = 0.0;
    int it = 0;
    bool stopped = false;
    const int n_tiles = a.tiles_x * a.tiles_y;
    const unsigned s_r = (unsigned)__cvta_generic_to_shared(ring_r);
    const unsigned s_p = (unsigned)__cvta_generic_to_shared(ring_p);

    if (!zero_rhs && !aborted) {
        for (; it < a.max_iter; ++it) {
            const bool first = it == 0;
            const double* __restrict__ ps = (it & 1) ? a.P1 : a.P0;
            double* __restrict__ pd = (it & 1) ? a.P0 : a.P1;
            const double* p_halo_lo = a.my_p_lo[it & 1];
            const double* p_halo_hi = a.my_p_hi[it & 1];
            double* nb_lo = a.nb_lo_p_hi[(it + 1) & 1];
            double* nb_hi = a.nb_hi_p_lo[(it + 1) & 1];
            auto pnew = [&](double rv, double pv) -> double { return first ? pv : add_rn(rv, mul_rn(pv, beta)); };
            auto pnew2 = [&](double2 rv, double2 pv) -> double2 { return make_double2(pnew(rv.x, pv.x), pnew(rv.y, pv.y)); };
            double dacc = 0.0;
            // ================================================================ phase A (KA)
            for (int item = (int)bid; item < a.n_items; item += (int)G) {
                const int crank = item / n_tiles, tile = item - crank * n_tiles;
                const int chunk = a.n_chunks - 1 - crank;  // upper chunks first: their neighbour stores are then acknowledged early
                const int tyi = tile / a.tiles_x, txi = tile - tyi * a.tiles_x;
                const int i_lo = txi * 2 * TX, j_lo = tyi * TY;
                const int i0 = i_lo + 2 * tx, j = j_lo + ty;
                const int kb = chunk * a.kc, ke = min(kb + a.kc, g.nz_loc);
                const bool active = (i0 < g.nx) && (j < g.ny);
                const int64_t row = (int64_t)j * g.nx + i0;
                const int k_first = max(kb - 1, -g.k_off);
                const int k_last = min(ke, g.nz - 1 - g.k_off);
                unsigned c_sm[Q];
                int c_off[Q];
                bool c_ok[Q];
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    const int e = tid + q * NT;
                    const int jj = e / CPR, c = e - jj * CPR;
                    const int gj = j_lo - 1 + jj, gi = i_lo - 2 + 2 * c;
                    const bool corner = (jj == 0 || jj == TY + 1) && (c == 0 || c == CPR - 1);
                    c_ok[q] = e < NCH && !corner && gj >= 0 && gj < g.ny && gi >= 0 && gi < g.nx;
                    c_sm[q] = 8u * (unsigned)(jj * W + 2 * c);
                    c_off[q] = gj * g.nx + gi;
                    ATK_DEVICE_ASSERT(!c_ok[q] || (c_off[q] >= 0 && (int64_t)c_off[q] + 1 < sxy && c_sm[q] + 16u <= (unsigned)(STAGE * 8)));
                }
                ATK_DEVICE_ASSERT(kb >= 0 && kb < ke && ke <= g.nz_loc && tyi < a.tiles_y && chunk >= 0);
                auto issue_plane = [&](int k) {
                    if (k >= k_first && k <= k_last) {
                        const unsigned so = (unsigned)((k - kb + 1) & (DEPTH - 1)) * (unsigned)(STAGE * 8);
                        const double* rp = k < 0 ? a.r_halo_lo : (k >= g.nz_loc ? a.r_halo_hi : a.r + (int64_t)k * sxy);
                        const double* pp = k < 0 ? p_halo_lo : (k >= g.nz_loc ? p_halo_hi : ps + (int64_t)k * sxy);
                        ATK_DEVICE_ASSERT(k >= -1 && k <= g.nz_loc && rp != nullptr && pp != nullptr);  // a halo plane is only read where a neighbour exists
#pragma unroll
                        for (int q = 0; q < Q; ++q)
                            if (c_ok[q]) {
                                cp_async16_s(s_r + so + c_sm[q], rp + c_off[q]);
                                cp_async16_s(s_p + so + c_sm[q], pp + c_off[q]);
                            }
                    }
                    cp_async_commit();
                };
                const int centre = (ty + 1) * W + 2 * tx + 2;  // this thread's pair inside a stage
                auto ld2s = [](const double* q) -> double2 { return *reinterpret_cast<const double2*>(q); };
                const bool j_shell = (j == 0 || j == g.ny - 1);
                const bool a_shell_i = (i0 == 0 || i0 == g.nx - 1);
                const bool b_shell_i = (i0 + 1 == g.nx - 1);
                const double2 zero2 = make_double2(0.0, 0.0);
                issue_plane(kb - 1);
                issue_plane(kb);
                issue_plane(kb + 1);
                issue_plane(kb + 2);
                double2 xv = zero2;
                if (active && !first) xv = ldcg2(a.x + (int64_t)kb * sxy + row);
                cp_async_wait<2>();
                __syncthreads();
                double2 pnA = zero2, pnB = zero2, pnC = zero2, poB = zero2;
                if (active) {
                    if (g.k_off + kb - 1 >= 0) pnA = pnew2(ld2s(ring_r + centre), ld2s(ring_p + centre));  // stage 0 = plane kb-1
                    poB = ld2s(ring_p + STAGE + centre);                                                     // stage 1 = plane kb
                    pnB = pnew2(ld2s(ring_r + STAGE + centre), poB);
                }
                for (int k = kb; k < ke; ++k) {
                    const int kg = g.k_off + k;
                    const int64_t idx = (int64_t)k * sxy + row;
                    cp_async_wait<1>();
                    __syncthreads();
                    issue_plane(k + 3);
                    double2 xn = zero2;
                    if (active && !first && k + 1 < ke) xn = ldcg2(a.x + idx + sxy);
                    if (active) {
                        const int so = ((k - kb + 1) & (DEPTH - 1)) * STAGE, sn = ((k - kb + 2) & (DEPTH - 1)) * STAGE;
                        const double* rs = ring_r + so + centre;
                        const double* pss = ring_p + so + centre;
                        double2 poC = zero2;
                        if (kg + 1 < g.nz) {
                            poC = ld2s(ring_p + sn + centre);
                            pnC = pnew2(ld2s(ring_r + sn + centre), poC);
                        }
                        double2 out;
                        if (j_shell || kg == 0 || kg == g.nz - 1) {
                            out.x = shell_row(pnB.x);
                            out.y = shell_row(pnB.y);
                        } else {
                            const double2 jm = pnew2(ld2s(rs - W), ld2s(pss - W));
                            const double2 jp = pnew2(ld2s(rs + W), ld2s(pss + W));
                            if (a_shell_i) {
                                out.x = shell_row(pnB.x);
                            } else {
                                const double left = pnew(rs[-1], pss[-1]);
                                out.x = tap7(g, pnA.x, jm.x, left, pnB.x, pnB.y, jp.x, pnC.x);
                            }
                            if (b_shell_i) {
                                out.y = shell_row(pnB.y);
                            } else {
                                const double right = pnew(rs[2], pss[2]);
                                out.y = tap7(g, pnA.y, jm.y, pnB.x, pnB.y, right, jp.y, pnC.y);
                            }
                        }
                        *reinterpret_cast<double2*>(a.Ap + idx) = out;
                        *reinterpret_cast<double2*>(pd + idx) = pnB;
                        if (dist) {  // boundary planes of p_new go straight into the neighbours' halo planes over NVLink
                            if (k == 0 && nb_lo) *reinterpret_cast<double2*>(nb_lo + row) = pnB;
                            if (k == g.nz_loc - 1 && nb_hi) *reinterpret_cast<double2*>(nb_hi + row) = pnB;
                        }
                        dacc = add_rn(dacc, mul_rn(pnB.x, out.x));
                        dacc = add_rn(dacc, mul_rn(pnB.y, out.y));
                        if (!first) {
                            xv.x = add_rn(xv.x, mul_rn(poB.x, alpha));  // x = x + p*alpha of the previous iteration (:60)
                            xv.y = add_rn(xv.y, mul_rn(poB.y, alpha));
                            *reinterpret_cast<double2*>(a.x + idx) = xv;
                        }
                        pnA = pnB;
                        pnB = pnC;
                        poB = poC;
                    }
                    xv = xn;
                }
                cp_async_wait<0>();
                __syncthreads();  // the ring is free for the next item
            }
            const double pAp = reduce(dacc);  // :51
            stamp();
            if (aborted) break;
            last_pAp = pAp;
            if (bid == 0 && tid == 0 && a.pAp_trace && it < a.trace_cap) a.pAp_trace[it] = pAp;
            if (fabs(pAp) < 1e-12) {  // :53-57
                stopped = true;
                break;
            }
            alpha = __ddiv_rn(rs_old, pAp);  // :58
            // ================================================================ phase B (KB): r = r - Ap*alpha ; r.r
            double acc0 = 0.0, acc1 = 0.0;
            {
                // distributed: start with the last plane, then the first, so that the stores into the neighbours' halo
                // planes have long been acknowledged when the fence before the arrival is reached
                const int64_t off = dist ? n2 - plane2 : 0;
                auto one = [&](int64_t i, const double2 rv0, const double2 av) {
                    int64_t e = i + off;
                    if (e >= n2) e -= n2;
                    double2 rv = rv0;
                    rv.x = sub_rn(rv.x, mul_rn(av.x, alpha));
                    rv.y = sub_rn(rv.y, mul_rn(av.y, alpha));
                    *reinterpret_cast<double2*>(a.r + 2 * e) = rv;
                    if (dist) {
                        if (e < plane2 && a.nb_lo_r_hi) reinterpret_cast<double2*>(a.nb_lo_r_hi)[e] = rv;
                        if (e >= n2 - plane2 && a.nb_hi_r_lo) reinterpret_cast<double2*>(a.nb_hi_r_lo)[e - (n2 - plane2)] = rv;
                    }
                    acc0 = add_rn(acc0, mul_rn(rv.x, rv.x));
                    acc1 = add_rn(acc1, mul_rn(rv.y, rv.y));
                };
                auto at = [&](int64_t i) -> int64_t {
                    int64_t e = i + off;
                    if (e >= n2) e -= n2;
                    return 2 * e;
                };
                int64_t i = t_lin;
                for (; i + 3 * stride < n2; i += 4 * stride) {
                    const double2 r0 = ldcg2(a.r + at(i)), a0 = ldcg2(a.Ap + at(i));
                    const double2 r1 = ldcg2(a.r + at(i + stride)), a1 = ldcg2(a.Ap + at(i + stride));
                    const double2 r2 = ldcg2(a.r + at(i + 2 * stride)), a2 = ldcg2(a.Ap + at(i + 2 * stride));
                    const double2 r3 = ldcg2(a.r + at(i + 3 * stride)), a3 = ldcg2(a.Ap + at(i + 3 * stride));
                    one(i, r0, a0);
                    one(i + stride, r1, a1);
                    one(i + 2 * stride, r2, a2);
                    one(i + 3 * stride, r3, a3);
                }
                for (; i < n2; i += stride) one(i, ldcg2(a.r + at(i)), ldcg2(a.Ap + at(i)));
            }
            const double rs_new = reduce(add_rn(acc0, acc1));  // :63
            stamp();
            if (aborted) break;
            beta = __ddiv_rn(rs_new, rs_old);
            if (bid == 0 && tid == 0 && a.rs_trace && it < a.trace_cap) a.rs_trace[it] = rs_new;
            rs_old = rs_new;  // :80
        }
    }
    // ---- KF: bring the caller's x and p to the reference's state
    if (!aborted && (it > 0 || stopped)) {
        const int ka = stopped ? it + 1 : it;  // KA passes executed: the current direction lives in P[ka & 1]
        const double* pc = (ka & 1) ? a.P1 : a.P0;
        if (stopped) {  // x and r are complete; p keeps the value formed for this iteration
            if (pc != a.P0)
                for (int64_t i = t_lin; i < n2; i += stride) *reinterpret_cast<double2*>(a.P0 + 2 * i) = ldcg2(pc + 2 * i);
        } else {  // the last iteration still owes x = x + p*alpha (:60) and p = r + p*beta (:79)
            for (int64_t i = t_lin; i < n2; i += stride) {
                const double2 pv = ldcg2(pc + 2 * i), rv = ldcg2(a.r + 2 * i);
                double2 xv = ldcg2(a.x + 2 * i);
                xv.x = add_rn(xv.x, mul_rn(pv.x, alpha));
                xv.y = add_rn(xv.y, mul_rn(pv.y, alpha));
                *reinterpret_cast<double2*>(a.x + 2 * i) = xv;
                *reinterpret_cast<double2*>(a.P0 + 2 * i) = make_double2(add_rn(rv.x, mul_rn(pv.x, beta)), add_rn(rv.y, mul_rn(pv.y, beta)));
            }
        }
    }
    if (bid == 0 && tid == 0) {
        st->it = it;
        st->rs_old = rs_old;
        st->b_norm2 = bb;
        st->alpha = alpha;
        st->last_pAp = last_pAp;
        st->stopped_on_pAp = stopped ? 1 : 0;
        st->zero_rhs = zero_rhs ? 1 : 0;
        st->done = (stopped || zero_rhs) ? 1 : 0;
        st->first = 0;
        st->ka_count = stopped ? it + 1 : it;
        st->epoch = a.epoch;
        if (a.stamps && a.stamp_cap > 0) a.stamps[a.stamp_cap] = (unsigned long long)min(n_stamps, a.stamp_cap);
    }
}

template <int TX, int TY>
constexpr size_t fused_smem_bytes() { return (size_t)2 * 4 * (TY + 2) * (2 * TX + 4) * sizeof(double) + kRingShift * sizeof(double); }
template <int TX, int TY>
void allow_fused_smem() {
    ATK_CUDA_CHECK(cudaFuncSetAttribute(stencil7_cg_fused_smem_kernel<TX, TY>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)fused_smem_bytes<TX, TY>()));
}

struct Stencil7 : public Operator {
    StencilGeom g;
    double* nccl_halo_lo = nullptr;  // receive buffers of the NCCL exchange
    double* nccl_halo_hi = nullptr;
    const double* halo_lo = nullptr;  // the halo planes the kernels of the current apply read
    const double* halo_hi = nullptr;
    PeerHalo ph;                      // peer-memory exchange of the plain apply (Context::p2p)
    int kc = 64;
    int tile = 4;  // 0: 32x8 threads (64x8 points), 1: 64x4 (128x4), 2: 32x16 (64x16), 3: 64x8 (128x8)

    // fused CG: halo buffers for r and p (two planes each way) and the packed send buffer
    double* fsend = nullptr;  // [4*plane]: r_first p_first r_last p_last
    double* frecv = nullptr;  // halo arena [6*plane]: r_lo p_lo[0] r_hi p_hi[0] p_lo[1] p_hi[1] (the last two: peer mode)
    double* nb_lo_arena = nullptr;  // the neighbours' halo arenas mapped through CUDA IPC (peer mode)
    double* nb_hi_arena = nullptr;
    int fkc = 64;
    int ftile = 4;
    bool fsmem = true;
    bool use_tma = false;  // plain apply through TMA bulk copies (experiment switch ATK_STENCIL_TMA)

    bool arena_exported = false;

    ~Stencil7() override {
        cudaFree(oc_part);
        cudaFree(ps_part);
        cudaFree(nccl_halo_lo);
        cudaFree(nccl_halo_hi);
        ph.release();
        cudaFree(fsend);
        // the neighbours may still have this rank's halo arena mapped: unmap theirs now, free ours with the context
        ctx->ipc_close(nb_lo_arena);
        ctx->ipc_close(nb_hi_arena);
        if (arena_exported) ctx->arena_give(frecv);  // recycled by the next operator of the same plane size
        else cudaFree(frecv);
    }

    // Planes per z-chunk: the configured length on large grids; on small ones shorter chunks, so that there are about as
    // many CTAs as four per SM - but never below 8 planes, the pipeline fill costs 3.
    int chunk_planes(int configured, int planes, int tiles_xy) const {
        const int want_ctas = ctx->sm_count * 4;
        int kcc = std::min(configured, planes);
        while (kcc > 8 && (int64_t)tiles_xy * ((planes + kcc - 1) / kcc) < want_ctas) kcc = (kcc + 1) / 2;
        return std::max(1, kcc);
    }

    bool fused_cg_capable() const override { return true; }
    // ---- on-chip CG loop (single GPU, every CTA co-resident)
    double* oc_part = nullptr;  // [2 * oc_blocks] partial sums, then the barrier counter
    int oc_blocks = -1;         // -1: not evaluated yet, 0: not eligible
    int cg_onchip_blocks() const override {
        auto* self = const_cast<Stencil7*>(this);
        if (self->oc_blocks >= 0) return self->oc_blocks;
        self->oc_blocks = 0;
        const char* env = std::getenv("ATK_CG_ONCHIP");
        if (ctx->world != 1 || (env && env[0] == '0')) return 0;
        const int64_t rows_per_cta = (int64_t)kOnchipBlock * kOnchipRows;
        const int64_t blocks = (n_rows_local + rows_per_cta - 1) / rows_per_cta;
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cg_onchip_kernel, kOnchipBlock, 0) != cudaSuccess) return 0;
        int coop = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device);
        if (!coop || blocks < 1 || blocks > (int64_t)per_sm * ctx->sm_count) return 0;
        if (cudaMalloc(&self->oc_part, sizeof(double) * (2 * (size_t)blocks + 2)) != cudaSuccess) return 0;
        self->oc_blocks = (int)blocks;
        return self->oc_blocks;
    }
    void cg_onchip(CgFusedState* st, double* x, double* r, double* p, double* p_alt, int max_iter, double* pAp_trace,
                   double* rs_trace) override {
        ATK_REQUIRE(oc_blocks > 0, ATK_ERR_INVALID_ARGUMENT, "on-chip CG not available for this operator");
        unsigned* bar = reinterpret_cast<unsigned*>(oc_part + 2 * (size_t)oc_blocks);
        ATK_CUDA_CHECK(cudaMemsetAsync(bar, 0, sizeof(unsigned), ctx->stream));
        OnchipArgs oa{g, st, x, r, p, p_alt, oc_part, oc_part + oc_blocks, bar, max_iter, pAp_trace, rs_trace, n_rows_local};
        void* kargs[] = {&oa};
        ATK_CUDA_CHECK(cudaLaunchCooperativeKernel((void*)cg_onchip_kernel, dim3((unsigned)oc_blocks), dim3(kOnchipBlock), kargs, 0,
                                                   ctx->stream));
        count_launch(ctx);
    }
    // ---- persistent CG loop (one cooperative launch per solve; single GPU, or a z-slab in peer-memory mode)
    double* ps_part = nullptr;  // [ps_grid] per-CTA partials, then the 64-bit arrival counter
    int ps_grid = -1;           // -1: not evaluated yet, 0: not eligible
    int ps_kc = 0, ps_chunks = 0;
    int cg_persistent_blocks() const override {
        auto* self = const_cast<Stencil7*>(this);
        if (self->ps_grid >= 0) return self->ps_grid;
        self->ps_grid = 0;
        if ((g.nx & 1) || !fsmem || n_rows_local <= 0 || g.nz_loc < 1) return 0;
        if (ctx->world > 1 && !(ctx->p2p && frecv)) return 0;
        constexpr int TX = 32, TY = 4;
        int coop = 0, per_sm = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device);
        if (!coop) return 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cg_persistent_kernel<TX, TY>, TX * TY, fused_smem_bytes<TX, TY>()) !=
                cudaSuccess || per_sm < 1)
            return 0;
        const int64_t g_max = (int64_t)per_sm * ctx->sm_count;
        const int64_t tiles = (int64_t)((g.nx + 2 * TX - 1) / (2 * TX)) * ((g.ny + TY - 1) / TY);
        // z-chunk length: the one that minimises the planes the busiest CTA has to stream (every chunk re-reads its two
        // neighbour planes), among chunkings that give every SM something to do
        int min_chunks = 1;
        if (const char* e = std::getenv("ATK_PERSIST_SPLIT")) min_chunks = std::max(1, std::atoi(e));
        int64_t best = -1;
        int best_chunks = 1;
        for (int chunks = min_chunks; chunks <= std::max(min_chunks, g.nz_loc); ++chunks) {
            const int kcc = (g.nz_loc + chunks - 1) / chunks;
            if (kcc < 8 && chunks > min_chunks) break;
            if ((g.nz_loc + kcc - 1) / kcc != chunks) continue;
            const int64_t items = tiles * chunks;
            if (items > 0x3fffffff) break;
            const int64_t gu = std::min(g_max, items);
            const int64_t load = ((items + gu - 1) / gu) * (kcc + 2);
            if (best < 0 || load < best) {
                best = load;
                best_chunks = chunks;
            }
        }
        if (const char* e = std::getenv("ATK_PERSIST_KC")) {
            const int kcc = std::max(1, std::min(std::atoi(e), g.nz_loc));
            best_chunks = (g.nz_loc + kcc - 1) / kcc;
        }
        self->ps_chunks = best_chunks;
        self->ps_kc = (g.nz_loc + best_chunks - 1) / best_chunks;
        self->ps_chunks = (g.nz_loc + self->ps_kc - 1) / self->ps_kc;
        const int64_t items = tiles * self->ps_chunks;
        const int grid = (int)std::min(g_max, items);
        if (cudaMalloc(&self->ps_part, sizeof(double) * ((size_t)grid + 8)) != cudaSuccess) {
            cudaGetLastError();
            return 0;
        }
        self->ps_grid = grid;
        return grid;
    }
    void cg_persistent(CgFusedState* st, const double* b, double* x, double* r, double* Ap, double* P0, double* P1, int max_iter,
                       double* pAp_trace, double* rs_trace, int trace_cap, unsigned long long epoch, unsigned long long* stamps,
                       int stamp_cap) override {
        ATK_REQUIRE(ps_grid > 0, ATK_ERR_INVALID_ARGUMENT, "persistent CG loop not available for this operator");
        ATK_NVTX("atk::cg_persistent_kernel");
        constexpr int TX = 32, TY = 4;
        Context* c = ctx;
        const size_t plane = (size_t)g.nx * g.ny;
        PersistArgs pa{};
        pa.g = g;
        pa.st = st;
        pa.b = b; pa.x = x; pa.r = r; pa.Ap = Ap; pa.P0 = P0; pa.P1 = P1;
        pa.max_iter = max_iter;
        pa.pAp_trace = pAp_trace; pa.rs_trace = rs_trace; pa.trace_cap = trace_cap;
        pa.tiles_x = (g.nx + 2 * TX - 1) / (2 * TX);
        pa.tiles_y = (g.ny + TY - 1) / TY;
        pa.kc = ps_kc;
        pa.n_chunks = ps_chunks;
        pa.n_items = pa.tiles_x * pa.tiles_y * ps_chunks;
        pa.partials = ps_part;
        pa.arrive = reinterpret_cast<unsigned long long*>(ps_part + ps_grid);
        pa.ll = c->d_ll;
        pa.peer_ll = c->d_peer_ll;
        pa.world = c->world;
        pa.rank = c->rank;
        pa.epoch = epoch;
        if (c->world > 1) {
            pa.r_halo_lo = frecv;
            pa.r_halo_hi = frecv + 2 * plane;
            pa.my_p_lo[0] = frecv + plane;
            pa.my_p_hi[0] = frecv + 3 * plane;
            pa.my_p_lo[1] = frecv + 4 * plane;
            pa.my_p_hi[1] = frecv + 5 * plane;
            pa.nb_lo_r_hi = nb_lo_arena ? nb_lo_arena + 2 * plane : nullptr;
            pa.nb_hi_r_lo = nb_hi_arena ? nb_hi_arena : nullptr;
            pa.nb_lo_p_hi[0] = nb_lo_arena ? nb_lo_arena + 3 * plane : nullptr;
            pa.nb_lo_p_hi[1] = nb_lo_arena ? nb_lo_arena + 5 * plane : nullptr;
            pa.nb_hi_p_lo[0] = nb_hi_arena ? nb_hi_arena + plane : nullptr;
            pa.nb_hi_p_lo[1] = nb_hi_arena ? nb_hi_arena + 4 * plane : nullptr;
        }
        pa.stamps = stamps;
        pa.stamp_cap = stamp_cap;
        ATK_CUDA_CHECK(cudaMemsetAsync(pa.arrive, 0, sizeof(unsigned long long), c->stream));
        void* kargs[] = {&pa};
        ATK_CUDA_CHECK(cudaLaunchCooperativeKernel((void*)cg_persistent_kernel<TX, TY>, dim3((unsigned)ps_grid), dim3(TX, TY, 1), kargs,
                                                   fused_smem_bytes<TX, TY>(), c->stream));
        count_launch(c);
    }
    bool peer_mode = false;  // set by the CG driver for the duration of a solve (all ranks agree)

    // peer mode helpers for the CG driver
    bool peer_capable() const override {
        return ctx->world > 1 && ctx->p2p && fsmem && (g.nx % 2 == 0) && frecv != nullptr;
    }
    void set_peer_mode(bool on) override { peer_mode = on; }
    // where KB must mirror the first / last plane of r: the neighbours' r_hi / r_lo halo planes
    void peer_r_targets(double** lo_nb_r_hi, double** hi_nb_r_lo, int64_t* plane_elems) const override {
        const size_t plane = (size_t)g.nx * g.ny;
        *lo_nb_r_hi = nb_lo_arena ? nb_lo_arena + 2 * plane : nullptr;
        *hi_nb_r_lo = nb_hi_arena ? nb_hi_arena : nullptr;
        *plane_elems = (int64_t)plane;
    }
    // one NCCL exchange of the boundary planes of r and p (parity 0) at the start of a solve
    void initial_halo_exchange(const double* r, const double* p) override {
        Context* c = ctx;
        const size_t plane = (size_t)g.nx * g.ny;
        const size_t last = (size_t)(g.nz_loc - 1) * plane;
        ATK_CUDA_CHECK(cudaMemcpyAsync(fsend, r, sizeof(double) * plane, cudaMemcpyDeviceToDevice, c->stream));
        ATK_CUDA_CHECK(cudaMemcpyAsync(fsend + plane, p, sizeof(double) * plane, cudaMemcpyDeviceToDevice, c->stream));
        ATK_CUDA_CHECK(cudaMemcpyAsync(fsend + 2 * plane, r + last, sizeof(double) * plane, cudaMemcpyDeviceToDevice, c->stream));
        ATK_CUDA_CHECK(cudaMemcpyAsync(fsend + 3 * plane, p + last, sizeof(double) * plane, cudaMemcpyDeviceToDevice, c->stream));
        exchange_halo(c, fsend, fsend + 2 * plane, frecv, frecv + 2 * plane, 2 * plane, c->stream);
    }

    // 3-D tensor map over one slab-sized array for the TMA variant of KA: box (2TX+4) x (TY+2) x 1, zero fill outside
    template <int TX, int TY>
    CUtensorMap make_tensor_map(const double* base) const {
        typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        static EncodeFn encode = nullptr;
        if (!encode) {
            void* fn = nullptr;
            cudaDriverEntryPointQueryResult qres;
            ATK_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
            ATK_REQUIRE(fn && qres == cudaDriverEntryPointSuccess, ATK_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
            encode = reinterpret_cast<EncodeFn>(fn);
        }
        CUtensorMap tm;
        const cuuint64_t dims[3] = {(cuuint64_t)g.nx, (cuuint64_t)g.ny, (cuuint64_t)g.nz_loc};
        const cuuint64_t strides[2] = {(cuuint64_t)g.nx * 8u, (cuuint64_t)g.nx * (cuuint64_t)g.ny * 8u};
        const cuuint32_t box[3] = {(cuuint32_t)(2 * TX + 4), (cuuint32_t)(TY + 2), 1u};
        const cuuint32_t estr[3] = {1u, 1u, 1u};
        CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_128B;  // ATK_TMA_L2PROMO: 0 none, 1 64 B, 2 128 B, 3 256 B
        if (const char* e = std::getenv("ATK_TMA_L2PROMO")) {
            const int v = std::atoi(e);
            promo = v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : v == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                  : v == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
        }
        const CUresult rc = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(base), dims, strides, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        ATK_REQUIRE(rc == CUDA_SUCCESS, ATK_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)rc) + ")");
        return tm;
    }
    bool fused_tma = false;  // KA through tensor-map TMA (ATK_FUSED_TMA=1, single-GPU contexts)

    template <int TX, int TY>
    void launch_fused(const FusedArgs& fa, int k_lo, int k_hi, int partial_base, bool is_final, double* dot_out, int* blocks_out) {
        Context* c = ctx;
        const int planes = k_hi - k_lo;
        *blocks_out = 0;
        if (planes <= 0) return;
        const int kcc = chunk_planes(fkc, planes, ((g.nx + 2 * TX - 1) / (2 * TX)) * ((g.ny + TY - 1) / TY));
        dim3 grid((g.nx + 2 * TX - 1) / (2 * TX), (g.ny + TY - 1) / TY, (planes + kcc - 1) / kcc);
        dim3 block(TX, TY, 1);
        const int nblocks = (int)(grid.x * grid.y * grid.z);
        ATK_REQUIRE(partial_base + nblocks <= kMaxPartials, ATK_ERR_INVALID_ARGUMENT, "stencil grid exceeds the partial-sum scratch");
        auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
        const bool vec = (g.nx % 2 == 0) && al(fa.r) && al(fa.x) && al(fa.Ap) && al(fa.P0) && al(fa.P1);
        if (vec && fsmem && fused_tma && c->world == 1) {
            constexpr int TILE = (TY + 2) * (2 * TX + 4);
            constexpr size_t stage = (size_t)(TILE * 8 + 127) / 128 * 128;
            constexpr size_t smem = 2 * 4 * stage;
            // (the limit is per function AND per device: set on every launch - it is cheap and a solve captures KA once)
            ATK_CUDA_CHECK(cudaFuncSetAttribute(stencil7_cg_fused_tma_kernel<TX, TY>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)smem));
            const CUtensorMap tm_r = make_tensor_map<TX, TY>(fa.r), tm_p0 = make_tensor_map<TX, TY>(fa.P0),
                              tm_p1 = make_tensor_map<TX, TY>(fa.P1);
            stencil7_cg_fused_tma_kernel<TX, TY><<<grid, block, smem, c->stream>>>(tm_r, tm_p0, tm_p1, g, fa, k_lo, k_hi, kcc,
                                                                                   c->d_partials, partial_base, c->d_tickets + 5,
                                                                                   dot_out, is_final ? 1 : 0);
        } else if (vec && fsmem) {
            constexpr size_t smem = fused_smem_bytes<TX, TY>();
            stencil7_cg_fused_smem_kernel<TX, TY><<<grid, block, smem, c->stream>>>(g, fa, k_lo, k_hi, kcc, c->d_partials,
                                                                                    partial_base, c->d_tickets + 5, dot_out,
                                                                                    is_final ? 1 : 0);
        }
        else if (vec)
            stencil7_cg_fused_kernel<TX, TY, true><<<grid, block, 0, c->stream>>>(g, fa, k_lo, k_hi, kcc, c->d_partials, partial_base,
                                                                                  c->d_tickets + 5, dot_out, is_final ? 1 : 0);
        else
            stencil7_cg_fused_kernel<TX, TY, false><<<grid, block, 0, c->stream>>>(g, fa, k_lo, k_hi, kcc, c->d_partials, partial_base,
                                                                                   c->d_tickets + 5, dot_out, is_final ? 1 : 0);
        ATK_CUDA_CHECK(cudaGetLastError());
        count_launch(c);
        *blocks_out = nblocks;
    }
    void launch_fused_tile(const FusedArgs& fa, int k_lo, int k_hi, int partial_base, bool is_final, double* dot_out, int* nb) {
        switch (ftile) {
            case 0: launch_fused<32, 8>(fa, k_lo, k_hi, partial_base, is_final, dot_out, nb); break;
            case 1: launch_fused<64, 4>(fa, k_lo, k_hi, partial_base, is_final, dot_out, nb); break;
            case 2: launch_fused<32, 16>(fa, k_lo, k_hi, partial_base, is_final, dot_out, nb); break;
            case 4: launch_fused<32, 4>(fa, k_lo, k_hi, partial_base, is_final, dot_out, nb); break;
            case 5: launch_fused<64, 2>(fa, k_lo, k_hi, partial_base, is_final, dot_out, nb); break;
            default: launch_fused<64, 8>(fa, k_lo, k_hi, partial_base, is_final, dot_out, nb); break;
        }
    }

    void cg_fused_step(CgFusedState* st, const double* r, double* x, double* Ap, double* P0, double* P1, int rr_slot,
                       int dot_slot, double* rs_trace) override {
        Context* c = ctx;
        const size_t plane = (size_t)g.nx * g.ny;
        FusedArgs fa{r, x, Ap, P0, P1, nullptr, nullptr, nullptr, nullptr, st, c->slot(rr_slot), c->world, rs_trace};
        double* dot_out = c->slot(dot_slot) + c->rank;
        int nb = 0;
        fa.p2p = 0;
        if (c->world == 1) {
            launch_fused_tile(fa, 0, g.nz_loc, 0, true, dot_out, &nb);
        } else if (peer_mode) {
            // halo planes were written by the neighbours' previous KB (r) and KA (p); one launch over all planes
            fa.p2p = 1;
            fa.rank = c->rank;
            fa.ll = c->d_ll;
            fa.peer_ll = c->d_peer_ll;
            fa.r_halo_lo = frecv;
            fa.r_halo_hi = frecv + 2 * plane;
            fa.my_p_lo[0] = frecv + plane;
            fa.my_p_hi[0] = frecv + 3 * plane;
            fa.my_p_lo[1] = frecv + 4 * plane;
            fa.my_p_hi[1] = frecv + 5 * plane;
            fa.p_halo_lo = fa.my_p_lo[0];
            fa.p_halo_hi = fa.my_p_hi[0];
            fa.nb_lo_p_hi[0] = nb_lo_arena ? nb_lo_arena + 3 * plane : nullptr;
            fa.nb_lo_p_hi[1] = nb_lo_arena ? nb_lo_arena + 5 * plane : nullptr;
            fa.nb_hi_p_lo[0] = nb_hi_arena ? nb_hi_arena + plane : nullptr;
            fa.nb_hi_p_lo[1] = nb_hi_arena ? nb_hi_arena + 4 * plane : nullptr;
            launch_fused_tile(fa, 0, g.nz_loc, 0, true, dot_out, &nb);
            return;  // the p.Ap partials travel by peer stores: no all-gather
        } else {
            fa.r_halo_lo = frecv;
            fa.p_halo_lo = frecv + plane;
            fa.r_halo_hi = frecv + 2 * plane;
            fa.p_halo_hi = frecv + 3 * plane;
            pack_halo_kernel<<<c->sm_count * 2, 256, 0, c->stream>>>(st, r, P0, P1, (int64_t)plane,
                                                                     (int64_t)(g.nz_loc - 1) * (int64_t)plane, fsend);
            count_launch(c);
            ATK_CUDA_CHECK(cudaEventRecord(c->ev_fork, c->stream));
            ATK_CUDA_CHECK(cudaStreamWaitEvent(c->comm_stream, c->ev_fork, 0));
            exchange_halo(c, fsend, fsend + 2 * plane, frecv, frecv + 2 * plane, 2 * plane, c->comm_stream);
            ATK_CUDA_CHECK(cudaEventRecord(c->ev_join, c->comm_stream));
            int nb_int = 0, nb_lo = 0, nb_hi = 0;
            if (g.nz_loc > 2) launch_fused_tile(fa, 1, g.nz_loc - 1, 0, false, dot_out, &nb_int);
            ATK_CUDA_CHECK(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
            if (g.nz_loc >= 2) {
                launch_fused_tile(fa, 0, 1, nb_int, false, dot_out, &nb_lo);
                launch_fused_tile(fa, g.nz_loc - 1, g.nz_loc, nb_int + nb_lo, true, dot_out, &nb_hi);
            } else {
                launch_fused_tile(fa, 0, g.nz_loc, nb_int, true, dot_out, &nb_lo);
            }
        }
        allgather_slot(c, dot_slot, c->stream);
    }

    template <int TX, int TY>
    void launch(const double* x, double* y, int k_lo, int k_hi, bool dot, int partial_base, bool dot_final, double* dot_out,
                const int* skip_flag, int* blocks_out) {
        Context* c = ctx;
        const int planes = k_hi - k_lo;
        *blocks_out = 0;
        if (planes <= 0) return;
        const int kcc = chunk_planes(kc, planes, ((g.nx + 2 * TX - 1) / (2 * TX)) * ((g.ny + TY - 1) / TY));
        dim3 grid((g.nx + 2 * TX - 1) / (2 * TX), (g.ny + TY - 1) / TY, (planes + kcc - 1) / kcc);
        dim3 block(TX, TY, 1);
        const int nblocks = (int)(grid.x * grid.y * grid.z);
        ATK_REQUIRE(partial_base + nblocks <= kMaxPartials, ATK_ERR_INVALID_ARGUMENT, "stencil grid exceeds the partial-sum scratch");
        const bool vec = (g.nx % 2 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15u) == 0) &&
                         ((reinterpret_cast<uintptr_t>(y) & 15u) == 0);
        if (vec && fsmem && use_tma) {
            constexpr size_t smem = (size_t)4 * (TY + 2) * (2 * TX + 4) * sizeof(double) + kRingShift * sizeof(double);
            if (dot)
                stencil7_tma_kernel<TX, TY, true><<<grid, block, smem, c->stream>>>(g, x, y, halo_lo, halo_hi, k_lo, k_hi, kcc,
                                                                                    c->d_partials, partial_base, c->d_tickets + 2,
                                                                                    dot_out, dot_final ? 1 : 0, skip_flag);
            else
                stencil7_tma_kernel<TX, TY, false><<<grid, block, smem, c->stream>>>(g, x, y, halo_lo, halo_hi, k_lo, k_hi, kcc,
                                                                                     c->d_partials, partial_base, c->d_tickets + 2,
                                                                                     dot_out, dot_final ? 1 : 0, skip_flag);
            ATK_CUDA_CHECK(cudaGetLastError());
            count_launch(c);
            *blocks_out = nblocks;
            return;
        }
        if (vec && fsmem) {
            constexpr size_t smem = (size_t)4 * (TY + 2) * (2 * TX + 4) * sizeof(double) + kRingShift * sizeof(double);
            if (dot)
                stencil7_smem_kernel<TX, TY, true><<<grid, block, smem, c->stream>>>(g, x, y, halo_lo, halo_hi, k_lo, k_hi, kcc,
                                                                                     c->d_partials, partial_base, c->d_tickets + 2,
                                                                                     dot_out, dot_final ? 1 : 0, skip_flag);
            else
                stencil7_smem_kernel<TX, TY, false><<<grid, block, smem, c->stream>>>(g, x, y, halo_lo, halo_hi, k_lo, k_hi, kcc,
                                                                                      c->d_partials, partial_base, c->d_tickets + 2,
                                                                                      dot_out, dot_final ? 1 : 0, skip_flag);
            ATK_CUDA_CHECK(cudaGetLastError());
            count_launch(c);
            *blocks_out = nblocks;
            return;
        }
#define ATK_ST(V, D)                                                                                                     \
    stencil7_kernel<TX, TY, V, D><<<grid, block, 0, c->stream>>>(g, x, y, halo_lo, halo_hi, k_lo, k_hi, kcc, c->d_partials, \
                                                                 partial_base, c->d_tickets + 2, dot_out, dot_final ? 1 : 0, \
                                                                 skip_flag)
        if (vec) { if (dot) ATK_ST(true, true); else ATK_ST(true, false); }
        else     { if (dot) ATK_ST(false, true); else ATK_ST(false, false); }
#undef ATK_ST
        ATK_CUDA_CHECK(cudaGetLastError());
        count_launch(c);
        *blocks_out = nblocks;
    }

    void launch_tile(const double* x, double* y, int k_lo, int k_hi, bool dot, int partial_base, bool dot_final,
                     double* dot_out, const int* skip_flag, int* blocks_out) {
        switch (tile) {
            case 1: launch<64, 4>(x, y, k_lo, k_hi, dot, partial_base, dot_final, dot_out, skip_flag, blocks_out); break;
            case 2: launch<32, 16>(x, y, k_lo, k_hi, dot, partial_base, dot_final, dot_out, skip_flag, blocks_out); break;
            case 3: launch<64, 8>(x, y, k_lo, k_hi, dot, partial_base, dot_final, dot_out, skip_flag, blocks_out); break;
            case 0: launch<32, 8>(x, y, k_lo, k_hi, dot, partial_base, dot_final, dot_out, skip_flag, blocks_out); break;
            case 5: launch<64, 2>(x, y, k_lo, k_hi, dot, partial_base, dot_final, dot_out, skip_flag, blocks_out); break;
            default: launch<32, 4>(x, y, k_lo, k_hi, dot, partial_base, dot_final, dot_out, skip_flag, blocks_out); break;
        }
    }

    template <int TX, int TY>
    void launch_peer(const double* x, double* y, bool dot, double* dot_out, const int* skip_flag, const PeerApply& pa) {
        Context* c = ctx;
        const int planes = g.nz_loc;
        // at least four chunks where the slab allows it, so that the inner ones are computed while the boundary planes are
        // still crossing NVLink; plus the slice of push-only CTAs (blockIdx.z == 0)
        int kcc = chunk_planes(kc, planes, ((g.nx + 2 * TX - 1) / (2 * TX)) * ((g.ny + TY - 1) / TY));
        kcc = std::max(1, std::min(kcc, std::max(8, (planes + 3) / 4)));
        if (const char* e = std::getenv("ATK_PEER_APPLY_KC")) kcc = std::max(1, std::min(std::atoi(e), planes));
        dim3 grid((g.nx + 2 * TX - 1) / (2 * TX), (g.ny + TY - 1) / TY, (planes + kcc - 1) / kcc + 1);
        dim3 block(TX, TY, 1);
        const int nblocks = (int)(grid.x * grid.y * grid.z);
        ATK_REQUIRE(nblocks <= kMaxPartials, ATK_ERR_INVALID_ARGUMENT, "stencil grid exceeds the partial-sum scratch");
        constexpr size_t smem = (size_t)4 * (TY + 2) * (2 * TX + 4) * sizeof(double) + kRingShift * sizeof(double);
        if (dot)
            stencil7_smem_kernel<TX, TY, true, true><<<grid, block, smem, c->stream>>>(g, x, y, halo_lo, halo_hi, 0, planes, kcc,
                                                                                        c->d_partials, 0, c->d_tickets + 2, dot_out, 1,
                                                                                        skip_flag, pa);
        else
            stencil7_smem_kernel<TX, TY, false, true><<<grid, block, smem, c->stream>>>(g, x, y, halo_lo, halo_hi, 0, planes, kcc,
                                                                                         c->d_partials, 0, c->d_tickets + 2, dot_out, 1,
                                                                                         skip_flag, pa);
        ATK_CUDA_CHECK(cudaGetLastError());
        count_launch(c);
    }
    void launch_peer_tile(const double* x, double* y, bool dot, double* dot_out, const int* skip_flag, const PeerApply& pa) {
        switch (tile) {
            case 1: launch_peer<64, 4>(x, y, dot, dot_out, skip_flag, pa); break;
            case 2: launch_peer<32, 16>(x, y, dot, dot_out, skip_flag, pa); break;
            case 3: launch_peer<64, 8>(x, y, dot, dot_out, skip_flag, pa); break;
            case 0: launch_peer<32, 8>(x, y, dot, dot_out, skip_flag, pa); break;
            case 5: launch_peer<64, 2>(x, y, dot, dot_out, skip_flag, pa); break;
            default: launch_peer<32, 4>(x, y, dot, dot_out, skip_flag, pa); break;
        }
    }
    static bool apply_over_nccl() {
        const char* e = std::getenv("ATK_COMM");
        return e && std::string(e) == "nccl";
    }
    void apply(const double* x, double* y, int dot_slot, const int* skip_flag) override {
        ATK_NVTX("atk::stencil7_apply");
        Context* c = ctx;
        const bool dot = dot_slot >= 0;
        double* dot_out = dot ? c->slot(dot_slot) + c->rank : nullptr;
        int nb = 0;
        if (c->world == 1) {
            launch_tile(x, y, 0, g.nz_loc, dot, 0, true, dot_out, skip_flag, &nb);
        } else {
            // halo exchange on the communication stream, overlapped with the planes that do not need it
            const size_t plane = (size_t)g.nx * g.ny;
            const bool peer_push = ph.ready && !apply_over_nccl() && !c->capturing();
            const bool vec_ok = (g.nx % 2 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15u) == 0) &&
                                ((reinterpret_cast<uintptr_t>(y) & 15u) == 0);
            if (peer_push && vec_ok && fsmem && !use_tma && !std::getenv("ATK_APPLY_SPLIT")) {
                // one launch for the whole slab: its boundary CTAs do the neighbour stores and the waiting themselves
                const PeerHalo::Targets tg = ph.begin();
                halo_lo = ph.lo();
                halo_hi = ph.hi();
                PeerApply pa;
                pa.prev_dst = tg.prev_dst;
                pa.next_dst = tg.next_dst;
                pa.prev_flag = tg.prev_flag;
                pa.next_flag = tg.next_flag;
                pa.my_flags = tg.my_flags;
                pa.tickets = ph.d_ticket + 1;
                pa.seq = tg.seq;
                launch_peer_tile(x, y, dot, dot_out, skip_flag, pa);
                if (dot) allgather_slot(c, dot_slot, c->stream);
                return;
            }
            if (peer_push) {  // boundary planes stored straight into the neighbours' buffers over NVLink, flag released
                ph.push(x, plane, x + (size_t)(g.nz_loc - 1) * plane, plane);
                halo_lo = ph.lo();
                halo_hi = ph.hi();
            } else {
                halo_lo = nccl_halo_lo;
                halo_hi = nccl_halo_hi;
                ATK_CUDA_CHECK(cudaEventRecord(c->ev_fork, c->stream));
                ATK_CUDA_CHECK(cudaStreamWaitEvent(c->comm_stream, c->ev_fork, 0));
                exchange_halo(c, x, x + (size_t)(g.nz_loc - 1) * plane, nccl_halo_lo, nccl_halo_hi, plane, c->comm_stream);
                ATK_CUDA_CHECK(cudaEventRecord(c->ev_join, c->comm_stream));
            }
            int nb_int = 0, nb_lo = 0, nb_hi = 0;
            if (g.nz_loc > 2) launch_tile(x, y, 1, g.nz_loc - 1, dot, 0, false, dot_out, skip_flag, &nb_int);
            if (peer_push) ph.wait();
            else ATK_CUDA_CHECK(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
            if (g.nz_loc >= 2) {
                launch_tile(x, y, 0, 1, dot, nb_int, false, dot_out, skip_flag, &nb_lo);
                launch_tile(x, y, g.nz_loc - 1, g.nz_loc, dot, nb_int + nb_lo, true, dot_out, skip_flag, &nb_hi);
            } else {
                launch_tile(x, y, 0, g.nz_loc, dot, nb_int, true, dot_out, skip_flag, &nb_lo);
            }
        }
        if (dot) allgather_slot(c, dot_slot, c->stream);
    }
};

// ---------------------------------------------------------------------------------------------- 27-point operator
// Matrix-free form of the synthetic operator of BASELINE config 5: open boundaries (neighbours outside
// the grid are dropped), A(r, r+d) = coef[d].  Taps are added from 0.0 in ascending column order (dk, dj, di
// ascending), which is the order the assembled CSC matvec uses, so the result is bit-identical to the SELL operator
// built by atk_operator_stencil27_assembled.  Zero coefficients are skipped (they are never stored by addValue).
// Two kernels: stencil27_smem_kernel (nx even, 16-byte aligned vectors) streams x through the same cp.async ring of
// shared-memory plane tiles as the 7-point kernels and keeps the 3x3x4 neighbourhood of its two points in registers,
// loading only the 12 values of the incoming plane per step; rows, columns and planes outside the grid are zeros in
// the tiles, and a product with such a zero leaves the running sum unchanged, exactly as dropping the tap does.
// stencil27_kernel (neighbours through L1/L2, explicit tests) is the general fallback.
struct Coef27 {
    double v[27];
};
__global__ void __launch_bounds__(256) stencil27_kernel(StencilGeom g, Coef27 cf, const double* __restrict__ x,
                                                        double* __restrict__ y, const double* __restrict__ halo_lo,
                                                        const double* __restrict__ halo_hi, int k_lo, int k_hi) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= g.nx || j >= g.ny) return;
    const int64_t sxy = (int64_t)g.nx * g.ny;
    for (int k = k_lo + blockIdx.z; k < k_hi; k += gridDim.z) {
        const int kg = g.k_off + k;
        double acc = 0.0;
#pragma unroll
        for (int dk = -1; dk <= 1; ++dk) {
            const int kk = k + dk;
            if (kg + dk < 0 || kg + dk >= g.nz) continue;
            const double* pl = kk < 0 ? halo_lo : (kk >= g.nz_loc ? halo_hi : x + (int64_t)kk * sxy);
#pragma unroll
            for (int dj = -1; dj <= 1; ++dj) {
                const int jj = j + dj;
                if (jj < 0 || jj >= g.ny) continue;
#pragma unroll
                for (int di = -1; di <= 1; ++di) {
                    const int ii = i + di;
                    const double c = cf.v[(dk + 1) * 9 + (dj + 1) * 3 + (di + 1)];
                    if (ii < 0 || ii >= g.nx || c == 0.0) continue;
                    acc = add_rn(acc, mul_rn(c, pl[(int64_t)jj * g.nx + ii]));
                }
            }
        }
        y[(int64_t)k * sxy + (int64_t)j * g.nx + i] = acc;
    }
}

// DENSE: all 27 coefficients are non-zero (the config-5 operator), so the per-tap zero tests are compiled out - the kernel
// is issue-bound (ncu: 79 % issue-active, FP64 pipe 50 %).
template <int TX, int TY, bool DENSE, bool PEER = false>
__global__ void __launch_bounds__(TX* TY) stencil27_smem_kernel(StencilGeom g, Coef27 cf, const double* __restrict__ x,
                                                                 double* __restrict__ y, const double* __restrict__ halo_lo,
                                                                 const double* __restrict__ halo_hi, int k_lo, int k_hi, int kc,
                                                                 PeerApply pa = PeerApply()) {
    constexpr int DEPTH = 4;
    constexpr int W = 2 * TX + 4;
    constexpr int STAGE = (TY + 2) * W;
    constexpr int NT = TX * TY;
    constexpr int CPR = W / 2;
    constexpr int NCH = (TY + 2) * CPR;
    constexpr int Q = (NCH + NT - 1) / NT;
    extern __shared__ __align__(128) double smem_raw[];
    double* ring = smem_raw + kRingShift;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int tid = ty * TX + tx;
    const int i_lo = blockIdx.x * 2 * TX;
    const int j_lo = blockIdx.y * TY;
    const int i0 = i_lo + 2 * tx;
    const int j = j_lo + ty;
    int zc = blockIdx.z;
    const bool active = (i0 < g.nx) && (j < g.ny);
    const int64_t sxy = (int64_t)g.nx * g.ny;
    if constexpr (PEER) {  // push-only CTAs first, boundary chunks last: see stencil7_smem_kernel
        if (zc == 0) {
            const int64_t row = (int64_t)j * g.nx + i0;
            if (active) {
                if (pa.prev_dst) *reinterpret_cast<double2*>(pa.prev_dst + row) = *reinterpret_cast<const double2*>(x + row);
                if (pa.next_dst)
                    *reinterpret_cast<double2*>(pa.next_dst + row) = *reinterpret_cast<const double2*>(x + (int64_t)(g.nz_loc - 1) * sxy + row);
            }
            __threadfence_system();
            __syncthreads();
            if (tid == 0) {
                const unsigned n_tiles = gridDim.x * gridDim.y;
                if (atomicInc(pa.tickets, n_tiles - 1) == n_tiles - 1) {
                    __threadfence_system();
                    if (pa.prev_flag) st_release_sys(pa.prev_flag, pa.seq);
                    if (pa.next_flag) st_release_sys(pa.next_flag, pa.seq);
                }
            }
            return;
        }
        const int nchunk = (int)gridDim.z - 1;
        const int q = zc - 1;
        zc = nchunk >= 2 ? (q < nchunk - 2 ? q + 1 : (q == nchunk - 2 ? 0 : nchunk - 1)) : q;
    }
    const int kb = k_lo + zc * kc;
    const int ke = min(kb + kc, k_hi);
    if (kb >= ke) return;
    if constexpr (PEER) {
        const bool need_lo = kb == 0 && g.k_off > 0, need_hi = ke == g.nz_loc && g.k_off + g.nz_loc < g.nz;
        if (need_lo || need_hi) {
            if (tid == 0) {
                if (need_lo) peer_flag_wait(pa.my_flags, pa.seq);
                if (need_hi) peer_flag_wait(pa.my_flags + 1, pa.seq);
            }
            __syncthreads();
        }
    }
    // everything that is never a copy destination (rows / columns outside the grid) stays zero for the whole kernel
    for (int e = tid; e < DEPTH * STAGE; e += NT) ring[e] = 0.0;
    __syncthreads();
    const int k_first = max(kb - 1, -g.k_off);             // planes kb-1 .. ke are read; these two bound the existing ones
    const int k_last = min(ke, g.nz - 1 - g.k_off);
    const unsigned s0 = (unsigned)__cvta_generic_to_shared(ring);
    unsigned c_sm[Q];
    int64_t c_off[Q];
    bool c_ok[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const int e = tid + q * NT;
        const int jj = e / CPR, c = e - jj * CPR;
        const int gj = j_lo - 1 + jj, gi = i_lo - 2 + 2 * c;
        c_ok[q] = e < NCH && gj >= 0 && gj < g.ny && gi >= 0 && gi < g.nx;  // corners included: the 27 taps reach them
        c_sm[q] = 8u * (unsigned)(jj * W + 2 * c);
        c_off[q] = (int64_t)gj * g.nx + gi;
    }
    auto plane_of = [&](int k) -> const double* { return k < 0 ? halo_lo : (k >= g.nz_loc ? halo_hi : x + (int64_t)k * sxy); };
    // plane k -> ring stage (k - kb + 1) mod DEPTH; a plane outside the global grid is written as zeros
    auto issue_plane = [&](int k, const double* plane) {
        if (k >= kb - 1 && k <= ke) {
            const unsigned so = (unsigned)((k - kb + 1) & (DEPTH - 1)) * (unsigned)(STAGE * 8);
            if (k >= k_first && k <= k_last) {
#pragma unroll
                for (int q = 0; q < Q; ++q)
                    if (c_ok[q]) cp_async16_s(s0 + c_sm[q] + so, plane + c_off[q]);
            } else {
#pragma unroll
                for (int q = 0; q < Q; ++q)
                    if (c_ok[q]) *reinterpret_cast<double2*>(reinterpret_cast<char*>(ring) + c_sm[q] + so) = make_double2(0.0, 0.0);
            }
        }
        cp_async_commit();
    };
    // the 3 x 4 values of one plane this thread needs: tile rows ty .. ty+2, tile columns 2tx+1 .. 2tx+4
    struct Nb {
        double v[3][4];
    };
    const double* corner = ring + ty * W + 2 * tx + 1;
    auto load_nb = [&](Nb& n, unsigned m) {
        const double* st = corner + (m & (DEPTH - 1)) * STAGE;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            n.v[r][0] = st[r * W];
            const double2 mid = *reinterpret_cast<const double2*>(st + r * W + 1);
            n.v[r][1] = mid.x;
            n.v[r][2] = mid.y;
            n.v[r][3] = st[r * W + 3];
        }
    };
    auto taps = [&](const Nb& n, int dk, double& acc0, double& acc1) {
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const double c = cf.v[dk * 9 + r * 3 + d];
                if (DENSE || c != 0.0) {  // uniform: zero coefficients are not stored by the assembled form
                    acc0 = add_rn(acc0, mul_rn(c, n.v[r][d]));
                    acc1 = add_rn(acc1, mul_rn(c, n.v[r][d + 1]));
                }
            }
    };
    issue_plane(kb - 1, plane_of(kb - 1));
    issue_plane(kb, plane_of(kb));
    issue_plane(kb + 1, plane_of(kb + 1));
    const double* plane = plane_of(kb + 2);
    issue_plane(kb + 2, plane);
    cp_async_wait<2>();
    __syncthreads();
    Nb n0, n1, n2;
    load_nb(n0, 0);
    load_nb(n1, 1);
    double* yp = y + (int64_t)kb * sxy + (int64_t)j * g.nx + i0;
    unsigned m = 1;  // ring position of plane k
    int k = kb;
    auto step = [&](const Nb& A, const Nb& B, Nb& C) {
        cp_async_wait<1>();
        __syncthreads();
        plane = (k + 3 == g.nz_loc) ? halo_hi : plane + sxy;
        issue_plane(k + 3, plane);
        load_nb(C, m + 1);
        double acc0 = 0.0, acc1 = 0.0;
        taps(A, 0, acc0, acc1);
        taps(B, 1, acc0, acc1);
        taps(C, 2, acc0, acc1);
        if (active) *reinterpret_cast<double2*>(yp) = make_double2(acc0, acc1);
        ++k;
        ++m;
        yp += sxy;
    };
    while (k + 3 <= ke) {
        step(n0, n1, n2);
        step(n1, n2, n0);
        step(n2, n0, n1);
    }
    if (k < ke) {
        step(n0, n1, n2);
        if (k < ke) step(n1, n2, n0);
    }
    cp_async_wait<0>();
}

struct Stencil27 : public Operator {
    StencilGeom g;
    Coef27 cf;
    double* nccl_halo_lo = nullptr;
    double* nccl_halo_hi = nullptr;
    const double* halo_lo = nullptr;  // the halo planes the kernels of the current apply read
    const double* halo_hi = nullptr;
    PeerHalo ph;
    ~Stencil27() override {
        cudaFree(nccl_halo_lo);
        cudaFree(nccl_halo_hi);
        ph.release();
    }
    int kc27 = 64;  // planes per chunk of the shared-memory kernel (ATK_STENCIL27_KC)
    bool vec_ok(const double* x, const double* y) const {
        return (g.nx % 2 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15u) == 0) && ((reinterpret_cast<uintptr_t>(y) & 15u) == 0) &&
               !std::getenv("ATK_STENCIL27_SIMPLE");
    }
    void launch(const double* x, double* y, int k_lo, int k_hi, const PeerApply* pa = nullptr) {
        if (k_hi <= k_lo) return;
        const bool vec = vec_ok(x, y);
        if (vec && pa) {  // the whole slab in one launch, neighbour stores and waits inside (see stencil7_smem_kernel PEER)
            constexpr int TX = 32, TY = 4;
            const int tiles = ((g.nx + 2 * TX - 1) / (2 * TX)) * ((g.ny + TY - 1) / TY);
            int kcc = std::min(kc27, k_hi - k_lo);
            while (kcc > 8 && (int64_t)tiles * ((k_hi - k_lo + kcc - 1) / kcc) < (int64_t)ctx->sm_count * 8) kcc = (kcc + 1) / 2;
            kcc = std::max(1, std::min(kcc, std::max(8, (k_hi - k_lo + 3) / 4)));  // inner chunks run during the handshake
            dim3 grid((g.nx + 2 * TX - 1) / (2 * TX), (g.ny + TY - 1) / TY, (k_hi - k_lo + kcc - 1) / kcc + 1);  // + push-only slice
            constexpr size_t smem = (size_t)4 * (TY + 2) * (2 * TX + 4) * sizeof(double) + kRingShift * sizeof(double);
            bool dense = true;
            for (int t = 0; t < 27; ++t) dense = dense && cf.v[t] != 0.0;
            if (dense)
                stencil27_smem_kernel<TX, TY, true, true><<<grid, dim3(TX, TY, 1), smem, ctx->stream>>>(g, cf, x, y, halo_lo, halo_hi, k_lo,
                                                                                                        k_hi, kcc, *pa);
            else
                stencil27_smem_kernel<TX, TY, false, true><<<grid, dim3(TX, TY, 1), smem, ctx->stream>>>(g, cf, x, y, halo_lo, halo_hi,
                                                                                                         k_lo, k_hi, kcc, *pa);
            ATK_CUDA_CHECK(cudaGetLastError());
            count_launch(ctx);
            return;
        }
        if (vec) {
            constexpr int TX = 32, TY = 4;
            const int tiles = ((g.nx + 2 * TX - 1) / (2 * TX)) * ((g.ny + TY - 1) / TY);
            int kcc = std::min(kc27, k_hi - k_lo);  // enough CTAs for a few waves, but never chunks shorter than 8 planes
            while (kcc > 8 && (int64_t)tiles * ((k_hi - k_lo + kcc - 1) / kcc) < (int64_t)ctx->sm_count * 8) kcc = (kcc + 1) / 2;
            dim3 grid((g.nx + 2 * TX - 1) / (2 * TX), (g.ny + TY - 1) / TY, (k_hi - k_lo + kcc - 1) / kcc);
            constexpr size_t smem = (size_t)4 * (TY + 2) * (2 * TX + 4) * sizeof(double) + kRingShift * sizeof(double);
            bool dense = true;
            for (int t = 0; t < 27; ++t) dense = dense && cf.v[t] != 0.0;
            if (dense)
                stencil27_smem_kernel<TX, TY, true><<<grid, dim3(TX, TY, 1), smem, ctx->stream>>>(g, cf, x, y, halo_lo, halo_hi, k_lo,
                                                                                                  k_hi, kcc);
            else
                stencil27_smem_kernel<TX, TY, false><<<grid, dim3(TX, TY, 1), smem, ctx->stream>>>(g, cf, x, y, halo_lo, halo_hi, k_lo,
                                                                                                   k_hi, kcc);
        } else {
            dim3 block(32, 8, 1);
            dim3 grid((g.nx + 31) / 32, (g.ny + 7) / 8, std::min(k_hi - k_lo, 64));
            stencil27_kernel<<<grid, block, 0, ctx->stream>>>(g, cf, x, y, halo_lo, halo_hi, k_lo, k_hi);
        }
        ATK_CUDA_CHECK(cudaGetLastError());
        count_launch(ctx);
    }
    void apply(const double* x, double* y, int dot_slot, const int* skip_flag) override {
        ATK_NVTX("atk::stencil27_apply");
        Context* c = ctx;
        (void)skip_flag;  // only the CG loop passes a skip flag, and CG on this operator goes through launch_dot below
        if (c->world == 1) {
            launch(x, y, 0, g.nz_loc);
        } else {
            const size_t plane = (size_t)g.nx * g.ny;
            const char* comm = std::getenv("ATK_COMM");
            const bool peer_push = ph.ready && !(comm && std::string(comm) == "nccl") && !c->capturing();
            if (peer_push && vec_ok(x, y) && !std::getenv("ATK_APPLY_SPLIT")) {
                const PeerHalo::Targets tg = ph.begin();
                halo_lo = ph.lo();
                halo_hi = ph.hi();
                PeerApply pa;
                pa.prev_dst = tg.prev_dst;
                pa.next_dst = tg.next_dst;
                pa.prev_flag = tg.prev_flag;
                pa.next_flag = tg.next_flag;
                pa.my_flags = tg.my_flags;
                pa.tickets = ph.d_ticket + 1;
                pa.seq = tg.seq;
                launch(x, y, 0, g.nz_loc, &pa);
                if (dot_slot >= 0) launch_dot(c, n_rows_local, x, y, dot_slot, skip_flag);
                return;
            }
            if (peer_push) {
                ph.push(x, plane, x + (size_t)(g.nz_loc - 1) * plane, plane);
                halo_lo = ph.lo();
                halo_hi = ph.hi();
            } else {
                halo_lo = nccl_halo_lo;
                halo_hi = nccl_halo_hi;
                ATK_CUDA_CHECK(cudaEventRecord(c->ev_fork, c->stream));
                ATK_CUDA_CHECK(cudaStreamWaitEvent(c->comm_stream, c->ev_fork, 0));
                exchange_halo(c, x, x + (size_t)(g.nz_loc - 1) * plane, nccl_halo_lo, nccl_halo_hi, plane, c->comm_stream);
                ATK_CUDA_CHECK(cudaEventRecord(c->ev_join, c->comm_stream));
            }
            launch(x, y, 1, g.nz_loc - 1);  // planes that need no halo run while the exchange is in flight
            if (peer_push) ph.wait();
            else ATK_CUDA_CHECK(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
            launch(x, y, 0, 1);
            if (g.nz_loc > 1) launch(x, y, g.nz_loc - 1, g.nz_loc);
        }
        if (dot_slot >= 0) launch_dot(c, n_rows_local, x, y, dot_slot, skip_flag);
    }
};

}  // namespace

Operator* make_stencil7(Context* ctx, int nx, int ny, int nz, double cx, double cy, double cz) {
    ATK_REQUIRE(nx > 0 && ny > 0 && nz > 0, ATK_ERR_INVALID_ARGUMENT, "grid dimensions must be positive");
    ATK_REQUIRE(nz >= ctx->world, ATK_ERR_INVALID_ARGUMENT, "fewer z-planes than ranks");
    auto* S = new Stencil7();
    S->ctx = ctx;
    // balanced z-slab split: the first (nz % world) ranks get one extra plane
    int k_begin = 0, k_end = 0;
    slab_partition(nz, ctx->rank, ctx->world, &k_begin, &k_end);
    const int nz_loc = k_end - k_begin;
    S->g = StencilGeom{nx, ny, nz, k_begin, nz_loc, cx, cy, cz, -2.0 * (cx + cy + cz)};  // cc: Poisson.hpp:57
    const int64_t plane = (int64_t)nx * ny;
    S->n_rows_local = plane * nz_loc;
    S->n_cols = plane * nz;
    S->row_begin = plane * k_begin;
    const int64_t interior = (nx > 2 && ny > 2 && nz > 2) ? (int64_t)(nx - 2) * (ny - 2) * (nz - 2) : 0;
    S->nnz = (plane * nz - interior) + 7 * interior;
    // opt in to > 48 KB of dynamic shared memory for every tile shape of the cp.async kernel (per device)
    allow_fused_smem<32, 8>(); allow_fused_smem<64, 4>(); allow_fused_smem<32, 16>(); allow_fused_smem<32, 4>();
    allow_fused_smem<64, 2>(); allow_fused_smem<64, 8>();
    if (const char* e = std::getenv("ATK_STENCIL_KC")) S->kc = std::max(1, std::atoi(e));
    if (const char* e = std::getenv("ATK_STENCIL_TILE")) S->tile = std::atoi(e);
    if (const char* e = std::getenv("ATK_FUSED_KC")) S->fkc = std::max(1, std::atoi(e));
    if (const char* e = std::getenv("ATK_FUSED_TILE")) S->ftile = std::atoi(e);
    if (const char* e = std::getenv("ATK_FUSED_SMEM")) S->fsmem = std::atoi(e) != 0;
    if (const char* e = std::getenv("ATK_STENCIL_TMA")) S->use_tma = std::atoi(e) != 0;
    if (const char* e = std::getenv("ATK_FUSED_TMA")) S->fused_tma = std::atoi(e) != 0;
    try {
        if (ctx->world > 1) {
            ATK_CUDA_CHECK(cudaMalloc(&S->nccl_halo_lo, sizeof(double) * (size_t)plane));
            ATK_CUDA_CHECK(cudaMalloc(&S->nccl_halo_hi, sizeof(double) * (size_t)plane));
            ATK_CUDA_CHECK(cudaMemset(S->nccl_halo_lo, 0, sizeof(double) * (size_t)plane));
            ATK_CUDA_CHECK(cudaMemset(S->nccl_halo_hi, 0, sizeof(double) * (size_t)plane));
            S->ph.setup(ctx, (size_t)plane, (size_t)plane);  // collective; no-op unless the GPUs can map each other
            ATK_CUDA_CHECK(cudaMalloc(&S->fsend, sizeof(double) * 4 * (size_t)plane));
            if (ctx->p2p) {
                S->frecv = static_cast<double*>(ctx->arena_take(sizeof(double) * 6 * (size_t)plane));
                S->arena_exported = true;
                ATK_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
            } else {
                ATK_CUDA_CHECK(cudaMalloc(&S->frecv, sizeof(double) * 6 * (size_t)plane));
                ATK_CUDA_CHECK(cudaMemset(S->frecv, 0, sizeof(double) * 6 * (size_t)plane));
            }
            if (ctx->p2p) {  // collective: map the z-neighbours' halo arenas
                std::vector<void*> peers;
                std::vector<int> want;
                if (ctx->rank > 0) want.push_back(ctx->rank - 1);
                if (ctx->rank + 1 < ctx->world) want.push_back(ctx->rank + 1);
                ipc_share(ctx, S->frecv, peers, want);
                if (ctx->rank > 0) S->nb_lo_arena = static_cast<double*>(peers[(size_t)ctx->rank - 1]);
                if (ctx->rank + 1 < ctx->world) S->nb_hi_arena = static_cast<double*>(peers[(size_t)ctx->rank + 1]);
            }
        }
    } catch (...) {
        delete S;
        throw;
    }
    return S;
}

Operator* make_stencil27(Context* ctx, int nx, int ny, int nz, const double* coef) {
    ATK_REQUIRE(nx > 0 && ny > 0 && nz > 0 && coef, ATK_ERR_INVALID_ARGUMENT, "bad 27-point operator arguments");
    ATK_REQUIRE(nz >= ctx->world, ATK_ERR_INVALID_ARGUMENT, "fewer z-planes than ranks");
    auto* S = new Stencil27();
    S->ctx = ctx;
    int k_begin = 0, k_end = 0;
    slab_partition(nz, ctx->rank, ctx->world, &k_begin, &k_end);
    S->g = StencilGeom{nx, ny, nz, k_begin, k_end - k_begin, 0.0, 0.0, 0.0, 0.0};
    for (int t = 0; t < 27; ++t) S->cf.v[t] = coef[t];
    if (const char* e = std::getenv("ATK_STENCIL27_KC")) S->kc27 = std::max(1, std::atoi(e));
    const int64_t plane = (int64_t)nx * ny;
    S->n_rows_local = plane * (k_end - k_begin);
    S->n_cols = plane * nz;
    S->row_begin = plane * k_begin;
    // stored entries: for every non-zero coefficient d, the number of grid points whose neighbour +d is inside the grid
    int64_t nnz = 0;
    for (int dk = -1; dk <= 1; ++dk)
        for (int dj = -1; dj <= 1; ++dj)
            for (int di = -1; di <= 1; ++di)
                if (coef[(dk + 1) * 9 + (dj + 1) * 3 + (di + 1)] != 0.0)
                    nnz += (int64_t)(nx - std::abs(di)) * (ny - std::abs(dj)) * (nz - std::abs(dk));
    S->nnz = nnz;
    try {
        if (ctx->world > 1) {
            ATK_CUDA_CHECK(cudaMalloc(&S->nccl_halo_lo, sizeof(double) * (size_t)plane));
            ATK_CUDA_CHECK(cudaMalloc(&S->nccl_halo_hi, sizeof(double) * (size_t)plane));
            ATK_CUDA_CHECK(cudaMemset(S->nccl_halo_lo, 0, sizeof(double) * (size_t)plane));
            ATK_CUDA_CHECK(cudaMemset(S->nccl_halo_hi, 0, sizeof(double) * (size_t)plane));
            S->ph.setup(ctx, (size_t)plane, (size_t)plane);
        }
    } catch (...) {
        delete S;
        throw;
    }
    return S;
}

}  // namespace atk
