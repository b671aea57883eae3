// atk_cg.cu -- ConjugateGrad::solve on the device (reference src/LinearAlgebra/Krylov/ConjugateGradient.hpp:36-84).
//
// Reference loop (no tolerance test - it is commented out at :44-47,75-77):
//     rs_old = r.r ; if ||b|| < 1e-12 return
//     for i < maxIter:  Ap = A p ; pAp = p.Ap ; if |pAp| < 1e-12 return
//                       alpha = rs_old/pAp ; x = x + p*alpha ; r = r - Ap*alpha
//                       rs_new = r.r ; p = r + p*(rs_new/rs_old) ; rs_old = rs_new
//
// Device formulation: three kernels per iteration, no host round trip.
//   K1  operator apply fused with the p.Ap partial sums                (stencil: 16 B/row, CSR: +12 B/nnz)
//   K2  x += p*alpha ; r -= Ap*alpha ; partial sums of r.r              (48 B/row)
//   K3  p = r + p*beta                                                  (24 B/row)
// alpha, beta and the |pAp| < 1e-12 exit are evaluated ON THE DEVICE from the reduction slots (every CTA derives
// the same scalars in its prologue; the CTA that finishes last commits the state).  Once the exit fires, a device
// flag turns every later kernel into a no-op, so the host can enqueue iterations ahead and still stop at exactly
// the reference's iteration.  The iteration is captured once in a CUDA graph and replayed.
//
// Element arithmetic keeps the reference's two roundings (`p*alpha` is rounded, then added; VectorObj.hpp:64-68,
// 89-102).  Only the summation order of the two dot products differs (block tree vs. sequential), unless the
// context is in ATK_REDUCE_SEQUENTIAL mode.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "atk_internal.cuh"

namespace atk {

namespace {

using CgState = CgFusedState;  // one device-resident scalar block for both formulations (atk_internal.cuh)

// rs_old = r.r and the ||b|| < 1e-12 test (:37-42).  One CTA; the two dots were left in the slots by dot kernels.
__global__ void cg_begin_kernel(CgState* st, const double* slot_rr, const double* slot_bb, int world,
                                unsigned long long epoch) {
    if (threadIdx.x != 0) return;
    const double rr = slot_sum(slot_rr, world);
    const double bb = slot_sum(slot_bb, world);
    st->rs_old = rr;
    st->b_norm2 = bb;
    st->last_pAp = 0.0;
    st->it = 0;
    st->stopped_on_pAp = 0;
    st->alpha = 0.0;
    st->ka_count = 0;
    st->first = 1;
    st->epoch = epoch;
    st->peer_timeout = 0;
    const int zero = (sqrt(bb) < 1e-12) ? 1 : 0;
    st->zero_rhs = zero;
    st->done = zero;
}

// K2: x = x + p*alpha ; r = r - Ap*alpha ; rs_new partial.   (:59-63)
template <bool VEC2>
__global__ void __launch_bounds__(kReduceBlock) cg_update_xr_kernel(int64_t n, CgState* st, const double* __restrict__ slot_pAp,
                                                                    int world, double* __restrict__ x, double* __restrict__ r,
                                                                    const double* __restrict__ p, const double* __restrict__ Ap,
                                                                    double* partials, unsigned* ticket, double* rr_out,
                                                                    double* pAp_trace) {
    if (st->done) return;
    const double pAp = slot_sum(slot_pAp, world);
    if (fabs(pAp) < 1e-12) {  // :53-57 - leave x, r, p untouched; benign identical write from every CTA's thread 0
        if (threadIdx.x == 0 && blockIdx.x == 0) {
            st->last_pAp = pAp;
            if (pAp_trace && st->it < st->trace_cap) pAp_trace[st->it] = pAp;
            st->stopped_on_pAp = 1;
            __threadfence();
            st->done = 1;
        }
        return;
    }
    const double alpha = __ddiv_rn(st->rs_old, pAp);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double acc = 0.0;
    if constexpr (VEC2) {
        const int64_t n2 = n >> 1;
        double2* x2 = reinterpret_cast<double2*>(x);
        double2* r2 = reinterpret_cast<double2*>(r);
        const double2* p2 = reinterpret_cast<const double2*>(p);
        const double2* a2 = reinterpret_cast<const double2*>(Ap);
        double acc1 = 0.0;
        for (int64_t i = t; i < n2; i += stride) {
            double2 xv = x2[i], rv = r2[i];
            const double2 pv = p2[i], av = a2[i];
            xv.x = add_rn(xv.x, mul_rn(pv.x, alpha));
            xv.y = add_rn(xv.y, mul_rn(pv.y, alpha));
            rv.x = sub_rn(rv.x, mul_rn(av.x, alpha));
            rv.y = sub_rn(rv.y, mul_rn(av.y, alpha));
            x2[i] = xv;
            r2[i] = rv;
            acc = add_rn(acc, mul_rn(rv.x, rv.x));
            acc1 = add_rn(acc1, mul_rn(rv.y, rv.y));
        }
        acc = add_rn(acc, acc1);
        if (t == 0 && (n & 1)) {
            const int64_t i = n - 1;
            x[i] = add_rn(x[i], mul_rn(p[i], alpha));
            const double rn = sub_rn(r[i], mul_rn(Ap[i], alpha));
            r[i] = rn;
            acc = add_rn(acc, mul_rn(rn, rn));
        }
    } else {
        for (int64_t i = t; i < n; i += stride) {
            x[i] = add_rn(x[i], mul_rn(p[i], alpha));
            const double rn = sub_rn(r[i], mul_rn(Ap[i], alpha));
            r[i] = rn;
            acc = add_rn(acc, mul_rn(rn, rn));
        }
    }
    double total;
    if (grid_sum_last_block<kReduceBlock>(acc, partials, ticket, &total)) {
        if (threadIdx.x == 0) {
            *rr_out = total;
            st->last_pAp = pAp;
            if (pAp_trace && st->it < st->trace_cap) pAp_trace[st->it] = pAp;
        }
    }
}

// Sequential-order variant of the r.r sum for ATK_REDUCE_SEQUENTIAL: K2 without the reduction, followed by the
// one-CTA sequential dot kernel (atk_vector.cu).
template <bool VEC2>
__global__ void __launch_bounds__(kReduceBlock) cg_update_xr_nosum_kernel(int64_t n, CgState* st, const double* __restrict__ slot_pAp,
                                                                          int world, double* __restrict__ x, double* __restrict__ r,
                                                                          const double* __restrict__ p, const double* __restrict__ Ap,
                                                                          double* pAp_trace) {
    if (st->done) return;
    const double pAp = slot_sum(slot_pAp, world);
    if (fabs(pAp) < 1e-12) {
        if (threadIdx.x == 0 && blockIdx.x == 0) {
            st->last_pAp = pAp;
            if (pAp_trace && st->it < st->trace_cap) pAp_trace[st->it] = pAp;
            st->stopped_on_pAp = 1;
            __threadfence();
            st->done = 1;
        }
        return;
    }
    const double alpha = __ddiv_rn(st->rs_old, pAp);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        x[i] = add_rn(x[i], mul_rn(p[i], alpha));
        r[i] = sub_rn(r[i], mul_rn(Ap[i], alpha));
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        st->last_pAp = pAp;  // racing readers only read rs_old / done / it
        if (pAp_trace && st->it < st->trace_cap) pAp_trace[st->it] = pAp;
    }
}

// K3: p = r + p*(rs_new/rs_old) ; rs_old = rs_new.   (:79-80)
template <bool VEC2>
__global__ void __launch_bounds__(kReduceBlock) cg_update_p_kernel(int64_t n, CgState* st, const double* __restrict__ slot_rr,
                                                                   int world, const double* __restrict__ r, double* __restrict__ p,
                                                                   unsigned* ticket, double* rs_trace) {
    if (st->done) return;
    const double rs_new = slot_sum(slot_rr, world);
    const double beta = __ddiv_rn(rs_new, st->rs_old);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if constexpr (VEC2) {
        const int64_t n2 = n >> 1;
        const double2* r2 = reinterpret_cast<const double2*>(r);
        double2* p2 = reinterpret_cast<double2*>(p);
        for (int64_t i = t; i < n2; i += stride) {
            const double2 rv = r2[i];
            double2 pv = p2[i];
            pv.x = add_rn(rv.x, mul_rn(pv.x, beta));
            pv.y = add_rn(rv.y, mul_rn(pv.y, beta));
            p2[i] = pv;
        }
        if (t == 0 && (n & 1)) p[n - 1] = add_rn(r[n - 1], mul_rn(p[n - 1], beta));
    } else {
        for (int64_t i = t; i < n; i += stride) p[i] = add_rn(r[i], mul_rn(p[i], beta));
    }
    if (last_block_done(ticket)) {
        if (threadIdx.x == 0) {
            if (rs_trace && st->it < st->trace_cap) rs_trace[st->it] = rs_new;
            st->rs_old = rs_new;
            st->it = st->it + 1;
        }
    }
}

// ---------------------------------------------------------------------------------------------- fused formulation
// Peer-memory mode of KB / KF (Context::p2p): partial sums and halo planes travel by stores into the other GPUs'
// memory; sequence flags replace the NCCL all-gather.  `plane` doubles per xy-plane (even).
struct PeerKb {
    int on = 0;
    int rank = 0;
    const unsigned long long* ll = nullptr;
    unsigned long long* const* peer_ll = nullptr;
    double* lo_nb_r_hi = nullptr;
    double* hi_nb_r_lo = nullptr;
    int64_t plane = 0;
};

// KB: alpha = rs_old/pAp ; r = r - Ap*alpha ; partial sums of r.r   (:53-63 without the x update, which KA applies
// one iteration later).  24 B/row.
template <bool VEC2>
__global__ void __launch_bounds__(kReduceBlock) cg_fused_update_r_kernel(int64_t n, CgState* st, const double* __restrict__ slot_pAp,
                                                                         int world, double* __restrict__ r,
                                                                         const double* __restrict__ Ap, double* partials,
                                                                         unsigned* ticket, double* rr_out, double* pAp_trace,
                                                                         PeerKb pk) {
    if (st->done || st->peer_timeout) return;
    // peer mode: wait until every rank's p.Ap partial of this iteration has been stored into our packed words
    const double pAp = pk.on ? peer_wait_sum(pk.ll, world, kSlotPAp, st->epoch + (unsigned long long)st->ka_count, &st->peer_timeout)
                             : slot_sum(slot_pAp, world);
    if (pk.on && *reinterpret_cast<volatile int*>(&st->peer_timeout)) return;
    if (fabs(pAp) < 1e-12) {  // :53-57
        if (threadIdx.x == 0 && blockIdx.x == 0) {
            st->last_pAp = pAp;
            if (pAp_trace && st->it < st->trace_cap) pAp_trace[st->it] = pAp;
            st->stopped_on_pAp = 1;
            __threadfence();
            st->done = 1;
        }
        return;
    }
    const double alpha = __ddiv_rn(st->rs_old, pAp);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double acc = 0.0;
    if constexpr (VEC2) {
        const int64_t n2 = n >> 1;
        double2* r2 = reinterpret_cast<double2*>(r);
        const double2* a2 = reinterpret_cast<const double2*>(Ap);
        double acc1 = 0.0;
        const int64_t plane2 = pk.plane >> 1;
        for (int64_t i = t; i < n2; i += stride) {
            double2 rv = r2[i];
            const double2 av = a2[i];
            rv.x = sub_rn(rv.x, mul_rn(av.x, alpha));
            rv.y = sub_rn(rv.y, mul_rn(av.y, alpha));
            r2[i] = rv;
            if (pk.on) {  // mirror the first / last plane of the new r into the neighbours' halo planes (NVLink stores)
                if (i < plane2 && pk.lo_nb_r_hi) reinterpret_cast<double2*>(pk.lo_nb_r_hi)[i] = rv;
                if (i >= n2 - plane2 && pk.hi_nb_r_lo) reinterpret_cast<double2*>(pk.hi_nb_r_lo)[i - (n2 - plane2)] = rv;
            }
            acc = add_rn(acc, mul_rn(rv.x, rv.x));
            acc1 = add_rn(acc1, mul_rn(rv.y, rv.y));
        }
        acc = add_rn(acc, acc1);
        if (t == 0 && (n & 1)) {
            const double rn = sub_rn(r[n - 1], mul_rn(Ap[n - 1], alpha));
            r[n - 1] = rn;
            acc = add_rn(acc, mul_rn(rn, rn));
        }
    } else {
        for (int64_t i = t; i < n; i += stride) {
            const double rn = sub_rn(r[i], mul_rn(Ap[i], alpha));
            r[i] = rn;
            acc = add_rn(acc, mul_rn(rn, rn));
        }
    }
    if (pk.on) __threadfence_system();  // each thread orders its peer stores before its CTA's ticket
    double total;
    if (grid_sum_last_block<kReduceBlock>(acc, partials, ticket, &total)) {
        if (pk.on) __threadfence_system();  // system-scope release by the publishing threads (see KA)
        if (pk.on && threadIdx.x < 32)
            publish_partial_warp(pk.peer_ll, world, pk.rank, kSlotRR, total, st->epoch + (unsigned long long)st->ka_count);
        if (threadIdx.x == 0) {
            if (!pk.on) *rr_out = total;
            st->alpha = alpha;
            st->last_pAp = pAp;
            if (pAp_trace && st->it < st->trace_cap) pAp_trace[st->it] = pAp;
        }
    }
}

// KF, once after the loop: brings the caller's x and p to the reference's state.
//   exit fired (done): x and r are complete; the current direction is copied into the caller's p if it lives in the
//                      scratch buffer;
//   loop ran out:      the last iteration's x = x + p*alpha and p = r + p*beta are still pending - apply them.
__global__ void __launch_bounds__(kReduceBlock) cg_fused_flush_kernel(int64_t n, CgState* st, const double* __restrict__ slot_rr,
                                                                      int world, const double* __restrict__ r, double* x,
                                                                      const double* P0, const double* P1, double* p_user,
                                                                      unsigned* ticket, double* rs_trace, PeerKb pk) {
    if (st->first || st->peer_timeout) return;  // no KA ran: nothing pending, p_user untouched
    double rs_peer = 0.0;
    if (pk.on && !st->done) {
        rs_peer = peer_wait_sum(pk.ll, world, kSlotRR, st->epoch + (unsigned long long)st->ka_count, &st->peer_timeout);
        if (*reinterpret_cast<volatile int*>(&st->peer_timeout)) return;
    }
    const double* pc = (st->ka_count & 1) ? P1 : P0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (st->done) {
        if (pc != p_user)
            for (int64_t i = t; i < n; i += stride) p_user[i] = pc[i];
        return;
    }
    const double rs_new = pk.on ? rs_peer : slot_sum(slot_rr, world);
    const double beta = __ddiv_rn(rs_new, st->rs_old);
    const double alpha = st->alpha;
    if (((n & 1) == 0) && ((reinterpret_cast<uintptr_t>(pc) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(r) |
                            reinterpret_cast<uintptr_t>(p_user)) & 15u) == 0) {
        // 16-byte accesses, four independent element pairs in flight per thread (40 B/row: read x, p, r; write x, p)
        const int64_t n2 = n >> 1;
        const double2* pc2 = reinterpret_cast<const double2*>(pc);
        const double2* r2 = reinterpret_cast<const double2*>(r);
        double2* x2 = reinterpret_cast<double2*>(x);
        double2* pu2 = reinterpret_cast<double2*>(p_user);
        auto one = [&](int64_t i, const double2 pv, const double2 rv, double2 xv) {
            xv.x = add_rn(xv.x, mul_rn(pv.x, alpha));
            xv.y = add_rn(xv.y, mul_rn(pv.y, alpha));
            x2[i] = xv;
            pu2[i] = make_double2(add_rn(rv.x, mul_rn(pv.x, beta)), add_rn(rv.y, mul_rn(pv.y, beta)));
        };
        int64_t i = t;
        for (; i + 3 * stride < n2; i += 4 * stride) {
            const double2 p0 = pc2[i], p1 = pc2[i + stride], p2 = pc2[i + 2 * stride], p3 = pc2[i + 3 * stride];
            const double2 r0 = r2[i], r1 = r2[i + stride], r2v = r2[i + 2 * stride], r3 = r2[i + 3 * stride];
            const double2 x0 = x2[i], x1 = x2[i + stride], x2v = x2[i + 2 * stride], x3 = x2[i + 3 * stride];
            one(i, p0, r0, x0);
            one(i + stride, p1, r1, x1);
            one(i + 2 * stride, p2, r2v, x2v);
            one(i + 3 * stride, p3, r3, x3);
        }
        for (; i < n2; i += stride) one(i, pc2[i], r2[i], x2[i]);
    } else {
        for (int64_t i = t; i < n; i += stride) {
            const double pv = pc[i];
            x[i] = add_rn(x[i], mul_rn(pv, alpha));
            p_user[i] = add_rn(r[i], mul_rn(pv, beta));
        }
    }
    if (last_block_done(ticket)) {
        if (threadIdx.x == 0) {
            if (rs_trace && st->it < st->trace_cap) rs_trace[st->it] = rs_new;
            st->rs_old = rs_new;
            st->it = st->it + 1;
        }
    }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

void cg_solve(Context* ctx, Operator* op, const double* b, double* x, double* r, double* p, int max_iter,
              atk_cg_info* info) {
    ATK_REQUIRE(op && b && x && r && p && info, ATK_ERR_INVALID_ARGUMENT, "null argument to cg_solve");
    ATK_NVTX("atk::cg_solve");
    ATK_REQUIRE(op->n_cols == op->n_rows_local * 1 || ctx->world > 1, ATK_ERR_INVALID_ARGUMENT, "Dimension mismatch");
    const int64_t n = op->n_rows_local;
    cudaStream_t st = ctx->stream;
    const int64_t launches0 = ctx->launches;
    const int trace_cap = (info->pAp_trace || info->rs_trace) ? std::max(0, info->trace_capacity) : 0;

    CgState* d_state = nullptr;
    double *Ap = nullptr, *d_ptrace = nullptr, *d_rtrace = nullptr, *p_alt = nullptr;
    const char* env_fused = std::getenv("ATK_CG_FUSED");
    const bool fused = op->fused_cg_capable() && ctx->reduction == ATK_REDUCE_TREE && !(env_fused && env_fused[0] == '0');
    info->fused = fused ? 1 : 0;
    // small single-GPU systems: the whole loop as one cooperative launch with the vectors in registers
    const bool onchip = fused && ctx->world == 1 && !info->time_kernels && max_iter > 0 && op->cg_onchip_blocks() > 0;
    // peer-memory mode and the persistent one-launch loop: every rank must agree (they change who waits for whom)
    bool peer = false;
    const bool all_aligned = (n % 2 == 0) && aligned16(x) && aligned16(r) && aligned16(p) && aligned16(b);
    // Persistent one-launch loop: the default on distributed contexts, where it removes two kernel boundaries, two
    // last-CTA elections and every NCCL call from the iteration.  On one GPU the CUDA-graph loop of two kernels is ~1 %
    // faster at 512^3 (the hardware CTA scheduler balances 8192 small CTAs better than a static split of 1024 tiles over
    // 148 SMs), so there it is opt-in: ATK_CG_PERSISTENT=1 / 0 forces either.
    const char* env_persistent = std::getenv("ATK_CG_PERSISTENT");
    const bool want_persistent = env_persistent ? env_persistent[0] != '0' : ctx->world > 1;
    bool persistent = fused && !onchip && max_iter > 0 && all_aligned && want_persistent && op->cg_persistent_blocks() > 0;
    if (fused && ctx->world > 1) {
        const char* force_nccl = std::getenv("ATK_CG_FORCE_NCCL");
        const int peer_ok = op->peer_capable() && (n % 2 == 0) && aligned16(x) && aligned16(r) && aligned16(p) &&
                            !(force_nccl && force_nccl[0] == '1');
        int mine = (peer_ok ? 1 : 0) | (peer_ok && persistent ? 2 : 0);
        std::vector<int> all((size_t)ctx->world);
        allgather_host_bytes(ctx, &mine, all.data(), sizeof(int));
        int agreed = 3;
        for (int v : all) agreed &= v;
        peer = (agreed & 1) != 0;
        persistent = (agreed & 2) != 0;
    }
    op->set_peer_mode(peer);
    info->peer = peer ? 1 : 0;
    info->persistent = persistent ? 1 : 0;
    PeerKb pk;
    if (peer) {
        pk.on = 1;
        pk.rank = ctx->rank;
        pk.ll = ctx->d_ll;
        pk.peer_ll = ctx->d_peer_ll;
        op->peer_r_targets(&pk.lo_nb_r_hi, &pk.hi_nb_r_lo, &pk.plane);
    }
    int* h_done = nullptr;  // pinned: done flag of every poll
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t gexec = nullptr;
    std::vector<cudaEvent_t> poll_events;
    auto cleanup = [&]() {
        if (gexec) cudaGraphExecDestroy(gexec);
        if (graph) cudaGraphDestroy(graph);
        for (auto e : poll_events) if (e) cudaEventDestroy(e);
        ctx->ws_release(d_state); ctx->ws_release(Ap); ctx->ws_release(d_ptrace); ctx->ws_release(d_rtrace); ctx->ws_release(p_alt);
        if (h_done && h_done != ctx->h_flags) cudaFreeHost(h_done);
    };
    try {
        d_state = static_cast<CgState*>(ctx->ws_alloc(sizeof(CgState)));
        Ap = static_cast<double*>(ctx->ws_alloc(sizeof(double) * std::max<int64_t>(n, 1)));
        if (fused) p_alt = static_cast<double*>(ctx->ws_alloc(sizeof(double) * std::max<int64_t>(n, 1)));
        if (trace_cap > 0) {
            d_ptrace = static_cast<double*>(ctx->ws_alloc(sizeof(double) * trace_cap));
            d_rtrace = static_cast<double*>(ctx->ws_alloc(sizeof(double) * trace_cap));
            ATK_CUDA_CHECK(cudaMemsetAsync(d_ptrace, 0xff, sizeof(double) * trace_cap, st));  // NaN fill
            ATK_CUDA_CHECK(cudaMemsetAsync(d_rtrace, 0xff, sizeof(double) * trace_cap, st));
        }
        CgState h0{};
        h0.trace_cap = trace_cap;
        ATK_CUDA_CHECK(cudaMemcpyAsync(d_state, &h0, sizeof(CgState), cudaMemcpyHostToDevice, st));

        const bool vec = aligned16(x) && aligned16(r) && aligned16(p) && aligned16(Ap);
        const int64_t work = vec ? (n + 1) / 2 : n;
        const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((work + kReduceBlock - 1) / kReduceBlock, ctx->reduce_grid()));
        const bool seq = ctx->reduction == ATK_REDUCE_SEQUENTIAL;
        const int* done_flag = &d_state->done;

        std::vector<cudaEvent_t> kev;  // time_kernels mode: 4 events per iteration
        auto mark = [&]() {
            if (!info->time_kernels) return;
            cudaEvent_t e;
            ATK_CUDA_CHECK(cudaEventCreate(&e));
            ATK_CUDA_CHECK(cudaEventRecord(e, st));
            kev.push_back(e);
        };
        auto enqueue_fused = [&]() {
            mark();
            op->cg_fused_step(d_state, r, x, Ap, p, p_alt, kSlotRR, kSlotPAp, d_rtrace);  // KA (+ all-gather of p.Ap)
            mark();
            if (vec)
                cg_fused_update_r_kernel<true><<<grid, kReduceBlock, 0, st>>>(n, d_state, ctx->slot(kSlotPAp), ctx->world, r, Ap,
                                                                              ctx->d_partials, ctx->d_tickets + 3,
                                                                              ctx->slot(kSlotRR) + ctx->rank, d_ptrace, pk);
            else
                cg_fused_update_r_kernel<false><<<grid, kReduceBlock, 0, st>>>(n, d_state, ctx->slot(kSlotPAp), ctx->world, r, Ap,
                                                                               ctx->d_partials, ctx->d_tickets + 3,
                                                                               ctx->slot(kSlotRR) + ctx->rank, d_ptrace, pk);
            count_launch(ctx);
            if (!peer) allgather_slot(ctx, kSlotRR, st);
            mark();
            mark();
        };
        auto enqueue_iteration = [&]() {
            if (fused) return enqueue_fused();
            mark();
            op->apply(p, Ap, seq ? -1 : kSlotPAp, done_flag);  // K1 (+ all-gather of the partial)
            if (seq) launch_dot(ctx, n, p, Ap, kSlotPAp, done_flag);  // the fused tree sum is replaced by an ordered sum
            mark();
            if (seq) {
                cg_update_xr_nosum_kernel<false><<<grid, kReduceBlock, 0, st>>>(n, d_state, ctx->slot(kSlotPAp), ctx->world, x, r,
                                                                                p, Ap, d_ptrace);
                count_launch(ctx);
                launch_dot(ctx, n, r, r, kSlotRR, done_flag);
            } else {
                if (vec)
                    cg_update_xr_kernel<true><<<grid, kReduceBlock, 0, st>>>(n, d_state, ctx->slot(kSlotPAp), ctx->world, x, r, p,
                                                                             Ap, ctx->d_partials, ctx->d_tickets + 3,
                                                                             ctx->slot(kSlotRR) + ctx->rank, d_ptrace);
                else
                    cg_update_xr_kernel<false><<<grid, kReduceBlock, 0, st>>>(n, d_state, ctx->slot(kSlotPAp), ctx->world, x, r, p,
                                                                              Ap, ctx->d_partials, ctx->d_tickets + 3,
                                                                              ctx->slot(kSlotRR) + ctx->rank, d_ptrace);
                count_launch(ctx);
                allgather_slot(ctx, kSlotRR, st);
            }
            mark();
            if (vec)
                cg_update_p_kernel<true><<<grid, kReduceBlock, 0, st>>>(n, d_state, ctx->slot(kSlotRR), ctx->world, r, p,
                                                                        ctx->d_tickets + 4, d_rtrace);
            else
                cg_update_p_kernel<false><<<grid, kReduceBlock, 0, st>>>(n, d_state, ctx->slot(kSlotRR), ctx->world, r, p,
                                                                         ctx->d_tickets + 4, d_rtrace);
            count_launch(ctx);
            mark();
        };

        // ---- capture one iteration, replay it
        int launches_per_iter = 0;
        const bool use_graph = !info->time_kernels;
        if (max_iter > 0 && use_graph && !onchip && !persistent) {
            const int64_t l0 = ctx->launches;
            ATK_CUDA_CHECK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            try {
                enqueue_iteration();
            } catch (...) {
                cudaGraph_t g2 = nullptr;
                cudaStreamEndCapture(st, &g2);
                if (g2) cudaGraphDestroy(g2);
                throw;
            }
            ATK_CUDA_CHECK(cudaStreamEndCapture(st, &graph));
            ATK_CUDA_CHECK(cudaGraphInstantiate(&gexec, graph, 0));
            launches_per_iter = (int)(ctx->launches - l0);
            ctx->launches = l0;  // capture did not launch anything
        }

        // completion tickets wrap back to zero by themselves; clear them anyway in case an earlier solve was aborted
        ATK_CUDA_CHECK(cudaMemsetAsync(ctx->d_tickets, 0, sizeof(unsigned) * kNumTickets, st));
        // persistent mode: phase stamps for the per-phase times (time_kernels) - [cap] stamps + their count
        unsigned long long* d_stamps = nullptr;
        int stamp_cap = 0;
        if (persistent && info->time_kernels) {
#ifdef ATK_REDUCE_PROFILE
            stamp_cap = 5 * (2 * std::min(max_iter, 1 << 12) + 2) + 1;  // five stations per reduction (profile build)
#else
            stamp_cap = 2 * std::min(max_iter, 1 << 15) + 1;
#endif
            d_stamps = static_cast<unsigned long long*>(ctx->ws_alloc(sizeof(unsigned long long) * ((size_t)stamp_cap + 1)));
            ATK_CUDA_CHECK(cudaMemsetAsync(d_stamps, 0, sizeof(unsigned long long) * ((size_t)stamp_cap + 1), st));
        }
        struct StampFree { Context* c; unsigned long long*& p; ~StampFree() { if (p) c->ws_release(p); } } stamp_free{ctx, d_stamps};
        ATK_CUDA_CHECK(cudaEventRecord(ctx->ev_t0, st));
        if (persistent) {
            // the start-up sums, the initial halo planes, every iteration and the final pending updates: one launch
            op->cg_persistent(d_state, b, x, r, Ap, p, p_alt, max_iter, d_ptrace, d_rtrace, trace_cap, ctx->epoch, d_stamps, stamp_cap);
            ctx->epoch += 2ull * (unsigned long long)max_iter + 8ull;  // same advance on every rank
        } else {
        // rs_old = r.r ; ||b||   (:37-38)
        launch_dot(ctx, n, r, r, kSlotRR);
        launch_dot(ctx, n, b, b, kSlotBnorm);
        cg_begin_kernel<<<1, 32, 0, st>>>(d_state, ctx->slot(kSlotRR), ctx->slot(kSlotBnorm), ctx->world, ctx->epoch);
        count_launch(ctx);
        if (peer) op->initial_halo_exchange(r, p);  // boundary planes of the incoming r and p, once, through NCCL
        }

        // ---- run: iterations are enqueued in batches; the exit flag is read back after every batch and examined
        // two batches later (a blocking wait on that batch's event, so every rank of a distributed run takes the
        // same decision after the same batch and enqueues the same number of collectives).
        if (onchip) op->cg_onchip(d_state, x, r, p, p_alt, max_iter, d_ptrace, d_rtrace);
        constexpr int kBatch = 16, kLag = 2;
        const int n_batches = (onchip || persistent) ? 0 : (max_iter + kBatch - 1) / kBatch;
        if (n_batches <= kHostFlags) h_done = ctx->h_flags;
        else ATK_CUDA_CHECK(cudaMallocHost(&h_done, sizeof(int) * (size_t)n_batches));
        poll_events.resize(std::max(n_batches, 1), nullptr);
        int enq = 0;
        for (int bt = 0; bt < n_batches; ++bt) {
            if (bt >= kLag) {
                ATK_CUDA_CHECK(cudaEventSynchronize(poll_events[bt - kLag]));
                if (h_done[bt - kLag]) break;
            }
            const int cnt = std::min(kBatch, max_iter - enq);
            for (int q = 0; q < cnt; ++q) {
                if (use_graph) {
                    ATK_CUDA_CHECK(cudaGraphLaunch(gexec, st));
                    ctx->launches += launches_per_iter;
                } else {
                    enqueue_iteration();
                }
            }
            enq += cnt;
            ATK_CUDA_CHECK(cudaMemcpyAsync(h_done + bt, &d_state->done, sizeof(int), cudaMemcpyDeviceToHost, st));
            ATK_CUDA_CHECK(cudaEventCreateWithFlags(&poll_events[bt], cudaEventDisableTiming));
            ATK_CUDA_CHECK(cudaEventRecord(poll_events[bt], st));
        }
        if (fused && !onchip && !persistent) {  // KF: pending x / p updates of the last iteration, or p copied back after an exit
            cg_fused_flush_kernel<<<ctx->reduce_grid(), kReduceBlock, 0, st>>>(n, d_state, ctx->slot(kSlotRR), ctx->world, r, x, p,
                                                                               p_alt, p, ctx->d_tickets + 4, d_rtrace, pk);
            count_launch(ctx);
        }
        if (peer && !persistent) ctx->epoch += (unsigned long long)std::max(max_iter, 0) + 4ull;  // same advance on every rank
        ATK_CUDA_CHECK(cudaEventRecord(ctx->ev_t1, st));

        CgState hs{};
        ATK_CUDA_CHECK(cudaMemcpyAsync(&hs, d_state, sizeof(CgState), cudaMemcpyDeviceToHost, st));
        if (trace_cap > 0) {
            if (info->pAp_trace)
                ATK_CUDA_CHECK(cudaMemcpyAsync(info->pAp_trace, d_ptrace, sizeof(double) * trace_cap, cudaMemcpyDeviceToHost, st));
            if (info->rs_trace)
                ATK_CUDA_CHECK(cudaMemcpyAsync(info->rs_trace, d_rtrace, sizeof(double) * trace_cap, cudaMemcpyDeviceToHost, st));
        }
        std::vector<unsigned long long> h_stamps;
        if (d_stamps) {
            h_stamps.resize((size_t)stamp_cap + 1);
            ATK_CUDA_CHECK(cudaMemcpyAsync(h_stamps.data(), d_stamps, sizeof(unsigned long long) * h_stamps.size(), cudaMemcpyDeviceToHost, st));
        }
        ATK_CUDA_CHECK(cudaStreamSynchronize(st));
        float ms = 0.f;
        ATK_CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->ev_t0, ctx->ev_t1));
        info->kernel_ms[0] = info->kernel_ms[1] = info->kernel_ms[2] = 0.f;
#ifdef ATK_REDUCE_PROFILE
        if (!h_stamps.empty() && info->rs_trace && info->trace_capacity > 0) {
            // profile build: the raw station stamps (ns since the first one) go back through rs_trace, five per reduction,
            // starting with the two start-up reductions; tools/latency_breakdown.py turns them into averages
            const unsigned long long t0s = h_stamps[1];
            for (int q = 0; q < info->trace_capacity && q + 1 < stamp_cap; ++q)
                info->rs_trace[q] = h_stamps[(size_t)q + 1] ? (double)(h_stamps[(size_t)q + 1] - t0s) : -1.0;
        }
#else
        if (!h_stamps.empty()) {  // [0] end of start-up, [1+2i] after p.Ap of iteration i, [2+2i] after its r.r
            const int cnt = (int)h_stamps[(size_t)stamp_cap];
            for (int q = 0; q + 1 < cnt; ++q) info->kernel_ms[q & 1] += (float)((double)(h_stamps[(size_t)q + 1] - h_stamps[(size_t)q]) * 1e-6);
        }
#endif
        for (size_t q = 0; q + 3 < kev.size(); q += 4)
            for (int c = 0; c < 3; ++c) {
                float t = 0.f;
                ATK_CUDA_CHECK(cudaEventElapsedTime(&t, kev[q + c], kev[q + c + 1]));
                info->kernel_ms[c] += t;
            }
        for (auto e : kev) cudaEventDestroy(e);
        kev.clear();
        op->set_peer_mode(false);
        ATK_REQUIRE(!hs.peer_timeout, ATK_ERR_RUNTIME,
                    "peer-memory CG: a wait for another rank's partial sum timed out (rank crashed or out of step)");
        info->iterations = hs.it;
        info->stopped_on_pAp = hs.stopped_on_pAp;
        info->zero_rhs = hs.zero_rhs;
        info->last_pAp = hs.last_pAp;
        info->rs = hs.rs_old;
        info->device_ms = ms;
        info->kernel_launches = (int)(ctx->launches - launches0);
    } catch (...) {
        op->set_peer_mode(false);
        cleanup();
        throw;
    }
    cleanup();
}

}  // namespace atk
