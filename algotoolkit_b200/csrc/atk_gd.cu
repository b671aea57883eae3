// atk_gd.cu -- GradientDescent::solve on the device (reference src/LinearAlgebra/Solver/IterSolver.hpp:23-44), a
// "next" row of the scope table: a pure composition of the operator apply, dot and axpy kernels.
//
//     r = b - A*x ; residualNorm = ||r|| ; xOld = x
//     for iter < maxIter && residualNorm > tol:
//         Ar = A*r ; alpha = (r.r)/(r.Ar) ; x = x + r*alpha ; r = r - Ar*alpha
//         residualNorm = ||r|| ; if ||x - xOld|| < tol break ; xOld = x
//
// Device formulation per iteration: operator apply fused with r.Ar, one update kernel (x, r and the difference
// d = x_new - x_old, computed exactly as the reference forms it), two dot kernels (r.r, d.d), one single-thread kernel that
// takes the loop decisions.  As in the CG driver a device flag turns the kernels of later iterations into no-ops, so the
// host enqueues batches ahead and still stops at exactly the reference's iteration.
#include <algorithm>
#include <cmath>
#include <vector>

#include "atk_internal.cuh"

namespace atk {

namespace {

struct GdState {
    double rr;             // r.r of the current residual
    double residual_norm;
    double change_norm;
    int it;                // iterations whose update has been applied
    int done;
};

__global__ void gd_begin_kernel(GdState* st, const double* slot_rr, int world, double tol, int max_iter) {
    const double rr = slot_sum(slot_rr, world);
    st->rr = rr;
    st->residual_norm = sqrt(rr);
    st->change_norm = 0.0;
    st->it = 0;
    st->done = (max_iter > 0 && st->residual_norm > tol) ? 0 : 1;  // loop condition (:28)
}

// x = x + r*alpha ; r = r - Ar*alpha ; d = x_new - x_old   (:30-33,38)
__global__ void __launch_bounds__(256) gd_update_kernel(int64_t n, const GdState* st, const double* __restrict__ slot_rAr, int world,
                                                        double* __restrict__ x, double* __restrict__ r,
                                                        const double* __restrict__ Ar, double* __restrict__ d) {
    if (st->done) return;
    const double alpha = __ddiv_rn(st->rr, slot_sum(slot_rAr, world));
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double xo = x[i], rv = r[i];
        const double xn = add_rn(xo, mul_rn(rv, alpha));
        x[i] = xn;
        r[i] = sub_rn(rv, mul_rn(Ar[i], alpha));
        d[i] = sub_rn(xn, xo);
    }
}

__global__ void gd_decide_kernel(GdState* st, const double* slot_rr, const double* slot_dd, int world, double tol, int max_iter) {
    if (st->done) return;
    const double rr = slot_sum(slot_rr, world);
    const double ch = sqrt(slot_sum(slot_dd, world));
    st->rr = rr;
    st->residual_norm = sqrt(rr);  // :35
    st->change_norm = ch;          // :38
    st->it = st->it + 1;
    // :39-41 early exit on the change norm, else the loop condition of the next iteration (:28)
    if (ch < tol || !(st->it < max_iter && st->residual_norm > tol)) st->done = 1;
}

}  // namespace

void gd_solve(Context* ctx, Operator* op, const double* b, double* x, int max_iter, double tol, atk_gd_info* info) {
    ATK_NVTX("atk::gd_solve");
    ATK_REQUIRE(op && b && x && info, ATK_ERR_INVALID_ARGUMENT, "null argument to gd_solve");
    const int64_t n = op->n_rows_local;
    cudaStream_t st = ctx->stream;
    const int64_t launches0 = ctx->launches;
    WsBuffer r(ctx, sizeof(double) * (size_t)std::max<int64_t>(n, 1)), Ar(ctx, sizeof(double) * (size_t)std::max<int64_t>(n, 1)),
        d(ctx, sizeof(double) * (size_t)std::max<int64_t>(n, 1)), state(ctx, sizeof(GdState));
    GdState* d_state = state.as<GdState>();
    const int* done_flag = &d_state->done;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 8));
    std::vector<cudaEvent_t> events;
    auto cleanup = [&]() {
        for (auto e : events)
            if (e) cudaEventDestroy(e);
    };
    try {
        ATK_CUDA_CHECK(cudaEventRecord(ctx->ev_t0, st));
        op->apply(x, Ar.as<double>(), -1, nullptr);  // :24 r = b - A*x
        launch_sub(ctx, n, b, Ar.as<double>(), r.as<double>());
        launch_dot(ctx, n, r.as<double>(), r.as<double>(), kSlotRR);
        gd_begin_kernel<<<1, 1, 0, st>>>(d_state, ctx->slot(kSlotRR), ctx->world, tol, max_iter);
        count_launch(ctx);
        constexpr int kBatch = 16, kLag = 2;
        const int n_batches = (max_iter + kBatch - 1) / kBatch;
        ATK_REQUIRE(n_batches <= kHostFlags, ATK_ERR_INVALID_ARGUMENT, "max_iter too large");
        events.assign((size_t)std::max(n_batches, 1), nullptr);
        int enq = 0;
        for (int bt = 0; bt < n_batches; ++bt) {
            if (bt >= kLag) {
                ATK_CUDA_CHECK(cudaEventSynchronize(events[(size_t)(bt - kLag)]));
                if (ctx->h_flags[bt - kLag]) break;
            }
            const int cnt = std::min(kBatch, max_iter - enq);
            for (int q = 0; q < cnt; ++q) {
                op->apply(r.as<double>(), Ar.as<double>(), -1, done_flag);                         // :29
                launch_dot(ctx, n, r.as<double>(), Ar.as<double>(), kSlotPAp, done_flag);          // r.Ar
                gd_update_kernel<<<grid, 256, 0, st>>>(n, d_state, ctx->slot(kSlotPAp), ctx->world, x, r.as<double>(),
                                                       Ar.as<double>(), d.as<double>());
                count_launch(ctx);
                launch_dot(ctx, n, r.as<double>(), r.as<double>(), kSlotRR, done_flag);            // :35
                launch_dot(ctx, n, d.as<double>(), d.as<double>(), kSlotGeneric, done_flag);       // :38
                gd_decide_kernel<<<1, 1, 0, st>>>(d_state, ctx->slot(kSlotRR), ctx->slot(kSlotGeneric), ctx->world, tol, max_iter);
                count_launch(ctx);
            }
            enq += cnt;
            ATK_CUDA_CHECK(cudaMemcpyAsync(ctx->h_flags + bt, done_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
            ATK_CUDA_CHECK(cudaEventCreateWithFlags(&events[(size_t)bt], cudaEventDisableTiming));
            ATK_CUDA_CHECK(cudaEventRecord(events[(size_t)bt], st));
        }
        ATK_CUDA_CHECK(cudaEventRecord(ctx->ev_t1, st));
        GdState hs{};
        ATK_CUDA_CHECK(cudaMemcpyAsync(&hs, d_state, sizeof(GdState), cudaMemcpyDeviceToHost, st));
        ATK_CUDA_CHECK(cudaStreamSynchronize(st));
        float ms = 0.f;
        ATK_CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->ev_t0, ctx->ev_t1));
        info->iterations = hs.it;
        info->residual_norm = hs.residual_norm;
        info->change_norm = hs.change_norm;
        info->device_ms = ms;
        info->kernel_launches = (int)(ctx->launches - launches0);
    } catch (...) {
        cudaStreamSynchronize(st);
        cleanup();
        throw;
    }
    cleanup();
}

}  // namespace atk
