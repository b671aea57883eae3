// atk_vector.cu -- the VectorObj<double> kernels (reference src/Obj/VectorObj.hpp).
//
// All of them are pure streaming kernels bounded by HBM bandwidth:
//   dot           16 B/element (8 B for x.x), no store
//   scale / div   16 B/element
//   add/sub/axpy  24 B/element
// Elementwise kernels move 16-byte vectors per thread when the pointers allow it; every kernel uses a
// persistent grid (4 CTAs of 256 threads per SM) and a grid-stride loop, which also makes the shape of the
// reduction tree a function of the device only.
#include <algorithm>

#include "atk_internal.cuh"

namespace atk {

namespace {

// ---------------------------------------------------------------------------------- elementwise
enum class Ew { Scale, Div, Add, Sub, Axpy, Axmy };

template <Ew OP>
__device__ __forceinline__ double ew_apply(double x, double y, double a) {
    if constexpr (OP == Ew::Scale) return mul_rn(x, a);            // VectorObj.hpp:64-68
    else if constexpr (OP == Ew::Div) return __ddiv_rn(x, a);      // :75-80 (true division, not x*(1/a))
    else if constexpr (OP == Ew::Add) return add_rn(x, y);         // :89-94
    else if constexpr (OP == Ew::Sub) return sub_rn(x, y);         // :97-102
    else if constexpr (OP == Ew::Axpy) return add_rn(x, mul_rn(y, a));  // x + y*a, product rounded first
    else return sub_rn(x, mul_rn(y, a));                           // x - y*a
}

template <Ew OP>
__host__ __device__ constexpr bool ew_binary() { return OP == Ew::Add || OP == Ew::Sub || OP == Ew::Axpy || OP == Ew::Axmy; }

template <Ew OP, bool VEC2>
__global__ void __launch_bounds__(256) ew_kernel(int64_t n, const double* x, const double* y,
                                                 double a, double* out) {  // no __restrict__: out may alias x or y
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if constexpr (VEC2) {
        const int64_t n2 = n >> 1;
        const double2* x2 = reinterpret_cast<const double2*>(x);
        const double2* y2 = reinterpret_cast<const double2*>(y);
        double2* o2 = reinterpret_cast<double2*>(out);
        for (int64_t i = t; i < n2; i += stride) {
            const double2 xv = x2[i];
            double2 yv = make_double2(0.0, 0.0);
            if constexpr (ew_binary<OP>()) yv = y2[i];
            o2[i] = make_double2(ew_apply<OP>(xv.x, yv.x, a), ew_apply<OP>(xv.y, yv.y, a));
        }
        if (t == 0 && (n & 1)) {
            const int64_t i = n - 1;
            out[i] = ew_apply<OP>(x[i], ew_binary<OP>() ? y[i] : 0.0, a);
        }
    } else {
        for (int64_t i = t; i < n; i += stride) out[i] = ew_apply<OP>(x[i], ew_binary<OP>() ? y[i] : 0.0, a);
    }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <Ew OP>
void launch_ew(Context* ctx, int64_t n, const double* x, const double* y, double a, double* out) {
    if (n <= 0) return;
    const bool vec = aligned16(x) && aligned16(out) && (!ew_binary<OP>() || aligned16(y));
    const int64_t work = vec ? (n + 1) / 2 : n;
    int grid = (int)std::min<int64_t>((work + 255) / 256, (int64_t)ctx->sm_count * 8);
    if (grid < 1) grid = 1;
    if (vec) ew_kernel<OP, true><<<grid, 256, 0, ctx->stream>>>(n, x, y, a, out);
    else ew_kernel<OP, false><<<grid, 256, 0, ctx->stream>>>(n, x, y, a, out);
    ATK_CUDA_CHECK(cudaGetLastError());
    count_launch(ctx);
}

// ---------------------------------------------------------------------------------- dot (tree)
// Two-stage deterministic reduction: per-thread sequential accumulation over a grid-stride subsequence,
// fixed-shape block tree, per-block partial, fixed-order final pass by the last block.
template <bool VEC2>
__global__ void __launch_bounds__(kReduceBlock) dot_tree_kernel(int64_t n, const double* __restrict__ x,
                                                                const double* __restrict__ y, double* partials,
                                                                unsigned* ticket, double* out, const int* skip_flag) {
    if (skip_flag && *skip_flag) return;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double acc = 0.0;
    if constexpr (VEC2) {
        const int64_t n2 = n >> 1;
        const double2* x2 = reinterpret_cast<const double2*>(x);
        const double2* y2 = reinterpret_cast<const double2*>(y);
        double acc1 = 0.0;
        for (int64_t i = t; i < n2; i += stride) {
            const double2 xv = x2[i], yv = y2[i];
            acc = add_rn(acc, mul_rn(xv.x, yv.x));
            acc1 = add_rn(acc1, mul_rn(xv.y, yv.y));
        }
        acc = add_rn(acc, acc1);
        if (t == 0 && (n & 1)) acc = add_rn(acc, mul_rn(x[n - 1], y[n - 1]));
    } else {
        for (int64_t i = t; i < n; i += stride) acc = add_rn(acc, mul_rn(x[i], y[i]));
    }
    double total;
    if (grid_sum_last_block<kReduceBlock>(acc, partials, ticket, &total)) {
        if (linear_thread_id() == 0) *out = total;
    }
}

// ---------------------------------------------------------------------------------- dot (sequential)
// Bit-identical to std::inner_product(begin, end, other, 0.0): acc = acc + x[i]*y[i], i ascending.  One CTA;
// all threads form the rounded products of a 1024-element chunk in shared memory, thread 0 adds them in order.
__global__ void __launch_bounds__(1024) dot_sequential_kernel(int64_t n, const double* __restrict__ x,
                                                              const double* __restrict__ y, double* out,
                                                              const int* skip_flag) {
    if (skip_flag && *skip_flag) return;
    __shared__ double prod[2][1024];
    double acc = 0.0;
    int buf = 0;
    for (int64_t base = 0; base < n; base += 1024, buf ^= 1) {
        const int64_t i = base + threadIdx.x;
        if (i < n) prod[buf][threadIdx.x] = mul_rn(x[i], y[i]);
        __syncthreads();  // one barrier per chunk: the other buffer is being consumed only by thread 0's past loop
        if (threadIdx.x == 0) {
            const int cnt = (int)min((int64_t)1024, n - base);
#pragma unroll 8
            for (int k = 0; k < cnt; ++k) acc = add_rn(acc, prod[buf][k]);
        }
    }
    if (threadIdx.x == 0) *out = acc;
}

}  // namespace

void launch_dot(Context* ctx, int64_t n, const double* x, const double* y, int slot, const int* skip_flag) {
    ATK_NVTX("atk::dot");
    double* out = ctx->slot(slot) + ctx->rank;
    if (ctx->reduction == ATK_REDUCE_SEQUENTIAL) {
        dot_sequential_kernel<<<1, 1024, 0, ctx->stream>>>(n, x, y, out, skip_flag);
    } else {
        const bool vec = aligned16(x) && aligned16(y);
        const int64_t work = vec ? (n + 1) / 2 : n;
        int grid = (int)std::min<int64_t>(std::max<int64_t>((work + kReduceBlock - 1) / kReduceBlock, 1),
                                          (int64_t)ctx->reduce_grid());
        if (vec)
            dot_tree_kernel<true><<<grid, kReduceBlock, 0, ctx->stream>>>(n, x, y, ctx->d_partials, ctx->d_tickets + 0,
                                                                          out, skip_flag);
        else
            dot_tree_kernel<false><<<grid, kReduceBlock, 0, ctx->stream>>>(n, x, y, ctx->d_partials,
                                                                           ctx->d_tickets + 0, out, skip_flag);
    }
    ATK_CUDA_CHECK(cudaGetLastError());
    count_launch(ctx);
    allgather_slot(ctx, slot, ctx->stream);
}

void launch_scale(Context* ctx, int64_t n, const double* x, double s, double* out) { launch_ew<Ew::Scale>(ctx, n, x, nullptr, s, out); }
void launch_div(Context* ctx, int64_t n, const double* x, double s, double* out) { launch_ew<Ew::Div>(ctx, n, x, nullptr, s, out); }
void launch_add(Context* ctx, int64_t n, const double* x, const double* y, double* out) { launch_ew<Ew::Add>(ctx, n, x, y, 0.0, out); }
void launch_sub(Context* ctx, int64_t n, const double* x, const double* y, double* out) { launch_ew<Ew::Sub>(ctx, n, x, y, 0.0, out); }
void launch_axpy(Context* ctx, int64_t n, const double* x, const double* y, double a, double* out) { launch_ew<Ew::Axpy>(ctx, n, x, y, a, out); }
void launch_axmy(Context* ctx, int64_t n, const double* x, const double* y, double a, double* out) { launch_ew<Ew::Axmy>(ctx, n, x, y, a, out); }

}  // namespace atk
