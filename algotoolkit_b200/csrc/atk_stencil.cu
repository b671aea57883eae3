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
#include <algorithm>
#include <cstdlib>

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

struct Stencil7 : public Operator {
    StencilGeom g;
    double* halo_lo = nullptr;
    double* halo_hi = nullptr;
    int kc = 16;
    int tile = 0;  // 0: 32x8 threads (64x8 points), 1: 64x4 (128x4), 2: 32x16 (64x16), 3: 64x8 (128x8)

    ~Stencil7() override {
        cudaFree(halo_lo);
        cudaFree(halo_hi);
    }

    template <int TX, int TY>
    void launch(const double* x, double* y, int k_lo, int k_hi, bool dot, int partial_base, bool dot_final, double* dot_out,
                const int* skip_flag, int* blocks_out) {
        Context* c = ctx;
        const int planes = k_hi - k_lo;
        *blocks_out = 0;
        if (planes <= 0) return;
        const int kcc = std::min(kc, planes);
        dim3 grid((g.nx + 2 * TX - 1) / (2 * TX), (g.ny + TY - 1) / TY, (planes + kcc - 1) / kcc);
        dim3 block(TX, TY, 1);
        const int nblocks = (int)(grid.x * grid.y * grid.z);
        ATK_REQUIRE(partial_base + nblocks <= kMaxPartials, ATK_ERR_INVALID_ARGUMENT, "stencil grid exceeds the partial-sum scratch");
        const bool vec = (g.nx % 2 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15u) == 0) &&
                         ((reinterpret_cast<uintptr_t>(y) & 15u) == 0);
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
            default: launch<32, 8>(x, y, k_lo, k_hi, dot, partial_base, dot_final, dot_out, skip_flag, blocks_out); break;
        }
    }

    void apply(const double* x, double* y, int dot_slot, const int* skip_flag) override {
        Context* c = ctx;
        const bool dot = dot_slot >= 0;
        double* dot_out = dot ? c->slot(dot_slot) + c->rank : nullptr;
        int nb = 0;
        if (c->world == 1) {
            launch_tile(x, y, 0, g.nz_loc, dot, 0, true, dot_out, skip_flag, &nb);
        } else {
            // halo exchange on the communication stream, overlapped with the planes that do not need it
            const size_t plane = (size_t)g.nx * g.ny;
            ATK_CUDA_CHECK(cudaEventRecord(c->ev_fork, c->stream));
            ATK_CUDA_CHECK(cudaStreamWaitEvent(c->comm_stream, c->ev_fork, 0));
            exchange_halo(c, x, x + (size_t)(g.nz_loc - 1) * plane, halo_lo, halo_hi, plane, c->comm_stream);
            ATK_CUDA_CHECK(cudaEventRecord(c->ev_join, c->comm_stream));
            int nb_int = 0, nb_lo = 0, nb_hi = 0;
            if (g.nz_loc > 2) launch_tile(x, y, 1, g.nz_loc - 1, dot, 0, false, dot_out, skip_flag, &nb_int);
            ATK_CUDA_CHECK(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
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

}  // namespace

Operator* make_stencil7(Context* ctx, int nx, int ny, int nz, double cx, double cy, double cz) {
    ATK_REQUIRE(nx > 0 && ny > 0 && nz > 0, ATK_ERR_INVALID_ARGUMENT, "grid dimensions must be positive");
    ATK_REQUIRE(nz >= ctx->world, ATK_ERR_INVALID_ARGUMENT, "fewer z-planes than ranks");
    auto* S = new Stencil7();
    S->ctx = ctx;
    // balanced z-slab split: the first (nz % world) ranks get one extra plane
    const int base = nz / ctx->world, rem = nz % ctx->world;
    const int k_begin = ctx->rank * base + std::min(ctx->rank, rem);
    const int nz_loc = base + (ctx->rank < rem ? 1 : 0);
    S->g = StencilGeom{nx, ny, nz, k_begin, nz_loc, cx, cy, cz, -2.0 * (cx + cy + cz)};  // cc: Poisson.hpp:57
    const int64_t plane = (int64_t)nx * ny;
    S->n_rows_local = plane * nz_loc;
    S->n_cols = plane * nz;
    S->row_begin = plane * k_begin;
    const int64_t interior = (nx > 2 && ny > 2 && nz > 2) ? (int64_t)(nx - 2) * (ny - 2) * (nz - 2) : 0;
    S->nnz = (plane * nz - interior) + 7 * interior;
    if (const char* e = std::getenv("ATK_STENCIL_KC")) S->kc = std::max(1, std::atoi(e));
    if (const char* e = std::getenv("ATK_STENCIL_TILE")) S->tile = std::atoi(e);
    try {
        if (ctx->world > 1) {
            ATK_CUDA_CHECK(cudaMalloc(&S->halo_lo, sizeof(double) * (size_t)plane));
            ATK_CUDA_CHECK(cudaMalloc(&S->halo_hi, sizeof(double) * (size_t)plane));
            ATK_CUDA_CHECK(cudaMemset(S->halo_lo, 0, sizeof(double) * (size_t)plane));
            ATK_CUDA_CHECK(cudaMemset(S->halo_hi, 0, sizeof(double) * (size_t)plane));
        }
    } catch (...) {
        delete S;
        throw;
    }
    return S;
}

}  // namespace atk
