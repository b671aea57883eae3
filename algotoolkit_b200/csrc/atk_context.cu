// atk_context.cu -- context, scratch buffers, memory entry points and the NCCL plumbing.
//
// NCCL is resolved with dlopen() the first time a distributed context is created, so the single-GPU
// library has no link-time dependency on it.  Inside a torch process the already-loaded (torch-bundled)
// libnccl.so.2 is picked up by SONAME; in a plain C++ program the system library is.
#include <dlfcn.h>
#include <nccl.h>

#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <cctype>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "atk_internal.cuh"

namespace atk {

// ------------------------------------------------------------------------------------------ NCCL via dlopen
namespace {
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;
std::once_flag g_nccl_once;

void load_nccl() {
    std::call_once(g_nccl_once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (g_nccl.handle) break;
        }
        if (!g_nccl.handle) return;
#define ATK_SYM(field, sym) g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(g_nccl.handle, sym))
        ATK_SYM(GetUniqueId, "ncclGetUniqueId");
        ATK_SYM(CommInitRank, "ncclCommInitRank");
        ATK_SYM(CommDestroy, "ncclCommDestroy");
        ATK_SYM(AllGather, "ncclAllGather");
        ATK_SYM(Send, "ncclSend");
        ATK_SYM(Recv, "ncclRecv");
        ATK_SYM(GroupStart, "ncclGroupStart");
        ATK_SYM(GroupEnd, "ncclGroupEnd");
        ATK_SYM(GetErrorString, "ncclGetErrorString");
#undef ATK_SYM
    });
    ATK_REQUIRE(g_nccl.handle && g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.AllGather && g_nccl.Send &&
                    g_nccl.Recv && g_nccl.GroupStart && g_nccl.GroupEnd,
                ATK_ERR_NCCL, "libnccl.so.2 could not be loaded (needed for a distributed context)");
}

void nccl_check(ncclResult_t r, const char* what) {
    if (r != ncclSuccess) {
        std::string m = std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "nccl error");
        throw Error(ATK_ERR_NCCL, m);
    }
}
}  // namespace

void allgather_slot(Context* ctx, int slot, cudaStream_t stream) {
    if (ctx->world == 1) return;
    double* base = ctx->slot(slot);
    nccl_check(g_nccl.AllGather(base + ctx->rank, base, 1, ncclDouble, (ncclComm_t)ctx->nccl_comm, stream),
               "ncclAllGather");
}

// in-place all-gather of `count` doubles per rank: rank q's block lives at buf + q*count
void allgather_doubles(Context* ctx, double* buf, size_t count, cudaStream_t stream) {
    if (ctx->world == 1) return;
    nccl_check(g_nccl.AllGather(buf + (size_t)ctx->rank * count, buf, count, ncclDouble, (ncclComm_t)ctx->nccl_comm, stream),
               "ncclAllGather");
}

void exchange_halo(Context* ctx, const double* first_plane, const double* last_plane, double* halo_lo, double* halo_hi,
                   size_t plane_elems, cudaStream_t stream) {
    if (ctx->world == 1) return;
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    const int lo = ctx->rank - 1, hi = ctx->rank + 1;
    nccl_check(g_nccl.GroupStart(), "ncclGroupStart");
    if (lo >= 0) {
        nccl_check(g_nccl.Send(first_plane, plane_elems, ncclDouble, lo, comm, stream), "ncclSend");
        nccl_check(g_nccl.Recv(halo_lo, plane_elems, ncclDouble, lo, comm, stream), "ncclRecv");
    }
    if (hi < ctx->world) {
        nccl_check(g_nccl.Send(last_plane, plane_elems, ncclDouble, hi, comm, stream), "ncclSend");
        nccl_check(g_nccl.Recv(halo_hi, plane_elems, ncclDouble, hi, comm, stream), "ncclRecv");
    }
    nccl_check(g_nccl.GroupEnd(), "ncclGroupEnd");
}

void allgather_host_bytes(Context* ctx, const void* send, void* recv, size_t bytes) {
    if (ctx->world == 1) {
        std::memcpy(recv, send, bytes);
        return;
    }
    if (bytes <= kSetupBytes && ctx->d_setup) {  // the context's own staging buffers: no cudaMalloc / cudaFree per call
        std::memcpy(ctx->h_setup + bytes * (size_t)ctx->rank, send, bytes);
        char* d = ctx->d_setup;
        ATK_CUDA_CHECK(cudaMemcpyAsync(d + bytes * (size_t)ctx->rank, ctx->h_setup + bytes * (size_t)ctx->rank, bytes,
                                       cudaMemcpyHostToDevice, ctx->stream));
        nccl_check(g_nccl.AllGather(d + bytes * (size_t)ctx->rank, d, bytes, ncclChar, (ncclComm_t)ctx->nccl_comm, ctx->stream),
                   "ncclAllGather(setup)");
        ATK_CUDA_CHECK(cudaMemcpyAsync(ctx->h_setup, d, bytes * (size_t)ctx->world, cudaMemcpyDeviceToHost, ctx->stream));
        ATK_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        std::memcpy(recv, ctx->h_setup, bytes * (size_t)ctx->world);
        return;
    }
    char* d = nullptr;
    ATK_CUDA_CHECK(cudaMalloc(&d, bytes * (size_t)ctx->world));
    try {
        ATK_CUDA_CHECK(cudaMemcpyAsync(d + bytes * (size_t)ctx->rank, send, bytes, cudaMemcpyHostToDevice, ctx->stream));
        nccl_check(g_nccl.AllGather(d + bytes * (size_t)ctx->rank, d, bytes, ncclChar, (ncclComm_t)ctx->nccl_comm, ctx->stream),
                   "ncclAllGather(setup)");
        ATK_CUDA_CHECK(cudaMemcpyAsync(recv, d, bytes * (size_t)ctx->world, cudaMemcpyDeviceToHost, ctx->stream));
        ATK_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    } catch (...) {
        cudaFree(d);
        throw;
    }
    cudaFree(d);
}

void ipc_share(Context* ctx, void* local, std::vector<void*>& peers, const std::vector<int>& want) {
    peers.assign((size_t)ctx->world, nullptr);
    peers[(size_t)ctx->rank] = local;
    if (ctx->world == 1) return;
    cudaIpcMemHandle_t mine;
    ATK_CUDA_CHECK(cudaIpcGetMemHandle(&mine, local));
    std::vector<cudaIpcMemHandle_t> all((size_t)ctx->world);
    allgather_host_bytes(ctx, &mine, all.data(), sizeof(mine));
    for (int q = 0; q < ctx->world; ++q) {
        if (q == ctx->rank) continue;
        if (!want.empty() && std::find(want.begin(), want.end(), q) == want.end()) continue;
        void* p = nullptr;
        ATK_CUDA_CHECK(cudaIpcOpenMemHandle(&p, all[(size_t)q], cudaIpcMemLazyEnablePeerAccess));
        peers[(size_t)q] = p;
        ctx->ipc_opened.push_back(p);
    }
}

void* Context::arena_take(size_t bytes) {
    for (auto& a : arena_pool)
        if (!a.busy && a.bytes == bytes) {
            a.busy = true;
            ATK_CUDA_CHECK(cudaMemsetAsync(a.ptr, 0, bytes, stream));
            return a.ptr;
        }
    void* p = nullptr;
    ATK_CUDA_CHECK(cudaMalloc(&p, bytes));
    ATK_CUDA_CHECK(cudaMemset(p, 0, bytes));
    arena_pool.push_back(Arena{p, bytes, true});
    return p;
}
void Context::arena_give(void* ptr) {
    for (auto& a : arena_pool)
        if (a.ptr == ptr) a.busy = false;
}

// ------------------------------------------------------------------------------------------ PeerHalo
namespace {
__global__ void __launch_bounds__(256) halo_push_kernel(const double* __restrict__ to_prev, size_t n_to_prev,
                                                        const double* __restrict__ to_next, size_t n_to_next,
                                                        double* __restrict__ prev_dst, double* __restrict__ next_dst,
                                                        unsigned long long* prev_flag, unsigned long long* next_flag,
                                                        unsigned long long seq, unsigned* ticket) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    auto copy = [&](const double* src, double* dst, size_t n) {
        if (!dst || n == 0) return;
        if ((((uintptr_t)src | (uintptr_t)dst) & 15u) == 0) {
            const size_t n2 = n >> 1;
            for (size_t i = t; i < n2; i += stride) reinterpret_cast<double2*>(dst)[i] = reinterpret_cast<const double2*>(src)[i];
            if ((n & 1) && t == 0) dst[n - 1] = src[n - 1];
        } else {
            for (size_t i = t; i < n; i += stride) dst[i] = src[i];
        }
    };
    copy(to_prev, prev_dst, n_to_prev);
    copy(to_next, next_dst, n_to_next);
    __threadfence_system();  // this thread's NVLink stores are performed before its CTA takes a ticket
    if (last_block_done(ticket)) {
        if (threadIdx.x == 0) {
            __threadfence_system();
            if (prev_flag) st_release_sys(prev_flag, seq);
            if (next_flag) st_release_sys(next_flag, seq);
        }
    }
}
__global__ void halo_wait_kernel(const unsigned long long* flags, int has_prev, int has_next, unsigned long long seq) {
    const int lane = threadIdx.x;
    if ((lane == 0 && has_prev) || (lane == 1 && has_next)) {
        unsigned spins = 0;
        unsigned long long t0 = 0;
        while (ld_acquire_sys(flags + lane) < seq) {
            if ((++spins & 0xfffu) == 0) {
                const unsigned long long now = global_timer_ns();
                if (t0 == 0) t0 = now;
                if (now - t0 > kPeerWaitNs) asm volatile("trap;");  // a neighbour died or fell out of step: fail loudly
            }
        }
    }
}
}  // namespace

void PeerHalo::setup(Context* c, size_t lo, size_t hi) {
    ctx = c;
    cap_lo = lo;
    cap_hi = hi;
    ready = false;
    if (c->world == 1 || !c->p2p) return;
    struct Caps { unsigned long long lo, hi; } mine{lo, hi};
    std::vector<Caps> all((size_t)c->world);
    allgather_host_bytes(c, &mine, all.data(), sizeof(Caps));
    // same arena size on every rank keeps recycled arenas interchangeable; the flag words sit behind the buffers
    arena = static_cast<double*>(c->arena_take(sizeof(double) * arena_doubles(lo, hi)));
    ATK_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    std::vector<void*> peers;
    std::vector<int> want;
    if (c->rank > 0) want.push_back(c->rank - 1);
    if (c->rank + 1 < c->world) want.push_back(c->rank + 1);
    ipc_share(c, arena, peers, want);
    if (c->rank > 0) {
        prev_arena = static_cast<double*>(peers[(size_t)c->rank - 1]);
        prev_cap_lo = all[(size_t)c->rank - 1].lo;
        prev_cap_hi = all[(size_t)c->rank - 1].hi;
    }
    if (c->rank + 1 < c->world) {
        next_arena = static_cast<double*>(peers[(size_t)c->rank + 1]);
        next_cap_lo = all[(size_t)c->rank + 1].lo;
        next_cap_hi = all[(size_t)c->rank + 1].hi;
    }
    ATK_CUDA_CHECK(cudaMalloc(&d_ticket, sizeof(unsigned) * 4));
    ATK_CUDA_CHECK(cudaMemset(d_ticket, 0, sizeof(unsigned) * 4));
    seq = 0;
    ready = true;
}

void PeerHalo::release() {
    if (!ctx) return;
    ctx->ipc_close(prev_arena);
    ctx->ipc_close(next_arena);
    if (arena) ctx->arena_give(arena);
    cudaFree(d_ticket);
    arena = prev_arena = next_arena = nullptr;
    d_ticket = nullptr;
    ready = false;
}

PeerHalo::Targets PeerHalo::begin() {
    ++seq;
    const size_t par = (size_t)(seq & 1);
    Targets t;
    // the previous rank receives into its "hi" buffer of this parity, the next rank into its "lo" buffer
    t.prev_dst = prev_arena ? prev_arena + par * (prev_cap_lo + prev_cap_hi) + prev_cap_lo : nullptr;
    t.next_dst = next_arena ? next_arena + par * (next_cap_lo + next_cap_hi) : nullptr;
    t.prev_flag = prev_arena ? reinterpret_cast<unsigned long long*>(prev_arena + flag_offset(prev_cap_lo, prev_cap_hi)) + 1 : nullptr;
    t.next_flag = next_arena ? reinterpret_cast<unsigned long long*>(next_arena + flag_offset(next_cap_lo, next_cap_hi)) + 0 : nullptr;
    t.my_flags = reinterpret_cast<const unsigned long long*>(arena + flag_offset(cap_lo, cap_hi));
    t.seq = seq;
    return t;
}

void PeerHalo::push(const double* to_prev, size_t n_to_prev, const double* to_next, size_t n_to_next) {
    Context* c = ctx;
    const Targets tg = begin();
    double* prev_dst = tg.prev_dst;
    double* next_dst = tg.next_dst;
    unsigned long long* prev_flag = tg.prev_flag;
    unsigned long long* next_flag = tg.next_flag;
    ATK_REQUIRE(n_to_prev <= (prev_arena ? prev_cap_hi : n_to_prev) && n_to_next <= (next_arena ? next_cap_lo : n_to_next),
                ATK_ERR_INVALID_ARGUMENT, "peer halo: more entries than the neighbour's buffer holds");
    ATK_CUDA_CHECK(cudaEventRecord(c->ev_fork, c->stream));
    ATK_CUDA_CHECK(cudaStreamWaitEvent(c->comm_stream, c->ev_fork, 0));
    const size_t work = std::max(n_to_prev, n_to_next) / 2 + 1;
    const int grid = (int)std::max<size_t>(1, std::min<size_t>((work + 255) / 256, 64));
    halo_push_kernel<<<grid, 256, 0, c->comm_stream>>>(to_prev, n_to_prev, to_next, n_to_next, prev_dst, next_dst, prev_flag, next_flag,
                                                      seq, d_ticket);
    ATK_CUDA_CHECK(cudaGetLastError());
    count_launch(c);
    ATK_CUDA_CHECK(cudaEventRecord(c->ev_join, c->comm_stream));
}

void PeerHalo::wait() {
    Context* c = ctx;
    // our own push must have left before the caller may overwrite x; the neighbours' entries must have landed
    ATK_CUDA_CHECK(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
    const unsigned long long* flags = reinterpret_cast<const unsigned long long*>(arena + flag_offset(cap_lo, cap_hi));
    halo_wait_kernel<<<1, 32, 0, c->stream>>>(flags, prev_arena ? 1 : 0, next_arena ? 1 : 0, seq);
    ATK_CUDA_CHECK(cudaGetLastError());
    count_launch(c);
}

// ------------------------------------------------------------------------------------------ PeerGather
namespace {
__global__ void __launch_bounds__(256) gather_push_kernel(const double* __restrict__ x, const int* __restrict__ idx, GatherTable tb,
                                                          unsigned long long seq, unsigned* ticket, int release) {
    const int total = tb.begin[tb.n_dst];
    const int stride = gridDim.x * blockDim.x;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        int d = 0;
        while (t >= tb.begin[d + 1]) ++d;
        tb.dst[d][t - tb.begin[d]] = x[idx[t]];
    }
    if (!release) return;
    __threadfence_system();
    if (last_block_done(ticket)) {
        if (threadIdx.x < tb.n_dst && tb.flag[threadIdx.x]) {
            __threadfence_system();
            st_release_sys(tb.flag[threadIdx.x], seq);
        }
    }
}
__global__ void gather_wait_kernel(WaitTable wt, unsigned long long seq) {
    const int lane = threadIdx.x;
    if (lane < wt.n) {
        unsigned spins = 0;
        unsigned long long t0 = 0;
        while (ld_acquire_sys(wt.flag[lane]) < seq) {
            if ((++spins & 0xfffu) == 0) {
                const unsigned long long now = global_timer_ns();
                if (t0 == 0) t0 = now;
                if (now - t0 > kPeerWaitNs) asm volatile("trap;");  // a partner died or fell out of step: fail loudly
            }
        }
    }
}
__global__ void shift_idx_kernel(int n, int shift, int* __restrict__ idx) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) idx[t] -= shift;
}
}  // namespace

void PeerGather::setup(Context* c, const int* d_ext_cols, int n_ext_, const std::vector<int>& ext_cols_host,
                       const std::vector<std::pair<int, int>>& blocks) {
    ctx = c;
    n_ext = n_ext_;
    const int W = c->world, me = c->rank;
    recv_count.assign((size_t)W, 0);
    recv_off.assign((size_t)W + 1, 0);
    for (int q = 0; q < W; ++q) {  // ext_cols is ascending and the blocks are contiguous: each owner's entries form a run
        const auto lo = std::lower_bound(ext_cols_host.begin(), ext_cols_host.end(), blocks[(size_t)q].first);
        const auto hi = std::lower_bound(ext_cols_host.begin(), ext_cols_host.end(), blocks[(size_t)q].second);
        recv_count[(size_t)q] = (int)(hi - lo);
        recv_off[(size_t)q] = (int)(lo - ext_cols_host.begin());
    }
    recv_off[(size_t)W] = n_ext;
    ATK_REQUIRE(recv_count[(size_t)me] == 0, ATK_ERR_INVALID_ARGUMENT, "internal: an own column was classified as external");
    // need[p][q]: entries rank p receives from rank q
    std::vector<int> row(kMaxRanks, 0), need((size_t)W * kMaxRanks, 0);
    for (int q = 0; q < W; ++q) row[(size_t)q] = recv_count[(size_t)q];
    allgather_host_bytes(c, row.data(), need.data(), sizeof(int) * kMaxRanks);
    send_count.assign((size_t)W, 0);
    send_off.assign((size_t)W + 1, 0);
    peer_n_ext.assign((size_t)W, 0);
    off_in_peer.assign((size_t)W, 0);
    for (int p = 0; p < W; ++p) {
        send_count[(size_t)p] = need[(size_t)p * kMaxRanks + me];
        send_off[(size_t)p + 1] = send_off[(size_t)p] + send_count[(size_t)p];
        size_t tot = 0, before = 0;
        for (int q = 0; q < W; ++q) {
            if (q < me) before += (size_t)need[(size_t)p * kMaxRanks + q];
            tot += (size_t)need[(size_t)p * kMaxRanks + q];
        }
        peer_n_ext[(size_t)p] = tot;
        off_in_peer[(size_t)p] = before;
    }
    n_send = send_off[(size_t)W];
    partners.clear();
    for (int p = 0; p < W; ++p)
        if (p != me && (send_count[(size_t)p] > 0 || recv_count[(size_t)p] > 0)) partners.push_back(p);
    ATK_REQUIRE((int)partners.size() <= kMaxRanks, ATK_ERR_INVALID_ARGUMENT, "too many partner ranks");
    // every owner learns WHICH of its entries each requester wants: the requesters' column lists travel over NCCL
    ATK_CUDA_CHECK(cudaMalloc(&d_send_idx, sizeof(int) * (size_t)std::max(n_send, 1)));
    ATK_CUDA_CHECK(cudaMalloc(&d_pack, sizeof(double) * (size_t)std::max(n_send, 1)));
    ATK_CUDA_CHECK(cudaMalloc(&nccl_ext, sizeof(double) * (size_t)std::max(n_ext, 1)));
    {
        ncclComm_t comm = (ncclComm_t)c->nccl_comm;
        nccl_check(g_nccl.GroupStart(), "ncclGroupStart");
        for (int q = 0; q < W; ++q) {
            if (q == me) continue;
            if (recv_count[(size_t)q] > 0)
                nccl_check(g_nccl.Send(d_ext_cols + recv_off[(size_t)q], (size_t)recv_count[(size_t)q], ncclInt, q, comm, c->stream), "ncclSend");
            if (send_count[(size_t)q] > 0)
                nccl_check(g_nccl.Recv(d_send_idx + send_off[(size_t)q], (size_t)send_count[(size_t)q], ncclInt, q, comm, c->stream), "ncclRecv");
        }
        nccl_check(g_nccl.GroupEnd(), "ncclGroupEnd");
        if (n_send > 0) {
            shift_idx_kernel<<<(n_send + 255) / 256, 256, 0, c->stream>>>(n_send, blocks[(size_t)me].first, d_send_idx);
            count_launch(c);
        }
        ATK_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    }
    peer_ready = false;
    peer_arena.assign((size_t)W, nullptr);
    if (c->p2p) {
        arena = static_cast<double*>(c->arena_take(sizeof(double) * (flag_offset((size_t)n_ext) + kMaxRanks)));
        ATK_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        std::vector<void*> peers;
        // ipc_share maps "all ranks" when the want-list is empty: a rank without partners asks for itself only
        std::vector<int> want = partners.empty() ? std::vector<int>{me} : partners;
        ipc_share(c, arena, peers, want);
        for (int p : partners) peer_arena[(size_t)p] = static_cast<double*>(peers[(size_t)p]);
        ATK_CUDA_CHECK(cudaMalloc(&d_ticket, sizeof(unsigned)));
        ATK_CUDA_CHECK(cudaMemset(d_ticket, 0, sizeof(unsigned)));
        peer_ready = true;
    }
    seq = 0;
}

void PeerGather::release() {
    if (!ctx) return;
    for (double* p : peer_arena) ctx->ipc_close(p);
    peer_arena.clear();
    if (arena) ctx->arena_give(arena);
    arena = nullptr;
    cudaFree(d_ticket);
    cudaFree(d_send_idx);
    cudaFree(d_pack);
    cudaFree(nccl_ext);
    d_ticket = nullptr;
    d_send_idx = nullptr;
    d_pack = nccl_ext = nullptr;
    peer_ready = false;
}

bool PeerGather::begin_inline(GatherTable* tb, WaitTable* wt, unsigned long long* seq_out, const double** ext) {
    Context* c = ctx;
    const char* comm_env = std::getenv("ATK_COMM");
    if (!(peer_ready && !(comm_env && std::string(comm_env) == "nccl") && !c->capturing())) return false;
    last_was_peer = true;
    ++seq;
    const size_t par = (size_t)(seq & 1);
    tb->n_dst = (int)partners.size();
    int pos = 0;
    tb->begin[0] = 0;
    // d_send_idx is grouped by destination RANK; the partners are in ascending rank order, and ranks that are no
    // partners have nothing to receive, so walking the partners in order walks d_send_idx in order
    for (int d = 0; d < tb->n_dst; ++d) {
        const int p = partners[(size_t)d];
        ATK_REQUIRE(send_off[(size_t)p] == pos, ATK_ERR_RUNTIME, "internal: send list order");
        double* pa = peer_arena[(size_t)p];
        tb->dst[d] = send_count[(size_t)p] > 0 ? pa + par * peer_n_ext[(size_t)p] + off_in_peer[(size_t)p] : nullptr;
        tb->flag[d] = reinterpret_cast<unsigned long long*>(pa + flag_offset(peer_n_ext[(size_t)p])) + c->rank;
        pos += send_count[(size_t)p];
        tb->begin[d + 1] = pos;
    }
    wt->n = (int)partners.size();
    const unsigned long long* flags = reinterpret_cast<const unsigned long long*>(arena + flag_offset((size_t)n_ext));
    for (int d = 0; d < wt->n; ++d) wt->flag[d] = flags + partners[(size_t)d];
    *seq_out = seq;
    *ext = arena + par * (size_t)n_ext;
    return true;
}

const double* PeerGather::begin(const double* x) {
    Context* c = ctx;
    const char* comm_env = std::getenv("ATK_COMM");
    const bool peer = peer_ready && !(comm_env && std::string(comm_env) == "nccl") && !c->capturing();
    last_was_peer = peer;
    ATK_CUDA_CHECK(cudaEventRecord(c->ev_fork, c->stream));
    ATK_CUDA_CHECK(cudaStreamWaitEvent(c->comm_stream, c->ev_fork, 0));
    GatherTable tb;
    const double* result = nullptr;
    if (peer) {
        WaitTable wt_unused;
        unsigned long long seq_now = 0;
        begin_inline(&tb, &wt_unused, &seq_now, &result);
        const int grid = std::max(1, std::min((n_send + 255) / 256, 128));
        gather_push_kernel<<<grid, 256, 0, c->comm_stream>>>(x, d_send_idx, tb, seq, d_ticket, 1);
        ATK_CUDA_CHECK(cudaGetLastError());
        count_launch(c);
    } else {
        tb.n_dst = 1;
        tb.dst[0] = d_pack;
        tb.begin[0] = 0;
        tb.begin[1] = n_send;
        tb.flag[0] = nullptr;
        if (n_send > 0) {
            const int grid = std::max(1, std::min((n_send + 255) / 256, 128));
            gather_push_kernel<<<grid, 256, 0, c->comm_stream>>>(x, d_send_idx, tb, 0ull, nullptr, 0);
            ATK_CUDA_CHECK(cudaGetLastError());
            count_launch(c);
        }
        ncclComm_t comm = (ncclComm_t)c->nccl_comm;
        nccl_check(g_nccl.GroupStart(), "ncclGroupStart");
        for (int q : partners) {
            if (send_count[(size_t)q] > 0)
                nccl_check(g_nccl.Send(d_pack + send_off[(size_t)q], (size_t)send_count[(size_t)q], ncclDouble, q, comm, c->comm_stream), "ncclSend");
            if (recv_count[(size_t)q] > 0)
                nccl_check(g_nccl.Recv(nccl_ext + recv_off[(size_t)q], (size_t)recv_count[(size_t)q], ncclDouble, q, comm, c->comm_stream), "ncclRecv");
        }
        nccl_check(g_nccl.GroupEnd(), "ncclGroupEnd");
        result = nccl_ext;
    }
    ATK_CUDA_CHECK(cudaEventRecord(c->ev_join, c->comm_stream));
    return result;
}

void PeerGather::wait() {
    Context* c = ctx;
    ATK_CUDA_CHECK(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
    if (last_was_peer && !partners.empty()) {
        WaitTable wt;
        wt.n = (int)partners.size();
        const unsigned long long* flags = reinterpret_cast<const unsigned long long*>(arena + flag_offset((size_t)n_ext));
        for (int d = 0; d < wt.n; ++d) wt.flag[d] = flags + partners[(size_t)d];
        gather_wait_kernel<<<1, 32, 0, c->stream>>>(wt, seq);
        ATK_CUDA_CHECK(cudaGetLastError());
        count_launch(c);
    }
}

void Context::ipc_close(void* mapped) {
    if (!mapped) return;
    auto it = std::find(ipc_opened.begin(), ipc_opened.end(), mapped);
    if (it == ipc_opened.end()) return;
    cudaIpcCloseMemHandle(mapped);
    ipc_opened.erase(it);
}

// Decide collectively whether peer-memory signalling can be used and set it up.
static void setup_p2p(Context& c) {
    c.p2p = false;
    if (c.world == 1) return;
    const char* env = std::getenv("ATK_COMM");
    int ok = !(env && std::string(env) == "nccl");
    // ranks are assumed to be one per device of this node, device index = local rank; check the mapping is possible
    int ndev = 0;
    cudaGetDeviceCount(&ndev);
    // all ranks exchange (device, hostname hash) and everyone checks peer access to every other device
    struct Info { int device; unsigned long long host; } mine{c.device, 0ull};
    char hn[256] = {0};
    gethostname(hn, sizeof(hn) - 1);
    for (const char* q = hn; *q; ++q) mine.host = mine.host * 1315423911ull + (unsigned char)*q;
    std::vector<Info> all((size_t)c.world);
    allgather_host_bytes(&c, &mine, all.data(), sizeof(Info));
    for (int q = 0; q < c.world && ok; ++q) {
        if (q == c.rank) continue;
        if (all[(size_t)q].host != mine.host || all[(size_t)q].device == c.device || all[(size_t)q].device >= ndev) { ok = 0; break; }
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, c.device, all[(size_t)q].device) != cudaSuccess || !can) ok = 0;
    }
    std::vector<int> oks((size_t)c.world);
    allgather_host_bytes(&c, &ok, oks.data(), sizeof(int));
    for (int v : oks) ok = ok && v;
    if (!ok) return;
    std::vector<void*> peers;
    ipc_share(&c, c.d_gather, peers, {});  // d_gather and d_flags are one allocation (init_context)
    std::vector<double*> pg((size_t)c.world);
    std::vector<unsigned long long*> pf((size_t)c.world), pl((size_t)c.world);
    for (int q = 0; q < c.world; ++q) {
        pg[(size_t)q] = static_cast<double*>(peers[(size_t)q]);
        pf[(size_t)q] = reinterpret_cast<unsigned long long*>(pg[(size_t)q] + (size_t)kNumSlots * kMaxRanks);
        pl[(size_t)q] = pf[(size_t)q] + (size_t)kNumSlots * kMaxRanks;
    }
    ATK_CUDA_CHECK(cudaMalloc(&c.d_peer_ll, sizeof(unsigned long long*) * (size_t)c.world));
    ATK_CUDA_CHECK(cudaMemcpy(c.d_peer_ll, pl.data(), sizeof(unsigned long long*) * (size_t)c.world, cudaMemcpyHostToDevice));
    ATK_CUDA_CHECK(cudaMalloc(&c.d_peer_gather, sizeof(double*) * (size_t)c.world));
    ATK_CUDA_CHECK(cudaMalloc(&c.d_peer_flags, sizeof(unsigned long long*) * (size_t)c.world));
    ATK_CUDA_CHECK(cudaMemcpy(c.d_peer_gather, pg.data(), sizeof(double*) * (size_t)c.world, cudaMemcpyHostToDevice));
    ATK_CUDA_CHECK(cudaMemcpy(c.d_peer_flags, pf.data(), sizeof(unsigned long long*) * (size_t)c.world, cudaMemcpyHostToDevice));
    c.p2p = true;
}

double read_slot(Context* ctx, int slot) {
    ATK_CUDA_CHECK(cudaMemcpyAsync(ctx->h_pinned, ctx->slot(slot), sizeof(double) * ctx->world, cudaMemcpyDeviceToHost,
                                   ctx->stream));
    ATK_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    // rank-ordered sum; volatile keeps the host compiler from reassociating (it may not anyway)
    volatile double s = ctx->h_pinned[0];
    for (int q = 1; q < ctx->world; ++q) s = s + ctx->h_pinned[q];
    return s;
}

// ------------------------------------------------------------------------------------------ pinned host memory
// One process per GPU: each rank's staging buffers should live on the socket its GPU hangs off, otherwise the uploads of
// the far ranks all squeeze through the inter-socket link (measured in round 1: 12-14 ms per call at 4 and 8 ranks that
// were neither loop nor PCIe time).  libnuma is not assumed: mbind / get_mempolicy / move_pages are raw syscalls.
namespace {
std::mutex g_pin_mutex;
struct PinRec { void* p; size_t bytes; };
std::vector<PinRec> g_pinned;  // buffers that came from mmap + cudaHostRegister

int device_numa_node() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), dev) != cudaSuccess) return -1;
    for (char* q = bus; *q; ++q) *q = (char)std::tolower((unsigned char)*q);
    const std::string path = std::string("/sys/bus/pci/devices/") + bus + "/numa_node";
    FILE* f = std::fopen(path.c_str(), "r");
    if (!f) return -1;
    int node = -1;
    if (std::fscanf(f, "%d", &node) != 1) node = -1;
    std::fclose(f);
    return node;
}
}  // namespace

void* host_alloc_pinned(size_t bytes) {
    const char* env = std::getenv("ATK_NUMA_PIN");
    const int node = (env && env[0] == '0') ? -1 : device_numa_node();
    if (node >= 0 && node < 1024) {
        const size_t page = 2u << 20;  // round to 2 MiB so that transparent huge pages can back the buffer
        const size_t len = (bytes + page - 1) / page * page;
        void* p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (p != MAP_FAILED) {
            unsigned long mask[16] = {0};
            mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
            // MPOL_PREFERRED = 1: fall back to other nodes rather than fail when the node is full
            (void)syscall(SYS_mbind, p, len, 1 /*MPOL_PREFERRED*/, mask, sizeof(mask) * 8, 0u);
            if (cudaHostRegister(p, len, cudaHostRegisterPortable) == cudaSuccess) {
                std::lock_guard<std::mutex> lk(g_pin_mutex);
                g_pinned.push_back(PinRec{p, len});
                return p;
            }
            cudaGetLastError();
            munmap(p, len);
        }
    }
    void* p = nullptr;
    ATK_CUDA_CHECK(cudaMallocHost(&p, bytes));
    return p;
}

void host_free_pinned(void* p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(g_pin_mutex);
        for (size_t i = 0; i < g_pinned.size(); ++i)
            if (g_pinned[i].p == p) {
                const size_t len = g_pinned[i].bytes;
                g_pinned.erase(g_pinned.begin() + (long)i);
                cudaHostUnregister(p);
                munmap(p, len);
                return;
            }
    }
    ATK_CUDA_CHECK(cudaFreeHost(p));
}

void host_numa_query(const void* p, int* buffer_node, int* device_node) {
    if (device_node) *device_node = device_numa_node();
    if (buffer_node) {
        *buffer_node = -1;
        if (p) {
            int node = -1;
            // get_mempolicy(MPOL_F_NODE | MPOL_F_ADDR): the node the page at `p` is allocated on
            if (syscall(SYS_get_mempolicy, &node, nullptr, 0ul, const_cast<void*>(p), 3ul /*MPOL_F_NODE|MPOL_F_ADDR*/) == 0)
                *buffer_node = node;
        }
    }
}

// ------------------------------------------------------------------------------------------ workspace cache
void* Context::ws_alloc(size_t bytes) {
    if (bytes == 0) bytes = 8;
    int best = -1;
    for (int i = 0; i < (int)ws_blocks.size(); ++i)
        if (!ws_blocks[i].busy && ws_blocks[i].bytes >= bytes && ws_blocks[i].bytes <= 2 * bytes + (1u << 20) &&
            (best < 0 || ws_blocks[i].bytes < ws_blocks[best].bytes))
            best = i;
    if (best >= 0) {
        ws_blocks[best].busy = true;
        return ws_blocks[best].ptr;
    }
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaErrorMemoryAllocation) {  // give the idle cache back and retry once
        cudaGetLastError();
        ws_trim();
        e = cudaMalloc(&p, bytes);
    }
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        throw Error(ATK_ERR_NOMEM, "out of device memory");
    }
    ATK_CUDA_CHECK(e);
    ws_blocks.push_back(WsBlock{p, bytes, true});
    return p;
}
void Context::ws_release(void* ptr) {
    for (auto& b : ws_blocks)
        if (b.ptr == ptr) {
            b.busy = false;
            return;
        }
}
void Context::ws_trim() {
    std::vector<WsBlock> keep;
    for (auto& b : ws_blocks) {
        if (b.busy) keep.push_back(b);
        else cudaFree(b.ptr);
    }
    ws_blocks.swap(keep);
}

// ------------------------------------------------------------------------------------------ context
static void init_context(Context& c, int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        throw Error(ATK_ERR_CUDA, std::string("no CUDA device available (") + cudaGetErrorString(e) +
                                      "); libatk_cuda has no CPU fallback");
    ATK_REQUIRE(device >= 0 && device < count, ATK_ERR_INVALID_ARGUMENT, "device index out of range");
    c.device = device;
    ATK_CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    ATK_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    c.sm_count = prop.multiProcessorCount;
    ATK_CUDA_CHECK(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    {  // communication stream at the highest priority: its small push kernels must not queue behind a full compute grid
        int prio_lo = 0, prio_hi = 0;
        ATK_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        ATK_CUDA_CHECK(cudaStreamCreateWithPriority(&c.comm_stream, cudaStreamNonBlocking, prio_hi));
    }
    ATK_CUDA_CHECK(cudaEventCreateWithFlags(&c.ev_fork, cudaEventDisableTiming));
    ATK_CUDA_CHECK(cudaEventCreateWithFlags(&c.ev_join, cudaEventDisableTiming));
    ATK_CUDA_CHECK(cudaEventCreate(&c.ev_t0));
    ATK_CUDA_CHECK(cudaEventCreate(&c.ev_t1));
    ATK_CUDA_CHECK(cudaEventCreate(&c.ev_h0));
    ATK_CUDA_CHECK(cudaEventCreate(&c.ev_h1));
    ATK_CUDA_CHECK(cudaEventCreate(&c.ev_h2));
    ATK_CUDA_CHECK(cudaEventCreate(&c.ev_h3));
    ATK_CUDA_CHECK(cudaMalloc(&c.d_setup, kSetupBytes * kMaxRanks));
    ATK_CUDA_CHECK(cudaMallocHost(&c.h_setup, kSetupBytes * kMaxRanks));
    ATK_CUDA_CHECK(cudaMalloc(&c.d_partials, sizeof(double) * kMaxPartials));
    ATK_CUDA_CHECK(cudaMalloc(&c.d_tickets, sizeof(unsigned) * kNumTickets));
    // reduction slots followed by their sequence flags: one allocation, so that one IPC handle shares both
    // [gather doubles | flags | packed value/sequence words x2]
    const size_t arena = (sizeof(double) + 3 * sizeof(unsigned long long)) * kNumSlots * kMaxRanks;
    ATK_CUDA_CHECK(cudaMalloc(&c.d_gather, arena));
    c.d_flags = reinterpret_cast<unsigned long long*>(c.d_gather + (size_t)kNumSlots * kMaxRanks);
    c.d_ll = c.d_flags + (size_t)kNumSlots * kMaxRanks;
    ATK_CUDA_CHECK(cudaMemset(c.d_tickets, 0, sizeof(unsigned) * kNumTickets));
    ATK_CUDA_CHECK(cudaMemset(c.d_gather, 0, arena));
    // ATK_EPOCH_SEED: start value of the sequence-number base (stress tests seed it just below 2^32 so that the 32-bit
    // tags of the packed words wrap during the run); must be the same on every rank
    if (const char* e = std::getenv("ATK_EPOCH_SEED")) c.epoch = std::strtoull(e, nullptr, 0);
    ATK_CUDA_CHECK(cudaMallocHost(&c.h_pinned, sizeof(double) * 64));
    ATK_CUDA_CHECK(cudaMallocHost(&c.h_flags, sizeof(int) * kHostFlags));
    ATK_CUDA_CHECK(cudaDeviceSynchronize());
}

atk_context_s* context_create(int device) {
    auto* h = new atk_context_s();
    try {
        init_context(h->c, device);
    } catch (...) {
        delete h;
        throw;
    }
    return h;
}

atk_context_s* context_create_distributed(int device, int rank, int world, const void* uid, size_t uid_bytes) {
    ATK_REQUIRE(world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world, ATK_ERR_INVALID_ARGUMENT,
                "bad rank / world size");
    auto* h = new atk_context_s();
    try {
        init_context(h->c, device);
        h->c.rank = rank;
        h->c.world = world;
        if (world > 1) {
            load_nccl();
            ATK_REQUIRE(uid && uid_bytes == sizeof(ncclUniqueId), ATK_ERR_INVALID_ARGUMENT, "bad NCCL unique id blob");
            ncclUniqueId id;
            std::memcpy(&id, uid, sizeof(id));
            ncclComm_t comm;
            nccl_check(g_nccl.CommInitRank(&comm, world, id, rank), "ncclCommInitRank");
            h->c.nccl_comm = comm;
            setup_p2p(h->c);
        }
    } catch (...) {
        delete h;
        throw;
    }
    return h;
}

void nccl_unique_id(void* buf, size_t cap, size_t* bytes) {
    load_nccl();
    ATK_REQUIRE(buf && cap >= sizeof(ncclUniqueId), ATK_ERR_INVALID_ARGUMENT, "unique-id buffer too small (need 128 bytes)");
    ncclUniqueId id;
    nccl_check(g_nccl.GetUniqueId(&id), "ncclGetUniqueId");
    std::memcpy(buf, &id, sizeof(id));
    if (bytes) *bytes = sizeof(id);
}

void context_destroy(atk_context_s* h) {
    if (!h) return;
    Context& c = h->c;
    cudaSetDevice(c.device);
    cudaDeviceSynchronize();
    for (auto& b : c.ws_blocks) cudaFree(b.ptr);
    c.ws_blocks.clear();
    for (void* p : c.ipc_opened) cudaIpcCloseMemHandle(p);
    for (void* p : c.ipc_exported) cudaFree(p);  // exported allocations: kept alive until now for the peers' sake
    for (auto& a : c.arena_pool) cudaFree(a.ptr);
    cudaFree(c.d_peer_gather);
    cudaFree(c.d_peer_flags);
    cudaFree(c.d_peer_ll);
    if (c.nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)c.nccl_comm);
    cudaFree(c.d_partials);
    cudaFree(c.d_tickets);
    cudaFree(c.d_gather);
    cudaFreeHost(c.h_pinned);
    cudaFreeHost(c.h_flags);
    cudaEventDestroy(c.ev_fork);
    cudaEventDestroy(c.ev_join);
    cudaEventDestroy(c.ev_t0);
    cudaEventDestroy(c.ev_t1);
    cudaEventDestroy(c.ev_h0);
    cudaEventDestroy(c.ev_h1);
    cudaEventDestroy(c.ev_h2);
    cudaEventDestroy(c.ev_h3);
    cudaFree(c.d_setup);
    cudaFreeHost(c.h_setup);
    cudaStreamDestroy(c.stream);
    cudaStreamDestroy(c.comm_stream);
    delete h;
}

}  // namespace atk
