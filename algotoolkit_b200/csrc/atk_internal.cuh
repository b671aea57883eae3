// atk_internal.cuh -- shared declarations of libatk_cuda.so (not part of the public ABI).
//
// Layout of the library:
//   atk_context.cu   context, streams, scratch, memory entry points, NCCL (dlopen'ed lazily), peer-memory exchanges
//                    (PeerHalo, PeerGather), NUMA-local pinned host buffers
//   atk_vector.cu    VectorObj kernels: dot / nrm2 / scale / div / add / sub / axpy
//   atk_spmv.cu      CSC -> CSR -> SELL-32 conversion on the device, SELL and CSR-vector SpMV
//   atk_stencil.cu   matrix-free 7-point Poisson operator (+ fused p.Ap), device-side assembly
//   atk_cg.cu        ConjugateGrad::solve device loop
//   atk_gmres.cu     GMRES::solve device loop
//   atk_ilu.cu       ILUPreconditioner (dense-in-disguise elimination + triangular sweeps) for GMRES::enablePreconditioner
//   atk_ilu0.cu      sparse ILU(0) (the same elimination restricted to A's pattern), level-scheduled factorisation / solves
//   atk_gd.cu        GradientDescent::solve device loop
//   atk_api.cu       extern "C" wrappers (exceptions -> status codes)
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <exception>
#include <string>
#include <utility>
#include <vector>

#include "../../include/atk_cuda.h"

#include <nvtx3/nvToolsExt.h>  // header-only NVTX v3: ranges cost nothing unless a profiler is attached

namespace atk {

// One NVTX range per solver call and per kernel class (nsys / ncu --nvtx group the launches by it).
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};
#define ATK_NVTX_CAT2(a, b) a##b
#define ATK_NVTX_CAT(a, b) ATK_NVTX_CAT2(a, b)
#define ATK_NVTX(name) ::atk::NvtxRange ATK_NVTX_CAT(atk_nvtx_range_, __LINE__)(name)

// Checked build (python -m algotoolkit_b200.build --checked -> libatk_cuda_checked.so, ATK_LIB=checked selects it): device-side
// assertions on tile / halo indices and on the sequence numbers of the peer protocols; a violation prints and traps.
#ifdef ATK_CHECKED
#define ATK_DEVICE_ASSERT(cond)                                                                        \
    do {                                                                                               \
        if (!(cond)) {                                                                                 \
            printf("ATK_CHECKED: %s failed at %s:%d (block %d thread %d)\n", #cond, __FILE__, __LINE__, \
                   (int)(blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)),               \
                   (int)(threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z)));         \
            asm volatile("trap;");                                                                     \
        }                                                                                              \
    } while (0)
#else
#define ATK_DEVICE_ASSERT(cond) ((void)0)
#endif

// ------------------------------------------------------------------------------------------ errors
struct Error : public std::exception {
    int code;
    std::string msg;
    Error(int c, std::string m) : code(c), msg(std::move(m)) {}
    const char* what() const noexcept override { return msg.c_str(); }
};

#define ATK_CUDA_CHECK(expr)                                                                             \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess)                                                                           \
            throw ::atk::Error(ATK_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + \
                                                 __FILE__ + ":" + std::to_string(__LINE__) + ")");       \
    } while (0)

#define ATK_REQUIRE(cond, code, text)                    \
    do {                                                 \
        if (!(cond)) throw ::atk::Error((code), (text)); \
    } while (0)

// ------------------------------------------------------------------------------------------ constants
constexpr int kMaxRanks = 16;          // ranks per communicator (one box: <= 8)
constexpr int kReduceBlock = 256;      // block size of every kernel that ends in a grid-wide sum
constexpr int kMaxPartials = 1 << 16;  // per-block partial sums a single kernel may produce
constexpr int kNumSlots = 8;           // independent reduction result slots
constexpr int kNumTickets = 16;
constexpr int kHostFlags = 1 << 16;
constexpr size_t kSetupBytes = 256;    // largest per-rank record of a setup-time host all-gather

// Reduction slots: slot s lives at gather[s*kMaxRanks + q], q = rank.  Each rank writes its
// own entry; on a distributed context an in-place all-gather fills the others; consumers add
// the entries in rank order, so every rank derives bit-identical scalars.
enum Slot : int { kSlotPAp = 0, kSlotRR = 1, kSlotGeneric = 2, kSlotBnorm = 3, kSlotH = 4, kSlotNorm = 5 };

struct Context {
    int device = 0;
    int rank = 0, world = 1;
    int sm_count = 148;
    cudaStream_t stream = nullptr;       // compute
    cudaStream_t comm_stream = nullptr;  // halo exchange
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
    cudaEvent_t ev_h0 = nullptr, ev_h1 = nullptr, ev_h2 = nullptr, ev_h3 = nullptr;  // phases of the *_host entry points
    char* d_setup = nullptr;             // [kSetupBytes * kMaxRanks] staging of allgather_host_bytes (no cudaMalloc per call)
    char* h_setup = nullptr;             // pinned mirror
    atk_reduction reduction = ATK_REDUCE_TREE;
    double* d_partials = nullptr;  // [kMaxPartials]
    unsigned* d_tickets = nullptr; // [kNumTickets], self-resetting (atomicInc wrap)
    double* d_gather = nullptr;    // [kNumSlots * kMaxRanks]
    double* h_pinned = nullptr;    // [64] pinned scratch for scalar read-back
    int* h_flags = nullptr;        // [kHostFlags] pinned: exit flags read back by the CG driver
    int64_t launches = 0;
    void* nccl_comm = nullptr;

    // ---- peer-memory signalling (distributed contexts whose GPUs can all map each other; NVLink / NVSwitch) ----
    // d_gather and d_flags form one allocation that every peer maps through CUDA IPC.  A rank publishes a partial sum
    // by storing it into EVERY rank's d_gather[slot][my_rank] and then releasing a sequence number into the matching
    // d_flags entry; consumers spin on their LOCAL flags.  No NCCL call is needed inside an iteration.
    bool p2p = false;
    unsigned long long* d_flags = nullptr;       // [kNumSlots * kMaxRanks], local
    unsigned long long* d_ll = nullptr;          // [kNumSlots * kMaxRanks * 2], local: value halves packed with their sequence
    double** d_peer_gather = nullptr;            // device array [world]: every rank's d_gather as mapped here
    unsigned long long** d_peer_flags = nullptr; // device array [world]
    unsigned long long** d_peer_ll = nullptr;    // device array [world]
    std::vector<void*> ipc_opened;               // mappings to close at destruction
    std::vector<void*> ipc_exported;             // allocations other ranks may still have mapped: freed with the context
    // Exported halo arenas are recycled, not freed, when their operator dies (a peer may still have them mapped): creating
    // and destroying operators in a loop then costs one arena per distinct size, not one per operator.
    struct Arena { void* ptr; size_t bytes; bool busy; };
    std::vector<Arena> arena_pool;
    void* arena_take(size_t bytes);              // a free pooled arena of exactly `bytes`, else a new zeroed allocation
    void arena_give(void* ptr);                  // back to the pool
    void ipc_close(void* mapped);                // close one mapping early (operator destruction)
    unsigned long long epoch = 1;                // sequence-number base, advanced identically on all ranks

    // ---- workspace cache: solver-internal device buffers are recycled between calls instead of going through
    // cudaMalloc / cudaFree (each costs ~0.1-1 ms at GB sizes and synchronises the device) ----
    struct WsBlock { void* ptr; size_t bytes; bool busy; };
    std::vector<WsBlock> ws_blocks;
    void* ws_alloc(size_t bytes);   // smallest free cached block that fits, else a new cudaMalloc
    void ws_release(void* ptr);     // back to the cache
    void ws_trim();                 // cudaFree every idle block

    // true while ctx->stream is being captured into a CUDA graph (the generic CG loop captures one iteration): work whose
    // kernel arguments change from call to call (PeerHalo sequence numbers) must then take its capturable NCCL form
    bool capturing() const {
        cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
        return cudaStreamIsCapturing(stream, &st) == cudaSuccess && st != cudaStreamCaptureStatusNone;
    }
    double* slot(int s) const { return d_gather + (size_t)s * kMaxRanks; }
    int reduce_grid() const { return sm_count * 4; }
};

// RAII handle on a cached workspace block
struct WsBuffer {
    Context* ctx = nullptr;
    void* ptr = nullptr;
    WsBuffer() = default;
    WsBuffer(Context* c, size_t bytes) : ctx(c), ptr(c->ws_alloc(bytes)) {}
    WsBuffer(const WsBuffer&) = delete;
    WsBuffer& operator=(const WsBuffer&) = delete;
    ~WsBuffer() { if (ptr) ctx->ws_release(ptr); }
    template <typename T> T* as() const { return static_cast<T*>(ptr); }
};

inline void count_launch(Context* c, int n = 1) { c->launches += n; }

// In-place all-gather of one reduction slot across ranks (no-op when world == 1).
void allgather_slot(Context* ctx, int slot, cudaStream_t stream);
void allgather_doubles(Context* ctx, double* buf, size_t count, cudaStream_t stream);
// Exchange one xy-plane with the z-neighbours: sends `first_plane` down / `last_plane` up, receives into
// halo_lo / halo_hi.  Runs on `stream`.
void exchange_halo(Context* ctx, const double* first_plane, const double* last_plane, double* halo_lo, double* halo_hi,
                   size_t plane_elems, cudaStream_t stream);
// Host value of a reduction slot (sum over ranks in rank order); synchronises the compute stream.
double read_slot(Context* ctx, int slot);
// Collective: every rank passes a device allocation of its own (same size everywhere); on return peers[q] is rank q's
// allocation mapped into this process (peers[rank] = local).  Only ranks listed in `want` are mapped (empty = all).
void ipc_share(Context* ctx, void* local, std::vector<void*>& peers, const std::vector<int>& want);
// Pinned host memory on the NUMA node of the current CUDA device (atk_context.cu).
void* host_alloc_pinned(size_t bytes);
void host_free_pinned(void* p);
void host_numa_query(const void* p, int* buffer_node, int* device_node);
// Collective all-gather of `bytes` host bytes per rank (setup-time helper, goes through NCCL).
void allgather_host_bytes(Context* ctx, const void* send, void* recv, size_t bytes);

// ------------------------------------------------------------------------------------------ peer-memory halo exchange
// Neighbour exchange of an operator apply without NCCL (distributed contexts with Context::p2p): every rank exports one
// arena [lo0 | hi0 | lo1 | hi1 | flags]; an apply PUSHES its boundary entries straight into the neighbours' buffers of the
// current parity with NVLink stores (one small kernel on the high-priority communication stream, system fence, the CTA
// that finishes last releases a sequence number into the neighbours' flag words) while the rows that need no remote
// entry are computed; a one-warp wait kernel on the compute stream then spins on the LOCAL flags before the boundary rows
// run.  Two parities suffice: a rank can only push apply n+2 after it has seen its neighbour's flag n+1, which the
// neighbour released after its own reads of apply n.
struct PeerHalo {
    Context* ctx = nullptr;
    size_t cap_lo = 0, cap_hi = 0;   // doubles received from the previous / next rank per apply
    double* arena = nullptr;         // exported allocation
    double* prev_arena = nullptr;    // neighbours' arenas mapped here
    double* next_arena = nullptr;
    size_t prev_cap_lo = 0, prev_cap_hi = 0, next_cap_lo = 0, next_cap_hi = 0;
    unsigned* d_ticket = nullptr;
    unsigned long long seq = 0;      // applies so far (identical on every rank: applies are collective)
    bool ready = false;
    static size_t flag_offset(size_t lo, size_t hi) { return (2 * (lo + hi) + 15) / 16 * 16; }  // in doubles
    static size_t arena_doubles(size_t lo, size_t hi) { return flag_offset(lo, hi) + 16; }
    void setup(Context* c, size_t lo, size_t hi);  // collective over all ranks
    void release();
    // entries for the previous rank (its "hi" side) and the next rank (its "lo" side); runs on ctx->comm_stream after
    // everything enqueued on ctx->stream so far, and starts the next apply (parity flips)
    void push(const double* to_prev, size_t n_to_prev, const double* to_next, size_t n_to_next);
    // For operators whose own kernel does the stores and the waiting: starts the next apply (parity flips) and returns
    // where its entries go, which flag words to release / poll, and the sequence number.  d_ticket[1], [2] are free for
    // that kernel's arrival counters.
    struct Targets {
        double* prev_dst;
        double* next_dst;
        unsigned long long* prev_flag;
        unsigned long long* next_flag;
        const unsigned long long* my_flags;
        unsigned long long seq;
    };
    Targets begin();
    void wait();  // on ctx->stream: returns (on the device) once both neighbours' entries of this apply have landed
    const double* lo() const { return arena + (seq & 1) * (cap_lo + cap_hi); }
    const double* hi() const { return arena + (seq & 1) * (cap_lo + cap_hi) + cap_lo; }
};

// ------------------------------------------------------------------------------------------ general gather exchange
// Row-block partitioned assembled operators with ARBITRARY sparsity (SparseObj.hpp:131-152 takes any pattern): a rank's
// rows may reference columns owned by any other rank.  Local column numbering is [own block | external entries in
// ascending global column order]; per apply every owner gathers the entries each requester needs from its x block and
//   peer mode: stores them straight into the requester's receive buffer of the apply's parity (NVLink stores, one kernel
//              for all destinations), fences, and the CTA that finishes last releases the sequence number into the flag
//              word it owns in every PARTNER's arena (partner = traffic in either direction, so the two-parity
//              write-after-read argument of PeerHalo holds pairwise);
//   NCCL mode: packs them and exchanges the packed ranges with grouped ncclSend / ncclRecv (also while a CUDA graph is
//              being captured).
struct GatherTable {
    double* dst[kMaxRanks];               // where destination d's entries go (null: nothing to send to that partner)
    int begin[kMaxRanks + 1];             // destination d serves send_idx[begin[d] .. begin[d+1])
    unsigned long long* flag[kMaxRanks];  // flag word to release per partner (null: none)
    int n_dst;
};
struct WaitTable {
    const unsigned long long* flag[kMaxRanks];
    int n;
};
struct PeerGather {
    Context* ctx = nullptr;
    int n_ext = 0;                          // external entries this rank receives per apply
    int n_send = 0;                         // entries this rank serves per apply (all destinations)
    std::vector<int> recv_count, recv_off;  // per source rank: how many, and where in the receive buffer
    std::vector<int> send_count, send_off;  // per destination rank: how many, and where in d_send_idx / d_pack
    std::vector<int> partners;
    int* d_send_idx = nullptr;              // [n_send] local row indices, grouped by destination rank
    double* d_pack = nullptr;               // [n_send] NCCL mode: packed entries
    double* nccl_ext = nullptr;             // [n_ext]  NCCL mode: receive buffer
    double* arena = nullptr;                // peer mode, exported: [buf parity 0 | buf parity 1 | flags (kMaxRanks words)]
    std::vector<double*> peer_arena;        // partners' arenas mapped here
    std::vector<size_t> peer_n_ext;         // their n_ext (layout of their arenas)
    std::vector<size_t> off_in_peer;        // where this rank's entries start inside partner p's receive buffer
    unsigned* d_ticket = nullptr;
    unsigned long long seq = 0;
    bool peer_ready = false;
    bool last_was_peer = false;
    static size_t flag_offset(size_t n) { return (2 * n + 15) / 16 * 16; }
    // ext_cols: this rank's external global columns, ascending (device, n_ext_ entries; a valid allocation even when empty);
    // blocks: [row_begin, row_end) of every rank.  Collective.
    void setup(Context* c, const int* d_ext_cols, int n_ext_, const std::vector<int>& ext_cols_host,
               const std::vector<std::pair<int, int>>& blocks);
    void release();
    // starts the exchange of an apply after everything enqueued on ctx->stream so far; returns the buffer the external
    // entries of THIS apply arrive in (valid for the kernels launched after wait())
    const double* begin(const double* x);
    void wait();
    // For operators whose own kernel gathers, stores, releases and waits (single-launch apply): true when the peer path can
    // be taken for the apply that starts now.  Starts it (parity flips) and fills the tables of its stores and waits;
    // *ext = the buffer its external entries arrive in.  d_ticket is free for that kernel's arrival counter.
    bool begin_inline(GatherTable* tb, WaitTable* wt, unsigned long long* seq_out, const double** ext);
};

// ------------------------------------------------------------------------------------------ fused CG state
// Device-resident scalars of the fused two-kernel CG iteration (atk_cg.cu drives it, the operator supplies KA).
struct CgFusedState {
    double rs_old;    // r.r of the current residual
    double alpha;     // alpha of the last r update; the matching x += alpha*p is still pending
    double last_pAp;
    double b_norm2;
    int ka_count;     // KA steps executed in this solve: the current search direction lives in P[ka_count & 1]
    int it;           // completed iterations (the reference's loop index)
    int done;
    int stopped_on_pAp;
    int zero_rhs;
    int first;        // 1 until the first KA of this solve has run: p is taken as is, no pending x update
    int trace_cap;
    unsigned long long epoch;  // peer-signalling sequence base of this solve: flags carry epoch + ka_count
    int peer_timeout;          // set when a wait on another rank's partial gave up (a rank died / fell out of step)
};

// ------------------------------------------------------------------------------------------ operators
struct Operator {
    Context* ctx = nullptr;
    int64_t n_rows_local = 0;  // rows owned by this rank
    int64_t n_cols = 0;        // global columns
    int64_t row_begin = 0;     // first global row owned
    int64_t nnz = 0;
    virtual ~Operator() {}
    // y = A x on the compute stream; if dot_slot >= 0 also leaves this rank's partial of x.y in that slot
    // (and all-gathers it).  `skip_flag` (device int, may be null): when *skip_flag != 0 the kernels return at once.
    virtual void apply(const double* x, double* y, int dot_slot, const int* skip_flag) = 0;
    // Fused CG step KA (matrix-free operators).  With p_cur = P[ka_count&1], p_alt the other buffer:
    //   p_new = r + beta*p_cur (beta = rr_slot / rs_old; p_new = p_cur on the first step), x += alpha*p_cur,
    //   Ap = A p_new, partial p_new.Ap into dot_slot, p_new written to p_alt; the last CTA commits the scalars
    //   (rs_old, it, ka_count).  The two buffers are needed because neighbouring CTAs still read p_cur.
    virtual bool fused_cg_capable() const { return false; }
    virtual void cg_fused_step(CgFusedState*, const double* /*r*/, double* /*x*/, double* /*Ap*/, double* /*P0*/,
                               double* /*P1*/, int /*rr_slot*/, int /*dot_slot*/, double* /*rs_trace*/) {
        throw Error(ATK_ERR_INVALID_ARGUMENT, "operator has no fused CG step");
    }
    // peer-memory mode of the fused step (distributed stencil operators on NVLink-connected GPUs)
    // Whole CG loop in ONE cooperative launch with x, r, p in registers (small single-GPU systems; see atk_stencil.cu).
    // cg_onchip_blocks: CTAs the launch would use, or 0 when this operator / size is not eligible.
    virtual int cg_onchip_blocks() const { return 0; }
    virtual void cg_onchip(CgFusedState*, double* /*x*/, double* /*r*/, double* /*p*/, double* /*p_alt*/, int /*max_iter*/,
                           double* /*pAp_trace*/, double* /*rs_trace*/) {
        throw Error(ATK_ERR_INVALID_ARGUMENT, "operator has no on-chip CG loop");
    }
    // Whole CG loop as ONE cooperative launch per rank at any size (phases separated by grid + cross-GPU reductions
    // instead of kernel boundaries; see atk_stencil.cu).  cg_persistent_blocks: CTAs of the launch, 0 when not eligible.
    virtual int cg_persistent_blocks() const { return 0; }
    virtual void cg_persistent(CgFusedState*, const double* /*b*/, double* /*x*/, double* /*r*/, double* /*Ap*/, double* /*P0*/,
                               double* /*P1*/, int /*max_iter*/, double* /*pAp_trace*/, double* /*rs_trace*/, int /*trace_cap*/,
                               unsigned long long /*epoch*/, unsigned long long* /*stamps*/, int /*stamp_cap*/) {
        throw Error(ATK_ERR_INVALID_ARGUMENT, "operator has no persistent CG loop");
    }
    virtual bool peer_capable() const { return false; }
    virtual void set_peer_mode(bool) {}
    virtual void peer_r_targets(double** lo, double** hi, int64_t* plane_elems) const { *lo = *hi = nullptr; *plane_elems = 0; }
    virtual void initial_halo_exchange(const double* /*r*/, const double* /*p*/) {}
    // device CSR arrays (ascending columns inside a row) of an assembled operator; false when there are none
    virtual bool device_csr(const int** /*row_ptr*/, const int** /*col*/, const double** /*val*/) const { return false; }
    virtual bool has_csr() const { return false; }
    virtual void export_csr(int*, int*, double*) const { throw Error(ATK_ERR_INVALID_ARGUMENT, "operator has no CSR arrays"); }
};

// balanced split of nz planes over `world` ranks: the first nz % world ranks get one extra plane
inline void slab_partition(int nz, int rank, int world, int* k_begin, int* k_end) {
    const int base = nz / world, rem = nz % world;
    *k_begin = rank * base + (rank < rem ? rank : rem);
    *k_end = *k_begin + base + (rank < rem ? 1 : 0);
}

Operator* make_operator_from_csc(Context* ctx, int n_rows, int n_cols, int64_t nnz, const int* col_ptr, const int* row_idx,
                                 const double* vals, bool on_device, atk_format format);
Operator* make_operator_from_csc_rows(Context* ctx, int n_rows_global, int n_cols, int row_begin, int row_end, int64_t nnz,
                                      const int* col_ptr, const int* row_idx, const double* vals, bool on_device, atk_format format);
Operator* make_stencil7(Context* ctx, int nx, int ny, int nz, double cx, double cy, double cz);
Operator* make_poisson_assembled(Context* ctx, int nx, int ny, int nz, double cx, double cy, double cz, atk_format format);
Operator* make_stencil27(Context* ctx, int nx, int ny, int nz, const double* coef);
Operator* make_stencil27_assembled(Context* ctx, int nx, int ny, int nz, const double* coef, atk_format format);

// ------------------------------------------------------------------------------------------ vector kernels (host launchers)
void launch_dot(Context* ctx, int64_t n, const double* x, const double* y, int slot, const int* skip_flag = nullptr);
void launch_scale(Context* ctx, int64_t n, const double* x, double s, double* out);
void launch_div(Context* ctx, int64_t n, const double* x, double s, double* out);
void launch_add(Context* ctx, int64_t n, const double* x, const double* y, double* out);
void launch_sub(Context* ctx, int64_t n, const double* x, const double* y, double* out);
void launch_axpy(Context* ctx, int64_t n, const double* x, const double* y, double a, double* out);
void launch_axmy(Context* ctx, int64_t n, const double* x, const double* y, double a, double* out);

// ------------------------------------------------------------------------------------------ solvers
void cg_solve(Context* ctx, Operator* op, const double* b, double* x, double* r, double* p, int max_iter, atk_cg_info* info);
void gd_solve(Context* ctx, Operator* op, const double* b, double* x, int max_iter, double tol, atk_gd_info* info);
void arnoldi(Context* ctx, Operator* op, double* Q, int64_t ldq, int m, double* H, double tol, int* breakdown);
// Left preconditioner handle: x = M^-1 b on ctx->stream (a set skip_flag turns the launch into a no-op).
struct Preconditioner {
    Context* ctx = nullptr;
    int64_t n = 0;
    virtual ~Preconditioner() {}
    virtual void solve(const double* b, double* x, const int* skip_flag) = 0;
    virtual void export_factors(double* L, double* U) = 0;  // dense n x n column-major host arrays (either may be null)
};
Preconditioner* make_ilu(Context* ctx, int64_t n_rows, int64_t n_cols, int64_t nnz, const int32_t* col_ptr, const int32_t* row_idx,
                         const double* vals);
// sparse ILU(0) (the reference's elimination restricted to the pattern of A) with level-scheduled solves: atk_ilu0.cu
Preconditioner* make_ilu0(Context* ctx, int64_t n_rows, int64_t n_cols, int64_t nnz, const int32_t* col_ptr, const int32_t* row_idx,
                          const double* vals);
void gmres_solve(Context* ctx, Operator* op, Preconditioner* pc, const double* b, double* x, int64_t n_x, int max_iter,
                 int krylov_dim, double tol, atk_gmres_ortho ortho, atk_gmres_callback cb, void* user, atk_gmres_info* info);

// ------------------------------------------------------------------------------------------ device helpers
#ifdef __CUDACC__

// Multiply and add that can never be contracted into an FMA (the reference is built without FMA).
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }

__device__ __forceinline__ unsigned linear_block_id() {
    return blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
}
__device__ __forceinline__ unsigned total_blocks() { return gridDim.x * gridDim.y * gridDim.z; }
__device__ __forceinline__ unsigned linear_thread_id() {
    return threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
}

// Fixed-shape block sum: xor-shuffle tree inside each warp, then the warp results added in warp order.
// Every thread returns the same value.  BLOCK = threads per block (multiple of 32).
template <int BLOCK>
__device__ __forceinline__ double block_sum(double v) {
    __shared__ double warp_part[BLOCK / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = add_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
    const unsigned tid = linear_thread_id();
    __syncthreads();  // protect warp_part against a previous use
    if ((tid & 31) == 0) warp_part[tid >> 5] = v;
    __syncthreads();
    double t = warp_part[0];
#pragma unroll
    for (int w = 1; w < BLOCK / 32; ++w) t = add_rn(t, warp_part[w]);
    return t;
}

// Sum of `count` per-block partials by one CTA, in an order that depends only on (count, BLOCK): every thread runs
// four independent accumulation chains over an interleaved subsequence (so the L2 loads pipeline), the chains are
// combined in a fixed order, then the fixed-shape block tree finishes.
template <int BLOCK>
__device__ __forceinline__ double partials_sum(const double* partials, unsigned count) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    unsigned i = linear_thread_id();
    for (; i + 3 * BLOCK < count; i += 4 * BLOCK) {
        const double v0 = __ldcg(partials + i), v1 = __ldcg(partials + i + BLOCK);
        const double v2 = __ldcg(partials + i + 2 * BLOCK), v3 = __ldcg(partials + i + 3 * BLOCK);
        s0 = add_rn(s0, v0);
        s1 = add_rn(s1, v1);
        s2 = add_rn(s2, v2);
        s3 = add_rn(s3, v3);
    }
    for (; i < count; i += BLOCK) s0 = add_rn(s0, __ldcg(partials + i));
    return block_sum<BLOCK>(add_rn(add_rn(s0, s1), add_rn(s2, s3)));
}

// Grid-wide deterministic sum: each block leaves its partial in partials[block]; the block that arrives last adds
// all partials in a fixed order.  Returns true (for every thread) in that last block, with the total in *total.
// The ticket wraps back to 0 by itself (atomicInc), so it is ready for the next launch.
template <int BLOCK>
__device__ __forceinline__ bool grid_sum_last_block(double v, double* partials, unsigned* ticket, double* total) {
    __shared__ bool is_last;
    const double bs = block_sum<BLOCK>(v);
    const unsigned nb = total_blocks();
    if (linear_thread_id() == 0) {
        partials[linear_block_id()] = bs;
        __threadfence();
        const unsigned t = atomicInc(ticket, nb - 1);
        is_last = (t == nb - 1);
    }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
    *total = partials_sum<BLOCK>(partials, nb);
    return true;
}

// The same for a sum spread over several launches on one stream: every launch leaves its blocks' sums at
// partials[partial_base + block]; only the launch marked final takes tickets, and its last block adds ALL partials
// [0, partial_base + blocks of this launch) in index order (the earlier launches have completed: stream order).
template <int BLOCK>
__device__ __forceinline__ bool grid_sum_last_block_chained(double v, double* partials, int partial_base, int is_final,
                                                            unsigned* ticket, double* total) {
    __shared__ bool is_last_c;
    const double bs = block_sum<BLOCK>(v);
    const unsigned nb = total_blocks();
    if (linear_thread_id() == 0) {
        partials[partial_base + linear_block_id()] = bs;
        is_last_c = false;
        if (is_final) {
            __threadfence();
            const unsigned t = atomicInc(ticket, nb - 1);
            is_last_c = (t == nb - 1);
        }
    }
    __syncthreads();
    if (!is_last_c) return false;
    __threadfence();
    *total = partials_sum<BLOCK>(partials, partial_base + nb);
    return true;
}

// Software grid barrier for cooperative launches (all CTAs co-resident): `counter` starts at 0 at launch, every CTA keeps
// its own running `target`.
__device__ __forceinline__ void coop_grid_barrier(unsigned* counter, unsigned& target) {
    __syncthreads();
    if (gridDim.x > 1) {
        if (threadIdx.x == 0) {
            target += gridDim.x;
            __threadfence();
            atomicAdd(counter, 1u);
            while (*reinterpret_cast<volatile unsigned*>(counter) < target) {}
            __threadfence();
        }
        __syncthreads();
    }
}

// Election without a sum: true in the block that finishes last.
__device__ __forceinline__ bool last_block_done(unsigned* ticket) {
    __shared__ bool is_last_e;
    __syncthreads();
    if (linear_thread_id() == 0) {
        __threadfence();
        const unsigned nb = total_blocks();
        const unsigned t = atomicInc(ticket, nb - 1);
        is_last_e = (t == nb - 1);
    }
    __syncthreads();
    return is_last_e;
}

// The same among the first `nb` blocks of a launch only (the others never call it).
__device__ __forceinline__ bool last_block_done_among(unsigned* ticket, unsigned nb) {
    __shared__ bool is_last_a;
    __syncthreads();
    if (linear_thread_id() == 0) {
        __threadfence();
        const unsigned t = atomicInc(ticket, nb - 1);
        is_last_a = (t == nb - 1);
    }
    __syncthreads();
    return is_last_a;
}

// ---- peer signalling primitives ----
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// Scalars travel between GPUs as two self-validating 8-byte words per (slot, source rank): {low half | seq32 << 32} and
// {high half | seq32 << 32}.  An aligned 8-byte store is delivered whole, so a word whose upper half equals the expected
// sequence number carries valid data and no fence is needed on the value path (the scheme of NCCL's low-latency protocol).
// Producer: first warp of the CTA that finished last; lane q serves rank q.  Ordering of the HALO planes other CTAs
// stored into the peers is given by the system-scope fence every CTA executes before it takes its ticket: those stores
// are performed before the ticket is visible, hence before this function is even entered.
// 32-bit tag of a sequence number: the top bit is always set, so a word that still holds its initial zeros can never
// validate, whatever the sequence number (the low 31 bits wrap around freely; tags are only compared for equality).
__device__ __forceinline__ unsigned long long ll_tag(unsigned long long seq) { return 0x80000000ull | (seq & 0x7fffffffull); }
__device__ __forceinline__ void publish_partial_warp(unsigned long long* const* peer_ll, int world, int rank, int slot,
                                                     double value, unsigned long long seq) {
    const int lane = threadIdx.x & 31;
    if (lane < world) {
        const unsigned long long bits = (unsigned long long)__double_as_longlong(value);
        const unsigned long long tag = ll_tag(seq) << 32;
        unsigned long long* dst = peer_ll[lane] + (size_t)(slot * kMaxRanks + rank) * 2;
        st_relaxed_sys(dst, (bits & 0xffffffffull) | tag);
        st_relaxed_sys(dst + 1, (bits >> 32) | tag);
    }
}
// Consumer: waits until every rank's pair of `slot` carries `seq`, then returns the rank-ordered sum (same bits on every
// rank).  Lane q of the first warp polls rank q's pair in LOCAL memory; the fence afterwards orders this CTA's later reads
// of peer-written halo planes behind the observation.
//
// A wait that lasts longer than kPeerWaitNs (a rank crashed or the ranks fell out of step) gives up instead of hanging the
// GPU: it raises *timeout_flag, every later kernel of the solve turns into a no-op and the host reports ATK_ERR_RUNTIME.
constexpr unsigned long long kPeerWaitNs = 10ull * 1000ull * 1000ull * 1000ull;
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ double peer_wait_sum(const unsigned long long* ll, int world, int slot, unsigned long long seq,
                                                int* timeout_flag) {
    __shared__ double peer_vals[kMaxRanks];
    if (linear_thread_id() < 32) {
        const int lane = linear_thread_id();
        if (lane < world) {
            const unsigned long long* src = ll + (size_t)(slot * kMaxRanks + lane) * 2;
            const unsigned long long tag = ll_tag(seq);
            unsigned long long w0, w1;
            unsigned spins = 0;
            unsigned long long t0 = 0;
            for (;;) {
                w0 = ld_relaxed_sys(src);
                w1 = ld_relaxed_sys(src + 1);
                if ((w0 >> 32) == tag && (w1 >> 32) == tag) break;
                if ((++spins & 0xfffu) == 0) {  // look at the clock only now and then
                    const unsigned long long now = global_timer_ns();
                    if (t0 == 0) t0 = now;
                    if (now - t0 > kPeerWaitNs || *reinterpret_cast<volatile int*>(timeout_flag)) {
                        *timeout_flag = 1;
                        break;
                    }
                }
            }
            peer_vals[lane] = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
        }
        __syncwarp();
        __threadfence_system();
    }
    __syncthreads();
    double s = peer_vals[0];
    for (int q = 1; q < world; ++q) s = add_rn(s, peer_vals[q]);
    __syncthreads();  // peer_vals may be reused by a later call
    return s;
}

// Sum of a reduction slot over ranks, in rank order (what every consumer kernel does in its prologue).
__device__ __forceinline__ double slot_sum(const double* slot, int world) {
    double s = __ldcg(slot);
    for (int q = 1; q < world; ++q) s = add_rn(s, __ldcg(slot + q));
    return s;
}

#endif  // __CUDACC__

}  // namespace atk

struct atk_context_s {
    atk::Context c;
};
struct atk_operator_s {
    atk::Operator* op;
};
