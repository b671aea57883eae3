/* include/atk_cuda.h -- C ABI of libatk_cuda.so: the B200 (sm_100a) device path behind
 * AlgoToolkit's sparse iterative-solver hot path.
 *
 * The reference (Yin169/AlgoToolkit) has no FFI seam: its "operator API" is a set of C++17
 * templates plus a pybind11 module.  The seam is created here: the double instantiations of
 * the reference's L0 operators and its three solve() loops become calls into this library,
 * and the reference-named host classes in algotoolkit_b200/host/ (VectorObj, SparseMatrixCSC,
 * ConjugateGrad, GMRES, Poisson3DSolver, utils::readMatrixMarket, the `fastsolver` module)
 * are thin owners of host std::vectors above it.  INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - plain C, opaque handles, caller-owned buffers, no exceptions, no C++/torch types;
 *   - every function returns an atk_status; the codes 1..3 map 1:1 onto the exception types
 *     the reference throws (std::invalid_argument / std::runtime_error / std::out_of_range);
 *     atk_last_error() returns the message of the last failure on the calling thread;
 *   - `const double* x` style vector arguments are DEVICE pointers (from atk_malloc, or any
 *     CUDA allocation such as a torch tensor's data_ptr()) unless the function name ends in
 *     `_host`, in which case they are host pointers and the call does its own H2D/D2H;
 *   - all arithmetic is IEEE fp64, multiply and add rounded separately (the reference is
 *     built without FMA), indices are int32 (nnz < 2^31 at every BASELINE config);
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     ATK_ERR_CUDA.
 *
 * Reference paths below are relative to the reference root (/root/reference).
 */
#ifndef ATK_CUDA_H
#define ATK_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ATK_ABI_VERSION 2

typedef enum atk_status {
    ATK_OK = 0,
    ATK_ERR_INVALID_ARGUMENT = 1, /* std::invalid_argument: SparseObj.hpp:133, GMRES.hpp:30-36, VectorObj.hpp:90,98,106 */
    ATK_ERR_RUNTIME = 2,          /* std::runtime_error:    VectorObj.hpp:76,83 (divide by zero), GMRES.hpp:163 */
    ATK_ERR_OUT_OF_RANGE = 3,     /* std::out_of_range:     VectorObj.hpp:38,43, SparseObj.hpp:69,107, Poisson.hpp:162 */
    ATK_ERR_CUDA = 4,             /* CUDA runtime failure, or no device */
    ATK_ERR_NCCL = 5,             /* NCCL failure / libnccl not loadable */
    ATK_ERR_NOMEM = 6
} atk_status;

typedef struct atk_context_s* atk_context_t;   /* device, streams, reduction scratch, NCCL communicator */
typedef struct atk_operator_s* atk_operator_t; /* a device-resident linear operator (row block of it when distributed) */

/* How dot / norm reductions are ordered. */
typedef enum atk_reduction {
    ATK_REDUCE_TREE = 0,       /* fixed-shape block tree + fixed-order final pass: deterministic run to run */
    ATK_REDUCE_SEQUENTIAL = 1  /* strictly sequential from 0.0 in index order: bit-identical to std::inner_product
                                  (VectorObj.hpp:52-54,105-108); one thread does the adds, for parity runs only */
} atk_reduction;

/* Device storage format of an assembled matrix. */
typedef enum atk_format {
    ATK_FORMAT_SELL = 0,       /* SELL-32-sigma, one thread per row, taps added in ascending column order from 0.0:
                                  bit-identical to the CSC matvec of SparseObj.hpp:131-152 */
    ATK_FORMAT_CSR_VECTOR = 1  /* CSR, 2..32 lanes per row + shuffle tree: deterministic, not reference-ordered */
} atk_format;

/* ------------------------------------------------------------------------------------------------ context */
int atk_abi_version(void);
const char* atk_last_error(void);
int atk_device_count(int* count);

/* One context per process and GPU. */
int atk_context_create(int device, atk_context_t* out);
/* Distributed context: `world` processes, one GPU each, row-slab partition.  `nccl_uid` is the blob produced by
 * atk_nccl_unique_id() on rank 0 and broadcast by the caller (torch.distributed, MPI, a file...). */
int atk_nccl_unique_id(void* buf, size_t capacity, size_t* bytes);
int atk_context_create_distributed(int device, int rank, int world, const void* nccl_uid, size_t uid_bytes,
                                   atk_context_t* out);
/* Destroy the operators and preconditioners created on a context before the context itself. */
int atk_context_destroy(atk_context_t ctx);
int atk_context_rank(atk_context_t ctx, int* rank, int* world);
int atk_context_set_reduction(atk_context_t ctx, atk_reduction mode);
int atk_context_synchronize(atk_context_t ctx);
/* Number of kernels this library has launched on the context since creation (for bench.py's gpu_launches). */
int atk_context_launch_count(atk_context_t ctx, int64_t* launches);

/* ------------------------------------------------------------------------------------------------ memory */
int atk_malloc(atk_context_t ctx, size_t bytes, void** dptr);
int atk_free(atk_context_t ctx, void* dptr);
int atk_memcpy_h2d(atk_context_t ctx, void* dst_device, const void* src_host, size_t bytes);
int atk_memcpy_d2h(atk_context_t ctx, void* dst_host, const void* src_device, size_t bytes);
int atk_memcpy_d2d(atk_context_t ctx, void* dst_device, const void* src_device, size_t bytes);
int atk_memset_zero(atk_context_t ctx, void* dptr, size_t bytes);
/* Pinned host memory for the `_host` entry points.  The pages are placed on the NUMA node the CURRENT CUDA device hangs
 * off (mmap + mbind + cudaHostRegister; plain cudaMallocHost when the node is unknown), so that with one process per GPU
 * the 8 concurrent uploads of a box do not all cross the socket interconnect. */
int atk_host_alloc(size_t bytes, void** hptr);
int atk_host_free(void* hptr);
/* Diagnostics: NUMA node the first page of `hptr` lives on and the node of the current CUDA device (-1 = unknown). */
int atk_host_numa_node(const void* hptr, int* buffer_node, int* device_node);

/* ------------------------------------------------------------------------------------------------ VectorObj ops
 * Replaces src/Obj/VectorObj.hpp for T=double.  n is the LOCAL length; on a distributed context dot/nrm2 return
 * the global value (rank-ordered sum of the per-rank partials, identical on every rank). */
int atk_dot(atk_context_t ctx, int64_t n, const double* x, const double* y, double* result_host);   /* :105-108 */
int atk_nrm2(atk_context_t ctx, int64_t n, const double* x, double* result_host);                   /* :52-54   */
int atk_vec_scale(atk_context_t ctx, int64_t n, const double* x, double s, double* out);            /* :64-68  x*s */
int atk_vec_div(atk_context_t ctx, int64_t n, const double* x, double s, double* out);              /* :75-80  x/s; s==0 -> ATK_ERR_RUNTIME */
int atk_vec_add(atk_context_t ctx, int64_t n, const double* x, const double* y, double* out);       /* :89-94   */
int atk_vec_sub(atk_context_t ctx, int64_t n, const double* x, const double* y, double* out);       /* :97-102  */
/* out = x + y*a with the product rounded first (the `x + p*alpha` idiom of ConjugateGradient.hpp:60,79) */
int atk_vec_axpy(atk_context_t ctx, int64_t n, const double* x, const double* y, double a, double* out);
/* out = x - y*a (ConjugateGradient.hpp:61, GMRES.hpp:76) */
int atk_vec_axmy(atk_context_t ctx, int64_t n, const double* x, const double* y, double a, double* out);

/* ------------------------------------------------------------------------------------------------ operators */
/* SparseMatrixCSC<double> (src/Obj/SparseObj.hpp:22-39): takes the finished CSC arrays (values, row_indices,
 * col_ptr after finalize(), :76-100) and converts them ON THE DEVICE to CSR (stable in column order) and then to
 * the requested format.  `on_device` != 0 means the three arrays are already device pointers.
 * On a distributed context (SELL format, square matrices): every rank passes the WHOLE matrix and keeps a balanced block
 * of rows (atk_operator_row_range tells which; the grid operators below keep whole z-planes); vectors are the local row
 * blocks.  ANY sparsity is accepted, as in the reference matvec: a rank's rows may reference columns owned by any other
 * rank; the owners gather those entries and store them into the requester's memory (NVLink peer stores + sequence flags,
 * grouped ncclSend/ncclRecv where the GPUs cannot map each other) while the rows that need none of them are computed. */
int atk_operator_from_csc(atk_context_t ctx, int n_rows, int n_cols, int64_t nnz, const int* col_ptr,
                          const int* row_idx, const double* vals, int on_device, atk_format format,
                          atk_operator_t* out);
/* The same from a ROW SLICE per rank, so that no rank ever holds the whole matrix: the CSC arrays contain only the entries
 * of rows [row_begin, row_end) (global row numbers in row_idx, col_ptr over all n_cols columns, nnz = entries of the
 * slice).  The ranks' blocks must tile [0, n_rows) in rank order (ATK_ERR_INVALID_ARGUMENT otherwise; a row outside the
 * block is ATK_ERR_OUT_OF_RANGE).  On a single-GPU context the slice must be the whole matrix. */
int atk_operator_from_csc_rows(atk_context_t ctx, int n_rows, int n_cols, int row_begin, int row_end, int64_t nnz,
                               const int* col_ptr, const int* row_idx, const double* vals, int on_device, atk_format format,
                               atk_operator_t* out);
/* Matrix-free operator of Poisson3DSolver::buildLaplacianMatrix (src/PDEs/FDM/Poisson.hpp:49-95): identity on the
 * Dirichlet shell, 7 taps (cz,cy,cx,cc=-2(cx+cy+cz),cx,cy,cz) inside, idx = i + j*nx + k*nx*ny (:98-100).
 * On a distributed context the operator is the rank's z-slab (planes [k_begin,k_end) of nz, balanced split). */
int atk_operator_stencil7(atk_context_t ctx, int nx, int ny, int nz, double cx, double cy, double cz,
                          atk_operator_t* out);
/* The same Poisson operator ASSEMBLED: the CSC arrays buildLaplacianMatrix+finalize would produce are generated on
 * the device (never on the host) and sent through the same CSC->CSR->format conversion as atk_operator_from_csc. */
int atk_operator_poisson_assembled(atk_context_t ctx, int nx, int ny, int nz, double cx, double cy, double cz,
                                   atk_format format, atk_operator_t* out);
/* Synthetic 27-point operator of BASELINE config 5: open boundaries, coef[13] on the diagonal and
 * coef[(dk+1)*9+(dj+1)*3+(di+1)] for neighbour (i+di,j+dj,k+dk); generated on the device. */
/* The same 27-point operator MATRIX-FREE (bit-identical results); on a distributed context it is the rank's z-slab with
 * one halo plane exchanged per apply, like atk_operator_stencil7. */
int atk_operator_stencil27(atk_context_t ctx, int nx, int ny, int nz, const double coef[27], atk_operator_t* out);
int atk_operator_stencil27_assembled(atk_context_t ctx, int nx, int ny, int nz, const double coef[27],
                                     atk_format format, atk_operator_t* out);
/* The balanced z-slab split used by the distributed stencil operators: planes [k_begin, k_end) of nz for `rank` of
 * `world` (the first nz % world ranks own one extra plane).  Pure host arithmetic, needs no device. */
int atk_slab_partition(int nz, int rank, int world, int* k_begin, int* k_end);
int atk_operator_destroy(atk_operator_t op);
/* local rows, columns (global), stored entries (without padding) */
int atk_operator_shape(atk_operator_t op, int64_t* n_rows_local, int64_t* n_cols, int64_t* nnz);
/* first global row owned by this rank and the local row count (0, n_rows on a single-GPU context) */
int atk_operator_row_range(atk_operator_t op, int64_t* row_begin, int64_t* n_rows_local);
/* CSR arrays of an assembled operator copied to HOST buffers (row_ptr[n_rows+1], col_idx[nnz], vals[nnz]); lets
 * tests check the device conversion against the oracle. */
int atk_operator_export_csr(atk_operator_t op, int* row_ptr, int* col_idx, double* vals);
/* y = A*x  (SparseMatrixCSC::operator*(VectorObj), SparseObj.hpp:131-152).  x and y are local row blocks. */
int atk_operator_apply(atk_context_t ctx, atk_operator_t op, const double* x, double* y);

/* ------------------------------------------------------------------------------------------------ ConjugateGrad */
typedef struct atk_cg_info {
    int iterations;        /* loop index at which the solver stopped: == max_iter, or the i at which |p.Ap| < 1e-12 */
    int stopped_on_pAp;    /* 1 if the |p.Ap| < 1e-12 exit fired (ConjugateGradient.hpp:53-57) */
    int zero_rhs;          /* 1 if ||b|| < 1e-12 returned immediately (:38-42) */
    double last_pAp;       /* p.Ap of the last iteration entered */
    double rs;             /* r.r at exit */
    float device_ms;       /* CUDA-event time of the device loop, excluding any host<->device copies */
    int kernel_launches;   /* kernels this call launched */
    /* optional per-iteration traces (host pointers, may be NULL), filled up to trace_capacity entries */
    double* pAp_trace;
    double* rs_trace;
    int trace_capacity;
    /* in: when non-zero the iterations are launched kernel by kernel (no CUDA graph) with CUDA events around each
     * kernel class, and kernel_ms[] returns the summed device time of K1 (operator apply + p.Ap), K2 (x/r update +
     * r.r) and K3 (p update) over all iterations; used by bench.py for the per-kernel roofline */
    int time_kernels;
    float kernel_ms[3];
    /* out: 1 when the fused two-kernel formulation ran (matrix-free operator, tree reductions): kernel_ms[0] is then
     * KA (p update + pending x update + operator apply + p.Ap, 48 B/row), kernel_ms[1] KB (r update + r.r, 24 B/row) */
    int fused;
    /* out: 1 when the iteration ran in peer-memory mode on a distributed context: halo planes and partial sums move by
     * direct stores into the neighbouring GPUs' memory (CUDA IPC over NVLink) and sequence flags, no NCCL call inside
     * the loop.  0: NCCL send/recv + all-gather (also selectable with ATK_COMM=nccl). */
    int peer;
    /* out: 1 when the whole loop ran as ONE persistent cooperative launch per rank (matrix-free operator, tree reductions,
     * 16-byte aligned vectors; single GPU or peer-memory mode): phases A (= KA) and B (= KB) are separated by grid-wide /
     * cross-GPU reductions inside the kernel instead of kernel boundaries.  With time_kernels set, kernel_ms[0] / [1] are
     * then the summed phase A / phase B times (their reductions included) from in-kernel %globaltimer stamps.
     * ATK_CG_PERSISTENT=0 selects the CUDA-graph loop of two kernels per iteration instead. */
    int persistent;
    /* out (atk_cg_solve_host only): wall-clock milliseconds of the phases of the call - host->device copies (b and the
     * state), device loop incl. its setup, device->host copies - measured with CUDA events on the context's stream */
    float h2d_ms, solve_ms, d2h_ms;
} atk_cg_info;

/* ConjugateGrad::solve (src/LinearAlgebra/Krylov/ConjugateGradient.hpp:36-84).  The solver state x, r, p lives in
 * caller-owned device vectors and is updated in place, so a second call continues where the first stopped, as the
 * reference's public members do.  A fresh solver is x=0, r=b, p=b (what the ctor leaves, :26-30).  There is no
 * tolerance test (the reference's is commented out, :44-47,75-77). */
int atk_cg_solve(atk_context_t ctx, atk_operator_t op, const double* b, double* x, double* r, double* p,
                 int max_iter, atk_cg_info* info);
/* Same through HOST buffers: uploads b and the state, runs, downloads x (and r, p when non-NULL).
 * `fresh` != 0 ignores the incoming x/r/p contents and starts from x=0, r=p=b. */
int atk_cg_solve_host(atk_context_t ctx, atk_operator_t op, const double* b, double* x, double* r, double* p,
                      int fresh, int max_iter, atk_cg_info* info);

/* ------------------------------------------------------------------------------------------------ GMRES */
typedef enum atk_gmres_event {
    ATK_GMRES_INITIAL_RESIDUAL = 0, /* GMRES.hpp:48  "Initial residual norm: " */
    ATK_GMRES_INITIAL_GOOD = 1,     /* :50           "Initial guess is good enough." */
    ATK_GMRES_KRYLOV_UPDATE = 2,    /* :83           "KrylovUpdate : " (value = j) */
    ATK_GMRES_RESTART_RESIDUAL = 3, /* :105          "Residual norm after restart: " */
    ATK_GMRES_CONVERGED = 4,        /* :108          "Converged after restart at iteration " (value = iter) */
    ATK_GMRES_MAX_ITER = 5          /* :114          "Reached maximum iterations without convergence." */
} atk_gmres_event;
/* progress callback: the host class turns these into the reference's exact std::cout lines */
typedef void (*atk_gmres_callback)(int event, double value, void* user);

typedef enum atk_gmres_ortho {
    ATK_GMRES_MGS = 0,         /* the reference's loop: for i<j { h = V[i].w; w = w - V[i]*h } (GMRES.hpp:73-77), evaluated
                                  projection by projection (persistent on-chip kernel where the slices fit) - parity path */
    ATK_GMRES_MGS_BATCHED = 1  /* the same projection with two passes over V[0:j] and two reductions per Arnoldi step:
                                  c = V^T w, g = V^T v_j (Gram row), solve (I + strict_lower(G)) h = c, w -= V h.
                                  Mathematically identical to the loop above even for the reference's non-orthonormal
                                  basis; rounding differs at the 1e-14 level on well-conditioned operators.  Plain
                                  classical Gram-Schmidt (h = c) is NOT equivalent to the reference and is not offered. */
} atk_gmres_ortho;

typedef struct atk_gmres_info {
    int restarts;          /* restart cycles run */
    int status;            /* 0 reached max_iter, 1 converged after a restart, 2 initial guess good enough */
    int arnoldi_steps;     /* total Arnoldi steps executed */
    double initial_residual;
    double final_residual;
    float device_ms;
    int kernel_launches;
    double* residual_trace; /* optional host buffer: [0]=initial, then one per restart */
    int trace_capacity;
    int trace_count;
} atk_gmres_info;

/* GMRES::solve (src/LinearAlgebra/Krylov/GMRES.hpp:27-115), unpreconditioned, with the reference's behaviour:
 * projection against V[0..j-1] only (:73), H allocated once per call and never cleared (:55), no inner convergence
 * test, absolute tolerance, Givens from the current H(j,j) (:89-94), x += V[i]*y[i] added in ascending i (:167-170).
 * x is the initial guess and the result.  Errors: ATK_ERR_INVALID_ARGUMENT for the checks at :28-37,
 * ATK_ERR_RUNTIME for a zero pivot (:162-164) or a division by zero (VectorObj.hpp:76). */
int atk_gmres_solve(atk_context_t ctx, atk_operator_t op, const double* b, double* x, int64_t n_x, int max_iter,
                    int krylov_dim, double tol, atk_gmres_ortho ortho, atk_gmres_callback cb, void* user,
                    atk_gmres_info* info);
int atk_gmres_solve_host(atk_context_t ctx, atk_operator_t op, const double* b, double* x, int64_t n_x, int max_iter,
                         int krylov_dim, double tol, atk_gmres_ortho ortho, atk_gmres_callback cb, void* user,
                         atk_gmres_info* info);

/* ------------------------------------------------------------------------------------------------ "next" rows
 * compositions of the same kernels. */
typedef struct atk_gd_info {
    int iterations;        /* iterations whose update was applied */
    double residual_norm;  /* ||r|| after the last update */
    double change_norm;    /* ||x - xOld|| of the last update */
    float device_ms;
    int kernel_launches;
} atk_gd_info;
/* GradientDescent::solve (src/LinearAlgebra/Solver/IterSolver.hpp:23-44): steepest descent with the reference's loop
 * condition (residualNorm > tol) and early exit (||x - xOld|| < tol).  x is the initial guess and the result. */
int atk_gd_solve(atk_context_t ctx, atk_operator_t op, const double* b, double* x, int max_iter, double tol,
                 atk_gd_info* info);
int atk_gd_solve_host(atk_context_t ctx, atk_operator_t op, const double* b, double* x, int max_iter, double tol,
                      atk_gd_info* info);
/* Krylov::Arnoldi (src/LinearAlgebra/Krylov/KrylovSubspace.hpp:11-38): textbook MGS Arnoldi (projects against ALL
 * previous vectors).  Q holds m columns of the local length, `ldq` doubles apart, Q[0] given, the others written;
 * H is a HOST array, m x (m-1) column-major, of which only the entries the reference writes are touched;
 * *breakdown_at = i when ||A q_(i-1) - ...|| < tol stopped the recursion at step i (Q[i] untouched), else -1. */
int atk_arnoldi(atk_context_t ctx, atk_operator_t op, double* Q, int64_t ldq, int m, double* H, double tol,
                int* breakdown_at);
int atk_arnoldi_host(atk_context_t ctx, atk_operator_t op, double* Q, int64_t ldq, int m, double* H, double tol,
                     int* breakdown_at);

/* ------------------------------------------------------------------------------------------------ "next" row
 * the enablePreconditioner() branch of GMRES::solve.
 *
 * ILUPreconditioner (src/LinearAlgebra/Preconditioner/ILU.hpp).  atk_ilu_compute = compute() (:21-81) on finished CSC
 * arrays (host pointers): the reference's full elimination without pivoting with its 1e-12 drop rules, carried out on
 * the device on a dense n x n array - the factors are bit-identical to getLFactor()/getUFactor().  Errors:
 * ATK_ERR_INVALID_ARGUMENT "Matrix must be square for ILU factorization" (:24-26), ATK_ERR_RUNTIME "Matrix is
 * numerically singular" (:38-40).  Single-GPU contexts, n <= 13500.
 * atk_ilu_solve = solve() (:84-112): x = U^-1 L^-1 b.  The forward sweep always adds in the reference's order; the
 * backward sweep does so under ATK_REDUCE_SEQUENTIAL (bit-exact) and accumulates column by column under
 * ATK_REDUCE_TREE (same result up to rounding, far less serial).  n must equal the factor size
 * (ATK_ERR_INVALID_ARGUMENT "Vector size does not match matrix size", :88-90).
 * atk_ilu_export: L and U as dense n x n column-major host arrays (either may be NULL). */
typedef struct atk_preconditioner* atk_preconditioner_t;
int atk_ilu_compute(atk_context_t ctx, int64_t n_rows, int64_t n_cols, int64_t nnz, const int32_t* col_ptr,
                    const int32_t* row_idx, const double* vals, atk_preconditioner_t* out);
/* Sparse ILU(0) - what SURVEY 8(f) rank 4 names as the next step: the SAME elimination (loops and 1e-12 drop rules of
 * ILU.hpp:21-81) restricted to the positions A stores, i.e. without fill-in, for systems of any size; level-scheduled
 * factorisation (one launch per level) and triangular solves (one cooperative launch for both sweeps, a grid barrier per
 * level), every sum in ascending column order.  The reference has no such factorisation; on a matrix whose full
 * elimination creates no fill-in (tridiagonal ...) the factors and solves are bit-identical to atk_ilu_compute / the
 * reference.  The handle works with atk_ilu_solve[_host], atk_ilu_export and atk_gmres_solve_precond[_host].  Same errors
 * as atk_ilu_compute (a missing diagonal entry counts as a zero pivot).  Single-GPU contexts. */
int atk_ilu0_compute(atk_context_t ctx, int64_t n_rows, int64_t n_cols, int64_t nnz, const int32_t* col_ptr,
                     const int32_t* row_idx, const double* vals, atk_preconditioner_t* out);
int atk_ilu_solve(atk_context_t ctx, atk_preconditioner_t pc, int64_t n, const double* b, double* x);      /* device */
int atk_ilu_solve_host(atk_context_t ctx, atk_preconditioner_t pc, int64_t n, const double* b, double* x); /* host */
int atk_ilu_export(atk_preconditioner_t pc, double* L, double* U);
int atk_preconditioner_destroy(atk_preconditioner_t pc);
/* GMRES::solve with usePreconditioner (GMRES.hpp:39-45,67-68,100-102): r = M^-1 (b - A x), w = M^-1 (A v_j); everything
 * else as atk_gmres_solve.  pc == NULL is the unpreconditioned solve. */
int atk_gmres_solve_precond(atk_context_t ctx, atk_operator_t op, atk_preconditioner_t pc, const double* b, double* x,
                            int64_t n_x, int max_iter, int krylov_dim, double tol, atk_gmres_ortho ortho,
                            atk_gmres_callback cb, void* user, atk_gmres_info* info);
int atk_gmres_solve_precond_host(atk_context_t ctx, atk_operator_t op, atk_preconditioner_t pc, const double* b, double* x,
                                 int64_t n_x, int max_iter, int krylov_dim, double tol, atk_gmres_ortho ortho,
                                 atk_gmres_callback cb, void* user, atk_gmres_info* info);

#ifdef __cplusplus
}
#endif
#endif /* ATK_CUDA_H */
