/* oracle/krylov_oracle.h -- CPU restatement of the reference's sparse Krylov hot path.
 *
 * TEST INFRASTRUCTURE ONLY: a plain-C, single-threaded re-statement of what the
 * reference computes, used as the checker in tests/, __graft_entry__.smoke() and as
 * bench.py's `cpu_baseline` ("port") when oracle/_ref is unavailable.  The product
 * (algotoolkit_b200/, include/) never links, imports or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function below against
 * (a) the reference's own known-answer tests and notebook outputs (the JSON fixtures under tests/golden)
 * and (b) the unmodified reference compiled into oracle/_ref/libatk_ref.so, bit for bit.
 *
 * Every function cites the reference file:line it follows (paths relative to the
 * reference root).  All arithmetic is IEEE double, multiply-then-add (the reference is
 * built without FMA: CMakeLists.txt:7 has no -march), compiled with -ffp-contract=off.
 */
#ifndef ATK_KRYLOV_ORACLE_H
#define ATK_KRYLOV_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* splitmix64 finaliser -> [0,1): the deterministic generator of SURVEY.md 8(d). */
double orc_u01(uint64_t s);

/* ---- VectorObj (src/Obj/VectorObj.hpp) ---- */
double orc_dot(long n, const double* a, const double* b);           /* :105-108 sequential from 0 */
double orc_nrm2(long n, const double* a);                           /* :52-54 */
void orc_vec_scale(long n, const double* a, double s, double* out); /* :64-68  a*s */
int orc_vec_div(long n, const double* a, double s, double* out);    /* :75-80  a/s, returns 2 on s==0 */
void orc_vec_add(long n, const double* a, const double* b, double* out); /* :89-94 */
void orc_vec_sub(long n, const double* a, const double* b, double* out); /* :97-102 */

/* ---- SparseMatrixCSC (src/Obj/SparseObj.hpp) ---- */
/* addValue+finalize (:55-100): drop explicit zeros, sort by (col,row), count, prefix-sum.
 * Triplets must be unique in (row,col) (the reference's order among duplicates is
 * unspecified: std::sort is unstable).  Returns nnz, or -3 on an out-of-range index. */
long orc_csc_from_triplets(int n_rows, int n_cols, long nt, const int* rows, const int* cols, const double* vals,
                           int* col_ptr, int* row_idx, double* out_vals);
/* operator()(i,j) (:121-129): lower_bound in the column; 0 when absent; -3 out of range */
int orc_csc_get(int n_rows, int n_cols, const int* col_ptr, const int* row_idx, const double* vals, int i, int j,
                double* out);
/* operator*(VectorObj) (:131-152): y=0; for col ascending, skip x[col]==0, y[row]+=val*x[col] */
void orc_spmv_csc(int n_rows, int n_cols, const int* col_ptr, const int* row_idx, const double* vals, const double* x,
                  double* y);

/* ---- utils::readMatrixMarket (src/utils.hpp:30-71) ----
 * Skips leading '%' lines, reads "M N L", then L triplets (1-based), IGNORING the
 * symmetric/pattern qualifiers.  Two-call protocol: first with col_ptr==NULL to get
 * sizes, then with buffers.  Returns nnz (>=0), -1 if the file cannot be opened. */
long orc_read_matrix_market(const char* path, int* n_rows, int* n_cols, long* n_entries, int* col_ptr, int* row_idx,
                            double* vals);

/* ---- Poisson3DSolver (src/PDEs/FDM/Poisson.hpp) ---- */
/* coefficients (:30-57): h=l/(n-1), c=1/h^2, cc=-2(cx+cy+cz) */
void orc_poisson_coeffs(int nx, int ny, int nz, double lx, double ly, double lz, double* cx, double* cy, double* cz,
                        double* cc);
/* nnz of buildLaplacianMatrix (:49-95) */
long orc_poisson_nnz(int nx, int ny, int nz);
/* assembled CSC exactly as buildLaplacianMatrix + finalize leave it */
void orc_poisson_csc(int nx, int ny, int nz, double lx, double ly, double lz, int* col_ptr, int* row_idx, double* vals);
/* matrix-free application of the same operator: identity on the Dirichlet shell, taps
 * accumulated in ascending column order k-1,j-1,i-1,c,i+1,j+1,k+1 from 0.0 */
void orc_poisson_apply(int nx, int ny, int nz, double cx, double cy, double cz, const double* x, double* y);
/* b after setRHS(f) then setDirichletBC(g) (:103-151).
 * rhs_kind 0: README f=-3pi^2 sin sin sin, g=sin sin sin;  1: RHS-R f=2*u01(seed+idx)-1, g=0 */
void orc_poisson_rhs(int nx, int ny, int nz, double lx, double ly, double lz, int rhs_kind, uint64_t seed, double* b);

/* ---- synthetic 27-point operator of BASELINE config 5 ----
 * open boundaries; A(r,r)=40; A(r,c) = -1 + 0.3*(di + 0.5*dj + 0.25*dk) */
long orc_op27_nnz(int nx, int ny, int nz);
void orc_op27_csc(int nx, int ny, int nz, int* col_ptr, int* row_idx, double* vals);

/* ---- ConjugateGrad (src/LinearAlgebra/Krylov/ConjugateGradient.hpp:23-84) ----
 * Operator given either as CSC (col_ptr != NULL) or as the matrix-free Poisson stencil
 * (col_ptr == NULL, grid/coeffs in `st`).  State x,r,p is in/out: pass x=0,r=b,p=b for
 * a fresh solver (what the ctor leaves, :26-30); a second call continues.
 * iters = loop index at exit (max_iter if the loop ran out); traces may be NULL. */
typedef struct {
    int nx, ny, nz;
    double cx, cy, cz;
} orc_stencil7;
int orc_cg_solve(int n, const int* col_ptr, const int* row_idx, const double* vals, const orc_stencil7* st,
                 const double* b, int max_iter, double* x, double* r, double* p, int* iters, double* last_pAp,
                 double* pAp_trace, double* rs_trace, int trace_cap);

/* ---- GMRES (src/LinearAlgebra/Krylov/GMRES.hpp:27-171), unpreconditioned ----
 * Restated WITH the reference's behaviour: projection loop i<j (:73), H allocated once
 * per solve and never cleared (:55), no inner convergence test, absolute tol, Givens
 * from the (possibly stale) H(j,j) (:89-94), x = x + V[i]*y[i] one sweep per i (:167-170).
 * resid[0] = initial ||r||, resid[1..] = ||r|| after each restart (n_resid entries).
 * status: 0 reached max_iter, 1 converged after a restart, 2 initial guess good enough.
 * returns 0 ok, 1 invalid argument, 2 runtime_error (zero H diagonal / division by zero) */
int orc_gmres_solve(int n, const int* col_ptr, const int* row_idx, const double* vals, const double* b, double* x,
                    int n_x, int max_iter, int krylov_dim, double tol, double* resid, int resid_cap, int* n_resid,
                    int* status, int* last_kry_update);

/* ---- "next" row: the enablePreconditioner() branch ----
 * ILUPreconditioner::compute / solve / getLFactor / getUFactor (src/LinearAlgebra/Preconditioner/ILU.hpp:21-115) on a
 * dense column-major n x n array, and GMRES::solve with usePreconditioner (GMRES.hpp:39-45,67-68,100-102).
 * orc_ilu_compute returns 0 ok, 2 "Matrix is numerically singular". */
int orc_ilu_compute(int n, const int* col_ptr, const int* row_idx, const double* vals, double* LU);
void orc_ilu_solve(int n, const double* LU, const double* b, double* x);
void orc_ilu_factors(int n, const double* LU, double* L, double* U);
int orc_gmres_solve_precond(int n, const int* col_ptr, const int* row_idx, const double* vals, const double* b, double* x,
                            int n_x, int max_iter, int krylov_dim, double tol, double* resid, int resid_cap, int* n_resid,
                            int* status, int* last_kry_update);

/* ILU(0): the reference's elimination (loops and drop rules of ILU.hpp:21-81) restricted to the positions A stores - no
 * fill-in.  The reference itself has no such factorisation; where its full elimination creates no fill-in the two agree
 * bit for bit.  LU is a dense column-major n x n array like orc_ilu_compute's; orc_ilu_solve / orc_ilu_factors apply. */
int orc_ilu0_compute(int n, const int* col_ptr, const int* row_idx, const double* vals, double* LU);
int orc_gmres_solve_precond_ilu0(int n, const int* col_ptr, const int* row_idx, const double* vals, const double* b, double* x,
                                 int n_x, int max_iter, int krylov_dim, double tol, double* resid, int resid_cap, int* n_resid,
                                 int* status, int* last_kry_update);

/* ---- "next" rows: compositions of the same kernels ---- */
/* GradientDescent::solve (src/LinearAlgebra/Solver/IterSolver.hpp:23-44): steepest descent with the residual-norm loop
 * condition and the change-norm early exit.  x in/out; iters = iterations whose update was applied. */
int orc_gd_solve(int n, const int* col_ptr, const int* row_idx, const double* vals, const double* b, double* x, int max_iter,
                 double tol, int* iters, double* residual_norm, double* change_norm);
/* Krylov::Arnoldi (src/LinearAlgebra/Krylov/KrylovSubspace.hpp:11-38): textbook MGS Arnoldi (projects against ALL previous
 * vectors, unlike the loop inside GMRES).  Q: m columns of n (column-major, Q[0] given); H: m x (m-1) column-major, only
 * the entries the reference writes are touched.  breakdown = iteration index at which ||Av|| < tol stopped it, or -1. */
int orc_arnoldi(int n, const int* col_ptr, const int* row_idx, const double* vals, double* Q, int m, double* H, double tol,
                int* breakdown);

#ifdef __cplusplus
}
#endif
#endif
