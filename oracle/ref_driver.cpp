// oracle/ref_driver.cpp -- C-ABI shim around the UNMODIFIED reference headers.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under algotoolkit_b200/ or include/ may
// call into this file; it exists so that tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference leg can run the reference's own
// CPU implementation of the sparse Krylov hot path as a checker / baseline.
//
// It is compiled against the reference sources *where they lie*
// (-I/root/reference/src, see oracle/Makefile) into oracle/_ref/libatk_ref.so;
// no reference source is copied into this repository.  The only additions are
// (a) oracle/stubs/cblas.h, because DenseObj.hpp includes "cblas.h" and
// OpenBLAS is not installed, and (b) the extern "C" wrappers below.
//
// Reference entry points exercised (paths relative to /root/reference):
//   src/Obj/VectorObj.hpp:52-108        dot / L2norm / scale / div / add / sub
//   src/Obj/SparseObj.hpp:55-152        addValue / finalize / operator() / SpMV
//   src/LinearAlgebra/Krylov/ConjugateGradient.hpp:23-84
//   src/LinearAlgebra/Krylov/GMRES.hpp:27-171
//   src/PDEs/FDM/Poisson.hpp:30-170
//   src/utils.hpp:30-71                 readMatrixMarket
//   src/LinearAlgebra/Solver/IterSolver.hpp:23-44      GradientDescent   ("next" row)
//   src/LinearAlgebra/Krylov/KrylovSubspace.hpp:11-38  Krylov::Arnoldi   ("next" row)
//   src/LinearAlgebra/Preconditioner/ILU.hpp:21-115    ILUPreconditioner + GMRES::enablePreconditioner ("next" row)
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <numeric>
#include <random>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "Obj/VectorObj.hpp"
#include "Obj/DenseObj.hpp"
#include "Obj/SparseObj.hpp"
#include "LinearAlgebra/Krylov/ConjugateGradient.hpp"
#include "LinearAlgebra/Krylov/GMRES.hpp"
#include "LinearAlgebra/Krylov/KrylovSubspace.hpp"
#include "LinearAlgebra/Solver/IterSolver.hpp"
// Poisson3DSolver keeps A and b private; the checker needs to read them.
#define private public
#include "PDEs/FDM/Poisson.hpp"
#undef private
#include "utils.hpp"

namespace {

using Vec = VectorObj<double>;
using Csc = SparseMatrixCSC<double>;
using Dense = DenseObj<double>;

// splitmix64 finaliser -> [0,1); SURVEY.md section 8(d) synthetic generator.
inline double u01(uint64_t s) {
    s += 0x9E3779B97F4A7C15ull;
    s = (s ^ (s >> 30)) * 0xBF58476D1CE4E5B9ull;
    s = (s ^ (s >> 27)) * 0x94D049BB133111EBull;
    s ^= s >> 31;
    return (double)(s >> 11) * (1.0 / 9007199254740992.0);
}

Vec make_vec(const double* p, int n) {
    Vec v(n, 0.0);
    if (n > 0) std::memcpy(v.element(), p, sizeof(double) * (size_t)n);
    return v;
}
void copy_out(const Vec& v, double* out) {
    if (v.size()) std::memcpy(out, v.element(), sizeof(double) * v.size());
}

// Redirect std::cout for the lifetime of the object (GMRES / readMatrixMarket print).
struct CoutCapture {
    std::ostringstream os;
    std::streambuf* old;
    explicit CoutCapture(int precision) : old(std::cout.rdbuf(os.rdbuf())) {
        if (precision > 0) os << std::setprecision(precision);
        // the precision flag lives in std::cout's ios_base, not in the buffer
        saved_precision = std::cout.precision();
        if (precision > 0) std::cout.precision(precision);
    }
    ~CoutCapture() {
        std::cout.rdbuf(old);
        std::cout.precision(saved_precision);
    }
    std::streamsize saved_precision;
};

void copy_log(const std::string& s, char* log, int cap) {
    if (!log || cap <= 0) return;
    size_t n = std::min((size_t)cap - 1, s.size());
    std::memcpy(log, s.data(), n);
    log[n] = 0;
}

std::function<double(double, double, double)> rhs_f(int kind, uint64_t seed, int nx, int ny, double hx, double hy,
                                                     double hz) {
    if (kind == 0) {  // README: f = -3 pi^2 sin(pi x) sin(pi y) sin(pi z)
        return [](double x, double y, double z) {
            return -3.0 * M_PI * M_PI * std::sin(M_PI * x) * std::sin(M_PI * y) * std::sin(M_PI * z);
        };
    }
    // RHS-R: f = 2*u01(seed + idx) - 1 with idx recovered from the coordinates
    return [=](double x, double y, double z) {
        long long i = std::llround(x / hx), j = std::llround(y / hy), k = std::llround(z / hz);
        uint64_t idx = (uint64_t)(i + j * (long long)nx + k * (long long)nx * ny);
        return 2.0 * u01(seed + idx) - 1.0;
    };
}
std::function<double(double, double, double)> bc_g(int kind) {
    if (kind == 0) {
        return [](double x, double y, double z) { return std::sin(M_PI * x) * std::sin(M_PI * y) * std::sin(M_PI * z); };
    }
    return [](double, double, double) { return 0.0; };
}

}  // namespace

extern "C" {

int ref_abi_version(void) { return 1; }

double ref_u01(uint64_t s) { return u01(s); }

// ---------------------------------------------------------------- L0 vectors
double ref_dot(int n, const double* a, const double* b) { return make_vec(a, n) * make_vec(b, n); }
double ref_nrm2(int n, const double* a) { return make_vec(a, n).L2norm(); }
void ref_vec_scale(int n, const double* a, double s, double* out) { copy_out(make_vec(a, n) * s, out); }
int ref_vec_div(int n, const double* a, double s, double* out) {
    try {
        copy_out(make_vec(a, n) / s, out);
    } catch (const std::runtime_error&) {
        return 2;
    }
    return 0;
}
void ref_vec_add(int n, const double* a, const double* b, double* out) { copy_out(make_vec(a, n) + make_vec(b, n), out); }
void ref_vec_sub(int n, const double* a, const double* b, double* out) { copy_out(make_vec(a, n) - make_vec(b, n), out); }

// ---------------------------------------------------------------- matrices
void* ref_mat_from_triplets(int n_rows, int n_cols, long nt, const int* rows, const int* cols, const double* vals) {
    try {
        Csc* A = new Csc(n_rows, n_cols);
        for (long t = 0; t < nt; ++t) A->addValue(rows[t], cols[t], vals[t]);
        A->finalize();
        return A;
    } catch (...) {
        return nullptr;
    }
}
// Adopt finished CSC arrays as-is (also fills the retained construction buffer
// the way finalize() leaves it, SparseObj.hpp:99).
void* ref_mat_from_csc(int n_rows, int n_cols, long nnz, const int* col_ptr, const int* row_idx, const double* vals) {
    Csc* A = new Csc(n_rows, n_cols);
    A->col_ptr.assign(col_ptr, col_ptr + n_cols + 1);
    A->row_indices.assign(row_idx, row_idx + nnz);
    A->values.assign(vals, vals + nnz);
    A->construction_buffer.resize(nnz);
    for (int c = 0; c < n_cols; ++c)
        for (int k = col_ptr[c]; k < col_ptr[c + 1]; ++k) A->construction_buffer[k] = {row_idx[k], c, vals[k]};
    return A;
}
void* ref_mat_read_mm(const char* path, char* log, int log_cap) {
    CoutCapture cap(0);
    Csc* A = new Csc(1, 1);
    utils::readMatrixMarket<double, Csc>(std::string(path), *A);
    copy_log(cap.os.str(), log, log_cap);
    return A;
}
void* ref_mat_poisson(int nx, int ny, int nz, double lx, double ly, double lz) {
    Poisson3DSolver<double> s(nx, ny, nz, lx, ly, lz);
    return new Csc(s.A);
}
void ref_mat_free(void* h) { delete static_cast<Csc*>(h); }
void ref_mat_dims(void* h, int* n_rows, int* n_cols, long* nnz) {
    Csc* A = static_cast<Csc*>(h);
    *n_rows = A->getRows();
    *n_cols = A->getCols();
    *nnz = (long)A->values.size();
}
void ref_mat_export(void* h, int* col_ptr, int* row_idx, double* vals) {
    Csc* A = static_cast<Csc*>(h);
    std::copy(A->col_ptr.begin(), A->col_ptr.end(), col_ptr);
    std::copy(A->row_indices.begin(), A->row_indices.end(), row_idx);
    std::copy(A->values.begin(), A->values.end(), vals);
}
int ref_mat_get(void* h, int i, int j, double* out) {
    try {
        *out = (*static_cast<Csc*>(h))(i, j);
    } catch (const std::out_of_range&) {
        return 3;
    }
    return 0;
}
int ref_mat_spmv(void* h, int n_x, const double* x, double* y) {
    try {
        copy_out((*static_cast<Csc*>(h)) * make_vec(x, n_x), y);
    } catch (const std::invalid_argument&) {
        return 1;
    }
    return 0;
}

// ---------------------------------------------------------------- CG
// The unmodified class: ctor + solve(); state x,r,p are public members.
int ref_cg_solve(void* h, const double* b, int max_iter, double tol, double* x_out, double* r_out, double* p_out) {
    Csc* A = static_cast<Csc*>(h);
    Vec bb = make_vec(b, A->getRows());
    ConjugateGrad<double, Csc, Vec> cg(*A, bb, max_iter, tol);
    Vec x(A->getRows(), 0.0);
    cg.solve(x);
    copy_out(x, x_out);
    if (r_out) copy_out(cg.r, r_out);
    if (p_out) copy_out(cg.p, p_out);
    return 0;
}
// Same loop written with the reference's own operators, instrumented: the
// reference solve() does not report how many iterations ran.
// tests check this is bit-identical to ref_cg_solve.
int ref_cg_trace(void* h, const double* b, int max_iter, double* x_out, double* r_out, double* p_out, int* iters,
                 double* last_pAp, double* pAp_trace, double* rs_trace, int trace_cap) {
    Csc& A = *static_cast<Csc*>(h);
    Vec bb = make_vec(b, A.getRows());
    Vec x(bb.size(), 0.0);
    Vec r = bb - (A * x);
    Vec p = r;
    double rs_old = r * r;
    *iters = 0;
    *last_pAp = 0.0;
    if (bb.L2norm() >= 1e-12) {
        int i = 0;
        for (; i < max_iter; ++i) {
            Vec Ap = A * p;
            double pAp = p * Ap;
            *last_pAp = pAp;
            if (i < trace_cap && pAp_trace) pAp_trace[i] = pAp;
            if (std::abs(pAp) < 1e-12) break;
            double alpha = rs_old / pAp;
            x = x + p * alpha;
            r = r - Ap * alpha;
            double rs_new = r * r;
            if (i < trace_cap && rs_trace) rs_trace[i] = rs_new;
            p = r + p * (rs_new / rs_old);
            rs_old = rs_new;
        }
        *iters = i;
    }
    copy_out(x, x_out);
    if (r_out) copy_out(r, r_out);
    if (p_out) copy_out(p, p_out);
    return 0;
}

// ---------------------------------------------------------------- GMRES
// dense_h = 0: GMRES<double, SparseMatrixCSC> exactly as the pybind module
//              instantiates it (pybind.cpp:88-94); H is a sparse matrix.
// dense_h = 1: GMRES<double, DenseObj>: A and H dense; bit-identical results,
//              ~80x faster.  Needs n*n doubles.
// return: 0 ok, 1 invalid_argument, 2 runtime_error, 3 out_of_range
int ref_gmres_solve(void* h, const double* b, double* x_inout, int n_x, int max_iter, int krylov_dim, double tol,
                    int dense_h, int log_precision, char* log, int log_cap) {
    Csc& A = *static_cast<Csc*>(h);
    const int n = A.getRows();
    Vec bb = make_vec(b, n);
    Vec x = make_vec(x_inout, n_x);
    int rc = 0;
    CoutCapture cap(log_precision);
    try {
        if (dense_h) {
            Dense Ad(A.getRows(), A.getCols());
            for (int c = 0; c < A.getCols(); ++c)
                for (int k = A.col_ptr[c]; k < A.col_ptr[c + 1]; ++k) Ad(A.row_indices[k], c) = A.values[k];
            GMRES<double, Dense, Vec> g;
            g.solve(Ad, bb, x, max_iter, krylov_dim, tol);
        } else {
            GMRES<double, Csc, Vec> g;
            g.solve(A, bb, x, max_iter, krylov_dim, tol);
        }
    } catch (const std::invalid_argument&) {
        rc = 1;
    } catch (const std::out_of_range&) {
        rc = 3;
    } catch (const std::runtime_error&) {
        rc = 2;
    }
    copy_log(cap.os.str(), log, log_cap);
    if (rc == 0 || rc == 2) copy_out(x, x_inout);
    return rc;
}

// ---------------------------------------------------------------- GradientDescent / Krylov::Arnoldi ("next" rows)
// ---------------------------------------------------------------- ILU preconditioner ("next" row)
// compute() on the sparse instantiation; rc 0 ok, 1 invalid_argument (non-square), 2 runtime_error (singular).
void* ref_ilu_compute(void* h, int* rc) {
    auto* ilu = new ILUPreconditioner<double, Csc>();
    *rc = 0;
    try {
        ilu->compute(*static_cast<Csc*>(h));
    } catch (const std::invalid_argument&) {
        *rc = 1;
    } catch (const std::out_of_range&) {
        *rc = 3;
    } catch (const std::runtime_error&) {
        *rc = 2;
    }
    if (*rc) {
        delete ilu;
        return nullptr;
    }
    return ilu;
}
void ref_ilu_free(void* p) { delete static_cast<ILUPreconditioner<double, Csc>*>(p); }
// which = 0: copy of getLFactor(), 1: copy of getUFactor(); free with ref_mat_free
void* ref_ilu_factor(void* p, int which) {
    auto* ilu = static_cast<ILUPreconditioner<double, Csc>*>(p);
    return new Csc(which == 0 ? ilu->getLFactor() : ilu->getUFactor());
}
int ref_ilu_solve(void* p, int n, const double* b, double* x) {
    auto* ilu = static_cast<ILUPreconditioner<double, Csc>*>(p);
    try {
        copy_out(ilu->solve(make_vec(b, n)), x);
    } catch (const std::invalid_argument&) {
        return 1;
    } catch (const std::runtime_error&) {
        return 2;
    }
    return 0;
}
// GMRES<double, SparseMatrixCSC, VectorObj> with enablePreconditioner() (GMRES.hpp:23-25,39-45,67-68,100-102)
int ref_gmres_solve_precond(void* h, const double* b, double* x_inout, int n_x, int max_iter, int krylov_dim, double tol,
                            int log_precision, char* log, int log_cap) {
    Csc& A = *static_cast<Csc*>(h);
    const int n = A.getRows();
    Vec bb = make_vec(b, n);
    Vec x = make_vec(x_inout, n_x);
    int rc = 0;
    CoutCapture cap(log_precision);
    try {
        GMRES<double, Csc, Vec> g;
        g.enablePreconditioner();
        g.solve(A, bb, x, max_iter, krylov_dim, tol);
    } catch (const std::invalid_argument&) {
        rc = 1;
    } catch (const std::out_of_range&) {
        rc = 3;
    } catch (const std::runtime_error&) {
        rc = 2;
    }
    copy_log(cap.os.str(), log, log_cap);
    if (rc == 0 || rc == 2) copy_out(x, x_inout);
    return rc;
}

int ref_gd_solve(void* h, const double* b, double* x_inout, int max_iter, double tol) {
    Csc& A = *static_cast<Csc*>(h);
    Vec bb = make_vec(b, A.getRows());
    Vec x = make_vec(x_inout, A.getRows());
    GradientDescent<double, Csc, Vec> gd(A, bb, max_iter, tol);
    gd.solve(x);
    copy_out(x, x_inout);
    return 0;
}
// Q: m columns of n (column-major), Q[0] given; H: m x (m-1) column-major (DenseObj storage order).
// The reference instantiation is the one its pybind module binds: Arnoldi<double, DenseObj, VectorObj> (pybind.cpp:47).
int ref_arnoldi(void* h, double* Q, int m, double* H, double tol, char* log, int log_cap) {
    Csc& A = *static_cast<Csc*>(h);
    const int n = A.getRows();
    Dense Ad(n, n);
    for (int c = 0; c < A.getCols(); ++c)
        for (int k = A.col_ptr[c]; k < A.col_ptr[c + 1]; ++k) Ad(A.row_indices[k], c) = A.values[k];
    std::vector<Vec> Qv;
    for (int i = 0; i < m; ++i) Qv.push_back(make_vec(Q + (size_t)i * n, n));
    Dense Hd(m, m - 1);
    for (int i = 0; i < m * (m - 1); ++i) Hd[i] = H[i];
    CoutCapture cap(0);
    Krylov::Arnoldi<double, Dense, Vec>(Ad, Qv, Hd, tol);
    copy_log(cap.os.str(), log, log_cap);
    for (int i = 0; i < m; ++i) copy_out(Qv[i], Q + (size_t)i * n);
    for (int i = 0; i < m * (m - 1); ++i) H[i] = Hd[i];
    return 0;
}

// ---------------------------------------------------------------- Poisson3DSolver
// rhs_kind 0 = README sin RHS + matching Dirichlet data; 1 = RHS-R (seeded), g = 0.
// Writes the RHS the solver ended up with (b) and the solution (u).
int ref_poisson_solve(int nx, int ny, int nz, double lx, double ly, double lz, int rhs_kind, uint64_t seed, int max_iter,
                      double tol, double* b_out, double* u_out, double* seconds_build, double* seconds_solve) {
    using clk = std::chrono::steady_clock;
    auto t0 = clk::now();
    Poisson3DSolver<double> s(nx, ny, nz, lx, ly, lz);
    auto t1 = clk::now();
    s.setRHS(rhs_f(rhs_kind, seed, nx, ny, s.getDx(), s.getDy(), s.getDz()));
    s.setDirichletBC(bc_g(rhs_kind));
    if (b_out) copy_out(s.b, b_out);
    auto t2 = clk::now();
    s.solve(max_iter, tol);
    auto t3 = clk::now();
    if (u_out) copy_out(s.getSolution(), u_out);
    if (seconds_build) *seconds_build = std::chrono::duration<double>(t1 - t0).count();
    if (seconds_solve) *seconds_solve = std::chrono::duration<double>(t3 - t2).count();
    return 0;
}
// b only (no solve), for building device inputs.
int ref_poisson_rhs(int nx, int ny, int nz, double lx, double ly, double lz, int rhs_kind, uint64_t seed, double* b_out) {
    const double hx = lx / (nx - 1), hy = ly / (ny - 1), hz = lz / (nz - 1);
    auto f = rhs_f(rhs_kind, seed, nx, ny, hx, hy, hz);
    auto g = bc_g(rhs_kind);
    for (int k = 0; k < nz; k++)
        for (int j = 0; j < ny; j++)
            for (int i = 0; i < nx; i++) {
                const double x = i * hx, y = j * hy, z = k * hz;
                const bool shell = (i == 0 || i == nx - 1 || j == 0 || j == ny - 1 || k == 0 || k == nz - 1);
                b_out[i + (size_t)j * nx + (size_t)k * nx * ny] = shell ? g(x, y, z) : f(x, y, z);
            }
    return 0;
}
int ref_poisson_get_solution_oob(void) {
    Poisson3DSolver<double> s(3, 3, 3, 1, 1, 1);
    try {
        (void)s.getSolution(3, 0, 0);
    } catch (const std::out_of_range&) {
        return 3;
    }
    return 0;
}

}  // extern "C"
