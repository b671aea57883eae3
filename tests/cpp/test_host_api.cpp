// The reference's own hot-path unit tests (tests/SparseMatrixCSCTest.cpp, ConjugateGradient_test.cpp,
// GMRES_test.cpp, debuglogger.cpp of AlgoToolkit), restated as plain checks against the CUDA-backed drop-in headers.
// Same class names, same calls, same expectations; gtest is not installed here, hence the small CHECK macros.
// Build: see tests/test_host_cpp.py.  Needs a GPU to run (there is no CPU fallback).
#include <cmath>
#include <cstdio>
#include <sstream>
#include <string>

#include "LinearAlgebra/Krylov/ConjugateGradient.hpp"
#include "LinearAlgebra/Krylov/GMRES.hpp"
#include "LinearAlgebra/Krylov/KrylovSubspace.hpp"
#include "LinearAlgebra/Solver/IterSolver.hpp"
#include "Obj/DenseObj.hpp"
#include "Obj/SparseObj.hpp"
#include "Obj/VectorObj.hpp"
#include "PDEs/FDM/Poisson.hpp"
#include "utils.hpp"

static int g_fail = 0, g_checks = 0;
#define CHECK(cond)                                                          \
    do {                                                                     \
        ++g_checks;                                                          \
        if (!(cond)) {                                                       \
            ++g_fail;                                                        \
            std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond);      \
        }                                                                    \
    } while (0)
#define CHECK_NEAR(a, b, tol) CHECK(std::fabs((a) - (b)) <= (tol))
#define CHECK_THROWS(expr, Ex)        \
    do {                              \
        bool thrown = false;          \
        try {                         \
            (void)(expr);             \
        } catch (const Ex&) {         \
            thrown = true;            \
        } catch (...) {               \
        }                             \
        CHECK(thrown);                \
    } while (0)

static SparseMatrixCSC<double> tridiag3() {  // ConjugateGradient_test.cpp:14-22
    SparseMatrixCSC<double> m(3, 3);
    m.addValue(0, 0, 4);
    m.addValue(1, 0, -1);
    m.addValue(0, 1, -1);
    m.addValue(1, 1, 4);
    m.addValue(2, 1, -1);
    m.addValue(1, 2, -1);
    m.addValue(2, 2, 4);
    m.finalize();
    return m;
}

static void sparse_matrix_tests() {  // SparseMatrixCSCTest.cpp:32-64,136-139
    SparseMatrixCSC<double> mat(3, 3);
    mat.addValue(0, 2, 3.0);
    mat.addValue(1, 0, 4.0);
    mat.addValue(2, 1, 5.0);
    mat.addValue(2, 2, 6.0);
    mat.finalize();
    CHECK(mat(0, 2) == 3.0);
    CHECK(mat(1, 0) == 4.0);
    CHECK(mat(2, 1) == 5.0);
    CHECK(mat(2, 2) == 6.0);
    CHECK(mat(0, 0) == 0.0);
    VectorObj<double> vec(3);
    vec[0] = 1.0;
    vec[1] = 2.0;
    vec[2] = 3.0;
    VectorObj<double> result = mat * vec;
    CHECK(result[0] == 9.0);
    CHECK(result[1] == 4.0);
    CHECK(result[2] == 28.0);
    CHECK_THROWS(mat(3, 0), std::out_of_range);
    CHECK_THROWS(mat(0, 3), std::out_of_range);
    VectorObj<double> bad(4);
    CHECK_THROWS(mat * bad, std::invalid_argument);
    // overwrite of a stored entry after finalize (SparseObj.hpp:59-66)
    mat.addValue(2, 2, 7.0);
    CHECK(mat(2, 2) == 7.0);
    // int instantiation keeps the container interface (SparseMatrixCSCTest.cpp:17)
    SparseMatrixCSC<int> mi(3, 3);
    mi.addValue(0, 0, 1);
    mi.addValue(1, 1, 2);
    mi.finalize();
    CHECK(mi(1, 1) == 2 && mi(0, 1) == 0);
}

static void vector_tests() {
    VectorObj<double> a(3), b(3);
    for (int i = 0; i < 3; ++i) {
        a[i] = i + 1.0;
        b[i] = 2.0 * i - 1.0;
    }
    CHECK((a * b) == (1.0 * -1.0 + 2.0 * 1.0 + 3.0 * 3.0));
    CHECK(a.L2norm() == std::sqrt(14.0));
    VectorObj<double> c = a + b, d = a - b, e = a * 0.5, f = a / 4.0;
    CHECK(c[2] == 6.0 && d[0] == 2.0 && e[1] == 1.0 && f[2] == 0.75);
    CHECK_THROWS(a / 0.0, std::runtime_error);
    CHECK_THROWS(a[3], std::out_of_range);
    VectorObj<double> g(4);
    CHECK_THROWS(a + g, std::invalid_argument);
    CHECK_THROWS(a * g, std::invalid_argument);
}

static void cg_tests() {  // ConjugateGradient_test.cpp:67-100, debuglogger.cpp
    SparseMatrixCSC<double> matrix = tridiag3();
    VectorObj<double> rhs(3);
    rhs[0] = 1;
    rhs[1] = 5;
    rhs[2] = 0;
    const double expect[3] = {0.625, 1.5, 0.375};
    {
        ConjugateGrad<double, SparseMatrixCSC<double>, VectorObj<double>> cg(matrix, rhs, 1000, 1e-6);
        VectorObj<double> sol(3, 0.0);
        cg.solve(sol);
        for (int i = 0; i < 3; ++i) CHECK_NEAR(sol[i], expect[i], 1e-5);
    }
    {
        VectorObj<double> zero(3, 0.0), sol(3, 0.0);
        ConjugateGrad<double, SparseMatrixCSC<double>, VectorObj<double>> cg(matrix, zero, 1000, 1e-6);
        cg.solve(sol);
        for (int i = 0; i < 3; ++i) CHECK_NEAR(sol[i], 0.0, 1e-12);
    }
    {
        ConjugateGrad<double, SparseMatrixCSC<double>, VectorObj<double>> cg(matrix, rhs, 1, 1e-6);
        VectorObj<double> sol(3, 0.0);
        cg.solve(sol);
        CHECK(std::fabs(rhs[0] - matrix(0, 0) * sol[0]) > 1e-6);
        CHECK(cg.last.iterations == 1);
        // state persists: a second solve() continues (bench_Solver.cpp:74-77)
        cg._maxIter = 50;
        cg.solve(sol);
        for (int i = 0; i < 3; ++i) CHECK_NEAR(sol[i], expect[i], 1e-5);
    }
    {
        SparseMatrixCSC<double> A(4, 4);
        A.addValue(0, 0, 4.0);
        A.addValue(1, 1, 3.0);
        A.addValue(2, 2, 2.0);
        A.addValue(3, 3, 1.0);
        A.finalize();
        VectorObj<double> b(4), x(4, 0.0);
        b[0] = 8.0; b[1] = 9.0; b[2] = 4.0; b[3] = 1.0;
        ConjugateGrad<double, SparseMatrixCSC<double>, VectorObj<double>> cg(A, b, 10, 1e-6);
        cg.solve(x);
        CHECK_NEAR(x[0], 2.0, 1e-6);
        CHECK_NEAR(x[1], 3.0, 1e-6);
        CHECK_NEAR(x[2], 2.0, 1e-6);
        CHECK_NEAR(x[3], 1.0, 1e-6);
    }
}

static void gmres_tests() {  // GMRES_test.cpp:43-92 (DenseObj A, as in the reference test) + the sparse instantiation
    DenseObj<double> matrix(3, 3);
    matrix(0, 0) = 4.0; matrix(0, 1) = -1.0; matrix(0, 2) = 0.0;
    matrix(1, 0) = -1.0; matrix(1, 1) = 4.0; matrix(1, 2) = -1.0;
    matrix(2, 0) = 0.0; matrix(2, 1) = -1.0; matrix(2, 2) = 4.0;
    VectorObj<double> b(3);
    b[0] = 1.0; b[1] = 5.0; b[2] = 0.0;
    std::ostringstream sink;
    std::streambuf* old = std::cout.rdbuf(sink.rdbuf());
    {
        VectorObj<double> x(3, 0.0);
        GMRES<double, DenseObj<double>> solver;
        solver.solve(matrix, b, x, 10000, 3, 1e-12);
        CHECK_NEAR(x[0], 0.625, 1e-5);
        CHECK_NEAR(x[1], 1.5, 1e-5);
        CHECK_NEAR(x[2], 0.375, 1e-5);
        VectorObj<double> res = b - matrix * x;
        CHECK(res.L2norm() < 1e-6);
        VectorObj<double> wrong(4, 0.0);
        CHECK_THROWS(solver.solve(matrix, b, wrong, 10000, 3, 1e-12), std::invalid_argument);
        for (int dim = 1; dim <= 3; ++dim) {
            VectorObj<double> xt(3, 0.0);
            solver.solve(matrix, b, xt, 10000, dim, 1e-12);
            VectorObj<double> rr = b - matrix * xt;
            CHECK(rr.L2norm() < 1e-6);
        }
        VectorObj<double> zb(3, 0.0), zx(3, 0.0);
        solver.solve(matrix, zb, zx, 10000, 3, 1e-12);
        CHECK_NEAR(zx[0], 0.0, 1e-10);
    }
    {
        SparseMatrixCSC<double> A = tridiag3();
        VectorObj<double> x(3, 0.0);
        GMRES<double> solver;  // default MatrixType = SparseMatrixCSC<double>
        solver.solve(A, b, x, 10000, 3, 1e-12);
        CHECK_NEAR(x[1], 1.5, 1e-5);
        CHECK(solver.last.status == 1);
    }
    std::cout.rdbuf(old);
    const std::string log = sink.str();
    CHECK(log.find("Initial residual norm: ") == 0);
    CHECK(log.find("Residual norm after restart: ") != std::string::npos);
    CHECK(log.find("Converged after restart at iteration ") != std::string::npos);
    CHECK(log.find("Initial guess is good enough.") != std::string::npos);
}

static void gd_arnoldi_tests() {  // tests/itersolver_test.cpp:5-39, tests/KrylovSubspace_test.cpp (orthogonality, A Q = Q H)
    SparseMatrixCSC<double> A = tridiag3();
    VectorObj<double> b(3), x(3, 0.0);
    b[0] = 1.0; b[1] = 5.0; b[2] = 0.0;
    GradientDescent<double, SparseMatrixCSC<double>, VectorObj<double>> gd(A, b, 1000, 1e-6);
    gd.solve(x);
    CHECK_NEAR(x[0], 0.625, 1e-5);
    CHECK_NEAR(x[1], 1.5, 1e-5);
    CHECK_NEAR(x[2], 0.375, 1e-5);
    CHECK(gd.last.iterations == 15);  // the reference stops after 15 updates on this system (oracle)

    const int n = 6, m = 4;
    DenseObj<double> D(n, n);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) D(i, j) = (i == j) ? 4.0 + i : 1.0 / (1.0 + i + j);
    std::vector<VectorObj<double>> Q(m, VectorObj<double>(n, 0.0));
    Q[0][0] = 1.0;
    DenseObj<double> H(m, m - 1);
    Krylov::Arnoldi<double, DenseObj<double>, VectorObj<double>>(D, Q, H, 1e-10);
    for (int a = 0; a < m; ++a)
        for (int c = 0; c < m; ++c) CHECK_NEAR(Q[a] * Q[c], a == c ? 1.0 : 0.0, 1e-8);
    for (int c = 0; c + 1 < m; ++c) {  // A q_c = sum_r H(r,c) q_r
        VectorObj<double> lhs = D * Q[c];
        VectorObj<double> rhs(n, 0.0);
        for (int r = 0; r <= c + 1; ++r) rhs = rhs + Q[r] * H(r, c);
        CHECK((lhs - rhs).L2norm() < 1e-8);
    }
}

static void poisson_tests() {  // README / code/PDE.ipynb: 1 CG iteration, errors 8.529918e-04 / 2.886627e-04 at 32^3
    const int n = 32;
    Poisson3DSolver<double> s(n, n, n, 1.0, 1.0, 1.0);
    const double pi = 3.14159265358979323846;
    s.setRHS([pi](double x, double y, double z) { return -3.0 * pi * pi * std::sin(pi * x) * std::sin(pi * y) * std::sin(pi * z); });
    s.setDirichletBC([pi](double x, double y, double z) { return std::sin(pi * x) * std::sin(pi * y) * std::sin(pi * z); });
    s.solve(2000, 1e-6);
    CHECK(s.last.iterations == 1);
    double max_err = 0.0, sq = 0.0;
    for (int k = 0; k < n; ++k)
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < n; ++i) {
                const double an = std::sin(pi * i * s.getDx()) * std::sin(pi * j * s.getDy()) * std::sin(pi * k * s.getDz());
                const double e = std::fabs(s.getSolution(i, j, k) - an);
                max_err = std::fmax(max_err, e);
                sq += e * e;
            }
    CHECK_NEAR(max_err, 8.529918e-04, 5e-10);
    CHECK_NEAR(std::sqrt(sq / (n * n * n)), 2.886627e-04, 5e-10);
    CHECK_THROWS(s.getSolution(n, 0, 0), std::out_of_range);
    // no Dirichlet direction: stencil rows on the boundary with missing neighbours dropped (device-assembled)
    Poisson3DSolver<double> open(6, 6, 6, 1.0, 1.0, 1.0, BoundaryType::Neumann, BoundaryType::Neumann, BoundaryType::Periodic);
    open.setRHS([](double x, double y, double z) { return 1.0 + x + 2.0 * y - z; });
    open.solve(5, 1e-6);
    CHECK(open.last.iterations == 5);
}

static void ilu_tests() {  // ILU_test.cpp:52-153 + GMRES::enablePreconditioner()
    using MatrixType = SparseMatrixCSC<double>;
    using VectorType = VectorObj<double>;
    auto make = [] {
        MatrixType A(3, 3);
        A.addValue(0, 0, 4.0);  A.addValue(0, 1, -1.0); A.addValue(0, 2, 0.0);
        A.addValue(1, 0, -1.0); A.addValue(1, 1, 4.0);  A.addValue(1, 2, -1.0);
        A.addValue(2, 0, 0.0);  A.addValue(2, 1, -1.0); A.addValue(2, 2, 4.0);
        A.finalize();
        return A;
    };
    {
        ILUPreconditioner<double> ilu;
        ilu.compute(make());  // Initialization
        MatrixType nonsquare(3, 4);
        CHECK_THROWS(ilu.compute(nonsquare), std::invalid_argument);
    }
    {
        MatrixType singular(3, 3);
        singular.addValue(0, 1, 1.0);
        singular.finalize();
        ILUPreconditioner<double> ilu;
        CHECK_THROWS(ilu.compute(singular), std::runtime_error);
    }
    {
        ILUPreconditioner<double> ilu;
        MatrixType A = make();
        ilu.compute(A);
        VectorType x_true(3);
        x_true[0] = 1.0; x_true[1] = 2.0; x_true[2] = 3.0;
        VectorType b = A * x_true;
        VectorType x = ilu.solve(b);
        for (int i = 0; i < 3; ++i) CHECK_NEAR(x[i], x_true[i], 1e-10);
        const auto& L = ilu.getLFactor();
        const auto& U = ilu.getUFactor();
        CHECK(L.getRows() == 3 && L.getCols() == 3 && U.getRows() == 3 && U.getCols() == 3);
        for (int i = 0; i < 3; ++i) {
            CHECK_NEAR(L(i, i), 1.0, 1e-10);
            for (int j = i + 1; j < 3; ++j) CHECK_NEAR(L(i, j), 0.0, 1e-10);
            for (int j = 0; j < i; ++j) CHECK_NEAR(U(i, j), 0.0, 1e-10);
        }
        CHECK(L(1, 0) == -0.25 && U(0, 0) == 4.0 && U(1, 1) == 3.75);
        VectorType wrong(4, 1.0);
        CHECK_THROWS(ilu.solve(wrong), std::invalid_argument);
        ILUPreconditioner<double> fresh;
        VectorType b3(3, 1.0);
        CHECK_THROWS(fresh.solve(b3), std::runtime_error);
        // preconditioned GMRES: M^-1 A = I and the reference's i<j projection breaks down at j = 1 in every cycle, so
        // each restart only shrinks the residual by a factor (oracle: 2.3e-6 after 60 restarts)
        GMRES<double, MatrixType, VectorType> gmres;
        gmres.enablePreconditioner();
        VectorType x0(3, 0.0);
        std::ostringstream sink;
        std::streambuf* old = std::cout.rdbuf(sink.rdbuf());
        gmres.solve(A, b, x0, 60, 3, 1e-12);
        std::cout.rdbuf(old);
        for (int i = 0; i < 3; ++i) CHECK_NEAR(x0[i], x_true[i], 1e-5);
        CHECK(sink.str().find("Initial residual norm: ") == 0 && sink.str().find("KrylovUpdate : 1") != std::string::npos);
    }
}

int main(int argc, char** argv) {
    sparse_matrix_tests();
    vector_tests();
    cg_tests();
    gmres_tests();
    gd_arnoldi_tests();
    ilu_tests();
    poisson_tests();
    if (argc > 1) {  // Matrix Market round trip: path to bcsstk01.mtx
        SparseMatrixCSC<double> A(1, 1);
        std::ostringstream sink;
        std::streambuf* old = std::cout.rdbuf(sink.rdbuf());
        utils::readMatrixMarket<double, SparseMatrixCSC<double>>(argv[1], A);
        std::cout.rdbuf(old);
        CHECK(A.getRows() == 48 && A.getCols() == 48 && A.values.size() == 224);  // lower triangle only: header ignored
        CHECK(sink.str() == "M: 48 N: 48 L: 224\nMatrix created 48 48\n");
    }
    std::printf("%d checks, %d failed\n", g_checks, g_fail);
    return g_fail ? 1 : 0;
}
