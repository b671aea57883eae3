// ConjugateGrad -- drop-in for the reference's src/LinearAlgebra/Krylov/ConjugateGradient.hpp.
//
// Same public members (x, r, p, _tol, _maxIter, A, b) and the same two entry points.  solve() hands the whole loop to
// libatk_cuda (atk_cg_solve_host): the operator is converted and kept on the GPU, x/r/p/Ap never leave it during the
// iteration, and the |p.Ap| < 1e-12 exit is detected on the device.  As in the reference there is NO tolerance test
// (it is commented out there, :44-47,75-77; _tol is stored and unused) and the state x, r, p persists, so calling
// solve() again continues the iteration.
#pragma once

#include <type_traits>
#include <utility>

#include "../../Obj/DenseObj.hpp"
#include "../../Obj/SparseObj.hpp"
#include "../../Obj/VectorObj.hpp"

namespace atk_host {
// what the device solve reported for the last call (the reference's solve() returns nothing)
struct CgReport {
    int iterations = 0;
    bool stopped_on_pAp = false;
    bool zero_rhs = false;
    double last_pAp = 0.0;
    double rs = 0.0;
    float device_ms = 0.f;
};
inline OperatorHandle to_device(const SparseMatrixCSC<double>& A) { return A.device_operator(); }
inline OperatorHandle to_device(const DenseObj<double>& A) { return SparseMatrixCSC<double>(A).device_operator(); }
}  // namespace atk_host

template <typename TNum, typename MatrixType, typename VectorType>
class ConjugateGrad {
    static_assert(std::is_same<TNum, double>::value, "ConjugateGrad runs on the GPU in double precision only");

public:
    VectorType x, r, p;
    double _tol = 1e-6;  // kept for source compatibility; the reference never reads it either
    int _maxIter = 0;
    MatrixType A;
    VectorType b;
    atk_host::CgReport last;  // extension: iteration count etc. of the most recent solve()

    ConjugateGrad() = default;

    // reference :23-31 - copies A and b, x = 0, r = b - A*x, p = r
    explicit ConjugateGrad(const MatrixType& A_in, const VectorType& b_in, int maxIter, double tol)
        : _tol(tol), _maxIter(maxIter), A(A_in), b(b_in) {
        x = VectorType((int)b.size(), 0.0);
        r = b - (A * x);
        p = r;
    }
    virtual ~ConjugateGrad() = default;

    // reference :36-84
    void solve(VectorType& x_out) {
        atk_host::OperatorHandle op = atk_host::to_device(A);
        atk_cg_info info{};
        atk_host::check(atk_cg_solve_host(atk_host::context(), op.get(), b.element(), x.element(), r.element(), p.element(),
                                          /*fresh=*/0, _maxIter, &info));
        last.iterations = info.iterations;
        last.stopped_on_pAp = info.stopped_on_pAp != 0;
        last.zero_rhs = info.zero_rhs != 0;
        last.last_pAp = info.last_pAp;
        last.rs = info.rs;
        last.device_ms = info.device_ms;
        x_out = x;
    }
};
