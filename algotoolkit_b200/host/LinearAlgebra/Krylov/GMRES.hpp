// GMRES -- drop-in for the reference's src/LinearAlgebra/Krylov/GMRES.hpp.
//
// solve() forwards to libatk_cuda (atk_gmres_solve_host), which reproduces the reference loop including its
// non-textbook behaviour (projection over i<j only, H never cleared between restarts, absolute tolerance, no inner
// convergence test).  The progress messages the reference writes to std::cout - which its notebooks parse - are
// re-emitted from the library's callback with the same text and the stream's current formatting.
// enablePreconditioner() (SparseMatrixCSC instantiation): the ILU factors are computed on the device once per solve()
// (:39-41) and applied inside the device loop (atk_gmres_solve_precond_host).
#pragma once

#include <iostream>
#include <stdexcept>
#include <type_traits>

#include "../../Obj/DenseObj.hpp"
#include "../../Obj/SparseObj.hpp"
#include "../../Obj/VectorObj.hpp"
#include "../Preconditioner/ILU.hpp"
#include "ConjugateGradient.hpp"  // atk_host::to_device

template <typename TNum, typename MatrixType = SparseMatrixCSC<TNum>, typename VectorType = VectorObj<TNum>>
class GMRES {
    static_assert(std::is_same<TNum, double>::value, "GMRES runs on the GPU in double precision only");
    ILUPreconditioner<double, SparseMatrixCSC<double>> preconditioner;
    bool usePreconditioner = false;

    template <typename M>
    atk_preconditioner_t factorize(const M& A, std::true_type) {
        preconditioner.compute(A);  // :40 - once per solve()
        return preconditioner.device_handle();
    }
    template <typename M>
    atk_preconditioner_t factorize(const M&, std::false_type) {
        throw std::logic_error("GMRES: the preconditioned branch of the CUDA build takes SparseMatrixCSC<double> only");
    }

    static void on_event(int event, double value, void*) {
        switch (event) {  // reference GMRES.hpp:48,50,83,105,108,114 - text kept verbatim
            case ATK_GMRES_INITIAL_RESIDUAL: std::cout << "Initial residual norm: " << value << std::endl; break;
            case ATK_GMRES_INITIAL_GOOD: std::cout << "Initial guess is good enough." << std::endl; break;
            case ATK_GMRES_KRYLOV_UPDATE: std::cout << "KrylovUpdate : " << (int)value << std::endl; break;
            case ATK_GMRES_RESTART_RESIDUAL: std::cout << "Residual norm after restart: " << value << " " << std::endl; break;
            case ATK_GMRES_CONVERGED: std::cout << "Converged after restart at iteration " << (int)value << std::endl; break;
            case ATK_GMRES_MAX_ITER: std::cout << "Reached maximum iterations without convergence." << std::endl; break;
            default: break;
        }
    }

public:
    atk_gmres_info last{};  // extension: counters of the most recent solve()

    GMRES() = default;
    virtual ~GMRES() = default;

    void enablePreconditioner() { usePreconditioner = true; }  // :23-25

    // reference :27-115; x is the initial guess and the result
    void solve(const MatrixType& A, const VectorType& b, VectorType& x, int maxIter, int KrylovDim, double tol) {
        const int n = (int)b.size();
        if (A.getRows() != n || A.getCols() != n) throw std::invalid_argument("Matrix dimensions must match vector size");
        if ((int)x.size() != n) throw std::invalid_argument("Initial guess vector must match system size");
        if (KrylovDim <= 0 || KrylovDim > n) throw std::invalid_argument("Invalid Krylov subspace dimension");
        atk_preconditioner_t pc = nullptr;
        if (usePreconditioner) pc = factorize(A, std::is_same<MatrixType, SparseMatrixCSC<double>>{});
        atk_host::OperatorHandle op = atk_host::to_device(A);
        last = atk_gmres_info{};
        atk_host::check(atk_gmres_solve_precond_host(atk_host::context(), op.get(), pc, b.element(), x.element(), (int64_t)x.size(),
                                                     maxIter, KrylovDim, tol, ATK_GMRES_MGS, &GMRES::on_event, nullptr, &last));
    }
};
