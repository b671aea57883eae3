// GradientDescent -- drop-in for the steepest-descent solver of the reference's src/LinearAlgebra/Solver/IterSolver.hpp
// (:12-45).  Same constructor (references to A and b are kept, as in the reference) and solve(x); the loop runs on the GPU
// through atk_gd_solve_host with the reference's loop condition (||r|| > tol) and early exit (||x - xOld|| < tol).
// The SOR class of the same reference header is a sequential Gauss-Seidel sweep and is out of scope.
#pragma once

#include <type_traits>

#include "../../Obj/SparseObj.hpp"
#include "../../Obj/VectorObj.hpp"
#include "../Krylov/ConjugateGradient.hpp"  // atk_host::to_device

template <typename TNum, typename SparseMatrixType, typename VectorType>
class GradientDescent {
    static_assert(std::is_same<TNum, double>::value, "GradientDescent runs on the GPU in double precision only");
    const SparseMatrixType& A;
    const VectorType& b;
    int maxIter;
    TNum tol;

public:
    atk_gd_info last{};  // extension: what the device loop reported

    GradientDescent(const SparseMatrixType& matrix, const VectorType& rhs, int iterations, TNum tolerance)
        : A(matrix), b(rhs), maxIter(iterations), tol(tolerance) {}

    void solve(VectorType& x) {
        if ((size_t)A.getCols() != x.size()) throw std::invalid_argument("Dimension mismatch");  // A * x, SparseObj.hpp:133
        if (b.size() != (size_t)A.getRows()) throw std::invalid_argument("Vector dimensions do not match.");  // b - A*x
        atk_host::OperatorHandle op = atk_host::to_device(A);
        last = atk_gd_info{};
        atk_host::check(atk_gd_solve_host(atk_host::context(), op.get(), b.element(), x.element(), maxIter, tol, &last));
    }
};
