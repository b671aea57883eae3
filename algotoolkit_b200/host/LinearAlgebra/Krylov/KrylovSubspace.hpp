// Krylov::Arnoldi -- drop-in for the reference's src/LinearAlgebra/Krylov/KrylovSubspace.hpp (:11-38): the textbook
// modified Gram-Schmidt Arnoldi recursion (it projects against ALL previous basis vectors, unlike the loop inside
// GMRES::solve).  Q[0] is the start vector; Q[1..m-1] and the entries H(j,i-1), H(i,i-1) the reference writes are filled
// in.  The recursion runs on the GPU (atk_arnoldi_host): one persistent kernel per step where the vectors fit on chip.
// H must offer `H(r,c) = value` (the reference's code needs the same, i.e. DenseObj in practice).
#pragma once

#include <cassert>
#include <iostream>
#include <vector>

#include "../../Obj/DenseObj.hpp"
#include "../../Obj/SparseObj.hpp"
#include "../../Obj/VectorObj.hpp"
#include "ConjugateGradient.hpp"  // atk_host::to_device

namespace Krylov {

template <typename TNum, typename MatrixType, typename VectorType>
void Arnoldi(MatrixType& A, std::vector<VectorType>& Q, MatrixType& H, TNum tol) {
    static_assert(std::is_same<TNum, double>::value, "Krylov::Arnoldi runs on the GPU in double precision only");
    const int m = (int)Q.size();
    const int n = A.getRows();
    assert(H.getRows() == m && H.getCols() == m - 1);
    assert((int)Q[0].size() == n);
    if (m < 2) return;
    std::vector<double> Qflat((size_t)m * n, 0.0), Hflat((size_t)m * (m - 1), 0.0);
    for (int i = 0; i < m; ++i)
        if ((int)Q[i].size() == n) std::copy(Q[i].element(), Q[i].element() + n, Qflat.begin() + (size_t)i * n);
    atk_host::OperatorHandle op = atk_host::to_device(A);
    int breakdown = -1;
    atk_host::check(atk_arnoldi_host(atk_host::context(), op.get(), Qflat.data(), n, m, Hflat.data(), (double)tol, &breakdown));
    const int last = breakdown < 0 ? m - 1 : breakdown;  // steps 1..last ran; a breakdown step wrote its projections only
    for (int i = 1; i <= last; ++i) {
        for (int j = 0; j < i; ++j) H(j, i - 1) = Hflat[(size_t)j + (size_t)(i - 1) * m];
        if (i == breakdown) break;
        H(i, i - 1) = Hflat[(size_t)i + (size_t)(i - 1) * m];
        Q[i] = VectorType(Qflat.data() + (size_t)i * n, n);
    }
    if (breakdown >= 0) std::cout << "Breakdown: Norm below tolerance at iteration " << breakdown << "\n";  // :31
}

}  // namespace Krylov
