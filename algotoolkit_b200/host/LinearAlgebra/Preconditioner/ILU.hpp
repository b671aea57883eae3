// ILUPreconditioner -- drop-in for the reference's src/LinearAlgebra/Preconditioner/ILU.hpp (SparseMatrixCSC<double>
// instantiation, the one GMRES::enablePreconditioner() uses).
//
// compute() hands the finished CSC arrays to libatk_cuda (atk_ilu_compute), which carries out the reference's
// elimination - a full LU without pivoting with its 1e-12 drop rules - on the device; solve() is the device's forward
// and back substitution (atk_ilu_solve_host).  getLFactor()/getUFactor() rebuild SparseMatrixCSC objects from the
// device factors, entry for entry what the reference extracts (:64-79).  Same exceptions as the reference:
// invalid_argument for a non-square matrix (:24-26) or a wrong-sized vector (:88-90), runtime_error for a numerically
// singular matrix (:38-40) or solve() before compute() (:85-87).
#pragma once

#include <memory>
#include <stdexcept>
#include <type_traits>
#include <vector>

#include "../../Obj/SparseObj.hpp"
#include "../../Obj/VectorObj.hpp"
#include "../../atk_runtime.hpp"

template <typename TNum, typename MatrixType = SparseMatrixCSC<TNum>>
class ILUPreconditioner {
    static_assert(std::is_same<TNum, double>::value && std::is_same<MatrixType, SparseMatrixCSC<double>>::value,
                  "ILUPreconditioner runs on the GPU for SparseMatrixCSC<double> only");
    struct Release {
        void operator()(atk_preconditioner* p) const { atk_preconditioner_destroy(p); }
    };
    std::shared_ptr<atk_preconditioner> handle_;  // copies of the preconditioner share the device factors
    mutable MatrixType L, U;
    mutable bool factorsFetched = false;
    int n;
    bool isComputed;

    void fetchFactors() const {
        if (factorsFetched) return;
        std::vector<double> l((size_t)n * n), u((size_t)n * n);
        atk_host::check(atk_ilu_export(handle_.get(), l.data(), u.data()));
        L = MatrixType(n, n);
        U = MatrixType(n, n);
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) {
                const double lv = l[(size_t)i + (size_t)j * n], uv = u[(size_t)i + (size_t)j * n];
                if (lv != 0.0) L.addValue(i, j, lv);
                if (uv != 0.0) U.addValue(i, j, uv);
            }
        L.finalize();
        U.finalize();
        factorsFetched = true;
    }

public:
    ILUPreconditioner() : n(0), isComputed(false) {}

    void compute(const MatrixType& Matrix) {
        n = Matrix.getRows();
        if (n != Matrix.getCols()) throw std::invalid_argument("Matrix must be square for ILU factorization");
        isComputed = false;
        factorsFetched = false;
        atk_preconditioner_t h = nullptr;
        atk_host::check(atk_ilu_compute(atk_host::context(), Matrix.getRows(), Matrix.getCols(), (int64_t)Matrix.values.size(),
                                        Matrix.col_ptr.data(), Matrix.row_indices.data(), Matrix.values.data(), &h));
        handle_ = std::shared_ptr<atk_preconditioner>(h, Release{});
        isComputed = true;
    }

    VectorObj<TNum> solve(const VectorObj<TNum>& b) const {
        if (!isComputed) throw std::runtime_error("ILU factorization not computed");
        if (static_cast<int>(b.size()) != n) throw std::invalid_argument("Vector size does not match matrix size");
        VectorObj<TNum> x(n);
        atk_host::check(atk_ilu_solve_host(atk_host::context(), handle_.get(), (int64_t)n, b.element(), x.element()));
        return x;
    }

    const MatrixType& getLFactor() const {
        if (isComputed) fetchFactors();
        return L;
    }
    const MatrixType& getUFactor() const {
        if (isComputed) fetchFactors();
        return U;
    }

    // extension: the device handle, for GMRES::solve
    atk_preconditioner_t device_handle() const { return handle_.get(); }
};
