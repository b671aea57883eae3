// DenseObj<T> -- column-major dense matrix container, the subset of the reference's src/Obj/DenseObj.hpp that the
// sparse solver path touches (GMRES is instantiated with it in the reference's own unit test, and
// SparseMatrixCSC has a constructor from it).  The dense BLAS path itself (dgemm, reference :139-150) is OUT OF
// SCOPE of this build; the matrix-vector product is routed through the sparse device kernel by listing the non-zero
// entries column by column, which accumulates in the same column-ascending order as a reference dgemv.
#pragma once

#include <stdexcept>
#include <type_traits>
#include <vector>

#include "VectorObj.hpp"

template <typename TObj>
class SparseMatrixCSC;

template <typename TObj>
class DenseObj {
    int rows_ = 0, cols_ = 0;
    std::vector<TObj> a_;  // column-major

    void bounds(int r, int c) const {
        if (r < 0 || c < 0 || r >= rows_ || c >= cols_) throw std::out_of_range("Matrix indices are out of range.");
    }

public:
    DenseObj() = default;
    DenseObj(int n, int m) : rows_(n), cols_(m), a_((size_t)n * m, TObj(0)) {}
    DenseObj(const VectorObj<TObj>& flat, int n, int m) : rows_(n), cols_(m) {
        if ((int)flat.size() != n * m) throw std::invalid_argument("Data size does not match matrix dimensions.");
        a_.assign(flat.element(), flat.element() + (size_t)n * m);
    }
    DenseObj(const std::vector<VectorObj<TObj>>& columns, int n, int m) : rows_(n), cols_(m) {
        if ((int)columns.size() != m) throw std::invalid_argument("Data size does not match matrix dimensions.");
        for (const auto& c : columns) a_.insert(a_.end(), c.element(), c.element() + n);
    }
    DenseObj(const std::vector<TObj>& flat, int n, int m) : rows_(n), cols_(m), a_(flat) {
        if ((int)flat.size() != n * m) throw std::invalid_argument("Data size does not match matrix dimensions.");
    }

    TObj& operator()(int r, int c) {
        bounds(r, c);
        return a_[(size_t)r + (size_t)c * rows_];
    }
    const TObj& operator()(int r, int c) const {
        bounds(r, c);
        return a_[(size_t)r + (size_t)c * rows_];
    }
    TObj& operator[](int i) {
        if (i < 0 || i >= rows_ * cols_) throw std::out_of_range("Matrix indices are out of range.");
        return a_[(size_t)i];
    }
    const TObj& operator[](int i) const {
        if (i < 0 || i >= rows_ * cols_) throw std::out_of_range("Matrix indices are out of range.");
        return a_[(size_t)i];
    }
    TObj* data() { return a_.data(); }
    const TObj* data() const { return a_.data(); }
    int getRows() const { return rows_; }
    int getCols() const { return cols_; }
    int size() const { return (int)a_.size(); }
    void zero() { a_.assign(a_.size(), TObj()); }
    void resize(int n, int m) {
        rows_ = n;
        cols_ = m;
        a_.resize((size_t)n * m);
    }
    // same call shape as the sparse class, so generic code (utils::setMatrixValue, readMatrixMarket) works on both
    void addValue(int r, int c, TObj v) { (*this)(r, c) = v; }
    void finalize() {}

    VectorObj<TObj> getColumn(int c) const {
        if (c < 0 || c >= cols_) throw std::out_of_range("Column index is out of range.");
        return VectorObj<TObj>(a_.data() + (size_t)c * rows_, rows_);
    }

    // y = A*x through the sparse device kernel (defined in SparseObj.hpp, which includes this header)
    VectorObj<TObj> operator*(const VectorObj<TObj>& x) const;
};
