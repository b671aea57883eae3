// SparseMatrixCSC<T> -- drop-in for the solver-path subset of the reference's src/Obj/SparseObj.hpp.
//
// Host side: the same public storage (_n, _m, values, row_indices, col_ptr, construction_buffer) and the same
// build protocol - addValue() buffers (row, col, value) entries, finalize() sorts them by (col, row) and emits the
// CSC arrays, keeping the buffer so that a later addValue() on a stored position overwrites in place.
// Device side: operator*(VectorObj) uploads the CSC arrays, lets libatk_cuda convert them to CSR and SELL-32 on the
// GPU and runs the SpMV kernel there (atk_operator_from_csc / atk_operator_apply).  Solvers call device_operator()
// once and keep the handle for the whole solve.
//
// Not carried over (outside the solver hot path): sparse*sparse, +, -, Transpose, swapRows.
#pragma once

#include <algorithm>
#include <numeric>
#include <stdexcept>
#include <type_traits>
#include <vector>

#include "../atk_runtime.hpp"
#include "DenseObj.hpp"
#include "VectorObj.hpp"

template <typename TObj>
class SparseMatrixCSC {
public:
    int _n = 0, _m = 0;  // rows, columns
    std::vector<TObj> values;
    std::vector<int> row_indices;
    std::vector<int> col_ptr;

    struct Entry {
        int row;
        int col;
        TObj value;
        bool operator<(const Entry& o) const { return col != o.col ? col < o.col : row < o.row; }
    };
    std::vector<Entry> construction_buffer;

    SparseMatrixCSC() = default;
    SparseMatrixCSC(int rows, int cols) : _n(rows), _m(cols), col_ptr((size_t)cols + 1, 0) {}
    SparseMatrixCSC(const DenseObj<TObj>& dense) : SparseMatrixCSC(dense.getRows(), dense.getCols()) {
        for (int i = 0; i < _n; ++i)
            for (int j = 0; j < _m; ++j)
                if (dense(i, j) != TObj(0)) addValue(i, j, dense(i, j));
        finalize();
    }

    int getRows() const { return _n; }
    int getCols() const { return _m; }

    // Position of (row, col) in the finished arrays, or -1.  Throws out_of_range like the reference's SearchList.
    long find(int row, int col) const {
        if (row < 0 || row >= _n || col < 0 || col >= _m) throw std::out_of_range("Index out of range");
        const auto first = row_indices.begin() + col_ptr[(size_t)col];
        const auto last = row_indices.begin() + col_ptr[(size_t)col + 1];
        const auto it = std::lower_bound(first, last, row);
        return (it != last && *it == row) ? (long)(it - row_indices.begin()) : -1L;
    }

    // reference :55-74 - overwrite a stored entry, otherwise buffer a new non-zero one until finalize()
    void addValue(int row, int col, TObj value) {
        const long pos = find(row, col);
        if (pos >= 0) {
            values[(size_t)pos] = value;
            if ((size_t)pos < construction_buffer.size() && construction_buffer[(size_t)pos].row == row_indices[(size_t)pos])
                construction_buffer[(size_t)pos].value = value;
            return;
        }
        if (value != TObj()) construction_buffer.push_back(Entry{row, col, value});
    }

    // reference :76-100
    void finalize() {
        if (construction_buffer.empty()) return;
        std::sort(construction_buffer.begin(), construction_buffer.end());
        col_ptr.assign((size_t)_m + 1, 0);
        for (const Entry& e : construction_buffer) ++col_ptr[(size_t)e.col + 1];
        std::partial_sum(col_ptr.begin(), col_ptr.end(), col_ptr.begin());
        const size_t nnz = construction_buffer.size();
        values.resize(nnz);
        row_indices.resize(nnz);
        for (size_t k = 0; k < nnz; ++k) {
            values[k] = construction_buffer[k].value;
            row_indices[k] = construction_buffer[k].row;
        }
    }

    // reference :121-129
    TObj operator()(int row, int col) const {
        const long pos = find(row, col);
        return pos >= 0 ? values[(size_t)pos] : TObj();
    }

    // Upload + on-device CSC -> CSR -> SELL conversion.  The handle is independent of this object afterwards.
    atk_host::OperatorHandle device_operator(atk_format format = ATK_FORMAT_SELL) const {
        static_assert(std::is_same<TObj, double>::value, "the device path is provided for double only");
        atk_operator_t h = nullptr;
        atk_host::check(atk_operator_from_csc(atk_host::context(), _n, _m, (int64_t)values.size(), col_ptr.data(),
                                              row_indices.data(), values.data(), /*on_device=*/0, format, &h));
        return atk_host::OperatorHandle(h);
    }

    // y = A*x  (reference :131-152) on the GPU
    VectorObj<TObj> operator*(const VectorObj<TObj>& x) const {
        if ((size_t)_m != x.size()) throw std::invalid_argument("Dimension mismatch");
        atk_host::OperatorHandle op = device_operator();
        atk_host::DeviceArray dx(x.element(), x.size()), dy((size_t)_n);
        atk_host::check(atk_operator_apply(atk_host::context(), op.get(), dx.get(), dy.get()));
        VectorObj<TObj> y(_n);
        dy.download(y.element());
        return y;
    }
};

// DenseObj * vector via the sparse kernel: non-zeros listed column by column (rows ascending inside a column).
template <typename TObj>
VectorObj<TObj> DenseObj<TObj>::operator*(const VectorObj<TObj>& x) const {
    if (cols_ != (int)x.size()) throw std::invalid_argument("Vector size does not match matrix columns.");
    return SparseMatrixCSC<TObj>(*this) * x;
}
