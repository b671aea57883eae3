// utils -- drop-in for the solver-path subset of the reference's src/utils.hpp: the Matrix Market reader and
// setMatrixValue.  Host-side I/O only; it keeps the reference's observable behaviour: every leading '%' line is
// skipped INCLUDING the "%%MatrixMarket ... symmetric" header (so a symmetric file yields only its stored triangle),
// indices are 1-based, the two progress lines are printed, and a missing file prints "Unable to open" and returns.
#pragma once

#include <fstream>
#include <iostream>
#include <string>

#include "Obj/DenseObj.hpp"
#include "Obj/SparseObj.hpp"

namespace utils {

template <typename T = double, typename MatrixType = DenseObj<T>>
void setMatrixValue(MatrixType& M, int i, int j, T value) {
    M.addValue(i, j, value);
    M.finalize();
}

// reference :30-71
template <typename T, typename MatrixType>
void readMatrixMarket(std::string filename, MatrixType& matrix) {
    std::ifstream in(filename, std::ios::binary);
    if (!in.is_open()) {
        std::cerr << "Unable to open" << std::endl;
        return;
    }
    while (in.peek() == '%') in.ignore(2048, '\n');
    int rows = 0, cols = 0, entries = 0;
    in >> rows >> cols >> entries;
    std::cout << "M: " << rows << " N: " << cols << " L: " << entries << std::endl;
    matrix = MatrixType(rows, cols);
    std::cout << "Matrix created " << matrix.getRows() << " " << matrix.getCols() << std::endl;
    for (int e = 0; e < entries; ++e) {
        int r = 0, c = 0;
        T v{};
        in >> r >> c >> v;
        matrix.addValue(r - 1, c - 1, v);
    }
    matrix.finalize();
}

}  // namespace utils
