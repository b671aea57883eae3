// utils -- drop-in for the solver-path subset of the reference's src/utils.hpp: the Matrix Market reader and
// setMatrixValue.  Host-side I/O only; it keeps the reference's observable behaviour: every leading '%' line is
// skipped INCLUDING the "%%MatrixMarket ... symmetric" header (so a symmetric file yields only its stored triangle),
// indices are 1-based, the two progress lines are printed, and a missing file prints "Unable to open" and returns.
// Opt-in addition: expand_symmetric = true reads the banner and mirrors the stored triangle of a
// "symmetric" / "skew-symmetric" file (and gives "pattern" entries the value 1), so SuiteSparse SPD inputs such as
// bcsstk08 load as the full matrix.  The default keeps the reference's header-ignoring behaviour.
#pragma once

#include <algorithm>
#include <cctype>
#include <fstream>
#include <iostream>
#include <sstream>
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
void readMatrixMarket(std::string filename, MatrixType& matrix, bool expand_symmetric = false) {
    std::ifstream in(filename, std::ios::binary);
    if (!in.is_open()) {
        std::cerr << "Unable to open" << std::endl;
        return;
    }
    int mirror = 0;        // +1 symmetric, -1 skew-symmetric, 0 general (or expansion not asked for)
    bool pattern = false;  // entries carry no value
    if (expand_symmetric && in.peek() == '%') {
        std::string banner;
        std::getline(in, banner);
        std::transform(banner.begin(), banner.end(), banner.begin(), [](unsigned char ch) { return (char)std::tolower(ch); });
        std::istringstream words(banner);
        std::string tag, object, format, field, symmetry;
        words >> tag >> object >> format >> field >> symmetry;
        if (tag == "%%matrixmarket") {
            if (format != "coordinate") throw std::invalid_argument("readMatrixMarket: only coordinate files are supported");
            if (field == "complex" || symmetry == "hermitian")
                throw std::invalid_argument("readMatrixMarket: complex matrices are not supported");
            pattern = field == "pattern";
            mirror = symmetry == "symmetric" ? 1 : symmetry == "skew-symmetric" ? -1 : 0;
        }
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
        in >> r >> c;
        if (pattern) v = T(1);
        else in >> v;
        matrix.addValue(r - 1, c - 1, v);
        if (mirror != 0 && r != c) matrix.addValue(c - 1, r - 1, mirror > 0 ? v : -v);
    }
    matrix.finalize();
}

}  // namespace utils
