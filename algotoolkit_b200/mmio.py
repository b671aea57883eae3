"""Matrix Market reader with the REFERENCE's semantics (src/utils.hpp:30-71), in numpy, for Python-side drivers:
every leading '%' line is skipped including the "%%MatrixMarket ... symmetric" header (a symmetric file therefore
yields only its stored triangle), indices are 1-based, explicit zeros are dropped, and entries end up as CSC sorted by
(column, row).  Host-side input handling only - no solver arithmetic.

Opt-in addition: expand_symmetric=True reads the banner and mirrors the stored triangle of a
"symmetric" / "skew-symmetric" file ("pattern" entries get the value 1); the default keeps the reference's behaviour."""
from __future__ import annotations

import numpy as np


def read_matrix_market_csc(path: str, expand_symmetric: bool = False):
    with open(path, "rb") as f:
        lines = f.read().decode().splitlines()
    mirror, pattern = 0, False
    if expand_symmetric and lines and lines[0].lower().startswith("%%matrixmarket"):
        words = lines[0].lower().split()
        fmt, field, symmetry = (words + [""] * 5)[2:5]
        if fmt != "coordinate":
            raise ValueError("readMatrixMarket: only coordinate files are supported")
        if field == "complex" or symmetry == "hermitian":
            raise ValueError("readMatrixMarket: complex matrices are not supported")
        pattern = field == "pattern"
        mirror = 1 if symmetry == "symmetric" else -1 if symmetry == "skew-symmetric" else 0
    k = 0
    while k < len(lines) and lines[k].startswith("%"):
        k += 1
    n_rows, n_cols, n_entries = (int(t) for t in lines[k].split()[:3])
    width = 2 if pattern else 3
    body = np.array(" ".join(lines[k + 1:]).split(), dtype=np.float64)[: width * n_entries].reshape(n_entries, width)
    rows = body[:, 0].astype(np.int64) - 1
    cols = body[:, 1].astype(np.int64) - 1
    vals = np.ones(n_entries) if pattern else body[:, 2].copy()
    if mirror:
        off = rows != cols  # file order is kept: each entry is followed by its mirror, as the C++ reader adds them
        idx = np.arange(n_entries)
        order = np.argsort(np.concatenate([idx, idx[off] + 0.5]), kind="stable")
        rows, cols, vals = (np.concatenate([rows, cols[off]])[order], np.concatenate([cols, rows[off]])[order],
                            np.concatenate([vals, mirror * vals[off]])[order])
    if n_entries and (rows.min() < 0 or rows.max() >= n_rows or cols.min() < 0 or cols.max() >= n_cols):
        raise IndexError("Index out of range")
    keep = vals != 0.0
    rows, cols, vals = rows[keep], cols[keep], vals[keep]
    order = np.lexsort((rows, cols))
    rows, cols, vals = rows[order], cols[order], vals[order]
    col_ptr = np.zeros(n_cols + 1, np.int32)
    np.add.at(col_ptr, cols + 1, 1)
    col_ptr = np.cumsum(col_ptr).astype(np.int32)
    return dict(n_rows=n_rows, n_cols=n_cols, col_ptr=col_ptr, row_idx=rows.astype(np.int32), vals=vals)
