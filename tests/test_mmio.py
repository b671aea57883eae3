"""Matrix Market readers: the reference's header-ignoring default and the opt-in symmetric
expansion, numpy reader and the C++ host reader behind fastsolver.read_matrix_market.  Host logic only - no GPU."""
import os
import sys

import numpy as np
import pytest
import scipy.io
import scipy.sparse as sp

from algotoolkit_b200.mmio import read_matrix_market_csc
from tests.conftest import ROOT, mm_path


def _scipy(d):
    return sp.csc_matrix((d["vals"], d["row_idx"], d["col_ptr"]), shape=(d["n_rows"], d["n_cols"]))


@pytest.mark.parametrize("name,stored", [("bcsstk01", 224), ("bcsstk08", 7017)])
def test_default_keeps_reference_behaviour(port, name, stored):
    d = read_matrix_market_csc(mm_path(name))
    assert d["vals"].size == stored  # the stored triangle only (reference utils.hpp:30-71 skips the banner)
    A = port.read_matrix_market(mm_path(name))
    assert np.array_equal(A.col_ptr, d["col_ptr"]) and np.array_equal(A.row_idx, d["row_idx"]) and np.array_equal(A.vals, d["vals"])


@pytest.mark.parametrize("name", ["bcsstk01", "bcsstk08"])
def test_symmetric_expansion_matches_scipy(name):
    d = read_matrix_market_csc(mm_path(name), expand_symmetric=True)
    A = _scipy(d)
    ref = sp.csc_matrix(scipy.io.mmread(mm_path(name), spmatrix=True))
    assert abs(A - ref).max() == 0.0 and abs(A - A.T).max() == 0.0
    assert np.all(np.diff(d["col_ptr"]) >= 1)
    for c in range(d["n_cols"]):  # (column, row) order as SparseMatrixCSC::finalize leaves it
        seg = d["row_idx"][d["col_ptr"][c]:d["col_ptr"][c + 1]]
        assert np.all(np.diff(seg) > 0)


def test_skew_pattern_and_general(tmp_path):
    skew = tmp_path / "skew.mtx"
    skew.write_text("%%MatrixMarket matrix coordinate real skew-symmetric\n% comment\n3 3 2\n2 1 5.0\n3 2 -1.5\n")
    A = _scipy(read_matrix_market_csc(str(skew), expand_symmetric=True)).toarray()
    assert np.array_equal(A, np.array([[0, -5.0, 0], [5.0, 0, 1.5], [0, -1.5, 0]]))
    pat = tmp_path / "pat.mtx"
    pat.write_text("%%MatrixMarket matrix coordinate pattern symmetric\n3 3 3\n1 1\n3 1\n2 2\n")
    P = _scipy(read_matrix_market_csc(str(pat), expand_symmetric=True)).toarray()
    assert np.array_equal(P, np.array([[1.0, 0, 1], [0, 1, 0], [1, 0, 0]]))
    gen = tmp_path / "gen.mtx"
    gen.write_text("%%MatrixMarket matrix coordinate real general\n2 3 2\n1 3 2.5\n2 1 -1\n")
    for flag in (False, True):
        G = _scipy(read_matrix_market_csc(str(gen), expand_symmetric=flag)).toarray()
        assert np.array_equal(G, np.array([[0, 0, 2.5], [-1.0, 0, 0]]))
    cplx = tmp_path / "c.mtx"
    cplx.write_text("%%MatrixMarket matrix coordinate complex hermitian\n1 1 1\n1 1 1.0 0.0\n")
    with pytest.raises(ValueError):
        read_matrix_market_csc(str(cplx), expand_symmetric=True)


def _fastsolver():
    sys.path.insert(0, os.path.join(ROOT, "algotoolkit_b200"))
    try:
        import fastsolver
    except ImportError:
        pytest.skip("fastsolver module not built (python -m algotoolkit_b200.build_host)")
    return fastsolver


@pytest.mark.parametrize("name", ["bcsstk01", "bcsstk08"])
def test_cpp_reader_matches_numpy_reader(name, tmp_path, capsys):
    fs = _fastsolver()
    for flag in (False, True):
        d = read_matrix_market_csc(mm_path(name), expand_symmetric=flag)
        M = fs.SparseMatrix(1, 1)
        if flag:
            fs.read_matrix_market(mm_path(name), M, expand_symmetric=True)
        else:
            fs.read_matrix_market(mm_path(name), M)  # the reference's two-argument call
        assert (M.rows(), M.cols(), M.nnz()) == (d["n_rows"], d["n_cols"], d["vals"].size)
        A = _scipy(d)
        rng = np.random.default_rng(5)
        for i, j in rng.integers(0, d["n_rows"], size=(200, 2)):
            assert M(int(i), int(j)) == A[int(i), int(j)]
    out = capsys.readouterr().out
    assert "M: %d N: %d L: " % (d["n_rows"], d["n_cols"]) in out and "Matrix created" in out
    skew = tmp_path / "skew.mtx"
    skew.write_text("%%MatrixMarket matrix coordinate real skew-symmetric\n3 3 2\n2 1 5.0\n3 2 -1.5\n")
    S = fs.SparseMatrix(1, 1)
    fs.read_matrix_market(str(skew), S, True)
    assert (S(0, 1), S(1, 0), S(1, 2), S(2, 1)) == (-5.0, 5.0, 1.5, -1.5)
