"""The reference's README / notebook examples through the drop-in `fastsolver` Python module (GPU)."""
import os
import sys

import numpy as np
import pytest

from tests.conftest import mm_path

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "algotoolkit_b200"))


@pytest.fixture(scope="module")
def fs():
    from algotoolkit_b200 import build_host

    build_host.build()
    import fastsolver

    return fastsolver


def test_vector_and_sparse_basics(fs):
    v = fs.Vector(np.array([3.0, 4.0]))
    assert v.size() == 2 and v.norm() == 5.0 and v[1] == 4.0
    assert np.array_equal(np.array(v), [3.0, 4.0]) and np.array_equal(v.to_numpy(), [3.0, 4.0])  # numpy>=2 __array__
    v[0] = 1.0
    assert v[0] == 1.0
    with pytest.raises(IndexError):
        v[2]
    A = fs.SparseMatrix(3, 3)  # tests/SparseMatrixCSCTest.cpp:47-64
    for (i, j, x) in [(0, 2, 3.0), (1, 0, 4.0), (2, 1, 5.0), (2, 2, 6.0)]:
        A.addValue(i, j, x)
    A.finalize()
    y = A * fs.Vector(np.array([1.0, 2.0, 3.0]))
    assert y.to_numpy().tolist() == [9.0, 4.0, 28.0]
    assert fs.matvec_mul(A, fs.Vector(np.array([1.0, 2.0, 3.0]))).to_numpy().tolist() == [9.0, 4.0, 28.0]
    assert A(2, 1) == 5.0 and A(0, 0) == 0.0 and A.rows() == 3 and A.cols() == 3
    with pytest.raises(IndexError):
        A(3, 0)
    with pytest.raises(ValueError):
        A * fs.Vector(4)
    with pytest.raises(RuntimeError):
        fs.Vector(np.ones(2)) / 0.0


def test_notebook_cg_2x2(fs, golden):
    A = fs.SparseMatrix(2, 2)  # code/test.ipynb cell 2
    for (i, j, x) in [(0, 0, 4.0), (0, 1, 1.0), (1, 0, 1.0), (1, 1, 3.0)]:
        A.addValue(i, j, x)
    A.finalize()
    b = fs.Vector(np.array([1.0, 2.0]))
    cg = fs.ConjugateGrad(A, b, 100, 1e-6)
    x = fs.Vector(2)
    cg.solve(x)
    assert np.allclose(x.to_numpy(), golden["kat"]["cg_2x2_notebook"]["expect"], rtol=1e-12)


def test_notebook_gmres_2x2_stdout(fs, golden, capsys):
    g = golden["kat"]["gmres_2x2_notebook"]  # code/test.ipynb cell 6
    A = fs.SparseMatrix(2, 2)
    for (i, j, x) in [(0, 0, 4.0), (0, 1, 1.0), (1, 0, 1.0), (1, 1, 3.0)]:
        A.addValue(i, j, x)
    A.finalize()
    b, x = fs.Vector(np.array([1.0, 2.0])), fs.Vector(2)
    fs.set_reduction("sequential")
    try:
        fs.GMRES().solve(A, b, x, 100, 1, 1e-6)
    finally:
        fs.set_reduction("tree")
    out = capsys.readouterr().out.splitlines()
    assert out[:6] == g["printed_log_head"]           # the notebooks parse these exact lines
    assert out[-2:] == g["printed_log_tail"]
    assert x.to_numpy().tolist() == g["expect"]
    with pytest.raises(ValueError):
        fs.GMRES().solve(A, b, fs.Vector(3), 10, 1, 1e-6)


def test_readme_gmres_bcsstk08(fs, port, golden, capsys):
    """README.md:64-86 with a seeded x_ext and a small Krylov dimension (K=300 is covered through the C ABI)."""
    matrix = fs.SparseMatrix(1, 1)
    fs.read_matrix_market(mm_path("bcsstk08"), matrix)
    assert capsys.readouterr().out == golden["gmres_mm"]["bcsstk08"]["printed"]
    assert matrix.rows() == 1074 and matrix.nnz() == 7017  # lower triangle only: the header is ignored
    x_ext = fs.Vector(port.u01_array(777, matrix.cols()))
    b = matrix * x_ext
    x = fs.Vector(matrix.cols())
    fs.set_reduction("sequential")
    try:
        fs.GMRES().solve(matrix, b, x, 2, 30, 1e-6)
    finally:
        fs.set_reduction("tree")
    run = golden["gmres_mm"]["bcsstk08"]["runs"]["K30_it2"]
    lines = capsys.readouterr().out.splitlines()
    assert lines[0] == "Initial residual norm: %g" % run["resid"]["all"][0]
    assert lines[-1] == "Reached maximum iterations without convergence."
    import hashlib

    assert hashlib.sha256(x.to_numpy().tobytes()).hexdigest() == run["x_sha256"]


def test_readme_poisson_example(fs, golden):
    """README.md:88-160 / code/PDE.ipynb cell 1, with the Python callbacks of the reference example."""
    nx = ny = nz = 32
    solver = fs.Poisson3DSolver(nx, ny, nz, 1.0, 1.0, 1.0)
    analytic = lambda x, y, z: np.sin(np.pi * x) * np.sin(np.pi * y) * np.sin(np.pi * z)
    solver.setRHS(lambda x, y, z: -3 * np.pi ** 2 * np.sin(np.pi * x) * np.sin(np.pi * y) * np.sin(np.pi * z))
    solver.setDirichletBC(analytic)
    solver.solve(2000, 1e-6)
    assert solver.iterations == 1
    num = solver.getSolutionArray().reshape(nz, ny, nx)
    assert num[3, 4, 5] == solver.getSolution(5, 4, 3)
    g = np.arange(nx) / (nx - 1)
    an = np.sin(np.pi * g)[None, None, :] * np.sin(np.pi * g)[None, :, None] * np.sin(np.pi * g)[:, None, None]
    err = np.abs(num - an)
    assert f"{err.max():.6e}" == golden["poisson_readme"]["notebook_32"]["max_err"]
    assert f"{np.sqrt(np.mean(err ** 2)):.6e}" == golden["poisson_readme"]["notebook_32"]["l2_err"]
    with pytest.raises(IndexError):
        solver.getSolution(32, 0, 0)
    assert solver.getNx() == 32 and abs(solver.getDx() - 1 / 31) < 1e-16


def test_poisson_array_rhs_matches_reference_iterations(fs, port, golden):
    n = 32
    solver = fs.Poisson3DSolver(n, n, n, 1.0, 1.0, 1.0)
    solver.setRHSArray(port.poisson_rhs(n, n, n, rhs_kind=1, seed=12345))
    solver.solve(100000, 1e-6)
    g = golden["poisson_rhsr"]["32"]
    assert abs(solver.iterations - g["iters"]) <= 2
    x = solver.getSolutionArray()
    assert abs(np.linalg.norm(x) - g["x_norm"]) < 1e-10 * g["x_norm"]
