// fastsolver -- Python module with the reference's names (python/pybind.cpp of AlgoToolkit) for the solver hot path,
// built on the CUDA-backed host classes in ../host.  Bound here: Vector, SparseMatrix, DenseMatrix (container),
// ConjugateGrad, GMRES, Poisson3DSolver, BoundaryType, read_matrix_market, matvec_mul.  The reference's other
// bindings (power_iter, AMG, LU, quadrature, RK, Navier-Stokes, FFT Poisson ...) are outside the hot path and are
// not part of this module.
//
// Exceptions map as in the reference build: std::invalid_argument -> ValueError, std::out_of_range -> IndexError,
// std::runtime_error -> RuntimeError.  Differences a Python user can see, all additive:
//   * Vector.__array__ accepts numpy>=2's dtype= / copy= keywords (the reference binding rejects them);
//   * Poisson3DSolver.setRHSArray / getSolutionArray move whole ndarrays instead of nx*ny*nz callbacks;
//   * set_reduction("sequential" | "tree") selects reference-ordered or fast deterministic dot products;
//   * ConjugateGrad.iterations / Poisson3DSolver.iterations report what the device loop did.
#include <pybind11/functional.h>
#include <pybind11/iostream.h>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <cstring>

#include "../host/LinearAlgebra/Krylov/ConjugateGradient.hpp"
#include "../host/LinearAlgebra/Krylov/GMRES.hpp"
#include "../host/LinearAlgebra/Krylov/KrylovSubspace.hpp"
#include "../host/LinearAlgebra/Solver/IterSolver.hpp"
#include "../host/Obj/DenseObj.hpp"
#include "../host/Obj/SparseObj.hpp"
#include "../host/Obj/VectorObj.hpp"
#include "../host/PDEs/FDM/Poisson.hpp"
#include "../host/utils.hpp"

namespace py = pybind11;

class fastsolverError : public std::runtime_error {
public:
    using std::runtime_error::runtime_error;
};

using Vec = VectorObj<double>;
using Csc = SparseMatrixCSC<double>;
using Dense = DenseObj<double>;
using Cg = ConjugateGrad<double, Csc, Vec>;
using Gm = GMRES<double, Csc, Vec>;
using Poisson = Poisson3DSolver<double>;

static py::array_t<double> to_numpy(const Vec& v) {
    py::array_t<double> out((py::ssize_t)v.size());
    if (v.size()) std::memcpy(out.mutable_data(), v.element(), v.size() * sizeof(double));
    return out;
}

PYBIND11_MODULE(fastsolver, m) {
    m.doc() = "CUDA (sm_100a) build of the FASTSolver sparse iterative-solver path";
    py::register_exception<fastsolverError>(m, "fastsolverError");

    m.def("set_reduction", [](const std::string& mode) {
        if (mode == "sequential") atk_host::Runtime::instance().set_reduction(ATK_REDUCE_SEQUENTIAL);
        else if (mode == "tree") atk_host::Runtime::instance().set_reduction(ATK_REDUCE_TREE);
        else throw std::invalid_argument("set_reduction: expected 'sequential' or 'tree'");
    }, "Order of dot-product sums: 'sequential' (bit-identical to the CPU reference) or 'tree' (fast, deterministic)", py::arg("mode"));

    m.def("matvec_mul", [](const Csc& A, const Vec& x) { return A * x; }, "Matrix-vector multiplication", py::arg("A"), py::arg("x"));

    py::class_<Vec>(m, "Vector")
        .def(py::init<int>(), py::arg("size"))
        .def(py::init([](py::array_t<double, py::array::c_style | py::array::forcecast> a) {
            auto buf = a.request();
            if (buf.ndim != 1) throw std::runtime_error("Number of dimensions must be 1");
            Vec v((int)buf.shape[0]);
            if (buf.shape[0]) std::memcpy(v.element(), buf.ptr, (size_t)buf.shape[0] * sizeof(double));
            return v;
        }))
        .def("__array__", [](const Vec& v, py::object /*dtype*/, py::object /*copy*/) { return to_numpy(v); },
             py::arg("dtype") = py::none(), py::arg("copy") = py::none())
        .def("to_numpy", &to_numpy)
        .def("__getitem__", [](const Vec& v, size_t i) -> double {
            if (i >= v.size()) throw py::index_error("Index out of range");
            return v.element()[i];
        })
        .def("__setitem__", [](Vec& v, size_t i, double x) {
            if (i >= v.size()) throw py::index_error("Index out of range");
            v.element()[i] = x;
        })
        .def("__len__", &Vec::size)
        .def("size", &Vec::size, "Get the size of the vector.")
        .def("norm", &Vec::L2norm, "Compute the L2 norm of the vector.")
        .def("dot", [](const Vec& a, const Vec& b) { return a * b; }, py::arg("other"))
        .def("__add__", [](const Vec& a, const Vec& b) { return a + b; }, py::is_operator())
        .def("__sub__", [](const Vec& a, const Vec& b) { return a - b; }, py::is_operator())
        .def("__mul__", [](const Vec& a, double s) { return a * s; }, py::is_operator())
        .def("__truediv__", [](const Vec& a, double s) { return a / s; }, py::is_operator());

    py::class_<Dense>(m, "DenseMatrix")
        .def(py::init<int, int>(), py::arg("rows"), py::arg("cols"))
        .def("__setitem__", [](Dense& A, std::pair<int, int> ij, double v) { A(ij.first, ij.second) = v; })
        .def("__getitem__", [](const Dense& A, std::pair<int, int> ij) { return A(ij.first, ij.second); })
        .def("rows", &Dense::getRows)
        .def("cols", &Dense::getCols)
        .def("__len__", &Dense::getRows)
        .def("__mul__", [](const Dense& A, const Vec& x) { return A * x; }, py::is_operator());

    py::class_<Csc>(m, "SparseMatrix")
        .def(py::init<int, int>(), py::arg("rows"), py::arg("cols"))
        .def(py::init<const Dense&>(), py::arg("dense_matrix"))
        .def("addValue", &Csc::addValue, py::arg("row"), py::arg("col"), py::arg("value"))
        .def("finalize", &Csc::finalize, "Finalize the sparse matrix construction.")
        .def("rows", &Csc::getRows, "Get the number of rows.")
        .def("cols", &Csc::getCols, "Get the number of columns.")
        .def("nnz", [](const Csc& A) { return A.values.size(); })
        .def("__mul__", [](const Csc& A, const Vec& x) { return A * x; }, py::is_operator())
        .def("__call__", [](const Csc& A, int i, int j) { return A(i, j); }, py::arg("i"), py::arg("j"));

    m.def("read_matrix_market", [](const std::string& filename, py::object matrix, bool expand_symmetric) {
        py::scoped_ostream_redirect out(std::cout, py::module_::import("sys").attr("stdout"));
        if (py::isinstance<Dense>(matrix)) utils::readMatrixMarket<double, Dense>(filename, matrix.cast<Dense&>(), expand_symmetric);
        else if (py::isinstance<Csc>(matrix)) utils::readMatrixMarket<double, Csc>(filename, matrix.cast<Csc&>(), expand_symmetric);
        else throw fastsolverError("Unsupported matrix type");
    }, "Read matrix from Matrix Market format file (expand_symmetric=True mirrors the stored triangle of a symmetric file)",
          py::arg("filename"), py::arg("matrix"), py::arg("expand_symmetric") = false);

    py::class_<Cg>(m, "ConjugateGrad")
        .def(py::init<const Csc&, const Vec&, int, double>(), py::arg("A"), py::arg("b"), py::arg("max_iter"), py::arg("tol"))
        .def("solve", &Cg::solve, "Solve the linear system using the Conjugate Gradient method.", py::arg("x"))
        .def_property_readonly("iterations", [](const Cg& c) { return c.last.iterations; })
        .def_property_readonly("stopped_on_pAp", [](const Cg& c) { return c.last.stopped_on_pAp; });

    using Gd = GradientDescent<double, Csc, Vec>;
    py::class_<Gd>(m, "GradientDescent")
        .def(py::init<const Csc&, const Vec&, int, double>(), py::arg("A"), py::arg("b"), py::arg("max_iter"), py::arg("tol"),
             py::keep_alive<1, 2>(), py::keep_alive<1, 3>())  // the solver keeps references, as the reference class does
        .def("solve", &Gd::solve, "Solve the linear system using Gradient Descent.", py::arg("x"))
        .def_property_readonly("iterations", [](const Gd& g) { return g.last.iterations; });

    m.def("arnoldi", [](Dense& A, std::vector<Vec>& Q, Dense& H, double tol) {
        py::scoped_ostream_redirect out(std::cout, py::module_::import("sys").attr("stdout"));
        Krylov::Arnoldi<double, Dense, Vec>(A, Q, H, tol);
        return Q;  // pybind11 passes std::vector by value: hand the filled basis back as the result
    }, "Arnoldi iteration for generating an orthogonal basis of the Krylov subspace (returns the filled basis).",
       py::arg("A"), py::arg("Q"), py::arg("H"), py::arg("tol"));

    py::class_<Gm>(m, "GMRES")
        .def(py::init<>())
        .def("enablePreconditioner", &Gm::enablePreconditioner, "Enable preconditioning for GMRES.")
        .def("solve", [](Gm& g, const Csc& A, const Vec& b, Vec& x, int maxIter, int KrylovDim, double tol) {
            py::scoped_ostream_redirect out(std::cout, py::module_::import("sys").attr("stdout"));
            g.solve(A, b, x, maxIter, KrylovDim, tol);
        }, "Solve the linear system using GMRES.", py::arg("A"), py::arg("b"), py::arg("x"), py::arg("maxIter"),
           py::arg("KrylovDim"), py::arg("tol"))
        .def_property_readonly("restarts", [](const Gm& g) { return g.last.restarts; })
        .def_property_readonly("status", [](const Gm& g) { return g.last.status; });

    py::enum_<BoundaryType>(m, "BoundaryType")
        .value("Dirichlet", BoundaryType::Dirichlet)
        .value("Neumann", BoundaryType::Neumann)
        .value("Periodic", BoundaryType::Periodic)
        .export_values();

    py::class_<Poisson>(m, "Poisson3DSolver")
        .def(py::init<int, int, int, double, double, double, BoundaryType, BoundaryType, BoundaryType>(), py::arg("nx"),
             py::arg("ny"), py::arg("nz"), py::arg("lx"), py::arg("ly"), py::arg("lz"),
             py::arg("bcTypeX") = BoundaryType::Dirichlet, py::arg("bcTypeY") = BoundaryType::Dirichlet,
             py::arg("bcTypeZ") = BoundaryType::Dirichlet)
        .def("setRHS", [](Poisson& s, py::function f) {
            s.setRHS([f](double x, double y, double z) -> double { return py::cast<double>(f(x, y, z)); });
        }, "Set the right-hand side function f", py::arg("f"))
        .def("setDirichletBC", [](Poisson& s, py::function g) {
            s.setDirichletBC([g](double x, double y, double z) -> double { return py::cast<double>(g(x, y, z)); });
        }, "Set Dirichlet boundary conditions", py::arg("g"))
        .def("setRHSArray", [](Poisson& s, py::array_t<double, py::array::c_style | py::array::forcecast> a) {
            if ((size_t)a.size() != s.getRHS().size()) throw std::invalid_argument("setRHSArray: expected nx*ny*nz values");
            s.setRHSValues(a.data());
        }, "Set b directly (index i + j*nx + k*nx*ny); replaces setRHS + setDirichletBC for large grids", py::arg("values"))
        .def("solve", &Poisson::solve, "Solve the Poisson equation using Conjugate Gradient", py::arg("maxIter") = 1000,
             py::arg("tol") = 1e-6)
        .def("getSolution", py::overload_cast<int, int, int>(&Poisson::getSolution, py::const_), py::arg("i"), py::arg("j"),
             py::arg("k"))
        .def("getSolution", [](const Poisson& s) { return s.getSolution(); }, "Get the entire solution vector")
        .def("getSolutionArray", [](const Poisson& s) { return to_numpy(s.getSolution()); })
        .def_property_readonly("iterations", [](const Poisson& s) { return s.last.iterations; })
        .def("getDx", &Poisson::getDx)
        .def("getDy", &Poisson::getDy)
        .def("getDz", &Poisson::getDz)
        .def("getNx", &Poisson::getNx)
        .def("getNy", &Poisson::getNy)
        .def("getNz", &Poisson::getNz);
}
