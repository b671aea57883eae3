// Poisson3DSolver -- drop-in for the reference's src/PDEs/FDM/Poisson.hpp.
//
// Same constructor, setRHS / setDirichletBC callbacks, solve(maxIter, tol), getSolution, grid getters and
// BoundaryType enum.  The operator is never assembled on the host: with a Dirichlet boundary in any direction (the
// default, and the only case the reference really implements - :69-73) solve() uses the matrix-free 7-point
// kernel of libatk_cuda, which evaluates exactly the rows buildLaplacianMatrix would have produced (identity on the
// boundary shell, 7 taps inside).  Without any Dirichlet direction the reference keeps stencil rows on the boundary
// with the missing neighbours dropped; that variant is assembled ON THE DEVICE.  $ATK_POISSON_ASSEMBLED=1 forces the
// assembled (SELL) operator for the Dirichlet case too, e.g. to compare the two formulations.
#pragma once

#include <cstdlib>
#include <functional>
#include <stdexcept>
#include <type_traits>

#include "../../LinearAlgebra/Krylov/ConjugateGradient.hpp"
#include "../../Obj/SparseObj.hpp"
#include "../../Obj/VectorObj.hpp"

enum class BoundaryType { Dirichlet, Neumann, Periodic };

template <typename TNum>
class Poisson3DSolver {
    static_assert(std::is_same<TNum, double>::value, "Poisson3DSolver runs on the GPU in double precision only");
    int nx, ny, nz;
    double lx, ly, lz;
    double hx, hy, hz;
    BoundaryType bcTypeX, bcTypeY, bcTypeZ;
    VectorObj<TNum> b, u;

    bool dirichlet_shell() const {
        return bcTypeX == BoundaryType::Dirichlet || bcTypeY == BoundaryType::Dirichlet || bcTypeZ == BoundaryType::Dirichlet;
    }
    bool on_boundary(int i, int j, int k) const {
        return i == 0 || i == nx - 1 || j == 0 || j == ny - 1 || k == 0 || k == nz - 1;
    }
    size_t index(int i, int j, int k) const { return (size_t)i + (size_t)j * nx + (size_t)k * nx * ny; }  // reference :98-100

    atk_host::OperatorHandle make_operator() const {
        // coefficients exactly as the reference forms them (:53-57)
        const double cx = 1.0 / (hx * hx), cy = 1.0 / (hy * hy), cz = 1.0 / (hz * hz);
        atk_operator_t h = nullptr;
        if (dirichlet_shell()) {
            const char* force = std::getenv("ATK_POISSON_ASSEMBLED");
            if (force && force[0] == '1')
                atk_host::check(atk_operator_poisson_assembled(atk_host::context(), nx, ny, nz, cx, cy, cz, ATK_FORMAT_SELL, &h));
            else
                atk_host::check(atk_operator_stencil7(atk_host::context(), nx, ny, nz, cx, cy, cz, &h));
        } else {
            // stencil rows everywhere, neighbours outside the grid dropped (:75-88): a 7-entry coefficient table
            double coef[27] = {0.0};
            const double cc = -2.0 * (cx + cy + cz);
            coef[13] = cc;
            coef[12] = coef[14] = cx;  // di = -1, +1
            coef[10] = coef[16] = cy;  // dj = -1, +1
            coef[4] = coef[22] = cz;   // dk = -1, +1
            atk_host::check(atk_operator_stencil27_assembled(atk_host::context(), nx, ny, nz, coef, ATK_FORMAT_SELL, &h));
        }
        return atk_host::OperatorHandle(h);
    }

public:
    atk_host::CgReport last;  // extension: what the device CG reported for the most recent solve()

    Poisson3DSolver(int nx_, int ny_, int nz_, double lx_, double ly_, double lz_,
                    BoundaryType bcX = BoundaryType::Dirichlet, BoundaryType bcY = BoundaryType::Dirichlet,
                    BoundaryType bcZ = BoundaryType::Dirichlet)
        : nx(nx_), ny(ny_), nz(nz_), lx(lx_), ly(ly_), lz(lz_), hx(lx_ / (nx_ - 1)), hy(ly_ / (ny_ - 1)), hz(lz_ / (nz_ - 1)),
          bcTypeX(bcX), bcTypeY(bcY), bcTypeZ(bcZ), b(nx_ * ny_ * nz_, 0.0), u(nx_ * ny_ * nz_, 0.0) {}

    // b = f everywhere (reference :103-129 - both branches assign f)
    void setRHS(const std::function<TNum(double, double, double)>& f) {
        double* bp = b.element();
        for (int k = 0; k < nz; ++k)
            for (int j = 0; j < ny; ++j)
                for (int i = 0; i < nx; ++i) bp[index(i, j, k)] = f(i * hx, j * hy, k * hz);
    }
    // b = g on the boundary shell (reference :132-151)
    void setDirichletBC(const std::function<TNum(double, double, double)>& g) {
        double* bp = b.element();
        for (int k = 0; k < nz; ++k)
            for (int j = 0; j < ny; ++j)
                for (int i = 0; i < nx; ++i)
                    if (on_boundary(i, j, k)) bp[index(i, j, k)] = g(i * hx, j * hy, k * hz);
    }
    // extension: take the right-hand side as an array instead of nx*ny*nz callbacks
    void setRHSValues(const TNum* values) { std::copy(values, values + b.size(), b.element()); }

    // reference :154-157 - a fresh ConjugateGrad(A, b, maxIter, tol) solved into u; here the whole solve is one call
    void solve(int maxIter = 1000, double /*tol*/ = 1e-6) {
        atk_host::OperatorHandle op = make_operator();
        atk_cg_info info{};
        atk_host::check(atk_cg_solve_host(atk_host::context(), op.get(), b.element(), u.element(), nullptr, nullptr,
                                          /*fresh=*/1, maxIter, &info));
        last.iterations = info.iterations;
        last.stopped_on_pAp = info.stopped_on_pAp != 0;
        last.zero_rhs = info.zero_rhs != 0;
        last.last_pAp = info.last_pAp;
        last.rs = info.rs;
        last.device_ms = info.device_ms;
    }

    TNum getSolution(int i, int j, int k) const {
        if (i < 0 || i >= nx || j < 0 || j >= ny || k < 0 || k >= nz) throw std::out_of_range("Grid indices out of range");
        return u.element()[index(i, j, k)];
    }
    const VectorObj<TNum>& getSolution() const { return u; }
    const VectorObj<TNum>& getRHS() const { return b; }  // extension, read-only

    double getDx() const { return hx; }
    double getDy() const { return hy; }
    double getDz() const { return hz; }
    int getNx() const { return nx; }
    int getNy() const { return ny; }
    int getNz() const { return nz; }
};
