/* oracle/krylov_oracle.c -- see krylov_oracle.h.  TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference CPU algorithm for the sparse Krylov hot path.
 * Parity status: PINNED against the reference's KATs / notebook outputs and against the
 * unmodified reference (oracle/_ref) by tests/test_oracle.py.
 */
#include "krylov_oracle.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

double orc_u01(uint64_t s) {
    s += 0x9E3779B97F4A7C15ull;
    s = (s ^ (s >> 30)) * 0xBF58476D1CE4E5B9ull;
    s = (s ^ (s >> 27)) * 0x94D049BB133111EBull;
    s ^= s >> 31;
    return (double)(s >> 11) * (1.0 / 9007199254740992.0);
}

/* ------------------------------------------------------------------ VectorObj */
/* VectorObj.hpp:105-108  std::inner_product(begin,end,other,T(0)): acc = acc + a[i]*b[i] */
double orc_dot(long n, const double* a, const double* b) {
    double acc = 0.0;
    for (long i = 0; i < n; ++i) acc = acc + a[i] * b[i];
    return acc;
}
/* VectorObj.hpp:52-54 */
double orc_nrm2(long n, const double* a) { return sqrt(orc_dot(n, a, a)); }
/* VectorObj.hpp:64-68 */
void orc_vec_scale(long n, const double* a, double s, double* out) {
    for (long i = 0; i < n; ++i) out[i] = a[i] * s;
}
/* VectorObj.hpp:75-80: true division, throws on zero */
int orc_vec_div(long n, const double* a, double s, double* out) {
    if (s == 0.0) return 2;
    for (long i = 0; i < n; ++i) out[i] = a[i] / s;
    return 0;
}
/* VectorObj.hpp:89-94 */
void orc_vec_add(long n, const double* a, const double* b, double* out) {
    for (long i = 0; i < n; ++i) out[i] = a[i] + b[i];
}
/* VectorObj.hpp:97-102 */
void orc_vec_sub(long n, const double* a, const double* b, double* out) {
    for (long i = 0; i < n; ++i) out[i] = a[i] - b[i];
}

/* ------------------------------------------------------------------ SparseMatrixCSC */
typedef struct {
    int row, col;
    double v;
} trip;
static int trip_cmp(const void* pa, const void* pb) { /* SparseObj.hpp:34-36 Entry::operator< */
    const trip *a = (const trip*)pa, *b = (const trip*)pb;
    if (a->col != b->col) return a->col < b->col ? -1 : 1;
    if (a->row != b->row) return a->row < b->row ? -1 : 1;
    return 0;
}
/* SparseObj.hpp:55-100 */
long orc_csc_from_triplets(int n_rows, int n_cols, long nt, const int* rows, const int* cols, const double* vals,
                           int* col_ptr, int* row_idx, double* out_vals) {
    trip* buf = (trip*)malloc(sizeof(trip) * (size_t)(nt > 0 ? nt : 1));
    long m = 0;
    for (long t = 0; t < nt; ++t) {
        if (rows[t] < 0 || rows[t] >= n_rows || cols[t] < 0 || cols[t] >= n_cols) { /* :68-70 */
            free(buf);
            return -3;
        }
        if (vals[t] != 0.0) { /* :71-73 explicit zeros are not stored */
            buf[m].row = rows[t];
            buf[m].col = cols[t];
            buf[m].v = vals[t];
            ++m;
        }
    }
    qsort(buf, (size_t)m, sizeof(trip), trip_cmp);                  /* :81 */
    for (int c = 0; c <= n_cols; ++c) col_ptr[c] = 0;               /* :83 */
    for (long t = 0; t < m; ++t) ++col_ptr[buf[t].col + 1];         /* :84-86 */
    for (int c = 0; c < n_cols; ++c) col_ptr[c + 1] += col_ptr[c];  /* :88 */
    for (long t = 0; t < m; ++t) {                                  /* :94-97 */
        out_vals[t] = buf[t].v;
        row_idx[t] = buf[t].row;
    }
    free(buf);
    return m;
}
/* SparseObj.hpp:103-129 */
int orc_csc_get(int n_rows, int n_cols, const int* col_ptr, const int* row_idx, const double* vals, int i, int j,
                double* out) {
    if (i < 0 || i >= n_rows || j < 0 || j >= n_cols) return -3;
    int lo = col_ptr[j], hi = col_ptr[j + 1];
    const int end = hi;
    while (lo < hi) { /* std::lower_bound */
        int mid = lo + (hi - lo) / 2;
        if (row_idx[mid] < i) lo = mid + 1;
        else hi = mid;
    }
    *out = (lo != end && row_idx[lo] == i) ? vals[lo] : 0.0;
    return 0;
}
/* SparseObj.hpp:131-152 */
void orc_spmv_csc(int n_rows, int n_cols, const int* col_ptr, const int* row_idx, const double* vals, const double* x,
                  double* y) {
    for (int r = 0; r < n_rows; ++r) y[r] = 0.0;
    for (int c = 0; c < n_cols; ++c) {
        const double xc = x[c];
        if (xc == 0.0) continue; /* :141 */
        for (int k = col_ptr[c]; k < col_ptr[c + 1]; ++k) y[row_idx[k]] += vals[k] * xc;
    }
}

/* ------------------------------------------------------------------ readMatrixMarket */
/* utils.hpp:30-71 */
long orc_read_matrix_market(const char* path, int* n_rows, int* n_cols, long* n_entries, int* col_ptr, int* row_idx,
                            double* vals) {
    FILE* f = fopen(path, "rb");
    if (!f) return -1; /* :34-37 prints and returns */
    int c;
    while ((c = fgetc(f)) == '%') { /* :43 skip every leading '%' line, header included */
        while ((c = fgetc(f)) != EOF && c != '\n') {}
    }
    if (c != EOF) ungetc(c, f);
    int M = 0, N = 0, L = 0;
    if (fscanf(f, "%d %d %d", &M, &N, &L) != 3) { /* :46 */
        fclose(f);
        return -1;
    }
    *n_rows = M;
    *n_cols = N;
    *n_entries = L;
    if (!col_ptr) {
        fclose(f);
        return 0;
    }
    int* rr = (int*)malloc(sizeof(int) * (size_t)(L > 0 ? L : 1));
    int* cc = (int*)malloc(sizeof(int) * (size_t)(L > 0 ? L : 1));
    double* vv = (double*)malloc(sizeof(double) * (size_t)(L > 0 ? L : 1));
    for (int l = 0; l < L; ++l) { /* :62-67: 1-based -> 0-based, symmetry qualifier ignored */
        int m = 0, n = 0;
        double d = 0.0;
        if (fscanf(f, "%d %d %lf", &m, &n, &d) != 3) { m = 1; n = 1; d = 0.0; }
        rr[l] = m - 1;
        cc[l] = n - 1;
        vv[l] = d;
    }
    fclose(f);
    long nnz = orc_csc_from_triplets(M, N, L, rr, cc, vv, col_ptr, row_idx, vals);
    free(rr);
    free(cc);
    free(vv);
    return nnz;
}

/* ------------------------------------------------------------------ Poisson3DSolver */
/* Poisson.hpp:36,53-57 */
void orc_poisson_coeffs(int nx, int ny, int nz, double lx, double ly, double lz, double* cx, double* cy, double* cz,
                        double* cc) {
    const double hx = lx / (nx - 1), hy = ly / (ny - 1), hz = lz / (nz - 1);
    *cx = 1.0 / (hx * hx);
    *cy = 1.0 / (hy * hy);
    *cz = 1.0 / (hz * hz);
    *cc = -2.0 * (*cx + *cy + *cz);
}
static int is_shell(int i, int j, int k, int nx, int ny, int nz) { /* :66-68 */
    return i == 0 || i == nx - 1 || j == 0 || j == ny - 1 || k == 0 || k == nz - 1;
}
long orc_poisson_nnz(int nx, int ny, int nz) {
    const long total = (long)nx * ny * nz;
    const long interior = (nx > 2 && ny > 2 && nz > 2) ? (long)(nx - 2) * (ny - 2) * (nz - 2) : 0;
    return (total - interior) + 7 * interior;
}
/* Poisson.hpp:49-95 followed by finalize(): emitted column by column, rows ascending.
 * Column c=(i,j,k) holds: its own diagonal, plus one entry from every INTERIOR row that
 * neighbours it (boundary rows are identity rows, :73). */
void orc_poisson_csc(int nx, int ny, int nz, double lx, double ly, double lz, int* col_ptr, int* row_idx, double* vals) {
    double cx, cy, cz, cc;
    orc_poisson_coeffs(nx, ny, nz, lx, ly, lz, &cx, &cy, &cz, &cc);
    const long sxy = (long)nx * ny;
    long p = 0;
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
                const long c = i + (long)j * nx + (long)k * sxy;
                col_ptr[c] = (int)p;
                /* rows in ascending order: (k-1),(j-1),(i-1),self,(i+1),(j+1),(k+1) */
                if (k - 1 >= 0 && !is_shell(i, j, k - 1, nx, ny, nz)) { row_idx[p] = (int)(c - sxy); vals[p++] = cz; }
                if (j - 1 >= 0 && !is_shell(i, j - 1, k, nx, ny, nz)) { row_idx[p] = (int)(c - nx); vals[p++] = cy; }
                if (i - 1 >= 0 && !is_shell(i - 1, j, k, nx, ny, nz)) { row_idx[p] = (int)(c - 1); vals[p++] = cx; }
                row_idx[p] = (int)c;
                vals[p++] = is_shell(i, j, k, nx, ny, nz) ? 1.0 : cc;
                if (i + 1 < nx && !is_shell(i + 1, j, k, nx, ny, nz)) { row_idx[p] = (int)(c + 1); vals[p++] = cx; }
                if (j + 1 < ny && !is_shell(i, j + 1, k, nx, ny, nz)) { row_idx[p] = (int)(c + nx); vals[p++] = cy; }
                if (k + 1 < nz && !is_shell(i, j, k + 1, nx, ny, nz)) { row_idx[p] = (int)(c + sxy); vals[p++] = cz; }
            }
    col_ptr[(long)nx * ny * nz] = (int)p;
}
/* The operator of Poisson.hpp:49-95 applied through SparseObj.hpp:131-152 semantics:
 * each y[row] starts at 0.0 and receives val*x[col] in ascending col order, and a column
 * whose x is exactly zero is skipped. */
void orc_poisson_apply(int nx, int ny, int nz, double cx, double cy, double cz, const double* x, double* y) {
    const double cc = -2.0 * (cx + cy + cz);
    const long sxy = (long)nx * ny;
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
                const long c = i + (long)j * nx + (long)k * sxy;
                double acc = 0.0;
                if (is_shell(i, j, k, nx, ny, nz)) {
                    if (x[c] != 0.0) acc += 1.0 * x[c];
                } else {
                    if (x[c - sxy] != 0.0) acc += cz * x[c - sxy];
                    if (x[c - nx] != 0.0) acc += cy * x[c - nx];
                    if (x[c - 1] != 0.0) acc += cx * x[c - 1];
                    if (x[c] != 0.0) acc += cc * x[c];
                    if (x[c + 1] != 0.0) acc += cx * x[c + 1];
                    if (x[c + nx] != 0.0) acc += cy * x[c + nx];
                    if (x[c + sxy] != 0.0) acc += cz * x[c + sxy];
                }
                y[c] = acc;
            }
}
/* Poisson.hpp:103-151 with the two RHS families of SURVEY 8(d) */
void orc_poisson_rhs(int nx, int ny, int nz, double lx, double ly, double lz, int rhs_kind, uint64_t seed, double* b) {
    const double hx = lx / (nx - 1), hy = ly / (ny - 1), hz = lz / (nz - 1);
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
                const double x = i * hx, y = j * hy, z = k * hz;
                const long idx = i + (long)j * nx + (long)k * nx * ny;
                const int shell = is_shell(i, j, k, nx, ny, nz);
                double v;
                if (rhs_kind == 0) {
                    const double s = sin(M_PI * x) * sin(M_PI * y) * sin(M_PI * z);
                    v = shell ? s : -3.0 * M_PI * M_PI * sin(M_PI * x) * sin(M_PI * y) * sin(M_PI * z);
                } else {
                    /* idx recovered from coordinates as a callback would: llround(x/hx) == i */
                    v = shell ? 0.0 : 2.0 * orc_u01(seed + (uint64_t)idx) - 1.0;
                }
                b[idx] = v;
            }
}

/* ------------------------------------------------------------------ 27-point operator */
long orc_op27_nnz(int nx, int ny, int nz) {
    /* sum over points of (#neighbours in grid incl. self) = prod over axes of (3n-2) */
    return (long)(3 * nx - 2) * (3 * ny - 2) * (3 * nz - 2);
}
void orc_op27_csc(int nx, int ny, int nz, int* col_ptr, int* row_idx, double* vals) {
    const long sxy = (long)nx * ny;
    long p = 0;
    for (int kc = 0; kc < nz; ++kc)
        for (int jc = 0; jc < ny; ++jc)
            for (int ic = 0; ic < nx; ++ic) {
                const long c = ic + (long)jc * nx + (long)kc * sxy;
                col_ptr[c] = (int)p;
                /* row r = c - d, d = (di,dj,dk) the offset from row to column; r ascending
                 * means d descending */
                for (int dk = 1; dk >= -1; --dk)
                    for (int dj = 1; dj >= -1; --dj)
                        for (int di = 1; di >= -1; --di) {
                            const int ir = ic - di, jr = jc - dj, kr = kc - dk;
                            if (ir < 0 || ir >= nx || jr < 0 || jr >= ny || kr < 0 || kr >= nz) continue;
                            row_idx[p] = (int)(ir + (long)jr * nx + (long)kr * sxy);
                            vals[p] = (di == 0 && dj == 0 && dk == 0)
                                          ? 40.0
                                          : -1.0 + 0.3 * ((double)di + 0.5 * (double)dj + 0.25 * (double)dk);
                            ++p;
                        }
            }
    col_ptr[(long)nx * ny * nz] = (int)p;
}

/* ------------------------------------------------------------------ ConjugateGrad */
static void apply_op(int n, const int* col_ptr, const int* row_idx, const double* vals, const orc_stencil7* st,
                     const double* x, double* y) {
    if (col_ptr) orc_spmv_csc(n, n, col_ptr, row_idx, vals, x, y);
    else orc_poisson_apply(st->nx, st->ny, st->nz, st->cx, st->cy, st->cz, x, y);
}
/* ConjugateGradient.hpp:36-84 */
int orc_cg_solve(int n, const int* col_ptr, const int* row_idx, const double* vals, const orc_stencil7* st,
                 const double* b, int max_iter, double* x, double* r, double* p, int* iters, double* last_pAp,
                 double* pAp_trace, double* rs_trace, int trace_cap) {
    double rs_old = orc_dot(n, r, r);              /* :37 */
    const double b_norm = orc_nrm2(n, b);          /* :38 */
    *iters = 0;
    *last_pAp = 0.0;
    if (b_norm < 1e-12) return 0;                  /* :39-42 */
    double* Ap = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    int i = 0;
    for (; i < max_iter; ++i) {                    /* :49 */
        apply_op(n, col_ptr, row_idx, vals, st, p, Ap); /* :50 */
        const double pAp = orc_dot(n, p, Ap);      /* :51 */
        *last_pAp = pAp;
        if (pAp_trace && i < trace_cap) pAp_trace[i] = pAp;
        if (fabs(pAp) < 1e-12) break;              /* :53-57 */
        const double alpha = rs_old / pAp;         /* :59 */
        for (int t = 0; t < n; ++t) {              /* :60-61: scale first, then add/sub */
            const double pa = p[t] * alpha;
            x[t] = x[t] + pa;
            const double aa = Ap[t] * alpha;
            r[t] = r[t] - aa;
        }
        const double rs_new = orc_dot(n, r, r);    /* :63 */
        if (rs_trace && i < trace_cap) rs_trace[i] = rs_new;
        const double beta = rs_new / rs_old;       /* :79 */
        for (int t = 0; t < n; ++t) {
            const double pb = p[t] * beta;
            p[t] = r[t] + pb;
        }
        rs_old = rs_new;                           /* :80 */
    }
    *iters = i;
    free(Ap);
    return 0;
}

/* ------------------------------------------------------------------ GMRES */
/* GMRES.hpp:27-171 (usePreconditioner == false) */
/* ---- ILUPreconditioner (src/LinearAlgebra/Preconditioner/ILU.hpp) ----
 * compute() (:21-81) is a full elimination without pivoting carried out through SparseMatrixCSC lookups, so a dense
 * n x n array (column-major, absent entry == 0.0) restates it exactly:
 *   step j: |A(j,j)| < 1e-12 -> "numerically singular" (:38-40); for i > j: factor = A(i,j)/A(j,j); only when
 *   |factor| > 1e-12 (:50) is A(i,j) overwritten by the multiplier (:51) and row i updated, and only entries whose
 *   new value exceeds 1e-12 in magnitude are written (:55-57) - a smaller update leaves the OLD value in place.
 * Entries become L (i > j) or U (i <= j) only when |value| > 1e-12 (:67-76); L has a unit diagonal (:65). */
#define ORC_ILU_EPS 1e-12
int orc_ilu_compute(int n, const int* col_ptr, const int* row_idx, const double* vals, double* LU) {
    const size_t N = (size_t)n;
    memset(LU, 0, sizeof(double) * N * N);
    for (int c = 0; c < n; ++c)
        for (int k = col_ptr[c]; k < col_ptr[c + 1]; ++k) LU[(size_t)row_idx[k] + (size_t)c * N] = vals[k];
#define A_(i, j) LU[(size_t)(i) + (size_t)(j) * N]
    for (int j = 0; j < n; ++j) {
        const double pivot = A_(j, j);
        if (fabs(pivot) < ORC_ILU_EPS) return 2;
        for (int i = j + 1; i < n; ++i) {
            const double factor = A_(i, j) / pivot;
            if (fabs(factor) > ORC_ILU_EPS) {
                A_(i, j) = factor;
                for (int k = j + 1; k < n; ++k) {
                    const double prod = factor * A_(j, k);
                    const double update = A_(i, k) - prod;
                    if (fabs(update) > ORC_ILU_EPS) A_(i, k) = update;
                }
            }
        }
    }
#undef A_
    return 0;
}

/* solve() (:84-112): y_i = b_i - sum_{j<i} L(i,j) y_j ; x_i = (y_i - sum_{j>i} U(i,j) x_j) / U(i,i), every sum from 0.0
 * in ascending j with one rounding per product and per add, over ALL j (zeros included), as the reference loops. */
void orc_ilu_solve(int n, const double* LU, const double* b, double* x) {
    const size_t N = (size_t)n;
    double* y = (double*)malloc(sizeof(double) * (N ? N : 1));
#define M_(i, j) (fabs(LU[(size_t)(i) + (size_t)(j) * N]) > ORC_ILU_EPS ? LU[(size_t)(i) + (size_t)(j) * N] : 0.0)
    for (int i = 0; i < n; ++i) {
        double sum = 0.0;
        for (int j = 0; j < i; ++j) { const double prod = M_(i, j) * y[j]; sum = sum + prod; }
        y[i] = b[i] - sum;
    }
    for (int i = n - 1; i >= 0; --i) {
        double sum = 0.0;
        for (int j = i + 1; j < n; ++j) { const double prod = M_(i, j) * x[j]; sum = sum + prod; }
        x[i] = (y[i] - sum) / M_(i, i);
    }
#undef M_
    free(y);
}

/* dense n x n exports of the two factors as the reference's getLFactor()/getUFactor() hold them */
void orc_ilu_factors(int n, const double* LU, double* L, double* U) {
    const size_t N = (size_t)n;
    for (size_t j = 0; j < N; ++j)
        for (size_t i = 0; i < N; ++i) {
            const double v = LU[i + j * N];
            const double m = fabs(v) > ORC_ILU_EPS ? v : 0.0;
            L[i + j * N] = i == j ? 1.0 : (i > j ? m : 0.0);
            U[i + j * N] = i <= j ? m : 0.0;
        }
}

static int gmres_impl(int n, const int* col_ptr, const int* row_idx, const double* vals, const double* LU, const double* b,
                      double* x, int n_x, int max_iter, int krylov_dim, double tol, double* resid, int resid_cap,
                      int* n_resid, int* status, int* last_kry_update) {
    const int K = krylov_dim;
    *n_resid = 0;
    *status = 0;
    if (last_kry_update) *last_kry_update = K;
    if (n_x != n) return 1;            /* :32-34 */
    if (K <= 0 || K > n) return 1;     /* :35-37 */
    const size_t N = (size_t)n, ld = (size_t)K + 1;
    double* r = (double*)malloc(sizeof(double) * N);
    double* w = (double*)malloc(sizeof(double) * N);
    double* t = (double*)malloc(sizeof(double) * N);
    int rc = 0;
    /* :42 r = b - A*x */
    orc_spmv_csc(n, n, col_ptr, row_idx, vals, x, t);
    orc_vec_sub(n, b, t, r);
    if (LU) { orc_ilu_solve(n, LU, r, t); memcpy(r, t, sizeof(double) * N); }   /* :43-45 r = M^-1 (b - A x) */
    double beta = orc_nrm2(n, r);      /* :47 */
    if (*n_resid < resid_cap) resid[(*n_resid)++] = beta;
    if (beta < tol) {                  /* :49-52 */
        *status = 2;
        free(r); free(w); free(t);
        return 0;
    }
    double* V = (double*)malloc(sizeof(double) * N * ld);          /* :54 */
    double* H = (double*)calloc(ld * ld, sizeof(double));          /* :55 allocated ONCE, never cleared */
    double* cs = (double*)malloc(sizeof(double) * (size_t)K);
    double* sn = (double*)malloc(sizeof(double) * (size_t)K);
    double* rho = (double*)malloc(sizeof(double) * (size_t)K);
    double* e1 = (double*)calloc(ld, sizeof(double));
    double* y = (double*)malloc(sizeof(double) * ld);
    for (int i = 0; i < K; ++i) { cs[i] = 1.0; sn[i] = 0.0; rho[i] = 0.0; }   /* :56-58 */
    e1[0] = beta;                                                  /* :60 */
#define HH(i, j) H[(size_t)(i) + (size_t)(j) * ld]
    for (int iter = 0; iter < max_iter; ++iter) {                  /* :62 */
        if (orc_vec_div(n, r, beta, V) != 0) { rc = 2; goto done; } /* :63 */
        int kry = K;
        for (int j = 0; j < K; ++j) {                              /* :65 */
            if (LU) {                                                  /* :67-68 w = M^-1 (A v_j) */
                orc_spmv_csc(n, n, col_ptr, row_idx, vals, V + N * (size_t)j, t);
                orc_ilu_solve(n, LU, t, w);
            } else
                orc_spmv_csc(n, n, col_ptr, row_idx, vals, V + N * (size_t)j, w);   /* :70 */
            for (int i = 0; i < j; ++i) {                          /* :73  NOTE i<j */
                const double* vi = V + N * (size_t)i;
                const double h = orc_dot(n, vi, w);                /* :74 */
                HH(i, j) = h;                                      /* :75 */
                orc_vec_scale(n, vi, h, t);                        /* :76  w = w - V[i]*h */
                orc_vec_sub(n, w, t, w);
            }
            const double w_norm = orc_nrm2(n, w);                  /* :79 */
            HH(j + 1, j) = w_norm;                                 /* :80 */
            if (w_norm < tol) {                                    /* :81-85 */
                kry = j;
                break;
            }
            if (orc_vec_div(n, w, HH(j + 1, j), V + N * (size_t)(j + 1)) != 0) { rc = 2; goto done; } /* :86 */
            for (int i = 0; i < j; ++i) {                          /* :89-91 */
                const double tmp = HH(i, j);
                HH(i, j) = cs[i] * HH(i, j) + sn[i] * HH(i + 1, j);
                HH(i + 1, j) = -sn[i] * tmp + cs[i] * HH(i + 1, j);
            }
            {                                                      /* :92 / :139-151 */
                const double dx = HH(j, j), dy = HH(j + 1, j);     /* H(j,j): 0 or STALE rho */
                double rh = sqrt(dx * dx + dy * dy);
                if (fabs(rh) < DBL_EPSILON) { cs[j] = 1.0; sn[j] = 0.0; rh = 0.0; }
                else { cs[j] = dx / rh; sn[j] = dy / rh; }
                rho[j] = rh;
            }
            HH(j, j) = rho[j];                                     /* :93 */
            HH(j + 1, j) = 0.0;                                    /* :94 */
            {                                                      /* :95 / :128-132 */
                const double tmp = e1[j];
                e1[j] = cs[j] * tmp + sn[j] * e1[j + 1];
                e1[j + 1] = -sn[j] * tmp + cs[j] * e1[j + 1];
            }
        }
        if (last_kry_update) *last_kry_update = kry;
        /* :98 / :153-171 updateSolution */
        for (int i = kry - 1; i >= 0; --i) {
            y[i] = e1[i];
            for (int j = i + 1; j < kry; ++j) y[i] = y[i] - HH(i, j) * y[j];
            if (fabs(HH(i, i)) < DBL_EPSILON) { rc = 2; goto done; }
            y[i] = y[i] / HH(i, i);
        }
        for (int i = 0; i < kry; ++i) {                            /* :167-170 one sweep per i */
            orc_vec_scale(n, V + N * (size_t)i, y[i], t);
            orc_vec_add(n, x, t, x);
        }
        orc_spmv_csc(n, n, col_ptr, row_idx, vals, x, t);          /* :99 */
        orc_vec_sub(n, b, t, r);
        if (LU) { orc_ilu_solve(n, LU, r, t); memcpy(r, t, sizeof(double) * N); }   /* :100-102 */
        beta = orc_nrm2(n, r);                                     /* :104 */
        if (*n_resid < resid_cap) resid[(*n_resid)++] = beta;
        if (beta < tol) { *status = 1; goto done; }                /* :107-110 */
        memset(e1, 0, sizeof(double) * ld);                        /* :111-112 */
        e1[0] = beta;
    }
#undef HH
done:
    free(V); free(H); free(cs); free(sn); free(rho); free(e1); free(y);
    free(r); free(w); free(t);
    return rc;
}

int orc_gmres_solve(int n, const int* col_ptr, const int* row_idx, const double* vals, const double* b, double* x,
                    int n_x, int max_iter, int krylov_dim, double tol, double* resid, int resid_cap, int* n_resid,
                    int* status, int* last_kry_update) {
    return gmres_impl(n, col_ptr, row_idx, vals, NULL, b, x, n_x, max_iter, krylov_dim, tol, resid, resid_cap, n_resid, status,
                      last_kry_update);
}

int orc_gmres_solve_precond(int n, const int* col_ptr, const int* row_idx, const double* vals, const double* b, double* x,
                            int n_x, int max_iter, int krylov_dim, double tol, double* resid, int resid_cap, int* n_resid,
                            int* status, int* last_kry_update) {
    *n_resid = 0;
    *status = 0;
    if (n_x != n) return 1;                       /* argument checks come before compute() (:28-41) */
    if (krylov_dim <= 0 || krylov_dim > n) return 1;
    double* LU = (double*)malloc(sizeof(double) * (size_t)n * (size_t)n);
    int rc = orc_ilu_compute(n, col_ptr, row_idx, vals, LU);
    if (rc == 0)
        rc = gmres_impl(n, col_ptr, row_idx, vals, LU, b, x, n_x, max_iter, krylov_dim, tol, resid, resid_cap, n_resid, status,
                        last_kry_update);
    free(LU);
    return rc;
}

/* ---- ILU(0): the same elimination restricted to the sparsity pattern of A (no fill-in) ----
 * The reference has no sparse incomplete factorisation (its ILUPreconditioner is the full elimination above); SURVEY 8(f)
 * rank 4 names one as the "next" step.  Definition used here and by the device code: ILU.hpp:21-81's loops and drop rules
 * with every write confined to positions A stores - a multiplier is formed only where (i,j) is stored, an update applied
 * only where (i,k) is stored.  On a matrix whose full elimination creates no fill-in (e.g. tridiagonal) the result is
 * bit-identical to orc_ilu_compute / the reference; tests/test_ilu.py pins exactly that. */
int orc_ilu0_compute(int n, const int* col_ptr, const int* row_idx, const double* vals, double* LU) {
    const size_t N = (size_t)n;
    unsigned char* pat = (unsigned char*)calloc(N * N ? N * N : 1, 1);
    memset(LU, 0, sizeof(double) * N * N);
    for (int c = 0; c < n; ++c)
        for (int k = col_ptr[c]; k < col_ptr[c + 1]; ++k) {
            LU[(size_t)row_idx[k] + (size_t)c * N] = vals[k];
            pat[(size_t)row_idx[k] + (size_t)c * N] = 1;
        }
#define A_(i, j) LU[(size_t)(i) + (size_t)(j) * N]
#define P_(i, j) pat[(size_t)(i) + (size_t)(j) * N]
    int rc = 0;
    for (int j = 0; j < n && rc == 0; ++j) {
        const double pivot = A_(j, j);
        if (fabs(pivot) < ORC_ILU_EPS) { rc = 2; break; }
        for (int i = j + 1; i < n; ++i) {
            if (!P_(i, j)) continue;
            const double factor = A_(i, j) / pivot;
            if (fabs(factor) > ORC_ILU_EPS) {
                A_(i, j) = factor;
                for (int k = j + 1; k < n; ++k) {
                    if (!P_(i, k) || !P_(j, k)) continue;  /* an unstored A(j,k) is 0.0: the update would not change A(i,k) */
                    const double prod = factor * A_(j, k);
                    const double update = A_(i, k) - prod;
                    if (fabs(update) > ORC_ILU_EPS) A_(i, k) = update;
                }
            }
        }
    }
#undef A_
#undef P_
    free(pat);
    return rc;
}

/* GMRES::solve with M = the ILU(0) factors (same loop as orc_gmres_solve_precond, other preconditioner) */
int orc_gmres_solve_precond_ilu0(int n, const int* col_ptr, const int* row_idx, const double* vals, const double* b, double* x,
                                 int n_x, int max_iter, int krylov_dim, double tol, double* resid, int resid_cap, int* n_resid,
                                 int* status, int* last_kry_update) {
    *n_resid = 0;
    *status = 0;
    if (n_x != n) return 1;
    if (krylov_dim <= 0 || krylov_dim > n) return 1;
    double* LU = (double*)malloc(sizeof(double) * (size_t)n * (size_t)n);
    int rc = orc_ilu0_compute(n, col_ptr, row_idx, vals, LU);
    if (rc == 0)
        rc = gmres_impl(n, col_ptr, row_idx, vals, LU, b, x, n_x, max_iter, krylov_dim, tol, resid, resid_cap, n_resid, status,
                        last_kry_update);
    free(LU);
    return rc;
}

/* ------------------------------------------------------------------ GradientDescent */
/* IterSolver.hpp:23-44 */
int orc_gd_solve(int n, const int* col_ptr, const int* row_idx, const double* vals, const double* b, double* x, int max_iter,
                 double tol, int* iters, double* residual_norm, double* change_norm) {
    const size_t N = (size_t)(n > 0 ? n : 1);
    double* r = (double*)malloc(sizeof(double) * N);
    double* Ar = (double*)malloc(sizeof(double) * N);
    double* t = (double*)malloc(sizeof(double) * N);
    double* xold = (double*)malloc(sizeof(double) * N);
    orc_spmv_csc(n, n, col_ptr, row_idx, vals, x, t);            /* :24 r = b - A*x */
    orc_vec_sub(n, b, t, r);
    double rn = orc_nrm2(n, r);                                   /* :25 */
    double cn = 0.0;
    memcpy(xold, x, sizeof(double) * (size_t)n);                  /* :26 */
    int it = 0;
    for (; it < max_iter && rn > tol; ++it) {                     /* :28 */
        orc_spmv_csc(n, n, col_ptr, row_idx, vals, r, Ar);        /* :29 */
        const double alpha = orc_dot(n, r, r) / orc_dot(n, r, Ar); /* :30 */
        orc_vec_scale(n, r, alpha, t);                            /* :32 x = x + r*alpha */
        orc_vec_add(n, x, t, x);
        orc_vec_scale(n, Ar, alpha, t);                           /* :33 r = r - Ar*alpha */
        orc_vec_sub(n, r, t, r);
        rn = orc_nrm2(n, r);                                      /* :35 */
        orc_vec_sub(n, x, xold, t);                               /* :38 */
        cn = orc_nrm2(n, t);
        if (cn < tol) { ++it; break; }                            /* :39-41 (the update of this iteration was applied) */
        memcpy(xold, x, sizeof(double) * (size_t)n);              /* :42 */
    }
    *iters = it;
    *residual_norm = rn;
    *change_norm = cn;
    free(r); free(Ar); free(t); free(xold);
    return 0;
}

/* ------------------------------------------------------------------ Krylov::Arnoldi */
/* KrylovSubspace.hpp:11-38 */
int orc_arnoldi(int n, const int* col_ptr, const int* row_idx, const double* vals, double* Q, int m, double* H, double tol,
                int* breakdown) {
    const size_t N = (size_t)(n > 0 ? n : 1);
    double* av = (double*)malloc(sizeof(double) * N);
    double* t = (double*)malloc(sizeof(double) * N);
    *breakdown = -1;
    for (int i = 1; i < m; ++i) {                                              /* :18 */
        orc_spmv_csc(n, n, col_ptr, row_idx, vals, Q + (size_t)(i - 1) * N, av); /* :19 */
        for (int j = 0; j < i; ++j) {                                          /* :22-26 */
            const double* qj = Q + (size_t)j * N;
            const double h = orc_dot(n, qj, av);
            H[(size_t)j + (size_t)(i - 1) * (size_t)m] = h;
            orc_vec_scale(n, qj, h, t);
            orc_vec_sub(n, av, t, av);
        }
        const double h = orc_nrm2(n, av);                                      /* :29 */
        if (h < tol) { *breakdown = i; break; }                                /* :30-33 */
        H[(size_t)i + (size_t)(i - 1) * (size_t)m] = h;                        /* :35 */
        orc_vec_div(n, av, h, Q + (size_t)i * N);                              /* :36 */
    }
    free(av); free(t);
    return 0;
}
