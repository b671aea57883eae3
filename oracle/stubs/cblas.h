/* Minimal CBLAS stand-in, used ONLY to build the checker in oracle/_ref.
 *
 * TEST INFRASTRUCTURE - not part of the product.  The reference's sparse hot
 * path never calls BLAS, but its SparseObj.hpp includes DenseObj.hpp which
 * includes "cblas.h" (reference src/Obj/DenseObj.hpp:10), and OpenBLAS is not
 * installed in this image.  These two routines are plain column-ascending
 * loops so that GMRES<double, DenseObj<double>> accumulates in the same order
 * as the CSC matvec (reference src/Obj/SparseObj.hpp:139-149).
 */
#ifndef ATK_ORACLE_STUB_CBLAS_H
#define ATK_ORACLE_STUB_CBLAS_H

#ifdef __cplusplus
extern "C" {
#endif

enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 };
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 };

/* y = alpha*A*x + beta*y, column-major, no transpose, unit strides only. */
static inline void cblas_dgemv(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE trans,
                               int M, int N, double alpha, const double* A, int lda,
                               const double* X, int incX, double beta, double* Y, int incY) {
    (void)order; (void)trans; (void)incX; (void)incY;
    for (int i = 0; i < M; ++i) Y[i] = (beta == 0.0) ? 0.0 : beta * Y[i];
    for (int j = 0; j < N; ++j) {
        const double xj = alpha * X[j];
        const double* col = A + (long)j * lda;
        for (int i = 0; i < M; ++i) Y[i] += col[i] * xj;
    }
}

/* C = alpha*A*B + beta*C, column-major, no transposes. */
static inline void cblas_dgemm(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE ta, enum CBLAS_TRANSPOSE tb,
                               int M, int N, int K, double alpha, const double* A, int lda,
                               const double* B, int ldb, double beta, double* C, int ldc) {
    (void)order; (void)ta; (void)tb;
    for (int j = 0; j < N; ++j) {
        double* cj = C + (long)j * ldc;
        for (int i = 0; i < M; ++i) cj[i] = (beta == 0.0) ? 0.0 : beta * cj[i];
        for (int k = 0; k < K; ++k) {
            const double bkj = alpha * B[k + (long)j * ldb];
            const double* ak = A + (long)k * lda;
            for (int i = 0; i < M; ++i) cj[i] += ak[i] * bkj;
        }
    }
}

#ifdef __cplusplus
}
#endif
#endif
