/* TEST INFRASTRUCTURE ONLY (oracle build) -- not part of the product path.
 * Stand-in for Intel MKL's "mkl.h" (included unconditionally by the reference, src/util/shared.h:24).
 * Maps the five CBLAS/LAPACKE entry points the reference's hot path calls
 * (src/blas/interface.hpp:54,74,92; src/lapack/interface.hpp:39,54) onto the LP64 `scipy_`-prefixed
 * symbols of the OpenBLAS that ships inside scipy's wheel (no MKL / cblas.h / lapacke.h in this image). */
#ifndef CAPITAL_ORACLE_MKL_SHIM_H
#define CAPITAL_ORACLE_MKL_SHIM_H
#ifdef __cplusplus
extern "C" {
#endif
typedef enum { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_ORDER;
typedef enum { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 } CBLAS_TRANSPOSE;
typedef enum { CblasUpper = 121, CblasLower = 122 } CBLAS_UPLO;
typedef enum { CblasNonUnit = 131, CblasUnit = 132 } CBLAS_DIAG;
typedef enum { CblasLeft = 141, CblasRight = 142 } CBLAS_SIDE;
typedef CBLAS_ORDER CBLAS_LAYOUT;
#define LAPACK_ROW_MAJOR 101
#define LAPACK_COL_MAJOR 102
typedef int lapack_int;
void scipy_cblas_dgemm(CBLAS_ORDER, CBLAS_TRANSPOSE, CBLAS_TRANSPOSE, int m, int n, int k, double alpha,
                       const double* a, int lda, const double* b, int ldb, double beta, double* c, int ldc);
void scipy_cblas_dtrmm(CBLAS_ORDER, CBLAS_SIDE, CBLAS_UPLO, CBLAS_TRANSPOSE, CBLAS_DIAG, int m, int n, double alpha,
                       const double* a, int lda, double* b, int ldb);
void scipy_cblas_dsyrk(CBLAS_ORDER, CBLAS_UPLO, CBLAS_TRANSPOSE, int n, int k, double alpha, const double* a, int lda,
                       double beta, double* c, int ldc);
int scipy_LAPACKE_dpotrf(int layout, char uplo, int n, double* a, int lda);
int scipy_LAPACKE_dtrtri(int layout, char uplo, char diag, int n, double* a, int lda);
int scipy_LAPACKE_dgeqrf(int layout, int m, int n, double* a, int lda, double* tau);
int scipy_LAPACKE_dorgqr(int layout, int m, int n, int k, double* a, int lda, const double* tau);
void scipy_openblas_set_num_threads(int);
char* scipy_openblas_get_config(void);
#ifdef __cplusplus
}
#endif
#define cblas_dgemm scipy_cblas_dgemm
#define cblas_dtrmm scipy_cblas_dtrmm
#define cblas_dsyrk scipy_cblas_dsyrk
#define LAPACKE_dpotrf scipy_LAPACKE_dpotrf
#define LAPACKE_dtrtri scipy_LAPACKE_dtrtri
#define LAPACKE_dgeqrf scipy_LAPACKE_dgeqrf
#define LAPACKE_dorgqr scipy_LAPACKE_dorgqr
#endif
