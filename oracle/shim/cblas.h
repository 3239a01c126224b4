/* oracle/shim/cblas.h -- TEST INFRASTRUCTURE ONLY.
 * Declarations of the CBLAS entry points the reference's hp_numeric wrappers call
 * (include/qlten/framework/hp_numeric/blas_level1.h, blas_level3.h, blas_extensions.h);
 * resolved at link time by the OpenBLAS 0.3.15 shared object that ships in this image. */
#ifndef QLB200_ORACLE_SHIM_CBLAS_H
#define QLB200_ORACLE_SHIM_CBLAS_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_ORDER;
typedef CBLAS_ORDER CBLAS_LAYOUT;
typedef enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113, CblasConjNoTrans = 114 } CBLAS_TRANSPOSE;
typedef int blasint;
void openblas_set_num_threads(int);
int openblas_get_num_threads(void);
#define QLB200_GEMM(p, T, S) void cblas_##p##gemm(CBLAS_ORDER, CBLAS_TRANSPOSE, CBLAS_TRANSPOSE, blasint, blasint, blasint, S alpha, const T *A, blasint lda, const T *B, blasint ldb, S beta, T *C, blasint ldc);
QLB200_GEMM(s, float, float) QLB200_GEMM(d, double, double) QLB200_GEMM(c, void, const void *) QLB200_GEMM(z, void, const void *)
void cblas_saxpy(blasint, float, const float *, blasint, float *, blasint);
void cblas_daxpy(blasint, double, const double *, blasint, double *, blasint);
void cblas_caxpy(blasint, const void *, const void *, blasint, void *, blasint);
void cblas_zaxpy(blasint, const void *, const void *, blasint, void *, blasint);
void cblas_scopy(blasint, const float *, blasint, float *, blasint);
void cblas_dcopy(blasint, const double *, blasint, double *, blasint);
void cblas_ccopy(blasint, const void *, blasint, void *, blasint);
void cblas_zcopy(blasint, const void *, blasint, void *, blasint);
void cblas_sscal(blasint, float, float *, blasint);
void cblas_dscal(blasint, double, double *, blasint);
void cblas_cscal(blasint, const void *, void *, blasint);
void cblas_zscal(blasint, const void *, void *, blasint);
float cblas_sdot(blasint, const float *, blasint, const float *, blasint);
double cblas_ddot(blasint, const double *, blasint, const double *, blasint);
void cblas_cdotc_sub(blasint, const void *, blasint, const void *, blasint, void *);
void cblas_zdotc_sub(blasint, const void *, blasint, const void *, blasint, void *);
float cblas_snrm2(blasint, const float *, blasint);
double cblas_dnrm2(blasint, const double *, blasint);
float cblas_scnrm2(blasint, const void *, blasint);
double cblas_dznrm2(blasint, const void *, blasint);
void cblas_sger(CBLAS_ORDER, blasint, blasint, float, const float *, blasint, const float *, blasint, float *, blasint);
void cblas_dger(CBLAS_ORDER, blasint, blasint, double, const double *, blasint, const double *, blasint, double *, blasint);
void cblas_cgeru(CBLAS_ORDER, blasint, blasint, const void *, const void *, blasint, const void *, blasint, void *, blasint);
void cblas_zgeru(CBLAS_ORDER, blasint, blasint, const void *, const void *, blasint, const void *, blasint, void *, blasint);
void cblas_somatcopy(CBLAS_ORDER, CBLAS_TRANSPOSE, blasint, blasint, float, const float *, blasint, float *, blasint);
void cblas_domatcopy(CBLAS_ORDER, CBLAS_TRANSPOSE, blasint, blasint, double, const double *, blasint, double *, blasint);
void cblas_comatcopy(CBLAS_ORDER, CBLAS_TRANSPOSE, blasint, blasint, const float *, const float *, blasint, float *, blasint);
void cblas_zomatcopy(CBLAS_ORDER, CBLAS_TRANSPOSE, blasint, blasint, const double *, const double *, blasint, double *, blasint);
#ifdef __cplusplus
}
#endif
#endif
