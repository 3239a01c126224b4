/* oracle/shim/lapacke.h -- TEST INFRASTRUCTURE ONLY.
 * Declarations for the LAPACKE symbols named by include/qlten/framework/hp_numeric/lapack.h.
 * Not on the Contract path; only needed so the reference headers parse and link. */
#ifndef QLB200_ORACLE_SHIM_LAPACKE_H
#define QLB200_ORACLE_SHIM_LAPACKE_H
#include <complex>
#define LAPACK_ROW_MAJOR 101
#define LAPACK_COL_MAJOR 102
typedef int lapack_int;
typedef std::complex<float> lapack_complex_float;
typedef std::complex<double> lapack_complex_double;
extern "C" {
#define QLB200_SVD(p, T, R) \
  lapack_int LAPACKE_##p##gesdd(int, char, lapack_int, lapack_int, T *, lapack_int, R *, T *, lapack_int, T *, lapack_int); \
  lapack_int LAPACKE_##p##gesvd(int, char, char, lapack_int, lapack_int, T *, lapack_int, R *, T *, lapack_int, T *, lapack_int, R *); \
  lapack_int LAPACKE_##p##geqrf(int, lapack_int, lapack_int, T *, lapack_int, T *); \
  lapack_int LAPACKE_##p##gelqf(int, lapack_int, lapack_int, T *, lapack_int, T *);
QLB200_SVD(s, float, float) QLB200_SVD(d, double, double)
QLB200_SVD(c, lapack_complex_float, float) QLB200_SVD(z, lapack_complex_double, double)
#define QLB200_ORG(name, T) \
  lapack_int LAPACKE_##name(int, lapack_int, lapack_int, lapack_int, T *, lapack_int, const T *);
QLB200_ORG(sorgqr, float) QLB200_ORG(dorgqr, double) QLB200_ORG(sorglq, float) QLB200_ORG(dorglq, double)
QLB200_ORG(cungqr, lapack_complex_float) QLB200_ORG(zungqr, lapack_complex_double)
QLB200_ORG(cunglq, lapack_complex_float) QLB200_ORG(zunglq, lapack_complex_double)
lapack_int LAPACKE_ssyev(int, char, char, lapack_int, float *, lapack_int, float *);
lapack_int LAPACKE_dsyev(int, char, char, lapack_int, double *, lapack_int, double *);
lapack_int LAPACKE_cheev(int, char, char, lapack_int, lapack_complex_float *, lapack_int, float *);
lapack_int LAPACKE_zheev(int, char, char, lapack_int, lapack_complex_double *, lapack_int, double *);
}
#endif
