// cblas declarations bound to the OpenBLAS inside the SciPy wheel (symbols scipy_cblas_*), the reference's BLAS := open
#pragma once
enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 };
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 };
#define VV_CBLAS(name) scipy_cblas_##name
#define cblas_sgemm VV_CBLAS(sgemm)
#define cblas_dgemm VV_CBLAS(dgemm)
#define cblas_sgemv VV_CBLAS(sgemv)
#define cblas_dgemv VV_CBLAS(dgemv)
#define cblas_saxpy VV_CBLAS(saxpy)
#define cblas_daxpy VV_CBLAS(daxpy)
#define cblas_sscal VV_CBLAS(sscal)
#define cblas_dscal VV_CBLAS(dscal)
#define cblas_sdot VV_CBLAS(sdot)
#define cblas_ddot VV_CBLAS(ddot)
#define cblas_sasum VV_CBLAS(sasum)
#define cblas_dasum VV_CBLAS(dasum)
#define cblas_scopy VV_CBLAS(scopy)
#define cblas_dcopy VV_CBLAS(dcopy)
void cblas_sgemm(const enum CBLAS_ORDER, const enum CBLAS_TRANSPOSE, const enum CBLAS_TRANSPOSE, const int, const int, const int, const float, const float*, const int, const float*, const int, const float, float*, const int);
void cblas_dgemm(const enum CBLAS_ORDER, const enum CBLAS_TRANSPOSE, const enum CBLAS_TRANSPOSE, const int, const int, const int, const double, const double*, const int, const double*, const int, const double, double*, const int);
void cblas_sgemv(const enum CBLAS_ORDER, const enum CBLAS_TRANSPOSE, const int, const int, const float, const float*, const int, const float*, const int, const float, float*, const int);
void cblas_dgemv(const enum CBLAS_ORDER, const enum CBLAS_TRANSPOSE, const int, const int, const double, const double*, const int, const double*, const int, const double, double*, const int);
void cblas_saxpy(const int, const float, const float*, const int, float*, const int);
void cblas_daxpy(const int, const double, const double*, const int, double*, const int);
void cblas_sscal(const int, const float, float*, const int);
void cblas_dscal(const int, const double, double*, const int);
float cblas_sdot(const int, const float*, const int, const float*, const int);
double cblas_ddot(const int, const double*, const int, const double*, const int);
float cblas_sasum(const int, const float*, const int);
double cblas_dasum(const int, const double*, const int);
void cblas_scopy(const int, const float*, const int, float*, const int);
void cblas_dcopy(const int, const double*, const int, double*, const int);
