/* Minimal CBLAS interface used as the CPU BLAS provider when building the UNMODIFIED reference
 * (MXNet 1.4 / GeoMX) offline: no OpenBLAS/ATLAS/MKL headers exist in this image.
 * Only the entry points the reference's CPU code paths call are provided (baseline/refbuild/miniblas.c).
 * None of this is on the GPU path that bench.py --impl reference measures (that is cuBLAS + NCCL). */
#ifndef GEOMX_BASELINE_CBLAS_H_
#define GEOMX_BASELINE_CBLAS_H_
#ifdef __cplusplus
extern "C" {
#endif
enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 };
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 };
enum CBLAS_UPLO { CblasUpper = 121, CblasLower = 122 };
enum CBLAS_DIAG { CblasNonUnit = 131, CblasUnit = 132 };
enum CBLAS_SIDE { CblasLeft = 141, CblasRight = 142 };
typedef enum CBLAS_ORDER CBLAS_LAYOUT;

void cblas_sgemm(const enum CBLAS_ORDER order, const enum CBLAS_TRANSPOSE ta, const enum CBLAS_TRANSPOSE tb,
                 const int M, const int N, const int K, const float alpha, const float *A, const int lda,
                 const float *B, const int ldb, const float beta, float *C, const int ldc);
void cblas_dgemm(const enum CBLAS_ORDER order, const enum CBLAS_TRANSPOSE ta, const enum CBLAS_TRANSPOSE tb,
                 const int M, const int N, const int K, const double alpha, const double *A, const int lda,
                 const double *B, const int ldb, const double beta, double *C, const int ldc);
void cblas_sgemv(const enum CBLAS_ORDER order, const enum CBLAS_TRANSPOSE ta, const int M, const int N,
                 const float alpha, const float *A, const int lda, const float *X, const int incX,
                 const float beta, float *Y, const int incY);
void cblas_dgemv(const enum CBLAS_ORDER order, const enum CBLAS_TRANSPOSE ta, const int M, const int N,
                 const double alpha, const double *A, const int lda, const double *X, const int incX,
                 const double beta, double *Y, const int incY);
void cblas_sger(const enum CBLAS_ORDER order, const int M, const int N, const float alpha, const float *X,
                const int incX, const float *Y, const int incY, float *A, const int lda);
void cblas_dger(const enum CBLAS_ORDER order, const int M, const int N, const double alpha, const double *X,
                const int incX, const double *Y, const int incY, double *A, const int lda);
float cblas_sdot(const int N, const float *X, const int incX, const float *Y, const int incY);
double cblas_ddot(const int N, const double *X, const int incX, const double *Y, const int incY);

#define GX_DECL_L3(P, T) \
void cblas_##P##trsm(const enum CBLAS_ORDER order, const enum CBLAS_SIDE side, const enum CBLAS_UPLO uplo, \
                     const enum CBLAS_TRANSPOSE ta, const enum CBLAS_DIAG diag, const int M, const int N, \
                     const T alpha, const T *A, const int lda, T *B, const int ldb); \
void cblas_##P##trmm(const enum CBLAS_ORDER order, const enum CBLAS_SIDE side, const enum CBLAS_UPLO uplo, \
                     const enum CBLAS_TRANSPOSE ta, const enum CBLAS_DIAG diag, const int M, const int N, \
                     const T alpha, const T *A, const int lda, T *B, const int ldb); \
void cblas_##P##syrk(const enum CBLAS_ORDER order, const enum CBLAS_UPLO uplo, const enum CBLAS_TRANSPOSE trans, \
                     const int N, const int K, const T alpha, const T *A, const int lda, const T beta, T *C, \
                     const int ldc);
GX_DECL_L3(s, float)
GX_DECL_L3(d, double)
#undef GX_DECL_L3
#ifdef __cplusplus
}
#endif
#endif
