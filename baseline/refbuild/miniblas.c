/* miniblas: the handful of CBLAS entry points the reference's CPU paths need (see cblas.h).
 * Straightforward loops with an i-k-j order so gcc -O3 vectorises the inner loop; correctness first.
 * Built as libopenblas.so only because the reference's mshadow.mk links -lopenblas for USE_BLAS=openblas. */
#include "cblas.h"
#include <stdlib.h>
#include <string.h>

#define GEMM_IMPL(NAME, T)                                                                                   \
  static void NAME##_rm(int ta, int tb, int M, int N, int K, T alpha, const T *A, int lda, const T *B,        \
                        int ldb, T beta, T *C, int ldc) {                                                     \
    for (int i = 0; i < M; ++i) {                                                                             \
      T *c = C + (size_t)i * ldc;                                                                             \
      if (beta == (T)0) { for (int j = 0; j < N; ++j) c[j] = 0; }                                             \
      else if (beta != (T)1) { for (int j = 0; j < N; ++j) c[j] *= beta; }                                    \
      for (int k = 0; k < K; ++k) {                                                                           \
        T a = alpha * (ta ? A[(size_t)k * lda + i] : A[(size_t)i * lda + k]);                                 \
        if (!tb) { const T *b = B + (size_t)k * ldb; for (int j = 0; j < N; ++j) c[j] += a * b[j]; }          \
        else { for (int j = 0; j < N; ++j) c[j] += a * B[(size_t)j * ldb + k]; }                              \
      }                                                                                                       \
    }                                                                                                         \
  }

GEMM_IMPL(sgemm, float)
GEMM_IMPL(dgemm, double)

void cblas_sgemm(const enum CBLAS_ORDER order, const enum CBLAS_TRANSPOSE ta, const enum CBLAS_TRANSPOSE tb,
                 const int M, const int N, const int K, const float alpha, const float *A, const int lda,
                 const float *B, const int ldb, const float beta, float *C, const int ldc) {
  if (order == CblasRowMajor) sgemm_rm(ta != CblasNoTrans, tb != CblasNoTrans, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
  else sgemm_rm(tb != CblasNoTrans, ta != CblasNoTrans, N, M, K, alpha, B, ldb, A, lda, beta, C, ldc); /* C^T = B^T A^T */
}
void cblas_dgemm(const enum CBLAS_ORDER order, const enum CBLAS_TRANSPOSE ta, const enum CBLAS_TRANSPOSE tb,
                 const int M, const int N, const int K, const double alpha, const double *A, const int lda,
                 const double *B, const int ldb, const double beta, double *C, const int ldc) {
  if (order == CblasRowMajor) dgemm_rm(ta != CblasNoTrans, tb != CblasNoTrans, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
  else dgemm_rm(tb != CblasNoTrans, ta != CblasNoTrans, N, M, K, alpha, B, ldb, A, lda, beta, C, ldc);
}

#define GEMV_IMPL(NAME, T)                                                                                   \
  void NAME(const enum CBLAS_ORDER order, const enum CBLAS_TRANSPOSE ta, const int M, const int N,            \
            const T alpha, const T *A, const int lda, const T *X, const int incX, const T beta, T *Y,         \
            const int incY) {                                                                                 \
    int trans = (ta != CblasNoTrans);                                                                         \
    if (order == CblasColMajor) trans = !trans;                                                               \
    int rows = (order == CblasRowMajor) ? M : N, cols = (order == CblasRowMajor) ? N : M;                     \
    /* A is rows x cols row-major with leading dim lda; y = alpha*op(A)*x + beta*y */                         \
    int ylen = trans ? cols : rows, xlen = trans ? rows : cols;                                               \
    for (int i = 0; i < ylen; ++i) Y[(size_t)i * incY] = (beta == (T)0) ? (T)0 : beta * Y[(size_t)i * incY];  \
    if (!trans) {                                                                                             \
      for (int i = 0; i < rows; ++i) { T s = 0; for (int j = 0; j < cols; ++j) s += A[(size_t)i * lda + j] * X[(size_t)j * incX]; \
        Y[(size_t)i * incY] += alpha * s; }                                                                   \
    } else {                                                                                                  \
      for (int i = 0; i < rows; ++i) { T x = alpha * X[(size_t)i * incX];                                     \
        for (int j = 0; j < cols; ++j) Y[(size_t)j * incY] += x * A[(size_t)i * lda + j]; }                   \
    }                                                                                                         \
    (void)xlen;                                                                                               \
  }
GEMV_IMPL(cblas_sgemv, float)
GEMV_IMPL(cblas_dgemv, double)

#define GER_IMPL(NAME, T)                                                                                    \
  void NAME(const enum CBLAS_ORDER order, const int M, const int N, const T alpha, const T *X, const int incX, \
            const T *Y, const int incY, T *A, const int lda) {                                                \
    for (int i = 0; i < M; ++i) for (int j = 0; j < N; ++j) {                                                 \
      T v = alpha * X[(size_t)i * incX] * Y[(size_t)j * incY];                                                \
      if (order == CblasRowMajor) A[(size_t)i * lda + j] += v; else A[(size_t)j * lda + i] += v;              \
    }                                                                                                         \
  }
GER_IMPL(cblas_sger, float)
GER_IMPL(cblas_dger, double)

float cblas_sdot(const int N, const float *X, const int incX, const float *Y, const int incY) {
  float s = 0; for (int i = 0; i < N; ++i) s += X[(size_t)i * incX] * Y[(size_t)i * incY]; return s;
}
double cblas_ddot(const int N, const double *X, const int incX, const double *Y, const int incY) {
  double s = 0; for (int i = 0; i < N; ++i) s += X[(size_t)i * incX] * Y[(size_t)i * incY]; return s;
}

/* Level-3 triangular / rank-k routines (row-major only: the reference always passes CblasRowMajor).
 * op(A)(i,j) is read through TA(); "lower" below means op(A) is lower triangular. */
#define L3_IMPL(P, T)                                                                                        \
  void cblas_##P##trsm(const enum CBLAS_ORDER order, const enum CBLAS_SIDE side, const enum CBLAS_UPLO uplo,  \
                       const enum CBLAS_TRANSPOSE ta, const enum CBLAS_DIAG diag, const int M, const int N,   \
                       const T alpha, const T *A, const int lda, T *B, const int ldb) {                       \
    if (order != CblasRowMajor) abort();                                                                      \
    int tr = (ta != CblasNoTrans), lower = ((uplo == CblasLower) != tr), unit = (diag == CblasUnit);          \
    for (int i = 0; i < M; ++i) for (int j = 0; j < N; ++j) B[(size_t)i * ldb + j] *= alpha;                  \
    if (side == CblasLeft) { /* op(A) X = B, A is MxM; column by column of B */                               \
      for (int c = 0; c < N; ++c) {                                                                           \
        if (lower) for (int i = 0; i < M; ++i) { T s = B[(size_t)i * ldb + c];                                \
            for (int k = 0; k < i; ++k) s -= (tr ? A[(size_t)k * lda + i] : A[(size_t)i * lda + k]) * B[(size_t)k * ldb + c]; \
            B[(size_t)i * ldb + c] = unit ? s : s / A[(size_t)i * lda + i]; }                                 \
        else for (int i = M - 1; i >= 0; --i) { T s = B[(size_t)i * ldb + c];                                 \
            for (int k = i + 1; k < M; ++k) s -= (tr ? A[(size_t)k * lda + i] : A[(size_t)i * lda + k]) * B[(size_t)k * ldb + c]; \
            B[(size_t)i * ldb + c] = unit ? s : s / A[(size_t)i * lda + i]; }                                 \
      }                                                                                                       \
    } else { /* X op(A) = B, A is NxN; row by row of B: x_j = (b_j - sum_k x_k a(k,j)) / a(j,j) */            \
      for (int r = 0; r < M; ++r) {                                                                           \
        T *b = B + (size_t)r * ldb;                                                                           \
        if (lower) for (int j = N - 1; j >= 0; --j) { T s = b[j];                                             \
            for (int k = j + 1; k < N; ++k) s -= b[k] * (tr ? A[(size_t)j * lda + k] : A[(size_t)k * lda + j]); \
            b[j] = unit ? s : s / A[(size_t)j * lda + j]; }                                                   \
        else for (int j = 0; j < N; ++j) { T s = b[j];                                                        \
            for (int k = 0; k < j; ++k) s -= b[k] * (tr ? A[(size_t)j * lda + k] : A[(size_t)k * lda + j]);   \
            b[j] = unit ? s : s / A[(size_t)j * lda + j]; }                                                   \
      }                                                                                                       \
    }                                                                                                         \
  }                                                                                                           \
  void cblas_##P##trmm(const enum CBLAS_ORDER order, const enum CBLAS_SIDE side, const enum CBLAS_UPLO uplo,  \
                       const enum CBLAS_TRANSPOSE ta, const enum CBLAS_DIAG diag, const int M, const int N,   \
                       const T alpha, const T *A, const int lda, T *B, const int ldb) {                       \
    if (order != CblasRowMajor) abort();                                                                      \
    int tr = (ta != CblasNoTrans), lower = ((uplo == CblasLower) != tr), unit = (diag == CblasUnit);          \
    int n = (side == CblasLeft) ? M : N;                                                                      \
    T *tmp = (T *)malloc(sizeof(T) * (size_t)M * N);                                                          \
    for (int i = 0; i < M; ++i) for (int j = 0; j < N; ++j) {                                                 \
      T s = 0;                                                                                                \
      if (side == CblasLeft) { /* sum_k op(A)(i,k) B(k,j) */                                                  \
        int k0 = lower ? 0 : i, k1 = lower ? i : n - 1;                                                       \
        for (int k = k0; k <= k1; ++k) { T a = (k == i && unit) ? (T)1 : (tr ? A[(size_t)k * lda + i] : A[(size_t)i * lda + k]); \
          s += a * B[(size_t)k * ldb + j]; }                                                                  \
      } else { /* sum_k B(i,k) op(A)(k,j) */                                                                  \
        int k0 = lower ? j : 0, k1 = lower ? n - 1 : j;                                                       \
        for (int k = k0; k <= k1; ++k) { T a = (k == j && unit) ? (T)1 : (tr ? A[(size_t)j * lda + k] : A[(size_t)k * lda + j]); \
          s += B[(size_t)i * ldb + k] * a; }                                                                  \
      }                                                                                                       \
      tmp[(size_t)i * N + j] = alpha * s;                                                                     \
    }                                                                                                         \
    for (int i = 0; i < M; ++i) memcpy(B + (size_t)i * ldb, tmp + (size_t)i * N, sizeof(T) * N);              \
    free(tmp);                                                                                                \
  }                                                                                                           \
  void cblas_##P##syrk(const enum CBLAS_ORDER order, const enum CBLAS_UPLO uplo, const enum CBLAS_TRANSPOSE trans, \
                       const int N, const int K, const T alpha, const T *A, const int lda, const T beta, T *C, \
                       const int ldc) {                                                                       \
    if (order != CblasRowMajor) abort();                                                                      \
    int tr = (trans != CblasNoTrans);                                                                         \
    for (int i = 0; i < N; ++i) {                                                                             \
      int j0 = (uplo == CblasLower) ? 0 : i, j1 = (uplo == CblasLower) ? i : N - 1;                           \
      for (int j = j0; j <= j1; ++j) { T s = 0;                                                               \
        for (int k = 0; k < K; ++k) s += tr ? A[(size_t)k * lda + i] * A[(size_t)k * lda + j]                 \
                                            : A[(size_t)i * lda + k] * A[(size_t)j * lda + k];                \
        C[(size_t)i * ldc + j] = alpha * s + (beta == (T)0 ? (T)0 : beta * C[(size_t)i * ldc + j]); }         \
    }                                                                                                         \
  }
L3_IMPL(s, float)
L3_IMPL(d, double)
