// Minimal stand-in for <mkl.h>: forwards the two CBLAS entry points compute_groundtruth.cpp uses
// (cblas_sgemm :92-100,118 and cblas_sdot :73) to the LP64 Fortran BLAS symbols that
// libtorch_cpu.so exports from its statically linked oneMKL. TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cstddef>
extern "C" {
void sgemm_(const char *ta, const char *tb, const int *m, const int *n, const int *k, const float *alpha,
            const float *a, const int *lda, const float *b, const int *ldb, const float *beta, float *c,
            const int *ldc);
float sdot_(const int *n, const float *x, const int *incx, const float *y, const int *incy);
}
#define MKL_INT int
enum CBLAS_LAYOUT { CblasRowMajor = 101, CblasColMajor = 102 };
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 };
static inline void cblas_sgemm(CBLAS_LAYOUT layout, CBLAS_TRANSPOSE ta, CBLAS_TRANSPOSE tb, size_t m, size_t n,
                               size_t k, float alpha, const float *a, size_t lda, const float *b, size_t ldb,
                               float beta, float *c, size_t ldc) {
    (void)layout;  // only CblasColMajor is used by the reference
    const char cta = (ta == CblasNoTrans) ? 'N' : 'T', ctb = (tb == CblasNoTrans) ? 'N' : 'T';
    int im = (int)m, in = (int)n, ik = (int)k, ilda = (int)lda, ildb = (int)ldb, ildc = (int)ldc;
    sgemm_(&cta, &ctb, &im, &in, &ik, &alpha, a, &ilda, b, &ildb, &beta, c, &ildc);
}
static inline float cblas_sdot(size_t n, const float *x, size_t incx, const float *y, size_t incy) {
    int in = (int)n, ix = (int)incx, iy = (int)incy;
    return sdot_(&in, x, &ix, y, &iy);
}
