/* Minimal stand-in for <fftw3.h>: ONLY what /root/reference/chebyshev.c calls, so that the reference's own source file
 * compiles unmodified here (FFTW is not installed in this image).  The two r2r kinds are implemented straight from their
 * definitions in the FFTW manual ("1d Real-even DFTs (DCTs)", "1d Real-odd DFTs (DSTs)"), O(n^2) per line, long double
 * accumulation:
 *   REDFT00:  Y_k = X_0 + (-1)^k X_{n-1} + 2 sum_{j=1}^{n-2} X_j cos(pi j k / (n-1))
 *   RODFT00:  Y_k = 2 sum_{j=0}^{n-1} X_j sin(pi (j+1)(k+1) / (n+1))
 * Test infrastructure (oracle/), never part of the product. */
#ifndef SB200_STUB_FFTW3_H
#define SB200_STUB_FFTW3_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct { int n, is, os; } fftw_iodim;
typedef enum { FFTW_REDFT00 = 3, FFTW_RODFT00 = 7 } fftw_r2r_kind;
typedef struct sb200_stub_plan* fftw_plan;
#define FFTW_MEASURE (0U)
#define FFTW_DESTROY_INPUT (1U << 0)
#define FFTW_PRESERVE_INPUT (1U << 4)
#define FFTW_ESTIMATE (1U << 6)
void* fftw_malloc(size_t n);
void fftw_free(void* p);
fftw_plan fftw_plan_r2r_1d(int n, double* in, double* out, fftw_r2r_kind kind, unsigned flags);
fftw_plan fftw_plan_guru_r2r(int rank, const fftw_iodim* dims, int howmany_rank, const fftw_iodim* howmany_dims, double* in, double* out,
                             const fftw_r2r_kind* kind, unsigned flags);
void fftw_execute_r2r(const fftw_plan p, double* in, double* out);
void fftw_destroy_plan(fftw_plan p);
int fftw_import_system_wisdom(void);
#ifdef __cplusplus
}
#endif
#endif
