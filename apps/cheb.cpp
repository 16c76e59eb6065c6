// ./cheb - the reference's test program of the differentiation code (cheb.c) on the B200 path:
//
//     apps/cheb [-m1 5] [-m 8 -n 7 -p 1 -d 0]
//
// prints, like cheb.c:95-112, the max-norm error of  d/dx exp(x) = exp(x)  on m1 Chebyshev-Gauss-Lobatto nodes through the 1-D
// operator MatCreateChebD1 (cheb.c:47,68-70), then that of the derivative along axis -d of  exp(x) + exp(y) + exp(z)  on an
// (m, n, p) grid through MatCreateCheb (cheb.c:58,77-91).  The operators run on the GPU through the reference's own names
// (include/sb200_reference_api.h); the fields are filled and the norms taken on the host, as the reference does.
#include <algorithm>

#include "common.h"

int main(int argc, char** argv) {
  app::Options o;
  if (int rc = o.parse(argc, argv)) return rc;
  const int m1 = o.integer("m1", 5), m = o.integer("m", 8), n = o.integer("n", 7), p = o.integer("p", 1), d = o.integer("d", 0);  // cheb.c:27
  if (m1 < 1 || m < 1 || n < 1 || p < 1) {
    fprintf(stderr, "error: extents must be positive\n");
    return 83;
  }
  const double PI = 3.14159265358979323846;
  auto node = [PI](int i, int ext) { return ext == 1 ? 0.0 : cos(i * PI / (ext - 1)); };  // cheb.c:79-83

  {  // 1-D: u_i = exp(cos(i pi / (m1 - 1))), D u = u
    Vec u, b;
    Mat A;
    CHK(VecCreateSeqCUDA(PETSC_COMM_WORLD, m1, &u));
    CHK(VecDuplicate(u, &b));
    CHK(MatCreateChebD1(PETSC_COMM_WORLD, u, b, FFTW_ESTIMATE, &A));
    std::vector<double> a(m1), r(m1);
    for (int i = 0; i < m1; i++) a[i] = exp(cos(i * PI / (m1 - 1)));
    CHK(VecSetValuesHost(u, a.data()));
    CHK(MatMult(A, u, b));
    CHK(VecGetValuesHost(b, r.data()));
    double norm = 0;
    for (int i = 0; i < m1; i++) norm = fmax(norm, fabs(r[i] - a[i]));
    printf("Norm of error %g\n", norm);
    CHK(MatDestroy(A));
    CHK(VecDestroy(u));
    CHK(VecDestroy(b));
  }
  {  // 3-D: u = exp(x) + exp(y) + exp(z), derivative along axis d
    const int N = m * n * p;
    int dims[3] = {m, n, p};
    Vec u2, b2;
    Mat A2;
    CHK(VecCreateSeqCUDA(PETSC_COMM_WORLD, N, &u2));
    CHK(VecDuplicate(u2, &b2));
    CHK(MatCreateCheb(PETSC_COMM_WORLD, 3, d, dims, FFTW_ESTIMATE, u2, b2, &A2));
    std::vector<double> a(N), e(N, 0.0), r(N);
    for (int i = 0; i < m; i++)
      for (int j = 0; j < n; j++)
        for (int k = 0; k < p; k++) {
          const double x = node(i, m), y = node(j, n), z = node(k, p);
          a[(i * n + j) * p + k] = exp(x) + exp(y) + exp(z);
          e[(i * n + j) * p + k] = d == 0 ? exp(x) : (d == 1 ? exp(y) : exp(z));
        }
    CHK(VecSetValuesHost(u2, a.data()));
    CHK(MatMult(A2, u2, b2));
    CHK(VecGetValuesHost(b2, r.data()));
    double norm = 0;
    for (int i = 0; i < N; i++) norm = fmax(norm, fabs(r[i] - e[i]));
    printf("Norm of error %g\n", norm);
    CHK(MatDestroy(A2));
    CHK(VecDestroy(u2));
    CHK(VecDestroy(b2));
  }
  o.warn_unused();
  return 0;
}
