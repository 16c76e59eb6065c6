#include "cheb_matrix.h"

#include <cmath>

namespace sb200 {

static std::vector<long double> cgl_diff_matrix_ld(int P) {
  const int n = P - 1;
  std::vector<long double> D((size_t)P * P, 0.0L);
  if (n < 1) return D;
  const long double pi = 3.14159265358979323846264338327950288L;
  auto cbar = [n](int i) { return (i == 0 || i == n) ? 2.0L : 1.0L; };
  for (int i = 0; i <= n; i++) {
    for (int j = 0; j <= n; j++) {
      long double v;
      if (i != j) {
        // x_i - x_j = cos(a) - cos(b) = -2 sin((a+b)/2) sin((a-b)/2)
        long double sp = sinl(pi * (long double)(i + j) / (2.0L * n));
        long double sm = sinl(pi * (long double)(i - j) / (2.0L * n));
        long double dx = -2.0L * sp * sm;
        long double sgn = ((i + j) & 1) ? -1.0L : 1.0L;
        v = (cbar(i) / cbar(j)) * sgn / dx;
      } else if (i == 0) {
        v = (2.0L * n * n + 1.0L) / 6.0L;
      } else if (i == n) {
        v = -(2.0L * n * n + 1.0L) / 6.0L;
      } else {
        long double s = sinl(pi * (long double)i / n);
        long double c = cosl(pi * (long double)i / n);
        v = -c / (2.0L * s * s);
      }
      D[(size_t)i * P + j] = v;
    }
  }
  return D;
}

std::vector<double> cgl_diff_matrix(int P) {
  std::vector<long double> L = cgl_diff_matrix_ld(P);
  std::vector<double> D(L.size());
  for (size_t k = 0; k < L.size(); k++) D[k] = (double)L[k];
  return D;
}

void cgl_even_odd(int P, std::vector<double>& Ae, std::vector<double>& Bo) {
  const int n = P - 1, h = P / 2;
  std::vector<long double> L = cgl_diff_matrix_ld(P);
  Ae.assign((size_t)h * h, 0.0);
  Bo.assign((size_t)h * h, 0.0);
  for (int i = 0; i < h; i++)
    for (int j = 0; j < h; j++) {
      const long double p = L[(size_t)i * P + j], q = L[(size_t)i * P + (n - j)];
      Ae[(size_t)i * h + j] = (double)(0.5L * (p + q));
      Bo[(size_t)i * h + j] = (double)(0.5L * (p - q));
    }
}

}  // namespace sb200
