#include "cheb_matrix.h"

#include <cmath>

namespace sb200 {

std::vector<double> cgl_diff_matrix(int P) {
  const int n = P - 1;
  std::vector<double> D((size_t)P * P, 0.0);
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
      D[(size_t)i * P + j] = (double)v;
    }
  }
  return D;
}

}  // namespace sb200
