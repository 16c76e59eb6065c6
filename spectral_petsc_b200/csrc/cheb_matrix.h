// Host-side construction of the Chebyshev-Gauss-Lobatto differentiation matrix.
#pragma once
#include <vector>

namespace sb200 {
// Row-major P x P matrix D with (D u)_i = u'(x_i), x_i = cos(i*pi/(P-1)), i = 0..P-1 (node 0 is
// x = +1, as in chebyshev.c:154).  This is, in exact arithmetic, the operator that ChebMult
// (chebyshev.c:142-199) applies through DCT-I -> *k -> DST-I -> 1/(2n sin); entries are formed in
// 80-bit long double with the product-to-sum form of x_i - x_j and rounded once to fp64.
std::vector<double> cgl_diff_matrix(int P);
// Even-odd halves for even P (h = P/2, row-major h x h), formed in long double and rounded once:
//   Ae[i][j] = (D[i][j] + D[i][P-1-j])/2,  Bo[i][j] = (D[i][j] - D[i][P-1-j])/2   (see chain.cuh).
void cgl_even_odd(int P, std::vector<double>& Ae, std::vector<double>& Bo);
}  // namespace sb200
