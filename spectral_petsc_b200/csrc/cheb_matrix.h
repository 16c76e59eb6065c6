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
// The same for any P >= 2 (odd P: the middle node is its own pair), zero padded to HP x HP row-major, HP >= ceil(P/2):
// the operands of the generalised even-odd kernel (deriv_eo.cu).  For P % 16 == 0 and HP = P/2 identical to cgl_even_odd.
void cgl_even_odd_padded(int P, int HP, std::vector<double>& Ae, std::vector<double>& Bo);
}  // namespace sb200
