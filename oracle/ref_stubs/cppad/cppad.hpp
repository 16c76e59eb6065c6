// Minimal stand-in for <cppad/cppad.hpp>: stokes.C hard-codes "#define WITH_CPPAD 1", so the automatic-differentiation
// preconditioner (-pcvel 3, StokesPCSetUp3 / StokesComputeNodalJacobian) has to COMPILE; it is never called through the
// stand-in (README:57-60 reports that variant as no better and buggy; out of scope).  Values propagate, derivatives do not:
// ADFun::Jacobian returns zeros.  Test infrastructure.
#ifndef SB200_STUB_CPPAD_HPP
#define SB200_STUB_CPPAD_HPP
#include <vector>
namespace CppAD {
using std::vector;
template <class B>
struct AD {
  B v;
  AD() : v(0) {}
  AD(B x) : v(x) {}
  AD(int x) : v((B)x) {}
  AD& operator+=(const AD& o) { v += o.v; return *this; }
  AD& operator-=(const AD& o) { v -= o.v; return *this; }
  AD& operator*=(const AD& o) { v *= o.v; return *this; }
};
template <class B> AD<B> operator+(const AD<B>& a, const AD<B>& b) { return AD<B>(a.v + b.v); }
template <class B> AD<B> operator-(const AD<B>& a, const AD<B>& b) { return AD<B>(a.v - b.v); }
template <class B> AD<B> operator*(const AD<B>& a, const AD<B>& b) { return AD<B>(a.v * b.v); }
template <class B> AD<B> operator/(const AD<B>& a, const AD<B>& b) { return AD<B>(a.v / b.v); }
template <class B> AD<B> operator*(B a, const AD<B>& b) { return AD<B>(a * b.v); }
template <class B> AD<B> operator*(const AD<B>& a, B b) { return AD<B>(a.v * b); }
template <class B> AD<B> operator*(int a, const AD<B>& b) { return AD<B>(a * b.v); }
template <class B> AD<B> operator+(B a, const AD<B>& b) { return AD<B>(a + b.v); }
template <class B> AD<B> operator-(B a, const AD<B>& b) { return AD<B>(a - b.v); }
template <class V> void Independent(V&) {}
template <class B>
struct ADFun {
  size_t n, m;
  template <class V1, class V2> ADFun(const V1& x, const V2& y) : n(x.size()), m(y.size()) {}
  vector<B> Jacobian(const vector<B>&) { return vector<B>(n * m, B(0)); }
};
}  // namespace CppAD
#endif
