// TEST DOUBLE - never built into libspectral_b200.so, never shipped, never used by bench.py or the package.
//
// A plain-CPU stand-in for the DEVICE-side entry points of include/spectral_b200.h (memory helpers, the elliptic shells, the
// FGMRES), so that the HOST layer of the product - host/reference_api.cpp, host/petsc_shim.cpp, host/host_ilu.cpp,
// csrc/exact.cpp, csrc/fd_rows.h and the native executable apps/elliptic.cpp - can be linked UNCHANGED against it and driven
// end to end in the CPU test suite (tests/test_native_cpu_double.py): option handling, Newton loop, PC refresh, printed lines,
// iteration counts.  "Device" pointers are host pointers here.  The arithmetic follows the same definitions as the oracle
// (dense CGL differentiation matrix per axis, reference operation order), which is all a host-logic test needs; GPU parity
// is tested on the GPU (tests/test_gpu_*.py), never through this file.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/spectral_b200.h"
#include "../../spectral_petsc_b200/csrc/cheb_matrix.h"
#include "../../spectral_petsc_b200/csrc/fd_rows.h"

namespace sb200 {
static thread_local std::string g_err;
void set_last_error(const std::string& msg) { g_err = msg; }
}  // namespace sb200
using namespace sb200;

#define FAIL(code, msg)    \
  do {                     \
    set_last_error(msg);   \
    return (code);         \
  } while (0)

struct sb200_cheb {
  int P;
  long long O, R, N;
  std::vector<double> D;
};

struct sb200_elliptic {
  int d;
  std::vector<int> dim;
  long long m, g;
  std::vector<std::vector<double>> D;      // per axis
  std::vector<long long> stride, ixG, ixD; // local index of interior / boundary nodes in walk order
  std::vector<double> eta, deta, dirichlet, b;
  std::vector<std::vector<double>> gradu;
  double gamma = 0.0, exponent = 2.0;
};

struct sb200_ksp {
  long long n;
  int restart;
  sb200_apply_fn op = nullptr, pc = nullptr;
  void *op_ctx = nullptr, *pc_ctx = nullptr;
  double rtol = 1e-5, atol = 1e-50, dtol = 1e5, rnorm = 0, bnorm = 0;
  int maxits = 10000, its = 0, reason = 0;
  std::vector<double> history;
};

namespace {

// y = d x / d xi_axis on the full grid: dense matrix along one axis
void deriv(const sb200_elliptic* e, int axis, const double* x, double* y) {
  const int P = e->dim[axis];
  const long long R = e->stride[axis], O = e->m / (R * P);
  const double* D = e->D[axis].data();
  for (long long o = 0; o < O; o++)
    for (long long r = 0; r < R; r++) {
      const double* xl = x + o * P * R + r;
      double* yl = y + o * P * R + r;
      for (int i = 0; i < P; i++) {
        double s = 0;
        for (int j = 0; j < P; j++) s += D[(size_t)i * P + j] * xl[(long long)j * R];
        yl[(long long)i * R] = s;
      }
    }
}

void pad(const sb200_elliptic* e, const double* U, bool with_dirichlet, double* w0) {
  for (long long q = 0; q < e->g; q++) w0[e->ixG[q]] = U[q];
  for (size_t q = 0; q < e->ixD.size(); q++) w0[e->ixD[q]] = with_dirichlet ? e->dirichlet[q] : 0.0;
}

}  // namespace

extern "C" {

const char* sb200_last_error(void) { return g_err.c_str(); }
int sb200_malloc(void** p, size_t bytes) {
  *p = std::calloc(bytes ? bytes : 8, 1);
  return *p ? 0 : SB200_ERR_CUDA;
}
int sb200_free(void* p) {
  std::free(p);
  return 0;
}
int sb200_memcpy_h2d(void* d, const void* s, size_t n, void*) { std::memcpy(d, s, n); return 0; }
int sb200_memcpy_d2h(void* d, const void* s, size_t n, void*) { std::memcpy(d, s, n); return 0; }
int sb200_memcpy_d2d(void* d, const void* s, size_t n, void*) { std::memcpy(d, s, n); return 0; }
int sb200_memset0(void* d, size_t n, void*) { std::memset(d, 0, n); return 0; }
int sb200_stream_sync(void*) { return 0; }

// ---- ChebMult -------------------------------------------------------------------------------------------------------------
int sb200_cheb_create(int rank, int tr, const int* dims, long long n_total, sb200_cheb** out) {
  *out = nullptr;
  if (n_total < 2) FAIL(SB200_ERR_USER, "n must be >= 2");
  if (tr < 0 || tr >= rank) FAIL(SB200_ERR_USER, "tdim out of range");
  long long s = 1;
  for (int r = 0; r < rank; r++) s *= dims[r];
  if (s != n_total) FAIL(SB200_ERR_USER, "dimensions do not agree");
  sb200_cheb* c = new sb200_cheb();
  c->P = dims[tr];
  c->N = n_total;
  c->R = 1;
  for (int r = tr + 1; r < rank; r++) c->R *= dims[r];
  c->O = n_total / (c->R * c->P);
  c->D = cgl_diff_matrix(c->P);
  *out = c;
  return 0;
}
int sb200_cheb_apply(sb200_cheb* c, const double* x, double* y, void*) {
  if (x == y) FAIL(SB200_ERR_ARG, "x and y must not alias");
  for (long long o = 0; o < c->O; o++)
    for (long long r = 0; r < c->R; r++)
      for (int i = 0; i < c->P; i++) {
        double s = 0;
        for (int j = 0; j < c->P; j++) s += c->D[(size_t)i * c->P + j] * x[o * c->P * c->R + (long long)j * c->R + r];
        y[o * c->P * c->R + (long long)i * c->R + r] = s;
      }
  return 0;
}
int sb200_cheb_destroy(sb200_cheb* c) {
  delete c;
  return 0;
}

// ---- elliptic shells (elliptic.C:297-339, 481-533, 537-590) ---------------------------------------------------------------
int sb200_elliptic_create(int d, const int* dim, sb200_elliptic** out) {
  *out = nullptr;
  if (d < 1 || d > 10) FAIL(SB200_ERR_USER, "dimension count must be in [1,10]");
  for (int j = 0; j < d; j++)
    if (dim[j] < 3) FAIL(SB200_ERR_USER, "each extent must be >= 3");
  sb200_elliptic* e = new sb200_elliptic();
  e->d = d;
  e->dim.assign(dim, dim + d);
  e->stride.assign(d, 1);
  e->m = 1;
  for (int j = d - 1; j >= 0; j--) {
    e->stride[j] = e->m;
    e->m *= dim[j];
  }
  std::vector<int> ind(d, 0);
  for (long long node = 0; node < e->m; node++) {  // SetupBC walk (elliptic.C:386-415)
    bool bdy = false;
    for (int j = 0; j < d; j++) bdy = bdy || ind[j] == 0 || ind[j] == dim[j] - 1;
    (bdy ? e->ixD : e->ixG).push_back(node);
    for (int j = d - 1; j >= 0; j--) {
      if (++ind[j] < dim[j]) break;
      ind[j] = 0;
    }
  }
  e->g = (long long)e->ixG.size();
  for (int j = 0; j < d; j++) e->D.push_back(cgl_diff_matrix(dim[j]));
  e->eta.assign(e->m, 1.0);
  e->deta.assign(e->m, 0.0);
  e->gradu.assign(d, std::vector<double>(e->m, 0.0));
  e->dirichlet.assign(e->ixD.size(), 0.0);
  e->b.assign(e->g, 0.0);
  *out = e;
  return 0;
}
int sb200_elliptic_sizes(const sb200_elliptic* e, long long* m, long long* g, long long* nd) {
  if (m) *m = e->m;
  if (g) *g = e->g;
  if (nd) *nd = e->m - e->g;
  return 0;
}
int sb200_elliptic_set_params(sb200_elliptic* e, double gamma, double exponent) {
  e->gamma = gamma;
  e->exponent = exponent;
  return 0;
}
int sb200_elliptic_set_dirichlet(sb200_elliptic* e, const double* v, void*) {
  e->dirichlet.assign(v, v + e->ixD.size());
  return 0;
}
int sb200_elliptic_set_rhs(sb200_elliptic* e, const double* b, void*) {
  e->b.assign(b, b + e->g);
  return 0;
}
int sb200_elliptic_matmult(sb200_elliptic* e, const double* U, double* V, void*) {
  if (!U || !V || U == V) FAIL(SB200_ERR_ARG, "MatMult_Elliptic: U and V must be distinct non-null vectors");
  const long long m = e->m;
  std::vector<double> w0(m), out(m, 0.0), t(m);
  std::vector<std::vector<double>> w(e->d, std::vector<double>(m));
  pad(e, U, false, w0.data());
  for (int k = 0; k < e->d; k++) deriv(e, k, w0.data(), w[k].data());
  for (int k = 0; k < e->d; k++)
    for (long long i = 0; i < m; i++) w[k][i] = e->eta[i] * w[k][i] + e->deta[i] * w0[i] * e->gradu[k][i];
  for (int k = 0; k < e->d; k++) {
    deriv(e, k, w[k].data(), t.data());
    for (long long i = 0; i < m; i++) out[i] -= t[i];
  }
  for (long long q = 0; q < e->g; q++) V[q] = out[e->ixG[q]];
  return 0;
}
int sb200_elliptic_function(sb200_elliptic* e, const double* U, double* F, void*) {
  if (!U || !F || U == F) FAIL(SB200_ERR_ARG, "FormFunction: U and F must be distinct non-null vectors");
  const long long m = e->m;
  std::vector<double> w0(m), out(m, 0.0), t(m), f(m);
  pad(e, U, true, w0.data());
  for (int k = 0; k < e->d; k++) deriv(e, k, w0.data(), e->gradu[k].data());
  for (long long i = 0; i < m; i++) {
    e->eta[i] = 1.0 + e->gamma * pow(w0[i], e->exponent);
    e->deta[i] = e->exponent * e->gamma * pow(w0[i], e->exponent - 1.0);
  }
  for (int k = 0; k < e->d; k++) {
    for (long long i = 0; i < m; i++) f[i] = e->eta[i] * e->gradu[k][i];
    deriv(e, k, f.data(), t.data());
    for (long long i = 0; i < m; i++) out[i] -= t[i];
  }
  for (long long q = 0; q < e->g; q++) F[q] = out[e->ixG[q]] - e->b[q];
  return 0;
}
int sb200_elliptic_jacobian_sizes(sb200_elliptic* e, long long* nrows, long long* nnz) {
  FdGrid G;
  fd_grid_init(&G, e->d, e->dim.data());
  if (nrows) *nrows = G.g;
  if (nnz) *nnz = fd_total_entries(G);
  return 0;
}
int sb200_elliptic_jacobian_csr(sb200_elliptic* e, int* rowptr, int* colidx, double* vals, void*) {
  FdGrid G;
  fd_grid_init(&G, e->d, e->dim.data());
  std::vector<double> x;
  for (int j = 0; j < e->d; j++)
    for (int i = 0; i < e->dim[j]; i++) x.push_back(cos(i * M_PI / (e->dim[j] - 1)));
  FdFields F;
  F.xtab = x.data();
  F.eta = e->eta.data();
  F.deta = e->deta.data();
  for (int j = 0; j < SB200_FD_MAX_DIM; j++) F.gradu[j] = j < e->d ? e->gradu[j].data() : nullptr;
  for (long long r = 0; r < G.g; r++) {
    int k[SB200_FD_MAX_DIM];
    long long cols[2 * SB200_FD_MAX_DIM + 1];
    double v[2 * SB200_FD_MAX_DIM + 1];
    const long long node = fd_decode(G, r, k);
    const int n = fd_row(G, F, r, k, node, cols, v);
    const long long o = fd_row_offset(G, k, r);
    if (rowptr) rowptr[r] = (int)o;
    for (int q = 0; q < n; q++) {
      if (colidx) colidx[o + q] = (int)cols[q];
      vals[o + q] = v[q];
    }
  }
  if (rowptr) rowptr[G.g] = (int)fd_total_entries(G);
  return 0;
}
int sb200_elliptic_destroy(sb200_elliptic* e) {
  delete e;
  return 0;
}

// ---- FGMRES(restart): PETSc's algorithm as oracle/fgmres.py and csrc/ksp.cu restate it --------------------------------------
int sb200_ksp_create(long long n, int restart, sb200_ksp** out) {
  sb200_ksp* k = new sb200_ksp();
  k->n = n;
  k->restart = restart;
  *out = k;
  return 0;
}
int sb200_ksp_set_operators(sb200_ksp* k, sb200_apply_fn op, void* op_ctx, sb200_apply_fn pc, void* pc_ctx) {
  k->op = op;
  k->op_ctx = op_ctx;
  k->pc = pc;
  k->pc_ctx = pc_ctx;
  return 0;
}
int sb200_ksp_set_tolerances(sb200_ksp* k, double rtol, double atol, double dtol, int maxits) {
  k->rtol = rtol;
  k->atol = atol;
  k->dtol = dtol;
  k->maxits = maxits;
  return 0;
}
int sb200_ksp_solve(sb200_ksp* K, const double* b, double* x, int guess_nonzero, void* stream) {
  const long long n = K->n;
  const int m = K->restart;
  auto dot = [&](const double* a, const double* c) {
    double s = 0;
    for (long long i = 0; i < n; i++) s += a[i] * c[i];
    return s;
  };
  if (!guess_nonzero) std::memset(x, 0, sizeof(double) * n);
  K->bnorm = sqrt(dot(b, b));
  const double ttol = fmax(K->rtol * K->bnorm, K->atol);
  K->its = 0;
  K->reason = 0;
  K->history.clear();
  std::vector<double> r(n), w(n), V((size_t)(m + 1) * n), Z((size_t)m * n), H((size_t)(m + 1) * m), cs(m), sn(m), g(m + 1), h(m + 1), y(m);
  while (true) {
    if (guess_nonzero || K->its > 0) {
      if (int rc = K->op(K->op_ctx, x, w.data(), stream)) return rc;
      for (long long i = 0; i < n; i++) r[i] = b[i] - w[i];
    } else {
      std::memcpy(r.data(), b, sizeof(double) * n);
    }
    const double beta = sqrt(dot(r.data(), r.data()));
    if (K->its == 0) K->history.push_back(beta);
    K->rnorm = beta;
    if (beta <= ttol) {
      K->reason = 2;
      return 0;
    }
    for (long long i = 0; i < n; i++) V[i] = r[i] / beta;
    std::fill(g.begin(), g.end(), 0.0);
    g[0] = beta;
    int k = 0;
    bool done = false;
    while (k < m && !done) {
      double* zk = &Z[(size_t)k * n];
      const double* vk = &V[(size_t)k * n];
      if (K->pc) {
        if (int rc = K->pc(K->pc_ctx, vk, zk, stream)) return rc;
      } else {
        std::memcpy(zk, vk, sizeof(double) * n);
      }
      if (int rc = K->op(K->op_ctx, zk, w.data(), stream)) return rc;
      for (int i = 0; i <= k; i++) h[i] = dot(&V[(size_t)i * n], w.data());  // classical Gram-Schmidt
      for (int i = 0; i <= k; i++)
        for (long long q = 0; q < n; q++) w[q] -= h[i] * V[(size_t)i * n + q];
      const double hn = sqrt(dot(w.data(), w.data()));
      for (int i = 0; i <= k; i++) H[(size_t)i * m + k] = h[i];
      for (int i = 0; i < k; i++) {
        const double a = H[(size_t)i * m + k], c = H[(size_t)(i + 1) * m + k];
        H[(size_t)i * m + k] = cs[i] * a + sn[i] * c;
        H[(size_t)(i + 1) * m + k] = -sn[i] * a + cs[i] * c;
      }
      const double a = H[(size_t)k * m + k], rr = hypot(a, hn);
      cs[k] = rr > 0 ? a / rr : 1.0;
      sn[k] = rr > 0 ? hn / rr : 0.0;
      H[(size_t)k * m + k] = rr;
      g[k + 1] = -sn[k] * g[k];
      g[k] = cs[k] * g[k];
      for (long long q = 0; q < n; q++) V[(size_t)(k + 1) * n + q] = hn > 0 ? w[q] / hn : 0.0;
      const double rn = fabs(g[k + 1]);
      K->its++;
      k++;
      K->history.push_back(rn);
      K->rnorm = rn;
      if (rn <= ttol) K->reason = 2, done = true;
      else if (rn >= K->dtol * K->bnorm) K->reason = -4, done = true;
      else if (K->its >= K->maxits) K->reason = -3, done = true;
    }
    for (int i = k - 1; i >= 0; i--) {  // back substitution on the rotated Hessenberg
      double s = g[i];
      for (int j = i + 1; j < k; j++) s -= H[(size_t)i * m + j] * y[j];
      y[i] = s / H[(size_t)i * m + i];
    }
    for (int i = 0; i < k; i++)
      for (long long q = 0; q < n; q++) x[q] += y[i] * Z[(size_t)i * n + q];
    if (done) return 0;
  }
}
int sb200_ksp_get_result(const sb200_ksp* k, int* its, double* rnorm, double* bnorm, int* reason) {
  if (its) *its = k->its;
  if (rnorm) *rnorm = k->rnorm;
  if (bnorm) *bnorm = k->bnorm;
  if (reason) *reason = k->reason;
  return 0;
}
int sb200_ksp_destroy(sb200_ksp* k) {
  delete k;
  return 0;
}

// ---- the Stokes shells are not doubled: the host layer links, a call reports "not supported" --------------------------------
#define NOSTOKES(name, ...) \
  int name(__VA_ARGS__) { FAIL(SB200_ERR_SUP, #name ": not provided by the CPU test double"); }
NOSTOKES(sb200_stokes_create, int, const int*, sb200_stokes**)
NOSTOKES(sb200_stokes_destroy, sb200_stokes*)
NOSTOKES(sb200_stokes_sizes, const sb200_stokes*, long long*, long long*, long long*, long long*, long long*)
NOSTOKES(sb200_stokes_set_rheology, sb200_stokes*, int, double, double, double, double)
NOSTOKES(sb200_stokes_set_dirichlet, sb200_stokes*, const double*, void*)
NOSTOKES(sb200_stokes_set_force, sb200_stokes*, const double*, void*)
NOSTOKES(sb200_stokes_matmult, sb200_stokes*, const double*, double*, void*)
NOSTOKES(sb200_stokes_matmult_vv, sb200_stokes*, const double*, double*, void*)
NOSTOKES(sb200_stokes_matmult_pv, sb200_stokes*, const double*, double*, void*)
NOSTOKES(sb200_stokes_matmult_vp, sb200_stokes*, const double*, double*, void*)
NOSTOKES(sb200_stokes_get_diagonal_schur, sb200_stokes*, double*, void*)
NOSTOKES(sb200_stokes_matmult_schur, sb200_stokes*, const double*, double*, sb200_velocity_solve_fn, void*, void*)
NOSTOKES(sb200_stokes_function, sb200_stokes*, const double*, double*, void*)
NOSTOKES(sb200_stokes_pc_velocity_sizes, sb200_stokes*, long long*, long long*)
NOSTOKES(sb200_stokes_pc_velocity_csr, sb200_stokes*, int*, int*, double*, void*)

}  // extern "C"
