// TEST DOUBLE - never built into libspectral_b200.so, never shipped, never used by bench.py or the package.
//
// A plain-CPU stand-in for the DEVICE-side entry points of include/spectral_b200.h (memory helpers, the elliptic and Stokes shells,
// the FGMRES), so that the HOST layer of the product - host/reference_api.cpp, host/petsc_shim.cpp, host/host_ilu.cpp,
// csrc/exact.cpp, csrc/fd_rows.h and the native executable apps/elliptic.cpp - can be linked UNCHANGED against it and driven
// end to end in the CPU test suite (tests/test_native_cpu_double.py): option handling, Newton loop, PC refresh, printed lines,
// iteration counts.  "Device" pointers are host pointers here.  The arithmetic follows the same definitions as the oracle
// (dense CGL differentiation matrix per axis, reference operation order), which is all a host-logic test needs; GPU parity
// is tested on the GPU (tests/test_gpu_*.py), never through this file.
#include <algorithm>
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

struct sb200_stokes {
  int d;
  std::vector<int> dim;
  long long m, gp;
  std::vector<long long> stride, ixI, ixB;  // interior / boundary nodes in walk order
  std::vector<std::vector<double>> D, xnode, strain;
  std::vector<double> eta, deta, dirichlet, force;
  int rheology = 0;
  double hardness = 1.0, exponent = 1.0, reg = 1.0, gamma0 = 1.0, min_eta = 1.0, max_eta = 1.0;
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

// ---- Stokes helpers ---------------------------------------------------------------------------------------------------------
// derivative along `axis` of a field with nc interleaved components per node
void st_deriv(const sb200_stokes* s, int axis, const double* x, int nc, double* y) {
  const int P = s->dim[axis];
  const long long R = s->stride[axis], O = s->m / (R * P);
  const double* D = s->D[axis].data();
  for (long long o = 0; o < O; o++)
    for (long long r = 0; r < R; r++)
      for (int c = 0; c < nc; c++)
        for (int i = 0; i < P; i++) {
          double acc = 0;
          for (int j = 0; j < P; j++) acc += D[(size_t)i * P + j] * x[((o * P + j) * R + r) * nc + c];
          y[((o * P + i) * R + r) * nc + c] = acc;
        }
}

std::vector<double> st_vel_local(const sb200_stokes* s, const double* vG, bool with_dirichlet) {
  const int d = s->d;
  std::vector<double> xL(s->m * d, 0.0);
  for (long long q = 0; q < s->gp; q++)
    for (int k = 0; k < d; k++) xL[s->ixI[q] * d + k] = vG[q * d + k];
  if (with_dirichlet)
    for (size_t q = 0; q < s->ixB.size(); q++)
      for (int k = 0; k < d; k++) xL[s->ixB[q] * d + k] = s->dirichlet[q * d + k];
  return xL;
}

// Neville evaluation of the interpolant through (x_i, f_i), i < n, at the two points t0 and t1 (util.C:129-144 computes the same)
void neville2(int n, const double* x, const double* f, double t0, double t1, double* f0, double* f1) {
  std::vector<double> a(f, f + n), b(f, f + n);
  for (int lvl = 1; lvl < n; lvl++)
    for (int i = 0; i < n - lvl; i++) {
      const double den = x[i] - x[i + lvl];
      a[i] = ((t0 - x[i + lvl]) * a[i] + (x[i] - t0) * a[i + 1]) / den;
      b[i] = ((t1 - x[i + lvl]) * b[i] + (x[i] - t1) * b[i + 1]) / den;
    }
  *f0 = a[0];
  *f1 = b[0];
}

// StokesPressureReduceOrder (stokes.C:1029-1080): extend the interior pressure to the boundary nodes, z lines then y lines
// (planes i >= 1), then x lines; later passes read what earlier ones wrote
void st_reduce_order(const sb200_stokes* s, double* pres) {
  const int d = s->d, m = s->dim[0], n = s->dim[1], p = d == 2 ? 1 : s->dim[2];
  std::vector<double> f(std::max(std::max(m, n), p));
  for (int i = 1; i < m; i++) {
    if (p > 1)
      for (int j = 1; j < n; j++) {
        double* line = pres + ((long long)i * n + j) * p;
        for (int k = 1; k < p - 1; k++) f[k - 1] = line[k];
        neville2(p - 2, s->xnode[2].data() + 1, f.data(), s->xnode[2][0], s->xnode[2][p - 1], &line[0], &line[p - 1]);
      }
    for (int k = 0; k < p; k++) {
      double* line = pres + (long long)i * n * p + k;
      for (int j = 1; j < n - 1; j++) f[j - 1] = line[(long long)j * p];
      neville2(n - 2, s->xnode[1].data() + 1, f.data(), s->xnode[1][0], s->xnode[1][n - 1], &line[0], &line[(long long)(n - 1) * p]);
    }
  }
  for (int j = 0; j < n; j++)
    for (int k = 0; k < p; k++) {
      double* line = pres + (long long)j * p + k;
      const long long st = (long long)n * p;
      for (int i = 1; i < m - 1; i++) f[i - 1] = line[i * st];
      neville2(m - 2, s->xnode[0].data() + 1, f.data(), s->xnode[0][0], s->xnode[0][m - 1], &line[0], &line[(m - 1) * st]);
    }
}

std::vector<double> st_vv(const sb200_stokes* s, const double* xG) {  // StokesMatMultVV (stokes.C:623-676)
  const int d = s->d;
  const long long m = s->m;
  const std::vector<double> xL = st_vel_local(s, xG, false);
  std::vector<std::vector<double>> V(d, std::vector<double>(m * d)), W(d, std::vector<double>(m * d));
  for (int j = 0; j < d; j++) st_deriv(s, j, xL.data(), d, V[j].data());
  for (long long i = 0; i < m; i++) {
    double e[3][3], z = 0;
    for (int j = 0; j < d; j++)
      for (int k = 0; k < d; k++) {
        e[j][k] = 0.5 * (V[j][i * d + k] + V[k][i * d + j]);
        z += e[j][k] * s->strain[j][i * d + k];
      }
    for (int j = 0; j < d; j++)
      for (int k = 0; k < d; k++) W[j][i * d + k] = s->eta[i] * e[j][k] + s->deta[i] * s->strain[j][i * d + k] * z;
  }
  std::vector<double> yL(m * d, 0.0), t(m * d), y(s->gp * d);
  for (int j = 0; j < d; j++) {
    st_deriv(s, j, W[j].data(), d, t.data());
    for (long long i = 0; i < m * d; i++) yL[i] -= t[i];
  }
  for (long long q = 0; q < s->gp; q++)
    for (int k = 0; k < d; k++) y[q * d + k] = yL[s->ixI[q] * d + k];
  return y;
}

std::vector<double> st_div(const sb200_stokes* s, const double* vG, bool with_dirichlet) {  // StokesDivergence (stokes.C:570-595)
  const int d = s->d;
  const long long m = s->m;
  const std::vector<double> xL = st_vel_local(s, vG, with_dirichlet);
  std::vector<double> comp(m), t(m), acc(m, 0.0), y(s->gp);
  for (int k = 0; k < d; k++) {
    for (long long i = 0; i < m; i++) comp[i] = xL[i * d + k];
    st_deriv(s, k, comp.data(), 1, t.data());
    for (long long i = 0; i < m; i++) acc[i] += t[i];
  }
  for (long long q = 0; q < s->gp; q++) y[q] = acc[s->ixI[q]];
  return y;
}

std::vector<double> st_grad(const sb200_stokes* s, const double* pG) {  // StokesMatMultVP (stokes.C:599-619)
  const int d = s->d;
  const long long m = s->m;
  std::vector<double> pL(m, 0.0), t(m), y(s->gp * d);
  for (long long q = 0; q < s->gp; q++) pL[s->ixI[q]] = pG[q];
  st_reduce_order(s, pL.data());
  for (int k = 0; k < d; k++) {
    st_deriv(s, k, pL.data(), 1, t.data());
    for (long long q = 0; q < s->gp; q++) y[q * d + k] = t[s->ixI[q]];
  }
  return y;
}

}  // namespace

extern "C" {

const char* sb200_last_error(void) { return g_err.c_str(); }
long long sb200_launch_count(void) { return 0; }  // nothing is launched here
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

// ---- vector helpers (csrc/vecops.cu) ------------------------------------------------------------------------------------------
int sb200_vec_split(long long nodes, int d, const double* x, double* v, double* p, void*) {
  if (nodes < 0 || d < 1 || !x || (!v && !p)) FAIL(SB200_ERR_ARG, "sb200_vec_split: bad arguments");
  for (long long q = 0; q < nodes; q++) {
    if (v)
      for (int k = 0; k < d; k++) v[q * d + k] = x[q * (d + 1) + k];
    if (p) p[q] = x[q * (d + 1) + d];
  }
  return 0;
}
int sb200_vec_merge(long long nodes, int d, const double* v, const double* p, double* x, void*) {
  if (nodes < 0 || d < 1 || !x || (!v && !p)) FAIL(SB200_ERR_ARG, "sb200_vec_merge: bad arguments");
  for (long long q = 0; q < nodes; q++) {
    if (v)
      for (int k = 0; k < d; k++) x[q * (d + 1) + k] = v[q * d + k];
    if (p) x[q * (d + 1) + d] = p[q];
  }
  return 0;
}
int sb200_vec_axpby(long long n, double a, const double* x, double b, double* y, void*) {
  if (n < 0 || !y || (!x && a != 0.0)) FAIL(SB200_ERR_ARG, "sb200_vec_axpby: bad arguments");
  for (long long i = 0; i < n; i++) y[i] = a == 0.0 ? b * y[i] : (b == 0.0 ? a * x[i] : a * x[i] + b * y[i]);
  return 0;
}
int sb200_vec_pointwise_divide(long long n, const double* x, const double* dg, double* y, void*) {
  if (n < 0 || !x || !dg || !y) FAIL(SB200_ERR_ARG, "sb200_vec_pointwise_divide: bad arguments");
  for (long long i = 0; i < n; i++) y[i] = x[i] / dg[i];
  return 0;
}
int sb200_csr_diagonal(long long nrows, const int* rowptr, const int* colidx, const double* vals, double* diag, void*) {
  if (nrows < 0 || !rowptr || !colidx || !vals || !diag) FAIL(SB200_ERR_ARG, "sb200_csr_diagonal: bad arguments");
  for (long long r = 0; r < nrows; r++) {
    double v = 0.0;
    for (int q = rowptr[r]; q < rowptr[r + 1]; q++)
      if (colidx[q] == r) v = vals[q];
    diag[r] = v;
  }
  return 0;
}
int sb200_vec_remove_mean(long long n, int stride, int offset, double* x, double* scratch, void*) {
  if (n < 0 || stride < 1 || offset < 0 || offset >= stride || !x || !scratch) FAIL(SB200_ERR_ARG, "sb200_vec_remove_mean: bad arguments");
  double s = 0;
  for (long long i = 0; i < n; i++) s += x[offset + i * stride];
  s /= (double)n;
  for (long long i = 0; i < n; i++) x[offset + i * stride] -= s;
  return 0;
}
int sb200_vec_sum_count(long long n, int stride, int offset, const double* x, double* scratch, double* out2, void*) {
  if (n < 0 || stride < 1 || offset < 0 || offset >= stride || !scratch || !out2) FAIL(SB200_ERR_ARG, "sb200_vec_sum_count: bad arguments");
  double s = 0;
  for (long long i = 0; i < n; i++) s += x[offset + i * stride];
  out2[0] = s;
  out2[1] = (double)n;
  return 0;
}
int sb200_vec_shift_mean(long long n, int stride, int offset, double* x, const double* sums2, void*) {
  if (n < 0 || stride < 1 || offset < 0 || offset >= stride || !sums2) FAIL(SB200_ERR_ARG, "sb200_vec_shift_mean: bad arguments");
  const double mean = sums2[0] / sums2[1];
  for (long long i = 0; i < n; i++) x[offset + i * stride] -= mean;
  return 0;
}
int sb200_ksp_allreduce_sum(sb200_ksp*, double*, int, void*) { return 0; }  // single rank: the sum over the ranks is the value itself
int sb200_ksp_ipc_export(sb200_ksp*, void*) { FAIL(SB200_ERR_SUP, "test double: single rank only"); }
int sb200_ksp_ipc_attach(sb200_ksp*, int, const void*) { FAIL(SB200_ERR_SUP, "test double: single rank only"); }
int sb200_ksp_attach_local(sb200_ksp*, int, sb200_ksp*) { FAIL(SB200_ERR_SUP, "test double: single rank only"); }
int sb200_stream_sync(void*) { return 0; }
// measurement helper: the double has no tensor pipe; a nominal figure keeps bench.py's dry run going
int sb200_fp64_dmma_peak(double, double* tflops, double* measured_ms) {
  if (!tflops) FAIL(SB200_ERR_ARG, "null pointer");
  *tflops = 37.1;
  if (measured_ms) *measured_ms = 0.0;
  return 0;
}

// ---- ChebMult -------------------------------------------------------------------------------------------------------------
int sb200_cheb_create(int rank, int tr, const int* dims, long long n_total, sb200_cheb** out) {
  *out = nullptr;
  if (n_total < 2) FAIL(SB200_ERR_USER, "n must be >= 2");
  if (tr < 0 || tr >= rank) FAIL(SB200_ERR_USER, "tdim out of range");
  long long s = 1;
  for (int r = 0; r < rank; r++) s *= dims[r];
  if (s != n_total) FAIL(SB200_ERR_USER, "dimensions do not agree");
  if (dims[tr] < 2) FAIL(SB200_ERR_USER, "transformed extent must be >= 2");
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
const char* sb200_elliptic_last_kernel(const sb200_elliptic*) { return "cpu test double"; }
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
int sb200_ksp_set_lookahead(sb200_ksp*, int depth) { return (depth == 0 || depth == 1) ? 0 : 83; }  // a scheduling choice of the CUDA path; same iterates
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

// ---- Stokes shells (stokes.C:499-758, 1029-1080, 1160-1240, 1920-1944), -boundary 0 ------------------------------------------
int sb200_stokes_create(int d, const int* dim, sb200_stokes** out) {
  *out = nullptr;
  if (d != 2 && d != 3) FAIL(SB200_ERR_USER, "the Stokes shells need 2 or 3 dimensions (stokes.C:1036)");
  for (int j = 0; j < d; j++)
    if (dim[j] < 3) FAIL(SB200_ERR_USER, "each extent must be >= 3");
  sb200_stokes* s = new sb200_stokes();
  s->d = d;
  s->dim.assign(dim, dim + d);
  s->stride.assign(d, 1);
  s->m = 1;
  for (int j = d - 1; j >= 0; j--) {
    s->stride[j] = s->m;
    s->m *= dim[j];
  }
  std::vector<int> ind(d, 0);
  for (long long node = 0; node < s->m; node++) {  // StokesSetupDomain walk (stokes.C:791-879)
    bool bdy = false;
    for (int j = 0; j < d; j++) bdy = bdy || ind[j] == 0 || ind[j] == dim[j] - 1;
    (bdy ? s->ixB : s->ixI).push_back(node);
    for (int j = d - 1; j >= 0; j--) {
      if (++ind[j] < dim[j]) break;
      ind[j] = 0;
    }
  }
  s->gp = (long long)s->ixI.size();
  for (int j = 0; j < d; j++) {
    s->D.push_back(cgl_diff_matrix(dim[j]));
    std::vector<double> x(dim[j]);
    for (int i = 0; i < dim[j]; i++) x[i] = cos(i * M_PI / (dim[j] - 1));
    s->xnode.push_back(x);
  }
  s->eta.assign(s->m, 1.0);
  s->deta.assign(s->m, 0.0);
  s->strain.assign(d, std::vector<double>(s->m * d, 0.0));
  s->dirichlet.assign(s->ixB.size() * d, 0.0);
  s->force.assign(s->gp * (d + 1), 0.0);
  *out = s;
  return 0;
}
int sb200_stokes_destroy(sb200_stokes* s) {
  delete s;
  return 0;
}
int sb200_stokes_sizes(const sb200_stokes* s, long long* m, long long* g, long long* gp, long long* gv, long long* dv) {
  if (m) *m = s->m;
  if (g) *g = s->gp * (s->d + 1);
  if (gp) *gp = s->gp;
  if (gv) *gv = s->gp * s->d;
  if (dv) *dv = (long long)s->ixB.size() * s->d;
  return 0;
}
int sb200_stokes_set_rheology(sb200_stokes* s, int type, double hardness, double exponent, double regularization, double gamma0) {
  if (type != 0 && type != 1) FAIL(SB200_ERR_SUP, "Rheology type not implemented");
  s->rheology = type;
  s->hardness = hardness;
  s->exponent = exponent;
  s->reg = regularization;
  s->gamma0 = gamma0;
  return 0;
}
int sb200_stokes_set_dirichlet(sb200_stokes* s, const double* v, void*) {
  s->dirichlet.assign(v, v + s->dirichlet.size());
  return 0;
}
int sb200_stokes_set_force(sb200_stokes* s, const double* f, void*) {
  s->force.assign(f, f + s->force.size());
  return 0;
}
int sb200_stokes_matmult_vv(sb200_stokes* s, const double* x, double* y, void*) {
  std::vector<double> v = st_vv(s, x);
  std::copy(v.begin(), v.end(), y);
  return 0;
}
int sb200_stokes_matmult_pv(sb200_stokes* s, const double* x, double* y, void*) {
  std::vector<double> p = st_div(s, x, false);
  std::copy(p.begin(), p.end(), y);
  return 0;
}
int sb200_stokes_set_fold_pressure(sb200_stokes*, int) { return 0; }      // likewise
int sb200_stokes_set_graph(sb200_stokes*, int) { return 0; }              // likewise
int sb200_stokes_set_trace_divergence(sb200_stokes*, int) { return 0; }  // an implementation choice of the CUDA path; nothing to switch here
int sb200_stokes_divergence(sb200_stokes* s, int with_dirichlet, const double* x, double* y, void*) {
  std::vector<double> p = st_div(s, x, with_dirichlet != 0);
  std::copy(p.begin(), p.end(), y);
  return 0;
}
int sb200_stokes_matmult_vp(sb200_stokes* s, const double* x, double* y, void*) {
  std::vector<double> v = st_grad(s, x);
  std::copy(v.begin(), v.end(), y);
  return 0;
}
int sb200_stokes_matmult(sb200_stokes* s, const double* x, double* y, void*) {  // stokes.C:499-519
  const int d = s->d;
  std::vector<double> v(s->gp * d), p(s->gp);
  for (long long q = 0; q < s->gp; q++) {
    for (int k = 0; k < d; k++) v[q * d + k] = x[q * (d + 1) + k];
    p[q] = x[q * (d + 1) + d];
  }
  const std::vector<double> vv = st_vv(s, v.data()), pv = st_div(s, v.data(), false), vp = st_grad(s, p.data());
  for (long long q = 0; q < s->gp; q++) {
    for (int k = 0; k < d; k++) y[q * (d + 1) + k] = vv[q * d + k] + vp[q * d + k];
    y[q * (d + 1) + d] = pv[q];
  }
  return 0;
}
int sb200_stokes_get_diagonal_schur(sb200_stokes* s, double* y, void*) {
  for (long long q = 0; q < s->gp; q++) y[q] = 1.0 / s->eta[s->ixI[q]];
  return 0;
}
int sb200_stokes_matmult_schur(sb200_stokes* s, const double* x, double* y, sb200_velocity_solve_fn solve, void* solve_ctx, void* stream) {
  if (!solve) FAIL(SB200_ERR_ARG, "StokesMatMultSchur needs the inner velocity solve");
  std::vector<double> v0 = st_grad(s, x), v1(v0.size());
  if (int rc = solve(solve_ctx, v0.data(), v1.data(), stream)) return rc;
  std::vector<double> p = st_div(s, v1.data(), false);
  for (long long q = 0; q < s->gp; q++) y[q] = -p[q];
  return 0;
}
int sb200_stokes_function(sb200_stokes* s, const double* x, double* y, void*) {  // stokes.C:680-758
  const int d = s->d;
  const long long m = s->m;
  std::vector<double> v(s->gp * d), p(s->gp);
  for (long long q = 0; q < s->gp; q++) {
    for (int k = 0; k < d; k++) v[q * d + k] = x[q * (d + 1) + k];
    p[q] = x[q * (d + 1) + d];
  }
  std::vector<double> xL = st_vel_local(s, v.data(), true);
  std::vector<std::vector<double>> raw(d, std::vector<double>(m * d)), V(d, std::vector<double>(m * d));
  for (int j = 0; j < d; j++) st_deriv(s, j, xL.data(), d, raw[j].data());
  s->min_eta = 1e300;
  s->max_eta = -1e300;
  for (long long i = 0; i < m; i++) {
    double gamma = 0;
    for (int j = 0; j < d; j++)
      for (int k = 0; k < d; k++) {
        const double e = 0.5 * (raw[j][i * d + k] + raw[k][i * d + j]);
        s->strain[j][i * d + k] = e;
        gamma += 0.5 * (e * e);
      }
    if (s->rheology == 0) {
      s->eta[i] = 1.0;
      s->deta[i] = 0.0;
    } else {  // StokesRheologyPower (stokes.C:1930-1944)
      const double n = s->exponent, pw = (1.0 - n) / (2.0 * n), base = s->reg + gamma / s->gamma0;
      s->eta[i] = s->hardness * pow(base, pw);
      s->deta[i] = fabs(n) > 1.0e-5 ? s->hardness * pw / s->gamma0 * pow(base, pw - 1.0) : 0.0;
    }
    s->min_eta = fmin(s->min_eta, s->eta[i]);
    s->max_eta = fmax(s->max_eta, s->eta[i]);
    for (int j = 0; j < d; j++)
      for (int k = 0; k < d; k++) V[j][i * d + k] = s->eta[i] * s->strain[j][i * d + k];
  }
  std::vector<double> yL(m * d, 0.0), t(m * d);
  for (int j = 0; j < d; j++) {
    st_deriv(s, j, V[j].data(), d, t.data());
    for (long long i = 0; i < m * d; i++) yL[i] -= t[i];
  }
  const std::vector<double> pv = st_div(s, v.data(), true), vp = st_grad(s, p.data());
  for (long long q = 0; q < s->gp; q++) {
    for (int k = 0; k < d; k++) y[q * (d + 1) + k] = yL[s->ixI[q] * d + k] + vp[q * d + k] - s->force[q * (d + 1) + k];
    y[q * (d + 1) + d] = pv[q] - s->force[q * (d + 1) + d];
  }
  return 0;
}
int sb200_stokes_eta_minmax(sb200_stokes* s, double* mn, double* mx, void*) {
  *mn = s->min_eta;
  *mx = s->max_eta;
  return 0;
}
int sb200_stokes_get_state(sb200_stokes* s, int which, double* out, void*) {
  if (which == 0) std::copy(s->eta.begin(), s->eta.end(), out);
  else if (which == 1) std::copy(s->deta.begin(), s->deta.end(), out);
  else if (which >= 2 && which < 2 + s->d) std::copy(s->strain[which - 2].begin(), s->strain[which - 2].end(), out);
  else FAIL(SB200_ERR_USER, "state selector out of range");
  return 0;
}
int sb200_stokes_pressure_reduce_order(sb200_stokes* s, double* pL, void*) {
  st_reduce_order(s, pL);
  return 0;
}
int sb200_stokes_pc_velocity_sizes(sb200_stokes* s, long long* nrows, long long* nnz) {
  FdGrid G;
  fd_grid_init(&G, s->d, s->dim.data());
  if (nrows) *nrows = G.g * s->d;
  if (nnz) *nnz = fd_total_entries(G) * s->d;
  return 0;
}
int sb200_stokes_pc_velocity_csr(sb200_stokes* s, int* rowptr, int* colidx, double* vals, void*) {
  FdGrid G;
  fd_grid_init(&G, s->d, s->dim.data());
  const int nc = s->d;
  std::vector<double> x;
  for (int j = 0; j < s->d; j++) x.insert(x.end(), s->xnode[j].begin(), s->xnode[j].end());
  FdFields F;
  F.xtab = x.data();
  F.eta = s->eta.data();
  F.deta = nullptr;
  for (int j = 0; j < SB200_FD_MAX_DIM; j++) F.gradu[j] = nullptr;
  for (long long r = 0; r < G.g; r++) {
    int k[SB200_FD_MAX_DIM];
    long long cols[2 * SB200_FD_MAX_DIM + 1];
    double v[2 * SB200_FD_MAX_DIM + 1];
    const long long node = fd_decode(G, r, k);
    const int n = fd_row(G, F, r, k, node, cols, v);
    const long long base = fd_row_offset(G, k, r) * nc;
    for (int f = 0; f < nc; f++) {
      const long long o = base + (long long)f * n;
      if (rowptr) rowptr[r * nc + f] = (int)o;
      for (int q = 0; q < n; q++) {
        if (colidx) colidx[o + q] = (int)(cols[q] * nc + f);
        vals[o + q] = v[q];
      }
    }
  }
  if (rowptr) rowptr[G.g * nc] = (int)(fd_total_entries(G) * nc);
  return 0;
}

// ---- what the Python harness (spectral_petsc_b200/capi.py) calls besides the above, so that the GPU tests' own logic can be
// dry-run on CPU over this double (tests/test_gpu_dry_run_cpu.py); single rank only ------------------------------------------------
int sb200_version(void) { return 100; }
int sb200_ksp_create_slab(long long n_local, int restart, int rank, int nranks, sb200_ksp** out) {
  if (rank != 0 || nranks != 1) FAIL(SB200_ERR_SUP, "test double: single rank only");
  return sb200_ksp_create(n_local, restart, out);
}
int sb200_ksp_get_history(const sb200_ksp* k, double* h_hist, int cap, int* n) {
  *n = (int)k->history.size();
  if (h_hist)
    for (int i = 0; i < cap && i < *n; i++) h_hist[i] = k->history[i];
  return 0;
}
int sb200_ksp_get_times(const sb200_ksp*, double* a, double* b, double* c) {
  if (a) *a = 0;
  if (b) *b = 0;
  if (c) *c = 0;
  return 0;
}
int sb200_stokes_slab_info(const sb200_stokes* s, int* rank, int* nranks, int* i0, int* nloc, long long* goff_nodes) {
  if (rank) *rank = 0;
  if (nranks) *nranks = 1;
  if (i0) *i0 = 0;
  if (nloc) *nloc = s->dim[0];
  if (goff_nodes) *goff_nodes = 0;
  return 0;
}
int sb200_elliptic_slab_info(const sb200_elliptic* e, int* rank, int* nranks, int* i0, int* nloc, long long* goff, long long* gtotal) {
  if (rank) *rank = 0;
  if (nranks) *nranks = 1;
  if (i0) *i0 = 0;
  if (nloc) *nloc = e->dim[0];
  if (goff) *goff = 0;
  if (gtotal) *gtotal = e->g;
  return 0;
}
int sb200_elliptic_set_path(sb200_elliptic*, int path) { return path >= 0 && path <= 4 ? 0 : SB200_ERR_USER; }  // one CPU path here
// the remaining single-rank entry points of the Python harness (dry runs of the GPU test modules)
int sb200_device_count(int* n) { *n = 1; return 0; }
int sb200_set_device(int) { return 0; }
int sb200_cheb_apply_host(sb200_cheb* c, const double* h_x, double* h_y) { return sb200_cheb_apply(c, h_x, h_y, nullptr); }
int sb200_elliptic_function_host(sb200_elliptic* e, const double* h_U, double* h_F) { return sb200_elliptic_function(e, h_U, h_F, nullptr); }
int sb200_stokes_matmult_host(sb200_stokes* s, const double* h_x, double* h_y) { return sb200_stokes_matmult(s, h_x, h_y, nullptr); }
int sb200_stokes_function_host(sb200_stokes* s, const double* h_x, double* h_y) { return sb200_stokes_function(s, h_x, h_y, nullptr); }
int sb200_elliptic_create_slab(int d, const int* dim, int rank, int nranks, sb200_elliptic** out) {
  if (rank != 0 || nranks != 1) FAIL(SB200_ERR_SUP, "test double: single rank only");
  return sb200_elliptic_create(d, dim, out);
}
int sb200_stokes_create_slab(int d, const int* dim, int rank, int nranks, sb200_stokes** out) {
  if (rank != 0 || nranks != 1) FAIL(SB200_ERR_SUP, "test double: single rank only");
  return sb200_stokes_create(d, dim, out);
}
int sb200_elliptic_get_state(sb200_elliptic* e, int which, double* out, void*) {  // 0 eta, 1 deta, 2+j gradu[j]
  if (which < 0 || which > 1 + e->d) FAIL(SB200_ERR_USER, "get_state: which out of range");
  const std::vector<double>& a = which == 0 ? e->eta : (which == 1 ? e->deta : e->gradu[which - 2]);
  std::copy(a.begin(), a.end(), out);
  return 0;
}
int sb200_elliptic_pad(sb200_elliptic* e, const double* U, int with_dirichlet, double* local, void*) {
  std::fill(local, local + e->m, 0.0);
  for (long long q = 0; q < e->g; q++) local[e->ixG[q]] = U[q];
  for (size_t q = 0; q < e->ixD.size(); q++) local[e->ixD[q]] = with_dirichlet ? e->dirichlet[q] : 0.0;
  return 0;
}
int sb200_elliptic_crop(sb200_elliptic* e, const double* local, double* U, void*) {
  for (long long q = 0; q < e->g; q++) U[q] = local[e->ixG[q]];
  return 0;
}
// host-buffer forms: "device" memory is host memory here, so they are the operator itself; the queue completes at submit
static int g_pending = 0;
int sb200_elliptic_matmult_host(sb200_elliptic* e, const double* h_U, double* h_V) { return sb200_elliptic_matmult(e, h_U, h_V, nullptr); }
int sb200_elliptic_matmult_host_submit(sb200_elliptic* e, const double* h_U, double* h_V) {
  if (g_pending >= 4) FAIL(SB200_ERR_USER, "host queue full: call sb200_elliptic_matmult_host_wait first");
  g_pending++;
  return sb200_elliptic_matmult(e, h_U, h_V, nullptr);
}
int sb200_elliptic_matmult_host_wait(sb200_elliptic*) {
  if (g_pending <= 0) FAIL(SB200_ERR_USER, "host queue empty: nothing was submitted");
  g_pending--;
  return 0;
}
int sb200_elliptic_matmult_host_pending(const sb200_elliptic*, int* pending) {
  *pending = g_pending;
  return 0;
}
int sb200_apply_elliptic_matmult(void* ctx, const double* x, double* y, void* stream) { return sb200_elliptic_matmult((sb200_elliptic*)ctx, x, y, stream); }
int sb200_apply_stokes_matmult(void* ctx, const double* x, double* y, void* stream) { return sb200_stokes_matmult((sb200_stokes*)ctx, x, y, stream); }
int sb200_apply_stokes_matmult_vv(void* ctx, const double* x, double* y, void* stream) { return sb200_stokes_matmult_vv((sb200_stokes*)ctx, x, y, stream); }

}  // extern "C"
