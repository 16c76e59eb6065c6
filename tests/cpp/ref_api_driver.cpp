// Drives the B200 path through the reference's own interface only (include/sb200_reference_api.h):
// the checks are the reference's self-checks - cheb.c (d/dx e^x = e^x), the CHECK_EXACT block of
// elliptic.C:193-209, and stokes.C:190-212 (exact residual + constant-pressure null space).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/sb200_reference_api.h"
#include "../../include/spectral_b200.h"

#define CHK(expr)                                                                          \
  do {                                                                                     \
    PetscErrorCode _e = (expr);                                                            \
    if (_e) {                                                                              \
      fprintf(stderr, "%s:%d error %d: %s\n", __FILE__, __LINE__, _e, sb200_last_error()); \
      return _e;                                                                           \
    }                                                                                      \
  } while (0)

static double norm_inf(Vec v) {
  PetscInt n;
  VecGetSize(v, &n);
  std::vector<double> h(n);
  VecGetValuesHost(v, h.data());
  double m = 0;
  for (double x : h) m = fmax(m, fabs(x));
  return m;
}

// The reference has no self-check for its preconditioning matrices beyond the solve converging, so use what the stencil
// guarantees: with eta = 1, eta' = 0 the rows are the 3-point finite differences of -Laplace on the Chebyshev nodes, which are
// exact for quadratics: (P q)_r = -2 for q = x_axis^2 on every row with a full stencil, and its row sum vanishes.
static int check_fd_matrix(Mat P, int d, const int* dim, int ncomp, const char* tag) {
  PetscInt nz = 0, rows = 0;
  CHK(MatGetSize(P, &rows, PETSC_NULL));
  CHK(MatSeqAIJGetCSRHost(P, &nz, PETSC_NULL, PETSC_NULL, PETSC_NULL));
  std::vector<PetscInt> rp(rows + 1), ci(nz);
  std::vector<double> v(nz);
  CHK(MatSeqAIJGetCSRHost(P, PETSC_NULL, rp.data(), ci.data(), v.data()));
  const double PI = 3.14159265358979323846;
  const int axis = d - 2;  // an axis that is neither the slowest nor the fastest one in 3-D
  long long istride = 1;
  for (int j = d - 1; j > axis; j--) istride *= dim[j] - 2;
  auto coord = [&](PetscInt row) {
    const long long node = row / ncomp;
    const int k = (int)((node / istride) % (dim[axis] - 2));
    return cos((k + 1) * PI / (dim[axis] - 1));
  };
  double worst_q = 0, worst_sum = 0;
  long long full = 0;
  bool sorted = rp[0] == 0 && rp[rows] == nz;
  for (PetscInt r = 0; r < rows; r++) {
    for (PetscInt e = rp[r] + 1; e < rp[r + 1]; e++) sorted = sorted && ci[e] > ci[e - 1];
    if (rp[r + 1] - rp[r] != 2 * d + 1) continue;
    full++;
    double q = 0, sum = 0;
    for (PetscInt e = rp[r]; e < rp[r + 1]; e++) {
      const double x = coord(ci[e]);
      q += v[e] * x * x;
      sum += v[e];
    }
    worst_q = fmax(worst_q, fabs(q + 2.0));
    worst_sum = fmax(worst_sum, fabs(sum));
  }
  printf("%s P rows %d nz %d sorted %d full-stencil rows %lld  max |P x^2 + 2| = %.3e  max |row sum| = %.3e\n", tag, rows, nz, (int)sorted, full,
         worst_q, worst_sum);
  return 0;
}

static int test_cheb() {  // cheb.c:68-112 with the defaults m1=5, (m,n,p)=(8,7,6), all axes
  const double PI = 3.14159265358979323846;
  {
    int m1 = 5;
    Vec u, b;
    Mat A;
    CHK(VecCreateSeqCUDA(PETSC_COMM_WORLD, m1, &u));
    CHK(VecDuplicate(u, &b));
    int dims[1] = {m1};
    CHK(MatCreateCheb(PETSC_COMM_WORLD, 1, 0, dims, FFTW_ESTIMATE, u, b, &A));
    std::vector<double> a(m1), r(m1);
    for (int i = 0; i < m1; i++) a[i] = exp(cos(i * PI / (m1 - 1)));
    CHK(VecSetValuesHost(u, a.data()));
    CHK(MatMult(A, u, b));
    CHK(VecGetValuesHost(b, r.data()));
    double norm = 0;
    for (int i = 0; i < m1; i++) norm = fmax(norm, fabs(r[i] - a[i]));
    printf("cheb1d Norm of error %.12e\n", norm);
    CHK(MatDestroy(A));
    CHK(VecDestroy(u));
    CHK(VecDestroy(b));
  }
  int m = 8, n = 7, p = 6;
  for (int d = 0; d < 3; d++) {
    Vec u2, b2;
    Mat A2;
    CHK(VecCreateSeqCUDA(PETSC_COMM_WORLD, m * n * p, &u2));
    CHK(VecDuplicate(u2, &b2));
    int dims[3] = {m, n, p};
    CHK(MatCreateCheb(PETSC_COMM_WORLD, 3, d, dims, FFTW_ESTIMATE, u2, b2, &A2));
    std::vector<double> a(m * n * p), e(m * n * p), r(m * n * p);
    for (int i = 0; i < m; i++) {
      double x = cos(i * PI / (m - 1));
      for (int j = 0; j < n; j++) {
        double y = cos(j * PI / (n - 1));
        for (int k = 0; k < p; k++) {
          double z = cos(k * PI / (p - 1));
          a[(i * n + j) * p + k] = exp(x) + exp(y) + exp(z);
          e[(i * n + j) * p + k] = d == 0 ? exp(x) : (d == 1 ? exp(y) : exp(z));
        }
      }
    }
    CHK(VecSetValuesHost(u2, a.data()));
    CHK(MatMult(A2, u2, b2));
    CHK(VecGetValuesHost(b2, r.data()));
    double norm = 0;
    for (size_t i = 0; i < r.size(); i++) norm = fmax(norm, fabs(r[i] - e[i]));
    printf("cheb3d axis %d Norm of error %.12e\n", d, norm);
    CHK(MatDestroy(A2));
    CHK(VecDestroy(u2));
    CHK(VecDestroy(b2));
  }
  // error behaviour: tr out of range -> PETSC_ERR_USER (chebyshev.c:106)
  Vec u, b;
  Mat A;
  int dims[2] = {4, 4};
  VecCreateSeqCUDA(PETSC_COMM_WORLD, 16, &u);
  VecDuplicate(u, &b);
  printf("cheb bad tr -> %d\n", MatCreateCheb(PETSC_COMM_WORLD, 2, 2, dims, FFTW_ESTIMATE, u, b, &A));
  VecDestroy(u);
  VecDestroy(b);
  return 0;
}

static int test_elliptic(int d, int* dim, int exact) {  // elliptic.C:159-209
  AppCtx ac;
  ac.d = d;
  ac.dim = dim;
  ac.exact = exact;
  ac.gamma = 0.0;
  ac.exponent = 2.0;
  ac.debug = 0;
  Vec u, u2, r;
  Mat A;
  SNES snes;
  CHK(MatCreate_Elliptic(PETSC_COMM_WORLD, d, dim, FFTW_ESTIMATE, DirichletBdy, &u, &A));
  PetscInt m, n;
  CHK(MatGetSize(A, &m, &n));
  CHK(VecDuplicate(u, &u2));
  CHK(VecDuplicate(u, &r));
  CHK(VecDuplicate(u, &ac.b));
  CHK(SNESCreate(PETSC_COMM_WORLD, &snes));
  CHK(SNESSetApplicationContext(snes, &ac));
  ac.A = A;
  CHK(CreateExactSolution(snes, u, u2, 0.0));
  CHK(FormFunction(snes, u, r, &ac));
  printf("elliptic global dofs %d\n", m);
  printf("%-25s: abs = %8e\n", "Norm of exact residual", norm_inf(r));
  CHK(MatMult(A, u, r));
  printf("elliptic |A u| = %8e\n", norm_inf(r));
  {  // SNESSetJacobian(snes, A, P, FormJacobian, ac) (elliptic.C:167-178): P about the state FormFunction just cached
    Mat P;
    MatStructure flag = DIFFERENT_NONZERO_PATTERN;
    CHK(MatCreateSeqAIJ(PETSC_COMM_SELF, m, n, 1 + 2 * d, PETSC_NULL, &P));
    CHK(FormJacobian(snes, u, &A, &P, &flag, &ac));
    CHK(check_fd_matrix(P, d, dim, 1, "elliptic"));
    std::vector<double> v0(P->nz), v1(P->nz);
    CHK(MatSeqAIJGetCSRHost(P, PETSC_NULL, PETSC_NULL, PETSC_NULL, v0.data()));
    CHK(FormJacobian(snes, u, &A, &P, &flag, &ac));  // second Newton step: values refreshed into the same pattern
    CHK(MatSeqAIJGetCSRHost(P, PETSC_NULL, PETSC_NULL, PETSC_NULL, v1.data()));
    double diff = 0;
    for (size_t i = 0; i < v0.size(); i++) diff = fmax(diff, fabs(v0[i] - v1[i]));
    printf("elliptic P refresh flag %d max diff %.3e\n", (int)flag, diff);
    CHK(MatDestroy(P));
  }
  CHK(SNESDestroy(snes));
  CHK(MatDestroy(A));
  CHK(VecDestroy(u));
  CHK(VecDestroy(u2));
  CHK(VecDestroy(r));
  CHK(VecDestroy(ac.b));
  return 0;
}

static int test_stokes(int n0) {  // stokes.C:139-212
  StokesOptionsB200 opt;
  opt.numDims = 3;
  opt.dim[0] = opt.dim[1] = opt.dim[2] = n0;
  opt.exact = 2;
  opt.rheology = 0;
  opt.hardness = 1.0;
  opt.exponent = 1.0;
  opt.regularization = 1.0;
  opt.gamma0 = 1.0;
  Mat A;
  Vec x, r, u, u2;
  StokesCtxB200* ctx;
  SNES snes;
  CHK(StokesCreate(PETSC_COMM_SELF, &opt, &A, &x, &ctx));
  CHK(VecDuplicate(x, &r));
  CHK(VecDuplicate(x, &u));
  CHK(VecDuplicate(x, &u2));
  CHK(SNESCreate(PETSC_COMM_SELF, &snes));
  CHK(SNESSetApplicationContext(snes, ctx));
  CHK(StokesCreateExactSolution(snes, u, u2));
  CHK(StokesFunction(snes, u, r, ctx));
  printf("Norm of solution %9.3e  norm of forcing %9.3e  norm of residual %9.3e\n", norm_inf(u), norm_inf(u2), norm_inf(r));
  // MatNullSpaceTest(ns, A): A [0; 1_p] = 0 (stokes.C:206-212, 1013-1023)
  PetscInt g;
  VecGetSize(x, &g);
  std::vector<double> ns(g, 0.0);
  for (PetscInt i = 3; i < g; i += 4) ns[i] = 1.0;
  CHK(VecSetValuesHost(x, ns.data()));
  CHK(MatMult(A, x, r));
  printf("Null space test |A ns| = %9.3e\n", norm_inf(r));
  {  // PCShellSetContext(pc, ctx); PCShellSetSetUp(pc, StokesPCSetUp0) (stokes.C:163-166)
    PC pc;
    Mat MatVVPC;
    CHK(PCCreate(PETSC_COMM_SELF, &pc));
    CHK(PCShellSetContext(pc, ctx));
    CHK(StokesPCSetUp0(pc));
    CHK(StokesGetPCMatrix(ctx, &MatVVPC));
    CHK(check_fd_matrix(MatVVPC, 3, opt.dim, 3, "stokes"));
    CHK(PCDestroy(pc));
  }
  CHK(SNESDestroy(snes));
  CHK(StokesDestroy(ctx));
  CHK(MatDestroy(A));
  CHK(VecDestroy(x));
  CHK(VecDestroy(r));
  CHK(VecDestroy(u));
  CHK(VecDestroy(u2));
  return 0;
}

int main() {
  if (test_cheb()) return 1;
  int d3[3] = {16, 16, 16};
  if (test_elliptic(3, d3, 2)) return 1;
  int d5[5] = {12, 12, 12, 12, 12};
  if (test_elliptic(5, d5, 2)) return 1;
  if (test_stokes(20)) return 1;
  printf("launches %lld\n", sb200_launch_count());
  return 0;
}
