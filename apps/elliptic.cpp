// ./elliptic - the reference's executable (elliptic.C:116-247) on the B200 path, in C++ with no Python anywhere:
//
//     apps/elliptic -dim 16,16,16 -exact 2 -ksp_rtol 1e-10          (BASELINE.json configs[0])
//
// Same options, same order of work, same printed lines.  What runs where:
//   * MatCreate_Elliptic / CreateExactSolution / FormFunction / FormJacobian / MatMult: the reference's own names
//     (include/sb200_reference_api.h) over the C-ABI library - operators, residual and the finite-difference matrix on the GPU;
//   * KSPFGMRES (elliptic.C:182): the device-resident sb200_ksp_* with the MatShell's MULT as its operator callback;
//   * PCILU with 2 levels (elliptic.C:183-184): PETSc's own, out of scope - applied here by the HOST stand-in sb200_host_ilu_*
//     (the matrix values come down, the factor is refreshed on the same pattern, each application copies the vector down and up);
//   * SNES: full Newton steps with the step halved while the residual norm does not decrease (PETSc's cubic line search is
//     PETSc's); its vector updates are done on the host copy of x (one 8 g-byte transfer per Newton step).
// The Python command line (python -m spectral_petsc_b200.elliptic) runs the identical flow; tests compare the two.
#include "common.h"

using app::HostPc;
using app::norm2;
using app::Options;

namespace {

// ---- KSP callbacks ---------------------------------------------------------------------------------------------------------
// MatMult(A, x, y) on raw device arrays, the way PETSc's KSP calls the MatShell
int op_matmult(void* ctx, const double* d_x, double* d_y, void*) {
  Mat A = (Mat)ctx;
  PetscInt n;
  MatGetSize(A, &n, PETSC_NULL);
  Vec x = nullptr, y = nullptr;
  PetscErrorCode rc = VecCreateSeqCUDAWithArray(PETSC_COMM_SELF, n, (double*)d_x, &x);
  if (!rc) rc = VecCreateSeqCUDAWithArray(PETSC_COMM_SELF, n, d_y, &y);
  if (!rc) rc = MatMult(A, x, y);
  if (x) VecDestroy(x);
  if (y) VecDestroy(y);
  return rc;
}

int pc_apply(void* ctx, const double* d_x, double* d_y, void* stream) { return ((HostPc*)ctx)->apply_device(d_x, d_y, stream); }

}  // namespace

int main(int argc, char** argv) {
  Options o;
  if (int rc = o.parse(argc, argv)) return rc;
  // ---- options (elliptic.C:137-149) -------------------------------------------------------------------------------------
  AppCtx ac;
  int dim[10] = {8, 6};
  ac.d = 2;
  if (const int nd = o.int_array("dim", dim, 10)) {
    if (nd < 0) {
      fprintf(stderr, "error: -dim takes 1..10 comma-separated integers (elliptic.C:138)\n");
      return 83;
    }
    ac.d = nd;
  }
  ac.dim = dim;
  ac.debug = o.integer("debug", 0);
  ac.exact = o.integer("exact", 0);
  ac.gamma = o.real("gamma", 0.0);
  ac.exponent = o.real("exponent", 2.0);
  const bool have_cos = o.has("cos_scale");
  const double cos_scale = o.real("cos_scale", 0.0);
  if ((ac.exact == 0 || ac.exact == 3) && !have_cos) {
    fprintf(stderr, "error: -exact %d needs -cos_scale (the reference reads it without a default, elliptic.C:607-609)\n", (int)ac.exact);
    return 83;
  }
  const double ksp_rtol = o.real("ksp_rtol", 1e-5), snes_rtol = o.real("snes_rtol", 1e-8), snes_atol = o.real("snes_atol", 1e-50);
  const int ksp_max_it = o.integer("ksp_max_it", 10000), restart = o.integer("ksp_gmres_restart", 30), snes_max_it = o.integer("snes_max_it", 50);
  HostPc pc;
  pc.type = o.str("pc_type", "ilu");  // PCSetType(pc, PCILU); PCFactorSetLevels(pc, 2): elliptic.C:183-184
  pc.levels = o.integer("pc_factor_levels", 2);
  if (!HostPc::known(pc.type)) {
    fprintf(stderr, "error: unknown PC type '%s' (have: ilu, jacobi, none)\n", pc.type.c_str());
    return 83;
  }
  if (o.str("ksp_type", "fgmres") != "fgmres") {
    fprintf(stderr, "error: only -ksp_type fgmres (the type the reference sets in code, elliptic.C:182) is built\n");
    return 83;
  }
  const bool ksp_monitor = o.has("ksp_monitor"), snes_monitor = o.has("snes_monitor");

  printf("Elliptic problem  dims = [");
  for (int i = 0; i < ac.d; i++) printf("%s%d", i ? "," : "", dim[i]);
  printf("]    gamma = %f    exponent = %8f\n", ac.gamma, ac.exponent);

  // ---- objects (elliptic.C:159-186) -------------------------------------------------------------------------------------
  Vec x, r, u, u2;
  Mat A, P;
  SNES snes;
  CHK(MatCreate_Elliptic(PETSC_COMM_WORLD, ac.d, ac.dim, FFTW_ESTIMATE, DirichletBdy, &u, &A));
  PetscInt m, n;
  CHK(MatGetSize(A, &m, &n));
  {
    long long local = 1;
    for (int i = 0; i < ac.d; i++) local *= dim[i];
    printf("DOF distribution: %8lld local     %8d global     %8lld dirichlet\n", local, m, local - m);  // elliptic.C:424
  }
  CHK(MatCreateSeqAIJ(PETSC_COMM_SELF, m, n, 1 + 2 * ac.d, PETSC_NULL, &P));
  CHK(VecDuplicate(u, &u2));
  CHK(VecDuplicate(u, &r));
  CHK(VecDuplicate(u, &x));
  CHK(VecDuplicate(u, &ac.b));
  CHK(SNESCreate(PETSC_COMM_WORLD, &snes));
  CHK(SNESSetApplicationContext(snes, &ac));
  ac.A = A;
  if (ac.exact < 0 || ac.exact > 2) {
    fprintf(stderr, "error: Choose an exact solution.\n");  // elliptic.C:657
    return 83;
  }
  CHK(CreateExactSolution(snes, u, u2, cos_scale));
  std::vector<double> hu(m), hu2(m), hr(m);
  CHK(VecGetValuesHost(u, hu.data()));
  CHK(VecGetValuesHost(u2, hu2.data()));

  // ---- CHECK_EXACT (elliptic.C:192-209) ---------------------------------------------------------------------------------
  CHK(FormFunction(snes, u, r, &ac));
  CHK(VecGetValuesHost(r, hr.data()));
  {
    double norm = 0, rnorm = 0;
    for (int i = 0; i < m; i++) {
      norm = fmax(norm, fabs(hr[i]));
      const double q = fabs(hr[i] / hu2[i]);  // VecPointwiseDivide(r, r, u2)
      if (q == q) rnorm = fmax(rnorm, q);
    }
    printf("%-25s: abs = %8e   rel = %8e\n", "Norm of exact residual", norm, rnorm);
  }

  // ---- SOLVE (elliptic.C:211-228) ---------------------------------------------------------------------------------------
  sb200_ksp* ksp = nullptr;
  CHK(sb200_ksp_create(m, restart, &ksp));
  CHK(sb200_ksp_set_operators(ksp, op_matmult, A, pc.type == "none" ? nullptr : pc_apply, &pc));
  CHK(sb200_ksp_set_tolerances(ksp, ksp_rtol, 1e-50, 1e5, ksp_max_it));
  Vec rhs, dx, xn;
  CHK(VecDuplicate(u, &rhs));
  CHK(VecDuplicate(u, &dx));
  CHK(VecDuplicate(u, &xn));
  std::vector<double> hx(m, 0.0), hdx(m), hxn(m), hF(m), hFn(m), hist;
  CHK(VecSetValuesHost(x, hx.data()));  // VecSet(x, 0.0)
  CHK(FormFunction(snes, x, r, &ac));
  CHK(VecGetValuesHost(r, hF.data()));
  double fn = norm2(hF);
  const double f0 = fn;
  hist.push_back(fn);
  int its = 0;
  std::vector<int> kits;
  while (its < snes_max_it && fn > fmax(snes_rtol * f0, snes_atol)) {
    MatStructure flag;
    CHK(FormJacobian(snes, x, &A, &P, &flag, &ac));  // about the state the last FormFunction(x) cached
    pc.n = m;
    CHK(pc.setup(P));
    for (int i = 0; i < m; i++) hr[i] = -hF[i];
    CHK(VecSetValuesHost(rhs, hr.data()));
    CHK(sb200_memset0(dx->d_array, (size_t)m * sizeof(double), nullptr));
    CHK(sb200_ksp_solve(ksp, rhs->d_array, dx->d_array, 0, nullptr));
    int k = 0, kreason = 0;
    CHK(sb200_ksp_get_result(ksp, &k, nullptr, nullptr, &kreason));
    kits.push_back(k);
    if (ksp_monitor) printf("    KSP iterations %d reason %d\n", k, kreason);
    CHK(VecGetValuesHost(dx, hdx.data()));
    double lam = 1.0, fnn = 0;
    while (true) {  // x <- x + lam dx, lam halved while the residual norm does not decrease
      for (int i = 0; i < m; i++) hxn[i] = hx[i] + lam * hdx[i];
      CHK(VecSetValuesHost(xn, hxn.data()));
      CHK(FormFunction(snes, xn, r, &ac));
      CHK(VecGetValuesHost(r, hFn.data()));
      fnn = norm2(hFn);
      if (fnn < fn || lam < 1e-3) break;
      lam *= 0.5;
    }
    hx.swap(hxn);
    hF.swap(hFn);
    CHK(VecSetValuesHost(x, hx.data()));
    fn = fnn;
    hist.push_back(fn);
    its++;
  }
  if (snes_monitor)
    for (size_t i = 0; i < hist.size(); i++) printf("  %d SNES Function norm %.12e\n", (int)i, hist[i]);
  const char* reason = hist.back() <= snes_atol ? "CONVERGED_FNORM_ABS" : (hist.back() <= snes_rtol * f0 ? "CONVERGED_FNORM_RELATIVE" : "DIVERGED_MAX_IT");
  double norm = 0, rnorm = 0;
  for (int i = 0; i < m; i++) {  // VecAXPY(x, -1, u); VecPointwiseDivide(x, x, u)
    const double e = hx[i] - hu[i];
    norm = fmax(norm, fabs(e));
    const double q = fabs(e / hu[i]);
    if (q == q) rnorm = fmax(rnorm, q);
  }
  printf("Number of nonlinear iterations = %d\n", its);
  printf("Reason for solver termination: %s\n", reason);
  printf("%-25s: abs = %8e   rel = %8e\n", "Norm of error", norm, rnorm);
  printf("KSP iterations per Newton step:");
  for (int k : kits) printf(" %d", k);
  printf("\n");
  o.warn_unused();

  sb200_ksp_destroy(ksp);
  CHK(SNESDestroy(snes));
  CHK(MatDestroy(A));
  CHK(MatDestroy(P));
  for (Vec v : {u, u2, x, r, rhs, dx, xn, ac.b}) CHK(VecDestroy(v));
  return 0;
}
