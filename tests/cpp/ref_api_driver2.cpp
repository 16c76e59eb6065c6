// Second driver over the reference's own interface (include/sb200_reference_api.h): the 1-D operator MatCreateChebD1 /
// ChebD1Mult (chebyshev.c:8-85, used by cheb.c:47) and the Schur shell StokesMatMultSchur (stokes.C:318, 523-535).
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

// stands for KSPSolve(KSPSchurVelocity, rhs, sol) with -svel_ksp_type preonly -svel_pc_type none: sol = rhs
static PetscErrorCode identity_solve(void* calls, Vec rhs, Vec sol) {
  ++*(int*)calls;
  PetscInt n;
  VecGetSize(rhs, &n);
  const PetscScalar* a;
  PetscScalar* b;
  VecCUDAGetArrayRead(rhs, &a);
  VecCUDAGetArrayWrite(sol, &b);
  return sb200_memcpy_d2d(b, a, (size_t)n * sizeof(double), nullptr);
}


static int test_chebd1() {
  const double PI = 3.14159265358979323846;
  int m1 = 5;
  Vec u, b;
  Mat A;
  CHK(VecCreateSeqCUDA(PETSC_COMM_WORLD, m1, &u));
  CHK(VecDuplicate(u, &b));
  int dims[1] = {m1};
  CHK(MatCreateCheb(PETSC_COMM_WORLD, 1, 0, dims, FFTW_ESTIMATE, u, b, &A));
  std::vector<double> a(m1), r(m1);
  for (int i = 0; i < m1; i++) a[i] = exp(cos(i * PI / (m1 - 1)));  // cheb.c:68-70
  CHK(VecSetValuesHost(u, a.data()));
  CHK(MatMult(A, u, b));
  CHK(VecGetValuesHost(b, r.data()));
    {  // cheb.c creates the 1-D operator with MatCreateChebD1(comm, u, b, FFTW_ESTIMATE, &A): same numbers
      Mat A1;
      CHK(MatCreateChebD1(PETSC_COMM_WORLD, u, b, FFTW_ESTIMATE, &A1));
      CHK(MatMult(A1, u, b));
      std::vector<double> r1(m1);
      CHK(VecGetValuesHost(b, r1.data()));
      double diff = 0;
      for (int i = 0; i < m1; i++) diff = fmax(diff, fabs(r1[i] - r[i]));
      printf("chebD1 vs cheb max diff %.3e\n", diff);
      CHK(MatDestroy(A1));
      Vec one;
      CHK(VecCreateSeqCUDA(PETSC_COMM_WORLD, 1, &one));
      printf("chebD1 n=1 -> %d\n", MatCreateChebD1(PETSC_COMM_WORLD, one, one, FFTW_ESTIMATE, &A1));  // chebyshev.c:18
      CHK(VecDestroy(one));
    }
  CHK(MatDestroy(A));
  CHK(VecDestroy(u));
  CHK(VecDestroy(b));
  return 0;
}

static int test_schur(int n0) {
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
  Vec x;
  StokesCtxB200* ctx;
  CHK(StokesCreate(PETSC_COMM_SELF, &opt, &A, &x, &ctx));
  {  // the Schur shell (stokes.C:318, 523-535): with the inner solve replaced by the identity, S p = -PV (VP p)
    Mat MatVV, MatPV, MatVP, MatSchur;
    CHK(StokesGetShells(ctx, &MatVV, &MatPV, &MatVP, &MatSchur));
    PetscInt gp, gv;
    CHK(MatGetSize(MatPV, &gp, &gv));
    Vec p, sp, v, q;
    CHK(VecCreateSeqCUDA(PETSC_COMM_SELF, gp, &p));
    CHK(VecDuplicate(p, &sp));
    CHK(VecDuplicate(p, &q));
    CHK(VecCreateSeqCUDA(PETSC_COMM_SELF, gv, &v));
    std::vector<double> hp(gp), hs(gp), hq(gp);
    for (PetscInt i = 0; i < gp; i++) hp[i] = sin(0.37 * i) + 0.01 * (i % 7);
    CHK(VecSetValuesHost(p, hp.data()));
    printf("Schur without an inner solve -> %d\n", MatMult(MatSchur, p, sp));
    int calls = 0;
    CHK(StokesSetSchurVelocitySolve(ctx, identity_solve, &calls));
    CHK(MatMult(MatSchur, p, sp));
    CHK(MatMult(MatVP, p, v));
    CHK(MatMult(MatPV, v, q));
    CHK(VecGetValuesHost(sp, hs.data()));
    CHK(VecGetValuesHost(q, hq.data()));
    double diff = 0, big = 0;
    for (PetscInt i = 0; i < gp; i++) {
      diff = fmax(diff, fabs(hs[i] + hq[i]));
      big = fmax(big, fabs(hq[i]));
    }
    printf("Schur identity-solve calls %d  max |S p + PV VP p| / max |PV VP p| = %.3e\n", calls, diff / big);
    {  // StokesDivergence (stokes.C:570-595): without the Dirichlet data it is the PV shell; with it, the manufactured velocity
       // (u, v) = (sin cos, -cos sin) is divergence free up to the truncation error of 12 nodes per axis
      SNES snes;
      Vec U, U2, vex, d0, d1;
      CHK(SNESCreate(PETSC_COMM_SELF, &snes));
      CHK(SNESSetApplicationContext(snes, ctx));
      CHK(VecDuplicate(x, &U));
      CHK(VecDuplicate(x, &U2));
      CHK(StokesCreateExactSolution(snes, U, U2));
      PetscInt g;
      CHK(VecGetSize(U, &g));
      std::vector<double> hU(g), hv(gv), h0(gp), h1(gp);
      CHK(VecGetValuesHost(U, hU.data()));
      for (PetscInt q = 0; q < gp; q++)
        for (int k = 0; k < 3; k++) hv[q * 3 + k] = hU[q * 4 + k];
      CHK(VecCreateSeqCUDA(PETSC_COMM_SELF, gv, &vex));
      CHK(VecDuplicate(p, &d0));
      CHK(VecDuplicate(p, &d1));
      CHK(VecSetValuesHost(vex, hv.data()));
      CHK(StokesDivergence(ctx, PETSC_FALSE, vex, d0));
      CHK(MatMult(MatPV, vex, q));
      CHK(StokesDivergence(ctx, PETSC_TRUE, vex, d1));
      CHK(VecGetValuesHost(d0, h0.data()));
      CHK(VecGetValuesHost(q, hq.data()));
      CHK(VecGetValuesHost(d1, h1.data()));
      double same = 0, div_exact = 0, div_nobc = 0;
      for (PetscInt i = 0; i < gp; i++) {
        same = fmax(same, fabs(h0[i] - hq[i]));
        div_exact = fmax(div_exact, fabs(h1[i]));
        div_nobc = fmax(div_nobc, fabs(h0[i]));
      }
      printf("StokesDivergence: |div(no bc) - PV| = %.3e  |div(exact, with bc)| = %.3e  |div(exact, zero bc)| = %.3e\n", same, div_exact, div_nobc);
      for (Vec w : {U, U2, vex, d0, d1}) CHK(VecDestroy(w));
      CHK(SNESDestroy(snes));
    }
    CHK(VecDestroy(p));
    CHK(VecDestroy(sp));
    CHK(VecDestroy(q));
    CHK(VecDestroy(v));
  }
  CHK(StokesDestroy(ctx));
  CHK(MatDestroy(A));
  CHK(VecDestroy(x));
  return 0;
}

// ---- StokesPCApply0..3 (stokes.C:1714-1817) through the reference's names, everything on the device -------------------------------
// With the three inner solves run to 1e-12 the block-LU preconditioner is the inverse of StokesMatMult on the complement of the
// constant-pressure null space, and the triangular / diagonal variants satisfy their block identities.
static int maxdiff(Vec a, Vec b, PetscInt n, double* out, double* scale) {
  std::vector<double> ha(n), hb(n);
  CHK(VecGetValuesHost(a, ha.data()));
  CHK(VecGetValuesHost(b, hb.data()));
  *out = 0;
  *scale = 0;
  for (PetscInt i = 0; i < n; i++) {
    *out = fmax(*out, fabs(ha[i] - hb[i]));
    *scale = fmax(*scale, fabs(hb[i]));
  }
  return 0;
}

static int test_saddle(int n0) {
  StokesOptionsB200 opt;
  opt.numDims = 3;
  opt.dim[0] = opt.dim[1] = opt.dim[2] = n0;
  opt.exact = 2;
  opt.rheology = 1;  // power law: a variable eta, so the Jacobi "diagonal" 1/eta of the Schur solve matters
  opt.hardness = 1.0;
  opt.exponent = 2.0;
  opt.regularization = 0.5;
  opt.gamma0 = 1.0;
  Mat A, MatVV, MatPV, MatVP;
  Vec x, r, y[4], Ay, U, U2;
  StokesCtxB200* ctx;
  SNES snes;
  PC pc;
  CHK(StokesCreate(PETSC_COMM_SELF, &opt, &A, &x, &ctx));
  CHK(StokesGetShells(ctx, &MatVV, &MatPV, &MatVP, PETSC_NULL));
  PetscInt g, gp, gv;
  CHK(StokesGetSizes(ctx, PETSC_NULL, &g, &gp, &gv, PETSC_NULL));
  CHK(SNESCreate(PETSC_COMM_SELF, &snes));
  CHK(SNESSetApplicationContext(snes, ctx));
  CHK(PCCreate(PETSC_COMM_SELF, &pc));
  CHK(PCShellSetContext(pc, ctx));
  for (Vec* w : {&r, &y[0], &y[1], &y[2], &y[3], &Ay, &U, &U2}) CHK(VecDuplicate(x, w));
  CHK(StokesCreateExactSolution(snes, U, U2));
  CHK(StokesFunction(snes, U, r, ctx));  // linearisation state: eta, deta, strain of the manufactured solution
  PetscReal mn, mx;
  CHK(StokesGetEtaMinMax(ctx, &mn, &mx));
  // inner solves to 1e-12, no preconditioner on the velocity block (PCNONE), KSPSchurVelocity a full GMRES
  CHK(StokesSetVelocityPC(ctx, PETSC_NULL, PETSC_NULL, PETSC_NULL, PETSC_NULL));
  CHK(StokesSetInnerSolves(ctx, 1e-12, 2000, 1e-12, 500, PETSC_FALSE));
  CHK(StokesSetSchurVelocityTolerances(ctx, 1e-12, 2000));  // -svel_ksp_rtol / -svel_ksp_max_it: KSPSchurVelocity has its own prefix
  std::vector<double> hx(g), hz(g);
  srand(11);
  for (PetscInt i = 0; i < g; i++) hx[i] = rand() / (double)RAND_MAX - 0.5;
  CHK(VecSetValuesHost(x, hx.data()));
  CHK(StokesNullSpaceRemove(ctx, x));  // zero-mean pressure: the part StokesPCApply0 can recover
  CHK(VecGetValuesHost(x, hz.data()));
  double pmean = 0, vsame = 0;
  for (PetscInt q = 0; q < gp; q++) {
    pmean += hz[q * 4 + 3];
    for (int k = 0; k < 3; k++) vsame = fmax(vsame, fabs(hz[q * 4 + k] - hx[q * 4 + k]));
  }
  printf("saddle: eta in [%.3f, %.3f]  null-space removal: pressure mean %.1e  velocity change %.1e\n", mn, mx, fabs(pmean / gp), vsame);
  CHK(MatMult(A, x, r));                                      // r = J x
  PetscErrorCode (*apply[4])(PC, Vec, Vec) = {StokesPCApply0, StokesPCApply1, StokesPCApply2, StokesPCApply3};
  for (int t = 0; t < 4; t++) CHK(apply[t](pc, r, y[t]));
  double d, sc;
  CHK(StokesNullSpaceRemove(ctx, y[0]));
  CHK(maxdiff(y[0], x, g, &d, &sc));
  printf("saddle type 0 (block LU, exact inner solves): max |PC(J x) - x| / max |x| = %.3e\n", d / sc);
  // block identities of the other variants, checked with the full operator: J y = [VV y_v + VP y_p ; PV y_v]
  std::vector<double> hr(g), hy(g), hJ(g);
  CHK(VecGetValuesHost(r, hr.data()));
  {  // type 1 (upper): y_p = S^-1 r_p, VV y_v + VP y_p = r_v
    CHK(MatMult(A, y[1], Ay));
    CHK(VecGetValuesHost(Ay, hJ.data()));
    double e = 0, big = 0;
    for (PetscInt q = 0; q < gp; q++)
      for (int k = 0; k < 3; k++) {
        e = fmax(e, fabs(hJ[q * 4 + k] - hr[q * 4 + k]));
        big = fmax(big, fabs(hr[q * 4 + k]));
      }
    printf("saddle type 1 (upper): max |VV y_v + VP y_p - r_v| / max |r_v| = %.3e\n", e / big);
  }
  {  // types 2 (diagonal) and 3 (lower): VV y_v = r_v; their pressure parts repeat the Schur solves of types 1 and 0
    std::vector<double> h1(g), h2(g), h3(g), h0(g), hv(gv), hw(gv);
    CHK(VecGetValuesHost(y[1], h1.data()));
    CHK(VecGetValuesHost(y[2], h2.data()));
    CHK(VecGetValuesHost(y[3], h3.data()));
    Vec v, w;
    CHK(VecCreateSeqCUDA(PETSC_COMM_SELF, gv, &v));
    CHK(VecDuplicate(v, &w));
    double e23 = 0, ev = 0, big = 0, p12 = 0, pbig = 0;
    for (PetscInt q = 0; q < gp; q++) {
      for (int k = 0; k < 3; k++) {
        hv[q * 3 + k] = h2[q * 4 + k];
        e23 = fmax(e23, fabs(h2[q * 4 + k] - h3[q * 4 + k]));
      }
      p12 = fmax(p12, fabs(h1[q * 4 + 3] - h2[q * 4 + 3]));
      pbig = fmax(pbig, fabs(h2[q * 4 + 3]));
    }
    CHK(VecSetValuesHost(v, hv.data()));
    CHK(MatMult(MatVV, v, w));
    CHK(VecGetValuesHost(w, hw.data()));
    for (PetscInt q = 0; q < gp; q++)
      for (int k = 0; k < 3; k++) {
        ev = fmax(ev, fabs(hw[q * 3 + k] - hr[q * 4 + k]));
        big = fmax(big, fabs(hr[q * 4 + k]));
      }
    printf("saddle type 2 (diagonal): max |VV y_v - r_v| / max |r_v| = %.3e   |y_v(2) - y_v(3)| = %.3e   |y_p(1) - y_p(2)| / max = %.3e\n", ev / big, e23,
           p12 / pbig);
    // type 3 (lower): PV y_v + S y_p = r_p up to a constant; S y_p = J-row: use type 0, whose pressure part is type 3's
    CHK(apply[0](pc, r, y[0]));
    CHK(VecGetValuesHost(y[0], h0.data()));
    double p03 = 0;
    for (PetscInt q = 0; q < gp; q++) p03 = fmax(p03, fabs(h0[q * 4 + 3] - h3[q * 4 + 3]));
    printf("saddle type 3 (lower): |y_p(3) - y_p(0)| = %.3e\n", p03);
    CHK(VecDestroy(v));
    CHK(VecDestroy(w));
  }
  PetscInt iv, is;
  CHK(StokesGetInnerIterations(ctx, &iv, &is));
  printf("saddle inner iterations: velocity %d  schur %d\n", iv, is);
  Vec small;
  CHK(VecCreateSeqCUDA(PETSC_COMM_SELF, gp, &small));
  printf("saddle wrong-size Vec -> %d   same Vec twice -> %d\n", StokesPCApply0(pc, small, y[0]), StokesPCApply0(pc, r, r));
  CHK(VecDestroy(small));
  for (Vec w : {r, y[0], y[1], y[2], y[3], Ay, U, U2, x}) CHK(VecDestroy(w));
  CHK(PCDestroy(pc));
  CHK(SNESDestroy(snes));
  CHK(StokesDestroy(ctx));
  CHK(MatDestroy(A));
  return 0;
}

int main() {
  if (test_chebd1()) return 1;
  if (test_schur(12)) return 1;
  if (test_saddle(7)) return 1;
  return 0;
}
