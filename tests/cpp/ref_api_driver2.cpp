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

int main() {
  if (test_chebd1()) return 1;
  if (test_schur(12)) return 1;
  return 0;
}
