// The reference's operator interface on the B200 path (include/sb200_reference_api.h): same
// names, same argument meaning, same error behaviour; each function cites what it replaces.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/sb200_reference_api.h"
#include "../../include/spectral_b200.h"

#define CHK(expr)            \
  do {                       \
    PetscErrorCode _e = (expr); \
    if (_e) return _e;       \
  } while (0)

// Shared by FormJacobian and StokesPCSetUp0: the first assembly allocates the device CSR of the SeqAIJ matrix and writes
// pattern + values, later ones refresh the values alone.
template <class Ctx>
static PetscErrorCode assemble_aij(Mat P, Ctx* h, int (*sizes)(Ctx*, long long*, long long*), int (*csr)(Ctx*, int*, int*, double*, void*)) {
  if (!P || !P->is_aij) return SB200_ERR_ARG;
  long long nrows = 0, nnz = 0;
  CHK(sizes(h, &nrows, &nnz));
  if (P->m != (PetscInt)nrows || P->n != (PetscInt)nrows) return SB200_ERR_USER;  // MatSetValues would fail on the row ids
  if (!P->d_vals) {
    CHK(sb200_malloc((void**)&P->d_rowptr, (size_t)(nrows + 1) * sizeof(PetscInt)));
    CHK(sb200_malloc((void**)&P->d_colidx, (size_t)nnz * sizeof(PetscInt)));
    CHK(sb200_malloc((void**)&P->d_vals, (size_t)nnz * sizeof(PetscScalar)));
    P->nz = (PetscInt)nnz;
    return csr(h, P->d_rowptr, P->d_colidx, P->d_vals, nullptr);
  }
  return csr(h, nullptr, nullptr, P->d_vals, nullptr);
}

extern "C" {

// ---- chebyshev.c ---------------------------------------------------------------------------
// MatCreateCheb (chebyshev.c:89-138): N x N shell with MULT and DESTROY operations (:133-135).
PetscErrorCode MatCreateCheb(MPI_Comm comm, int rank, int tr, int* dims, unsigned, Vec vx, Vec vy, Mat* A) {
  PetscInt n = 0;
  CHK(VecGetSize(vx, &n));
  (void)vy;
  sb200_cheb* c = nullptr;
  CHK(sb200_cheb_create(rank, tr, dims, n, &c));
  CHK(MatCreateShell(comm, n, n, n, n, c, A));
  CHK(MatShellSetOperation(*A, MATOP_MULT, (void (*)(void))ChebMult));
  CHK(MatShellSetOperation(*A, MATOP_DESTROY, (void (*)(void))ChebDestroy));
  return 0;
}

// ChebMult (chebyshev.c:142-199)
PetscErrorCode ChebMult(Mat A, Vec vx, Vec vy) {
  sb200_cheb* c = nullptr;
  const PetscScalar* x;
  PetscScalar* y;
  CHK(MatShellGetContext(A, (void**)&c));
  CHK(VecCUDAGetArrayRead(vx, &x));
  CHK(VecCUDAGetArrayWrite(vy, &y));
  CHK(sb200_cheb_apply(c, x, y, nullptr));
  CHK(VecCUDARestoreArrayRead(vx, &x));
  CHK(VecCUDARestoreArrayWrite(vy, &y));
  return 0;
}

// ChebDestroy (chebyshev.c:223-235)
PetscErrorCode ChebDestroy(Mat A) {
  sb200_cheb* c = nullptr;
  CHK(MatShellGetContext(A, (void**)&c));
  return sb200_cheb_destroy(c);
}

// MatCreateChebD1 / ChebD1Mult / ChebD1Destroy (chebyshev.c:8-85): the 1-D operator; "n = %d but must be >= 2" and the
// vx / vy size check are the reference's (chebyshev.c:18)
PetscErrorCode MatCreateChebD1(MPI_Comm comm, Vec vx, Vec vy, unsigned flag, Mat* A) {
  PetscInt n = 0, ny = 0;
  CHK(VecGetSize(vx, &n));
  CHK(VecGetSize(vy, &ny));
  if (n != ny || n < 2) return SB200_ERR_USER;
  int dims[1] = {(int)n};
  CHK(MatCreateCheb(comm, 1, 0, dims, flag, vx, vy, A));
  CHK(MatShellSetOperation(*A, MATOP_MULT, (void (*)(void))ChebD1Mult));
  CHK(MatShellSetOperation(*A, MATOP_DESTROY, (void (*)(void))ChebD1Destroy));
  return 0;
}
PetscErrorCode ChebD1Mult(Mat A, Vec vx, Vec vy) { return ChebMult(A, vx, vy); }
PetscErrorCode ChebD1Destroy(Mat A) { return ChebDestroy(A); }

// ---- elliptic.C ------------------------------------------------------------------------------
struct MatEllipticB200 {  // MatElliptic (elliptic.C:78-86): the device state lives behind `e`
  sb200_elliptic* e;
  int d;
  std::vector<int> dim;
  long long m, g, nd;
};

PetscErrorCode DirichletBdy(int, double*, double*, BdyCond* bc) {  // elliptic.C:470-477
  bc->type = BDY_DIRICHLET;
  bc->value = 0.0;
  return 0;
}

PetscErrorCode MatCreate_Elliptic(MPI_Comm comm, int d, int* dim, unsigned, BdyFunc bf, Vec* vG, Mat* A) {
  // elliptic.C:250-293.  SetupBC (elliptic.C:372-466) only supports Dirichlet ("Neumann not implemented", :406)
  if (bf) {
    BdyCond bc;
    double x0[10] = {0}, n0[10] = {0};
    CHK(bf(d, x0, n0, &bc));
    if (bc.type != BDY_DIRICHLET) return SB200_ERR_SUP;
  }
  MatEllipticB200* c = new MatEllipticB200();
  PetscErrorCode rc = sb200_elliptic_create(d, dim, &c->e);
  if (rc) {
    delete c;
    return rc;
  }
  c->d = d;
  c->dim.assign(dim, dim + d);
  sb200_elliptic_sizes(c->e, &c->m, &c->g, &c->nd);
  CHK(VecCreateSeqCUDA(comm, (PetscInt)c->g, vG));
  CHK(MatCreateShell(comm, (PetscInt)c->g, (PetscInt)c->g, (PetscInt)c->g, (PetscInt)c->g, c, A));
  CHK(MatShellSetOperation(*A, MATOP_MULT, (void (*)(void))MatMult_Elliptic));
  CHK(MatShellSetOperation(*A, MATOP_DESTROY, (void (*)(void))MatDestroy_Elliptic));
  return 0;
}

PetscErrorCode MatMult_Elliptic(Mat A, Vec U, Vec V) {  // elliptic.C:297-339
  MatEllipticB200* c = nullptr;
  CHK(MatShellGetContext(A, (void**)&c));
  const PetscScalar* u;
  PetscScalar* v;
  CHK(VecCUDAGetArrayRead(U, &u));
  CHK(VecCUDAGetArrayWrite(V, &v));
  CHK(sb200_elliptic_matmult(c->e, u, v, nullptr));
  CHK(VecCUDARestoreArrayRead(U, &u));
  CHK(VecCUDARestoreArrayWrite(V, &v));
  return 0;
}

PetscErrorCode MatDestroy_Elliptic(Mat A) {  // elliptic.C:343-368
  MatEllipticB200* c = nullptr;
  CHK(MatShellGetContext(A, (void**)&c));
  PetscErrorCode rc = sb200_elliptic_destroy(c->e);
  delete c;
  return rc;
}

PetscErrorCode FormFunction(SNES, Vec U, Vec rhs, void* void_ac) {  // elliptic.C:481-533
  AppCtx* ac = (AppCtx*)void_ac;
  MatEllipticB200* c = nullptr;
  CHK(MatShellGetContext(ac->A, (void**)&c));
  CHK(sb200_elliptic_set_params(c->e, ac->gamma, ac->exponent));
  const PetscScalar *u, *b;
  PetscScalar* r;
  CHK(VecCUDAGetArrayRead(ac->b, &b));
  CHK(sb200_elliptic_set_rhs(c->e, b, nullptr));
  CHK(VecCUDAGetArrayRead(U, &u));
  CHK(VecCUDAGetArrayWrite(rhs, &r));
  CHK(sb200_elliptic_function(c->e, u, r, nullptr));
  return 0;
}

PetscErrorCode FormJacobian(SNES, Vec, Mat* A, Mat* P, MatStructure* flag, void*) {  // elliptic.C:537-590
  MatEllipticB200* c = nullptr;
  CHK(MatShellGetContext(*A, (void**)&c));
  CHK(assemble_aij<sb200_elliptic>(*P, c->e, sb200_elliptic_jacobian_sizes, sb200_elliptic_jacobian_csr));
  if (flag) *flag = SAME_NONZERO_PATTERN;  // :588
  return 0;
}

PetscErrorCode CreateExactSolution(SNES snes, Vec u, Vec u2, PetscReal cos_scale) {  // elliptic.C:594-677
  AppCtx* ac = nullptr;
  MatEllipticB200* c = nullptr;
  CHK(SNESGetApplicationContext(snes, (void**)&ac));
  CHK(MatShellGetContext(ac->A, (void**)&c));
  std::vector<double> hu((size_t)c->g), hu2((size_t)c->g), hd((size_t)c->nd);
  CHK(sb200_elliptic_exact_solution(c->d, c->dim.data(), (int)ac->exact, cos_scale, ac->gamma, ac->exponent, hu.data(), hu2.data(), hd.data()));
  CHK(VecSetValuesHost(u, hu.data()));
  CHK(VecSetValuesHost(u2, hu2.data()));
  CHK(VecSetValuesHost(ac->b, hu2.data()));  // VecCopy(u2, ac->b) elliptic.C:674
  // scatterLD into c->dirichlet (elliptic.C:672)
  void* dd = nullptr;
  CHK(sb200_malloc(&dd, hd.size() * sizeof(double) + 8));
  CHK(sb200_memcpy_h2d(dd, hd.data(), hd.size() * sizeof(double), nullptr));
  CHK(sb200_elliptic_set_dirichlet(c->e, (const double*)dd, nullptr));
  CHK(sb200_stream_sync(nullptr));
  CHK(sb200_free(dd));
  return 0;
}

// ---- stokes.C ----------------------------------------------------------------------------------
struct StokesCtxB200 {
  sb200_stokes* s;
  StokesOptionsB200 opt;
  long long m, g, gp, gv, dv;
  Mat MatVV, MatPV, MatVP, MatSchur, MatVVPC;
  StokesVelocitySolve svel = nullptr;  // KSPSolve(KSPSchurVelocity, ., .) (stokes.C:531)
  void* svel_ksp = nullptr;
  // StokesPCApply0..3: one device-resident composition per saddle type, created at its first application
  sb200_saddle* saddle[4] = {nullptr, nullptr, nullptr, nullptr};
  StokesVelocitySolve vel_pc = nullptr, svel_pc = nullptr;  // PCApply on MatVVPC for KSPVelocity / KSPSchurVelocity
  void* vel_pc_ctx = nullptr;
  void* svel_pc_ctx = nullptr;
  std::vector<double> h_force;  // host copy of c->force (stokes.C:1001), what StokesStateView writes as vel_force / div_force
  double vel_rtol = 1e-5, schur_rtol = 1e-5, svel_rtol = 1e-5;  // KSP defaults
  int vel_max_it = 10000, schur_max_it = 10000, svel_max_it = 10000, svel_preonly = 0;
};

static PetscErrorCode stokes_fill(StokesCtxB200* c, std::vector<double>* U, std::vector<double>* U2, std::vector<double>* D) {
  // StokesExact0..2 (stokes.C:1948-2012) at every node in walk order
  return sb200_stokes_exact_solution((int)c->opt.numDims, c->opt.dim, (int)c->opt.exact, U ? U->data() : nullptr, U2 ? U2->data() : nullptr,
                                     D ? D->data() : nullptr);
}

PetscErrorCode StokesCreate(MPI_Comm comm, const StokesOptionsB200* opt, Mat* A, Vec* x, StokesCtxB200** ctx) {
  // stokes.C:257-345 with StokesProcessOptions' switches (stokes.C:434-493)
  if (opt->exact < 0 || opt->exact > 3) return SB200_ERR_SUP;      // stokes.C:452; StokesExact3 itself refuses d != 2 (stokes.C:2021)
  if (opt->rheology < 0 || opt->rheology > 1) return SB200_ERR_SUP;  // stokes.C:492
  StokesCtxB200* c = new StokesCtxB200();
  c->opt = *opt;
  int dim[3] = {opt->dim[0], opt->dim[1], opt->dim[2]};
  PetscErrorCode rc = sb200_stokes_create(opt->numDims, dim, &c->s);
  if (rc) {
    delete c;
    return rc;
  }
  sb200_stokes_sizes(c->s, &c->m, &c->g, &c->gp, &c->gv, &c->dv);
  CHK(sb200_stokes_set_rheology(c->s, opt->rheology, opt->hardness, opt->exponent, opt->regularization, opt->gamma0));
  // Dirichlet values = exact solution on the boundary (StokesDirichlet, stokes.C:2039-2050)
  std::vector<double> D((size_t)c->dv);
  rc = stokes_fill(c, nullptr, nullptr, &D);
  if (rc) {  // e.g. -exact 3 in three dimensions
    sb200_stokes_destroy(c->s);
    delete c;
    return rc;
  }
  void* dd = nullptr;
  CHK(sb200_malloc(&dd, D.size() * sizeof(double) + 8));
  CHK(sb200_memcpy_h2d(dd, D.data(), D.size() * sizeof(double), nullptr));
  CHK(sb200_stokes_set_dirichlet(c->s, (const double*)dd, nullptr));
  CHK(sb200_stream_sync(nullptr));
  CHK(sb200_free(dd));
  CHK(VecCreateSeqCUDA(comm, (PetscInt)c->g, x));
  CHK(MatCreateShell(comm, (PetscInt)c->g, (PetscInt)c->g, 0, 0, c, A));
  CHK(MatShellSetOperation(*A, MATOP_MULT, (void (*)(void))StokesMatMult));  // stokes.C:309
  CHK(MatCreateShell(comm, (PetscInt)c->gp, (PetscInt)c->gp, 0, 0, c, &c->MatSchur));
  CHK(MatShellSetOperation(c->MatSchur, MATOP_MULT, (void (*)(void))StokesMatMultSchur));                 // :318
  CHK(MatShellSetOperation(c->MatSchur, MATOP_GET_DIAGONAL, (void (*)(void))StokesMatGetDiagonalSchur));  // :319
  CHK(MatCreateShell(comm, (PetscInt)c->gp, (PetscInt)c->gv, 0, 0, c, &c->MatPV));
  CHK(MatShellSetOperation(c->MatPV, MATOP_MULT, (void (*)(void))StokesMatMultPV));  // :321
  CHK(MatCreateShell(comm, (PetscInt)c->gv, (PetscInt)c->gp, 0, 0, c, &c->MatVP));
  CHK(MatShellSetOperation(c->MatVP, MATOP_MULT, (void (*)(void))StokesMatMultVP));  // :323
  CHK(MatCreateShell(comm, (PetscInt)c->gv, (PetscInt)c->gv, 0, 0, c, &c->MatVV));
  CHK(MatShellSetOperation(c->MatVV, MATOP_MULT, (void (*)(void))StokesMatMultVV));  // :325
  CHK(MatCreateSeqAIJ(comm, (PetscInt)c->gv, (PetscInt)c->gv, 1 + 2 * opt->numDims, PETSC_NULL, &c->MatVVPC));  // :326
  *ctx = c;
  return 0;
}

PetscErrorCode StokesDestroy(StokesCtxB200* c) {  // stokes.C:348-388
  if (!c) return 0;
  MatDestroy(c->MatVVPC);
  MatDestroy(c->MatSchur);
  MatDestroy(c->MatPV);
  MatDestroy(c->MatVP);
  MatDestroy(c->MatVV);
  for (sb200_saddle* p : c->saddle) sb200_saddle_destroy(p);
  PetscErrorCode rc = sb200_stokes_destroy(c->s);
  delete c;
  return rc;
}

#define STOKES_SHELL(NAME, CALL)                                 \
  PetscErrorCode NAME(Mat A, Vec xG, Vec yG) {                   \
    StokesCtxB200* c = nullptr;                                  \
    CHK(MatShellGetContext(A, (void**)&c));                      \
    const PetscScalar* x;                                        \
    PetscScalar* y;                                              \
    CHK(VecCUDAGetArrayRead(xG, &x));                            \
    CHK(VecCUDAGetArrayWrite(yG, &y));                           \
    return CALL(c->s, x, y, nullptr);                            \
  }
STOKES_SHELL(StokesMatMult, sb200_stokes_matmult)        // stokes.C:499-519
STOKES_SHELL(StokesMatMultVV, sb200_stokes_matmult_vv)   // stokes.C:623-676
STOKES_SHELL(StokesMatMultPV, sb200_stokes_matmult_pv)   // stokes.C:557-566
STOKES_SHELL(StokesMatMultVP, sb200_stokes_matmult_vp)   // stokes.C:599-619

PetscErrorCode StokesMatGetDiagonalSchur(Mat S, Vec y) {  // stokes.C:542-553
  StokesCtxB200* c = nullptr;
  CHK(MatShellGetContext(S, (void**)&c));
  PetscScalar* a;
  CHK(VecCUDAGetArrayWrite(y, &a));
  return sb200_stokes_get_diagonal_schur(c->s, a, nullptr);
}

PetscErrorCode StokesDivergence(StokesCtxB200* c, PetscTruth withDirichlet, Vec xG, Vec yG) {  // stokes.C:570-595
  const PetscScalar* x;
  PetscScalar* y;
  if (xG->n != (PetscInt)c->gv || yG->n != (PetscInt)c->gp) return SB200_ERR_USER;
  CHK(VecCUDAGetArrayRead(xG, &x));
  CHK(VecCUDAGetArrayWrite(yG, &y));
  return sb200_stokes_divergence(c->s, withDirichlet ? 1 : 0, x, y, nullptr);
}

// StokesExact0..3 (stokes.C:1948-2034) and StokesDirichlet (stokes.C:2039-2050): the manufactured solutions as the per-point host
// callbacks StokesOptions stores; StokesCreateExactSolution evaluates the same expressions at every node.
PetscErrorCode StokesExact0(PetscInt d, PetscReal* coord, PetscReal* value, PetscReal* rhs, void*) { return sb200_stokes_exact_eval(0, d, coord, value, rhs); }
PetscErrorCode StokesExact1(PetscInt d, PetscReal* coord, PetscReal* value, PetscReal* rhs, void*) { return sb200_stokes_exact_eval(1, d, coord, value, rhs); }
PetscErrorCode StokesExact2(PetscInt d, PetscReal* coord, PetscReal* value, PetscReal* rhs, void*) { return sb200_stokes_exact_eval(2, d, coord, value, rhs); }
PetscErrorCode StokesExact3(PetscInt d, PetscReal* coord, PetscReal* value, PetscReal* rhs, void*) { return sb200_stokes_exact_eval(3, d, coord, value, rhs); }
PetscErrorCode StokesDirichlet(PetscInt d, PetscReal* coord, PetscReal*, StokesBdyType* type, PetscReal* value, void* void_ctx) {
  StokesExactBoundaryCtx* ctx = (StokesExactBoundaryCtx*)void_ctx;
  *type = DIRICHLET;
  PetscReal full[4];  // the exact solution has d + 1 values; the boundary condition wants what the reference's callee writes
  PetscErrorCode rc = ctx->exact(d, coord, full, PETSC_NULL, ctx->exactCtx);
  for (PetscInt k = 0; !rc && k <= d; k++) value[k] = full[k];
  return rc;
}

PetscErrorCode StokesRheologyLinear(PetscInt, PetscReal, PetscReal* eta, PetscReal* deta, void*) {  // stokes.C:1920-1926
  *eta = 1.0;
  *deta = 0.0;
  return 0;
}

PetscErrorCode StokesRheologyPower(PetscInt, PetscReal gamma, PetscReal* eta, PetscReal* deta, void* ctx) {  // stokes.C:1930-1944
  const StokesOptionsB200* o = (const StokesOptionsB200*)ctx;
  const double n = o->exponent, p = (1.0 - n) / (2.0 * n), base = o->regularization + gamma / o->gamma0;
  *eta = o->hardness * pow(base, p);
  *deta = fabs(n) > 1.0e-5 ? o->hardness * p / o->gamma0 * pow(base, p - 1.0) : 0.0;  // "Avoid a singularity for the special case"
  return 0;
}

PetscErrorCode polyInterp(const PetscInt n, const PetscReal* x, PetscScalar* w, const PetscReal x0, const PetscReal x1, PetscScalar* f0, PetscScalar* f1) {
  // util.C:129-144.  Columns 0 / 1 of the width-4 table hold the values; each level combines neighbours in place.
  if (n < 1) return SB200_ERR_USER;
  for (PetscInt lvl = 1; lvl < n; lvl++)
    for (PetscInt i = 0; i < n - lvl; i++) {
      const double den = x[i] - x[i + lvl];
      w[4 * i] = ((x0 - x[i + lvl]) * w[4 * i] + (x[i] - x0) * w[4 * (i + 1)]) / den;
      w[4 * i + 1] = ((x1 - x[i + lvl]) * w[4 * i + 1] + (x[i] - x1) * w[4 * (i + 1) + 1]) / den;
    }
  *f0 = w[0];
  *f1 = w[1];
  return 0;
}

PetscErrorCode StokesJacobian(SNES, Vec, Mat*, Mat*, MatStructure* flag, void*) {  // stokes.C:761-769
  *flag = DIFFERENT_NONZERO_PATTERN;  // "The nonlinear term has already been fixed up by StokesFunction() so we do nothing here."
  return 0;
}

PetscErrorCode StokesSetSchurVelocitySolve(StokesCtxB200* c, StokesVelocitySolve solve, void* ksp) {
  c->svel = solve;
  c->svel_ksp = ksp;
  return 0;
}

// the C ABI hands the inner solve raw device pointers; wrap them as Vecs for the PETSc-side callback
static int schur_velocity_trampoline(void* vctx, const double* d_rhs, double* d_sol, void*) {
  StokesCtxB200* c = (StokesCtxB200*)vctx;
  Vec rhs = nullptr, sol = nullptr;
  PetscErrorCode rc = VecCreateSeqCUDAWithArray(PETSC_COMM_SELF, (PetscInt)c->gv, (double*)d_rhs, &rhs);
  if (!rc) rc = VecCreateSeqCUDAWithArray(PETSC_COMM_SELF, (PetscInt)c->gv, d_sol, &sol);
  if (!rc) rc = c->svel(c->svel_ksp, rhs, sol);
  if (rhs) VecDestroy(rhs);
  if (sol) VecDestroy(sol);
  return rc;
}

PetscErrorCode StokesMatMultSchur(Mat S, Vec xG, Vec yG) {  // stokes.C:523-535
  StokesCtxB200* c = nullptr;
  CHK(MatShellGetContext(S, (void**)&c));
  if (!c->svel) return SB200_ERR_ARG;
  const PetscScalar* x;
  PetscScalar* y;
  CHK(VecCUDAGetArrayRead(xG, &x));
  CHK(VecCUDAGetArrayWrite(yG, &y));
  return sb200_stokes_matmult_schur(c->s, x, y, schur_velocity_trampoline, c, nullptr);
}

PetscErrorCode StokesFunction(SNES, Vec xG, Vec yG, void* ctx) {  // stokes.C:680-758
  StokesCtxB200* c = (StokesCtxB200*)ctx;
  const PetscScalar* x;
  PetscScalar* y;
  CHK(VecCUDAGetArrayRead(xG, &x));
  CHK(VecCUDAGetArrayWrite(yG, &y));
  return sb200_stokes_function(c->s, x, y, nullptr);
}

PetscErrorCode StokesCreateExactSolution(SNES snes, Vec U, Vec U2) {  // stokes.C:942-1003
  StokesCtxB200* c = nullptr;
  CHK(SNESGetApplicationContext(snes, (void**)&c));
  std::vector<double> hu((size_t)c->g), hu2((size_t)c->g);
  stokes_fill(c, &hu, &hu2, nullptr);
  CHK(VecSetValuesHost(U, hu.data()));
  CHK(VecSetValuesHost(U2, hu2.data()));
  c->h_force = hu2;
  const PetscScalar* f;
  CHK(VecCUDAGetArrayRead(U2, &f));
  return sb200_stokes_set_force(c->s, f, nullptr);  // VecCopy(U2, c->force) stokes.C:1001
}

PetscErrorCode StokesGetShells(StokesCtxB200* c, Mat* MatVV, Mat* MatPV, Mat* MatVP, Mat* MatSchur) {
  if (MatVV) *MatVV = c->MatVV;
  if (MatPV) *MatPV = c->MatPV;
  if (MatVP) *MatVP = c->MatVP;
  if (MatSchur) *MatSchur = c->MatSchur;
  return 0;
}

PetscErrorCode StokesGetPCMatrix(StokesCtxB200* c, Mat* MatVVPC) {
  *MatVVPC = c->MatVVPC;
  return 0;
}

PetscErrorCode StokesPCSetUp0(PC pc) {  // stokes.C:1160-1240 (the KSPSetOperators calls at :1232-1234 stay with the caller's KSPs)
  StokesCtxB200* c = nullptr;
  CHK(PCShellGetContext(pc, (void**)&c));
  return assemble_aij<sb200_stokes>(c->MatVVPC, c->s, sb200_stokes_pc_velocity_sizes, sb200_stokes_pc_velocity_csr);
}

PetscErrorCode StokesPressureReduceOrder(Vec pL, StokesCtxB200* c) {  // stokes.C:1029-1080
  PetscScalar* p;
  if (pL->n != (PetscInt)c->m) return SB200_ERR_USER;
  CHK(VecCUDAGetArrayWrite(pL, &p));
  return sb200_stokes_pressure_reduce_order(c->s, p, nullptr);
}

PetscErrorCode StokesGetEtaMinMax(StokesCtxB200* c, PetscReal* minEta, PetscReal* maxEta) {  // stokes.C:731-734
  return sb200_stokes_eta_minmax(c->s, minEta, maxEta, nullptr);
}

PetscErrorCode StokesGetSizes(StokesCtxB200* c, PetscInt* m, PetscInt* g, PetscInt* gp, PetscInt* gv, PetscInt* dv) {  // stokes.C:891
  if (m) *m = (PetscInt)c->m;
  if (g) *g = (PetscInt)c->g;
  if (gp) *gp = (PetscInt)c->gp;
  if (gv) *gv = (PetscInt)c->gv;
  if (dv) *dv = (PetscInt)c->dv;
  return 0;
}

PetscErrorCode StokesGetState(StokesCtxB200* c, PetscInt which, Vec out) {  // the fields StokesStateView writes (stokes.C:1868-1885)
  const long long need = which < 2 ? c->m : c->m * c->opt.numDims;
  if (out->n != (PetscInt)need) return SB200_ERR_USER;
  PetscScalar* a;
  CHK(VecCUDAGetArrayWrite(out, &a));
  return sb200_stokes_get_state(c->s, (int)which, a, nullptr);
}

PetscErrorCode StokesSetContinuation(StokesCtxB200* c, PetscReal exponent, PetscReal regularization) {
  c->opt.exponent = exponent;
  c->opt.regularization = regularization;
  return sb200_stokes_set_rheology(c->s, c->opt.rheology, c->opt.hardness, exponent, regularization, c->opt.gamma0);
}

// ---- the saddle-point PCShells (stokes.C:171-185 selects one by -pc_saddle_type) -----------------------------------------------
PetscErrorCode StokesSetVelocityPC(StokesCtxB200* c, StokesVelocitySolve vel_pc, void* vel_ctx, StokesVelocitySolve svel_pc, void* svel_ctx) {
  c->vel_pc = vel_pc;
  c->vel_pc_ctx = vel_ctx;
  c->svel_pc = svel_pc;
  c->svel_pc_ctx = svel_ctx;
  return 0;
}

PetscErrorCode StokesSetInnerSolves(StokesCtxB200* c, PetscReal vel_rtol, PetscInt vel_max_it, PetscReal schur_rtol, PetscInt schur_max_it, PetscTruth svel_preonly) {
  if (vel_rtol < 0 || schur_rtol < 0 || vel_max_it < 0 || schur_max_it < 0) return SB200_ERR_USER;
  c->vel_rtol = vel_rtol;
  c->vel_max_it = vel_max_it;
  c->schur_rtol = schur_rtol;
  c->schur_max_it = schur_max_it;
  c->svel_preonly = svel_preonly ? 1 : 0;
  return 0;
}

PetscErrorCode StokesSetSchurVelocityTolerances(StokesCtxB200* c, PetscReal svel_rtol, PetscInt svel_max_it) {
  if (svel_rtol < 0 || svel_max_it < 0) return SB200_ERR_USER;
  c->svel_rtol = svel_rtol;
  c->svel_max_it = svel_max_it;
  return 0;
}

// the C ABI hands a preconditioner raw device pointers; wrap them as Vecs for the PETSc-side PCApply
static int vec_trampoline(StokesCtxB200* c, StokesVelocitySolve f, void* fctx, const double* d_r, double* d_z) {
  Vec r = nullptr, z = nullptr;
  PetscErrorCode rc = VecCreateSeqCUDAWithArray(PETSC_COMM_SELF, (PetscInt)c->gv, (double*)d_r, &r);
  if (!rc) rc = VecCreateSeqCUDAWithArray(PETSC_COMM_SELF, (PetscInt)c->gv, d_z, &z);
  if (!rc) rc = f(fctx, r, z);
  if (r) VecDestroy(r);
  if (z) VecDestroy(z);
  return rc;
}
static int vel_pc_trampoline(void* vctx, const double* d_r, double* d_z, void*) {
  StokesCtxB200* c = (StokesCtxB200*)vctx;
  return vec_trampoline(c, c->vel_pc, c->vel_pc_ctx, d_r, d_z);
}
static int svel_pc_trampoline(void* vctx, const double* d_r, double* d_z, void*) {
  StokesCtxB200* c = (StokesCtxB200*)vctx;
  return vec_trampoline(c, c->svel_pc, c->svel_pc_ctx, d_r, d_z);
}

static PetscErrorCode stokes_saddle(StokesCtxB200* c, int type, sb200_saddle** out) {
  if (!c->saddle[type]) CHK(sb200_saddle_create(c->s, type, &c->saddle[type]));
  sb200_saddle* p = c->saddle[type];
  CHK(sb200_saddle_set_velocity_pc(p, c->vel_pc ? vel_pc_trampoline : nullptr, c, c->svel_pc ? svel_pc_trampoline : nullptr, c, 0));
  CHK(sb200_saddle_set_inner(p, c->vel_rtol, c->vel_max_it, c->schur_rtol, c->schur_max_it, c->svel_preonly));
  CHK(sb200_saddle_set_svel(p, c->svel_rtol, c->svel_max_it));
  *out = p;
  return 0;
}

static PetscErrorCode stokes_pc_apply(PC pc, int type, Vec x, Vec y) {
  StokesCtxB200* c = nullptr;
  CHK(PCShellGetContext(pc, (void**)&c));
  if (!c || x->n != (PetscInt)c->g || y->n != (PetscInt)c->g) return SB200_ERR_USER;
  sb200_saddle* p = nullptr;
  CHK(stokes_saddle(c, type, &p));
  const PetscScalar* a;
  PetscScalar* b;
  CHK(VecCUDAGetArrayRead(x, &a));
  CHK(VecCUDAGetArrayWrite(y, &b));
  return sb200_saddle_apply(p, a, b, nullptr);
}
PetscErrorCode StokesPCApply0(PC pc, Vec x, Vec y) { return stokes_pc_apply(pc, 0, x, y); }  // stokes.C:1714-1742  block LU
PetscErrorCode StokesPCApply1(PC pc, Vec x, Vec y) { return stokes_pc_apply(pc, 1, x, y); }  // stokes.C:1747-1768  upper triangular
PetscErrorCode StokesPCApply2(PC pc, Vec x, Vec y) { return stokes_pc_apply(pc, 2, x, y); }  // stokes.C:1773-1792  diagonal
PetscErrorCode StokesPCApply3(PC pc, Vec x, Vec y) { return stokes_pc_apply(pc, 3, x, y); }  // stokes.C:1797-1817  lower triangular

PetscErrorCode StokesNullSpaceRemove(StokesCtxB200* c, Vec x) {  // MatNullSpaceRemove with the vector of stokes.C:1013-1020
  if (x->n != (PetscInt)c->g) return SB200_ERR_USER;
  sb200_saddle* p = nullptr;
  CHK(stokes_saddle(c, 0, &p));
  PetscScalar* a;
  CHK(VecCUDAGetArrayWrite(x, &a));
  return sb200_saddle_remove_constant_pressure(p, a, nullptr);
}

PetscErrorCode StokesGetInnerIterations(StokesCtxB200* c, PetscInt* velocity, PetscInt* schur) {
  long long v = 0, s = 0;
  for (sb200_saddle* p : c->saddle)
    if (p) {
      long long a = 0, b = 0;
      CHK(sb200_saddle_get_inner_its(p, &a, &b));
      v += a;
      s += b;
    }
  if (velocity) *velocity = (PetscInt)v;
  if (schur) *schur = (PetscInt)s;
  return 0;
}

// ---- StokesStateView / StokesVecView (stokes.C:1821-1915): the VTK dump of -output_vtk ---------------------------------------------
// StokesVecView writes `nodes` lines of `perline` numbers, the first `pernode` of them from the array (stokes.C:1898-1915).
static void stokes_vec_view(FILE* f, const double* a, long long nodes, int pernode, int perline) {
  for (long long i = 0; i < nodes; i++) {
    for (int j = 0; j < pernode && j < perline; j++) fprintf(f, "%20e ", a[i * pernode + j]);
    for (int j = pernode; j < perline; j++) fprintf(f, "0 ");
    fprintf(f, "\n");
  }
}

PetscErrorCode StokesStateViewFile(StokesCtxB200* c, Vec state, const char* path) {
  const int d = (int)c->opt.numDims;
  const long long nodes = c->m;
  if (!state || state->n != (PetscInt)c->g || !path) return SB200_ERR_USER;
  if (c->h_force.size() != (size_t)c->g) return SB200_ERR_USER;  // c->force is set by StokesCreateExactSolution (stokes.C:1001)
  std::vector<double> hx((size_t)c->g), dirichlet((size_t)c->dv), coord((size_t)nodes * d);
  CHK(VecGetValuesHost(state, hx.data()));
  CHK(stokes_fill(c, nullptr, nullptr, &dirichlet));
  std::vector<char> bdy((size_t)nodes);
  {  // node coordinates and the boundary flag, in walk order (stokes.C:791-879)
    int ind[3] = {0, 0, 0};
    for (long long node = 0; node < nodes; node++) {
      bool b = false;
      for (int j = 0; j < d; j++) {
        coord[node * d + j] = cos(ind[j] * M_PI / (c->opt.dim[j] - 1));
        b = b || ind[j] == 0 || ind[j] == c->opt.dim[j] - 1;
      }
      bdy[node] = b;
      for (int j = d - 1; j >= 0; j--) {
        if (++ind[j] < c->opt.dim[j]) break;
        ind[j] = 0;
      }
    }
  }
  // velocity, pressure of the state and of the forcing on the full grid (stokes.C:1828-1850): interior from the global vector,
  // Dirichlet values on the boundary, pressure extrapolated to the boundary by StokesPressureReduceOrder
  std::vector<double> fields[4];
  const std::vector<double>* globals[2] = {&hx, &c->h_force};
  Vec dp = nullptr;
  CHK(VecCreateSeqCUDA(PETSC_COMM_SELF, (PetscInt)c->m, &dp));
  for (int w = 0; w < 2; w++) {
    std::vector<double>& vL = fields[2 * w];
    std::vector<double> pL((size_t)nodes, 0.0);
    vL.assign((size_t)nodes * d, 0.0);
    const std::vector<double>& G = *globals[w];
    long long qi = 0, qb = 0;
    for (long long node = 0; node < nodes; node++) {
      if (bdy[node]) {
        for (int k = 0; k < d; k++) vL[node * d + k] = dirichlet[qb * d + k];  // scatterDL
        qb++;
      } else {
        for (int k = 0; k < d; k++) vL[node * d + k] = G[qi * (d + 1) + k];  // scatterGV + scatterVL
        pL[node] = G[qi * (d + 1) + d];                                        // scatterGP + scatterPL
        qi++;
      }
    }
    PetscErrorCode rc = VecSetValuesHost(dp, pL.data());
    if (!rc) rc = StokesPressureReduceOrder(dp, c);
    fields[2 * w + 1].resize((size_t)nodes);
    if (!rc) rc = VecGetValuesHost(dp, fields[2 * w + 1].data());
    if (rc) {
      VecDestroy(dp);
      return rc;
    }
  }
  CHK(VecDestroy(dp));
  std::vector<double> eta((size_t)nodes), deta((size_t)nodes), strain[3];
  {
    Vec s0 = nullptr, s1 = nullptr;
    PetscErrorCode rc = VecCreateSeqCUDA(PETSC_COMM_SELF, (PetscInt)c->m, &s0);
    if (!rc) rc = VecCreateSeqCUDA(PETSC_COMM_SELF, (PetscInt)(c->m * d), &s1);
    if (!rc) rc = StokesGetState(c, 0, s0);
    if (!rc) rc = VecGetValuesHost(s0, eta.data());
    if (!rc) rc = StokesGetState(c, 1, s0);
    if (!rc) rc = VecGetValuesHost(s0, deta.data());
    for (int j = 0; j < d && !rc; j++) {
      strain[j].resize((size_t)nodes * d);
      rc = StokesGetState(c, 2 + j, s1);
      if (!rc) rc = VecGetValuesHost(s1, strain[j].data());
    }
    if (s0) VecDestroy(s0);
    if (s1) VecDestroy(s1);
    if (rc) return rc;
  }
  FILE* f = fopen(path, "w");
  if (!f) return SB200_ERR_USER;  // PETSC_ERR_FILE_OPEN is 65; the shim has no file error class of its own
  const int mm = c->opt.dim[0], nn = c->opt.dim[1], pp = d > 2 ? c->opt.dim[2] : 1;
  fprintf(f, "# vtk DataFile Version 2.0\nStokes Output\nASCII\nDATASET STRUCTURED_GRID\n");
  fprintf(f, "DIMENSIONS %d %d %d\nPOINTS %d double\n", mm, nn, pp, mm * nn * pp);
  stokes_vec_view(f, coord.data(), nodes, d, 3);
  fprintf(f, "\nPOINT_DATA %d\nVECTORS velocity double\n", mm * nn * pp);
  stokes_vec_view(f, fields[0].data(), nodes, d, 3);
  fprintf(f, "\nSCALARS pressure double 1\nLOOKUP_TABLE default\n");
  stokes_vec_view(f, fields[1].data(), nodes, 1, 1);
  fprintf(f, "\nVECTORS vel_force double\n");
  stokes_vec_view(f, fields[2].data(), nodes, d, 3);
  fprintf(f, "\nSCALARS div_force double 1\nLOOKUP_TABLE default\n");
  stokes_vec_view(f, fields[3].data(), nodes, 1, 1);
  fprintf(f, "\nSCALARS eta double 1\nLOOKUP_TABLE default\n");
  stokes_vec_view(f, eta.data(), nodes, 1, 1);
  fprintf(f, "\nSCALARS deta double 1\nLOOKUP_TABLE default\n");
  stokes_vec_view(f, deta.data(), nodes, 1, 1);
  fprintf(f, "\nTENSORS strain double\n");
  for (long long i = 0; i < nodes; i++) {
    for (int j = 0; j < 3; j++) {
      for (int k = 0; k < 3; k++) fprintf(f, "%20e ", (j < d && k < d) ? strain[j][i * d + k] : 0.0);
      fprintf(f, "\n");
    }
    fprintf(f, "\n");
  }
  fclose(f);
  return 0;
}

// The reference ignores its third argument (a label, "final state" at stokes.C:240) and always writes stokes.vtk (stokes.C:1856).
PetscErrorCode StokesStateView(StokesCtxB200* c, Vec state, const char*) { return StokesStateViewFile(c, state, "stokes.vtk"); }

sb200_stokes* StokesGetHandle(StokesCtxB200* c) { return c ? c->s : nullptr; }

}  // extern "C"
