/* sb200_reference_api.h - the reference's own operator interface, re-implemented on the B200 path.
 * Names, argument order and error behaviour are the reference's (file:line cited per function);
 * Vecs are device resident.  These are thin C++ wrappers over the C ABI in spectral_b200.h. */
#ifndef SB200_REFERENCE_API_H
#define SB200_REFERENCE_API_H

#include "sb200_petsc_shim.h"

#ifdef __cplusplus
extern "C" {
#endif

/* chebyshev.h:25-27: the 1-D operator (chebyshev.c:8-85), the same derivative as MatCreateCheb(rank 1, tr 0) */
PetscErrorCode MatCreateChebD1(MPI_Comm comm, Vec vx, Vec vy, unsigned flag, Mat* A);
PetscErrorCode ChebD1Mult(Mat A, Vec vx, Vec vy);
PetscErrorCode ChebD1Destroy(Mat A);
/* chebyshev.h:29-32 */
PetscErrorCode MatCreateCheb(MPI_Comm comm, int rank, int tr, int* dims, unsigned flag, Vec vx, Vec vy, Mat* A);
PetscErrorCode ChebMult(Mat A, Vec vx, Vec vy);
PetscErrorCode ChebDestroy(Mat A);

/* elliptic.C:88-112 */
typedef enum { BDY_DIRICHLET, BDY_NEUMANN } BdyType;
typedef struct {
  BdyType type;
  PetscScalar value;
} BdyCond;
typedef PetscErrorCode (*BdyFunc)(int, double*, double*, BdyCond*);
typedef struct { /* elliptic.C:88-94 */
  PetscInt exact, d, *dim;
  Mat A;
  Vec b;
  PetscReal gamma, exponent;
  int debug;
} AppCtx;
PetscErrorCode DirichletBdy(int d, double* x, double* n, BdyCond* bc);
PetscErrorCode MatCreate_Elliptic(MPI_Comm comm, int d, int* dim, unsigned flag, BdyFunc bf, Vec* vG, Mat* A);
PetscErrorCode MatMult_Elliptic(Mat A, Vec U, Vec V);
PetscErrorCode MatDestroy_Elliptic(Mat A);
PetscErrorCode FormFunction(SNES snes, Vec U, Vec rhs, void* void_ac);
/* FormJacobian (elliptic.C:537-590): fills the SeqAIJ matrix *P (MatCreateSeqAIJ, elliptic.C:167) on the device from the
 * eta / deta / gradu the last FormFunction cached; *flag = SAME_NONZERO_PATTERN (:588).  w is not read, as in the reference. */
PetscErrorCode FormJacobian(SNES snes, Vec w, Mat* A, Mat* P, MatStructure* flag, void* void_ac);
/* CreateExactSolution(snes, u, u2) (elliptic.C:594-677); -cos_scale is passed explicitly because the
 * reference reads it from the options database without a default (elliptic.C:607-609). */
PetscErrorCode CreateExactSolution(SNES snes, Vec u, Vec u2, PetscReal cos_scale);

/* stokes.C:67-79: the context is opaque here; options that StokesProcessOptions reads from the PETSc
 * options database (stokes.C:392-495) are passed in a plain struct. */
typedef struct {
  PetscInt numDims, dim[3];
  PetscInt exact, rheology;
  PetscReal hardness, exponent, regularization, gamma0;
} StokesOptionsB200;
typedef struct StokesCtxB200 StokesCtxB200;
PetscErrorCode StokesCreate(MPI_Comm comm, const StokesOptionsB200* opt, Mat* A, Vec* x, StokesCtxB200** ctx);
PetscErrorCode StokesDestroy(StokesCtxB200* ctx);
PetscErrorCode StokesMatMult(Mat A, Vec xG, Vec yG);
PetscErrorCode StokesMatMultVV(Mat A, Vec xG, Vec yG);
PetscErrorCode StokesMatMultPV(Mat A, Vec xG, Vec yG);
PetscErrorCode StokesMatMultVP(Mat A, Vec xG, Vec yG);
PetscErrorCode StokesMatGetDiagonalSchur(Mat S, Vec y);
/* StokesExact0..3 (stokes.C:1948-2034): value = [u_0..u_{d-1}, p], rhs = forcing at the point coord (either may be NULL);
 * StokesDirichlet (stokes.C:2039-2050) evaluates ctx->exact and reports a Dirichlet condition.  Host functions. */
typedef enum { DIRICHLET, NEUMANN, MIXED, OUTFLOW } StokesBdyType; /* stokes.C:14 `BdyType`; elliptic.C's enum of that name is above */
typedef PetscErrorCode (*StokesExactFunc)(PetscInt d, PetscReal* coord, PetscReal* value, PetscReal* rhs, void* ctx);
typedef struct {
  StokesExactFunc exact;
  void* exactCtx;
} StokesExactBoundaryCtx;
PetscErrorCode StokesExact0(PetscInt d, PetscReal* coord, PetscReal* value, PetscReal* rhs, void* ctx);
PetscErrorCode StokesExact1(PetscInt d, PetscReal* coord, PetscReal* value, PetscReal* rhs, void* ctx);
PetscErrorCode StokesExact2(PetscInt d, PetscReal* coord, PetscReal* value, PetscReal* rhs, void* ctx);
PetscErrorCode StokesExact3(PetscInt d, PetscReal* coord, PetscReal* value, PetscReal* rhs, void* ctx);
PetscErrorCode StokesDirichlet(PetscInt d, PetscReal* coord, PetscReal* normal, StokesBdyType* type, PetscReal* value, void* ctx);
/* StokesDivergence(ctx, withDirichlet, xG, yG) (stokes.C:570-595) */
PetscErrorCode StokesDivergence(StokesCtxB200* ctx, PetscTruth withDirichlet, Vec xG, Vec yG);
/* The rheologies as the host callbacks StokesOptions stores (stokes.C:1920-1944); ctx = the StokesOptionsB200 holding hardness,
 * exponent, regularization, gamma0.  The kernels evaluate the same expressions per node on the device. */
PetscErrorCode StokesRheologyLinear(PetscInt d, PetscReal gamma, PetscReal* eta, PetscReal* deta, void* ctx);
PetscErrorCode StokesRheologyPower(PetscInt d, PetscReal gamma, PetscReal* eta, PetscReal* deta, void* ctx);
/* polyInterp (util.C:129-144): Neville evaluation at x0 and x1 of the interpolant through (x[i], w[4*i]) / (x[i], w[4*i+1]),
 * i < n; w is the reference's width-4 work array and is overwritten like there. */
PetscErrorCode polyInterp(const PetscInt n, const PetscReal* x, PetscScalar* w, const PetscReal x0, const PetscReal x1, PetscScalar* f0, PetscScalar* f1);
/* StokesMatMultSchur (stokes.C:523-535): y = -PV * KSPSolve(KSPSchurVelocity, VP * x).  The inner KSP is PETSc's; it is registered
 * here as a callback on device Vecs (with PETSc: `return KSPSolve((KSP)ksp, rhs, sol);`).  Without one the shell's MULT fails
 * with PETSC_ERR_ARG_*, like a KSPSolve on an unset KSP would. */
typedef PetscErrorCode (*StokesVelocitySolve)(void* ksp, Vec rhs, Vec sol);
PetscErrorCode StokesSetSchurVelocitySolve(StokesCtxB200* ctx, StokesVelocitySolve solve, void* ksp);
PetscErrorCode StokesMatMultSchur(Mat S, Vec xG, Vec yG);
PetscErrorCode StokesFunction(SNES snes, Vec xG, Vec yG, void* ctx);
/* StokesJacobian (stokes.C:761-769): the Jacobian state was fixed up by StokesFunction; only reports DIFFERENT_NONZERO_PATTERN */
PetscErrorCode StokesJacobian(SNES snes, Vec w, Mat* Ashell, Mat* Pshell, MatStructure* flag, void* ctx);
PetscErrorCode StokesCreateExactSolution(SNES snes, Vec U, Vec U2);
/* the inner shells created by StokesCreate (stokes.C:308-325) and the SeqAIJ matrix MatVVPC (stokes.C:326) */
PetscErrorCode StokesGetShells(StokesCtxB200* ctx, Mat* MatVV, Mat* MatPV, Mat* MatVP, Mat* MatSchur);
PetscErrorCode StokesGetPCMatrix(StokesCtxB200* ctx, Mat* MatVVPC);
/* StokesPCSetUp0 (stokes.C:1160-1240), the PCShell set-up routine of -pcvel 0 (stokes.C:166): assembles MatVVPC on the device
 * from the eta the last StokesFunction cached.  The PC shell's context is the Stokes context (stokes.C:163). */
PetscErrorCode StokesPCSetUp0(PC pc);
/* StokesPCApply0..3 (stokes.C:1714-1817), the PCShell apply routines -pc_saddle_type selects (stokes.C:171-185); the PC shell's
 * context is the Stokes context (stokes.C:163).  x, y: device Vecs of g doubles.  The three inner KSPs of stokes.C:328-341 run on
 * the device (sb200_saddle_*); what stays with the caller is the preconditioner PETSc builds on MatVVPC for KSPVelocity /
 * KSPSchurVelocity, registered here as callbacks on device Vecs (with PETSc: `return PCApply((PC)pc, r, z);`; NULL = PCNONE),
 * and the inner tolerances the options database would supply (-vel_ksp_rtol, -vel_ksp_max_it, -schur_ksp_rtol,
 * -schur_ksp_max_it, -svel_ksp_type preonly). */
PetscErrorCode StokesSetVelocityPC(StokesCtxB200* ctx, StokesVelocitySolve vel_pc, void* vel_pc_ctx, StokesVelocitySolve svel_pc, void* svel_pc_ctx);
PetscErrorCode StokesSetInnerSolves(StokesCtxB200* ctx, PetscReal vel_rtol, PetscInt vel_max_it, PetscReal schur_rtol, PetscInt schur_max_it, PetscTruth svel_preonly);
/* -svel_ksp_rtol / -svel_ksp_max_it (KSPSchurVelocity's own prefix, stokes.C:338-341; defaults 1e-5 / 10000) */
PetscErrorCode StokesSetSchurVelocityTolerances(StokesCtxB200* ctx, PetscReal svel_rtol, PetscInt svel_max_it);
PetscErrorCode StokesPCApply0(PC pc, Vec x, Vec y);
PetscErrorCode StokesPCApply1(PC pc, Vec x, Vec y);
PetscErrorCode StokesPCApply2(PC pc, Vec x, Vec y);
PetscErrorCode StokesPCApply3(PC pc, Vec x, Vec y);
/* the null space StokesRemoveConstantPressure attaches to the outer KSP (stokes.C:1006-1025), applied to a global Vec in place */
PetscErrorCode StokesNullSpaceRemove(StokesCtxB200* ctx, Vec x);
PetscErrorCode StokesGetInnerIterations(StokesCtxB200* ctx, PetscInt* velocity, PetscInt* schur);
/* StokesStateView(ctx, state, label) (stokes.C:1821-1894, called for -output_vtk at :238-242): the VTK dump of velocity, pressure,
 * forcing, eta, deta and the strain tensor on the full grid.  Like the reference it ignores the label and writes ./stokes.vtk
 * (stokes.C:1856); StokesStateViewFile takes the path.  Set-up / output work: runs on host copies, with the boundary pressure
 * extrapolated on the device by StokesPressureReduceOrder. */
PetscErrorCode StokesStateView(StokesCtxB200* ctx, Vec state, const char* label);
PetscErrorCode StokesStateViewFile(StokesCtxB200* ctx, Vec state, const char* path);
/* the C-ABI handle behind the context (for sb200_saddle_create / sb200_ksp_set_operators without going through Vecs) */
struct sb200_stokes* StokesGetHandle(StokesCtxB200* ctx);
PetscErrorCode StokesSetContinuation(StokesCtxB200* ctx, PetscReal exponent, PetscReal regularization); /* stokes.C:218-219 */
/* StokesPressureReduceOrder(pL, ctx) (stokes.C:1029-1080) on a local pressure Vec of m doubles, in place */
PetscErrorCode StokesPressureReduceOrder(Vec pL, StokesCtxB200* ctx);
/* what the driver reads out of the context: VecMin / VecMax of c->eta (stokes.C:731-734), the DOF counts printed at :891, and
 * the cached fields StokesStateView dumps (:1821-1894): which = 0 eta (m), 1 deta (m), 2+j strain[j] (m*d) */
PetscErrorCode StokesGetEtaMinMax(StokesCtxB200* ctx, PetscReal* minEta, PetscReal* maxEta);
PetscErrorCode StokesGetSizes(StokesCtxB200* ctx, PetscInt* m, PetscInt* g, PetscInt* gp, PetscInt* gv, PetscInt* dv);
PetscErrorCode StokesGetState(StokesCtxB200* ctx, PetscInt which, Vec out);

#ifdef __cplusplus
}
#endif
#endif
