/* Inert stand-ins for the PETSc solver objects elliptic.C's main() touches (SNES / KSP / PC / options).  Only the
 * application-context slot of SNES carries data (CreateExactSolution reads it, elliptic.C:601).  Test infrastructure. */
#ifndef SB200_STUB_PETSCSNES_H
#define SB200_STUB_PETSCSNES_H
#include <petscmat.h>
PETSC_EXTERN_CXX_BEGIN
typedef struct _stub_SNES { void* appctx; }* SNES;
typedef struct _stub_KSP { int dummy; }* KSP;
typedef struct _stub_PC { void* ctx; }* PC;
typedef int SNESConvergedReason;
extern const char* const* SNESConvergedReasons;
#define KSPFGMRES "fgmres"
#define PCILU "ilu"
#define PCJACOBI "jacobi"
#define PCSHELL "shell"
PetscErrorCode PetscInitialize(int* argc, char*** args, const char* file, const char* help);
PetscErrorCode PetscFinalize(void);
/* option queries: -cos_scale (and any other real) can be preset with ref_set_option_real(); everything else reports "not set" */
#define PetscOptionsBegin(comm, prefix, title, sec) 0
#define PetscOptionsEnd() 0
PetscErrorCode PetscOptionsIntArray(const char* opt, const char* text, const char* man, PetscInt* v, PetscInt* n, PetscTruth* set);
PetscErrorCode PetscOptionsInt(const char* opt, const char* text, const char* man, PetscInt def, PetscInt* v, PetscTruth* set);
PetscErrorCode PetscOptionsReal(const char* opt, const char* text, const char* man, PetscReal def, PetscReal* v, PetscTruth* set);
PetscErrorCode PetscOptionsGetReal(const char* pre, const char* name, PetscReal* v, PetscTruth* set);
PetscErrorCode PetscOptionsGetInt(const char* pre, const char* name, PetscInt* v, PetscTruth* set);
PetscErrorCode PetscOptionsHasName(const char* pre, const char* name, PetscTruth* set);
/* the option database of the stand-in: preset by the ctypes drivers before the reference parses its options */
void ref_set_option_real(const char* name, double v);
void ref_set_option_int(const char* name, int v);
void ref_set_option_intarray(const char* name, int n, const int* v);
void ref_clear_options(void);
PetscErrorCode SNESCreate(MPI_Comm comm, SNES* s);
PetscErrorCode SNESDestroy(SNES s);
PetscErrorCode SNESSetJacobian(SNES s, Mat A, Mat P, PetscErrorCode (*f)(SNES, Vec, Mat*, Mat*, MatStructure*, void*), void* ctx);
PetscErrorCode SNESSetFunction(SNES s, Vec r, PetscErrorCode (*f)(SNES, Vec, Vec, void*), void* ctx);
PetscErrorCode SNESSetApplicationContext(SNES s, void* ctx);
PetscErrorCode SNESGetApplicationContext(SNES s, void** ctx);
PetscErrorCode SNESGetKSP(SNES s, KSP* k);
PetscErrorCode SNESSetFromOptions(SNES s);
PetscErrorCode SNESSolve(SNES s, Vec b, Vec x);
PetscErrorCode SNESGetIterationNumber(SNES s, PetscInt* its);
PetscErrorCode SNESGetConvergedReason(SNES s, SNESConvergedReason* r);
PetscErrorCode KSPCreate(MPI_Comm comm, KSP* k);
PetscErrorCode KSPDestroy(KSP k);
PetscErrorCode KSPSetOperators(KSP k, Mat A, Mat P, MatStructure f);
PetscErrorCode KSPSetOptionsPrefix(KSP k, const char* p);
PetscErrorCode KSPSetFromOptions(KSP k);
PetscErrorCode KSPSetNullSpace(KSP k, MatNullSpace ns);
PetscErrorCode KSPSolve(KSP k, Vec b, Vec x);  /* stand-in: x = b (the "solver" is the identity) */
PetscErrorCode PCShellSetContext(PC pc, void* ctx);
PetscErrorCode PCShellGetContext(PC pc, void** ctx);
PetscErrorCode PCShellSetSetUp(PC pc, PetscErrorCode (*f)(PC));
PetscErrorCode PCShellSetApply(PC pc, PetscErrorCode (*f)(PC, Vec, Vec));
PetscErrorCode KSPSetType(KSP k, const char* t);
PetscErrorCode KSPGetPC(KSP k, PC* pc);
PetscErrorCode KSPGetIterationNumber(KSP k, PetscInt* its);
PetscErrorCode PCSetType(PC pc, const char* t);
PetscErrorCode PCFactorSetLevels(PC pc, PetscInt l);
PETSC_EXTERN_CXX_END
#endif
