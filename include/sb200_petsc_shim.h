/* sb200_petsc_shim.h - the handful of PETSc names the reference's operator callbacks are written
 * against, so that the B200 host layer keeps the reference's exact function names and signatures
 * (chebyshev.h:27-34, elliptic.C:105-112, stokes.C:67-79) without PETSc being installed here.
 *
 * With a real CUDA-enabled PETSc this header is NOT used: include <petscsnes.h> instead, let Vec be
 * VECCUDA and use PETSc's own VecCUDAGetArrayRead/Write + Restore (same names as below); see INTEGRATION.md.
 * Only what the hot path touches exists: Seq CUDA vectors, shell matrices, the callback typedefs.
 */
#ifndef SB200_PETSC_SHIM_H
#define SB200_PETSC_SHIM_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int PetscErrorCode;
typedef int PetscInt;
typedef double PetscScalar;
typedef double PetscReal;
typedef int PetscTruth; /* PETSc 3.0 spelling used by the reference (elliptic.C:130) */
typedef int MPI_Comm;
#define PETSC_COMM_SELF 1
#define PETSC_COMM_WORLD 2
#define PETSC_NULL 0
#define PETSC_TRUE 1
#define PETSC_FALSE 0
#define FFTW_ESTIMATE (1U << 6) /* accepted and ignored: there is no planner (elliptic.C:159) */

typedef struct _p_Vec* Vec;
typedef struct _p_Mat* Mat;
typedef struct _p_SNES* SNES;
typedef struct _p_PC* PC;
typedef enum { MATOP_MULT = 3, MATOP_GET_DIAGONAL = 17, MATOP_DESTROY = 250 } MatOperation;
typedef enum { SAME_NONZERO_PATTERN, DIFFERENT_NONZERO_PATTERN } MatStructure;

struct _p_Vec {
  double* d_array; /* device pointer: the VECCUDA array */
  PetscInt n;
  int owns;
};
struct _p_Mat {
  void* ctx;
  PetscInt m, n;
  PetscErrorCode (*mult)(Mat, Vec, Vec);
  PetscErrorCode (*getdiagonal)(Mat, Vec);
  PetscErrorCode (*destroy)(Mat);
  /* MATSEQAIJ created by MatCreateSeqAIJ: device CSR (the MATSEQAIJCUSPARSE layout), filled by FormJacobian /
   * StokesPCSetUp0; the arrays are allocated at the first assembly, nz = entries stored */
  int is_aij;
  PetscInt nz;
  PetscInt* d_rowptr;
  PetscInt* d_colidx;
  PetscScalar* d_vals;
};
struct _p_PC {
  void* ctx; /* PCShell context (stokes.C:163) */
};
struct _p_SNES {
  void* appctx;
};

/* Vec (device resident) */
PetscErrorCode VecCreateSeqCUDA(MPI_Comm comm, PetscInt n, Vec* v);
PetscErrorCode VecCreateSeqCUDAWithArray(MPI_Comm comm, PetscInt n, double* d_array, Vec* v);
PetscErrorCode VecDuplicate(Vec v, Vec* w);
PetscErrorCode VecDestroy(Vec v); /* PETSc 3.0 signature, as the reference calls it (elliptic.C:237) */
PetscErrorCode VecGetSize(Vec v, PetscInt* n);
PetscErrorCode VecCUDAGetArrayRead(Vec v, const PetscScalar** a);
PetscErrorCode VecCUDARestoreArrayRead(Vec v, const PetscScalar** a);
PetscErrorCode VecCUDAGetArrayWrite(Vec v, PetscScalar** a);
PetscErrorCode VecCUDARestoreArrayWrite(Vec v, PetscScalar** a);
PetscErrorCode VecSetValuesHost(Vec v, const PetscScalar* h); /* whole-vector upload  */
PetscErrorCode VecGetValuesHost(Vec v, PetscScalar* h);       /* whole-vector download */

/* MatShell */
PetscErrorCode MatCreateShell(MPI_Comm comm, PetscInt m, PetscInt n, PetscInt M, PetscInt N, void* ctx, Mat* A);
PetscErrorCode MatShellSetOperation(Mat A, MatOperation op, void (*f)(void));
PetscErrorCode MatShellGetContext(Mat A, void** ctx);
PetscErrorCode MatMult(Mat A, Vec x, Vec y);
PetscErrorCode MatGetDiagonal(Mat A, Vec y);
PetscErrorCode MatGetSize(Mat A, PetscInt* m, PetscInt* n);
PetscErrorCode MatDestroy(Mat A); /* PETSc 3.0 signature (elliptic.C:235) */

/* SeqAIJ: the preconditioning matrices P (elliptic.C:167) and MatVVPC (stokes.C:326); nz / nnz are accepted and ignored
 * (the assembling kernel knows the pattern in closed form).  GetCSRHost downloads whatever outputs are non-NULL. */
PetscErrorCode MatCreateSeqAIJ(MPI_Comm comm, PetscInt m, PetscInt n, PetscInt nz, const PetscInt* nnz, Mat* A);
PetscErrorCode MatSeqAIJGetCSRHost(Mat A, PetscInt* nz, PetscInt* rowptr, PetscInt* colidx, PetscScalar* vals);

/* PCShell: only the context plumbing StokesPCSetUp0 uses (stokes.C:163-166,1169) */
PetscErrorCode PCCreate(MPI_Comm comm, PC* pc);
PetscErrorCode PCShellSetContext(PC pc, void* ctx);
PetscErrorCode PCShellGetContext(PC pc, void** ctx);
PetscErrorCode PCDestroy(PC pc);

/* SNES: only the application-context plumbing the callbacks use (elliptic.C:180,604) */
PetscErrorCode SNESCreate(MPI_Comm comm, SNES* snes);
PetscErrorCode SNESSetApplicationContext(SNES snes, void* ctx);
PetscErrorCode SNESGetApplicationContext(SNES snes, void** ctx);
PetscErrorCode SNESDestroy(SNES snes);

#ifdef __cplusplus
}
#endif
#endif
