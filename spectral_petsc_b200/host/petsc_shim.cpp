// Minimal stand-ins for the PETSc objects the reference's callbacks use (include/sb200_petsc_shim.h).
#include <cstdlib>

#include "../../include/sb200_petsc_shim.h"
#include "../../include/spectral_b200.h"

extern "C" {

PetscErrorCode VecCreateSeqCUDA(MPI_Comm, PetscInt n, Vec* v) {
  if (!v || n < 0) return SB200_ERR_ARG;
  Vec x = (Vec)std::calloc(1, sizeof(_p_Vec));
  void* d = nullptr;
  int rc = sb200_malloc(&d, (size_t)n * sizeof(double));
  if (rc) {
    std::free(x);
    return rc;
  }
  sb200_memset0(d, (size_t)n * sizeof(double), nullptr);
  x->d_array = (double*)d;
  x->n = n;
  x->owns = 1;
  *v = x;
  return 0;
}

PetscErrorCode VecCreateSeqCUDAWithArray(MPI_Comm, PetscInt n, double* d_array, Vec* v) {
  if (!v) return SB200_ERR_ARG;
  Vec x = (Vec)std::calloc(1, sizeof(_p_Vec));
  x->d_array = d_array;
  x->n = n;
  x->owns = 0;
  *v = x;
  return 0;
}

PetscErrorCode VecDuplicate(Vec v, Vec* w) { return VecCreateSeqCUDA(PETSC_COMM_SELF, v->n, w); }

PetscErrorCode VecDestroy(Vec v) {
  if (!v) return 0;
  if (v->owns) sb200_free(v->d_array);
  std::free(v);
  return 0;
}

PetscErrorCode VecGetSize(Vec v, PetscInt* n) {
  *n = v->n;
  return 0;
}
PetscErrorCode VecCUDAGetArrayRead(Vec v, const PetscScalar** a) {
  *a = v->d_array;
  return 0;
}
PetscErrorCode VecCUDARestoreArrayRead(Vec, const PetscScalar** a) {
  *a = nullptr;
  return 0;
}
PetscErrorCode VecCUDAGetArrayWrite(Vec v, PetscScalar** a) {
  *a = v->d_array;
  return 0;
}
PetscErrorCode VecCUDARestoreArrayWrite(Vec, PetscScalar** a) {
  *a = nullptr;
  return 0;
}
PetscErrorCode VecSetValuesHost(Vec v, const PetscScalar* h) {
  int rc = sb200_memcpy_h2d(v->d_array, h, (size_t)v->n * sizeof(double), nullptr);
  return rc ? rc : sb200_stream_sync(nullptr);
}
PetscErrorCode VecGetValuesHost(Vec v, PetscScalar* h) {
  int rc = sb200_memcpy_d2h(h, v->d_array, (size_t)v->n * sizeof(double), nullptr);
  return rc ? rc : sb200_stream_sync(nullptr);
}

PetscErrorCode MatCreateShell(MPI_Comm, PetscInt m, PetscInt n, PetscInt, PetscInt, void* ctx, Mat* A) {
  Mat a = (Mat)std::calloc(1, sizeof(_p_Mat));
  a->ctx = ctx;
  a->m = m;
  a->n = n;
  *A = a;
  return 0;
}

PetscErrorCode MatShellSetOperation(Mat A, MatOperation op, void (*f)(void)) {
  switch (op) {
    case MATOP_MULT: A->mult = (PetscErrorCode(*)(Mat, Vec, Vec))f; return 0;
    case MATOP_GET_DIAGONAL: A->getdiagonal = (PetscErrorCode(*)(Mat, Vec))f; return 0;
    case MATOP_DESTROY: A->destroy = (PetscErrorCode(*)(Mat))f; return 0;
  }
  return SB200_ERR_SUP;
}

PetscErrorCode MatShellGetContext(Mat A, void** ctx) {
  *ctx = A->ctx;
  return 0;
}
PetscErrorCode MatMult(Mat A, Vec x, Vec y) { return A->mult ? A->mult(A, x, y) : SB200_ERR_SUP; }
PetscErrorCode MatGetDiagonal(Mat A, Vec y) { return A->getdiagonal ? A->getdiagonal(A, y) : SB200_ERR_SUP; }
PetscErrorCode MatGetSize(Mat A, PetscInt* m, PetscInt* n) {
  if (m) *m = A->m;
  if (n) *n = A->n;
  return 0;
}
PetscErrorCode MatDestroy(Mat A) {
  if (!A) return 0;
  PetscErrorCode rc = A->destroy ? A->destroy(A) : 0;
  if (A->d_rowptr) sb200_free(A->d_rowptr);
  if (A->d_colidx) sb200_free(A->d_colidx);
  if (A->d_vals) sb200_free(A->d_vals);
  std::free(A);
  return rc;
}

PetscErrorCode MatCreateSeqAIJ(MPI_Comm, PetscInt m, PetscInt n, PetscInt, const PetscInt*, Mat* A) {
  Mat a = (Mat)std::calloc(1, sizeof(_p_Mat));
  a->m = m;
  a->n = n;
  a->is_aij = 1;
  *A = a;
  return 0;
}

PetscErrorCode MatSeqAIJGetCSRHost(Mat A, PetscInt* nz, PetscInt* rowptr, PetscInt* colidx, PetscScalar* vals) {
  if (!A || !A->is_aij) return SB200_ERR_ARG;
  if (nz) *nz = A->nz;
  if ((rowptr || colidx || vals) && !A->d_vals) return SB200_ERR_USER;  // not assembled yet
  PetscErrorCode rc = 0;
  if (rowptr) rc = sb200_memcpy_d2h(rowptr, A->d_rowptr, (size_t)(A->m + 1) * sizeof(PetscInt), nullptr);
  if (!rc && colidx) rc = sb200_memcpy_d2h(colidx, A->d_colidx, (size_t)A->nz * sizeof(PetscInt), nullptr);
  if (!rc && vals) rc = sb200_memcpy_d2h(vals, A->d_vals, (size_t)A->nz * sizeof(PetscScalar), nullptr);
  return rc ? rc : sb200_stream_sync(nullptr);
}

PetscErrorCode PCCreate(MPI_Comm, PC* pc) {
  *pc = (PC)std::calloc(1, sizeof(_p_PC));
  return 0;
}
PetscErrorCode PCShellSetContext(PC pc, void* ctx) {
  pc->ctx = ctx;
  return 0;
}
PetscErrorCode PCShellGetContext(PC pc, void** ctx) {
  *ctx = pc->ctx;
  return 0;
}
PetscErrorCode PCDestroy(PC pc) {
  std::free(pc);
  return 0;
}

PetscErrorCode SNESCreate(MPI_Comm, SNES* snes) {
  *snes = (SNES)std::calloc(1, sizeof(_p_SNES));
  return 0;
}
PetscErrorCode SNESSetApplicationContext(SNES snes, void* ctx) {
  snes->appctx = ctx;
  return 0;
}
PetscErrorCode SNESGetApplicationContext(SNES snes, void** ctx) {
  *ctx = snes->appctx;
  return 0;
}
PetscErrorCode SNESDestroy(SNES snes) {
  std::free(snes);
  return 0;
}

}  // extern "C"
