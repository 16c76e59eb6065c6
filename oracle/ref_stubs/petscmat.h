/* Minimal stand-in for PETSc (~3.0 API) - ONLY what /root/reference/chebyshev.c and elliptic.C use - so that the reference's
 * own source files compile unmodified here (PETSc is not installed in this image and cannot be: no network).  Sequential
 * Vec / IS / VecScatter / MatShell / triplet-recording SeqAIJ with the documented semantics of each call; the solver
 * objects (SNES, KSP, PC; petscsnes.h) are inert.  Test infrastructure (oracle/), never part of the product. */
#ifndef SB200_STUB_PETSCMAT_H
#define SB200_STUB_PETSCMAT_H
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef __cplusplus
#define PETSC_EXTERN_CXX_BEGIN extern "C" {
#define PETSC_EXTERN_CXX_END }
#else
#define PETSC_EXTERN_CXX_BEGIN
#define PETSC_EXTERN_CXX_END
#endif
PETSC_EXTERN_CXX_BEGIN
typedef int PetscErrorCode;
typedef int PetscInt;
typedef double PetscScalar;
typedef double PetscReal;
typedef int MPI_Comm;
typedef enum { PETSC_FALSE, PETSC_TRUE } PetscTruth;
#define PETSC_COMM_SELF 1
#define PETSC_COMM_WORLD 2
#define PETSC_NULL 0
#define PETSC_DECIDE (-1)
#define PETSC_ERR_USER 83
#define PETSC_ERR_SUP 56
#define PETSC_PI 3.14159265358979323846264338327950288419716939937510582
typedef enum { INSERT_VALUES = 1, ADD_VALUES = 2 } InsertMode;
typedef enum { SCATTER_FORWARD = 0, SCATTER_REVERSE = 1 } ScatterMode;
typedef enum { NORM_1 = 0, NORM_2 = 1, NORM_INFINITY = 3 } NormType;
typedef enum { SAME_NONZERO_PATTERN, DIFFERENT_NONZERO_PATTERN, SAME_PRECONDITIONER } MatStructure;
typedef enum { MAT_FLUSH_ASSEMBLY = 1, MAT_FINAL_ASSEMBLY = 0 } MatAssemblyType;
typedef enum { MATOP_MULT = 3, MATOP_GET_DIAGONAL = 17, MATOP_DESTROY = 60 } MatOperation;

typedef struct _stub_Vec { int n, bs, owns; double* a; }* Vec;
typedef struct _stub_IS { int n; int* idx; }* IS;
typedef struct _stub_Scatter { int n; int* from; int* to; }* VecScatter;  /* y[to[i]] (=|+=) x[from[i]] */
typedef struct _stub_Mat {
  void* ctx;
  PetscErrorCode (*mult)(struct _stub_Mat*, Vec, Vec);
  PetscErrorCode (*destroy)(struct _stub_Mat*);
  PetscErrorCode (*getdiag)(struct _stub_Mat*, Vec);
  int m, n;
  /* SeqAIJ stand-in: every MatSetValues entry recorded as a triplet (negative indices dropped, as PETSc does) */
  int nt, cap;
  int *ti, *tj;
  double* tv;
}* Mat;
typedef void* PetscObject;
typedef void* PetscViewer;
typedef struct _stub_NullSpace { int has_const, n; Vec* vecs; }* MatNullSpace;
typedef void* ISColoring;
typedef void* MatFDColoring;
#define MATCOLORING_ID "id"
#define PETSC_VIEWER_ASCII "ascii"
#define PETSC_VIEWER_ASCII_VTK 0
#define PETSC_VIEWER_STDOUT_SELF ((PetscViewer)0)
#define PETSC_VIEWER_STDOUT_WORLD ((PetscViewer)0)

extern char sb200_stub_last_error[256];
#define PetscFunctionBegin
#define PetscFunctionReturn(a) return (a)
#define CHKERRQ(e) do { if (e) return (e); } while (0)
#define SETERRQ(code, msg) do { snprintf(sb200_stub_last_error, 256, "%s", msg); return (code); } while (0)
#define SETERRQ1(code, msg, a) do { snprintf(sb200_stub_last_error, 256, msg, a); return (code); } while (0)
#define SETERRQ2(code, msg, a, b) do { snprintf(sb200_stub_last_error, 256, msg, a, b); return (code); } while (0)
#define PetscSqr(a) ((a) * (a))
#define PetscAbs(a) (((a) >= 0) ? (a) : -(a))
#define PetscMax(a, b) (((a) < (b)) ? (b) : (a))
#define PetscMin(a, b) (((a) < (b)) ? (a) : (b))
#define PetscMemcpy(d, s, n) (memcpy((d), (s), (n)), 0)
#define PetscMalloc(sz, pp) ((*(void**)(pp) = malloc((size_t)((sz) > 0 ? (sz) : 1))) ? 0 : 55)
#define PetscFree(p) (free(p), 0)
#define PetscMemzero(p, sz) (memset((p), 0, (sz)), 0)
#define PetscMalloc2(m1, t1, r1, m2, t2, r2) (PetscMalloc((m1) * sizeof(t1), r1) || PetscMalloc((m2) * sizeof(t2), r2))
#define PetscFree2(a, b) (free(a), free(b), 0)
#define PetscMalloc5(m1, t1, r1, m2, t2, r2, m3, t3, r3, m4, t4, r4, m5, t5, r5) \
  (PetscMalloc((m1) * sizeof(t1), r1) || PetscMalloc((m2) * sizeof(t2), r2) || PetscMalloc((m3) * sizeof(t3), r3) || \
   PetscMalloc((m4) * sizeof(t4), r4) || PetscMalloc((m5) * sizeof(t5), r5))
#define PetscFree5(a, b, c, d, e) (free(a), free(b), free(c), free(d), free(e), 0)
#define PetscMalloc6(m1, t1, r1, m2, t2, r2, m3, t3, r3, m4, t4, r4, m5, t5, r5, m6, t6, r6) \
  (PetscMalloc5(m1, t1, r1, m2, t2, r2, m3, t3, r3, m4, t4, r4, m5, t5, r5) || PetscMalloc((m6) * sizeof(t6), r6))
#define PetscFree6(a, b, c, d, e, f) (free(a), free(b), free(c), free(d), free(e), free(f), 0)

PetscErrorCode PetscPrintf(MPI_Comm comm, const char* fmt, ...);
PetscErrorCode PetscObjectGetComm(PetscObject o, MPI_Comm* comm);
PetscErrorCode PetscMallocSetDumpLog(void);
PetscErrorCode PetscMallocDumpLog(FILE* f);

PetscErrorCode VecCreateSeq(MPI_Comm comm, PetscInt n, Vec* v);
PetscErrorCode VecCreateSeqWithArray(MPI_Comm comm, PetscInt n, PetscScalar* a, Vec* v);
PetscErrorCode VecDuplicate(Vec v, Vec* w);
PetscErrorCode VecDuplicateVecs(Vec v, PetscInt n, Vec** w);
PetscErrorCode VecDestroyVecs(Vec* w, PetscInt n);
PetscErrorCode VecDestroy(Vec v);
PetscErrorCode VecSetBlockSize(Vec v, PetscInt bs);
PetscErrorCode VecGetSize(Vec v, PetscInt* n);
PetscErrorCode VecGetArray(Vec v, PetscScalar** a);
PetscErrorCode VecRestoreArray(Vec v, PetscScalar** a);
PetscErrorCode VecGetArrays(const Vec* v, PetscInt n, PetscScalar*** a);
PetscErrorCode VecRestoreArrays(const Vec* v, PetscInt n, PetscScalar*** a);
PetscErrorCode VecSet(Vec v, PetscScalar a);
PetscErrorCode VecZeroEntries(Vec v);
PetscErrorCode VecCopy(Vec x, Vec y);
PetscErrorCode VecAXPY(Vec y, PetscScalar a, Vec x);
PetscErrorCode VecScale(Vec x, PetscScalar a);
PetscErrorCode VecNorm(Vec x, NormType t, PetscReal* r);
PetscErrorCode VecPointwiseDivide(Vec w, Vec x, Vec y);
PetscErrorCode VecView(Vec v, PetscViewer vw);
PetscErrorCode VecCreate(MPI_Comm comm, Vec* v);
PetscErrorCode VecSetSizes(Vec v, PetscInt n, PetscInt N);
PetscErrorCode VecSetFromOptions(Vec v);
PetscErrorCode VecStrideGather(Vec v, PetscInt start, Vec s, InsertMode m);   /* s[i] = v[i*bs + start] */
PetscErrorCode VecStrideScatter(Vec s, PetscInt start, Vec v, InsertMode m);  /* v[i*bs + start] = s[i] */
PetscErrorCode VecMin(Vec v, PetscInt* p, PetscReal* val);
PetscErrorCode VecMax(Vec v, PetscInt* p, PetscReal* val);
PetscErrorCode VecReciprocal(Vec v);
PetscErrorCode VecNormalize(Vec v, PetscReal* val);
PetscErrorCode PetscIntView(PetscInt n, const PetscInt* idx, PetscViewer vw);
PetscErrorCode PetscRealView(PetscInt n, const PetscReal* idx, PetscViewer vw);
PetscErrorCode PetscViewerCreate(MPI_Comm comm, PetscViewer* v);
PetscErrorCode PetscViewerSetType(PetscViewer v, const char* t);
PetscErrorCode PetscViewerSetFormat(PetscViewer v, int f);
PetscErrorCode PetscViewerFileSetName(PetscViewer v, const char* name);
PetscErrorCode PetscViewerASCIIPrintf(PetscViewer v, const char* fmt, ...);
PetscErrorCode PetscViewerDestroy(PetscViewer v);

PetscErrorCode ISCreateGeneral(MPI_Comm comm, PetscInt n, const PetscInt* idx, IS* is);
PetscErrorCode ISDestroy(IS is);
PetscErrorCode ISGetIndices(IS is, const PetscInt** idx);
PetscErrorCode ISRestoreIndices(IS is, const PetscInt** idx);
PetscErrorCode ISView(IS is, PetscViewer vw);

PetscErrorCode VecScatterCreate(Vec x, IS ix, Vec y, IS iy, VecScatter* s);
PetscErrorCode VecScatterBegin(VecScatter s, Vec x, Vec y, InsertMode im, ScatterMode sm);
PetscErrorCode VecScatterEnd(VecScatter s, Vec x, Vec y, InsertMode im, ScatterMode sm);
PetscErrorCode VecScatterDestroy(VecScatter s);

PetscErrorCode MatCreateShell(MPI_Comm comm, PetscInt m, PetscInt n, PetscInt M, PetscInt N, void* ctx, Mat* A);
PetscErrorCode MatShellSetOperation(Mat A, MatOperation op, void (*f)(void));
PetscErrorCode MatShellGetContext(Mat A, void** ctx);
PetscErrorCode MatMult(Mat A, Vec x, Vec y);
PetscErrorCode MatDestroy(Mat A);
PetscErrorCode MatGetSize(Mat A, PetscInt* m, PetscInt* n);
PetscErrorCode MatCreate(MPI_Comm comm, Mat* A);
PetscErrorCode MatSetSizes(Mat A, PetscInt m, PetscInt n, PetscInt M, PetscInt N);
PetscErrorCode MatSetType(Mat A, const char* type);
PetscErrorCode MatSetFromOptions(Mat A);
PetscErrorCode MatCreateSeqAIJ(MPI_Comm comm, PetscInt m, PetscInt n, PetscInt nz, const PetscInt* nnz, Mat* A);
PetscErrorCode MatSetValues(Mat A, PetscInt m, const PetscInt* im, PetscInt n, const PetscInt* in, const PetscScalar* v, InsertMode mode);
PetscErrorCode MatAssemblyBegin(Mat A, MatAssemblyType t);
PetscErrorCode MatAssemblyEnd(Mat A, MatAssemblyType t);
PetscErrorCode MatZeroEntries(Mat A);
PetscErrorCode MatDiagonalScale(Mat A, Vec l, Vec r);
PetscErrorCode MatView(Mat A, PetscViewer vw);
PetscErrorCode MatNullSpaceCreate(MPI_Comm comm, PetscTruth has_const, PetscInt n, const Vec* vecs, MatNullSpace* ns);
PetscErrorCode MatNullSpaceDestroy(MatNullSpace ns);
PetscErrorCode MatNullSpaceRemove(MatNullSpace ns, Vec v, Vec* out);
PetscErrorCode MatNullSpaceTest(MatNullSpace ns, Mat A, PetscTruth* isNull);
PetscErrorCode MatGetColoring(Mat A, const char* type, ISColoring* c);
PetscErrorCode MatFDColoringCreate(Mat A, ISColoring c, MatFDColoring* f);
PetscErrorCode MatFDColoringSetFunction(MatFDColoring f, PetscErrorCode (*fn)(void), void* ctx);
PetscErrorCode MatFDColoringSetFromOptions(MatFDColoring f);
PetscErrorCode MatFDColoringApply(Mat A, MatFDColoring f, Vec x, MatStructure* flag, void* ctx);
PETSC_EXTERN_CXX_END
#endif
