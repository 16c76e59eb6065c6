/* Implementation of the FFTW / PETSc stand-ins (see fftw3.h, petscmat.h in this directory) plus the ctypes entry points
 * that drive the REFERENCE's own MatCreateCheb / ChebMult / ChebDestroy (chebyshev.c:89-235).  Test infrastructure. */
#include <fftw3.h>
#include <petscsnes.h>
#include <stdarg.h>

char sb200_stub_last_error[256];

/* ---- FFTW stand-in ---------------------------------------------------------------------------------------- */
struct sb200_stub_plan {
  fftw_r2r_kind kind;
  fftw_iodim t;        /* transformed dimension */
  int hr;              /* howmany rank */
  fftw_iodim h[16];
};

void* fftw_malloc(size_t n) { return malloc(n ? n : 1); }
void fftw_free(void* p) { free(p); }

fftw_plan fftw_plan_guru_r2r(int rank, const fftw_iodim* dims, int howmany_rank, const fftw_iodim* howmany_dims, double* in, double* out,
                             const fftw_r2r_kind* kind, unsigned flags) {
  (void)in; (void)out; (void)flags;
  if (rank != 1 || howmany_rank > 16) return NULL;
  fftw_plan p = (fftw_plan)malloc(sizeof(*p));
  p->kind = kind[0];
  p->t = dims[0];
  p->hr = howmany_rank;
  for (int i = 0; i < howmany_rank; i++) p->h[i] = howmany_dims[i];
  return p;
}

fftw_plan fftw_plan_r2r_1d(int n, double* in, double* out, fftw_r2r_kind kind, unsigned flags) {
  fftw_iodim d = {n, 1, 1};
  return fftw_plan_guru_r2r(1, &d, 0, NULL, in, out, &kind, flags);
}

static void one_line(const struct sb200_stub_plan* p, const double* in, double* out) {
  const int n = p->t.n, is = p->t.is, os = p->t.os;
  const long double pi = 3.14159265358979323846264338327950288L;
  long double* tmp = (long double*)malloc(sizeof(long double) * (n > 0 ? n : 1));
  if (p->kind == FFTW_REDFT00) {
    for (int k = 0; k < n; k++) {
      long double s = in[0] + ((k & 1) ? -1.0L : 1.0L) * in[(long)(n - 1) * is];
      for (int j = 1; j < n - 1; j++) s += 2.0L * in[(long)j * is] * cosl(pi * (long double)(((long long)j * k) % (2LL * (n - 1))) / (n - 1));
      tmp[k] = s;
    }
  } else {
    for (int k = 0; k < n; k++) {
      long double s = 0.0L;
      for (int j = 0; j < n; j++)
        s += 2.0L * in[(long)j * is] * sinl(pi * (long double)(((long long)(j + 1) * (k + 1)) % (2LL * (n + 1))) / (n + 1));
      tmp[k] = s;
    }
  }
  for (int k = 0; k < n; k++) out[(long)k * os] = (double)tmp[k];  /* after all reads: in and out may overlap */
  free(tmp);
}

void fftw_execute_r2r(const fftw_plan p, double* in, double* out) {
  int ind[16] = {0};
  for (;;) {
    long io = 0, oo = 0;
    for (int i = 0; i < p->hr; i++) {
      io += (long)ind[i] * p->h[i].is;
      oo += (long)ind[i] * p->h[i].os;
    }
    one_line(p, in + io, out + oo);
    int i = p->hr - 1;
    for (; i >= 0; i--) {
      if (++ind[i] < p->h[i].n) break;
      ind[i] = 0;
    }
    if (i < 0) break;
  }
}

void fftw_destroy_plan(fftw_plan p) { free(p); }
int fftw_import_system_wisdom(void) { return 0; }

/* ---- PETSc stand-in --------------------------------------------------------------------------------------- */
PetscErrorCode PetscPrintf(MPI_Comm comm, const char* fmt, ...) { (void)comm; (void)fmt; return 0; }  /* silent */
PetscErrorCode PetscObjectGetComm(PetscObject o, MPI_Comm* comm) { (void)o; *comm = PETSC_COMM_SELF; return 0; }
PetscErrorCode PetscMallocSetDumpLog(void) { return 0; }
PetscErrorCode PetscMallocDumpLog(FILE* f) { (void)f; return 0; }

PetscErrorCode VecCreateSeq(MPI_Comm comm, PetscInt n, Vec* v) {
  (void)comm;
  *v = (Vec)calloc(1, sizeof(**v));
  (*v)->n = n;
  (*v)->bs = 1;
  (*v)->owns = 1;
  (*v)->a = (double*)calloc(n > 0 ? n : 1, sizeof(double));
  return 0;
}
PetscErrorCode VecCreateSeqWithArray(MPI_Comm comm, PetscInt n, PetscScalar* a, Vec* v) {
  (void)comm;
  *v = (Vec)calloc(1, sizeof(**v));
  (*v)->n = n;
  (*v)->bs = 1;
  (*v)->owns = 0;
  (*v)->a = a;
  return 0;
}
PetscErrorCode VecDuplicate(Vec v, Vec* w) { PetscErrorCode e = VecCreateSeq(0, v->n, w); (*w)->bs = v->bs; return e; }
PetscErrorCode VecDuplicateVecs(Vec v, PetscInt n, Vec** w) {
  *w = (Vec*)malloc(sizeof(Vec) * (n > 0 ? n : 1));
  for (int i = 0; i < n; i++) VecDuplicate(v, &(*w)[i]);
  return 0;
}
PetscErrorCode VecDestroy(Vec v) { if (v) { if (v->owns) free(v->a); free(v); } return 0; }
PetscErrorCode VecDestroyVecs(Vec* w, PetscInt n) { for (int i = 0; i < n; i++) VecDestroy(w[i]); free(w); return 0; }
PetscErrorCode VecSetBlockSize(Vec v, PetscInt bs) { v->bs = bs; return 0; }
PetscErrorCode VecGetSize(Vec v, PetscInt* n) { *n = v->n; return 0; }
PetscErrorCode VecGetArray(Vec v, PetscScalar** a) { *a = v->a; return 0; }
PetscErrorCode VecRestoreArray(Vec v, PetscScalar** a) { (void)v; (void)a; return 0; }
PetscErrorCode VecGetArrays(const Vec* v, PetscInt n, PetscScalar*** a) {
  *a = (PetscScalar**)malloc(sizeof(PetscScalar*) * (n > 0 ? n : 1));
  for (int i = 0; i < n; i++) (*a)[i] = v[i]->a;
  return 0;
}
PetscErrorCode VecRestoreArrays(const Vec* v, PetscInt n, PetscScalar*** a) { (void)v; (void)n; free(*a); *a = NULL; return 0; }
PetscErrorCode VecSet(Vec v, PetscScalar a) { for (int i = 0; i < v->n; i++) v->a[i] = a; return 0; }
PetscErrorCode VecZeroEntries(Vec v) { return VecSet(v, 0.0); }
PetscErrorCode VecCopy(Vec x, Vec y) { memcpy(y->a, x->a, sizeof(double) * x->n); return 0; }
/* BLAS daxpy semantics: y[i] = y[i] + a * x[i], one rounding for the product and one for the sum (no contraction: -ffp-contract=off) */
PetscErrorCode VecAXPY(Vec y, PetscScalar a, Vec x) { for (int i = 0; i < y->n; i++) y->a[i] = y->a[i] + a * x->a[i]; return 0; }
PetscErrorCode VecScale(Vec x, PetscScalar a) { for (int i = 0; i < x->n; i++) x->a[i] *= a; return 0; }
PetscErrorCode VecNorm(Vec x, NormType t, PetscReal* r) {
  double s = 0.0;
  for (int i = 0; i < x->n; i++) {
    const double v = fabs(x->a[i]);
    if (t == NORM_INFINITY) s = v > s ? v : s;
    else if (t == NORM_1) s += v;
    else s += v * v;
  }
  *r = t == NORM_2 ? sqrt(s) : s;
  return 0;
}
PetscErrorCode VecPointwiseDivide(Vec w, Vec x, Vec y) { for (int i = 0; i < w->n; i++) w->a[i] = x->a[i] / y->a[i]; return 0; }
PetscErrorCode VecView(Vec v, PetscViewer vw) { (void)v; (void)vw; return 0; }

PetscErrorCode ISCreateGeneral(MPI_Comm comm, PetscInt n, const PetscInt* idx, IS* is) {
  (void)comm;
  *is = (IS)calloc(1, sizeof(**is));
  (*is)->n = n;
  (*is)->idx = (int*)malloc(sizeof(int) * (n > 0 ? n : 1));
  memcpy((*is)->idx, idx, sizeof(int) * n);
  return 0;
}
PetscErrorCode ISDestroy(IS is) { if (is) { free(is->idx); free(is); } return 0; }
PetscErrorCode ISGetIndices(IS is, const PetscInt** idx) { *idx = is->idx; return 0; }
PetscErrorCode ISRestoreIndices(IS is, const PetscInt** idx) { (void)is; (void)idx; return 0; }
PetscErrorCode ISView(IS is, PetscViewer vw) { (void)is; (void)vw; return 0; }

/* VecScatterCreate(x, ix, y, iy): entry i moves x[ix[i]] to y[iy[i]]; a NULL index set stands for all entries 0..n-1 */
PetscErrorCode VecScatterCreate(Vec x, IS ix, Vec y, IS iy, VecScatter* s) {
  const int n = ix ? ix->n : (iy ? iy->n : (x->n < y->n ? x->n : y->n));
  *s = (VecScatter)calloc(1, sizeof(**s));
  (*s)->n = n;
  (*s)->from = (int*)malloc(sizeof(int) * (n > 0 ? n : 1));
  (*s)->to = (int*)malloc(sizeof(int) * (n > 0 ? n : 1));
  for (int i = 0; i < n; i++) {
    (*s)->from[i] = ix ? ix->idx[i] : i;
    (*s)->to[i] = iy ? iy->idx[i] : i;
  }
  return 0;
}
PetscErrorCode VecScatterBegin(VecScatter s, Vec x, Vec y, InsertMode im, ScatterMode sm) {
  (void)sm;
  for (int i = 0; i < s->n; i++) {
    if (im == ADD_VALUES) y->a[s->to[i]] += x->a[s->from[i]];
    else y->a[s->to[i]] = x->a[s->from[i]];
  }
  return 0;
}
PetscErrorCode VecScatterEnd(VecScatter s, Vec x, Vec y, InsertMode im, ScatterMode sm) { (void)s; (void)x; (void)y; (void)im; (void)sm; return 0; }
PetscErrorCode VecScatterDestroy(VecScatter s) { if (s) { free(s->from); free(s->to); free(s); } return 0; }

PetscErrorCode MatCreateShell(MPI_Comm comm, PetscInt m, PetscInt n, PetscInt M, PetscInt N, void* ctx, Mat* A) {
  (void)comm; (void)M; (void)N;
  *A = (Mat)calloc(1, sizeof(**A));
  (*A)->ctx = ctx;
  (*A)->m = m;
  (*A)->n = n;
  return 0;
}
PetscErrorCode MatShellSetOperation(Mat A, MatOperation op, void (*f)(void)) {
  if (op == MATOP_MULT) A->mult = (PetscErrorCode(*)(Mat, Vec, Vec))f;
  else if (op == MATOP_DESTROY) A->destroy = (PetscErrorCode(*)(Mat))f;
  else if (op == MATOP_GET_DIAGONAL) A->getdiag = (PetscErrorCode(*)(Mat, Vec))f;
  return 0;
}
PetscErrorCode MatShellGetContext(Mat A, void** ctx) { *ctx = A->ctx; return 0; }
PetscErrorCode MatMult(Mat A, Vec x, Vec y) {
  if (!A->mult) SETERRQ(56, "MatMult: no MULT operation");
  return A->mult(A, x, y);
}
PetscErrorCode MatDestroy(Mat A) {
  if (!A) return 0;
  if (A->destroy) { PetscErrorCode e = A->destroy(A); if (e) return e; }
  free(A->ti); free(A->tj); free(A->tv);
  free(A);
  return 0;
}
PetscErrorCode MatGetSize(Mat A, PetscInt* m, PetscInt* n) { if (m) *m = A->m; if (n) *n = A->n; return 0; }
PetscErrorCode MatCreate(MPI_Comm comm, Mat* A) { return MatCreateShell(comm, 0, 0, 0, 0, NULL, A); }
PetscErrorCode MatSetSizes(Mat A, PetscInt m, PetscInt n, PetscInt M, PetscInt N) { (void)M; (void)N; A->m = m; A->n = n; return 0; }
PetscErrorCode MatSetType(Mat A, const char* type) { (void)A; (void)type; return 0; }
PetscErrorCode MatSetFromOptions(Mat A) { (void)A; return 0; }
PetscErrorCode MatCreateSeqAIJ(MPI_Comm comm, PetscInt m, PetscInt n, PetscInt nz, const PetscInt* nnz, Mat* A) {
  (void)nz; (void)nnz;
  return MatCreateShell(comm, m, n, m, n, NULL, A);
}
PetscErrorCode MatSetValues(Mat A, PetscInt m, const PetscInt* im, PetscInt n, const PetscInt* in, const PetscScalar* v, InsertMode mode) {
  (void)mode;
  for (int i = 0; i < m; i++)
    for (int j = 0; j < n; j++) {
      if (im[i] < 0 || in[j] < 0) continue;  /* negative indices are ignored (PETSc semantics the reference relies on) */
      if (A->nt == A->cap) {
        A->cap = A->cap ? 2 * A->cap : 1024;
        A->ti = (int*)realloc(A->ti, sizeof(int) * A->cap);
        A->tj = (int*)realloc(A->tj, sizeof(int) * A->cap);
        A->tv = (double*)realloc(A->tv, sizeof(double) * A->cap);
      }
      A->ti[A->nt] = im[i];
      A->tj[A->nt] = in[j];
      A->tv[A->nt] = v[i * n + j];
      A->nt++;
    }
  return 0;
}
PetscErrorCode MatAssemblyBegin(Mat A, MatAssemblyType t) { (void)A; (void)t; return 0; }
PetscErrorCode MatAssemblyEnd(Mat A, MatAssemblyType t) { (void)A; (void)t; return 0; }

/* ---- inert solver objects and options ---------------------------------------------------------------------------- */
static const char* const reasons_[] = {"(stub)"};
const char* const* SNESConvergedReasons = reasons_;
#define MAXOPT 32
static char opt_name[MAXOPT][32];
static double opt_val[MAXOPT];
static int opt_arr[MAXOPT][16], opt_arr_n[MAXOPT];
static int nopt = 0;
static int opt_find(const char* name, int create) {
  for (int i = 0; i < nopt; i++)
    if (!strcmp(opt_name[i], name)) return i;
  if (!create || nopt >= MAXOPT) return -1;
  snprintf(opt_name[nopt], 32, "%s", name);
  opt_arr_n[nopt] = 0;
  return nopt++;
}
void ref_clear_options(void) { nopt = 0; }
void ref_set_option_real(const char* name, double v) { const int i = opt_find(name, 1); if (i >= 0) opt_val[i] = v; }
void ref_set_option_int(const char* name, int v) { const int i = opt_find(name, 1); if (i >= 0) opt_val[i] = (double)v; }
void ref_set_option_intarray(const char* name, int n, const int* v) {
  const int i = opt_find(name, 1);
  if (i < 0) return;
  opt_arr_n[i] = n < 16 ? n : 16;
  for (int k = 0; k < opt_arr_n[i]; k++) opt_arr[i][k] = v[k];
}
PetscErrorCode PetscInitialize(int* argc, char*** args, const char* file, const char* help) { (void)argc; (void)args; (void)file; (void)help; return 0; }
PetscErrorCode PetscFinalize(void) { return 0; }
PetscErrorCode PetscOptionsIntArray(const char* o, const char* t, const char* m, PetscInt* v, PetscInt* n, PetscTruth* set) {
  (void)t; (void)m;
  const int i = opt_find(o, 0);
  if (i < 0 || !opt_arr_n[i]) { if (set) *set = PETSC_FALSE; return 0; }
  const int cnt = opt_arr_n[i] < *n ? opt_arr_n[i] : *n;
  for (int k = 0; k < cnt; k++) v[k] = opt_arr[i][k];
  *n = cnt;
  if (set) *set = PETSC_TRUE;
  return 0;
}
PetscErrorCode PetscOptionsInt(const char* o, const char* t, const char* m, PetscInt def, PetscInt* v, PetscTruth* set) {
  (void)t; (void)m;
  const int i = opt_find(o, 0);
  *v = i >= 0 ? (PetscInt)opt_val[i] : def;
  if (set) *set = i >= 0 ? PETSC_TRUE : PETSC_FALSE;
  return 0;
}
PetscErrorCode PetscOptionsReal(const char* o, const char* t, const char* m, PetscReal def, PetscReal* v, PetscTruth* set) {
  (void)t; (void)m;
  const int i = opt_find(o, 0);
  *v = i >= 0 ? opt_val[i] : def;
  if (set) *set = i >= 0 ? PETSC_TRUE : PETSC_FALSE;
  return 0;
}
PetscErrorCode PetscOptionsGetReal(const char* pre, const char* name, PetscReal* v, PetscTruth* set) {
  (void)pre;
  const int i = opt_find(name, 0);
  if (i >= 0) *v = opt_val[i];
  if (set) *set = i >= 0 ? PETSC_TRUE : PETSC_FALSE;
  return 0;
}
PetscErrorCode PetscOptionsGetInt(const char* pre, const char* name, PetscInt* v, PetscTruth* set) {
  (void)pre;
  const int i = opt_find(name, 0);
  if (i >= 0) *v = (PetscInt)opt_val[i];
  if (set) *set = i >= 0 ? PETSC_TRUE : PETSC_FALSE;
  return 0;
}
PetscErrorCode PetscOptionsHasName(const char* pre, const char* name, PetscTruth* set) { (void)pre; *set = opt_find(name, 0) >= 0 ? PETSC_TRUE : PETSC_FALSE; return 0; }

/* ---- more Vec / Mat operations used by stokes.C --------------------------------------------------------------------- */
PetscErrorCode VecCreate(MPI_Comm comm, Vec* v) { (void)comm; *v = (Vec)calloc(1, sizeof(**v)); (*v)->bs = 1; (*v)->owns = 1; return 0; }
PetscErrorCode VecSetSizes(Vec v, PetscInt n, PetscInt N) {
  (void)N;
  free(v->a);
  v->n = n;
  v->a = (double*)calloc(n > 0 ? n : 1, sizeof(double));
  return 0;
}
PetscErrorCode VecSetFromOptions(Vec v) { (void)v; return 0; }
PetscErrorCode VecStrideGather(Vec v, PetscInt start, Vec s, InsertMode m) {
  const int bs = v->bs;
  for (int i = 0; i < s->n; i++) {
    if (m == ADD_VALUES) s->a[i] += v->a[i * bs + start];
    else s->a[i] = v->a[i * bs + start];
  }
  return 0;
}
PetscErrorCode VecStrideScatter(Vec s, PetscInt start, Vec v, InsertMode m) {
  const int bs = v->bs;
  for (int i = 0; i < s->n; i++) {
    if (m == ADD_VALUES) v->a[i * bs + start] += s->a[i];
    else v->a[i * bs + start] = s->a[i];
  }
  return 0;
}
PetscErrorCode VecMin(Vec v, PetscInt* p, PetscReal* val) {
  int k = 0;
  for (int i = 1; i < v->n; i++) if (v->a[i] < v->a[k]) k = i;
  if (p) *p = k;
  *val = v->n ? v->a[k] : 0.0;
  return 0;
}
PetscErrorCode VecMax(Vec v, PetscInt* p, PetscReal* val) {
  int k = 0;
  for (int i = 1; i < v->n; i++) if (v->a[i] > v->a[k]) k = i;
  if (p) *p = k;
  *val = v->n ? v->a[k] : 0.0;
  return 0;
}
PetscErrorCode VecReciprocal(Vec v) { for (int i = 0; i < v->n; i++) if (v->a[i] != 0.0) v->a[i] = 1.0 / v->a[i]; return 0; }
PetscErrorCode VecNormalize(Vec v, PetscReal* val) {
  PetscReal nrm;
  VecNorm(v, NORM_2, &nrm);
  if (nrm > 0) VecScale(v, 1.0 / nrm);
  if (val) *val = nrm;
  return 0;
}
PetscErrorCode PetscIntView(PetscInt n, const PetscInt* idx, PetscViewer vw) { (void)n; (void)idx; (void)vw; return 0; }
PetscErrorCode PetscRealView(PetscInt n, const PetscReal* idx, PetscViewer vw) { (void)n; (void)idx; (void)vw; return 0; }
PetscErrorCode PetscViewerCreate(MPI_Comm comm, PetscViewer* v) { (void)comm; *v = NULL; return 0; }
PetscErrorCode PetscViewerSetType(PetscViewer v, const char* t) { (void)v; (void)t; return 0; }
PetscErrorCode PetscViewerSetFormat(PetscViewer v, int f) { (void)v; (void)f; return 0; }
PetscErrorCode PetscViewerFileSetName(PetscViewer v, const char* name) { (void)v; (void)name; return 0; }
PetscErrorCode PetscViewerASCIIPrintf(PetscViewer v, const char* fmt, ...) { (void)v; (void)fmt; return 0; }
PetscErrorCode PetscViewerDestroy(PetscViewer v) { (void)v; return 0; }
PetscErrorCode MatZeroEntries(Mat A) { A->nt = 0; return 0; }
PetscErrorCode MatDiagonalScale(Mat A, Vec l, Vec r) {
  for (int k = 0; k < A->nt; k++) {
    if (l) A->tv[k] *= l->a[A->ti[k]];
    if (r) A->tv[k] *= r->a[A->tj[k]];
  }
  return 0;
}
PetscErrorCode MatView(Mat A, PetscViewer vw) { (void)A; (void)vw; return 0; }
PetscErrorCode MatNullSpaceCreate(MPI_Comm comm, PetscTruth has_const, PetscInt n, const Vec* vecs, MatNullSpace* ns) {
  (void)comm;
  *ns = (MatNullSpace)calloc(1, sizeof(**ns));
  (*ns)->has_const = has_const;
  (*ns)->n = n;
  (*ns)->vecs = (Vec*)malloc(sizeof(Vec) * (n > 0 ? n : 1));
  for (int i = 0; i < n; i++) (*ns)->vecs[i] = vecs[i];
  return 0;
}
PetscErrorCode MatNullSpaceDestroy(MatNullSpace ns) { if (ns) { free(ns->vecs); free(ns); } return 0; }
PetscErrorCode MatNullSpaceRemove(MatNullSpace ns, Vec v, Vec* out) {
  (void)out;
  if (ns->has_const && v->n) {
    double s = 0.0;
    for (int i = 0; i < v->n; i++) s += v->a[i];
    s /= v->n;
    for (int i = 0; i < v->n; i++) v->a[i] -= s;
  }
  for (int k = 0; k < ns->n; k++) {
    double dot = 0.0;
    for (int i = 0; i < v->n; i++) dot += v->a[i] * ns->vecs[k]->a[i];
    for (int i = 0; i < v->n; i++) v->a[i] -= dot * ns->vecs[k]->a[i];
  }
  return 0;
}
PetscErrorCode MatNullSpaceTest(MatNullSpace ns, Mat A, PetscTruth* isNull) {
  *isNull = PETSC_TRUE;
  for (int k = 0; k < ns->n; k++) {
    Vec y;
    PetscReal nrm;
    VecDuplicate(ns->vecs[k], &y);
    PetscErrorCode e = MatMult(A, ns->vecs[k], y);
    if (e) return e;
    VecNorm(y, NORM_2, &nrm);
    VecDestroy(y);
    if (nrm > 1e-7) *isNull = PETSC_FALSE;
  }
  return 0;
}
PetscErrorCode MatGetColoring(Mat A, const char* type, ISColoring* c) { (void)A; (void)type; *c = NULL; return 0; }
PetscErrorCode MatFDColoringCreate(Mat A, ISColoring c, MatFDColoring* f) { (void)A; (void)c; *f = NULL; return 0; }
PetscErrorCode MatFDColoringSetFunction(MatFDColoring f, PetscErrorCode (*fn)(void), void* ctx) { (void)f; (void)fn; (void)ctx; return 0; }
PetscErrorCode MatFDColoringSetFromOptions(MatFDColoring f) { (void)f; return 0; }
PetscErrorCode MatFDColoringApply(Mat A, MatFDColoring f, Vec x, MatStructure* flag, void* ctx) {
  (void)A; (void)f; (void)x; (void)flag; (void)ctx;
  SETERRQ(PETSC_ERR_SUP, "MatFDColoringApply is not part of the stand-in (-pcvel 2)");
}
PetscErrorCode KSPCreate(MPI_Comm comm, KSP* k) { (void)comm; *k = (KSP)calloc(1, sizeof(**k)); return 0; }
PetscErrorCode KSPDestroy(KSP k) { free(k); return 0; }
PetscErrorCode KSPSetOperators(KSP k, Mat A, Mat P, MatStructure f) { (void)k; (void)A; (void)P; (void)f; return 0; }
PetscErrorCode KSPSetOptionsPrefix(KSP k, const char* p) { (void)k; (void)p; return 0; }
PetscErrorCode KSPSetFromOptions(KSP k) { (void)k; return 0; }
PetscErrorCode KSPSetNullSpace(KSP k, MatNullSpace ns) { (void)k; (void)ns; return 0; }
PetscErrorCode KSPSolve(KSP k, Vec b, Vec x) { (void)k; return VecCopy(b, x); }
PetscErrorCode PCShellSetContext(PC pc, void* ctx) { pc->ctx = ctx; return 0; }
PetscErrorCode PCShellGetContext(PC pc, void** ctx) { *ctx = pc->ctx; return 0; }
PetscErrorCode PCShellSetSetUp(PC pc, PetscErrorCode (*f)(PC)) { (void)pc; (void)f; return 0; }
PetscErrorCode PCShellSetApply(PC pc, PetscErrorCode (*f)(PC, Vec, Vec)) { (void)pc; (void)f; return 0; }
PetscErrorCode SNESCreate(MPI_Comm comm, SNES* s) { (void)comm; *s = (SNES)calloc(1, sizeof(**s)); return 0; }
PetscErrorCode SNESDestroy(SNES s) { free(s); return 0; }
PetscErrorCode SNESSetJacobian(SNES s, Mat A, Mat P, PetscErrorCode (*f)(SNES, Vec, Mat*, Mat*, MatStructure*, void*), void* ctx) {
  (void)s; (void)A; (void)P; (void)f; (void)ctx; return 0;
}
PetscErrorCode SNESSetFunction(SNES s, Vec r, PetscErrorCode (*f)(SNES, Vec, Vec, void*), void* ctx) { (void)s; (void)r; (void)f; (void)ctx; return 0; }
PetscErrorCode SNESSetApplicationContext(SNES s, void* ctx) { s->appctx = ctx; return 0; }
PetscErrorCode SNESGetApplicationContext(SNES s, void** ctx) { *ctx = s->appctx; return 0; }
PetscErrorCode SNESGetKSP(SNES s, KSP* k) { (void)s; static struct _stub_KSP kk; *k = &kk; return 0; }
PetscErrorCode SNESSetFromOptions(SNES s) { (void)s; return 0; }
PetscErrorCode SNESSolve(SNES s, Vec b, Vec x) { (void)s; (void)b; (void)x; SETERRQ(56, "SNESSolve: the solver stack is not part of the stand-in"); }
PetscErrorCode SNESGetIterationNumber(SNES s, PetscInt* its) { (void)s; *its = 0; return 0; }
PetscErrorCode SNESGetConvergedReason(SNES s, SNESConvergedReason* r) { (void)s; *r = 0; return 0; }
PetscErrorCode KSPSetType(KSP k, const char* t) { (void)k; (void)t; return 0; }
PetscErrorCode KSPGetPC(KSP k, PC* pc) { (void)k; static struct _stub_PC pp; *pc = &pp; return 0; }
PetscErrorCode KSPGetIterationNumber(KSP k, PetscInt* its) { (void)k; *its = 0; return 0; }
PetscErrorCode PCSetType(PC pc, const char* t) { (void)pc; (void)t; return 0; }
PetscErrorCode PCFactorSetLevels(PC pc, PetscInt l) { (void)pc; (void)l; return 0; }

/* ---- ctypes entry points: the reference's own functions, driven the way cheb.c drives them ---------------------- */
PetscErrorCode MatCreateCheb(MPI_Comm comm, int rank, int tr, int* dims, unsigned flag, Vec vx, Vec vy, Mat* A);
PetscErrorCode MatCreateChebD1(MPI_Comm comm, Vec vx, Vec vy, unsigned flag, Mat* A);

/* y = ChebMult(MatCreateCheb(rank, tr, dims), x); returns the PetscErrorCode of the first failing call */
int ref_cheb_mult(int rank, int tr, int* dims, int n, double* x, double* y) {
  struct _stub_Vec vx = {n, 1, 0, x}, vy = {n, 1, 0, y};
  Mat A = NULL;
  int e = MatCreateCheb(PETSC_COMM_SELF, rank, tr, dims, FFTW_ESTIMATE, &vx, &vy, &A);
  if (e) return e;
  e = A->mult(A, &vx, &vy);
  if (e) return e;
  return MatDestroy(A);
}

int ref_chebd1_mult(int n, double* x, double* y) {
  struct _stub_Vec vx = {n, 1, 0, x}, vy = {n, 1, 0, y};
  Mat A = NULL;
  int e = MatCreateChebD1(PETSC_COMM_SELF, &vx, &vy, FFTW_ESTIMATE, &A);
  if (e) return e;
  e = A->mult(A, &vx, &vy);
  if (e) return e;
  return MatDestroy(A);
}

const char* ref_last_error(void) { return sb200_stub_last_error; }
