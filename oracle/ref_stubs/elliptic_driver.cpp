// ctypes entry points that drive the REFERENCE's own elliptic.C (MatCreate_Elliptic, SetupBC, CreateExactSolution, FormFunction,
// MatMult_Elliptic, FormJacobian), compiled where it lies (textual include, never copied) against the PETSc / FFTW stand-ins in
// this directory.  The reference's main() is kept out of the way by renaming it.  Test infrastructure.
#define main ref_elliptic_main
#include "elliptic.C"
#undef main

struct RefElliptic {
  AppCtx ac;
  SNES snes;
  Vec u, u2;   // exact solution and its forcing (global vectors)
  Mat P;       // FormJacobian's preconditioning matrix (triplets)
  MatElliptic* c;
};

extern "C" {

void* ref_elliptic_create(int d, int* dim, double gamma, double exponent, int exact, double cos_scale) {
  RefElliptic* r = new RefElliptic();
  r->ac.d = d;
  r->ac.dim = (PetscInt*)malloc(sizeof(PetscInt) * d);
  for (int i = 0; i < d; i++) r->ac.dim[i] = dim[i];
  r->ac.exact = exact;
  r->ac.gamma = gamma;
  r->ac.exponent = exponent;
  r->ac.debug = 0;
  ref_set_option_real("-cos_scale", cos_scale);
  if (MatCreate_Elliptic(PETSC_COMM_WORLD, d, r->ac.dim, FFTW_ESTIMATE, DirichletBdy, &r->u, &r->ac.A)) return NULL;  // elliptic.C:158
  MatShellGetContext(r->ac.A, (void**)&r->c);
  VecDuplicate(r->u, &r->u2);
  VecDuplicate(r->u, &r->ac.b);
  SNESCreate(PETSC_COMM_WORLD, &r->snes);
  SNESSetApplicationContext(r->snes, &r->ac);
  PetscInt n;
  VecGetSize(r->u, &n);
  MatCreateSeqAIJ(PETSC_COMM_SELF, n, n, 1 + 2 * d, PETSC_NULL, &r->P);
  if (exact >= 0 && CreateExactSolution(r->snes, r->u, r->u2)) return NULL;  // elliptic.C:185
  return r;
}

void ref_elliptic_sizes(void* h, long long* m, long long* g, long long* nd) {
  RefElliptic* r = (RefElliptic*)h;
  *m = r->c->w[0]->n;
  *g = r->u->n;
  *nd = r->c->dirichlet->n;
}

// which: 0 exact u, 1 exact forcing u2, 2 dirichlet values, 3 rhs b, 4 eta, 5 deta, 6+k gradu[k], 100 coordinates (m*d)
int ref_elliptic_get(void* h, int which, double* out) {
  RefElliptic* r = (RefElliptic*)h;
  Vec v = NULL;
  if (which == 0) v = r->u;
  else if (which == 1) v = r->u2;
  else if (which == 2) v = r->c->dirichlet;
  else if (which == 3) v = r->ac.b;
  else if (which == 4) v = r->c->eta;
  else if (which == 5) v = r->c->deta;
  else if (which >= 6 && which < 6 + r->c->d) v = r->c->gradu[which - 6];
  else if (which == 100) v = r->c->x;
  if (!v) return 1;
  memcpy(out, v->a, sizeof(double) * v->n);
  return 0;
}

int ref_elliptic_set_dirichlet_rhs(void* h, const double* dir, const double* b) {
  RefElliptic* r = (RefElliptic*)h;
  if (dir) memcpy(r->c->dirichlet->a, dir, sizeof(double) * r->c->dirichlet->n);
  if (b) memcpy(r->ac.b->a, b, sizeof(double) * r->ac.b->n);
  return 0;
}

int ref_elliptic_function(void* h, double* U, double* F) {
  RefElliptic* r = (RefElliptic*)h;
  struct _stub_Vec vu = {r->u->n, 1, 0, U}, vf = {r->u->n, 1, 0, F};
  return FormFunction(r->snes, &vu, &vf, &r->ac);  // elliptic.C:481
}

int ref_elliptic_matmult(void* h, double* U, double* V) {
  RefElliptic* r = (RefElliptic*)h;
  struct _stub_Vec vu = {r->u->n, 1, 0, U}, vv = {r->u->n, 1, 0, V};
  return MatMult(r->ac.A, &vu, &vv);  // MatMult_Elliptic, elliptic.C:297
}

// FormJacobian (elliptic.C:537): returns the number of recorded entries; rows/cols/vals may be NULL to query the count
int ref_elliptic_jacobian(void* h, int cap, int* rows, int* cols, double* vals) {
  RefElliptic* r = (RefElliptic*)h;
  r->P->nt = 0;
  MatStructure flag;
  struct _stub_Vec dummy = {r->u->n, 1, 0, r->u->a};
  if (FormJacobian(r->snes, &dummy, &r->ac.A, &r->P, &flag, &r->ac)) return -1;
  const int n = r->P->nt;
  for (int i = 0; i < n && i < cap; i++) {
    rows[i] = r->P->ti[i];
    cols[i] = r->P->tj[i];
    vals[i] = r->P->tv[i];
  }
  return n;
}

void ref_elliptic_destroy(void* h) {
  RefElliptic* r = (RefElliptic*)h;
  if (!r) return;
  MatDestroy(r->ac.A);
  MatDestroy(r->P);
  VecDestroy(r->u);
  VecDestroy(r->u2);
  VecDestroy(r->ac.b);
  SNESDestroy(r->snes);
  free(r->ac.dim);
  delete r;
}

}  // extern "C"
