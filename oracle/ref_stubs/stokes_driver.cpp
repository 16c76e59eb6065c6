// ctypes entry points that drive the REFERENCE's own stokes.C (StokesCreate, StokesSetupDomain, StokesCreateExactSolution,
// StokesFunction, StokesMatMult{,VV,PV,VP,Schur}, StokesMatGetDiagonalSchur, StokesPressureReduceOrder, StokesPCSetUp0), compiled
// where it lies (textual include, never copied) against the PETSc / FFTW / CppAD stand-ins in this directory.  The reference's
// main() is renamed out of the way; its option parsing is fed through the stand-in's option table.  Test infrastructure.
#define main ref_stokes_main
#include "stokes.C"
#undef main

struct RefStokes {
  StokesCtx* c;
  Mat A;
  Vec x, u, u2;
  SNES snes;
  struct _stub_PC pc;
};

extern "C" {

void* ref_stokes_create(int d, int* dim, int exact, int rheology, double hardness, double exponent, double eps, double gamma0) {
  ref_clear_options();
  ref_set_option_intarray("-dim", d, dim);
  ref_set_option_int("-exact", exact);
  ref_set_option_int("-boundary", 0);
  ref_set_option_int("-rheology", rheology);
  ref_set_option_real("-hardness", hardness);
  ref_set_option_real("-exponent", exponent);
  ref_set_option_real("-eps", eps);
  ref_set_option_real("-gamma0", gamma0);
  RefStokes* r = new RefStokes();
  if (StokesCreate(PETSC_COMM_WORLD, &r->A, &r->x, &r->c)) { delete r; return NULL; }  // stokes.C:139
  VecDuplicate(r->x, &r->u);
  VecDuplicate(r->x, &r->u2);
  SNESCreate(PETSC_COMM_WORLD, &r->snes);
  SNESSetApplicationContext(r->snes, r->c);
  r->pc.ctx = r->c;
  if (StokesCreateExactSolution(r->snes, r->u, r->u2)) { delete r; return NULL; }  // stokes.C:190
  return r;
}

void ref_stokes_sizes(void* h, long long* m, long long* g, long long* gp, long long* gv, long long* dv) {
  RefStokes* r = (RefStokes*)h;
  *m = r->c->eta->n;
  *g = r->x->n;
  *gp = r->c->pG0->n;
  *gv = r->c->vG0->n;
  *dv = r->c->dirichlet->n;
}

// which: 0 exact u, 1 forcing u2, 2 dirichlet, 3 force, 4 eta, 5 deta, 6+j strain[j] (m*d)
int ref_stokes_get(void* h, int which, double* out) {
  RefStokes* r = (RefStokes*)h;
  Vec v = NULL;
  if (which == 0) v = r->u;
  else if (which == 1) v = r->u2;
  else if (which == 2) v = r->c->dirichlet;
  else if (which == 3) v = r->c->force;
  else if (which == 4) v = r->c->eta;
  else if (which == 5) v = r->c->deta;
  else if (which >= 6 && which < 6 + r->c->numDims) v = r->c->strain[which - 6];
  if (!v) return 1;
  memcpy(out, v->a, sizeof(double) * v->n);
  return 0;
}

void ref_stokes_set_rheology(void* h, double exponent, double regularization) {  // the continuation loop, stokes.C:218-219
  RefStokes* r = (RefStokes*)h;
  r->c->options->exponent = exponent;
  r->c->options->regularization = regularization;
}

static int apply(Mat A, int nin, double* x, int nout, double* y) {
  struct _stub_Vec vx = {nin, 1, 0, x}, vy = {nout, 1, 0, y};
  return MatMult(A, &vx, &vy);
}

int ref_stokes_function(void* h, double* x, double* y) {
  RefStokes* r = (RefStokes*)h;
  struct _stub_Vec vx = {r->x->n, 1, 0, x}, vy = {r->x->n, 1, 0, y};
  return StokesFunction(r->snes, &vx, &vy, r->c);  // stokes.C:680
}
int ref_stokes_matmult(void* h, double* x, double* y) { RefStokes* r = (RefStokes*)h; return apply(r->A, r->x->n, x, r->x->n, y); }
int ref_stokes_matmult_vv(void* h, double* x, double* y) { RefStokes* r = (RefStokes*)h; return apply(r->c->MatVV, r->c->vG0->n, x, r->c->vG0->n, y); }
int ref_stokes_matmult_pv(void* h, double* x, double* y) { RefStokes* r = (RefStokes*)h; return apply(r->c->MatPV, r->c->vG0->n, x, r->c->pG0->n, y); }
int ref_stokes_matmult_vp(void* h, double* x, double* y) { RefStokes* r = (RefStokes*)h; return apply(r->c->MatVP, r->c->pG0->n, x, r->c->vG0->n, y); }
// StokesMatMultSchur with the stand-in's KSPSolve (the identity) as the inner velocity solve
int ref_stokes_matmult_schur_identity(void* h, double* x, double* y) { RefStokes* r = (RefStokes*)h; return apply(r->c->MatSchur, r->c->pG0->n, x, r->c->pG0->n, y); }
int ref_stokes_diag_schur(void* h, double* y) {
  RefStokes* r = (RefStokes*)h;
  struct _stub_Vec vy = {r->c->pG0->n, 1, 0, y};
  return StokesMatGetDiagonalSchur(r->c->MatSchur, &vy);
}
int ref_stokes_reduce_order(void* h, double* pL) {
  RefStokes* r = (RefStokes*)h;
  struct _stub_Vec vp = {r->c->eta->n, 1, 0, pL};
  return StokesPressureReduceOrder(&vp, r->c);  // stokes.C:1029
}
// StokesPCSetUp0 (stokes.C:1160): the finite-difference velocity preconditioning matrix as triplets
int ref_stokes_pc_matrix(void* h, int cap, int* rows, int* cols, double* vals) {
  RefStokes* r = (RefStokes*)h;
  r->c->MatVVPC->nt = 0;
  if (StokesPCSetUp0(&r->pc)) return -1;
  const int n = r->c->MatVVPC->nt;
  for (int i = 0; i < n && i < cap; i++) {
    rows[i] = r->c->MatVVPC->ti[i];
    cols[i] = r->c->MatVVPC->tj[i];
    vals[i] = r->c->MatVVPC->tv[i];
  }
  return n;
}

void ref_stokes_destroy(void* h) {
  RefStokes* r = (RefStokes*)h;
  if (!r) return;
  VecDestroy(r->u);
  VecDestroy(r->u2);
  SNESDestroy(r->snes);
  delete r;  // the context itself is left to the process (StokesDestroy frees objects the stand-in never created)
}

const char* ref_stokes_last_error(void) { return sb200_stub_last_error; }

}  // extern "C"
