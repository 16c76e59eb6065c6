// StokesPCApply0..3 (stokes.C:1714-1817) and the inner solver stack they use (stokes.C:328-341), device resident: host code that
// only sequences C-ABI calls (shells, sb200_ksp, the vector helpers) on device vectors it owns - no vector crosses PCIe here.
// The reference's PCShell routines work on the context's shared vG0 / vG1 / pG0 / pG1 ("KSPSolve tampers with vG0 and vG1",
// stokes.C:1731); the work vectors below play those roles, one set per nesting level so an inner solve cannot clobber its caller.
//
// Inner solves are PETSc's default KSP: GMRES(30), LEFT preconditioning, zero initial guess, convergence on the preconditioned
// residual.  With the flexible right-preconditioned sb200_ksp that is GMRES on the operator M^-1 A with right-hand side M^-1 b
// and no PC; a null space (KSPSetNullSpace) is projected out after every preconditioner application.
#include <cstddef>
#include <string>

#include "../../include/spectral_b200.h"

namespace sb200 {
void set_last_error(const std::string& msg);
}

#define CHK(expr)      \
  do {                 \
    int _e = (expr);   \
    if (_e) return _e; \
  } while (0)

struct sb200_saddle {
  sb200_stokes* s = nullptr;
  int type = 0, d = 0;
  long long m = 0, g = 0, gp = 0, gv = 0, dv = 0;
  sb200_apply_fn vel_pc = nullptr, svel_pc = nullptr;
  void* vel_ctx = nullptr;
  void* svel_ctx = nullptr;
  double vel_rtol = 1e-5, schur_rtol = 1e-5, svel_rtol = 1e-5;       // PETSc's KSP defaults; KSPSchurVelocity has its own
  int vel_maxits = 10000, schur_maxits = 10000, svel_maxits = 10000;  // "svel_" options prefix (stokes.C:338-341)
  bool svel_preonly = false;
  sb200_ksp* kvel = nullptr;    // KSPVelocity      (stokes.C:334-337)
  sb200_ksp* kschur = nullptr;  // KSPSchur         (stokes.C:328-333)
  sb200_ksp* ksvel = nullptr;   // KSPSchurVelocity (stokes.C:338-341), unless preonly
  int r_vel = 0, r_schur = 0, r_svel = 0;  // Krylov space each solver was created with
  long long its_vel = 0, its_schur = 0;
  // device work vectors by role (enums below), the Jacobi "diagonal", one reduction scratch
  double* v[8] = {};
  double* p[5] = {};
  double* diag = nullptr;
  double* scratch = nullptr;
  double* sums = nullptr;  // {sum, count} of a mean removal (summed over the ranks on a slab partition)
  int rank = 0, nranks = 1;  // slab partition of the Stokes context (1 = single GPU)
};

namespace {

enum { XV = 0, V1 = 1, TV = 2, UV = 3, PBV = 4, WV = 5, SV0 = 6, SV1 = 7 };  // velocity-sized roles
enum { XP = 0, P1 = 1, TP = 2, WP = 3, SPB = 4 };                            // pressure-sized roles

int pc_or_copy(sb200_apply_fn pc, void* ctx, long long n, const double* r, double* z, void* stream) {
  if (pc) return pc(ctx, r, z, stream);
  return sb200_memcpy_d2d(z, r, (size_t)n * sizeof(double), stream);  // PCNONE
}

// --- operators of the inner Krylov solves: x -> M^-1 A x -------------------------------------------------------------------
int op_velocity(void* ctx, const double* x, double* y, void* stream) {  // KSPVelocity: A = MatVV, M = -vel_pc on MatVVPC
  sb200_saddle* P = (sb200_saddle*)ctx;
  CHK(sb200_stokes_matmult_vv(P->s, x, P->v[WV], stream));
  return pc_or_copy(P->vel_pc, P->vel_ctx, P->gv, P->v[WV], y, stream);
}

int op_schur_velocity(void* ctx, const double* x, double* y, void* stream) {  // KSPSchurVelocity, when it is not preonly
  sb200_saddle* P = (sb200_saddle*)ctx;
  CHK(sb200_stokes_matmult_vv(P->s, x, P->v[SV1], stream));
  return pc_or_copy(P->svel_pc, P->svel_ctx, P->gv, P->v[SV1], y, stream);
}

// GMRES(30) that may take at most maxits < 30 iterations never restarts, so it only needs maxits + 1 basis vectors: the solver is
// (re)created with restart = min(30, maxits) - identical iterates, a sixth of the memory for the usual -vel_ksp_max_it 4.
// On a slab partition the solvers are created by sb200_saddle_prepare (their arenas are exchanged between the ranks afterwards) and
// may not be re-created behind the peers' backs.
int ensure_ksp(sb200_saddle* P, sb200_ksp** k, int* have, long long n, int maxits, bool may_create) {
  const int want = maxits < 1 ? 1 : (maxits < 30 ? maxits : 30);
  if (*k && *have == want) return 0;
  if (!may_create) {
    sb200::set_last_error("slab-partitioned saddle PC: call sb200_saddle_prepare and attach the peers after changing the inner solvers' settings");
    return SB200_ERR_USER;
  }
  if (*k) sb200_ksp_destroy(*k);
  *k = nullptr;
  CHK(sb200_ksp_create_slab(n, want, P->rank, P->nranks, k));
  *have = want;
  return 0;
}

int ensure_all(sb200_saddle* P, bool may_create) {
  CHK(ensure_ksp(P, &P->kvel, &P->r_vel, P->gv, P->vel_maxits, may_create));
  CHK(ensure_ksp(P, &P->kschur, &P->r_schur, P->gp, P->schur_maxits, may_create));
  if (!P->svel_preonly) CHK(ensure_ksp(P, &P->ksvel, &P->r_svel, P->gv, P->svel_maxits, may_create));
  return 0;
}

// VecAXPY with the mean / MatNullSpaceRemove of the constant vector (stokes.C:1006-1025); on a slab partition the mean is global:
// local {sum, count} pairs are added over the ranks through the Schur solver's peer-memory all-reduce (rank order, same bits everywhere)
int remove_mean(sb200_saddle* P, long long n, int stride, int offset, double* x, void* stream) {
  if (P->nranks == 1) return sb200_vec_remove_mean(n, stride, offset, x, P->scratch, stream);
  CHK(sb200_vec_sum_count(n, stride, offset, x, P->scratch, P->sums, stream));
  CHK(sb200_ksp_allreduce_sum(P->kschur, P->sums, 2, stream));
  return sb200_vec_shift_mean(n, stride, offset, x, P->sums, stream);
}

// left-preconditioned GMRES: x = KSPSolve(b), Minv(b) staged in pb
int left_gmres(sb200_ksp* k, sb200_apply_fn op, sb200_saddle* P, const double* pb, double* x, double rtol, int maxits, long long* its_acc,
               void* stream) {
  CHK(sb200_ksp_set_operators(k, op, P, nullptr, nullptr));
  CHK(sb200_ksp_set_tolerances(k, rtol, 1e-50, 1e5, maxits));
  CHK(sb200_ksp_solve(k, pb, x, 0, stream));
  if (its_acc) {
    int its = 0;
    CHK(sb200_ksp_get_result(k, &its, nullptr, nullptr, nullptr));
    *its_acc += its;
  }
  return 0;
}

int solve_velocity(sb200_saddle* P, const double* rhs, double* x, void* stream) {
  CHK(pc_or_copy(P->vel_pc, P->vel_ctx, P->gv, rhs, P->v[PBV], stream));
  return left_gmres(P->kvel, op_velocity, P, P->v[PBV], x, P->vel_rtol, P->vel_maxits, &P->its_vel, stream);
}

int solve_schur_velocity(sb200_saddle* P, const double* rhs, double* x, void* stream) {
  if (P->svel_preonly) return pc_or_copy(P->svel_pc, P->svel_ctx, P->gv, rhs, x, stream);  // -svel_ksp_type preonly: one PC application
  CHK(pc_or_copy(P->svel_pc, P->svel_ctx, P->gv, rhs, P->v[UV], stream));
  return left_gmres(P->ksvel, op_schur_velocity, P, P->v[UV], x, P->svel_rtol, P->svel_maxits, nullptr, stream);
}

// StokesMatMultSchur (stokes.C:523-535): y = -PV * KSPSolve(KSPSchurVelocity, VP * x)
int schur_mult(sb200_saddle* P, const double* p, double* y, void* stream) {
  CHK(sb200_stokes_matmult_vp(P->s, p, P->v[SV0], stream));
  CHK(solve_schur_velocity(P, P->v[SV0], P->v[TV], stream));
  CHK(sb200_stokes_matmult_pv(P->s, P->v[TV], y, stream));
  return sb200_vec_axpby(P->gp, 0.0, nullptr, -1.0, y, stream);
}

int jacobi_project(sb200_saddle* P, const double* r, double* z, void* stream) {  // PCJacobi with 1/eta, then the constant null space
  CHK(sb200_vec_pointwise_divide(P->gp, r, P->diag, z, stream));
  return remove_mean(P, P->gp, 1, 0, z, stream);
}

int op_schur(void* ctx, const double* x, double* y, void* stream) {  // KSPSchur: A = the Schur shell, M = Jacobi
  sb200_saddle* P = (sb200_saddle*)ctx;
  CHK(schur_mult(P, x, P->p[WP], stream));
  return jacobi_project(P, P->p[WP], y, stream);
}

int solve_schur(sb200_saddle* P, const double* rhs, double* x, void* stream) {
  CHK(jacobi_project(P, rhs, P->p[SPB], stream));
  return left_gmres(P->kschur, op_schur, P, P->p[SPB], x, P->schur_rtol, P->schur_maxits, &P->its_schur, stream);
}

}  // namespace

extern "C" {

int sb200_saddle_create(sb200_stokes* s, int type, sb200_saddle** out) {
  if (!s || !out) {
    sb200::set_last_error("sb200_saddle_create: null pointer");
    return SB200_ERR_ARG;
  }
  *out = nullptr;
  if (type < 0 || type > 3) {
    sb200::set_last_error("pc_saddle_type not implemented (stokes.C:184)");
    return SB200_ERR_USER;
  }
  sb200_saddle* P = new sb200_saddle();
  P->s = s;
  P->type = type;
  int rc = sb200_stokes_sizes(s, &P->m, &P->g, &P->gp, &P->gv, &P->dv);
  if (!rc) rc = sb200_stokes_slab_info(s, &P->rank, &P->nranks, nullptr, nullptr, nullptr);
  if (!rc && P->gp <= 0) rc = SB200_ERR_USER;
  if (!rc) P->d = (int)(P->gv / P->gp);
  for (int i = 0; i < 8 && !rc; i++) rc = sb200_malloc((void**)&P->v[i], (size_t)P->gv * sizeof(double) + 16);
  for (int i = 0; i < 5 && !rc; i++) rc = sb200_malloc((void**)&P->p[i], (size_t)P->gp * sizeof(double) + 16);
  if (!rc) rc = sb200_malloc((void**)&P->diag, (size_t)P->gp * sizeof(double) + 16);
  if (!rc) rc = sb200_malloc((void**)&P->scratch, SB200_REDUCE_SCRATCH_DOUBLES * sizeof(double));
  if (!rc) rc = sb200_malloc((void**)&P->sums, 64 * sizeof(double));
  if (rc) {
    sb200_saddle_destroy(P);
    return rc;
  }
  *out = P;
  return 0;
}

int sb200_saddle_set_velocity_pc(sb200_saddle* P, sb200_apply_fn vel_pc, void* vel_ctx, sb200_apply_fn svel_pc, void* svel_ctx, int svel_same) {
  if (!P) return SB200_ERR_ARG;
  P->vel_pc = vel_pc;
  P->vel_ctx = vel_ctx;
  P->svel_pc = svel_same ? vel_pc : svel_pc;
  P->svel_ctx = svel_same ? vel_ctx : svel_ctx;
  return 0;
}

int sb200_saddle_set_inner(sb200_saddle* P, double vel_rtol, int vel_maxits, double schur_rtol, int schur_maxits, int svel_preonly) {
  if (!P) return SB200_ERR_ARG;
  if (vel_rtol < 0 || schur_rtol < 0 || vel_maxits < 0 || schur_maxits < 0) {
    sb200::set_last_error("sb200_saddle_set_inner: negative tolerance");
    return SB200_ERR_USER;
  }
  P->vel_rtol = vel_rtol;
  P->vel_maxits = vel_maxits;
  P->schur_rtol = schur_rtol;
  P->schur_maxits = schur_maxits;
  P->svel_preonly = svel_preonly != 0;
  return 0;
}

int sb200_saddle_set_svel(sb200_saddle* P, double svel_rtol, int svel_maxits) {
  if (!P) return SB200_ERR_ARG;
  if (svel_rtol < 0 || svel_maxits < 0) {
    sb200::set_last_error("sb200_saddle_set_svel: negative tolerance");
    return SB200_ERR_USER;
  }
  P->svel_rtol = svel_rtol;
  P->svel_maxits = svel_maxits;
  return 0;
}

int sb200_saddle_apply(sb200_saddle* P, const double* d_x, double* d_y, void* stream) {
  if (!P || !d_x || !d_y || d_x == d_y) {
    sb200::set_last_error("StokesPCApply: x and y must be distinct non-null vectors");
    return SB200_ERR_ARG;
  }
  CHK(ensure_all(P, P->nranks == 1));
  double *xv = P->v[XV], *xp = P->p[XP], *v1 = P->v[V1], *p1 = P->p[P1], *tp = P->p[TP];
  CHK(sb200_vec_split(P->gp, P->d, d_x, xv, xp, stream));              // scatterGV / scatterGP
  CHK(sb200_stokes_get_diagonal_schur(P->s, P->diag, stream));         // PCJacobi's MatGetDiagonal (1 / eta of the current state)
  switch (P->type) {
    case 0:  // block LU (stokes.C:1714-1742)
      CHK(solve_velocity(P, xv, v1, stream));                          // v1 <- A^-1 v0
      CHK(sb200_stokes_matmult_pv(P->s, v1, tp, stream));              // p0 <- B v1
      CHK(sb200_vec_axpby(P->gp, 1.0, xp, -1.0, tp, stream));          // p0 <- x_p - p0
      CHK(solve_schur(P, tp, p1, stream));                             // p1 <- S^-1 p0
      CHK(sb200_stokes_matmult_vp(P->s, p1, xv, stream));              // v0 <- B^T p1 (x_v is no longer needed)
      CHK(sb200_vec_axpby(P->gv, 0.0, nullptr, -1.0, xv, stream));     // v0 <- -v0
      CHK(solve_velocity(P, xv, P->v[TV], stream));                    // correction
      CHK(sb200_vec_axpby(P->gv, 1.0, P->v[TV], 1.0, v1, stream));     // ADD_VALUES into the velocity part
      break;
    case 1:  // block upper triangular (stokes.C:1747-1768)
      CHK(solve_schur(P, xp, p1, stream));
      CHK(sb200_stokes_matmult_vp(P->s, p1, P->v[TV], stream));
      CHK(sb200_vec_axpby(P->gv, 1.0, xv, -1.0, P->v[TV], stream));    // v0 <- x_v - B^T p1
      CHK(solve_velocity(P, P->v[TV], v1, stream));
      break;
    case 2:  // block diagonal (stokes.C:1773-1792)
      CHK(solve_velocity(P, xv, v1, stream));
      CHK(solve_schur(P, xp, p1, stream));
      break;
    default:  // block lower triangular (stokes.C:1797-1817)
      CHK(solve_velocity(P, xv, v1, stream));
      CHK(sb200_stokes_matmult_pv(P->s, v1, tp, stream));
      CHK(sb200_vec_axpby(P->gp, 1.0, xp, -1.0, tp, stream));
      CHK(solve_schur(P, tp, p1, stream));
      break;
  }
  return sb200_vec_merge(P->gp, P->d, v1, p1, d_y, stream);            // scatterVG / scatterPG
}

int sb200_saddle_remove_constant_pressure(sb200_saddle* P, double* d_x, void* stream) {
  if (!P || !d_x) return SB200_ERR_ARG;
  if (P->nranks > 1) CHK(ensure_all(P, false));
  return remove_mean(P, P->gp, P->d + 1, P->d, d_x, stream);
}

// ---- slab partition: the inner solvers' peer-mapped arenas -------------------------------------------------------------------
int sb200_saddle_prepare(sb200_saddle* P) {
  if (!P) return SB200_ERR_ARG;
  return ensure_all(P, true);
}

int sb200_saddle_ipc_export(sb200_saddle* P, void* handle192) {
  if (!P || !handle192) return SB200_ERR_ARG;
  CHK(ensure_all(P, P->nranks == 1));
  char* h = (char*)handle192;
  for (int i = 0; i < SB200_SADDLE_HANDLE_BYTES; i++) h[i] = 0;
  CHK(sb200_ksp_ipc_export(P->kvel, h));
  CHK(sb200_ksp_ipc_export(P->kschur, h + 64));
  if (P->ksvel) CHK(sb200_ksp_ipc_export(P->ksvel, h + 128));
  return 0;
}

int sb200_saddle_ipc_attach(sb200_saddle* P, int peer_rank, const void* handle192) {
  if (!P || !handle192) return SB200_ERR_ARG;
  CHK(ensure_all(P, P->nranks == 1));
  const char* h = (const char*)handle192;
  CHK(sb200_ksp_ipc_attach(P->kvel, peer_rank, h));
  CHK(sb200_ksp_ipc_attach(P->kschur, peer_rank, h + 64));
  if (P->ksvel) CHK(sb200_ksp_ipc_attach(P->ksvel, peer_rank, h + 128));
  return 0;
}

int sb200_saddle_attach_local(sb200_saddle* P, int peer_rank, sb200_saddle* peer) {
  if (!P || !peer) return SB200_ERR_ARG;
  CHK(ensure_all(P, false));
  CHK(ensure_all(peer, false));
  CHK(sb200_ksp_attach_local(P->kvel, peer_rank, peer->kvel));
  CHK(sb200_ksp_attach_local(P->kschur, peer_rank, peer->kschur));
  if (P->ksvel && peer->ksvel) CHK(sb200_ksp_attach_local(P->ksvel, peer_rank, peer->ksvel));
  return 0;
}

int sb200_apply_saddle(void* ctx, const double* d_x, double* d_y, void* stream) {
  sb200_saddle* P = (sb200_saddle*)ctx;
  CHK(sb200_saddle_apply(P, d_x, d_y, stream));
  return sb200_saddle_remove_constant_pressure(P, d_y, stream);
}

int sb200_saddle_get_inner_its(const sb200_saddle* P, long long* velocity, long long* schur) {
  if (!P) return SB200_ERR_ARG;
  if (velocity) *velocity = P->its_vel;
  if (schur) *schur = P->its_schur;
  return 0;
}

int sb200_saddle_destroy(sb200_saddle* P) {
  if (!P) return 0;
  for (double* a : P->v)
    if (a) sb200_free(a);
  for (double* a : P->p)
    if (a) sb200_free(a);
  if (P->diag) sb200_free(P->diag);
  if (P->scratch) sb200_free(P->scratch);
  if (P->sums) sb200_free(P->sums);
  if (P->kvel) sb200_ksp_destroy(P->kvel);
  if (P->kschur) sb200_ksp_destroy(P->kschur);
  if (P->ksvel) sb200_ksp_destroy(P->ksvel);
  delete P;
  return 0;
}

}  // extern "C"
