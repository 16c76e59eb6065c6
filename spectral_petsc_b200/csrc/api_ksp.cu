// C ABI of the device-resident FGMRES (include/spectral_b200.h, "KSP" section).
#include <cstring>

#include "../../include/spectral_b200.h"
#include "common.cuh"
#include "ksp.h"

using namespace sb200;

struct sb200_ksp {
  KspCtx* c = nullptr;
};

extern "C" {

int sb200_ksp_create(long long n, int restart, sb200_ksp** out) { return sb200_ksp_create_slab(n, restart, 0, 1, out); }

int sb200_ksp_create_slab(long long n_local, int restart, int rank, int nranks, sb200_ksp** out) {
  SB_CHECK(out, SB200_ERR_ARG, "null pointer");
  *out = nullptr;
  KspCtx* c = nullptr;
  SB_TRY(KspCtx::create(n_local, restart, rank, nranks, &c));
  sb200_ksp* k = new sb200_ksp();
  k->c = c;
  *out = k;
  return 0;
}

int sb200_ksp_set_operators(sb200_ksp* k, sb200_apply_fn op, void* op_ctx, sb200_apply_fn pc, void* pc_ctx) {
  SB_CHECK(k && op, SB200_ERR_ARG, "null pointer");
  k->c->op = op;
  k->c->op_ctx = op_ctx;
  k->c->pc = pc;
  k->c->pc_ctx = pc_ctx;
  return 0;
}

int sb200_ksp_set_tolerances(sb200_ksp* k, double rtol, double atol, double dtol, int maxits) {
  SB_CHECK(k, SB200_ERR_ARG, "null context");
  SB_CHECK(rtol >= 0 && atol >= 0 && dtol > 0 && maxits >= 0, SB200_ERR_USER, "KSPSetTolerances: negative tolerance");
  k->c->rtol = rtol;
  k->c->atol = atol;
  k->c->dtol = dtol;
  k->c->maxits = maxits;
  return 0;
}

int sb200_ksp_set_lookahead(sb200_ksp* k, int depth) {
  SB_CHECK(k, SB200_ERR_ARG, "null context");
  SB_CHECK(depth == 0 || depth == 1, SB200_ERR_USER, "KSP lookahead: 0 (read each iteration's norm before the next is enqueued) or 1");
  k->c->lookahead = depth;
  return 0;
}

int sb200_ksp_solve(sb200_ksp* k, const double* d_b, double* d_x, int guess_nonzero, void* stream) {
  SB_CHECK(k, SB200_ERR_ARG, "null context");
  SB_CHECK(!k->c->arena.failed(), SB200_ERR_CUDA, "slab partition: a device-side flag wait timed out earlier (a peer never arrived or the ranks' calls went out of step); the context refuses further work");
  return k->c->solve(d_b, d_x, guess_nonzero != 0, (cudaStream_t)stream);
}

int sb200_ksp_get_result(const sb200_ksp* k, int* its, double* rnorm, double* bnorm, int* reason) {
  SB_CHECK(k, SB200_ERR_ARG, "null context");
  if (its) *its = k->c->its;
  if (rnorm) *rnorm = k->c->rnorm;
  if (bnorm) *bnorm = k->c->bnorm;
  if (reason) *reason = k->c->reason;
  return 0;
}

int sb200_ksp_get_history(const sb200_ksp* k, double* h_hist, int cap, int* n) {
  SB_CHECK(k && n, SB200_ERR_ARG, "null pointer");
  const int have = (int)k->c->history.size();
  *n = have;
  if (h_hist && cap > 0) std::memcpy(h_hist, k->c->history.data(), sizeof(double) * (size_t)(have < cap ? have : cap));
  return 0;
}

int sb200_ksp_get_times(const sb200_ksp* k, double* ms_operator, double* ms_pc, double* ms_orthogonalisation) {
  SB_CHECK(k, SB200_ERR_ARG, "null context");
  if (ms_operator) *ms_operator = k->c->t_op;
  if (ms_pc) *ms_pc = k->c->t_pc;
  if (ms_orthogonalisation) *ms_orthogonalisation = k->c->t_orth;
  return 0;
}

int sb200_ksp_allreduce_sum(sb200_ksp* k, double* d_vals, int count, void* stream) {
  SB_CHECK(k && d_vals && count >= 1, SB200_ERR_ARG, "sb200_ksp_allreduce_sum: bad arguments");
  SB_CHECK(!k->c->arena.failed(), SB200_ERR_CUDA, "slab partition: a device-side flag wait timed out earlier; the context refuses further work");
  return k->c->allreduce(d_vals, count, (cudaStream_t)stream);
}

int sb200_ksp_ipc_export(sb200_ksp* k, void* handle) {
  SB_CHECK(k, SB200_ERR_ARG, "null context");
  return k->c->arena.export_handle(handle);
}

int sb200_ksp_ipc_attach(sb200_ksp* k, int peer_rank, const void* handle) {
  SB_CHECK(k, SB200_ERR_ARG, "null context");
  return k->c->arena.attach(peer_rank, handle);
}

int sb200_ksp_attach_local(sb200_ksp* k, int peer_rank, sb200_ksp* peer) {
  SB_CHECK(k && peer, SB200_ERR_ARG, "null context");
  SB_CHECK(peer->c->arena.rank == peer_rank && peer->c->arena.nranks == k->c->arena.nranks, SB200_ERR_USER,
           "attach_local: peer context has a different rank / partition");
  return k->c->arena.attach_ptr(peer_rank, peer->c->arena.base);
}

int sb200_ksp_destroy(sb200_ksp* k) {
  if (!k) return 0;
  delete k->c;
  delete k;
  return 0;
}

// MatShell operators as sb200_apply_fn, so a KSP can be pointed at a context without host glue
int sb200_apply_elliptic_matmult(void* ctx, const double* d_x, double* d_y, void* stream) {
  return sb200_elliptic_matmult((sb200_elliptic*)ctx, d_x, d_y, stream);
}
int sb200_apply_stokes_matmult(void* ctx, const double* d_x, double* d_y, void* stream) {
  return sb200_stokes_matmult((sb200_stokes*)ctx, d_x, d_y, stream);
}
int sb200_apply_stokes_matmult_vv(void* ctx, const double* d_x, double* d_y, void* stream) {
  return sb200_stokes_matmult_vv((sb200_stokes*)ctx, d_x, d_y, stream);
}

}  // extern "C"
