// Internal: device-resident flexible GMRES(m) - the Krylov solver the reference selects in code
// (KSPSetType(ksp, KSPFGMRES): elliptic.C:181-182, stokes.C:155-157) and the direct caller of the MatShells.
// PETSc's own implementation (src/ksp/ksp/impls/gmres/fgmres/fgmres.c, unpinned ~3.0 in the reference) is a
// third-party dependency; what is restated here is its published algorithm: right-preconditioned flexible
// Arnoldi with classical Gram-Schmidt (PETSc's default, no refinement), Givens-rotated Hessenberg, residual
// norm from the recurrence, KSPConvergedDefault (rnorm <= max(rtol*||b||, atol); divergence at dtol*||b||).
#pragma once
#include <cuda_runtime.h>

#include <vector>

#include "symm.h"

extern "C" typedef int (*sb200_apply_fn)(void* ctx, const double* d_x, double* d_y, void* stream);

namespace sb200 {

struct KspCtx {
  long long n = 0;     // local vector length
  long long ldv = 0;   // leading dimension of the Krylov bases (n rounded up to even)
  int restart = 30;    // KSPGMRESSetRestart default
  double rtol = 1e-5, atol = 1e-50, dtol = 1e5;  // KSP defaults
  int maxits = 10000;
  sb200_apply_fn op = nullptr, pc = nullptr;
  void* op_ctx = nullptr;
  void* pc_ctx = nullptr;

  // device state
  double* V = nullptr;   // (restart+1) basis vectors
  double* Z = nullptr;   // restart preconditioned vectors (aliases V when there is no PC)
  double* w = nullptr;   // work vector
  double* small = nullptr;  // Hessenberg (column-major, ld = restart+1), rotations, g, y, h scratch
  double* partial = nullptr;  // per-block partial sums of the reductions
  unsigned* counters = nullptr;
  double* h_rnorm = nullptr;  // pinned + mapped: the kernels report the recurrence residual norm here
  double* d_rnorm = nullptr;
  int nblocks = 0;

  // slab partition of the vectors: dot products are summed over the ranks through peer memory
  SymmArena arena;
  double* slots = nullptr;  // [2][nranks][64] exchange slots
  unsigned long long ar_epoch = 0;

  // results
  int its = 0, reason = 0;
  double rnorm = 0.0, bnorm = 0.0;
  std::vector<double> history;
  double t_op = 0.0, t_pc = 0.0, t_orth = 0.0;  // milliseconds (CUDA events), accumulated over the last solve
  static constexpr int NRING = 4;  // per-iteration event sets / residual-norm slots in flight (lookahead + 2 at most)
  cudaEvent_t evr[NRING][4] = {};
  int lookahead = 0;  // 1: iteration k+1 is enqueued before iteration k's norm is read (sb200_ksp_set_lookahead)

  static int create(long long n, int restart, int rank, int nranks, KspCtx** out);
  ~KspCtx();
  int solve(const double* b, double* x, bool guess_nonzero, cudaStream_t s);
  int allreduce(double* vals, int k, cudaStream_t s);  // sum over the slab ranks in rank order, in place (no-op on one rank)

 private:
  int init(long long n, int restart, int rank, int nranks);
  int dots(const double* x, const double* Y, long long ldy, int nv, double* out, cudaStream_t s);
  int norm(const double* x, double* out, cudaStream_t s);
};

}  // namespace sb200
