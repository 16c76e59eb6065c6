// Internal interface of the generic per-axis derivative kernel (deriv_generic.cu).
#pragma once
#include <cuda_runtime.h>

namespace sb200 {

enum DerivMode { DERIV_STORE = 0, DERIV_SUB = 1, DERIV_ADD = 2 };

struct DerivParams {
  const double* D;    // device, Pp x Pp row-major, zero padded
  const double* Ae = nullptr;  // even-odd halves of D ([P/2][P/2], P even); with `sync` they enable the even-odd
  const double* Bo = nullptr;  // persistent kernel (deriv_eo.cu) for P in {16, 32, 64, 128}
  unsigned* sync = nullptr;    // 2 zero-initialised counters owned by the calling context (ticket, exited warps)
  int P, Pp;          // extent of the differentiated axis, padded extent (multiple of 32)
  const double* x;    // input field
  double* y;          // output field (must not alias x)
  const double* yin;  // accumulation input for SUB/ADD (may alias y, may be null = 0)
  long long O, R;     // array factored as (O, P, R), row-major
  int xs, xoff;       // element e of x lives at x[e*xs + xoff]  (AoS component access)
  int ys, yoff;       // same for y / yin
  int mode;           // DerivMode: y = acc | yin - acc | yin + acc
  // Slab-partitioned axis (multi-GPU, O == 1): input row k is read from xpeer[k / nloc] at local row
  // k % nloc (peer memory over NVLink, the "all-gather" is the operand load itself); this rank computes
  // output rows [row0, row0 + nloc) and stores them at local rows 0..nloc-1.  npeer <= 1: off.
  int npeer = 0, nloc = 0, row0 = 0;
  const double* xpeer[8] = {};
};

int deriv_apply(const DerivParams& p, cudaStream_t stream);
bool deriv_eo_supported(const DerivParams& p);
int deriv_eo_apply(const DerivParams& p, unsigned* sync, cudaStream_t stream);
#define SB200_EO_MAX_JOBS 3
// Up to SB200_EO_MAX_JOBS derivatives sharing the matrix (same extent) in ONE launch; outputs must be distinct.
int deriv_eo_batch(const DerivParams* jobs, int n, unsigned* sync, cudaStream_t stream);

void count_launch(int n = 1);

struct SymmArena;
// Transpose-based derivative along the partitioned axis of a slab-distributed field (slab_deriv.cu).
bool slab_deriv0_pencil_supported(const SymmArena& a, const DerivParams& p);
int slab_deriv0_pencil(SymmArena& a, const DerivParams& p, int nloc, int i0, double* Xp, double* Yp, cudaStream_t s);
int slab_deriv0_pencil_begin(SymmArena& a, const DerivParams& p, int nloc, int i0, double* Xp, cudaStream_t s);
int slab_deriv0_pencil_finish(SymmArena& a, const DerivParams& p, int nloc, double* Xp, double* Yp, cudaStream_t s);

}  // namespace sb200
