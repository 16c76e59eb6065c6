// Fused MatMult_Elliptic (elliptic.C:297-339) for d-dimensional grids whose extents are all P with
// P % 16 == 0 (P = 32, 64, 128): one persistent "chain" kernel per axis,
//     out  = 0   - D_0 (eta D_0 w + deta w g0_0)          axis 0   (FIRST: no read of out)
//     out  = out - D_k (eta D_k w + deta w g0_k)          middle axes
//     V    = crop(out - D_l (eta D_l w + deta w g0_l))    last axis (R == 1; crop fused)
// preserving the reference's accumulation order (elliptic.C:331-334).  Each warp runs the whole
// chain for 8 grid lines out of shared memory (see chain.cuh); the even-odd halves of D are resident
// in shared memory for the life of the persistent CTA.
#include "../../include/spectral_b200.h"
#include "chain.cuh"
#include "deriv.h"
#include "elliptic.h"

namespace sb200 {

namespace {

constexpr int NWARPS = 16;

struct ChainParams {
  const double* Ae;
  const double* Bo;
  const double* w;     // padded local field (m)
  const double* eta;   // m
  const double* deta;  // m
  const double* g0;    // gradu[axis] (m)
  double* out;         // m: accumulator field
  double* V;           // g: cropped result (LAST only)
  LineGeom lg;
  // crop geometry (LAST): lines are indexed by the leading d-1 indices
  int d;
  int dim[SB200_MAX_DIM];
};

enum { POS_FIRST = 0, POS_MID = 1, POS_LAST = 2 };

template <int P, bool RIGHT, int POS>
__global__ void __launch_bounds__(NWARPS * 32, 1) chain_kernel(ChainParams p) {
  using E = EO<P>;
  extern __shared__ double sm[];
  double* Ae = sm;
  double* Bo = sm + E::H * E::LDM;
  constexpr int BE = RIGHT ? E::BLOCK_ELEMS_RIGHT : E::BLOCK_ELEMS_LEFT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  double* Xw = sm + E::MAT_ELEMS + warp * BE;

  const long long nblocks = (p.lg.nlines + 7) / 8;
  const long long per = (nblocks + gridDim.x - 1) / gridDim.x;
  const long long b_begin = (long long)blockIdx.x * per;
  const long long b_end = b_begin + per < nblocks ? b_begin + per : nblocks;

  load_matrices<P>(sm, p.Ae, p.Bo);
  long long blk = b_begin + warp;
  if (blk < b_end) load_block<P, RIGHT>(Xw, p.w, p.lg, blk * 8, lane);
  cp_async_wait<0>();
  __syncthreads();  // matrices visible to all warps (the only CTA-wide barrier)

  for (; blk < b_end; blk += NWARPS) {
    const long long n0 = blk * 8;
    // per-thread global bases of its two C-fragment columns (lines n0+2t, n0+2t+1)
    long long base[2];
    bool lok[2];
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const long long n = n0 + 2 * t + h;
      lok[h] = n < p.lg.nlines;
      base[h] = lok[h] ? p.lg.base(n) : 0;
    }
    double a[E::MT][2], b[E::MT][2];
    eo_gemm<P, RIGHT>(Ae, Bo, Xw, a, b, g, t);
    __syncwarp();
    // flux (elliptic.C:319-323), written back in place: f = eta*y + (deta*w)*g0
#pragma unroll
    for (int i = 0; i < E::MT; i++) {
      const int r = i * 8 + g;
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int c = 2 * t + h;
        const double ytop = a[i][h] + b[i][h];
        const double ybot = b[i][h] - a[i][h];
        const int mt = r, mb = P - 1 - r;
        double ft = 0.0, fb = 0.0;
        if (lok[h]) {
          const long long et = base[h] + (long long)mt * p.lg.R;
          const long long eb = base[h] + (long long)mb * p.lg.R;
          const double wt = Xw[xaddr<P, RIGHT>(mt, c)], wb = Xw[xaddr<P, RIGHT>(mb, c)];
          ft = __dadd_rn(__dmul_rn(__ldg(p.eta + et), ytop), __dmul_rn(__dmul_rn(__ldg(p.deta + et), wt), __ldg(p.g0 + et)));
          fb = __dadd_rn(__dmul_rn(__ldg(p.eta + eb), ybot), __dmul_rn(__dmul_rn(__ldg(p.deta + eb), wb), __ldg(p.g0 + eb)));
        }
        Xw[xaddr<P, RIGHT>(mt, c)] = ft;
        Xw[xaddr<P, RIGHT>(mb, c)] = fb;
      }
    }
    __syncwarp();
    eo_gemm<P, RIGHT>(Ae, Bo, Xw, a, b, g, t);
    __syncwarp();  // all lanes done reading the block: safe to refill it
    const long long nblk = blk + NWARPS;
    if (nblk < b_end) load_block<P, RIGHT>(Xw, p.w, p.lg, nblk * 8, lane);  // overlaps the epilogue

    // epilogue: out = (FIRST ? 0 : out) - D f ; LAST crops into V
    long long vrow[2] = {0, 0};
    bool vint[2] = {false, false};
    if (POS == POS_LAST) {
#pragma unroll
      for (int h = 0; h < 2; h++) {
        // line index = lexicographic index over the leading d-1 axes; interior id of the line
        long long n = n0 + 2 * t + h, gid = 0, mul = 1;
        bool interior = lok[h];
        for (int j = p.d - 2; j >= 0; j--) {
          const int ij = (int)(n % p.dim[j]);
          n /= p.dim[j];
          interior = interior && ij > 0 && ij < p.dim[j] - 1;
          gid += (long long)(ij - 1) * mul;
          mul *= p.dim[j] - 2;
        }
        vint[h] = interior;
        vrow[h] = gid * (P - 2);
      }
    }
#pragma unroll
    for (int i = 0; i < E::MT; i++) {
      const int r = i * 8 + g;
#pragma unroll
      for (int h = 0; h < 2; h++) {
        if (!lok[h]) continue;
        const double ytop = a[i][h] + b[i][h];
        const double ybot = b[i][h] - a[i][h];
        const int mt = r, mb = P - 1 - r;
        const long long et = base[h] + (long long)mt * p.lg.R;
        const long long eb = base[h] + (long long)mb * p.lg.R;
        if (POS == POS_FIRST) {
          p.out[et] = 0.0 - ytop;
          p.out[eb] = 0.0 - ybot;
        } else if (POS == POS_MID) {
          p.out[et] = p.out[et] - ytop;
          p.out[eb] = p.out[eb] - ybot;
        } else {
          if (vint[h]) {
            if (mt > 0) p.V[vrow[h] + mt - 1] = p.out[et] - ytop;  // mt < P/2 so never the far end
            if (mb < P - 1) p.V[vrow[h] + mb - 1] = p.out[eb] - ybot;
          }
        }
      }
    }
    if (nblk < b_end) cp_async_wait<0>();
    __syncwarp();
  }
}

template <int P, bool RIGHT, int POS>
int launch_chain(const ChainParams& p, cudaStream_t s) {
  using E = EO<P>;
  constexpr int BE = RIGHT ? E::BLOCK_ELEMS_RIGHT : E::BLOCK_ELEMS_LEFT;
  const size_t smem = (size_t)(E::MAT_ELEMS + NWARPS * BE) * sizeof(double);
  auto kern = chain_kernel<P, RIGHT, POS>;
  static bool attr = false;
  if (!attr) {
    SB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long nblocks = (p.lg.nlines + 7) / 8;
  // one persistent CTA per SM; fewer when there is not a block per warp to hand out
  long long grid = (nblocks + NWARPS - 1) / NWARPS;
  if (grid > sms) grid = sms;
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, NWARPS * 32, smem, s>>>(p);
  count_launch();
  SB_CUDA(cudaGetLastError());
  return 0;
}

template <int P>
int matmult_fused_P(EllipticCtx& e, const double* U, double* V, cudaStream_t s) {
  const int d = e.gd.d;
  SB_TRY(e.pad(U, false, e.w[0], s));
  for (int k = 0; k < d; k++) {
    ChainParams p;
    p.Ae = e.Dax[k]->d_Ae;
    p.Bo = e.Dax[k]->d_Bo;
    p.w = e.w[0];
    p.eta = e.eta;
    p.deta = e.deta;
    p.g0 = e.gradu[k];
    p.out = e.w[1];
    p.V = V;
    p.lg.R = e.gd.stride[k];
    p.lg.PR = (long long)P * e.gd.stride[k];
    p.lg.nlines = e.gd.m / P;
    p.d = d;
    for (int j = 0; j < d; j++) p.dim[j] = e.gd.dim[j];
    const bool last = (k == d - 1);
    if (k == 0 && !last) SB_TRY((launch_chain<P, false, POS_FIRST>(p, s)));
    else if (!last) SB_TRY((launch_chain<P, false, POS_MID>(p, s)));
    else SB_TRY((launch_chain<P, true, POS_LAST>(p, s)));
  }
  return 0;
}

}  // namespace

bool elliptic_fused_supported(const EllipticCtx& e) {
  const int d = e.gd.d;
  if (d < 2) return false;
  const int P = e.gd.dim[0];
  for (int j = 1; j < d; j++)
    if (e.gd.dim[j] != P) return false;
  return P == 32 || P == 64 || P == 128;
}

int elliptic_matmult_fused(EllipticCtx& e, const double* U, double* V, cudaStream_t s) {
  switch (e.gd.dim[0]) {
    case 32: return matmult_fused_P<32>(e, U, V, s);
    case 64: return matmult_fused_P<64>(e, U, V, s);
    case 128: return matmult_fused_P<128>(e, U, V, s);
  }
  set_last_error("fused path: unsupported extent");
  return SB200_ERR_SUP;
}

}  // namespace sb200
