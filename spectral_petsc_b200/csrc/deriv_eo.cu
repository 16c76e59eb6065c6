// Even-odd persistent derivative kernel: y = D_axis x (with the reference's AXPY accumulation fused in the
// epilogue) for P in {16, 32, 64, 128}, any (O, P, R) factorisation and any element stride / offset (the AoS
// velocity components of stokes.C:284-289,585,613).
//
// Same machinery as the fused chain kernels (chain.cuh): Ae / Bo (the two half-size matrices of the
// centro-antisymmetric CGL matrix) stay resident in shared memory, every warp owns an 8-line block, pulls
// tickets from a global counter and runs  load -> even-odd DMMA GEMM -> epilogue  with no CTA barrier after the
// matrix load.  Executed flops are half of the dense product that deriv_generic.cu performs.
#include "../../include/spectral_b200.h"
#include "chain.cuh"
#include "deriv.h"

namespace sb200 {

namespace {

// Up to three derivatives that share the matrix run as ONE launch: the tickets of job j are [start[j], end[j]).
// (The reference applies D_0, D_1, D_2 back to back - stokes.C:584-590,611-614,639,668-671 - each on the full
// grid; one launch keeps every SM busy through the tail of each and loads Ae / Bo once.)
struct EoJob {
  const double* x;
  double* y;
  const double* yin;
  long long R, nlines;
  int xs, xoff, ys, yoff, mode;
  int vec;  // R % 8 == 0, unit strides, 16-byte aligned: the block's 8 lines are adjacent in memory
  unsigned end;  // exclusive prefix of item counts
};
struct EoParams {
  EoJob job[SB200_EO_MAX_JOBS];
  int njobs;
  const double* Ae;
  const double* Bo;
  unsigned items;     // total
  unsigned* sync;     // [0] ticket, [1] exited warps
};

template <int P, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32, 1) eo_deriv_kernel(EoParams q) {
  using E = EO<P>;
  extern __shared__ double sm[];
  double* Ae = sm;
  double* Bo = sm + E::H * E::LDM;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  double* Xw = sm + E::MAT_ELEMS + warp * E::BLOCK_ELEMS_LEFT;
  load_matrices<P>(sm, q.Ae, q.Bo);

  auto grab = [&]() -> unsigned {
    unsigned tk = 0;
    if (lane == 0) tk = atomicAdd(q.sync, 1u);
    return __shfl_sync(0xffffffffu, tk, 0);
  };
  auto job_of = [&](unsigned tk) -> int {
    int j = 0;
    while (j + 1 < q.njobs && tk >= q.job[j].end) j++;
    return j;
  };
  auto line_base = [&](long long n, long long R) -> long long {  // element index of (line n, m = 0)
    const long long o = n / R, r = n - o * R;
    return o * (long long)P * R + r;
  };
  auto issue_load = [&](unsigned tk) {
    if (tk >= q.items) return;
    const int j = job_of(tk);
    const EoJob& p = q.job[j];
    const long long R = p.R;
    const long long n0 = (long long)(tk - (j ? q.job[j - 1].end : 0u)) * 8;
    if (p.vec) {
      const long long b0 = line_base(n0, R);
#pragma unroll 4
      for (int idx = lane; idx < P * 4; idx += 32) {
        const int m = idx >> 2, c2 = (idx & 3) * 2;
        cp_async16(Xw + xaddr<P, false>(m, c2), p.x + b0 + (long long)m * R + c2, true);
      }
    } else {
      const int c = lane & 7;
      const bool ok = (n0 + c) < p.nlines;
      const long long bc = ok ? line_base(n0 + c, R) : 0;
#pragma unroll 4
      for (int m = lane >> 3; m < P; m += 4)
        cp_async8(Xw + xaddr<P, false>(m, c), p.x + (ok ? (bc + (long long)m * R) * p.xs + p.xoff : 0), ok);
    }
    cp_async_commit();
  };

  unsigned tk = grab();
  issue_load(tk);
  cp_async_wait<0>();
  __syncthreads();  // matrices visible to all warps

  while (tk < q.items) {
    const int j = job_of(tk);
    const EoJob& p = q.job[j];
    const long long R = p.R;
    const long long n0 = (long long)(tk - (j ? q.job[j - 1].end : 0u)) * 8;
    cp_async_wait<0>();
    __syncwarp();
    double a[E::MT][2], b[E::MT][2];
    eo_gemm<P, false>(Ae, Bo, Xw, a, b, g, t);
    __syncwarp();  // block free: refill it while the epilogue drains
    const unsigned nxt = grab();
    issue_load(nxt);

    // thread-owned outputs: lines 2t, 2t+1; rows mt = i*8+g (a+b) and mb = P-1-mt (b-a)
    if (p.vec) {
      const long long base = line_base(n0, R) + 2 * t;
#pragma unroll
      for (int i = 0; i < E::MT; i++) {
        const long long et = base + (long long)(i * 8 + g) * R, eb = base + (long long)(P - 1 - i * 8 - g) * R;
        double2 vt = make_double2(a[i][0] + b[i][0], a[i][1] + b[i][1]);
        double2 vb = make_double2(b[i][0] - a[i][0], b[i][1] - a[i][1]);
        if (p.mode != DERIV_STORE) {
          const double2 yt = p.yin ? ld2(p.yin + et) : make_double2(0.0, 0.0);
          const double2 yb = p.yin ? ld2(p.yin + eb) : make_double2(0.0, 0.0);
          if (p.mode == DERIV_SUB) {
            vt = make_double2(yt.x - vt.x, yt.y - vt.y);
            vb = make_double2(yb.x - vb.x, yb.y - vb.y);
          } else {
            vt = make_double2(yt.x + vt.x, yt.y + vt.y);
            vb = make_double2(yb.x + vb.x, yb.y + vb.y);
          }
        }
        st2(p.y + et, vt.x, vt.y);
        st2(p.y + eb, vb.x, vb.y);
      }
    } else {
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const long long n = n0 + 2 * t + h;
        if (n >= p.nlines) continue;
        const long long lb = line_base(n, R);
#pragma unroll
        for (int i = 0; i < E::MT; i++) {
          const long long et = (lb + (long long)(i * 8 + g) * R) * p.ys + p.yoff;
          const long long eb = (lb + (long long)(P - 1 - i * 8 - g) * R) * p.ys + p.yoff;
          double vt = a[i][h] + b[i][h], vb = b[i][h] - a[i][h];
          if (p.mode == DERIV_SUB) {
            vt = (p.yin ? p.yin[et] : 0.0) - vt;
            vb = (p.yin ? p.yin[eb] : 0.0) - vb;
          } else if (p.mode == DERIV_ADD) {
            vt = (p.yin ? p.yin[et] : 0.0) + vt;
            vb = (p.yin ? p.yin[eb] : 0.0) + vb;
          }
          p.y[et] = vt;
          p.y[eb] = vb;
        }
      }
    }
    tk = nxt;
  }
  cp_async_wait<0>();
  if (lane == 0) {
    const unsigned gone = atomicAdd(q.sync + 1, 1u);
    if (gone == gridDim.x * NWARPS - 1) {  // the last warp to leave re-arms the counters
      q.sync[0] = 0;
      q.sync[1] = 0;
    }
  }
}

template <int P, int NWARPS>
int launch_eo(const EoParams& q, cudaStream_t s) {
  using E = EO<P>;
  auto kern = eo_deriv_kernel<P, NWARPS>;
  const size_t smem = (size_t)(E::MAT_ELEMS + NWARPS * E::BLOCK_ELEMS_LEFT) * sizeof(double);
  static bool attr = false;
  if (!attr) {
    SB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long grid = (q.items + NWARPS - 1) / NWARPS;
  if (grid > sms) grid = sms;
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, NWARPS * 32, smem, s>>>(q);
  count_launch();
  SB_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

bool deriv_eo_supported(const DerivParams& p) {
  if (!p.Ae || !p.Bo || p.npeer > 1) return false;
  return p.P == 16 || p.P == 32 || p.P == 64 || p.P == 128;
}

int deriv_eo_batch(const DerivParams* jobs, int n, unsigned* sync, cudaStream_t s) {
  SB_CHECK(n >= 1 && n <= SB200_EO_MAX_JOBS, SB200_ERR_USER, "even-odd derivative: bad job count");
  EoParams q;
  q.njobs = n;
  q.Ae = jobs[0].Ae;
  q.Bo = jobs[0].Bo;
  q.sync = sync;
  unsigned total = 0;
  for (int j = 0; j < n; j++) {
    const DerivParams& p = jobs[j];
    SB_CHECK(deriv_eo_supported(p) && p.P == jobs[0].P && p.Ae == q.Ae, SB200_ERR_USER, "even-odd derivative: jobs must share the matrix");
    SB_CHECK(p.x != p.y, SB200_ERR_ARG, "deriv: x and y must not alias (chebyshev.c:127)");
    EoJob& e = q.job[j];
    e.x = p.x;
    e.y = p.y;
    e.yin = p.yin;
    e.R = p.R;
    e.nlines = p.O * p.R;
    e.xs = p.xs;
    e.xoff = p.xoff;
    e.ys = p.ys;
    e.yoff = p.yoff;
    e.mode = p.mode;
    e.vec = (p.R % 8 == 0) && p.xs == 1 && p.ys == 1 && p.xoff == 0 && p.yoff == 0 &&
            ((reinterpret_cast<uintptr_t>(p.x) | reinterpret_cast<uintptr_t>(p.y) | reinterpret_cast<uintptr_t>(p.yin)) % 16 == 0);
    total += (unsigned)((e.nlines + 7) / 8);
    e.end = total;
  }
  q.items = total;
  switch (jobs[0].P) {
    case 16: return launch_eo<16, 16>(q, s);
    case 32: return launch_eo<32, 16>(q, s);
    case 64: return launch_eo<64, 16>(q, s);
    case 128: return launch_eo<128, 16>(q, s);
  }
  set_last_error("even-odd derivative: unsupported extent");
  return SB200_ERR_SUP;
}

int deriv_eo_apply(const DerivParams& p, unsigned* sync, cudaStream_t s) { return deriv_eo_batch(&p, 1, sync, s); }

}  // namespace sb200
