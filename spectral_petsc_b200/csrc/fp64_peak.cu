// FP64 tensor-pipe peak, measured in-run: the roofline denominator of bench.py ("FP64_peak measured on the box
// first").  Register-resident mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) chains, 8 independent accumulator pairs per warp,
// 8 warps per CTA, 4 CTAs per SM: the pipe is the only limiter.  2 * 8*8*4 = 512 flop per DMMA per warp.
#include "../../include/spectral_b200.h"
#include "common.cuh"
#include "deriv.h"

namespace sb200 {
namespace {

__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters) {
  double c[8][2];
  const double a = threadIdx.x * 1e-6, b = 1.0 + threadIdx.x * 1e-7;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    c[i][0] = i;
    c[i][1] = -i;
  }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace
}  // namespace sb200

extern "C" int sb200_fp64_dmma_peak(double target_ms, double* tflops, double* measured_ms) {
  using namespace sb200;
  SB_CHECK(tflops, SB200_ERR_ARG, "null pointer");
  int dev = 0, sms = 0;
  SB_CUDA(cudaGetDevice(&dev));
  SB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int blocks = sms * 4, threads = 256;
  double* out = nullptr;
  SB_CUDA(cudaMalloc((void**)&out, (size_t)blocks * threads * sizeof(double)));
  cudaEvent_t e0, e1;
  SB_CUDA(cudaEventCreate(&e0));
  SB_CUDA(cudaEventCreate(&e1));
  // calibrate: a short run, then scale the iteration count to the requested duration (clamped to 1..500 ms)
  if (!(target_ms > 1.0)) target_ms = 1.0;
  if (target_ms > 500.0) target_ms = 500.0;
  int iters = 2000;
  float ms = 0.f;
  for (int pass = 0; pass < 3; pass++) {
    SB_CUDA(cudaEventRecord(e0, 0));
    dmma_peak_kernel<<<blocks, threads>>>(out, iters);
    count_launch();
    SB_CUDA(cudaEventRecord(e1, 0));
    SB_CUDA(cudaEventSynchronize(e1));
    SB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (pass < 2) {
      const double scale = target_ms / (ms > 1e-3 ? ms : 1e-3);
      double it2 = iters * scale;
      if (it2 < 100) it2 = 100;
      if (it2 > 2e8) it2 = 2e8;
      iters = (int)it2;
    }
  }
  const double flop = 512.0 * 8.0 * iters * (double)blocks * (threads / 32);
  *tflops = flop / (ms * 1e-3) / 1e12;
  if (measured_ms) *measured_ms = ms;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  return 0;
}
