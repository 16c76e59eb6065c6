// Inner-loop feasibility microbench for the even-odd fused chain kernels: how many shared-memory
// fragment loads per DMMA can the SM sustain at full FP64 tensor rate?
// Variant<MT,NT>: per k4-step a warp loads 2*MT matrix fragments (Ae,Bo) + 2*NT field fragments (p,q -> s,d)
// and issues 2*MT*NT DMMAs.  8 warps per CTA, 1 CTA per SM (big smem), K=64 loop repeated.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

constexpr int LDM = 68;   // matrix leading dim (64 + 4)
constexpr int LDX = 68;   // field tile leading dim, 64 columns

template <int MT, int NT, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32) eo_kernel(double* out, int iters) {
  extern __shared__ double sm[];
  double* Ae = sm;                 // [64][LDM]
  double* Bo = Ae + 64 * LDM;      // [64][LDM]
  double* Xs = Bo + 64 * LDM;      // [128][LDX]
  for (int i = threadIdx.x; i < 64 * LDM * 2 + 128 * LDX; i += blockDim.x) sm[i] = 1e-3 * (i % 97);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int m0 = (warp * MT * 8) % 64;
  const int n0 = (warp * NT * 8) % 64;
  double a[MT][NT][2], b[MT][NT][2];
#pragma unroll
  for (int i = 0; i < MT; i++)
#pragma unroll
    for (int j = 0; j < NT; j++) { a[i][j][0] = a[i][j][1] = 0; b[i][j][0] = b[i][j][1] = 0; }
  for (int it = 0; it < iters; it++) {
#pragma unroll 4
    for (int ks = 0; ks < 16; ks++) {
      double fa[MT], fb[MT], s[NT], d[NT];
#pragma unroll
      for (int j = 0; j < NT; j++) {
        double p = Xs[(ks * 4 + t) * LDX + n0 + j * 8 + g];
        double q = Xs[(127 - ks * 4 - t) * LDX + n0 + j * 8 + g];
        s[j] = p + q; d[j] = p - q;
      }
#pragma unroll
      for (int i = 0; i < MT; i++) {
        fa[i] = Ae[(m0 + i * 8 + g) * LDM + ks * 4 + t];
        fb[i] = Bo[(m0 + i * 8 + g) * LDM + ks * 4 + t];
      }
#pragma unroll
      for (int i = 0; i < MT; i++)
#pragma unroll
        for (int j = 0; j < NT; j++) { dmma884(a[i][j][0], a[i][j][1], fa[i], s[j]); dmma884(b[i][j][0], b[i][j][1], fb[i], d[j]); }
    }
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < MT; i++)
#pragma unroll
    for (int j = 0; j < NT; j++) r += a[i][j][0] + a[i][j][1] + b[i][j][0] + b[i][j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

// explicitly double-buffered fragments: loads for k-step ks+1 are issued before the DMMAs of ks
template <int MT, int NT, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32) eo_kernel_pipe(double* out, int iters) {
  extern __shared__ double sm[];
  double* Ae = sm;
  double* Bo = Ae + 64 * LDM;
  double* Xs = Bo + 64 * LDM;
  for (int i = threadIdx.x; i < 64 * LDM * 2 + 128 * LDX; i += blockDim.x) sm[i] = 1e-3 * (i % 97);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int m0 = (warp * MT * 8) % 64;
  const int n0 = (warp * NT * 8) % 64;
  double a[MT][NT][2], b[MT][NT][2];
#pragma unroll
  for (int i = 0; i < MT; i++)
#pragma unroll
    for (int j = 0; j < NT; j++) { a[i][j][0] = a[i][j][1] = 0; b[i][j][0] = b[i][j][1] = 0; }
  double fa[2][MT], fb[2][MT], p[2][NT], q[2][NT];
  auto ld = [&](int buf, int ks) {
#pragma unroll
    for (int j = 0; j < NT; j++) {
      p[buf][j] = Xs[(ks * 4 + t) * LDX + n0 + j * 8 + g];
      q[buf][j] = Xs[(127 - ks * 4 - t) * LDX + n0 + j * 8 + g];
    }
#pragma unroll
    for (int i = 0; i < MT; i++) {
      fa[buf][i] = Ae[(m0 + i * 8 + g) * LDM + ks * 4 + t];
      fb[buf][i] = Bo[(m0 + i * 8 + g) * LDM + ks * 4 + t];
    }
  };
  for (int it = 0; it < iters; it++) {
    ld(0, 0);
#pragma unroll
    for (int ks = 0; ks < 16; ks++) {
      const int cur = ks & 1;
      if (ks + 1 < 16) ld(cur ^ 1, ks + 1);
      double s[NT], d[NT];
#pragma unroll
      for (int j = 0; j < NT; j++) { s[j] = p[cur][j] + q[cur][j]; d[j] = p[cur][j] - q[cur][j]; }
#pragma unroll
      for (int i = 0; i < MT; i++)
#pragma unroll
        for (int j = 0; j < NT; j++) { dmma884(a[i][j][0], a[i][j][1], fa[cur][i], s[j]); dmma884(b[i][j][0], b[i][j][1], fb[cur][i], d[j]); }
    }
  }
  double r = 0;
#pragma unroll
  for (int i = 0; i < MT; i++)
#pragma unroll
    for (int j = 0; j < NT; j++) r += a[i][j][0] + a[i][j][1] + b[i][j][0] + b[i][j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MT, int NT, int NWARPS, bool PIPE = false>
void run(const char* name, double* out, int sms) {
  size_t smem = (64 * LDM * 2 + 128 * LDX) * sizeof(double);
  auto k = PIPE ? eo_kernel_pipe<MT, NT, NWARPS> : eo_kernel<MT, NT, NWARPS>;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int iters = 400;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  k<<<sms, NWARPS * 32, smem>>>(out, iters); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < 5; r++) {
    CK(cudaEventRecord(e0)); k<<<sms, NWARPS * 32, smem>>>(out, iters); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  double fl = 2.0 * 256 * (2.0 * MT * NT) * 16 * iters * NWARPS * sms;
  printf("{\"bench\": \"%s\", \"MT\": %d, \"NT\": %d, \"warps\": %d, \"lds_per_dmma\": %.3f, \"ms\": %.3f, \"tflops\": %.2f}\n", name, MT, NT, NWARPS,
         (2.0 * MT + 2.0 * NT) / (2.0 * MT * NT), best, fl / best * 1e-9);
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 512));
  run<8, 1, 4>("eo_inner", out, sms);
  run<8, 1, 8>("eo_inner", out, sms);
  run<8, 1, 12>("eo_inner", out, sms);
  run<8, 1, 4, true>("eo_pipe", out, sms);
  run<8, 1, 8, true>("eo_pipe", out, sms);
  run<8, 1, 12, true>("eo_pipe", out, sms);
  run<8, 2, 4>("eo_inner", out, sms);
  run<8, 2, 4, true>("eo_pipe", out, sms);
  run<8, 2, 8, true>("eo_pipe", out, sms);
  run<4, 4, 4, true>("eo_pipe", out, sms);
  return 0;
}
