// FP64 pipe microbenchmark for B200 (sm_100a): the roofline denominator for the
// dense Chebyshev derivative.  Measures DFMA (vector) and DMMA (mma.sync f64)
// register-resident peak rates.  Build: see tools/Makefile.  Output: JSON lines.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void __launch_bounds__(256) dfma_kernel(double* out, int iters, double a, double b) {
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; i++) acc[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* c, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double* c, const double* a, const double* b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                 "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int NACC>
__global__ void __launch_bounds__(256) dmma884_kernel(double* out, int iters) {
  double c[NACC][2];
  double a = threadIdx.x * 1e-6, b = 1.0 + threadIdx.x * 1e-7;
#pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = i; c[i][1] = -i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) dmma1688_kernel(double* out, int iters) {
  double c[NACC][4];
  double a[4], b[2];
  for (int i = 0; i < 4; i++) a[i] = threadIdx.x * 1e-6 + i;
  for (int i = 0; i < 2; i++) b[i] = 1.0 + threadIdx.x * 1e-7 + i;
#pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) dmma1688(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) dmma16816_kernel(double* out, int iters) {
  double c[NACC][4];
  double a[8], b[4];
  for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-6 + i;
  for (int i = 0; i < 4; i++) b[i] = 1.0 + threadIdx.x * 1e-7 + i;
#pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) dmma16816(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DMMA fed from shared memory each k-step: warp tile 32x32 (4 A frags x 4 B frags = 16 DMMA per k4 step)
__global__ void __launch_bounds__(256) dmma_smem_kernel(double* out, int iters) {
  __shared__ double sA[8][32 * 16 + 8];
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = lane; i < 32 * 16; i += 32) { sA[warp][i] = i * 1e-6; }
  __syncwarp();
  double c[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) { c[i][j][0] = i; c[i][j][1] = j; }
  int g = lane >> 2, t = lane & 3;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int ks = 0; ks < 4; ks++) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; i++) a[i] = sA[warp][(i * 8 + g) * 16 + ks * 4 + t];
#pragma unroll
      for (int j = 0; j < 4; j++) b[j] = sA[warp][((j * 8 + g + 8) & 31) * 16 + ks * 4 + t];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) dmma884(c[i][j][0], c[i][j][1], a[i], b[j]);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) s += c[i][j][0] + c[i][j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
static float time_ms(F launch, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); launch();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(e0));
    launch();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  return best;
}

int main(int argc, char** argv) {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 16 * 256));
  int iters = 20000;
  printf("{\"device\": \"%s\", \"sms\": %d}\n", p.name, sms);
  for (int cps = 1; cps <= 8; cps *= 2) {
    int grid = sms * cps;
    float ms = time_ms([&] { dfma_kernel<<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
    double fl = 2.0 * 16 * iters * 256.0 * grid;
    printf("{\"bench\": \"dfma\", \"ctas_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", cps, ms, fl / ms * 1e-9);
  }
  for (int cps = 1; cps <= 8; cps *= 2) {
    int grid = sms * cps;
    float ms = time_ms([&] { dmma884_kernel<8><<<grid, 256>>>(out, iters); }, 5);
    double fl = 2.0 * 256 * 8 * iters * 8.0 * grid;
    printf("{\"bench\": \"dmma_m8n8k4_acc8\", \"ctas_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", cps, ms, fl / ms * 1e-9);
  }
  for (int cps = 1; cps <= 4; cps *= 2) {
    int grid = sms * cps;
    float ms = time_ms([&] { dmma884_kernel<16><<<grid, 256>>>(out, iters); }, 5);
    double fl = 2.0 * 256 * 16 * iters * 8.0 * grid;
    printf("{\"bench\": \"dmma_m8n8k4_acc16\", \"ctas_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", cps, ms, fl / ms * 1e-9);
  }
  for (int cps = 1; cps <= 4; cps *= 2) {
    int grid = sms * cps;
    float ms = time_ms([&] { dmma1688_kernel<8><<<grid, 256>>>(out, iters); }, 5);
    double fl = 2.0 * 16 * 8 * 8 * 8 * iters * 8.0 * grid;
    printf("{\"bench\": \"dmma_m16n8k8_acc8\", \"ctas_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", cps, ms, fl / ms * 1e-9);
  }
  for (int cps = 1; cps <= 4; cps *= 2) {
    int grid = sms * cps;
    float ms = time_ms([&] { dmma16816_kernel<8><<<grid, 256>>>(out, iters / 2); }, 5);
    double fl = 2.0 * 16 * 8 * 16 * 8 * (iters / 2) * 8.0 * grid;
    printf("{\"bench\": \"dmma_m16n8k16_acc8\", \"ctas_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", cps, ms, fl / ms * 1e-9);
  }
  for (int cps = 1; cps <= 4; cps *= 2) {
    int grid = sms * cps;
    float ms = time_ms([&] { dmma_smem_kernel<<<grid, 256>>>(out, iters / 8); }, 5);
    double fl = 2.0 * 256 * 16 * 4 * (iters / 8) * 8.0 * grid;
    printf("{\"bench\": \"dmma_m8n8k4_smem_32x32\", \"ctas_per_sm\": %d, \"ms\": %.3f, \"tflops\": %.2f}\n", cps, ms, fl / ms * 1e-9);
  }
  // long sustained DFMA run (power-capped clocks)
  {
    int grid = sms * 4;
    float ms = time_ms([&] { dfma_kernel<<<grid, 256>>>(out, iters * 20, 1.0000001, 1e-9); }, 3);
    double fl = 2.0 * 16 * iters * 20.0 * 256.0 * grid;
    printf("{\"bench\": \"dfma_sustained\", \"ctas_per_sm\": 4, \"ms\": %.3f, \"tflops\": %.2f}\n", ms, fl / ms * 1e-9);
    ms = time_ms([&] { dmma884_kernel<8><<<grid, 256>>>(out, iters * 20); }, 3);
    fl = 2.0 * 256 * 8 * iters * 20.0 * 8.0 * grid;
    printf("{\"bench\": \"dmma_m8n8k4_sustained\", \"ctas_per_sm\": 4, \"ms\": %.3f, \"tflops\": %.2f}\n", ms, fl / ms * 1e-9);
  }
  return 0;
}
