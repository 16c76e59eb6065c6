// Peer-memory bandwidth / latency of SM-issued loads and stores over NVLink between GPU 0 and GPU 1
// (one process, cudaDeviceEnablePeerAccess).  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/p2p_bw tools/p2p_bw.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void store16(double2* dst, long long n, double v) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = make_double2(v, v);
}
__global__ void copy16(double2* dst, const double2* src, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i];
}
// 64-byte rows scattered with a large stride (the chain epilogue pattern): each group of 4 threads writes one row
__global__ void store_rows64(double2* dst, long long nrows, long long row_stride16, double v) {
  const long long stride = (long long)gridDim.x * blockDim.x / 4;
  const int t = threadIdx.x & 3;
  for (long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / 4; r < nrows; r += stride) dst[r * row_stride16 + t] = make_double2(v, v);
}
__global__ void fence_latency(double2* dst, long long* out) {
  long long t0 = clock64();
  dst[threadIdx.x] = make_double2(1.0, 2.0);
  __threadfence_system();
  long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = t1 - t0;
}

template <class F>
float time_ms(F f, int reps) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  f();
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < reps; i++) f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms / reps;
}

int main() {
  int n = 0;
  CK(cudaGetDeviceCount(&n));
  if (n < 2) { printf("{\"error\": \"need 2 GPUs\"}\n"); return 0; }
  const long long bytes = 64ll << 20, n16 = bytes / 16;
  double2 *loc0, *loc0b, *rem1;
  CK(cudaSetDevice(1));
  CK(cudaMalloc(&rem1, bytes));
  CK(cudaMemset(rem1, 0, bytes));
  CK(cudaSetDevice(0));
  CK(cudaDeviceEnablePeerAccess(1, 0));
  CK(cudaMalloc(&loc0, bytes));
  CK(cudaMalloc(&loc0b, bytes));
  CK(cudaMemset(loc0, 0, bytes));
  long long* d_out;
  CK(cudaMalloc(&d_out, 64));
  for (int blocks : {148, 148 * 4, 148 * 8}) {
    for (long long sz : {4ll << 20, 64ll << 20}) {
      const long long m = sz / 16;
      float t_ls = time_ms([&] { store16<<<blocks, 256>>>(loc0, m, 1.0); }, 20);
      float t_rs = time_ms([&] { store16<<<blocks, 256>>>(rem1, m, 1.0); }, 20);
      float t_lc = time_ms([&] { copy16<<<blocks, 256>>>(loc0b, loc0, m); }, 20);
      float t_push = time_ms([&] { copy16<<<blocks, 256>>>(rem1, loc0, m); }, 20);
      float t_pull = time_ms([&] { copy16<<<blocks, 256>>>(loc0, rem1, m); }, 20);
      printf("{\"blocks\": %d, \"MiB\": %lld, \"local_store_GBs\": %.0f, \"remote_store_GBs\": %.0f, \"local_copy_GBs\": %.0f, \"push_copy_GBs\": %.0f, \"pull_copy_GBs\": %.0f, "
             "\"remote_store_us\": %.1f, \"push_us\": %.1f, \"pull_us\": %.1f}\n",
             blocks, sz >> 20, sz / t_ls / 1e6, sz / t_rs / 1e6, sz / t_lc / 1e6, sz / t_push / 1e6, sz / t_pull / 1e6, t_rs * 1e3, t_push * 1e3, t_pull * 1e3);
    }
  }
  {
    // 65536 rows of 64 B, row stride 1 KiB (4 MiB of payload inside a 64 MiB window)
    const long long nrows = 65536;
    float t_l = time_ms([&] { store_rows64<<<148 * 4, 256>>>(loc0, nrows, 64, 1.0); }, 20);
    float t_r = time_ms([&] { store_rows64<<<148 * 4, 256>>>(rem1, nrows, 64, 1.0); }, 20);
    printf("{\"pattern\": \"64B rows, 1KiB stride, 4MiB payload\", \"local_us\": %.1f, \"remote_us\": %.1f, \"remote_GBs\": %.0f}\n", t_l * 1e3, t_r * 1e3,
           nrows * 64 / t_r / 1e6);
  }
  fence_latency<<<1, 32>>>(rem1, d_out);
  long long h = 0;
  CK(cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost));
  fence_latency<<<1, 32>>>(rem1, d_out);
  CK(cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost));
  printf("{\"remote_store_plus_fence_sys_clk\": %lld}\n", h);
  fence_latency<<<1, 32>>>(loc0, d_out);
  CK(cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost));
  printf("{\"local_store_plus_fence_sys_clk\": %lld}\n", h);
  return 0;
}
