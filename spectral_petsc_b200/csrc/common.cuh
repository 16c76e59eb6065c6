// Shared device helpers for the sm_100a kernels: FP64 DMMA (mma.sync m8n8k4 -> SASS DMMA.8x8x4),
// cp.async staging, error plumbing.  No torch types; plain CUDA runtime.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>

namespace sb200 {

void set_last_error(const std::string& msg);

#define SB_CUDA(expr)                                                                     \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      sb200::set_last_error(std::string(#expr) + ": " + cudaGetErrorString(_e));          \
      return SB200_ERR_CUDA;                                                              \
    }                                                                                     \
  } while (0)

#define SB_CHECK(cond, code, msg)                                                         \
  do {                                                                                    \
    if (!(cond)) {                                                                        \
      sb200::set_last_error(msg);                                                         \
      return (code);                                                                      \
    }                                                                                     \
  } while (0)

#define SB_TRY(expr)                                                                      \
  do {                                                                                    \
    int _r = (expr);                                                                      \
    if (_r != 0) return _r;                                                               \
  } while (0)

// D(8x8) += A(8x4, row) * B(4x8, col).  Fragment ownership (lane = 4*g + t):
//   a = A[g][t], b = B[t][g], c0 = C[g][2t], c1 = C[g][2t+1].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// 8-byte cp.async with zero fill when !valid (src-size = 0).
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc, bool valid) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  int sz = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gsrc), "r"(sz));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gsrc), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

}  // namespace sb200
