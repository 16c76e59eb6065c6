// The small vector operations the saddle-point preconditioners of stokes.C:1714-1817 are composed from, on device arrays:
// the index scatters between the interleaved global vector and its velocity / pressure parts (scatterGV / GP / VG / PG,
// stokes.C:867-877), VecAXPBY / VecScale, VecPointwiseDivide (PCJacobi on StokesMatGetDiagonalSchur) and MatNullSpaceRemove
// with the constant vector (stokes.C:1013-1023).  All are single HBM passes; the reduction is two-stage with a fixed
// summation order (deterministic, no atomics).  C ABI: include/spectral_b200.h, "vector helpers".
#include "../../include/spectral_b200.h"
#include "common.cuh"
#include "deriv.h"

namespace sb200 {
namespace {

constexpr int TPB = 256;
constexpr int MAX_PARTIALS = SB200_REDUCE_SCRATCH_DOUBLES;

int blocks_for(long long n, int cap) {
  long long b = (n + TPB - 1) / TPB;
  if (b < 1) b = 1;
  return (int)(b < cap ? b : cap);
}

// x: global AoS [v_0..v_{d-1}, p] per interior node -> v (nodes*d), p (nodes); either output may be null
__global__ void __launch_bounds__(TPB) split_kernel(long long nodes, int d, const double* __restrict__ x, double* __restrict__ v,
                                                    double* __restrict__ p) {
  const long long total = nodes * (d + 1), stride = (long long)gridDim.x * TPB;
  for (long long i = (long long)blockIdx.x * TPB + threadIdx.x; i < total; i += stride) {
    const long long q = i / (d + 1);
    const int k = (int)(i - q * (d + 1));
    const double a = x[i];
    if (k < d) {
      if (v) v[q * d + k] = a;
    } else if (p) {
      p[q] = a;
    }
  }
}

__global__ void __launch_bounds__(TPB) merge_kernel(long long nodes, int d, const double* __restrict__ v, const double* __restrict__ p,
                                                    double* __restrict__ x) {
  const long long total = nodes * (d + 1), stride = (long long)gridDim.x * TPB;
  for (long long i = (long long)blockIdx.x * TPB + threadIdx.x; i < total; i += stride) {
    const long long q = i / (d + 1);
    const int k = (int)(i - q * (d + 1));
    if (k < d) {
      if (v) x[i] = v[q * d + k];
    } else if (p) {
      x[i] = p[q];
    }
  }
}

// y = a x + b y; b == 0 never reads y (so y may be uninitialised, like VecAXPBY's fast paths)
__global__ void __launch_bounds__(TPB) axpby_kernel(long long n, double a, const double* __restrict__ x, double b, double* __restrict__ y) {
  const long long stride = (long long)gridDim.x * TPB;
  for (long long i = (long long)blockIdx.x * TPB + threadIdx.x; i < n; i += stride) y[i] = b == 0.0 ? a * x[i] : a * x[i] + b * y[i];
}

__global__ void __launch_bounds__(TPB) scale_kernel(long long n, double a, double* __restrict__ y) {
  const long long stride = (long long)gridDim.x * TPB;
  for (long long i = (long long)blockIdx.x * TPB + threadIdx.x; i < n; i += stride) y[i] = a * y[i];
}

__global__ void __launch_bounds__(TPB) divide_kernel(long long n, const double* __restrict__ x, const double* __restrict__ dg,
                                                     double* __restrict__ y) {
  const long long stride = (long long)gridDim.x * TPB;
  for (long long i = (long long)blockIdx.x * TPB + threadIdx.x; i < n; i += stride) y[i] = x[i] / dg[i];
}

// diag[r] = the stored entry (r, r) of a CSR matrix, 0 if the row has none (MatGetDiagonal on the SeqAIJ preconditioning matrices:
// what PCJacobi needs of them); rows are short (2d + 1 entries), one thread per row
__global__ void __launch_bounds__(TPB) csr_diagonal_kernel(long long nrows, const int* __restrict__ rowptr, const int* __restrict__ colidx,
                                                           const double* __restrict__ vals, double* __restrict__ diag) {
  const long long stride = (long long)gridDim.x * TPB;
  for (long long r = (long long)blockIdx.x * TPB + threadIdx.x; r < nrows; r += stride) {
    double v = 0.0;
    for (int q = rowptr[r]; q < rowptr[r + 1]; q++)
      if (colidx[q] == r) v = vals[q];
    diag[r] = v;
  }
}

__device__ __forceinline__ double block_sum_all(double v, double* sm) {  // result on every thread, fixed order
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) sm[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int k = 0; k < TPB / 32; k++) t += sm[k];
  return t;
}

// partial[b] = sum over this block's grid-stride share of x[off + i * stride_x], i < n
__global__ void __launch_bounds__(TPB) sum_partial_kernel(long long n, int stride_x, int off, const double* __restrict__ x,
                                                          double* __restrict__ partial) {
  __shared__ double sm[TPB / 32];
  double acc = 0.0;
  const long long stride = (long long)gridDim.x * TPB;
  for (long long i = (long long)blockIdx.x * TPB + threadIdx.x; i < n; i += stride) acc += x[off + i * stride_x];
  const double t = block_sum_all(acc, sm);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

// every block adds the np partials in the same order, then subtracts the mean from its share
__global__ void __launch_bounds__(TPB) sub_mean_kernel(long long n, int stride_x, int off, double* __restrict__ x,
                                                       const double* __restrict__ partial, int np) {
  __shared__ double sm[TPB / 32];
  double acc = 0.0;
  for (int b = threadIdx.x; b < np; b += TPB) acc += partial[b];
  const double mean = block_sum_all(acc, sm) / (double)n;
  const long long stride = (long long)gridDim.x * TPB;
  for (long long i = (long long)blockIdx.x * TPB + threadIdx.x; i < n; i += stride) x[off + i * stride_x] -= mean;
}

// out[0] = sum of the np partials (one block, fixed order), out[1] = n: what the slab ranks add up before the mean is formed
__global__ void __launch_bounds__(TPB) sum_final_kernel(const double* __restrict__ partial, int np, long long n, double* __restrict__ out) {
  __shared__ double sm[TPB / 32];
  double acc = 0.0;
  for (int b = threadIdx.x; b < np; b += TPB) acc += partial[b];
  const double t = block_sum_all(acc, sm);
  if (threadIdx.x == 0) {
    out[0] = t;
    out[1] = (double)n;
  }
}

// x[off + i*stride] -= sums[0] / sums[1]
__global__ void __launch_bounds__(TPB) shift_kernel(long long n, int stride_x, int off, double* __restrict__ x, const double* __restrict__ sums) {
  const double mean = sums[0] / sums[1];
  const long long stride = (long long)gridDim.x * TPB;
  for (long long i = (long long)blockIdx.x * TPB + threadIdx.x; i < n; i += stride) x[off + i * stride_x] -= mean;
}

}  // namespace
}  // namespace sb200

using namespace sb200;

extern "C" {

int sb200_vec_split(long long nodes, int d, const double* d_x, double* d_v, double* d_p, void* stream) {
  SB_CHECK(nodes >= 0 && d >= 1 && d_x && (d_v || d_p), SB200_ERR_ARG, "sb200_vec_split: bad arguments");
  if (nodes == 0) return 0;
  split_kernel<<<blocks_for(nodes * (d + 1), 148 * 16), TPB, 0, (cudaStream_t)stream>>>(nodes, d, d_x, d_v, d_p);
  count_launch();
  SB_CUDA(cudaGetLastError());
  return 0;
}

int sb200_vec_merge(long long nodes, int d, const double* d_v, const double* d_p, double* d_x, void* stream) {
  SB_CHECK(nodes >= 0 && d >= 1 && d_x && (d_v || d_p), SB200_ERR_ARG, "sb200_vec_merge: bad arguments");
  if (nodes == 0) return 0;
  merge_kernel<<<blocks_for(nodes * (d + 1), 148 * 16), TPB, 0, (cudaStream_t)stream>>>(nodes, d, d_v, d_p, d_x);
  count_launch();
  SB_CUDA(cudaGetLastError());
  return 0;
}

int sb200_vec_axpby(long long n, double a, const double* d_x, double b, double* d_y, void* stream) {
  SB_CHECK(n >= 0 && d_y && (d_x || a == 0.0), SB200_ERR_ARG, "sb200_vec_axpby: bad arguments");
  if (n == 0) return 0;
  if (a == 0.0)
    scale_kernel<<<blocks_for(n, 148 * 16), TPB, 0, (cudaStream_t)stream>>>(n, b, d_y);
  else
    axpby_kernel<<<blocks_for(n, 148 * 16), TPB, 0, (cudaStream_t)stream>>>(n, a, d_x, b, d_y);
  count_launch();
  SB_CUDA(cudaGetLastError());
  return 0;
}

int sb200_vec_pointwise_divide(long long n, const double* d_x, const double* d_diag, double* d_y, void* stream) {
  SB_CHECK(n >= 0 && d_x && d_diag && d_y, SB200_ERR_ARG, "sb200_vec_pointwise_divide: bad arguments");
  if (n == 0) return 0;
  divide_kernel<<<blocks_for(n, 148 * 16), TPB, 0, (cudaStream_t)stream>>>(n, d_x, d_diag, d_y);
  count_launch();
  SB_CUDA(cudaGetLastError());
  return 0;
}

int sb200_csr_diagonal(long long nrows, const int* d_rowptr, const int* d_colidx, const double* d_vals, double* d_diag, void* stream) {
  SB_CHECK(nrows >= 0 && d_rowptr && d_colidx && d_vals && d_diag, SB200_ERR_ARG, "sb200_csr_diagonal: bad arguments");
  if (nrows == 0) return 0;
  csr_diagonal_kernel<<<blocks_for(nrows, 148 * 16), TPB, 0, (cudaStream_t)stream>>>(nrows, d_rowptr, d_colidx, d_vals, d_diag);
  count_launch();
  SB_CUDA(cudaGetLastError());
  return 0;
}

int sb200_vec_sum_count(long long n, int stride, int offset, const double* d_x, double* d_scratch, double* d_out2, void* stream) {
  SB_CHECK(n >= 0 && stride >= 1 && offset >= 0 && offset < stride && d_scratch && d_out2 && (d_x || n == 0), SB200_ERR_ARG, "sb200_vec_sum_count: bad arguments");
  const int np = n > 0 ? blocks_for(n, MAX_PARTIALS) : 0;
  if (np > 0) sum_partial_kernel<<<np, TPB, 0, (cudaStream_t)stream>>>(n, stride, offset, d_x, d_scratch);
  sum_final_kernel<<<1, TPB, 0, (cudaStream_t)stream>>>(d_scratch, np, n, d_out2);
  count_launch(np > 0 ? 2 : 1);
  SB_CUDA(cudaGetLastError());
  return 0;
}

int sb200_vec_shift_mean(long long n, int stride, int offset, double* d_x, const double* d_sums2, void* stream) {
  SB_CHECK(n >= 0 && stride >= 1 && offset >= 0 && offset < stride && d_sums2 && (d_x || n == 0), SB200_ERR_ARG, "sb200_vec_shift_mean: bad arguments");
  if (n == 0) return 0;
  shift_kernel<<<blocks_for(n, 148 * 16), TPB, 0, (cudaStream_t)stream>>>(n, stride, offset, d_x, d_sums2);
  count_launch();
  SB_CUDA(cudaGetLastError());
  return 0;
}

int sb200_vec_remove_mean(long long n, int stride, int offset, double* d_x, double* d_scratch, void* stream) {
  SB_CHECK(n >= 0 && stride >= 1 && offset >= 0 && offset < stride && d_x && d_scratch, SB200_ERR_ARG, "sb200_vec_remove_mean: bad arguments");
  if (n == 0) return 0;
  const int np = blocks_for(n, MAX_PARTIALS);
  sum_partial_kernel<<<np, TPB, 0, (cudaStream_t)stream>>>(n, stride, offset, d_x, d_scratch);
  sub_mean_kernel<<<blocks_for(n, 148 * 16), TPB, 0, (cudaStream_t)stream>>>(n, stride, offset, d_x, d_scratch, np);
  count_launch(2);
  SB_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
