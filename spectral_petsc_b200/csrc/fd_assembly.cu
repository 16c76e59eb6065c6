// Device assembly of the finite-difference preconditioning matrices: one thread per interior node evaluates its
// (2d+1)-point row (fd_rows.h, the reference's arithmetic) and stores it at a closed-form CSR offset, so there is no
// scan, no index array and no atomics; a Newton step refreshes the values alone (SAME_NONZERO_PATTERN).
// Streaming work: reads (2+d) fields once through L1/L2, writes 12 B per entry.
#include "fd_assembly.h"

#include <climits>
#include <cmath>
#include <vector>

#include "../../include/spectral_b200.h"
#include "common.cuh"
#include "deriv.h"

namespace sb200 {

namespace {

template <int D>
__global__ void __launch_bounds__(256) fd_assemble_kernel(FdGrid G, FdFields F, int ncomp, int* __restrict__ rowptr, int* __restrict__ colidx,
                                                         double* __restrict__ vals, long long nnz_total) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= G.g) return;
  constexpr int MD = D > 0 ? D : SB200_FD_MAX_DIM;
  int k[MD];
  long long cols[2 * MD + 1];
  double v[2 * MD + 1];
  const long long node = fd_decode<D>(G, r, k);
  const int n = fd_row<D>(G, F, r, k, node, cols, v);
  const long long base = fd_row_offset<D>(G, k, r) * ncomp;  // the ncomp rows of this node are stored back to back
  for (int f = 0; f < ncomp; f++) {
    const long long o = base + (long long)f * n;
    if (rowptr) rowptr[r * ncomp + f] = (int)o;
    for (int e = 0; e < n; e++) {
      if (colidx) colidx[o + e] = (int)(cols[e] * ncomp + f);
      vals[o + e] = v[e];
    }
  }
  if (rowptr && r == G.g - 1) rowptr[G.g * ncomp] = (int)nnz_total;
}

}  // namespace

int FdAssembler::create(int d, const int* dim, int ncomp, FdAssembler** out) {
  SB_CHECK(d >= 1 && d <= SB200_FD_MAX_DIM, SB200_ERR_USER, "dimension count must be in [1,10] (elliptic.C:138)");
  for (int j = 0; j < d; j++) SB_CHECK(dim[j] >= 3, SB200_ERR_USER, "each extent must be >= 3 (needs an interior node)");
  FdAssembler* a = new FdAssembler;
  fd_grid_init(&a->G, d, dim);
  a->ncomp = ncomp;
  a->nrows = a->G.g * ncomp;
  a->nnz = fd_total_entries(a->G) * ncomp;
  if (a->nnz > (long long)INT_MAX) {
    delete a;
    set_last_error("preconditioning matrix has more than 2^31-1 entries (32-bit PetscInt CSR)");
    return SB200_ERR_SUP;
  }
  std::vector<double> x;
  for (int j = 0; j < d; j++)
    for (int i = 0; i < dim[j]; i++) x.push_back(cos(i * M_PI / (dim[j] - 1)));  // elliptic.C:279, stokes.C:297
  cudaError_t e = cudaMalloc((void**)&a->d_xtab, x.size() * sizeof(double));
  if (e == cudaSuccess) e = cudaMemcpy(a->d_xtab, x.data(), x.size() * sizeof(double), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    delete a;
    set_last_error(std::string("FdAssembler: ") + cudaGetErrorString(e));
    return SB200_ERR_CUDA;
  }
  *out = a;
  return 0;
}

FdAssembler::~FdAssembler() {
  if (d_xtab) cudaFree(d_xtab);
}

int FdAssembler::assemble(const double* eta, const double* deta, const double* const* gradu, int* d_rowptr, int* d_colidx, double* d_vals,
                          cudaStream_t s) const {
  SB_CHECK(eta && d_vals, SB200_ERR_ARG, "null pointer");
  SB_CHECK((d_rowptr == nullptr) == (d_colidx == nullptr), SB200_ERR_ARG, "pass both index arrays or neither (values-only refresh)");
  FdFields F;
  F.xtab = d_xtab;
  F.eta = eta;
  F.deta = deta;
  for (int j = 0; j < SB200_FD_MAX_DIM; j++) F.gradu[j] = (deta && gradu && j < G.d) ? gradu[j] : nullptr;
  const int threads = 256;
  const long long blocks = (G.g + threads - 1) / threads;
  if (G.d == 3) fd_assemble_kernel<3><<<(unsigned)blocks, threads, 0, s>>>(G, F, ncomp, d_rowptr, d_colidx, d_vals, nnz);
  else if (G.d == 2) fd_assemble_kernel<2><<<(unsigned)blocks, threads, 0, s>>>(G, F, ncomp, d_rowptr, d_colidx, d_vals, nnz);
  else fd_assemble_kernel<0><<<(unsigned)blocks, threads, 0, s>>>(G, F, ncomp, d_rowptr, d_colidx, d_vals, nnz);
  SB_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace sb200
