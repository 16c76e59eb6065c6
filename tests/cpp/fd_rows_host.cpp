// Test harness only: runs the host/device row functions of spectral_petsc_b200/csrc/fd_rows.h in a plain CPU loop so
// tests/test_fd_rows_cpu.py can compare the index arithmetic and the stencil with the oracle (and, through it, with
// the reference's FormJacobian / StokesPCSetUp0) without a GPU.  The product path is fd_assembly.cu (one CUDA thread
// per row); nothing in the library calls this file.
#include <cmath>
#include <vector>

#include "../../spectral_petsc_b200/csrc/fd_rows.h"

using namespace sb200;

extern "C" long long fd_host_sizes(int d, const int* dim, int ncomp, long long* nrows) {
  FdGrid G;
  fd_grid_init(&G, d, dim);
  *nrows = G.g * ncomp;
  return fd_total_entries(G) * ncomp;
}

// gradu: d arrays of m doubles back to back (or null with deta null); same store pattern as fd_assemble_kernel
extern "C" int fd_host_assemble(int d, const int* dim, int ncomp, const double* eta, const double* deta, const double* gradu, int* rowptr,
                                int* colidx, double* vals, int unrolled) {
  FdGrid G;
  fd_grid_init(&G, d, dim);
  long long m = 1;
  for (int j = 0; j < d; j++) m *= dim[j];
  std::vector<double> x;
  for (int j = 0; j < d; j++)
    for (int i = 0; i < dim[j]; i++) x.push_back(cos(i * M_PI / (dim[j] - 1)));
  FdFields F;
  F.xtab = x.data();
  F.eta = eta;
  F.deta = deta;
  for (int j = 0; j < SB200_FD_MAX_DIM; j++) F.gradu[j] = (deta && gradu && j < d) ? gradu + j * m : nullptr;
  const long long nnz = fd_total_entries(G) * ncomp;
  for (long long r = 0; r < G.g; r++) {
    int k[SB200_FD_MAX_DIM];
    long long cols[2 * SB200_FD_MAX_DIM + 1];
    double v[2 * SB200_FD_MAX_DIM + 1];
    long long node, base;
    int n;
    if (unrolled && d == 3) {  // the instantiation the 3-D kernel uses
      node = fd_decode<3>(G, r, k);
      n = fd_row<3>(G, F, r, k, node, cols, v);
      base = fd_row_offset<3>(G, k, r) * ncomp;
    } else if (unrolled && d == 2) {
      node = fd_decode<2>(G, r, k);
      n = fd_row<2>(G, F, r, k, node, cols, v);
      base = fd_row_offset<2>(G, k, r) * ncomp;
    } else {
      node = fd_decode<0>(G, r, k);
      n = fd_row<0>(G, F, r, k, node, cols, v);
      base = fd_row_offset<0>(G, k, r) * ncomp;
    }
    for (int f = 0; f < ncomp; f++) {
      const long long o = base + (long long)f * n;
      if (o < 0 || o + n > nnz) return 1;  // an offset outside the matrix: the closed form is wrong
      rowptr[r * ncomp + f] = (int)o;
      for (int e = 0; e < n; e++) {
        colidx[o + e] = (int)(cols[e] * ncomp + f);
        vals[o + e] = v[e];
      }
    }
  }
  rowptr[G.g * ncomp] = (int)nnz;
  return 0;
}
