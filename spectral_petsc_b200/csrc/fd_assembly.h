// Internal: device assembly of the finite-difference preconditioning matrices (FormJacobian elliptic.C:537-590,
// StokesPCSetUp0 stokes.C:1160-1240) as CSR with 32-bit indices (PetscInt), rows in the global Vec order.
#pragma once
#include <cuda_runtime.h>

#include "fd_rows.h"

namespace sb200 {

struct FdAssembler {
  FdGrid G;
  int ncomp = 1;            // 1: scalar matrix (elliptic); d: one copy of the stencil per velocity component (Stokes)
  double* d_xtab = nullptr; // per-axis node coordinates
  long long nrows = 0, nnz = 0;

  static int create(int d, const int* dim, int ncomp, FdAssembler** out);
  ~FdAssembler();
  // rowptr (nrows+1) / colidx (nnz) may be null: SAME_NONZERO_PATTERN refresh of the values only (elliptic.C:588)
  int assemble(const double* eta, const double* deta, const double* const* gradu, int* d_rowptr, int* d_colidx, double* d_vals,
               cudaStream_t s) const;
};

}  // namespace sb200
