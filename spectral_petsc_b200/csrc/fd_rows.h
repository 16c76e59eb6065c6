// Rows of the finite-difference preconditioning matrices the reference assembles with MatSetValues:
//   FormJacobian    (elliptic.C:537-590)   scalar, with the eta' * grad(u0) terms
//   StokesPCSetUp0  (stokes.C:1160-1240)   velocity block, Dirichlet case (:1203-1224), one copy per component
// Both are a (2d+1)-point stencil on the interior nodes in the reference's walk order; neighbours that are
// Dirichlet nodes have a negative column id there and MatSetValues drops them.
//
// This header is plain C++ with host/device qualifiers only under nvcc: fd_assembly.cu runs it one thread per
// interior node; tests/cpp/fd_rows_host.cpp compiles the very same functions with g++ so the index arithmetic
// and the stencil are checked against the oracle without a GPU (a test harness, not a product path).
#pragma once

#ifdef __CUDACC__
#define SB_FD_HD __host__ __device__ __forceinline__
#define SB_FD_UNROLL _Pragma("unroll")
#else
#define SB_FD_HD inline
#define SB_FD_UNROLL
#endif

#define SB200_FD_MAX_DIM 10  // elliptic.C:138

namespace sb200 {

struct FdGrid {
  int d;
  int dim[SB200_FD_MAX_DIM];            // full extents (boundary nodes included)
  long long stride[SB200_FD_MAX_DIM];   // row-major strides of the full grid (last axis fastest)
  long long istride[SB200_FD_MAX_DIM];  // row-major strides of the interior grid (extents dim-2)
  long long g;                          // interior nodes = node-level rows
  int xoff[SB200_FD_MAX_DIM];           // axis j's node coordinates start at xtab[xoff[j]] (dim[j] values)
};

struct FdFields {
  const double* xtab;  // cos(i*pi/(dim[j]-1)) per axis, evaluated on the host like elliptic.C:279 / stokes.C:297
  const double* eta;   // m values
  const double* deta;  // m values, or null (Stokes velocity block: no eta' terms, stokes.C:1216-1218)
  const double* gradu[SB200_FD_MAX_DIM];  // gradu[j], m values each (only read when deta is set)
};

inline void fd_grid_init(FdGrid* G, int d, const int* dim) {
  G->d = d;
  long long m = 1, g = 1;
  int off = 0;
  for (int j = 0; j < d; j++) {
    G->dim[j] = dim[j];
    G->xoff[j] = off;
    off += dim[j];
  }
  for (int j = d - 1; j >= 0; j--) {
    G->stride[j] = m;
    G->istride[j] = g;
    m *= dim[j];
    g *= dim[j] - 2;
  }
  G->g = g;
}

// Entries of the whole node-level matrix: every interior node has its diagonal and, per axis, the two neighbours
// unless the node sits next to the boundary along that axis.
inline long long fd_total_entries(const FdGrid& G) {
  long long nnz = G.g;
  for (int j = 0; j < G.d; j++) {
    const long long n = G.dim[j] - 2;
    nnz += 2 * (n - 1) * (G.g / n);
  }
  return nnz;
}

// interior ordinal r (the global id of SetupBC, elliptic.C:408-409) -> interior indices k[j] in [0, dim[j]-2);
// returns the local (full-grid) index of the node
// (template parameter D: the dimension count when known at compile time, so the loops unroll and the per-row arrays
// stay in registers; D = 0 reads it from the grid)
template <int D = 0>
SB_FD_HD long long fd_decode(const FdGrid& G, long long r, int* k) {
  const int d = D > 0 ? D : G.d;
  long long node = 0;
SB_FD_UNROLL
  for (int j = 0; j < d; j++) {
    const long long q = r / G.istride[j];
    r -= q * G.istride[j];
    k[j] = (int)q;
    node += (q + 1) * G.stride[j];
  }
  return node;
}

// entries of this node's row
template <int D = 0>
SB_FD_HD int fd_row_entries(const FdGrid& G, const int* k) {
  const int d = D > 0 ? D : G.d;
  int n = 1;
SB_FD_UNROLL
  for (int j = 0; j < d; j++) n += (k[j] > 0) + (k[j] < G.dim[j] - 3);
  return n;
}

// Entries in the rows before r, in closed form (no scan, no index arrays): r full stencils minus, per axis, the
// predecessors in walk order whose index along that axis is the first (no M neighbour) or the last (no P neighbour)
// interior one.  Predecessors of k split by the first axis a where they differ (k'_a < k_a, later axes free).
template <int D = 0>
SB_FD_HD long long fd_row_offset(const FdGrid& G, const int* k, long long r) {
  const int d = D > 0 ? D : G.d;
  long long missing = 0;
SB_FD_UNROLL
  for (int j = 0; j < d; j++) {
    const int n = G.dim[j] - 2;
SB_FD_UNROLL
    for (int side = 0; side < 2; side++) {
      const int v = side ? n - 1 : 0;
      long long c = 0;
SB_FD_UNROLL
      for (int a = 0; a < j; a++) c += (long long)k[a] * (G.istride[a] / n);  // differs before j: k'_j free -> pinned to v
      if (v < k[j]) c += G.istride[j];                                        // differs at j with k'_j = v
      if (k[j] == v)
SB_FD_UNROLL
        for (int a = j + 1; a < d; a++) c += (long long)k[a] * G.istride[a];  // agrees through j, differs after
      missing += c;
    }
  }
  return r * (2 * d + 1) - missing;
}

// The row of interior node r: node-level column ids in increasing order and their values.  Order of the arithmetic is
// the reference's (diagonal accumulated over the axes j = 0..d-1, elliptic.C:565-575 / stokes.C:1206-1222).
// Returns the number of entries (<= 2d+1).
template <int D = 0>
SB_FD_HD int fd_row(const FdGrid& G, const FdFields& F, long long r, const int* k, long long node, long long* cols, double* vals) {
  const int d = D > 0 ? D : G.d;
  int nM = 0;
SB_FD_UNROLL
  for (int j = 0; j < d; j++) nM += (k[j] > 0);
  const int n = fd_row_entries<D>(G, k);
  // sorted layout: M neighbours of axes 0..d-1 (columns r - istride[j], increasing with j), the diagonal, then the P
  // neighbours of axes d-1..0 (columns r + istride[j])
  int posM = 0, posP = n - 1;
  double diag = 0.0;
  const double e0 = F.eta[node];
  const double de0 = F.deta ? F.deta[node] : 0.0;
SB_FD_UNROLL
  for (int j = 0; j < d; j++) {
    const long long iM = node - G.stride[j], iP = node + G.stride[j];
    const double* X = F.xtab + G.xoff[j];
    const double x0 = X[k[j] + 1], xMM = X[k[j]], xPP = X[k[j] + 2];
    const double xM = 0.5 * (xMM + x0), idxM = 1.0 / (x0 - xMM), xP = 0.5 * (x0 + xPP), idxP = 1.0 / (xPP - x0), idx = 1.0 / (xP - xM);
    const double eM = 0.5 * (F.eta[iM] + e0), eP = 0.5 * (F.eta[iP] + e0);
    double vM, vP;
    if (F.deta) {
      const double* gj = F.gradu[j];
      const double deM = 0.5 * (F.deta[iM] + de0), du0M = 0.5 * (gj[iM] + gj[node]);
      const double deP = 0.5 * (F.deta[iP] + de0), du0P = 0.5 * (gj[iP] + gj[node]);
      vM = -idx * (idxM * eM - 0.5 * deM * du0M);
      vP = -idx * (idxP * eP + 0.5 * deP * du0P);
      diag += idx * (idxP * eP + idxM * eM - 0.5 * (deP * du0P - deM * du0M));
    } else {
      vM = -idx * (idxM * eM);
      vP = -idx * (idxP * eP);
      diag += idx * (idxP * eP + idxM * eM);
    }
    if (k[j] > 0) {
      cols[posM] = r - G.istride[j];
      vals[posM] = vM;
      posM++;
    }
    if (k[j] < G.dim[j] - 3) {
      cols[posP] = r + G.istride[j];
      vals[posP] = vP;
      posP--;
    }
  }
  cols[nM] = r;
  vals[nM] = diag;
  return n;
}

}  // namespace sb200
