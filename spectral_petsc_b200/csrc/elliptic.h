// Internal: device-side context for the elliptic.C shells (replaces MatElliptic, elliptic.C:78-86).
#pragma once
#include <cuda_runtime.h>

#include <vector>

#include "symm.h"

#define SB200_MAX_DIM 10  // elliptic.C:138 "Maximum number of dimensions"
#define SB200_EO_MAX_P 160  // largest extent the generalised even-odd derivative kernel holds in shared memory (deriv_eo.cu)

namespace sb200 {

// Padded device copy of one P x P differentiation matrix (shared by all axes of equal extent).
struct DiffMatrix {
  int P = 0, Pp = 0;
  double* d_D = nullptr;
  double* d_Ae = nullptr;  // even-odd halves (P even), [P/2][P/2] row-major
  double* d_Bo = nullptr;
  // any P <= SB200_EO_MAX_P: the even-odd halves zero padded to HP x HP, HP = 8 * ceil(ceil(P/2) / 8) (cheb_matrix.h); for
  // P % 16 == 0 the same numbers as d_Ae / d_Bo
  int HP = 0;
  double* d_Aep = nullptr;
  double* d_Bop = nullptr;
  static int create(int P, DiffMatrix* out);
  void destroy();
};

// Row-major grid description, passed by value to kernels.
struct GridDesc {
  int d;
  int dim[SB200_MAX_DIM];
  long long stride[SB200_MAX_DIM];   // row-major strides of the full grid
  long long istride[SB200_MAX_DIM];  // row-major strides of the interior (dim-2) grid
  long long m, g;                    // local nodes, interior (global) nodes
  // Slab view of axis 0 (multi-GPU partition along the outermost axis): the local arrays hold planes
  // [i0, i0 + dim[0]) of a grid whose axis 0 has n0g nodes; goff = interior nodes in the planes before
  // i0 (the local global-vector starts there).  Single GPU: i0 = 0, n0g = dim[0], goff = 0.
  int i0, n0g;
  long long goff;
  int init(int d, const int* dim);
  // dim_global: the full grid; this rank keeps planes [rank*nloc, (rank+1)*nloc), nloc = dim_global[0]/nranks
  int init_slab(int d, const int* dim_global, int rank, int nranks);
  __host__ __device__ int gext(int j) const { return j == 0 ? n0g : dim[j]; }
  __host__ __device__ int gidx(int j, int i) const { return j == 0 ? i + i0 : i; }
};

struct EllipticCtx {
  GridDesc gd;
  int nw = 0;
  double* w[SB200_MAX_DIM + 2] = {};  // c->w (elliptic.C:263)
  double* gradu[SB200_MAX_DIM] = {};  // c->gradu
  double* eta = nullptr;
  double* deta = nullptr;
  double* dirichlet = nullptr;  // c->dirichlet (nd values, walk order)
  double* b = nullptr;          // ac->b
  double gamma = 0.0, exponent = 2.0;
  int path = 0;
  const char* last_kernel = "none";  // which kernel path the last MatMult_Elliptic ran (reported by bench.py)
  long long* trace = nullptr;  // debug: phase time stamps (SB200_TRACE builds)
  unsigned* sync = nullptr;  // ticket / completion counters of the persistent kernel
  DiffMatrix* Dax[SB200_MAX_DIM] = {};
  std::vector<DiffMatrix*> owned;
  // slab partition (nranks == 1: plain single-GPU context; all device arrays live in the arena)
  SymmArena arena;
  int gdim[SB200_MAX_DIM] = {};     // global extents
  long long gtot = 0;               // global Vec length over all ranks
  double* Wp = nullptr;             // axis-0 pencil [P][R0/G] of the padded input vector, pushed by all ranks
  double* eta_p = nullptr;          // axis-0 pencil copies of eta / deta / gradu[0] for the fused MatMult
  double* deta_p = nullptr;
  double* g0_p = nullptr;
  double* Xp = nullptr;             // pencil operand / result buffers of the generic axis-0 derivative
  double* Yp = nullptr;
  bool pencil_valid = false;
  unsigned long long mm_epoch = 0;  // fused MatMult epoch (SYMM_READY / SYMM_DONE flags)
  struct SlabMaps* tmaps = nullptr; // TMA descriptors of every rank's part[0] field (persist.h), built at the first fused slab MatMult
  bool tmaps_tried = false;

  static int create(int d, const int* dim, int rank, int nranks, EllipticCtx** out);
  int init(int d, const int* dim, int rank, int nranks);
  ~EllipticCtx();
  int refresh_pencils(cudaStream_t s);
  int deriv(int axis, const double* x, double* y, const double* yin, int mode, cudaStream_t s);
  bool fusable() const;
  struct EoLineMap line_map(int axis) const;
  struct DerivParams job(int axis, const double* x, double* y, const double* yin, int mode) const;
  int fused_tail(const double* rhs, double* V, cudaStream_t s);
  int pad(const double* U, bool with_dirichlet, double* local, cudaStream_t s);
  int crop(const double* local, const double* rhs, double* V, cudaStream_t s);
  int matmult(const double* U, double* V, cudaStream_t s);
  int matmult_generic(const double* U, double* V, cudaStream_t s);
  // path 4 (opt-in, single GPU): the generic path's launches captured once into a CUDA graph on fixed staging vectors and
  // replayed per application - for the small grids (BASELINE configs 1 and 3) where 2d + 3 launches of a few us each are the cost
  int matmult_graph(const double* U, double* V, cudaStream_t s);
  cudaGraphExec_t gexec = nullptr;
  cudaStream_t gstream = nullptr;
  double* gU = nullptr;
  double* gV = nullptr;
  int gnodes = 0;
  int function(const double* U, double* F, cudaStream_t s);
};

bool elliptic_fused_supported(const EllipticCtx& e);
int elliptic_matmult_fused(EllipticCtx& e, const double* U, double* V, cudaStream_t s);
bool elliptic_persist_supported(const EllipticCtx& e);
int elliptic_matmult_persist(EllipticCtx& e, const double* U, double* V, cudaStream_t s);
void free_slab_maps(struct SlabMaps* m);
bool elliptic_slab_fused_supported(const EllipticCtx& e);
int elliptic_matmult_slab_fused(EllipticCtx& e, const double* U, double* V, cudaStream_t s);
// x_pencil[m][nl] = x_slab(plane m, line rank*R/G + nl) pulled from the owners of the planes
int slab_to_pencil(const SymmArena& a, const double* x_slab, double* x_pencil, int P, long long R, cudaStream_t s);

}  // namespace sb200
