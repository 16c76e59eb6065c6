// Internal: device-side context for the elliptic.C shells (replaces MatElliptic, elliptic.C:78-86).
#pragma once
#include <cuda_runtime.h>

#include <vector>

#define SB200_MAX_DIM 10  // elliptic.C:138 "Maximum number of dimensions"

namespace sb200 {

// Padded device copy of one P x P differentiation matrix (shared by all axes of equal extent).
struct DiffMatrix {
  int P = 0, Pp = 0;
  double* d_D = nullptr;
  double* d_Ae = nullptr;  // even-odd halves (P even), [P/2][P/2] row-major
  double* d_Bo = nullptr;
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
  int init(int d, const int* dim);
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
  long long* trace = nullptr;  // debug: phase time stamps (SB200_TRACE builds)
  unsigned* sync = nullptr;  // ticket / completion counters of the persistent kernel
  DiffMatrix* Dax[SB200_MAX_DIM] = {};
  std::vector<DiffMatrix*> owned;

  static int create(int d, const int* dim, EllipticCtx** out);
  int init(int d, const int* dim);
  ~EllipticCtx();
  int deriv(int axis, const double* x, double* y, const double* yin, int mode, cudaStream_t s);
  int pad(const double* U, bool with_dirichlet, double* local, cudaStream_t s);
  int crop(const double* local, const double* rhs, double* V, cudaStream_t s);
  int matmult(const double* U, double* V, cudaStream_t s);
  int function(const double* U, double* F, cudaStream_t s);
};

bool elliptic_fused_supported(const EllipticCtx& e);
int elliptic_matmult_fused(EllipticCtx& e, const double* U, double* V, cudaStream_t s);
bool elliptic_persist_supported(const EllipticCtx& e);
int elliptic_matmult_persist(EllipticCtx& e, const double* U, double* V, cudaStream_t s);

}  // namespace sb200
