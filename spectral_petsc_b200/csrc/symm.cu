// Symmetric peer-mapped arena + device-side barrier (symm.h).
#include "symm.h"

#include <cstring>

#include "../../include/spectral_b200.h"
#include "common.cuh"
#include "deriv.h"

namespace sb200 {

namespace {

// One warp: lane q publishes this rank's epoch in slot `rank` of peer q, then waits until peer q's
// epoch shows up in the local slot q.  The fence orders every earlier write of this GPU (stream
// order already completed them) before the flag becomes visible at system scope.
__global__ void barrier_kernel(SymmFlags sf, unsigned long long epoch) {
  const int q = threadIdx.x;
  __threadfence_system();
  if (q < sf.nranks) {
    st_release_sys(sf.f[q] + SYMM_BAR + sf.rank, epoch);
    spin_until(sf.f[sf.rank] + SYMM_BAR + q, epoch, sf.f[sf.rank]);
  }
  __threadfence_system();
}

}  // namespace

int SymmArena::init(size_t nbytes, int rank_, int nranks_) {
  SB_CHECK(nranks_ >= 1 && nranks_ <= SB200_MAX_RANKS && rank_ >= 0 && rank_ < nranks_, SB200_ERR_USER,
           "slab partition: rank / nranks out of range (at most 8 ranks)");
  rank = rank_;
  nranks = nranks_;
  bytes = nbytes + SYMM_NFLAGS * sizeof(unsigned long long) + 4096;
  SB_CUDA(cudaMalloc((void**)&base, bytes));
  SB_CUDA(cudaMemset(base, 0, SYMM_NFLAGS * sizeof(unsigned long long)));
  SB_CUDA(cudaHostAlloc((void**)&h_fail, sizeof(unsigned long long), cudaHostAllocMapped));
  *h_fail = 0;
  {
    unsigned long long* d_fail = nullptr;
    SB_CUDA(cudaHostGetDevicePointer((void**)&d_fail, h_fail, 0));
    const unsigned long long v = (unsigned long long)(uintptr_t)d_fail;
    SB_CUDA(cudaMemcpy(base + SYMM_HOSTFAIL * sizeof(unsigned long long), &v, sizeof(v), cudaMemcpyHostToDevice));
  }
  SB_CUDA(cudaDeviceSynchronize());
  used = (SYMM_NFLAGS * sizeof(unsigned long long) + 255) / 256 * 256;
  for (int q = 0; q < SB200_MAX_RANKS; q++) {
    peer[q] = nullptr;
    opened[q] = false;
  }
  peer[rank] = base;
  return 0;
}

void SymmArena::destroy() {
  for (int q = 0; q < SB200_MAX_RANKS; q++)
    if (opened[q] && peer[q]) cudaIpcCloseMemHandle(peer[q]);
  if (base) cudaFree(base);
  base = nullptr;
  if (h_fail) cudaFreeHost(h_fail);
  h_fail = nullptr;
}

void* SymmArena::alloc(size_t nbytes) {
  const size_t a = (nbytes + 255) / 256 * 256;
  if (used + a > bytes) return nullptr;
  void* p = base + used;
  used += a;
  return p;
}

bool SymmArena::attached() const {
  for (int q = 0; q < nranks; q++)
    if (!peer[q]) return false;
  return true;
}

int SymmArena::export_handle(void* handle64) const {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  SB_CHECK(handle64 && base, SB200_ERR_ARG, "null pointer");
  cudaIpcMemHandle_t h;
  SB_CUDA(cudaIpcGetMemHandle(&h, base));
  std::memcpy(handle64, &h, sizeof(h));
  return 0;
}

int SymmArena::attach(int q, const void* handle64) {
  SB_CHECK(handle64 && q >= 0 && q < nranks, SB200_ERR_ARG, "bad peer rank");
  if (q == rank) return 0;
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle64, sizeof(h));
  void* p = nullptr;
  SB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  peer[q] = (char*)p;
  opened[q] = true;
  return 0;
}

int SymmArena::attach_ptr(int q, void* mapped_base) {
  SB_CHECK(mapped_base && q >= 0 && q < nranks, SB200_ERR_ARG, "bad peer rank");
  if (q != rank) peer[q] = (char*)mapped_base;
  return 0;
}

int SymmArena::barrier(cudaStream_t s) {
  if (nranks == 1) return 0;
  SB_CHECK(attached(), SB200_ERR_USER, "slab partition: peers are not attached (exchange the IPC handles first)");
  SymmFlags sf;
  for (int q = 0; q < SB200_MAX_RANKS; q++) sf.f[q] = q < nranks ? flags(q) : nullptr;
  sf.rank = rank;
  sf.nranks = nranks;
  barrier_kernel<<<1, 32, 0, s>>>(sf, ++bar_epoch);
  count_launch();
  SB_CUDA(cudaGetLastError());
  return 0;
}

int SymmArena::timeouts(cudaStream_t s, unsigned long long* n) {
  SB_CUDA(cudaMemcpyAsync(n, base + SYMM_TIMEOUT * sizeof(unsigned long long), sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
  SB_CUDA(cudaStreamSynchronize(s));
  return 0;
}

}  // namespace sb200
