// Internal: symmetric (peer-mapped) device arena for the slab-partitioned operators.
//
// One process per GPU.  Every rank allocates the SAME sequence of buffers from one cudaMalloc'ed
// arena, exports the arena with a CUDA IPC handle and maps the peers' arenas, so a local pointer
// translates to the peer's copy by base-offset arithmetic.  Kernels then read (pull) or write (push)
// peer memory directly over NVLink; ordering across GPUs is by epoch flags that live at the start of
// the arena (system-scope release/acquire), never by host synchronisation.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

#define SB200_MAX_RANKS 8

namespace sb200 {

// Flag words (unsigned long long) at the head of every arena.
enum SymmFlag {
  SYMM_BAR = 0,     // [0..7]   generic barrier: peer q wrote its epoch into slot q
  SYMM_READY = 8,   // [8..15]  fused MatMult: peer q's staged input vector is ready
  SYMM_DONE = 16,   // [16..23] fused MatMult: peer q has pushed all its axis-0 results
  SYMM_TIMEOUT = 24,  // number of flag waits that gave up (0 in a healthy run)
  SYMM_HOSTFAIL = 25, // device address of a host-mapped word that a timed-out wait sets (sticky failure seen by the host without a sync)
  SYMM_AR = 32,       // [32..39] small all-reduce (KSP dot products): peer q's contribution is in my slot q
  SYMM_NFLAGS = 48,
};

struct SymmArena {
  int rank = 0, nranks = 1;
  char* base = nullptr;
  size_t bytes = 0, used = 0;
  char* peer[SB200_MAX_RANKS] = {};  // peer[rank] == base; others set by attach()
  bool opened[SB200_MAX_RANKS] = {};
  unsigned long long bar_epoch = 0;
  unsigned long long* h_fail = nullptr;  // pinned, device-mapped: non-zero once any device-side flag wait of this rank has timed out

  int init(size_t bytes, int rank, int nranks);
  void destroy();
  // 256-byte aligned bump allocation; returns nullptr when the arena is exhausted.
  void* alloc(size_t nbytes);
  double* alloc_doubles(size_t n) { return (double*)alloc(n * sizeof(double)); }
  template <class T>
  T* on(int q, const T* p) const {
    return (T*)(peer[q] + ((const char*)p - base));
  }
  unsigned long long* flags(int q) const { return (unsigned long long*)peer[q]; }
  bool attached() const;
  int export_handle(void* handle64) const;
  int attach(int q, const void* handle64);
  int attach_ptr(int q, void* mapped_base);
  // Device-side barrier over all ranks on `s` (no-op for nranks == 1): everything enqueued on the
  // peers' streams before their matching barrier is visible to kernels enqueued after ours.
  int barrier(cudaStream_t s);
  // Synchronises `s` and returns the number of device-side flag waits that timed out so far.
  int timeouts(cudaStream_t s, unsigned long long* n);
  // Sticky failure, readable without synchronising: true once a device-side wait has given up (a peer never arrived, or the
  // ranks' collective calls went out of step).  Every slab entry point checks it first and refuses to run: results after a
  // timed-out wait are undefined.
  bool failed() const { return h_fail && *(volatile unsigned long long*)h_fail != 0; }
};

// Device helpers -----------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Flag store without its own fence: issue ONE __threadfence_system() and then publish to every peer with
// these (a st.release.sys per peer costs a full system-scope fence each, ~2 us, and they serialise).
__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Bounded spin on an epoch flag: a rank that never shows up (mismatched collective calls, a dead peer)
// must not hang the GPU.  After ~4 s the wait gives up and records the failure in the flag block's
// SYMM_TIMEOUT word (sb200_*_slab_status reports it); results are then undefined but the kernel exits.
#define SB200_SPIN_LIMIT (1ll << 33)
__device__ __forceinline__ void symm_record_timeout(unsigned long long* local_flags) {
  atomicAdd(local_flags + SYMM_TIMEOUT, 1ull);
  unsigned long long* h = reinterpret_cast<unsigned long long*>(local_flags[SYMM_HOSTFAIL]);
  if (h) *(volatile unsigned long long*)h = 1ull;  // host-mapped: the next API call on this context fails loudly
}
__device__ __forceinline__ void spin_until(const unsigned long long* f, unsigned long long epoch,
                                           unsigned long long* local_flags) {
  if (ld_acquire_sys(f) >= epoch) return;
  const long long t0 = clock64();
  while (ld_acquire_sys(f) < epoch) {
    if (clock64() - t0 > SB200_SPIN_LIMIT) {
      symm_record_timeout(local_flags);
      return;
    }
  }
}

// Flag pointers of every rank, passed to kernels by value.
struct SymmFlags {
  unsigned long long* f[SB200_MAX_RANKS];  // f[q] = flag block of rank q (f[rank] is local)
  int rank, nranks;
};

}  // namespace sb200
