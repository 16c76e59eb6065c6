// C ABI (include/spectral_b200.h): thin extern "C" layer over the device contexts.
#include <atomic>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/spectral_b200.h"
#include "cheb_matrix.h"
#include "common.cuh"
#include "deriv.h"
#include "elliptic.h"
#include "fd_assembly.h"

namespace sb200 {

static thread_local std::string g_last_error;
static std::atomic<long long> g_launches{0};

void set_last_error(const std::string& msg) { g_last_error = msg; }
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace sb200

using namespace sb200;

// ChebCtx replacement: (rank, tr, dims) -> (O, P, R) factorisation + the matrix.
struct sb200_cheb {
  int rank, tr;
  std::vector<int> dims;
  long long N, O, R;
  DiffMatrix D;
  double* d_x = nullptr;  // staging for *_host
  double* d_y = nullptr;
  unsigned* sync = nullptr;  // counters of the even-odd derivative kernel
};

struct sb200_elliptic {
  EllipticCtx* c = nullptr;
  double* d_in = nullptr;  // staging for *_host
  double* d_out = nullptr;
  double* h_pin_in = nullptr;
  double* h_pin_out = nullptr;
  // host-buffer queue (sb200_elliptic_matmult_host_submit / _wait): SB200_HOST_QUEUE_DEPTH slots, each with its own
  // device in / out vectors; copy-in, compute and copy-out run on three non-blocking streams chained by events
  struct HostSlot {
    double* d_in = nullptr;
    double* d_out = nullptr;
    cudaEvent_t in_done = nullptr, op_done = nullptr, out_done = nullptr;
  };
  HostSlot q[SB200_HOST_QUEUE_DEPTH];
  cudaStream_t q_in = nullptr, q_op = nullptr, q_out = nullptr;
  long long q_submitted = 0, q_waited = 0;
  // Ordering between FormFunction (any stream; it rewrites eta / deta / gradu) and the queue's applications (which read them):
  // state_ev is recorded behind every FormFunction and waited for by every submission; a FormFunction waits for the last
  // submitted application's op_done before it touches the state.
  cudaEvent_t state_ev = nullptr;
  bool state_recorded = false;
  FdAssembler* fd = nullptr;  // FormJacobian's matrix (built on first use)
};

extern "C" {

int sb200_version(void) { return 100; }

const char* sb200_last_error(void) { return g_last_error.c_str(); }

long long sb200_launch_count(void) { return g_launches.load(); }

int sb200_device_count(int* n) {
  SB_CHECK(n, SB200_ERR_ARG, "null pointer");
  cudaError_t e = cudaGetDeviceCount(n);
  if (e != cudaSuccess) {
    *n = 0;
    set_last_error(std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
    return SB200_ERR_CUDA;
  }
  return 0;
}

int sb200_set_device(int ordinal) {
  SB_CUDA(cudaSetDevice(ordinal));
  return 0;
}

int sb200_malloc(void** d_ptr, size_t bytes) {
  SB_CHECK(d_ptr, SB200_ERR_ARG, "null pointer");
  SB_CUDA(cudaMalloc(d_ptr, bytes ? bytes : 8));
  return 0;
}

int sb200_free(void* d_ptr) {
  if (d_ptr) SB_CUDA(cudaFree(d_ptr));
  return 0;
}

int sb200_memcpy_h2d(void* d_dst, const void* h_src, size_t bytes, void* stream) {
  SB_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return 0;
}

int sb200_memcpy_d2h(void* h_dst, const void* d_src, size_t bytes, void* stream) {
  SB_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return 0;
}

int sb200_memcpy_d2d(void* d_dst, const void* d_src, size_t bytes, void* stream) {
  SB_CUDA(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}

int sb200_memset0(void* d_dst, size_t bytes, void* stream) {
  SB_CUDA(cudaMemsetAsync(d_dst, 0, bytes, (cudaStream_t)stream));
  return 0;
}

int sb200_stream_sync(void* stream) {
  SB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}

// ---- Chebyshev ---------------------------------------------------------------------------
int sb200_cheb_create(int rank, int tr, const int* dims, long long n_total, sb200_cheb** out) {
  SB_CHECK(out && dims, SB200_ERR_ARG, "null pointer");
  *out = nullptr;
  if (n_total < 2) {  // chebyshev.c:98
    set_last_error("n = " + std::to_string(n_total) + " but must be >= 2");
    return SB200_ERR_USER;
  }
  SB_CHECK(rank >= 1, SB200_ERR_USER, "rank must be >= 1");
  SB_CHECK(0 <= tr && tr < rank, SB200_ERR_USER, "tdim out of range");  // chebyshev.c:106
  long long stride = 1;
  for (int r = 0; r < rank; r++) {
    SB_CHECK(dims[r] >= 1, SB200_ERR_USER, "extents must be positive");
    stride *= dims[r];
  }
  if (n_total != stride) {  // chebyshev.c:122
    set_last_error("dimensions do not agree: n = " + std::to_string(n_total) + " but stride = " + std::to_string(stride));
    return SB200_ERR_USER;
  }
  SB_CHECK(dims[tr] >= 2, SB200_ERR_USER, "transformed extent must be >= 2");
  sb200_cheb* c = new sb200_cheb();
  c->rank = rank;
  c->tr = tr;
  c->dims.assign(dims, dims + rank);
  c->N = n_total;
  c->R = 1;
  for (int r = tr + 1; r < rank; r++) c->R *= dims[r];
  c->O = n_total / (c->R * dims[tr]);
  int rc = DiffMatrix::create(dims[tr], &c->D);
  if (rc) {
    delete c;
    return rc;
  }
  cudaError_t ce = cudaMalloc((void**)&c->sync, 64);
  if (ce == cudaSuccess) ce = cudaMemset(c->sync, 0, 64);
  if (ce != cudaSuccess) {
    set_last_error(std::string("MatCreateCheb: ") + cudaGetErrorString(ce));
    sb200_cheb_destroy(c);
    return SB200_ERR_CUDA;
  }
  *out = c;
  return 0;
}

int sb200_cheb_apply(sb200_cheb* c, const double* d_x, double* d_y, void* stream) {
  SB_CHECK(c && d_x && d_y, SB200_ERR_ARG, "null pointer");
  DerivParams p;
  p.D = c->D.d_D;
  p.Ae = c->D.d_Aep;
  p.Bo = c->D.d_Bop;
  p.HP = c->D.HP;
  p.sync = c->sync;
  p.P = c->D.P;
  p.Pp = c->D.Pp;
  p.x = d_x;
  p.y = d_y;
  p.yin = nullptr;
  p.O = c->O;
  p.R = c->R;
  p.xs = p.ys = 1;
  p.xoff = p.yoff = 0;
  p.mode = DERIV_STORE;
  return deriv_apply(p, (cudaStream_t)stream);
}

int sb200_cheb_apply_host(sb200_cheb* c, const double* h_x, double* h_y) {
  SB_CHECK(c && h_x && h_y, SB200_ERR_ARG, "null pointer");
  const size_t bytes = (size_t)c->N * sizeof(double);
  if (!c->d_x) SB_CUDA(cudaMalloc((void**)&c->d_x, bytes));
  if (!c->d_y) SB_CUDA(cudaMalloc((void**)&c->d_y, bytes));
  SB_CUDA(cudaMemcpyAsync(c->d_x, h_x, bytes, cudaMemcpyHostToDevice, 0));
  SB_TRY(sb200_cheb_apply(c, c->d_x, c->d_y, nullptr));
  SB_CUDA(cudaMemcpyAsync(h_y, c->d_y, bytes, cudaMemcpyDeviceToHost, 0));
  SB_CUDA(cudaStreamSynchronize(0));
  return 0;
}

int sb200_cheb_destroy(sb200_cheb* c) {
  if (!c) return 0;
  c->D.destroy();
  if (c->sync) cudaFree(c->sync);
  if (c->d_x) cudaFree(c->d_x);
  if (c->d_y) cudaFree(c->d_y);
  delete c;
  return 0;
}

int sb200_cheb_matrix(int P, double* h_D) {
  SB_CHECK(P >= 2 && h_D, SB200_ERR_ARG, "bad arguments");
  std::vector<double> D = cgl_diff_matrix(P);
  std::memcpy(h_D, D.data(), D.size() * sizeof(double));
  return 0;
}

int sb200_cheb_even_odd(int P, int* HP, double* h_Ae, double* h_Bo) {
  SB_CHECK(P >= 2 && HP, SB200_ERR_ARG, "bad arguments");
  const int hp = ((P + 1) / 2 + 7) / 8 * 8;
  *HP = hp;
  if (!h_Ae && !h_Bo) return 0;
  SB_CHECK(h_Ae && h_Bo, SB200_ERR_ARG, "bad arguments");
  std::vector<double> Ae, Bo;
  cgl_even_odd_padded(P, hp, Ae, Bo);
  std::memcpy(h_Ae, Ae.data(), Ae.size() * sizeof(double));
  std::memcpy(h_Bo, Bo.data(), Bo.size() * sizeof(double));
  return 0;
}

// ---- elliptic ----------------------------------------------------------------------------
int sb200_elliptic_create(int d, const int* dim, sb200_elliptic** out) {
  SB_CHECK(out && dim, SB200_ERR_ARG, "null pointer");
  *out = nullptr;
  EllipticCtx* c = nullptr;
  SB_TRY(EllipticCtx::create(d, dim, 0, 1, &c));
  sb200_elliptic* e = new sb200_elliptic();
  e->c = c;
  *out = e;
  return 0;
}

int sb200_elliptic_create_slab(int d, const int* dim, int rank, int nranks, sb200_elliptic** out) {
  SB_CHECK(out && dim, SB200_ERR_ARG, "null pointer");
  *out = nullptr;
  EllipticCtx* c = nullptr;
  SB_TRY(EllipticCtx::create(d, dim, rank, nranks, &c));
  sb200_elliptic* e = new sb200_elliptic();
  e->c = c;
  *out = e;
  return 0;
}

int sb200_elliptic_slab_info(const sb200_elliptic* e, int* rank, int* nranks, int* i0, int* nloc, long long* goff,
                             long long* gtotal) {
  SB_CHECK(e, SB200_ERR_ARG, "null context");
  if (rank) *rank = e->c->arena.rank;
  if (nranks) *nranks = e->c->arena.nranks;
  if (i0) *i0 = e->c->gd.i0;
  if (nloc) *nloc = e->c->gd.dim[0];
  if (goff) *goff = e->c->gd.goff;
  if (gtotal) *gtotal = e->c->gtot;
  return 0;
}

int sb200_elliptic_slab_status(sb200_elliptic* e, long long* timeouts, void* stream) {
  SB_CHECK(e && timeouts, SB200_ERR_ARG, "null pointer");
  unsigned long long n = 0;
  SB_TRY(e->c->arena.timeouts((cudaStream_t)stream, &n));
  *timeouts = (long long)n;
  return 0;
}

int sb200_slab_geometry(int d, const int* dim, int rank, int nranks, int* i0, int* nloc, long long* goff, long long* g_local,
                        long long* m_local, long long* nd_local) {
  SB_CHECK(dim, SB200_ERR_ARG, "null pointer");
  SB_CHECK(nranks >= 1 && nranks <= SB200_MAX_RANKS && rank >= 0 && rank < nranks, SB200_ERR_USER,
           "slab partition: rank / nranks out of range (at most 8 ranks)");
  GridDesc gd;
  SB_TRY(gd.init_slab(d, dim, rank, nranks));
  if (i0) *i0 = gd.i0;
  if (nloc) *nloc = gd.dim[0];
  if (goff) *goff = gd.goff;
  if (g_local) *g_local = gd.g;
  if (m_local) *m_local = gd.m;
  if (nd_local) *nd_local = gd.m - gd.g;
  return 0;
}

int sb200_ipc_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

int sb200_elliptic_ipc_export(sb200_elliptic* e, void* handle) {
  SB_CHECK(e, SB200_ERR_ARG, "null context");
  return e->c->arena.export_handle(handle);
}

int sb200_elliptic_ipc_attach(sb200_elliptic* e, int peer_rank, const void* handle) {
  SB_CHECK(e, SB200_ERR_ARG, "null context");
  return e->c->arena.attach(peer_rank, handle);
}

int sb200_elliptic_attach_local(sb200_elliptic* e, int peer_rank, sb200_elliptic* peer) {
  SB_CHECK(e && peer, SB200_ERR_ARG, "null context");
  SB_CHECK(peer->c->arena.rank == peer_rank && peer->c->arena.nranks == e->c->arena.nranks, SB200_ERR_USER,
           "attach_local: peer context has a different rank / partition");
  return e->c->arena.attach_ptr(peer_rank, peer->c->arena.base);
}

int sb200_elliptic_sizes(const sb200_elliptic* e, long long* m, long long* g, long long* nd) {
  SB_CHECK(e, SB200_ERR_ARG, "null context");
  if (m) *m = e->c->gd.m;
  if (g) *g = e->c->gd.g;
  if (nd) *nd = e->c->gd.m - e->c->gd.g;
  return 0;
}

int sb200_elliptic_set_params(sb200_elliptic* e, double gamma, double exponent) {
  SB_CHECK(e, SB200_ERR_ARG, "null context");
  e->c->gamma = gamma;
  e->c->exponent = exponent;
  return 0;
}

int sb200_elliptic_set_dirichlet(sb200_elliptic* e, const double* d_values, void* stream) {
  SB_CHECK(e && d_values, SB200_ERR_ARG, "null pointer");
  const size_t bytes = (size_t)(e->c->gd.m - e->c->gd.g) * sizeof(double);
  SB_CUDA(cudaMemcpyAsync(e->c->dirichlet, d_values, bytes, cudaMemcpyDefault, (cudaStream_t)stream));
  return 0;
}

int sb200_elliptic_set_rhs(sb200_elliptic* e, const double* d_b, void* stream) {
  SB_CHECK(e && d_b, SB200_ERR_ARG, "null pointer");
  SB_CUDA(cudaMemcpyAsync(e->c->b, d_b, (size_t)e->c->gd.g * sizeof(double), cudaMemcpyDefault, (cudaStream_t)stream));
  return 0;
}

int sb200_elliptic_matmult(sb200_elliptic* e, const double* d_U, double* d_V, void* stream) {
  SB_CHECK(e, SB200_ERR_ARG, "null context");
  return e->c->matmult(d_U, d_V, (cudaStream_t)stream);
}

const char* sb200_elliptic_last_kernel(const sb200_elliptic* e) { return e ? e->c->last_kernel : "null context"; }

int sb200_elliptic_function(sb200_elliptic* e, const double* d_U, double* d_F, void* stream) {
  SB_CHECK(e, SB200_ERR_ARG, "null context");
  SB_CHECK(!e->c->arena.failed(), SB200_ERR_CUDA, "slab partition: a device-side flag wait timed out earlier; the context refuses further work");
  if (e->q_in && e->q_submitted > 0)  // queued applications still reading the old state finish first
    SB_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, e->q[(e->q_submitted - 1) % SB200_HOST_QUEUE_DEPTH].op_done, 0));
  SB_TRY(e->c->function(d_U, d_F, (cudaStream_t)stream));
  if (!e->state_ev) SB_CUDA(cudaEventCreateWithFlags(&e->state_ev, cudaEventDisableTiming));
  SB_CUDA(cudaEventRecord(e->state_ev, (cudaStream_t)stream));
  e->state_recorded = true;
  return 0;
}

static int elliptic_host_staging(sb200_elliptic* e) {
  const size_t bytes = (size_t)e->c->gd.g * sizeof(double);
  if (!e->d_in) SB_CUDA(cudaMalloc((void**)&e->d_in, bytes));
  if (!e->d_out) SB_CUDA(cudaMalloc((void**)&e->d_out, bytes));
  return 0;
}

int sb200_elliptic_matmult_host(sb200_elliptic* e, const double* h_U, double* h_V) {
  SB_CHECK(e && h_U && h_V, SB200_ERR_ARG, "null pointer");
  SB_TRY(elliptic_host_staging(e));
  const size_t bytes = (size_t)e->c->gd.g * sizeof(double);
  SB_CUDA(cudaMemcpyAsync(e->d_in, h_U, bytes, cudaMemcpyHostToDevice, 0));
  SB_TRY(e->c->matmult(e->d_in, e->d_out, 0));
  SB_CUDA(cudaMemcpyAsync(h_V, e->d_out, bytes, cudaMemcpyDeviceToHost, 0));
  SB_CUDA(cudaStreamSynchronize(0));
  return 0;
}

static int elliptic_host_queue(sb200_elliptic* e) {
  if (e->q_in) return 0;
  const size_t bytes = (size_t)e->c->gd.g * sizeof(double);
  // every piece is created only if missing, so a call that failed half way (out of memory) can be retried without leaking
  for (auto& s : e->q) {
    if (!s.d_in) SB_CUDA(cudaMalloc((void**)&s.d_in, bytes));
    if (!s.d_out) SB_CUDA(cudaMalloc((void**)&s.d_out, bytes));
    if (!s.in_done) SB_CUDA(cudaEventCreateWithFlags(&s.in_done, cudaEventDisableTiming));
    if (!s.op_done) SB_CUDA(cudaEventCreateWithFlags(&s.op_done, cudaEventDisableTiming));
    if (!s.out_done) SB_CUDA(cudaEventCreateWithFlags(&s.out_done, cudaEventDisableTiming));
  }
  if (!e->q_op) SB_CUDA(cudaStreamCreateWithFlags(&e->q_op, cudaStreamNonBlocking));
  if (!e->q_out) SB_CUDA(cudaStreamCreateWithFlags(&e->q_out, cudaStreamNonBlocking));
  SB_CUDA(cudaStreamCreateWithFlags(&e->q_in, cudaStreamNonBlocking));  // last: its presence marks the queue as built
  return 0;
}

int sb200_elliptic_matmult_host_submit(sb200_elliptic* e, const double* h_U, double* h_V) {
  SB_CHECK(e && h_U && h_V, SB200_ERR_ARG, "null pointer");
  SB_CHECK(h_U != h_V, SB200_ERR_ARG, "x must not alias y");
  SB_CHECK(e->q_submitted - e->q_waited < SB200_HOST_QUEUE_DEPTH, SB200_ERR_USER, "host queue full: call sb200_elliptic_matmult_host_wait first");
  SB_TRY(elliptic_host_queue(e));
  const size_t bytes = (size_t)e->c->gd.g * sizeof(double);
  auto& s = e->q[e->q_submitted % SB200_HOST_QUEUE_DEPTH];
  // every application runs behind the last FormFunction, whatever stream that was issued on and whether or not the queue is idle
  // (the non-blocking queue streams see no other stream's work by themselves)
  if (e->state_recorded) SB_CUDA(cudaStreamWaitEvent(e->q_op, e->state_ev, 0));
  if (e->q_submitted == e->q_waited) {
    // queue idle: also behind whatever else the caller enqueued on the legacy default stream (set_rhs / set_dirichlet copies)
    SB_CUDA(cudaEventRecord(s.op_done, 0));
    SB_CUDA(cudaStreamWaitEvent(e->q_op, s.op_done, 0));
  }
  // the slot's previous application was waited for (queue-full check above), so its vectors are free
  SB_CUDA(cudaMemcpyAsync(s.d_in, h_U, bytes, cudaMemcpyHostToDevice, e->q_in));
  SB_CUDA(cudaEventRecord(s.in_done, e->q_in));
  SB_CUDA(cudaStreamWaitEvent(e->q_op, s.in_done, 0));
  SB_TRY(e->c->matmult(s.d_in, s.d_out, e->q_op));  // the context's scratch is shared: applications serialise on q_op
  SB_CUDA(cudaEventRecord(s.op_done, e->q_op));
  SB_CUDA(cudaStreamWaitEvent(e->q_out, s.op_done, 0));
  SB_CUDA(cudaMemcpyAsync(h_V, s.d_out, bytes, cudaMemcpyDeviceToHost, e->q_out));
  SB_CUDA(cudaEventRecord(s.out_done, e->q_out));
  e->q_submitted++;
  return 0;
}

int sb200_elliptic_matmult_host_wait(sb200_elliptic* e) {
  SB_CHECK(e, SB200_ERR_ARG, "null pointer");
  SB_CHECK(e->q_waited < e->q_submitted, SB200_ERR_USER, "host queue empty: nothing was submitted");
  SB_CUDA(cudaEventSynchronize(e->q[e->q_waited % SB200_HOST_QUEUE_DEPTH].out_done));
  e->q_waited++;
  return 0;
}

int sb200_elliptic_matmult_host_pending(const sb200_elliptic* e, int* pending) {
  SB_CHECK(e && pending, SB200_ERR_ARG, "null pointer");
  *pending = (int)(e->q_submitted - e->q_waited);
  return 0;
}

int sb200_elliptic_function_host(sb200_elliptic* e, const double* h_U, double* h_F) {
  SB_CHECK(e && h_U && h_F, SB200_ERR_ARG, "null pointer");
  SB_TRY(elliptic_host_staging(e));
  const size_t bytes = (size_t)e->c->gd.g * sizeof(double);
  SB_CUDA(cudaMemcpyAsync(e->d_in, h_U, bytes, cudaMemcpyHostToDevice, 0));
  SB_TRY(e->c->function(e->d_in, e->d_out, 0));
  SB_CUDA(cudaMemcpyAsync(h_F, e->d_out, bytes, cudaMemcpyDeviceToHost, 0));
  SB_CUDA(cudaStreamSynchronize(0));
  return 0;
}

int sb200_elliptic_get_state(sb200_elliptic* e, int which, double* d_out, void* stream) {
  SB_CHECK(e && d_out, SB200_ERR_ARG, "null pointer");
  const double* src = nullptr;
  if (which == 0) src = e->c->eta;
  else if (which == 1) src = e->c->deta;
  else if (which >= 2 && which < 2 + e->c->gd.d) src = e->c->gradu[which - 2];
  SB_CHECK(src, SB200_ERR_USER, "state selector out of range");
  SB_CUDA(cudaMemcpyAsync(d_out, src, (size_t)e->c->gd.m * sizeof(double), cudaMemcpyDefault, (cudaStream_t)stream));
  return 0;
}

static int elliptic_fd(sb200_elliptic* e) {
  SB_CHECK(e->c->arena.nranks == 1, SB200_ERR_SUP, "the finite-difference preconditioning matrix is assembled for single-GPU contexts only");
  if (!e->fd) SB_TRY(FdAssembler::create(e->c->gd.d, e->c->gd.dim, 1, &e->fd));
  return 0;
}

int sb200_elliptic_jacobian_sizes(sb200_elliptic* e, long long* nrows, long long* nnz) {
  SB_CHECK(e, SB200_ERR_ARG, "null context");
  SB_TRY(elliptic_fd(e));
  if (nrows) *nrows = e->fd->nrows;
  if (nnz) *nnz = e->fd->nnz;
  return 0;
}

int sb200_elliptic_jacobian_csr(sb200_elliptic* e, int* d_rowptr, int* d_colidx, double* d_vals, void* stream) {
  SB_CHECK(e, SB200_ERR_ARG, "null context");
  SB_TRY(elliptic_fd(e));
  return e->fd->assemble(e->c->eta, e->c->deta, e->c->gradu, d_rowptr, d_colidx, d_vals, (cudaStream_t)stream);
}

int sb200_elliptic_pad(sb200_elliptic* e, const double* d_U, int with_dirichlet, double* d_local, void* stream) {
  SB_CHECK(e && d_U && d_local, SB200_ERR_ARG, "null pointer");
  return e->c->pad(d_U, with_dirichlet != 0, d_local, (cudaStream_t)stream);
}

int sb200_elliptic_crop(sb200_elliptic* e, const double* d_local, double* d_U, void* stream) {
  SB_CHECK(e && d_U && d_local, SB200_ERR_ARG, "null pointer");
  return e->c->crop(d_local, nullptr, d_U, (cudaStream_t)stream);
}

int sb200_elliptic_set_path(sb200_elliptic* e, int path) {
  SB_CHECK(e, SB200_ERR_ARG, "null context");
  SB_CHECK(path >= 0 && path <= 4, SB200_ERR_USER, "path must be 0..4");
  e->c->path = path;
  return 0;
}

int sb200_elliptic_debug_trace(sb200_elliptic* e, long long* d_buf) {
  SB_CHECK(e, SB200_ERR_ARG, "null context");
  e->c->trace = d_buf;
  return 0;
}

int sb200_elliptic_debug_timeline(sb200_elliptic* e, unsigned long long* h_out20, void* stream) {
  // slots (min,max) x 10 written by the slab kernels when SB200_XFLAGS has bit 64; read and re-armed here
  SB_CHECK(e && h_out20, SB200_ERR_ARG, "null pointer");
  SB_CHECK(e->c->sync, SB200_ERR_USER, "no persistent launch yet");
  unsigned long long* ts = reinterpret_cast<unsigned long long*>(e->c->sync) + 16;
  SB_CUDA(cudaMemcpyAsync(h_out20, ts, 20 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  SB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  unsigned long long init[20];
  for (int i = 0; i < 20; i++) init[i] = (i & 1) ? 0ull : ~0ull;
  SB_CUDA(cudaMemcpyAsync(ts, init, sizeof(init), cudaMemcpyHostToDevice, (cudaStream_t)stream));
  SB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}

int sb200_elliptic_destroy(sb200_elliptic* e) {
  if (!e) return 0;
  // applications still in flight use the context's scratch and write into the slots freed below: drain the queue first
  if (e->q_op) cudaStreamSynchronize(e->q_op);
  if (e->q_out) cudaStreamSynchronize(e->q_out);
  delete e->c;
  if (e->d_in) cudaFree(e->d_in);
  if (e->d_out) cudaFree(e->d_out);
  if (e->state_ev) cudaEventDestroy(e->state_ev);
  for (auto& s : e->q) {
    if (s.d_in) cudaFree(s.d_in);
    if (s.d_out) cudaFree(s.d_out);
    if (s.in_done) cudaEventDestroy(s.in_done);
    if (s.op_done) cudaEventDestroy(s.op_done);
    if (s.out_done) cudaEventDestroy(s.out_done);
  }
  if (e->q_in) cudaStreamDestroy(e->q_in);
  if (e->q_op) cudaStreamDestroy(e->q_op);
  if (e->q_out) cudaStreamDestroy(e->q_out);
  delete e->fd;
  delete e;
  return 0;
}

}  // extern "C"
