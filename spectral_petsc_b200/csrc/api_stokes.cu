// C ABI for the stokes.C shells (include/spectral_b200.h).
#include <string>

#include "../../include/spectral_b200.h"
#include "common.cuh"
#include "fd_assembly.h"
#include "stokes.h"

using namespace sb200;

struct sb200_stokes {
  StokesCtx* c = nullptr;
  double* d_in = nullptr;  // staging for *_host
  double* d_out = nullptr;
  FdAssembler* fd = nullptr;  // StokesPCSetUp0's MatVVPC (built on first use)
  // opt-in (sb200_stokes_set_graph): one CUDA graph per linear shell, captured on fixed staging vectors and replayed - for the
  // small grids (BASELINE config 4, 20^3) where a shell is 5-22 launches of a few us each
  struct GraphSlot {
    cudaGraphExec_t exec = nullptr;
    double* in = nullptr;
    double* out = nullptr;
    int nodes = 0;
  };
  GraphSlot graph[4];  // 0 StokesMatMult, 1 VV, 2 PV, 3 VP
  cudaStream_t gstream = nullptr;
  bool use_graph = false;
};

namespace {

void drop_graphs(sb200_stokes* s) {  // the captured launch sequence depends on the evaluation switches
  bool any = false;
  for (auto& g : s->graph) any = any || g.exec;
  if (any) cudaDeviceSynchronize();  // a replay may still be in flight
  for (auto& g : s->graph) {
    if (g.exec) cudaGraphExecDestroy(g.exec);
    g.exec = nullptr;
  }
}

// y = shell(x) through the slot's graph; `run` enqueues the shell's launches on a stream
template <class Run>
int run_graphed(sb200_stokes* s, int slot, long long n_in, long long n_out, const double* d_x, double* d_y, cudaStream_t stream, Run run) {
  sb200_stokes::GraphSlot& g = s->graph[slot];
  if (!g.exec) {
    if (!g.in) SB_CUDA(cudaMalloc((void**)&g.in, (size_t)n_in * sizeof(double)));
    if (!g.out) SB_CUDA(cudaMalloc((void**)&g.out, (size_t)n_out * sizeof(double)));
    if (!s->gstream) SB_CUDA(cudaStreamCreateWithFlags(&s->gstream, cudaStreamNonBlocking));
    SB_CUDA(cudaStreamSynchronize(stream));  // the warm-up below uses the context's scratch fields on another stream
    SB_CUDA(cudaMemsetAsync(g.in, 0, (size_t)n_in * sizeof(double), s->gstream));
    SB_TRY(run(g.in, g.out, s->gstream));  // once outside the capture: function attributes, lazy module loading
    SB_CUDA(cudaStreamSynchronize(s->gstream));
    SB_CUDA(cudaStreamBeginCapture(s->gstream, cudaStreamCaptureModeThreadLocal));
    const int rc = run(g.in, g.out, s->gstream);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(s->gstream, &graph);
    if (rc || ce != cudaSuccess || !graph) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      SB_CHECK(rc == 0, rc, "Stokes shell: the launch sequence failed while being captured");
      set_last_error(std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce));
      return SB200_ERR_CUDA;
    }
    size_t nn = 0;
    cudaGraphGetNodes(graph, nullptr, &nn);
    g.nodes = (int)nn;
    const cudaError_t ie = cudaGraphInstantiate(&g.exec, graph, 0);
    cudaGraphDestroy(graph);
    SB_CUDA(ie);
  }
  SB_CUDA(cudaMemcpyAsync(g.in, d_x, (size_t)n_in * sizeof(double), cudaMemcpyDeviceToDevice, stream));
  SB_CUDA(cudaGraphLaunch(g.exec, stream));
  SB_CUDA(cudaMemcpyAsync(d_y, g.out, (size_t)n_out * sizeof(double), cudaMemcpyDeviceToDevice, stream));
  count_launch(g.nodes);
  return 0;
}

bool graphed(const sb200_stokes* s) { return s->use_graph && s->c->arena.nranks == 1; }

}  // namespace

extern "C" {

int sb200_stokes_create(int d, const int* dim, sb200_stokes** out) {
  SB_CHECK(out && dim, SB200_ERR_ARG, "null pointer");
  *out = nullptr;
  StokesCtx* c = nullptr;
  SB_TRY(StokesCtx::create(d, dim, 0, 1, &c));
  sb200_stokes* s = new sb200_stokes();
  s->c = c;
  *out = s;
  return 0;
}

int sb200_stokes_create_slab(int d, const int* dim, int rank, int nranks, sb200_stokes** out) {
  SB_CHECK(out && dim, SB200_ERR_ARG, "null pointer");
  *out = nullptr;
  StokesCtx* c = nullptr;
  SB_TRY(StokesCtx::create(d, dim, rank, nranks, &c));
  sb200_stokes* s = new sb200_stokes();
  s->c = c;
  *out = s;
  return 0;
}

int sb200_stokes_slab_info(const sb200_stokes* s, int* rank, int* nranks, int* i0, int* nloc, long long* goff_nodes) {
  SB_CHECK(s, SB200_ERR_ARG, "null context");
  if (rank) *rank = s->c->arena.rank;
  if (nranks) *nranks = s->c->arena.nranks;
  if (i0) *i0 = s->c->gd.i0;
  if (nloc) *nloc = s->c->gd.dim[0];
  if (goff_nodes) *goff_nodes = s->c->gd.goff;
  return 0;
}

int sb200_stokes_ipc_export(sb200_stokes* s, void* handle) {
  SB_CHECK(s, SB200_ERR_ARG, "null context");
  return s->c->arena.export_handle(handle);
}

int sb200_stokes_ipc_attach(sb200_stokes* s, int peer_rank, const void* handle) {
  SB_CHECK(s, SB200_ERR_ARG, "null context");
  return s->c->arena.attach(peer_rank, handle);
}

int sb200_stokes_attach_local(sb200_stokes* s, int peer_rank, sb200_stokes* peer) {
  SB_CHECK(s && peer, SB200_ERR_ARG, "null context");
  SB_CHECK(peer->c->arena.rank == peer_rank && peer->c->arena.nranks == s->c->arena.nranks, SB200_ERR_USER,
           "attach_local: peer context has a different rank / partition");
  return s->c->arena.attach_ptr(peer_rank, peer->c->arena.base);
}

int sb200_stokes_slab_status(sb200_stokes* s, long long* timeouts, void* stream) {
  SB_CHECK(s && timeouts, SB200_ERR_ARG, "null pointer");
  unsigned long long n = 0;
  SB_TRY(s->c->arena.timeouts((cudaStream_t)stream, &n));
  *timeouts = (long long)n;
  return 0;
}

int sb200_stokes_sizes(const sb200_stokes* s, long long* m, long long* g, long long* gp, long long* gv, long long* dv) {
  SB_CHECK(s, SB200_ERR_ARG, "null context");
  if (m) *m = s->c->gd.m;
  if (g) *g = s->c->g;
  if (gp) *gp = s->c->gp;
  if (gv) *gv = s->c->gv;
  if (dv) *dv = s->c->dvn;
  return 0;
}

int sb200_stokes_set_rheology(sb200_stokes* s, int type, double hardness, double exponent, double regularization, double gamma0) {
  SB_CHECK(s, SB200_ERR_ARG, "null context");
  if (type != 0 && type != 1) {  // stokes.C:491
    set_last_error("Rheology type " + std::to_string(type) + " not implemented");
    return SB200_ERR_SUP;
  }
  s->c->rheology = type;
  s->c->hardness = hardness;
  s->c->exponent = exponent;
  s->c->regularization = regularization;
  s->c->gamma0 = gamma0;
  return 0;
}

int sb200_stokes_set_dirichlet(sb200_stokes* s, const double* d_values, void* stream) {
  SB_CHECK(s && d_values, SB200_ERR_ARG, "null pointer");
  SB_CUDA(cudaMemcpyAsync(s->c->dirichlet, d_values, (size_t)s->c->dvn * sizeof(double), cudaMemcpyDefault, (cudaStream_t)stream));
  return 0;
}

int sb200_stokes_set_force(sb200_stokes* s, const double* d_force, void* stream) {
  SB_CHECK(s && d_force, SB200_ERR_ARG, "null pointer");
  SB_CUDA(cudaMemcpyAsync(s->c->force, d_force, (size_t)s->c->g * sizeof(double), cudaMemcpyDefault, (cudaStream_t)stream));
  return 0;
}

int sb200_stokes_matmult(sb200_stokes* s, const double* d_x, double* d_y, void* stream) {
  SB_CHECK(s, SB200_ERR_ARG, "null context");
  SB_CHECK(!s->c->arena.failed(), SB200_ERR_CUDA, "slab partition: a device-side flag wait timed out earlier (a peer never arrived or the ranks' calls went out of step); the context refuses further work");
  if (graphed(s)) {
    SB_CHECK(d_x && d_y && d_x != d_y, SB200_ERR_ARG, "StokesMatMult: x and y must be distinct non-null vectors");
    StokesCtx* c = s->c;
    return run_graphed(s, 0, c->g, c->g, d_x, d_y, (cudaStream_t)stream, [c](const double* x, double* y, cudaStream_t st) { return c->matmult(x, y, st); });
  }
  return s->c->matmult(d_x, d_y, (cudaStream_t)stream);
}

int sb200_stokes_matmult_vv(sb200_stokes* s, const double* d_x, double* d_y, void* stream) {
  SB_CHECK(s && d_x && d_y && d_x != d_y, SB200_ERR_ARG, "StokesMatMultVV: bad vectors");
  SB_CHECK(!s->c->arena.failed(), SB200_ERR_CUDA, "slab partition: a device-side flag wait timed out earlier (a peer never arrived or the ranks' calls went out of step); the context refuses further work");
  if (graphed(s)) {
    StokesCtx* c = s->c;
    return run_graphed(s, 1, c->gv, c->gv, d_x, d_y, (cudaStream_t)stream,
                       [c](const double* x, double* y, cudaStream_t st) { return c->matmult_vv_into(x, c->gd.d, 0, y, c->gd.d, 0, st); });
  }
  return s->c->matmult_vv_into(d_x, s->c->gd.d, 0, d_y, s->c->gd.d, 0, (cudaStream_t)stream);
}

int sb200_stokes_matmult_pv(sb200_stokes* s, const double* d_x, double* d_y, void* stream) {
  SB_CHECK(s && d_x && d_y && d_x != d_y, SB200_ERR_ARG, "StokesMatMultPV: bad vectors");
  SB_CHECK(!s->c->arena.failed(), SB200_ERR_CUDA, "slab partition: a device-side flag wait timed out earlier (a peer never arrived or the ranks' calls went out of step); the context refuses further work");
  if (graphed(s)) {
    StokesCtx* c = s->c;
    return run_graphed(s, 2, c->gv, c->gp, d_x, d_y, (cudaStream_t)stream,
                       [c](const double* x, double* y, cudaStream_t st) { return c->divergence_into(x, c->gd.d, 0, false, y, 1, 0, st); });
  }
  return s->c->divergence_into(d_x, s->c->gd.d, 0, false, d_y, 1, 0, (cudaStream_t)stream);
}

int sb200_stokes_matmult_vp(sb200_stokes* s, const double* d_x, double* d_y, void* stream) {
  SB_CHECK(s && d_x && d_y && d_x != d_y, SB200_ERR_ARG, "StokesMatMultVP: bad vectors");
  SB_CHECK(!s->c->arena.failed(), SB200_ERR_CUDA, "slab partition: a device-side flag wait timed out earlier (a peer never arrived or the ranks' calls went out of step); the context refuses further work");
  if (graphed(s)) {
    StokesCtx* c = s->c;
    return run_graphed(s, 3, c->gp, c->gv, d_x, d_y, (cudaStream_t)stream,
                       [c](const double* x, double* y, cudaStream_t st) { return c->matmult_vp_into(x, 1, 0, y, c->gd.d, 0, false, nullptr, st); });
  }
  return s->c->matmult_vp_into(d_x, 1, 0, d_y, s->c->gd.d, 0, false, nullptr, (cudaStream_t)stream);
}

int sb200_stokes_get_diagonal_schur(sb200_stokes* s, double* d_y, void* stream) {
  SB_CHECK(s && d_y, SB200_ERR_ARG, "null pointer");
  return s->c->get_diagonal_schur(d_y, (cudaStream_t)stream);
}

int sb200_stokes_divergence(sb200_stokes* s, int with_dirichlet, const double* d_x, double* d_y, void* stream) {
  SB_CHECK(s && d_x && d_y, SB200_ERR_ARG, "null pointer");
  SB_CHECK(!s->c->arena.failed(), SB200_ERR_CUDA, "slab partition: a device-side flag wait timed out earlier (a peer never arrived or the ranks' calls went out of step); the context refuses further work");
  return s->c->divergence_into(d_x, s->c->gd.d, 0, with_dirichlet != 0, d_y, 1, 0, (cudaStream_t)stream);
}

int sb200_stokes_set_trace_divergence(sb200_stokes* s, int on) {
  SB_CHECK(s, SB200_ERR_ARG, "null context");
  s->c->trace_divergence = on != 0;
  drop_graphs(s);
  return 0;
}

int sb200_stokes_set_fold_pressure(sb200_stokes* s, int on) {
  SB_CHECK(s, SB200_ERR_ARG, "null context");
  s->c->fold_pressure = on != 0;
  drop_graphs(s);
  return 0;
}

int sb200_stokes_set_graph(sb200_stokes* s, int on) {
  SB_CHECK(s, SB200_ERR_ARG, "null context");
  s->use_graph = on != 0;
  return 0;
}

int sb200_stokes_matmult_schur(sb200_stokes* s, const double* d_x, double* d_y, sb200_velocity_solve_fn solve, void* solve_ctx, void* stream) {
  SB_CHECK(s && d_x && d_y, SB200_ERR_ARG, "null pointer");
  SB_CHECK(!s->c->arena.failed(), SB200_ERR_CUDA, "slab partition: a device-side flag wait timed out earlier (a peer never arrived or the ranks' calls went out of step); the context refuses further work");
  return s->c->matmult_schur(d_x, d_y, solve, solve_ctx, (cudaStream_t)stream);
}

int sb200_stokes_function(sb200_stokes* s, const double* d_x, double* d_y, void* stream) {
  SB_CHECK(s, SB200_ERR_ARG, "null context");
  SB_CHECK(!s->c->arena.failed(), SB200_ERR_CUDA, "slab partition: a device-side flag wait timed out earlier (a peer never arrived or the ranks' calls went out of step); the context refuses further work");
  return s->c->function(d_x, d_y, (cudaStream_t)stream);
}

static int stokes_host_staging(sb200_stokes* s) {
  if (s->d_in) return 0;
  SB_CUDA(cudaMalloc((void**)&s->d_in, (size_t)s->c->g * sizeof(double)));
  SB_CUDA(cudaMalloc((void**)&s->d_out, (size_t)s->c->g * sizeof(double)));
  return 0;
}

int sb200_stokes_matmult_host(sb200_stokes* s, const double* h_x, double* h_y) {
  SB_CHECK(s && h_x && h_y, SB200_ERR_ARG, "null pointer");
  SB_TRY(stokes_host_staging(s));
  const size_t bytes = (size_t)s->c->g * sizeof(double);
  SB_CUDA(cudaMemcpyAsync(s->d_in, h_x, bytes, cudaMemcpyHostToDevice, 0));
  SB_TRY(s->c->matmult(s->d_in, s->d_out, 0));
  SB_CUDA(cudaMemcpyAsync(h_y, s->d_out, bytes, cudaMemcpyDeviceToHost, 0));
  SB_CUDA(cudaStreamSynchronize(0));
  return 0;
}

int sb200_stokes_function_host(sb200_stokes* s, const double* h_x, double* h_y) {
  SB_CHECK(s && h_x && h_y, SB200_ERR_ARG, "null pointer");
  SB_TRY(stokes_host_staging(s));
  const size_t bytes = (size_t)s->c->g * sizeof(double);
  SB_CUDA(cudaMemcpyAsync(s->d_in, h_x, bytes, cudaMemcpyHostToDevice, 0));
  SB_TRY(s->c->function(s->d_in, s->d_out, 0));
  SB_CUDA(cudaMemcpyAsync(h_y, s->d_out, bytes, cudaMemcpyDeviceToHost, 0));
  SB_CUDA(cudaStreamSynchronize(0));
  return 0;
}

int sb200_stokes_eta_minmax(sb200_stokes* s, double* h_min, double* h_max, void* stream) {
  SB_CHECK(s && h_min && h_max, SB200_ERR_ARG, "null pointer");
  double mm[2];
  SB_CUDA(cudaMemcpyAsync(mm, s->c->minmax, sizeof(mm), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  SB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  *h_min = mm[0];
  *h_max = mm[1];
  return 0;
}

int sb200_stokes_get_state(sb200_stokes* s, int which, double* d_out, void* stream) {
  SB_CHECK(s && d_out, SB200_ERR_ARG, "null pointer");
  const double* src = nullptr;
  size_t n = (size_t)s->c->gd.m;
  if (which == 0) src = s->c->eta;
  else if (which == 1) src = s->c->deta;
  else if (which >= 2 && which < 2 + s->c->gd.d) {
    src = s->c->strain[which - 2];
    n *= s->c->gd.d;
  }
  SB_CHECK(src, SB200_ERR_USER, "state selector out of range");
  SB_CUDA(cudaMemcpyAsync(d_out, src, n * sizeof(double), cudaMemcpyDefault, (cudaStream_t)stream));
  return 0;
}

int sb200_stokes_pressure_reduce_order(sb200_stokes* s, double* d_pL, void* stream) {
  SB_CHECK(s && d_pL, SB200_ERR_ARG, "null pointer");
  return s->c->pressure_reduce_order(d_pL, (cudaStream_t)stream);
}

static int stokes_fd(sb200_stokes* s) {
  SB_CHECK(s->c->arena.nranks == 1, SB200_ERR_SUP, "the finite-difference preconditioning matrix is assembled for single-GPU contexts only");
  if (!s->fd) SB_TRY(FdAssembler::create(s->c->gd.d, s->c->gd.dim, s->c->gd.d, &s->fd));
  return 0;
}

int sb200_stokes_pc_velocity_sizes(sb200_stokes* s, long long* nrows, long long* nnz) {
  SB_CHECK(s, SB200_ERR_ARG, "null context");
  SB_TRY(stokes_fd(s));
  if (nrows) *nrows = s->fd->nrows;
  if (nnz) *nnz = s->fd->nnz;
  return 0;
}

int sb200_stokes_pc_velocity_csr(sb200_stokes* s, int* d_rowptr, int* d_colidx, double* d_vals, void* stream) {
  SB_CHECK(s, SB200_ERR_ARG, "null context");
  SB_TRY(stokes_fd(s));
  return s->fd->assemble(s->c->eta, nullptr, nullptr, d_rowptr, d_colidx, d_vals, (cudaStream_t)stream);
}

int sb200_stokes_destroy(sb200_stokes* s) {
  if (!s) return 0;
  drop_graphs(s);
  for (auto& g : s->graph) {
    if (g.in) cudaFree(g.in);
    if (g.out) cudaFree(g.out);
  }
  if (s->gstream) cudaStreamDestroy(s->gstream);
  delete s->fd;
  delete s->c;
  if (s->d_in) cudaFree(s->d_in);
  if (s->d_out) cudaFree(s->d_out);
  delete s;
  return 0;
}

}  // extern "C"
