// ./stokes - the reference's executable (stokes.C:114-255, options :392-495) on the B200 path, in C++ with no Python anywhere:
//
//     apps/stokes -exact 2 -cont0 1 -schur_ksp_max_it 3 -vel_ksp_max_it 4 -svel_ksp_type preonly -ksp_type fgmres -dim 20,20,20 -ksp_rtol 1e-10
//     apps/stokes -exact 2 -cont 4 -rheology 1 -eps 1e-4 -exponent 3 -schur_ksp_max_it 3 -vel_ksp_max_it 4 -svel_ksp_type preonly -dim 20,20,20
//
// Same options, same order of work, same printed lines, -boundary 0 (all Dirichlet).  What runs where:
//   * StokesCreate / StokesCreateExactSolution / StokesFunction / the MatShells / StokesPCSetUp0: the reference's own names
//     (include/sb200_reference_api.h) over the C-ABI library - operators, residual, the finite-difference matrix on the GPU;
//   * every Krylov solve (outer FGMRES, stokes.C:157; the inner KSPs vel_ / schur_ / svel_, :328-341): the device-resident
//     sb200_ksp_*; the inner solves are PETSc's default left-preconditioned GMRES, run here as GMRES on M^-1 A;
//   * the saddle-point preconditioners StokesPCApply0..3 (stokes.C:1714-1817), the constant-pressure null space (:1006-1025),
//     SNES (full Newton steps, halved while the residual norm does not decrease) and the continuation loop (:214-236): host
//     orchestration in this file; their small vector updates run on host copies (the velocity PC below moves every vector
//     through the host anyway);
//   * the PC on MatVVPC is PETSc's own (ILU(0) by default, hypre in README:44) and out of scope: the HOST stand-in of
//     apps/common.h applies it (-vel_pc_type / -svel_pc_type ilu | jacobi | none, -vel_pc_factor_levels k).  jacobi and none are
//     applied on the device (the diagonal of MatVVPC uploaded once per set-up), so with them no vector of a linear solve ever
//     crosses PCIe: outer FGMRES, StokesPCApply, the three inner Krylov solves and the PC all work on device vectors.
// The Python command line (python -m spectral_petsc_b200.stokes) runs the identical flow; tests compare the two.
#include <functional>
#include <memory>

#include "common.h"

using app::HostPc;
using app::norm2;
using app::norm_inf;
using app::Options;
typedef std::vector<double> Vecd;
typedef std::function<int(const Vecd&, Vecd&)> HostOp;  // y = f(x) on host vectors; non-zero = error code

namespace {

// ---- device scratch vectors, reused by size ---------------------------------------------------------------------------------
struct Pool {
  std::map<PetscInt, std::vector<Vec>> free_;
  ~Pool() {
    for (auto& kv : free_)
      for (Vec v : kv.second) VecDestroy(v);
  }
  Vec get(PetscInt n) {
    auto& f = free_[n];
    if (!f.empty()) {
      Vec v = f.back();
      f.pop_back();
      return v;
    }
    Vec v = nullptr;
    return VecCreateSeqCUDA(PETSC_COMM_SELF, n, &v) ? nullptr : v;
  }
  void put(Vec v) { free_[v->n].push_back(v); }
  void clear() {  // before main returns: static destructors run after the CUDA runtime has begun to unload
    for (auto& kv : free_)
      for (Vec v : kv.second) VecDestroy(v);
    free_.clear();
  }
};
Pool g_pool;
struct Tmp {  // RAII lease of a pooled device vector
  Vec v;
  explicit Tmp(PetscInt n) : v(g_pool.get(n)) {}
  ~Tmp() {
    if (v) g_pool.put(v);
  }
};

int upload(const Vecd& h, Vec d) { return VecSetValuesHost(d, h.data()); }
int download(Vec d, Vecd& h) {
  h.resize(d->n);
  return VecGetValuesHost(d, h.data());
}

// MatMult(M, x, y) with host vectors on both sides
int mat(Mat M, const Vecd& x, Vecd& y) {
  PetscInt rows, cols;
  CHK(MatGetSize(M, &rows, &cols));
  Tmp dx(cols), dy(rows);
  if (!dx.v || !dy.v) return SB200_ERR_CUDA;
  CHK(upload(x, dx.v));
  CHK(MatMult(M, dx.v, dy.v));
  return download(dy.v, y);
}

// ---- the Krylov engine: device FGMRES(restart) with host-level callbacks ------------------------------------------------------
struct Callback {
  const HostOp* f;
  PetscInt n;
  Vecd hx, hy;
};
int trampoline(void* ctx, const double* d_x, double* d_y, void* stream) {
  Callback* c = (Callback*)ctx;
  const size_t bytes = (size_t)c->n * sizeof(double);
  c->hx.resize(c->n);
  int rc = sb200_memcpy_d2h(c->hx.data(), d_x, bytes, stream);
  if (!rc) rc = sb200_stream_sync(stream);
  if (!rc) rc = (*c->f)(c->hx, c->hy);
  if (!rc && (PetscInt)c->hy.size() != c->n) rc = SB200_ERR_USER;
  if (!rc) rc = sb200_memcpy_h2d(d_y, c->hy.data(), bytes, stream);
  return rc ? rc : sb200_stream_sync(stream);
}
int native_matmult(void* ctx, const double* d_x, double* d_y, void*) {  // MatMult(A, x, y) on raw device arrays
  Mat A = (Mat)ctx;
  PetscInt n;
  MatGetSize(A, &n, PETSC_NULL);
  Vec x = nullptr, y = nullptr;
  PetscErrorCode rc = VecCreateSeqCUDAWithArray(PETSC_COMM_SELF, n, (double*)d_x, &x);
  if (!rc) rc = VecCreateSeqCUDAWithArray(PETSC_COMM_SELF, n, d_y, &y);
  if (!rc) rc = MatMult(A, x, y);
  if (x) VecDestroy(x);
  if (y) VecDestroy(y);
  return rc;
}

struct Krylov {
  std::map<std::pair<PetscInt, int>, sb200_ksp*> cache;  // one solver object per (n, restart), like PETSc's one KSP per role
  ~Krylov() {
    for (auto& kv : cache) sb200_ksp_destroy(kv.second);
  }
  // x = solve(op, b) from a zero initial guess; op is a host-level callback, or the MatShell `native` applied on the device;
  // pc may be null (PCNONE)
  int solve(PetscInt n, const HostOp* op, Mat native, const HostOp* pc, const Vecd& b, Vecd& x, double rtol, int maxits, int restart, int* its, int* reason,
            sb200_apply_fn native_pc = nullptr, void* native_pc_ctx = nullptr) {
    auto key = std::make_pair(n, restart);
    if (!cache.count(key)) {
      sb200_ksp* k = nullptr;
      CHK(sb200_ksp_create(n, restart, &k));
      cache[key] = k;
    }
    sb200_ksp* k = cache[key];
    Callback cop{op, n, {}, {}}, cpc{pc, n, {}, {}};
    if (native_pc)  // the preconditioner works on the device vectors themselves (sb200_apply_saddle)
      CHK(sb200_ksp_set_operators(k, native ? native_matmult : trampoline, native ? (void*)native : (void*)&cop, native_pc, native_pc_ctx));
    else
      CHK(sb200_ksp_set_operators(k, native ? native_matmult : trampoline, native ? (void*)native : (void*)&cop, pc ? trampoline : nullptr, &cpc));
    CHK(sb200_ksp_set_tolerances(k, rtol, 1e-50, 1e5, maxits));
    Tmp db(n), dx(n);
    if (!db.v || !dx.v) return SB200_ERR_CUDA;
    CHK(upload(b, db.v));
    CHK(sb200_memset0(dx.v->d_array, (size_t)n * sizeof(double), nullptr));
    CHK(sb200_ksp_solve(k, db.v->d_array, dx.v->d_array, 0, nullptr));
    CHK(sb200_ksp_get_result(k, its, nullptr, nullptr, reason));
    return download(dx.v, x);
  }
};

// ---- everything the flow needs, in one place ----------------------------------------------------------------------------------
struct Flow {
  int d = 0;
  PetscInt m = 0, g = 0, gp = 0, gv = 0, dv = 0;
  StokesCtxB200* ctx = nullptr;
  Mat A = nullptr, MatVV = nullptr, MatPV = nullptr, MatVP = nullptr, MatSchur = nullptr;
  Krylov krylov;
  HostPc vel, svel;
  bool svel_is_vel = true, svel_preonly = false;
  int saddle = 0, vel_max_it = 10000, schur_max_it = 10000;
  double vel_rtol = 1e-5, schur_rtol = 1e-5, svel_rtol = 1e-5;
  int svel_max_it = 10000;
  Vecd diag;  // StokesMatGetDiagonalSchur: 1/eta at the pressure nodes
  // StokesPCApply{saddle} on the device (sb200_saddle_*, host/saddle.cpp): the default; -saddle_on_host 1 runs the same
  // composition in this file on host copies instead (kept as the cross-check of the device-resident one)
  sb200_saddle* dev_saddle = nullptr;
  bool saddle_on_host = false;
  ~Flow() {
    if (dev_saddle) sb200_saddle_destroy(dev_saddle);
  }
  static int pc_device(void* ctx, const double* d_x, double* d_y, void* stream) { return ((HostPc*)ctx)->apply_device(d_x, d_y, stream); }
  int setup_device_saddle() {
    if (saddle_on_host) return 0;
    if (!dev_saddle) CHK(sb200_saddle_create(StokesGetHandle(ctx), saddle, &dev_saddle));
    CHK(sb200_saddle_set_inner(dev_saddle, vel_rtol, vel_max_it, schur_rtol, schur_max_it, svel_preonly ? 1 : 0));
    CHK(sb200_saddle_set_svel(dev_saddle, svel_rtol, svel_max_it));
    sb200_apply_fn fv = vel.type == "none" ? nullptr : pc_device;
    HostPc& sp = svel_is_vel ? vel : svel;
    sb200_apply_fn fs = sp.type == "none" ? nullptr : pc_device;
    return sb200_saddle_set_velocity_pc(dev_saddle, fv, &vel, fs, &sp, 0);
  }

  // scatterGV / scatterGP and back (stokes.C:867-877): global AoS [v_0..v_{d-1}, p] per interior node
  void split(const Vecd& x, Vecd& v, Vecd& p) const {
    v.resize(gv);
    p.resize(gp);
    for (PetscInt q = 0; q < gp; q++) {
      for (int k = 0; k < d; k++) v[(size_t)q * d + k] = x[(size_t)q * (d + 1) + k];
      p[q] = x[(size_t)q * (d + 1) + d];
    }
  }
  void merge(const Vecd& v, const Vecd& p, Vecd& x) const {
    x.resize(g);
    for (PetscInt q = 0; q < gp; q++) {
      for (int k = 0; k < d; k++) x[(size_t)q * (d + 1) + k] = v[(size_t)q * d + k];
      x[(size_t)q * (d + 1) + d] = p[q];
    }
  }
  static void remove_mean(Vecd& p) {  // MatNullSpaceRemove with the constant vector
    double s = 0;
    for (double a : p) s += a;
    s /= (double)p.size();
    for (double& a : p) a -= s;
  }
  void remove_constant_pressure(Vecd& x) const {  // StokesRemoveConstantPressure (stokes.C:1006-1025)
    double s = 0;
    for (PetscInt q = 0; q < gp; q++) s += x[(size_t)q * (d + 1) + d];
    s /= (double)gp;
    for (PetscInt q = 0; q < gp; q++) x[(size_t)q * (d + 1) + d] -= s;
  }

  // KSPGMRES with left preconditioning (PETSc's default inner solver): GMRES on M^-1 A with the preconditioned right-hand
  // side, zero initial guess; `project` (KSPSetNullSpace) after every preconditioner application
  int left_gmres(PetscInt n, const HostOp& Aop, const HostOp& Minv, const Vecd& b, Vecd& x, double rtol, int maxits, bool project) {
    Vecd pb, t;
    CHK(Minv(b, pb));
    if (project) remove_mean(pb);
    HostOp op = [&](const Vecd& in, Vecd& out) {
      int rc = Aop(in, t);
      if (!rc) rc = Minv(t, out);
      if (!rc && project) remove_mean(out);
      return rc;
    };
    int its = 0, reason = 0;
    return krylov.solve(n, &op, nullptr, nullptr, pb, x, rtol, maxits, 30, &its, &reason);
  }

  HostOp pc_op(HostPc& pc) {
    return [&pc](const Vecd& r, Vecd& z) {
      z.resize(r.size());
      return pc.apply_host(r.data(), z.data());
    };
  }
  HostOp shell_op(Mat M) {
    return [M](const Vecd& x, Vecd& y) { return mat(M, x, y); };
  }

  int solve_velocity(const Vecd& rhs, Vecd& x) {  // KSPVelocity (stokes.C:334-337)
    return left_gmres(gv, shell_op(MatVV), pc_op(vel), rhs, x, vel_rtol, vel_max_it, false);
  }
  int solve_schur_velocity(const Vecd& rhs, Vecd& x) {  // KSPSchurVelocity (stokes.C:338-341); -svel_ksp_type preonly = one PC application
    HostPc& pc = svel_is_vel ? vel : svel;
    if (svel_preonly) return pc_op(pc)(rhs, x);
    return left_gmres(gv, shell_op(MatVV), pc_op(pc), rhs, x, svel_rtol, svel_max_it, false);
  }
  int schur(const Vecd& p, Vecd& y) {  // StokesMatMultSchur (stokes.C:523-535)
    Vecd v0, v1;
    CHK(mat(MatVP, p, v0));
    CHK(solve_schur_velocity(v0, v1));
    CHK(mat(MatPV, v1, y));
    for (double& a : y) a = -a;
    return 0;
  }
  int solve_schur(const Vecd& rhs, Vecd& x) {  // KSPSchur: Jacobi from StokesMatGetDiagonalSchur, constant null space (stokes.C:328-333)
    HostOp S = [this](const Vecd& p, Vecd& y) { return schur(p, y); };
    HostOp jac = [this](const Vecd& r, Vecd& z) {
      z.resize(r.size());
      for (size_t i = 0; i < r.size(); i++) z[i] = r[i] / diag[i];
      return 0;
    };
    return left_gmres(gp, S, jac, rhs, x, schur_rtol, schur_max_it, true);
  }

  int refresh_diag() {
    Tmp y(gp);
    if (!y.v) return SB200_ERR_CUDA;
    CHK(MatGetDiagonal(MatSchur, y.v));
    return download(y.v, diag);
  }

  // StokesPCApply0..3 (stokes.C:1714-1817)
  int saddle_apply(const Vecd& x, Vecd& y) {
    Vecd xv, xp, v1, p1, t, u;
    split(x, xv, xp);
    CHK(refresh_diag());
    if (saddle == 0) {  // full block LU
      CHK(solve_velocity(xv, v1));
      CHK(mat(MatPV, v1, t));
      for (PetscInt q = 0; q < gp; q++) t[q] = xp[q] - t[q];
      CHK(solve_schur(t, p1));
      CHK(mat(MatVP, p1, t));
      for (double& a : t) a = -a;
      CHK(solve_velocity(t, u));
      for (PetscInt q = 0; q < gv; q++) v1[q] += u[q];
    } else if (saddle == 1) {  // block upper triangular
      CHK(solve_schur(xp, p1));
      CHK(mat(MatVP, p1, t));
      for (PetscInt q = 0; q < gv; q++) t[q] = xv[q] - t[q];
      CHK(solve_velocity(t, v1));
    } else if (saddle == 2) {  // block diagonal
      CHK(solve_velocity(xv, v1));
      CHK(solve_schur(xp, p1));
    } else {  // block lower triangular
      CHK(solve_velocity(xv, v1));
      CHK(mat(MatPV, v1, t));
      for (PetscInt q = 0; q < gp; q++) t[q] = xp[q] - t[q];
      CHK(solve_schur(t, p1));
    }
    merge(v1, p1, y);
    return 0;
  }
};

}  // namespace

int main(int argc, char** argv) {
  Options o;
  if (int rc = o.parse(argc, argv)) return rc;
  // ---- StokesProcessOptions (stokes.C:392-495) ------------------------------------------------------------------------------
  StokesOptionsB200 opt;
  int dim[10] = {8, 6};
  int nd = o.int_array("dim", dim, 10);
  if (nd < 0) {
    fprintf(stderr, "error: -dim takes 1..10 comma-separated integers\n");
    return 83;
  }
  if (nd == 0) nd = 2;
  o.integer("debug", 0);
  opt.exact = o.integer("exact", 0);
  const int boundary = o.integer("boundary", 0);
  opt.rheology = o.integer("rheology", 0);
  opt.hardness = o.real("hardness", 1.0);
  const double exponent = o.real("exponent", 1.0), regularization = o.real("eps", 1.0);
  opt.exponent = exponent;
  opt.regularization = regularization;
  opt.gamma0 = o.real("gamma0", 1.0);
  const int cont0 = o.integer("cont0", 0), cont = o.integer("cont", 1);
  o.real("scaleM", 1.0);
  o.real("scaleN", 1.0);
  o.integer("zeroN", 0);
  o.real("zeroV", 1.0);
  if (nd != 2 && nd != 3) {
    fprintf(stderr, "error: the Stokes driver needs 2 or 3 dimensions (StokesPressureReduceOrder, stokes.C:1036)\n");
    return 83;
  }
  if (boundary != 0) {
    fprintf(stderr, "error: Boundary type %d not implemented (README:64-68: the Neumann / mixed conditions are broken upstream)\n", boundary);
    return 83;
  }
  if (opt.rheology != 0 && opt.rheology != 1) {
    fprintf(stderr, "error: Rheology type %d not implemented\n", (int)opt.rheology);  // stokes.C:492
    return 83;
  }
  if (opt.exact < 0 || opt.exact > 3) {
    fprintf(stderr, "error: Exact solution %d not implemented\n", (int)opt.exact);  // stokes.C:452
    return 83;
  }
  if (opt.exact == 3 && nd != 2) {
    fprintf(stderr, "error: StokesExact3 only implemented for dimension 2 but %d given\n", nd);  // stokes.C:2021
    return 83;
  }
  Flow F;
  const int pcvel = o.integer("pcvel", 0);
  F.saddle = o.integer("pc_saddle_type", 0);
  if (pcvel != 0) {
    fprintf(stderr, "error: pcvel type number %d not implemented (only the finite-difference matrix, StokesPCSetUp0)\n", pcvel);
    return 83;
  }
  if (F.saddle < 0 || F.saddle > 3) {
    fprintf(stderr, "error: pc_saddle_type %d not implemented\n", F.saddle);  // stokes.C:184
    return 83;
  }
  if (o.str("ksp_type", "fgmres") != "fgmres") {
    fprintf(stderr, "error: only -ksp_type fgmres (the type the reference sets in code, stokes.C:157) is built\n");
    return 83;
  }
  const double ksp_rtol = o.real("ksp_rtol", 1e-5), snes_rtol = o.real("snes_rtol", 1e-8), snes_atol = o.real("snes_atol", 1e-50);
  const int ksp_max_it = o.integer("ksp_max_it", 10000), snes_max_it = o.integer("snes_max_it", 50);
  F.vel_max_it = o.integer("vel_ksp_max_it", 10000);
  F.vel_rtol = o.real("vel_ksp_rtol", 1e-5);
  F.schur_max_it = o.integer("schur_ksp_max_it", 10000);
  F.schur_rtol = o.real("schur_ksp_rtol", 1e-5);
  F.svel_preonly = o.str("svel_ksp_type", "gmres") == "preonly";
  F.svel_rtol = o.real("svel_ksp_rtol", 1e-5);  // KSPSchurVelocity's own prefix (stokes.C:338-341)
  F.svel_max_it = o.integer("svel_ksp_max_it", 10000);
  F.saddle_on_host = o.integer("saddle_on_host", 0) != 0;
  F.vel.type = o.str("vel_pc_type", "ilu");  // PETSc's default PC for the SeqAIJ matrix MatVVPC is ILU(0); README:44 overrides it with hypre
  F.svel.type = o.str("svel_pc_type", "ilu");
  F.vel.levels = o.integer("vel_pc_factor_levels", 0);
  F.svel.levels = o.integer("svel_pc_factor_levels", 0);
  for (const HostPc* pc : {&F.vel, &F.svel})
    if (!HostPc::known(pc->type)) {
      fprintf(stderr, "error: PC type '%s' is not available in the native executable (have: ilu with -vel_pc_factor_levels k, jacobi, none; "
                      "the Python command line also offers lu)\n", pc->type.c_str());
      return 83;
    }
  F.svel_is_vel = F.svel.type == F.vel.type && F.svel.levels == F.vel.levels;
  const bool ksp_monitor = o.has("ksp_monitor"), snes_monitor = o.has("snes_monitor"), want_vtk = o.has("output_vtk");
  const std::string vtk_path = o.str("output_vtk", "stokes.vtk");

  const int d = nd;
  opt.numDims = d;
  for (int j = 0; j < 3; j++) opt.dim[j] = j < d ? dim[j] : 1;
  printf("Stokes problem  dim = [");
  for (int i = 0; i < d; i++) printf("%s%d", i ? "," : "", dim[i]);
  printf("]\n  hardness = %f    exponent = %8f    regularization = %8f    gamma0 = %8f\n", opt.hardness, exponent, regularization, opt.gamma0);

  // ---- objects (stokes.C:131-160) ---------------------------------------------------------------------------------------------
  Vec x, r, u, u2;
  SNES snes;
  CHK(StokesCreate(PETSC_COMM_SELF, &opt, &F.A, &x, &F.ctx));
  F.d = d;
  // not options of the reference: how this implementation evaluates StokesMatMult / StokesFunction (same operator either way)
  CHK(sb200_stokes_set_trace_divergence(StokesGetHandle(F.ctx), o.integer("sb200_trace_divergence", 1)));
  CHK(sb200_stokes_set_fold_pressure(StokesGetHandle(F.ctx), o.integer("sb200_fold_pressure", 1)));
  CHK(sb200_stokes_set_graph(StokesGetHandle(F.ctx), o.integer("sb200_graph", 0)));  // shells replayed from CUDA graphs (small grids)
  CHK(StokesGetSizes(F.ctx, &F.m, &F.g, &F.gp, &F.gv, &F.dv));
  CHK(StokesGetShells(F.ctx, &F.MatVV, &F.MatPV, &F.MatVP, &F.MatSchur));
  printf("DOF distribution: %d global   %d/%d pressure    %d/%d velocity    %d dirichlet    %d mixed\n", F.g, F.gp, F.m, F.gv, F.m * d, F.dv, 0);  // stokes.C:891
  CHK(VecDuplicate(x, &r));
  CHK(VecDuplicate(x, &u));
  CHK(VecDuplicate(x, &u2));
  CHK(SNESCreate(PETSC_COMM_SELF, &snes));
  CHK(SNESSetApplicationContext(snes, F.ctx));
  CHK(StokesCreateExactSolution(snes, u, u2));
  Vecd hU, hU2;
  CHK(download(u, hU));
  CHK(download(u2, hU2));

  // the residual with the line every evaluation prints (stokes.C:731-734)
  auto function = [&](const Vecd& hx, Vecd& hF) -> int {
    CHK(upload(hx, x));
    CHK(StokesFunction(snes, x, r, F.ctx));
    CHK(download(r, hF));
    PetscReal mn, mx;
    CHK(StokesGetEtaMinMax(F.ctx, &mn, &mx));
    printf("Minimum eta = %9.3e   Maximum eta = %9.3e\n", mn, mx);
    return 0;
  };

  Vecd hF;
  CHK(function(hU, hF));  // stokes.C:179
  printf("Norm of solution %9.3e  norm of forcing %9.3e  norm of residual %9.3e\n", norm_inf(hU), norm_inf(hU2), norm_inf(hF));
  {  // MatNullSpaceTest(ns, A) (stokes.C:206-212) with the normalised constant-pressure vector (:1013-1020)
    Vecd ns(F.g, 0.0), y;
    for (PetscInt q = 0; q < F.gp; q++) ns[(size_t)q * (d + 1) + d] = 1.0 / sqrt((double)F.gp);
    CHK(mat(F.A, ns, y));
    if (!(norm_inf(y) < 1e-8)) {
      fprintf(stderr, "error: Null space test failed\n");
      return 1;
    }
  }

  // ---- the solve loop (stokes.C:214-236) ------------------------------------------------------------------------------------------
  Vecd hx(F.g, 0.0);  // VecSet(x, 0.0)
  for (int i = cont0; i < cont + 1; i++) {
    const double e_i = 1.0 + pow(1.0 * i / cont, 0.8) * (exponent - 1.0), r_i = exp(log(regularization) * i / cont);
    CHK(StokesSetContinuation(F.ctx, e_i, r_i));
    printf("## [%d/%d] Solving with exponent = %5f regularization %8.2e\n", i, cont, e_i, r_i);
    // SNESSolve: x <- x + lam J^-1 (-F(x))
    CHK(function(hx, hF));
    double fn = norm2(hF);
    const double f0 = fn;
    Vecd hist(1, fn), rhs, dx, hxn, hFn;
    std::vector<int> kits;
    int its = 0;
    while (its < snes_max_it && fn > fmax(snes_rtol * f0, snes_atol)) {
      {  // StokesPCSetUp0 + the PCs PETSc builds on MatVVPC for KSPVelocity / KSPSchurVelocity
        PC pc;
        Mat MatVVPC;
        CHK(PCCreate(PETSC_COMM_SELF, &pc));
        CHK(PCShellSetContext(pc, F.ctx));
        CHK(StokesPCSetUp0(pc));
        CHK(StokesGetPCMatrix(F.ctx, &MatVVPC));
        CHK(PCDestroy(pc));
        F.vel.n = F.gv;
        CHK(F.vel.setup(MatVVPC));
        if (!F.svel_is_vel) {
          F.svel.n = F.gv;
          CHK(F.svel.setup(MatVVPC));
        }
      }
      rhs.resize(F.g);
      for (PetscInt q = 0; q < F.g; q++) rhs[q] = -hF[q];
      HostOp pc_outer = [&](const Vecd& in, Vecd& out) {  // the PCShell apply + KSPSetNullSpace (stokes.C:155-160, 1022)
        int rc = F.saddle_apply(in, out);
        if (!rc) F.remove_constant_pressure(out);
        return rc;
      };
      int k = 0, kreason = 0;
      CHK(F.setup_device_saddle());
      if (F.saddle_on_host)
        CHK(F.krylov.solve(F.g, nullptr, F.A, &pc_outer, rhs, dx, ksp_rtol, ksp_max_it, 30, &k, &kreason));
      else
        CHK(F.krylov.solve(F.g, nullptr, F.A, nullptr, rhs, dx, ksp_rtol, ksp_max_it, 30, &k, &kreason, sb200_apply_saddle, F.dev_saddle));
      kits.push_back(k);
      if (ksp_monitor) printf("    KSP iterations %d reason %d\n", k, kreason);
      double lam = 1.0, fnn = 0;
      while (true) {
        hxn.resize(F.g);
        for (PetscInt q = 0; q < F.g; q++) hxn[q] = hx[q] + lam * dx[q];
        CHK(function(hxn, hFn));
        fnn = norm2(hFn);
        if (fnn < fn || lam < 1e-3) break;
        lam *= 0.5;
      }
      hx.swap(hxn);
      hF.swap(hFn);
      fn = fnn;
      hist.push_back(fn);
      its++;
    }
    if (snes_monitor)
      for (size_t q = 0; q < hist.size(); q++) printf("  %d SNES Function norm %.12e\n", (int)q, hist[q]);
    const char* reason = hist.back() <= snes_atol ? "CONVERGED_FNORM_ABS" : (hist.back() <= snes_rtol * f0 ? "CONVERGED_FNORM_RELATIVE" : "DIVERGED_MAX_IT");
    Vecd err(F.g);
    for (PetscInt q = 0; q < F.g; q++) err[q] = hx[q] - hU[q];  // VecAXPY(r, -1, u); MatNullSpaceRemove (stokes.C:224-226)
    F.remove_constant_pressure(err);
    printf("Number of nonlinear iterations = %d\n", its);
    printf("Reason for solver termination: %s\n", reason);
    printf("%-25s: abs = %8e\n", "Norm of error", norm_inf(err));
    printf("KSP iterations per Newton step:");
    for (int kk : kits) printf(" %d", kk);
    printf("\n");
  }

  if (want_vtk) {  // StokesStateView(ctx, x, "final state") (stokes.C:238-242, 1821-1894)
    CHK(upload(hx, x));
    if (StokesStateViewFile(F.ctx, x, vtk_path.c_str())) {
      fprintf(stderr, "error: cannot write %s\n", vtk_path.c_str());
      return 1;
    }
  }
  o.warn_unused();

  g_pool.clear();
  CHK(SNESDestroy(snes));
  CHK(StokesDestroy(F.ctx));
  CHK(MatDestroy(F.A));
  for (Vec v : {x, r, u, u2}) CHK(VecDestroy(v));
  return 0;
}
