/* spectral_b200.h - C ABI of the B200-native matrix-free Chebyshev collocation operator.
 *
 * This is the drop-in boundary for ONE path of jedbrown/spectral-petsc: the per-axis
 * Chebyshev-Gauss-Lobatto derivative (chebyshev.c) and the MatShell / SNES callbacks built on
 * it (elliptic.C, stokes.C).  Each entry point cites the reference interface it replaces.
 *
 * Conventions (mirroring the reference's PETSc conventions):
 *   - every function returns an int error code, 0 = success (PetscErrorCode convention,
 *     chebyshev.c:98 SETERRQ / CHKERRQ); sb200_last_error() gives the message.
 *   - contexts are opaque and NOT re-entrant (shared scratch, like ChebCtx.work / MatElliptic.w).
 *   - all arithmetic is IEEE fp64; vectors are contiguous doubles with the reference's Vec layout.
 *   - pointers named d_* are DEVICE pointers (the "VECCUDA-backed Vec"); h_* are HOST pointers.
 *     x must not alias y; x is preserved, y fully overwritten (chebyshev.c:127 PRESERVE_INPUT).
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Device entry points
 *     are asynchronous on that stream; *_host entry points copy in, run, copy out and synchronise.
 *   - there is no CPU fallback: without a CUDA device every compute call fails with SB200_ERR_CUDA.
 */
#ifndef SPECTRAL_B200_H
#define SPECTRAL_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB200_OK 0
#define SB200_ERR_USER 83     /* PETSC_ERR_USER: bad sizes / options (chebyshev.c:98,106,122) */
#define SB200_ERR_SUP 56      /* PETSC_ERR_SUP: not implemented (stokes.C:452, elliptic.C:406) */
#define SB200_ERR_CUDA 97     /* CUDA runtime failure or no device */
#define SB200_ERR_ARG 62      /* PETSC_ERR_ARG_*: null / aliased pointers */

typedef struct sb200_cheb sb200_cheb;         /* replaces ChebCtx      (chebyshev.h:18-24) */
typedef struct sb200_elliptic sb200_elliptic; /* replaces MatElliptic  (elliptic.C:78-86)  */
typedef struct sb200_stokes sb200_stokes;     /* replaces StokesCtx    (stokes.C:40-65)    */

/* ---- library ------------------------------------------------------------------------------ */
int sb200_version(void);
const char* sb200_last_error(void);
/* Number of kernels this library has launched in the calling process (bench.py "gpu_launches"). */
long long sb200_launch_count(void);
int sb200_device_count(int* n);
int sb200_set_device(int ordinal);
/* Device memory helpers so a non-torch host (the C++ PETSc shim) can own VECCUDA-like arrays. */
int sb200_malloc(void** d_ptr, size_t bytes);
int sb200_free(void* d_ptr);
int sb200_memcpy_h2d(void* d_dst, const void* h_src, size_t bytes, void* stream);
int sb200_memcpy_d2h(void* h_dst, const void* d_src, size_t bytes, void* stream);
int sb200_memcpy_d2d(void* d_dst, const void* d_src, size_t bytes, void* stream);
int sb200_memset0(void* d_dst, size_t bytes, void* stream);
int sb200_stream_sync(void* stream);
/* Measurement helper (bench.py roofline denominator): runs a register-resident mma.sync.m8n8k4.f64 (SASS DMMA)
 * loop on every SM of the current device for about target_ms milliseconds (clamped to 1..500) and returns the FP64
 * tensor-pipe rate in TFLOP/s; measured_ms (may be NULL) receives the timed kernel's duration. */
int sb200_fp64_dmma_peak(double target_ms, double* tflops, double* measured_ms);

/* ---- Chebyshev derivative: MatCreateCheb / ChebMult / ChebDestroy (chebyshev.h:31-34) ----- */
/* MatCreateCheb(comm, rank, tr, dims, flag, vx, vy, &A) (chebyshev.c:89-138).  dims is row-major,
 * last axis fastest; n_total is the Vec length and must equal prod(dims) (chebyshev.c:122).
 * Errors like the reference: n_total < 2, tr out of range, size mismatch -> SB200_ERR_USER. */
int sb200_cheb_create(int rank, int tr, const int* dims, long long n_total, sb200_cheb** out);
/* ChebMult(A, vx, vy) (chebyshev.c:142-199): y = d x / d xi_tr on the CGL nodes cos(i pi/n). */
int sb200_cheb_apply(sb200_cheb* c, const double* d_x, double* d_y, void* stream);
int sb200_cheb_apply_host(sb200_cheb* c, const double* h_x, double* h_y);
/* ChebDestroy(A) (chebyshev.c:223-235). */
int sb200_cheb_destroy(sb200_cheb* c);
/* The P x P differentiation matrix the kernels apply (row-major, host buffer of P*P doubles). */
int sb200_cheb_matrix(int P, double* h_D);
/* The two half-size matrices the even-odd kernels apply instead (any P >= 2; D is centro-antisymmetric): pair j couples nodes j and P-1-j
 * (for odd P the middle node with itself), Ae = (D[i][j] + D[i][P-1-j]) / 2, Bo = (D[i][j] - D[i][P-1-j]) / 2, zero padded to HP x HP with
 * HP = 8 * ceil(ceil(P/2) / 8); with s_j = u_j + u_{P-1-j}, d_j = u_j - u_{P-1-j}:  (D u)_i = (Ae s)_i + (Bo d)_i,  (D u)_{P-1-i} = (Bo d)_i - (Ae s)_i.
 * *HP receives the padded size; h_Ae / h_Bo (HP*HP doubles each, row-major) may be NULL to query it. */
int sb200_cheb_even_odd(int P, int* HP, double* h_Ae, double* h_Bo);

/* ---- elliptic.C: MatCreate_Elliptic / MatMult_Elliptic / FormFunction ---------------------- */
/* MatCreate_Elliptic(comm, d, dim, flag, bf, &vG, &A) (elliptic.C:250-293) with the all-Dirichlet
 * boundary function the reference uses (DirichletBdy, elliptic.C:470-477). */
int sb200_elliptic_create(int d, const int* dim, sb200_elliptic** out);
/* DOF distribution printed at elliptic.C:424: local (m), global (g), dirichlet (nd). */
int sb200_elliptic_sizes(const sb200_elliptic* e, long long* m, long long* g, long long* nd);
/* -gamma / -exponent (elliptic.C:147-148): eta = 1 + gamma*u^exponent. */
int sb200_elliptic_set_params(sb200_elliptic* e, double gamma, double exponent);
/* c->dirichlet (nd doubles, lexicographic boundary order; elliptic.C:672) and ac->b (g doubles, :674). */
int sb200_elliptic_set_dirichlet(sb200_elliptic* e, const double* d_values, void* stream);
int sb200_elliptic_set_rhs(sb200_elliptic* e, const double* d_b, void* stream);
/* CreateExactSolution(snes, u, u2) (elliptic.C:594-677) on the HOST (set-up work): -exact 0 (separable cosine, needs
 * -cos_scale, which the reference reads without a default, :607-609), 1 (quadratics) or 2 (polynomials); writes the exact
 * solution u (g), the forcing u2 (g, what the driver copies into ac->b, :674) and the Dirichlet values (nd), any of which
 * may be NULL.  Needs no device; upload with sb200_elliptic_set_dirichlet / _set_rhs. */
int sb200_elliptic_exact_solution(int d, const int* dim, int exact, double cos_scale, double gamma, double exponent, double* h_u,
                                  double* h_u2, double* h_dirichlet);
/* MatMult_Elliptic(A, U, V) (elliptic.C:297-339): Jacobian action, U and V of g doubles. */
int sb200_elliptic_matmult(sb200_elliptic* e, const double* d_U, double* d_V, void* stream);
int sb200_elliptic_matmult_host(sb200_elliptic* e, const double* h_U, double* h_V);
/* Queued host-buffer form of the same call for callers that apply the operator to a stream of host vectors
 * (block Krylov right-hand sides, the bench's end-to-end leg): submit() enqueues copy-in -> MatMult_Elliptic ->
 * copy-out on three streams of the context and returns at once; wait() blocks until the OLDEST submitted
 * application has landed in its h_V.  At most SB200_HOST_QUEUE_DEPTH applications are in flight (submit() then
 * fails with SB200_ERR_USER), each through its own device vectors, so the copy-in of one overlaps the kernels
 * and the copy-out of its predecessors (PCIe is full duplex).  h_U / h_V should be pinned and must stay valid
 * and untouched until the matching wait(); drain the queue before any other call on this context. */
#define SB200_HOST_QUEUE_DEPTH 4
int sb200_elliptic_matmult_host_submit(sb200_elliptic* e, const double* h_U, double* h_V);
int sb200_elliptic_matmult_host_wait(sb200_elliptic* e);
int sb200_elliptic_matmult_host_pending(const sb200_elliptic* e, int* pending);
/* FormFunction(snes, U, rhs, ctx) (elliptic.C:481-533): residual; refreshes eta/deta/gradu caches. */
/* Name of the kernel path the last sb200_elliptic_matmult of this context ran (static string; bench.py's roofline.kernel). */
const char* sb200_elliptic_last_kernel(const sb200_elliptic* e);
int sb200_elliptic_function(sb200_elliptic* e, const double* d_U, double* d_F, void* stream);
int sb200_elliptic_function_host(sb200_elliptic* e, const double* h_U, double* h_F);
/* Cached state read by FormJacobian (elliptic.C:550-553): which = 0 eta, 1 deta, 2+k gradu[k];
 * copies m doubles device->device into d_out. */
int sb200_elliptic_get_state(sb200_elliptic* e, int which, double* d_out, void* stream);
/* FormJacobian(snes, w, &A, &P, &flag, ctx) (elliptic.C:537-590): the finite-difference preconditioning matrix P
 * about the state cached by the last FormFunction (eta, deta, gradu), assembled on the device as CSR with 32-bit
 * indices (PetscInt): g rows in the global Vec order, columns increasing within a row, Dirichlet neighbours dropped
 * like MatSetValues drops negative ids.  d_rowptr (nrows+1) and d_colidx (nnz) may both be NULL to refresh the values
 * alone (the pattern never changes: SAME_NONZERO_PATTERN, elliptic.C:588).  What PC is built from P stays PETSc's. */
int sb200_elliptic_jacobian_sizes(sb200_elliptic* e, long long* nrows, long long* nnz);
int sb200_elliptic_jacobian_csr(sb200_elliptic* e, int* d_rowptr, int* d_colidx, double* d_vals, void* stream);
/* Scatter helpers with the semantics of scatterGL+scatterDL / scatterLG (elliptic.C:426-434). */
int sb200_elliptic_pad(sb200_elliptic* e, const double* d_U, int with_dirichlet, double* d_local, void* stream);
int sb200_elliptic_crop(sb200_elliptic* e, const double* d_local, double* d_U, void* stream);
/* Select kernel path: 0 = auto, 1 = generic per-axis kernels, 2 = one fused chain kernel per axis,
 * 3 = single persistent chain kernel (2 needs equal extents P in {32, 64, 128}, 3 equal extents P % 16 == 0 from 32 to 160), 4 = the generic kernels captured once
 * into a CUDA graph and replayed (opt-in for the small, launch-bound grids; single GPU; same arithmetic as path 1). */
int sb200_elliptic_set_path(sb200_elliptic* e, int path);
/* Debug hook (only active in SB200_TRACE builds): device buffer receiving per-item phase clocks. */
int sb200_elliptic_debug_trace(sb200_elliptic* e, long long* d_buf);
/* ---- slab partition over the GPUs of one node (no counterpart in the reference, which builds every Vec with
 * VecCreateSeq on PETSC_COMM_SELF, elliptic.C:167,262; this is what an MPI-parallel MatCreate_Elliptic would
 * be).  One process per GPU; rank r keeps planes [r*dim[0]/nranks, (r+1)*dim[0]/nranks) of every local array
 * and the matching contiguous range [goff, goff + g) of the global Vec (lexicographic interior order), so
 * sb200_elliptic_sizes() reports LOCAL sizes and every vector argument is the local part.  Axis-0 derivatives
 * read / write the peers' memory directly over NVLink; all collective entry points (matmult, function) must be
 * called by every rank in the same order.  dim[0] must be divisible by nranks (<= 8). */
int sb200_elliptic_create_slab(int d, const int* dim, int rank, int nranks, sb200_elliptic** out);
/* The partition arithmetic alone (host only, no device needed): planes [i0, i0+nloc), Vec range [goff, goff+g_local). */
int sb200_slab_geometry(int d, const int* dim, int rank, int nranks, int* i0, int* nloc, long long* goff, long long* g_local,
                        long long* m_local, long long* nd_local);
int sb200_elliptic_slab_info(const sb200_elliptic* e, int* rank, int* nranks, int* i0, int* nloc, long long* goff, long long* gtotal);
/* Synchronises the stream and reports how many device-side waits on a peer's flag gave up (~4 s each):
 * non-zero means the ranks did not make the same sequence of collective calls (or a peer died). */
int sb200_elliptic_slab_status(sb200_elliptic* e, long long* timeouts, void* stream);
/* Peer mapping: each rank exports one CUDA IPC handle (sb200_ipc_handle_bytes() = 64 bytes), the host code
 * exchanges them (MPI_Allgather / torch.distributed.all_gather) and attaches every peer's handle.
 * attach_local maps a peer context living in the SAME process (several ranks driven by one process). */
int sb200_ipc_handle_bytes(void);
int sb200_elliptic_ipc_export(sb200_elliptic* e, void* handle);
int sb200_elliptic_ipc_attach(sb200_elliptic* e, int peer_rank, const void* handle);
int sb200_elliptic_attach_local(sb200_elliptic* e, int peer_rank, sb200_elliptic* peer);
/* Debug: global-timer stamps (ns; 10 (min,max) pairs) of the slab step's milestones, recorded when the
 * environment has SB200_XFLAGS bit 64; reading re-arms them.  See tools/slab_timeline.py. */
int sb200_elliptic_debug_timeline(sb200_elliptic* e, unsigned long long* h_out20, void* stream);
/* MatDestroy_Elliptic (elliptic.C:343-368). */
int sb200_elliptic_destroy(sb200_elliptic* e);

/* ---- stokes.C: the five MatShells, the SNES residual, the rheology ---------------------------- */
/* StokesCreate(comm, &A, &x, &ctx) (stokes.C:257-345) for -boundary 0 (all Dirichlet, stokes.C:463-468);
 * d must be 2 or 3 (StokesPressureReduceOrder, stokes.C:1036). */
int sb200_stokes_create(int d, const int* dim, sb200_stokes** out);
/* DOF distribution printed at stokes.C:891: local nodes m, global g, pressure gp, velocity gv, dirichlet dv. */
int sb200_stokes_sizes(const sb200_stokes* s, long long* m, long long* g, long long* gp, long long* gv, long long* dv);
/* -rheology / -hardness / -exponent / -eps / -gamma0 (stokes.C:404-410); type 0 linear, 1 power law.
 * The continuation loop (stokes.C:217-221) calls this with the per-step exponent / regularisation. */
int sb200_stokes_set_rheology(sb200_stokes* s, int type, double hardness, double exponent, double regularization, double gamma0);
/* c->dirichlet (dv doubles: boundary nodes in walk order x d components, stokes.C:796-801) and c->force (g doubles). */
int sb200_stokes_set_dirichlet(sb200_stokes* s, const double* d_values, void* stream);
int sb200_stokes_set_force(sb200_stokes* s, const double* d_force, void* stream);
/* StokesCreateExactSolution (stokes.C:942-1003) on the HOST with StokesExact0..3 (:1948-2034): exact solution U and forcing
 * U2 (g doubles each, AoS [v, p]) and the Dirichlet velocities StokesDirichlet gives the boundary nodes (dv doubles, :2039-2050);
 * any output may be NULL.  -exact 3 is 2-D only (:2022).  Needs no device. */
int sb200_stokes_exact_solution(int d, const int* dim, int exact, double* h_u, double* h_u2, double* h_dirichlet);
/* one node of StokesExact{exact} (stokes.C:1948-2034): value = [u_0..u_{d-1}, p] and the forcing rhs at the point coord (host) */
int sb200_stokes_exact_eval(int exact, int d, const double* coord, double* value, double* rhs);
/* StokesMatMult(A, x, y) (stokes.C:499-519): x, y of g doubles, AoS [v_0..v_{d-1}, p] per interior node. */
int sb200_stokes_matmult(sb200_stokes* s, const double* d_x, double* d_y, void* stream);
int sb200_stokes_matmult_host(sb200_stokes* s, const double* h_x, double* h_y);
/* StokesMatMultVV (stokes.C:623-676): gv -> gv. */
int sb200_stokes_matmult_vv(sb200_stokes* s, const double* d_x, double* d_y, void* stream);
/* StokesMatMultPV (stokes.C:557-566): divergence, gv -> gp. */
int sb200_stokes_matmult_pv(sb200_stokes* s, const double* d_x, double* d_y, void* stream);
/* StokesDivergence(ctx, withDirichlet, xG, yG) (stokes.C:570-595): the same with the Dirichlet velocities inserted on the boundary
 * when with_dirichlet != 0 (what the residual uses, :746); with_dirichlet = 0 is StokesMatMultPV. */
int sb200_stokes_divergence(sb200_stokes* s, int with_dirichlet, const double* d_x, double* d_y, void* stream);
/* Evaluation switch (default 1 since round 2; 0 = the reference's literal sequence): StokesMatMult and StokesFunction take their pressure rows  sum_i D_i v_i  from the trace of the velocity
 * gradient their viscous part computes anyway, instead of running StokesDivergence on the same input a second time
 * (stokes.C:509,746); same values bit for bit, one pad pass and d derivative passes fewer per application. */
int sb200_stokes_set_trace_divergence(sb200_stokes* s, int on);
/* Evaluation switch (default 1 since round 2): StokesMatMult and StokesFunction subtract the boundary-extrapolated pressure from the diagonal of the viscous
 * flux (the stress eta*eps - p I), so the divergence of the viscous part also produces the pressure gradient of StokesMatMultVP
 * (stokes.C:512-513,747-750): d derivative passes and one read-modify-write crop fewer.  Same operator, ~1e-15 relative rounding
 * difference (the two terms are summed before the derivative instead of after). */
int sb200_stokes_set_fold_pressure(sb200_stokes* s, int on);
/* Opt-in (default 0, single GPU): the linear shells StokesMatMult / VV / PV / VP replay their launch sequence from a CUDA graph
 * captured once on fixed staging vectors (same kernels, same results) - for small grids such as BASELINE config 4 (20^3), where a
 * shell is 5-22 launches of a few microseconds each and the inner solves of the saddle-point PCs call them hundreds of times. */
int sb200_stokes_set_graph(sb200_stokes* s, int on);
/* StokesMatMultVP (stokes.C:599-619): pressure gradient with P_N - P_{N-2} extrapolation, gp -> gv. */
int sb200_stokes_matmult_vp(sb200_stokes* s, const double* d_x, double* d_y, void* stream);
/* StokesMatGetDiagonalSchur (stokes.C:542-553): y = 1/eta at pressure nodes (gp doubles). */
int sb200_stokes_get_diagonal_schur(sb200_stokes* s, double* d_y, void* stream);
/* StokesMatMultSchur (stokes.C:523-535): y = -PV * solve(VP * x); `solve` stands for
 * KSPSolve(KSPSchurVelocity, rhs, sol) on device vectors of gv doubles and returns 0 on success. */
typedef int (*sb200_velocity_solve_fn)(void* ctx, const double* d_rhs, double* d_sol, void* stream);
int sb200_stokes_matmult_schur(sb200_stokes* s, const double* d_x, double* d_y, sb200_velocity_solve_fn solve, void* solve_ctx, void* stream);
/* StokesFunction(snes, x, y, ctx) (stokes.C:680-758): residual; refreshes strain / eta / deta caches. */
int sb200_stokes_function(sb200_stokes* s, const double* d_x, double* d_y, void* stream);
int sb200_stokes_function_host(sb200_stokes* s, const double* h_x, double* h_y);
/* "Minimum eta / Maximum eta" of the last residual evaluation (stokes.C:731-734); synchronises the stream. */
int sb200_stokes_eta_minmax(sb200_stokes* s, double* h_min, double* h_max, void* stream);
/* Cached state (stokes.C:766): which = 0 eta (m), 1 deta (m), 2+j strain[j] (m*d); device->device copy. */
int sb200_stokes_get_state(sb200_stokes* s, int which, double* d_out, void* stream);
/* StokesPCSetUp0 (stokes.C:1160-1240), Dirichlet rows (:1203-1224): the finite-difference velocity matrix MatVVPC about
 * the eta cached by the last StokesFunction, as device CSR (gv rows, row = interior node * d + component, 32-bit
 * indices, columns increasing); index arrays may be NULL to refresh the values alone. */
int sb200_stokes_pc_velocity_sizes(sb200_stokes* s, long long* nrows, long long* nnz);
int sb200_stokes_pc_velocity_csr(sb200_stokes* s, int* d_rowptr, int* d_colidx, double* d_vals, void* stream);
/* StokesPressureReduceOrder (stokes.C:1029-1080) applied in place to a local pressure array of m doubles. */
int sb200_stokes_pressure_reduce_order(sb200_stokes* s, double* d_pL, void* stream);
/* Slab partition of the Stokes shells over the GPUs of one node (see sb200_elliptic_create_slab): rank r keeps the planes
 * [r*dim[0]/nranks, ...) of every local field and the matching contiguous range of the global Vec (AoS [v,p] per interior
 * node: (d+1)*goff_nodes values precede it).  All sizes reported by sb200_stokes_sizes are then LOCAL.  Axis-0 derivatives
 * pull their operand planes from the owners over NVLink; the axis-0 pass of the pressure extrapolation is a rank-ordered sum. */
int sb200_stokes_create_slab(int d, const int* dim, int rank, int nranks, sb200_stokes** out);
int sb200_stokes_slab_info(const sb200_stokes* s, int* rank, int* nranks, int* i0, int* nloc, long long* goff_nodes);
int sb200_stokes_ipc_export(sb200_stokes* s, void* handle);
int sb200_stokes_ipc_attach(sb200_stokes* s, int peer_rank, const void* handle);
int sb200_stokes_attach_local(sb200_stokes* s, int peer_rank, sb200_stokes* peer);
int sb200_stokes_slab_status(sb200_stokes* s, long long* timeouts, void* stream);
/* StokesDestroy (stokes.C:348-388). */
int sb200_stokes_destroy(sb200_stokes* s);

/* ---- KSP: device-resident FGMRES(m), the solver the reference selects in code and the direct caller of the
 * MatShells (KSPSetType(ksp, KSPFGMRES): elliptic.C:181-182, stokes.C:155-157; inner KSPs stokes.C:328-341).
 * PETSc's algorithm restated: right-preconditioned flexible Arnoldi, classical Gram-Schmidt, KSPConvergedDefault.
 * Operator and preconditioner are callbacks on DEVICE vectors (MatMult / PCApply); pc may be NULL (PCNONE).  The
 * PC itself (ILU / hypre / LU on the finite-difference matrix) stays PETSc's and is out of scope. */
typedef struct sb200_ksp sb200_ksp;
typedef int (*sb200_apply_fn)(void* ctx, const double* d_x, double* d_y, void* stream);
int sb200_ksp_create(long long n, int restart, sb200_ksp** out);                 /* KSPCreate + KSPGMRESSetRestart (default 30) */
/* vectors are slab-partitioned: n_local values here, dot products summed over the ranks through peer memory */
int sb200_ksp_create_slab(long long n_local, int restart, int rank, int nranks, sb200_ksp** out);
int sb200_ksp_set_operators(sb200_ksp* k, sb200_apply_fn op, void* op_ctx, sb200_apply_fn pc, void* pc_ctx); /* KSPSetOperators / PCShell */
int sb200_ksp_set_tolerances(sb200_ksp* k, double rtol, double atol, double dtol, int maxits);                 /* KSPSetTolerances */
/* Opt-in (default 0): with depth 1 the Arnoldi step k+1 is enqueued BEFORE the host reads the residual norm of step k, so the GPU
 * does not idle on the host's convergence decision (about 25 us per iteration at 128^3).  Same iterates, same iteration count and
 * history; the step enqueued behind the last one is discarded, i.e. a converged solve applies the operator (and the PC callback)
 * at most once more than PETSc would. */
int sb200_ksp_set_lookahead(sb200_ksp* k, int depth);
/* KSPSolve(ksp, b, x); guess_nonzero = KSPSetInitialGuessNonzero.  Synchronises the stream once per iteration. */
int sb200_ksp_solve(sb200_ksp* k, const double* d_b, double* d_x, int guess_nonzero, void* stream);
/* KSPGetIterationNumber / KSPGetResidualNorm / KSPGetConvergedReason (PETSc's reason codes: 2 rtol, 3 atol, -3 its, -4 dtol). */
int sb200_ksp_get_result(const sb200_ksp* k, int* its, double* rnorm, double* bnorm, int* reason);
int sb200_ksp_get_history(const sb200_ksp* k, double* h_hist, int cap, int* n);   /* KSPGetResidualHistory */
/* CUDA-event times of the last solve, split the way the north star asks: operator / PC ("timed separately") / KSP vector work. */
int sb200_ksp_get_times(const sb200_ksp* k, double* ms_operator, double* ms_pc, double* ms_orthogonalisation);
/* Sum of count <= 64 device doubles over the ranks of a slab-partitioned solver, in rank order (same bits on every rank), in place;
 * collective, no-op on one rank.  (The all-reduce FGMRES uses for its dot products: peer memory + epoch flags, no NCCL.) */
int sb200_ksp_allreduce_sum(sb200_ksp* k, double* d_vals, int count, void* stream);
int sb200_ksp_ipc_export(sb200_ksp* k, void* handle);
int sb200_ksp_ipc_attach(sb200_ksp* k, int peer_rank, const void* handle);
int sb200_ksp_attach_local(sb200_ksp* k, int peer_rank, sb200_ksp* peer);
int sb200_ksp_destroy(sb200_ksp* k);
/* The MatShell MULT operations in sb200_apply_fn form (ctx = the operator context). */
int sb200_apply_elliptic_matmult(void* ctx, const double* d_x, double* d_y, void* stream);
int sb200_apply_stokes_matmult(void* ctx, const double* d_x, double* d_y, void* stream);
int sb200_apply_stokes_matmult_vv(void* ctx, const double* d_x, double* d_y, void* stream);

/* ---- vector helpers: what the saddle-point preconditioners are composed from besides shells and inner solves -------------
 * scatterGV / scatterGP and scatterVG / scatterPG (stokes.C:867-877): the global Stokes vector holds [v_0..v_{d-1}, p] per
 * interior node; split / merge move the velocity (nodes*d) and pressure (nodes) parts out and back.  A NULL part is skipped
 * (merge then leaves those entries of d_x untouched). */
int sb200_vec_split(long long nodes, int d, const double* d_x, double* d_v, double* d_p, void* stream);
int sb200_vec_merge(long long nodes, int d, const double* d_v, const double* d_p, double* d_x, void* stream);
/* VecAXPBY: y = a x + b y (b == 0 does not read y; a == 0 is VecScale and does not read x) */
int sb200_vec_axpby(long long n, double a, const double* d_x, double b, double* d_y, void* stream);
/* VecPointwiseDivide: y = x / diag (PCJacobi on the "diagonal" of StokesMatGetDiagonalSchur, stokes.C:328-333) */
int sb200_vec_pointwise_divide(long long n, const double* d_x, const double* d_diag, double* d_y, void* stream);
/* MatGetDiagonal of a device CSR matrix (the SeqAIJ preconditioning matrices of FormJacobian / StokesPCSetUp0): what PCJacobi needs,
 * without bringing the matrix to the host; rows without a stored diagonal entry give 0 */
int sb200_csr_diagonal(long long nrows, const int* d_rowptr, const int* d_colidx, const double* d_vals, double* d_diag, void* stream);
/* MatNullSpaceRemove with the constant vector on the n entries x[offset + i*stride] (stokes.C:1013-1023: the pressure slots of
 * the global vector are stride d+1, offset d).  d_scratch: SB200_REDUCE_SCRATCH_DOUBLES doubles of device memory. */
#define SB200_REDUCE_SCRATCH_DOUBLES 1024
int sb200_vec_remove_mean(long long n, int stride, int offset, double* d_x, double* d_scratch, void* stream);
/* The same in two steps for vectors that are slab-partitioned over several ranks: d_out2 = {sum of the local entries, local count}
 * (fixed summation order); after the ranks have added their pairs (sb200_ksp_allreduce_sum), x -= d_sums2[0] / d_sums2[1]. */
int sb200_vec_sum_count(long long n, int stride, int offset, const double* d_x, double* d_scratch, double* d_out2, void* stream);
int sb200_vec_shift_mean(long long n, int stride, int offset, double* d_x, const double* d_sums2, void* stream);

/* ---- StokesPCApply0..3 (stokes.C:1714-1817): the saddle-point preconditioners, device resident -----------------------------
 * Composition of the PV / VP / VV shells with the three inner Krylov solves of stokes.C:328-341 - KSPVelocity, KSPSchur (on the
 * Schur shell, Jacobi from StokesMatGetDiagonalSchur, constant null space) and KSPSchurVelocity - each PETSc's default GMRES(30)
 * with LEFT preconditioning, run as sb200_ksp on M^-1 A.  Every vector stays on the device; only the preconditioner of the
 * velocity block (PETSc's PC on MatVVPC, out of scope) is a callback.  type = -pc_saddle_type (stokes.C:171-185): 0 block LU,
 * 1 upper triangular, 2 diagonal, 3 lower triangular. */
typedef struct sb200_saddle sb200_saddle;
int sb200_saddle_create(sb200_stokes* s, int type, sb200_saddle** out);
/* -vel_pc_* / -svel_pc_*: z = M^-1 r on device vectors of gv doubles; NULL = PCNONE.  svel_pc NULL with svel_same != 0 reuses vel_pc. */
int sb200_saddle_set_velocity_pc(sb200_saddle* p, sb200_apply_fn vel_pc, void* vel_ctx, sb200_apply_fn svel_pc, void* svel_ctx, int svel_same);
/* -vel_ksp_rtol / -vel_ksp_max_it, -schur_ksp_rtol / -schur_ksp_max_it, -svel_ksp_type preonly (PETSc defaults: 1e-5, 10000, gmres) */
int sb200_saddle_set_inner(sb200_saddle* p, double vel_rtol, int vel_maxits, double schur_rtol, int schur_maxits, int svel_preonly);
/* -svel_ksp_rtol / -svel_ksp_max_it: KSPSchurVelocity has its own options prefix (stokes.C:338-341); PETSc defaults 1e-5, 10000.
 * Only read when -svel_ksp_type is not preonly. */
int sb200_saddle_set_svel(sb200_saddle* p, double svel_rtol, int svel_maxits);
/* Slab-partitioned Stokes context (config 5 as a SOLVE over 2/4/8 GPUs): the three inner Krylov solvers are slab solvers whose dot products
 * cross the ranks, and the constant-pressure null space is removed with a cross-rank mean.  After sb200_saddle_set_inner / _set_svel call
 * sb200_saddle_prepare (creates the inner solvers), exchange the SB200_SADDLE_HANDLE_BYTES-byte handle of sb200_saddle_ipc_export between
 * the ranks (MPI_Allgather / torch.distributed.all_gather) and attach every peer's; sb200_saddle_attach_local maps a peer context that lives
 * in the same process (tests).  Every sb200_saddle_apply is then a collective.  On one rank none of this is needed. */
#define SB200_SADDLE_HANDLE_BYTES 192
int sb200_saddle_prepare(sb200_saddle* p);
int sb200_saddle_ipc_export(sb200_saddle* p, void* handle192);
int sb200_saddle_ipc_attach(sb200_saddle* p, int peer_rank, const void* handle192);
int sb200_saddle_attach_local(sb200_saddle* p, int peer_rank, sb200_saddle* peer);
int sb200_saddle_apply(sb200_saddle* p, const double* d_x, double* d_y, void* stream); /* y = StokesPCApply{type}(x), g doubles */
/* StokesRemoveConstantPressure's null space (stokes.C:1006-1025) applied to a global vector in place */
int sb200_saddle_remove_constant_pressure(sb200_saddle* p, double* d_x, void* stream);
/* sb200_apply_fn form for the pc slot of the outer KSP: PCApply followed by the null-space removal KSPSetNullSpace adds */
int sb200_apply_saddle(void* ctx, const double* d_x, double* d_y, void* stream);
/* inner iterations accumulated since creation (KSPVelocity, KSPSchur) */
int sb200_saddle_get_inner_its(const sb200_saddle* p, long long* velocity, long long* schur);
int sb200_saddle_destroy(sb200_saddle* p);

/* ---- host stand-in for PETSc's PCILU (NOT part of the B200 path; the PC stays PETSc's own) --------------------------------
 * The reference sets PCILU with 2 levels of fill on the finite-difference matrix in code (elliptic.C:183-184); PETSc's default
 * PC for the Stokes matrix MatVVPC is ILU(0).  So that the command-line drivers and the solver-level parity tests can run those
 * DEFAULT configurations without PETSc, this is the textbook level-of-fill ILU(k) on HOST CSR arrays (natural ordering, no
 * pivoting, no shift; columns must increase within a row).  solve() applies (LU)^-1; refactor() takes new values on the same
 * pattern (SAME_NONZERO_PATTERN). */
typedef struct sb200_host_ilu sb200_host_ilu;
int sb200_host_ilu_create(int n, const int* h_rowptr, const int* h_colidx, const double* h_vals, int levels, sb200_host_ilu** out);
int sb200_host_ilu_refactor(sb200_host_ilu* f, const double* h_vals);
int sb200_host_ilu_solve(const sb200_host_ilu* f, const double* h_b, double* h_x);
int sb200_host_ilu_nnz(const sb200_host_ilu* f, long long* nnz);
/* factor as CSR: rowptr (n+1), colidx / vals (nnz): strictly lower part = L (unit diagonal implied), diagonal and above = U */
int sb200_host_ilu_get(const sb200_host_ilu* f, int* h_rowptr, int* h_colidx, double* h_vals);
int sb200_host_ilu_destroy(sb200_host_ilu* f);

#ifdef __cplusplus
}
#endif
#endif /* SPECTRAL_B200_H */
