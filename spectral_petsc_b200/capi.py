"""ctypes binding of include/spectral_b200.h plus thin operator classes.

The classes keep the reference's names and argument meaning:
  Cheb      <-> MatCreateCheb / ChebMult / ChebDestroy       (chebyshev.c:89-235)
  Elliptic  <-> MatCreate_Elliptic / MatMult_Elliptic / FormFunction (elliptic.C:250-533)
Vectors are torch fp64 CUDA tensors (the VECCUDA-backed Vec); only their data_ptr() crosses the ABI.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# SB200_ABLATE_LIB=1 selects the diagnostic build (`make ablate`: ablation / timeline switches compiled in); the production library
# has none of them.  bench.py records every SB200_* variable it sees.
lib_path = os.path.join(_HERE, "libspectral_b200_ablate.so" if os.environ.get("SB200_ABLATE_LIB") == "1" else "libspectral_b200.so")


class SB200Error(RuntimeError):
    """Non-zero PetscErrorCode-style return from the C ABI."""

    def __init__(self, code, msg):
        super().__init__("sb200 error %d: %s" % (code, msg))
        self.code = code


_lib = None


def lib():
    """Load the CUDA library; fail loudly if it was not built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(lib_path):
            raise SB200Error(-1, "CUDA library %s is missing: run `make` or __graft_entry__.build()" % lib_path)
        L = ctypes.CDLL(lib_path)
        L.sb200_last_error.restype = ctypes.c_char_p
        L.sb200_launch_count.restype = ctypes.c_longlong
        _lib = L
    return _lib


def _ck(rc):
    if rc != 0:
        raise SB200Error(rc, lib().sb200_last_error().decode())


def launch_count():
    return int(lib().sb200_launch_count())


def _ptr(t):
    import torch

    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
        raise TypeError("expected a contiguous fp64 CUDA tensor")
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    import torch

    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _hptr(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.c_void_p)


def cheb_matrix(P):
    """The P x P CGL differentiation matrix the kernels apply (host, numpy)."""
    D = np.empty((P, P))
    _ck(lib().sb200_cheb_matrix(ctypes.c_int(P), _hptr(D)))
    return D


def cheb_even_odd(P):
    """The zero-padded half-size matrices (Ae, Bo) of the even-odd kernels (host, numpy; sb200_cheb_even_odd)."""
    hp = ctypes.c_int(0)
    _ck(lib().sb200_cheb_even_odd(ctypes.c_int(P), ctypes.byref(hp), None, None))
    Ae, Bo = np.empty((hp.value, hp.value)), np.empty((hp.value, hp.value))
    _ck(lib().sb200_cheb_even_odd(ctypes.c_int(P), ctypes.byref(hp), _hptr(Ae), _hptr(Bo)))
    return Ae, Bo


def elliptic_exact_solution(dim, exact, cos_scale=0.0, gamma=0.0, exponent=2.0):
    """CreateExactSolution (elliptic.C:594-677) on the host: (u, u2, dirichlet) as numpy arrays in the reference's Vec order."""
    dim = [int(v) for v in dim]
    g = int(np.prod([v - 2 for v in dim]))
    m = int(np.prod(dim))
    u, u2, dr = np.empty(g), np.empty(g), np.empty(m - g)
    _ck(lib().sb200_elliptic_exact_solution(ctypes.c_int(len(dim)), (ctypes.c_int * len(dim))(*dim), ctypes.c_int(exact), ctypes.c_double(cos_scale),
                                            ctypes.c_double(gamma), ctypes.c_double(exponent), _hptr(u), _hptr(u2), _hptr(dr)))
    return u, u2, dr


def stokes_exact_solution(dim, exact):
    """StokesCreateExactSolution + StokesDirichlet (stokes.C:942-1003, 1948-2050) on the host: (U, U2, dirichlet velocities)."""
    dim = [int(v) for v in dim]
    d = len(dim)
    gp = int(np.prod([v - 2 for v in dim]))
    m = int(np.prod(dim))
    U, U2, dr = np.empty(gp * (d + 1)), np.empty(gp * (d + 1)), np.empty((m - gp) * d)
    _ck(lib().sb200_stokes_exact_solution(ctypes.c_int(d), (ctypes.c_int * d)(*dim), ctypes.c_int(exact), _hptr(U), _hptr(U2), _hptr(dr)))
    return U, U2, dr


class HostILU:
    """ILU(levels) of a host CSR matrix: the stand-in for PETSc's PCILU (sb200_host_ilu_*; NOT part of the B200 path)."""

    def __init__(self, P, levels=0):
        P = P.tocsr().copy()
        P.sort_indices()
        self.n = P.shape[0]
        self._rowptr = np.ascontiguousarray(P.indptr, dtype=np.int32)
        self._colidx = np.ascontiguousarray(P.indices, dtype=np.int32)
        self._h = ctypes.c_void_p()
        vals = np.ascontiguousarray(P.data, dtype=np.float64)
        ip = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        _ck(lib().sb200_host_ilu_create(ctypes.c_int(self.n), ip(self._rowptr), ip(self._colidx), _hptr(vals), ctypes.c_int(levels), ctypes.byref(self._h)))

    def refactor(self, P):
        P = P.tocsr().copy()
        P.sort_indices()
        assert np.array_equal(P.indptr, self._rowptr) and np.array_equal(P.indices, self._colidx), "SAME_NONZERO_PATTERN required"
        _ck(lib().sb200_host_ilu_refactor(self._h, _hptr(np.ascontiguousarray(P.data, dtype=np.float64))))

    def solve(self, b):
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.empty_like(b)
        _ck(lib().sb200_host_ilu_solve(self._h, _hptr(b), _hptr(x)))
        return x

    @property
    def nnz(self):
        n = ctypes.c_longlong()
        _ck(lib().sb200_host_ilu_nnz(self._h, ctypes.byref(n)))
        return n.value

    def factor(self):
        """(rowptr, colidx, vals) of the combined factor: strictly lower = L (unit diagonal implied), rest = U."""
        rp, ci, v = np.empty(self.n + 1, dtype=np.int32), np.empty(self.nnz, dtype=np.int32), np.empty(self.nnz)
        ip = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        _ck(lib().sb200_host_ilu_get(self._h, ip(rp), ip(ci), _hptr(v)))
        return rp, ci, v

    def __del__(self):
        try:
            if self._h:
                lib().sb200_host_ilu_destroy(self._h)
                self._h = ctypes.c_void_p()
        except Exception:
            pass


def _csr_call(sizes_fn, csr_fn, handle, pattern):
    """Shared by Elliptic.jacobian_csr / Stokes.pc_velocity_csr: device CSR (int32 rowptr, int32 colidx, fp64 vals)."""
    import torch

    nrows, nnz = ctypes.c_longlong(), ctypes.c_longlong()
    _ck(sizes_fn(handle, ctypes.byref(nrows), ctypes.byref(nnz)))
    vals = torch.empty(nnz.value, dtype=torch.float64, device="cuda")
    if pattern is None:
        rowptr = torch.empty(nrows.value + 1, dtype=torch.int32, device="cuda")
        colidx = torch.empty(nnz.value, dtype=torch.int32, device="cuda")
        _ck(csr_fn(handle, ctypes.c_void_p(rowptr.data_ptr()), ctypes.c_void_p(colidx.data_ptr()), _ptr(vals), _stream()))
    else:
        rowptr, colidx = pattern
        _ck(csr_fn(handle, None, None, _ptr(vals), _stream()))
    return rowptr, colidx, vals


class Cheb:
    """MatCreateCheb(comm, rank, tr, dims, flag, vx, vy, &A): y = ChebMult(A, x)."""

    def __init__(self, rank, tr, dims, n_total=None):
        dims = [int(v) for v in dims]
        arr = (ctypes.c_int * len(dims))(*dims)
        n = int(np.prod(dims[:rank])) if n_total is None else int(n_total)
        self._h = ctypes.c_void_p()
        _ck(lib().sb200_cheb_create(ctypes.c_int(rank), ctypes.c_int(tr), arr, ctypes.c_longlong(n), ctypes.byref(self._h)))
        self.N = n

    def mult(self, x, y=None):
        import torch

        if y is None:
            y = torch.empty_like(x)
        if x.numel() != self.N or y.numel() != self.N:
            raise SB200Error(83, "vector length does not match the operator")
        _ck(lib().sb200_cheb_apply(self._h, _ptr(x), _ptr(y), _stream()))
        return y

    def mult_host(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty_like(x)
        _ck(lib().sb200_cheb_apply_host(self._h, _hptr(x), _hptr(y)))
        return y

    def destroy(self):
        if self._h:
            _ck(lib().sb200_cheb_destroy(self._h))
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class Elliptic:
    """MatCreate_Elliptic + the MatMult_Elliptic / FormFunction callbacks (elliptic.C)."""

    def __init__(self, dim, gamma=0.0, exponent=2.0, rank=0, nranks=1):
        """nranks > 1: slab partition along axis 0 (one rank per GPU); m, g, nd are then LOCAL sizes and
        every vector is the local part.  The peers must be attached (dist.attach_peers / attach_in_process)
        before the first collective call."""
        dim = [int(v) for v in dim]
        arr = (ctypes.c_int * len(dim))(*dim)
        self._h = ctypes.c_void_p()
        if nranks == 1:
            _ck(lib().sb200_elliptic_create(ctypes.c_int(len(dim)), arr, ctypes.byref(self._h)))
        else:
            _ck(lib().sb200_elliptic_create_slab(ctypes.c_int(len(dim)), arr, ctypes.c_int(rank), ctypes.c_int(nranks), ctypes.byref(self._h)))
        m, g, nd = ctypes.c_longlong(), ctypes.c_longlong(), ctypes.c_longlong()
        _ck(lib().sb200_elliptic_sizes(self._h, ctypes.byref(m), ctypes.byref(g), ctypes.byref(nd)))
        self.dim, self.d = dim, len(dim)
        self.m, self.g, self.nd = m.value, g.value, nd.value
        self.rank, self.nranks = rank, nranks
        i0, nloc, goff, gtot = ctypes.c_int(), ctypes.c_int(), ctypes.c_longlong(), ctypes.c_longlong()
        _ck(lib().sb200_elliptic_slab_info(self._h, None, None, ctypes.byref(i0), ctypes.byref(nloc), ctypes.byref(goff), ctypes.byref(gtot)))
        self.i0, self.nloc, self.goff, self.gtotal = i0.value, nloc.value, goff.value, gtot.value
        self.set_params(gamma, exponent)

    # peer mapping for the slab partition (see dist.py)
    def ipc_export(self):
        buf = ctypes.create_string_buffer(lib().sb200_ipc_handle_bytes())
        _ck(lib().sb200_elliptic_ipc_export(self._h, buf))
        return buf.raw

    def ipc_attach(self, peer_rank, handle):
        _ck(lib().sb200_elliptic_ipc_attach(self._h, ctypes.c_int(peer_rank), ctypes.c_char_p(handle)))

    def slab_timeouts(self):
        n = ctypes.c_longlong()
        _ck(lib().sb200_elliptic_slab_status(self._h, ctypes.byref(n), _stream()))
        return n.value

    def attach_local(self, peer_rank, peer):
        _ck(lib().sb200_elliptic_attach_local(self._h, ctypes.c_int(peer_rank), peer._h))

    def set_params(self, gamma, exponent):
        _ck(lib().sb200_elliptic_set_params(self._h, ctypes.c_double(gamma), ctypes.c_double(exponent)))

    def set_path(self, path):
        _ck(lib().sb200_elliptic_set_path(self._h, ctypes.c_int(path)))

    def kernel_name(self):
        """Which kernel path the last mat_mult ran (bench.py reports it as roofline.kernel)."""
        f = lib().sb200_elliptic_last_kernel
        f.restype = ctypes.c_char_p
        return f(self._h).decode()

    def set_dirichlet(self, values):
        assert values.numel() == self.nd
        _ck(lib().sb200_elliptic_set_dirichlet(self._h, _ptr(values), _stream()))

    def set_rhs(self, b):
        assert b.numel() == self.g
        _ck(lib().sb200_elliptic_set_rhs(self._h, _ptr(b), _stream()))

    def mat_mult(self, U, V=None):
        import torch

        if V is None:
            V = torch.empty_like(U)
        assert U.numel() == self.g and V.numel() == self.g
        _ck(lib().sb200_elliptic_matmult(self._h, _ptr(U), _ptr(V), _stream()))
        return V

    def form_function(self, U, F=None):
        import torch

        if F is None:
            F = torch.empty_like(U)
        assert U.numel() == self.g and F.numel() == self.g
        _ck(lib().sb200_elliptic_function(self._h, _ptr(U), _ptr(F), _stream()))
        return F

    def mat_mult_host(self, U):
        U = np.ascontiguousarray(U, dtype=np.float64)
        V = np.empty_like(U)
        _ck(lib().sb200_elliptic_matmult_host(self._h, _hptr(U), _hptr(V)))
        return V

    def form_function_host(self, U):
        U = np.ascontiguousarray(U, dtype=np.float64)
        F = np.empty_like(U)
        _ck(lib().sb200_elliptic_function_host(self._h, _hptr(U), _hptr(F)))
        return F

    HOST_QUEUE_DEPTH = 4  # SB200_HOST_QUEUE_DEPTH

    def mat_mult_host_submit(self, U, V):
        """Queued host-buffer MatMult (sb200_elliptic_matmult_host_submit): U, V are numpy views of (pinned) host
        memory that must stay alive and untouched until the matching mat_mult_host_wait()."""
        assert U.size == self.g and V.size == self.g
        _ck(lib().sb200_elliptic_matmult_host_submit(self._h, _hptr(U), _hptr(V)))

    def mat_mult_host_wait(self):
        _ck(lib().sb200_elliptic_matmult_host_wait(self._h))

    def mat_mult_host_pending(self):
        n = ctypes.c_int()
        _ck(lib().sb200_elliptic_matmult_host_pending(self._h, ctypes.byref(n)))
        return n.value

    def mat_mult_host_stream(self, Us, Vs):
        """Apply the operator to every vector of Us (host arrays) into Vs, keeping the queue full."""
        for U, V in zip(Us, Vs):
            if self.mat_mult_host_pending() == self.HOST_QUEUE_DEPTH:
                self.mat_mult_host_wait()
            self.mat_mult_host_submit(U, V)
        while self.mat_mult_host_pending():
            self.mat_mult_host_wait()

    def get_state(self, which):
        import torch

        out = torch.empty(self.m, dtype=torch.float64, device="cuda")
        _ck(lib().sb200_elliptic_get_state(self._h, ctypes.c_int(which), _ptr(out), _stream()))
        return out

    def jacobian_csr(self, pattern=None):
        """FormJacobian (elliptic.C:537-590): the FD preconditioning matrix about the state of the last form_function as
        device CSR (rowptr, colidx, vals).  pattern=(rowptr, colidx) from an earlier call refreshes the values only."""
        return _csr_call(lib().sb200_elliptic_jacobian_sizes, lib().sb200_elliptic_jacobian_csr, self._h, pattern)

    def pad(self, U, with_dirichlet=False):
        import torch

        out = torch.empty(self.m, dtype=torch.float64, device="cuda")
        _ck(lib().sb200_elliptic_pad(self._h, _ptr(U), ctypes.c_int(int(with_dirichlet)), _ptr(out), _stream()))
        return out

    def crop(self, local):
        import torch

        out = torch.empty(self.g, dtype=torch.float64, device="cuda")
        _ck(lib().sb200_elliptic_crop(self._h, _ptr(local), _ptr(out), _stream()))
        return out

    def destroy(self):
        if self._h:
            _ck(lib().sb200_elliptic_destroy(self._h))
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


_VSOLVE = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p)


class Stokes:
    """StokesCreate + the StokesMatMult{,VV,PV,VP,Schur} shells and StokesFunction (stokes.C), -boundary 0."""

    def __init__(self, dim, rheology=0, hardness=1.0, exponent=1.0, regularization=1.0, gamma0=1.0, rank=0, nranks=1):
        """nranks > 1: slab partition along axis 0 (see Elliptic); all sizes and vectors are then the local parts."""
        dim = [int(v) for v in dim]
        arr = (ctypes.c_int * len(dim))(*dim)
        self._h = ctypes.c_void_p()
        if nranks == 1:
            _ck(lib().sb200_stokes_create(ctypes.c_int(len(dim)), arr, ctypes.byref(self._h)))
        else:
            _ck(lib().sb200_stokes_create_slab(ctypes.c_int(len(dim)), arr, ctypes.c_int(rank), ctypes.c_int(nranks), ctypes.byref(self._h)))
        v = [ctypes.c_longlong() for _ in range(5)]
        _ck(lib().sb200_stokes_sizes(self._h, *[ctypes.byref(x) for x in v]))
        self.m, self.g, self.gp, self.gv, self.dv = [x.value for x in v]
        self.dim, self.d = dim, len(dim)
        self.rank, self.nranks = rank, nranks
        i0, nloc, goff = ctypes.c_int(), ctypes.c_int(), ctypes.c_longlong()
        _ck(lib().sb200_stokes_slab_info(self._h, None, None, ctypes.byref(i0), ctypes.byref(nloc), ctypes.byref(goff)))
        self.i0, self.nloc, self.goff = i0.value, nloc.value, goff.value
        self.set_rheology(rheology, hardness, exponent, regularization, gamma0)

    def ipc_export(self):
        buf = ctypes.create_string_buffer(lib().sb200_ipc_handle_bytes())
        _ck(lib().sb200_stokes_ipc_export(self._h, buf))
        return buf.raw

    def ipc_attach(self, peer_rank, handle):
        _ck(lib().sb200_stokes_ipc_attach(self._h, ctypes.c_int(peer_rank), ctypes.c_char_p(handle)))

    def attach_local(self, peer_rank, peer):
        _ck(lib().sb200_stokes_attach_local(self._h, ctypes.c_int(peer_rank), peer._h))

    def slab_timeouts(self):
        n = ctypes.c_longlong()
        _ck(lib().sb200_stokes_slab_status(self._h, ctypes.byref(n), _stream()))
        return n.value

    def set_rheology(self, rheology, hardness=1.0, exponent=1.0, regularization=1.0, gamma0=1.0):
        _ck(lib().sb200_stokes_set_rheology(self._h, ctypes.c_int(rheology), ctypes.c_double(hardness), ctypes.c_double(exponent),
                                            ctypes.c_double(regularization), ctypes.c_double(gamma0)))

    def set_trace_divergence(self, on):
        """Evaluation switch (on by default): pressure rows of mat_mult / function from the trace of the gradient the viscous part computes (same bits)."""
        _ck(lib().sb200_stokes_set_trace_divergence(self._h, ctypes.c_int(int(on))))

    def set_graph(self, on):
        """Opt-in: the linear shells replay their launches from CUDA graphs (small, launch-bound grids)."""
        _ck(lib().sb200_stokes_set_graph(self._h, ctypes.c_int(int(on))))

    def set_fold_pressure(self, on):
        """Evaluation switch (on by default): the pressure gradient of mat_mult / function comes out of the viscous divergence (flux = eta*eps - p I)."""
        _ck(lib().sb200_stokes_set_fold_pressure(self._h, ctypes.c_int(int(on))))

    def set_dirichlet(self, values):
        assert values.numel() == self.dv
        _ck(lib().sb200_stokes_set_dirichlet(self._h, _ptr(values), _stream()))

    def set_force(self, force):
        assert force.numel() == self.g
        _ck(lib().sb200_stokes_set_force(self._h, _ptr(force), _stream()))

    def _apply(self, fn, x, nin, nout, y=None):
        import torch

        assert x.numel() == nin
        if y is None:
            y = torch.empty(nout, dtype=torch.float64, device=x.device)
        _ck(fn(self._h, _ptr(x), _ptr(y), _stream()))
        return y

    def mat_mult(self, x, y=None):
        return self._apply(lib().sb200_stokes_matmult, x, self.g, self.g, y)

    def mat_mult_vv(self, x, y=None):
        return self._apply(lib().sb200_stokes_matmult_vv, x, self.gv, self.gv, y)

    def mat_mult_pv(self, x, y=None):
        return self._apply(lib().sb200_stokes_matmult_pv, x, self.gv, self.gp, y)

    def mat_mult_vp(self, x, y=None):
        return self._apply(lib().sb200_stokes_matmult_vp, x, self.gp, self.gv, y)

    def function(self, x, y=None):
        return self._apply(lib().sb200_stokes_function, x, self.g, self.g, y)

    def mat_mult_host(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty_like(x)
        _ck(lib().sb200_stokes_matmult_host(self._h, _hptr(x), _hptr(y)))
        return y

    def function_host(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty_like(x)
        _ck(lib().sb200_stokes_function_host(self._h, _hptr(x), _hptr(y)))
        return y

    def get_diagonal_schur(self):
        import torch

        y = torch.empty(self.gp, dtype=torch.float64, device="cuda")
        _ck(lib().sb200_stokes_get_diagonal_schur(self._h, _ptr(y), _stream()))
        return y

    def mat_mult_schur(self, x, velocity_solve):
        """velocity_solve(rhs_tensor) -> sol_tensor stands for KSPSolve(KSPSchurVelocity) (stokes.C:531)."""
        import torch

        y = torch.empty(self.gp, dtype=torch.float64, device=x.device)
        gv = self.gv

        def cb(_ctx, d_rhs, d_sol, _stream_):
            try:
                rhs = _wrap(d_rhs, gv)
                sol = _wrap(d_sol, gv)
                sol.copy_(velocity_solve(rhs))
                return 0
            except Exception:  # pragma: no cover
                return 1

        cfn = _VSOLVE(cb)
        _ck(lib().sb200_stokes_matmult_schur(self._h, _ptr(x), _ptr(y), cfn, None, _stream()))
        return y

    def eta_minmax(self):
        a, b = ctypes.c_double(), ctypes.c_double()
        _ck(lib().sb200_stokes_eta_minmax(self._h, ctypes.byref(a), ctypes.byref(b), _stream()))
        return a.value, b.value

    def get_state(self, which):
        import torch

        n = self.m if which < 2 else self.m * self.d
        out = torch.empty(n, dtype=torch.float64, device="cuda")
        _ck(lib().sb200_stokes_get_state(self._h, ctypes.c_int(which), _ptr(out), _stream()))
        return out

    def pc_velocity_csr(self, pattern=None):
        """StokesPCSetUp0 (stokes.C:1160-1240): the FD velocity matrix MatVVPC about the eta of the last function() call."""
        return _csr_call(lib().sb200_stokes_pc_velocity_sizes, lib().sb200_stokes_pc_velocity_csr, self._h, pattern)

    def pressure_reduce_order(self, pL):
        assert pL.numel() == self.m
        _ck(lib().sb200_stokes_pressure_reduce_order(self._h, _ptr(pL), _stream()))
        return pL

    def destroy(self):
        if self._h:
            _ck(lib().sb200_stokes_destroy(self._h))
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


_APPLY = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p)


class KSP:
    """KSPCreate + KSPSetType(KSPFGMRES) (elliptic.C:181-182, stokes.C:155-157): device-resident FGMRES(restart).

    The operator is a MatShell context of this package (Elliptic / Stokes: its MULT runs natively, no Python in
    the iteration) or a Python callable on torch tensors; the preconditioner is a Python callable (the PC stays
    outside this package, e.g. a host LU of the finite-difference matrix) or None."""

    def __init__(self, n, restart=30, rank=0, nranks=1):
        self._h = ctypes.c_void_p()
        _ck(lib().sb200_ksp_create_slab(ctypes.c_longlong(n), ctypes.c_int(restart), ctypes.c_int(rank), ctypes.c_int(nranks), ctypes.byref(self._h)))
        self.n, self.rank, self.nranks = n, rank, nranks
        self._keep = []

    def _as_fn(self, f):
        n = self.n

        def cb(_ctx, d_x, d_y, _stream):
            try:
                y = _wrap(d_y, n)
                y.copy_(f(_wrap(d_x, n)))
                return 0
            except Exception:  # pragma: no cover
                import traceback

                traceback.print_exc()
                return 1

        c = _APPLY(cb)
        self._keep.append(c)
        return c, None

    def set_operators(self, op, pc=None, stokes_block=None):
        L = lib()
        if isinstance(op, Elliptic):
            fn, ctx = ctypes.cast(L.sb200_apply_elliptic_matmult, ctypes.c_void_p), op._h
        elif isinstance(op, Stokes):
            name = "sb200_apply_stokes_matmult_vv" if stokes_block == "vv" else "sb200_apply_stokes_matmult"
            fn, ctx = ctypes.cast(getattr(L, name), ctypes.c_void_p), op._h
        else:
            fn, ctx = self._as_fn(op)
        self._keep.append(op)
        if isinstance(pc, StokesSaddle):  # StokesPCApply + null-space removal natively on the device vectors
            pfn, pctx = pc.as_ksp_pc()
            self._keep.append(pc)
        else:
            pfn, pctx = (None, None) if pc is None else self._as_fn(pc)
        _ck(L.sb200_ksp_set_operators(self._h, fn, ctx, pfn, pctx))

    def set_tolerances(self, rtol=1e-5, atol=1e-50, dtol=1e5, maxits=10000):
        _ck(lib().sb200_ksp_set_tolerances(self._h, ctypes.c_double(rtol), ctypes.c_double(atol), ctypes.c_double(dtol), ctypes.c_int(maxits)))

    def set_lookahead(self, depth):
        """1: enqueue Arnoldi step k+1 before reading the norm of step k (sb200_ksp_set_lookahead; same iterates and counts)."""
        _ck(lib().sb200_ksp_set_lookahead(self._h, ctypes.c_int(int(depth))))

    def solve(self, b, x=None, guess_nonzero=False):
        import torch

        if x is None:
            x = torch.zeros_like(b)
        assert b.numel() == self.n and x.numel() == self.n
        _ck(lib().sb200_ksp_solve(self._h, _ptr(b), _ptr(x), ctypes.c_int(int(guess_nonzero)), _stream()))
        return x

    @property
    def result(self):
        its, reason = ctypes.c_int(), ctypes.c_int()
        rn, bn = ctypes.c_double(), ctypes.c_double()
        _ck(lib().sb200_ksp_get_result(self._h, ctypes.byref(its), ctypes.byref(rn), ctypes.byref(bn), ctypes.byref(reason)))
        return {"its": its.value, "rnorm": rn.value, "bnorm": bn.value, "reason": reason.value}

    @property
    def history(self):
        n = ctypes.c_int()
        _ck(lib().sb200_ksp_get_history(self._h, None, 0, ctypes.byref(n)))
        buf = np.empty(max(n.value, 1))
        _ck(lib().sb200_ksp_get_history(self._h, _hptr(buf), ctypes.c_int(n.value), ctypes.byref(n)))
        return buf[:n.value]

    @property
    def times_ms(self):
        a, b, c = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        _ck(lib().sb200_ksp_get_times(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return {"operator": a.value, "pc": b.value, "ksp_vector_work": c.value}

    def ipc_export(self):
        buf = ctypes.create_string_buffer(lib().sb200_ipc_handle_bytes())
        _ck(lib().sb200_ksp_ipc_export(self._h, buf))
        return buf.raw

    def ipc_attach(self, peer_rank, handle):
        _ck(lib().sb200_ksp_ipc_attach(self._h, ctypes.c_int(peer_rank), ctypes.c_char_p(handle)))

    def attach_local(self, peer_rank, peer):
        _ck(lib().sb200_ksp_attach_local(self._h, ctypes.c_int(peer_rank), peer._h))

    def destroy(self):
        if self._h:
            _ck(lib().sb200_ksp_destroy(self._h))
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class StokesSaddle:
    """StokesPCApply0..3 (stokes.C:1714-1817) composed on the device (sb200_saddle_*, host/saddle.cpp): the PV / VP / VV shells,
    KSPVelocity / KSPSchur / KSPSchurVelocity as device GMRES, the index scatters and the null-space projection.
    velocity_pc / svel_pc: callables z = M^-1 r on torch tensors of gv doubles (None = PCNONE); svel_pc defaults to velocity_pc."""

    def __init__(self, stokes, saddle_type=0, velocity_pc=None, svel_pc="same", vel_rtol=1e-5, vel_max_it=10000, schur_rtol=1e-5, schur_max_it=10000,
                 svel_preonly=False, svel_rtol=1e-5, svel_max_it=10000):
        self._h = ctypes.c_void_p()
        self.stokes = stokes
        _ck(lib().sb200_saddle_create(stokes._h, ctypes.c_int(saddle_type), ctypes.byref(self._h)))
        _ck(lib().sb200_saddle_set_inner(self._h, ctypes.c_double(vel_rtol), ctypes.c_int(vel_max_it), ctypes.c_double(schur_rtol), ctypes.c_int(schur_max_it),
                                         ctypes.c_int(int(svel_preonly))))
        _ck(lib().sb200_saddle_set_svel(self._h, ctypes.c_double(svel_rtol), ctypes.c_int(svel_max_it)))
        self._keep = []
        fv = self._as_fn(velocity_pc)
        fs = fv if svel_pc == "same" else self._as_fn(svel_pc)
        _ck(lib().sb200_saddle_set_velocity_pc(self._h, fv, None, fs, None, ctypes.c_int(0)))
        # slab-partitioned Stokes context: the inner solvers are slab solvers; create them now so that their arenas can be exchanged
        # (spectral_petsc_b200.dist.attach_peers(pc) / attach_in_process([...])) before the first - collective - apply
        self.rank, self.nranks = getattr(stokes, "rank", 0), getattr(stokes, "nranks", 1)
        if self.nranks > 1:
            _ck(lib().sb200_saddle_prepare(self._h))

    HANDLE_BYTES = 192

    def ipc_export(self):
        buf = ctypes.create_string_buffer(self.HANDLE_BYTES)
        _ck(lib().sb200_saddle_ipc_export(self._h, buf))
        return buf.raw

    def ipc_attach(self, peer_rank, handle):
        _ck(lib().sb200_saddle_ipc_attach(self._h, ctypes.c_int(peer_rank), ctypes.c_char_p(handle)))

    def attach_local(self, peer_rank, peer):
        _ck(lib().sb200_saddle_attach_local(self._h, ctypes.c_int(peer_rank), peer._h))

    def _as_fn(self, f):
        if f is None:
            return None
        n = self.stokes.gv

        def cb(_ctx, d_x, d_y, _stream):
            try:
                _wrap(d_y, n).copy_(f(_wrap(d_x, n)))
                return 0
            except Exception:  # pragma: no cover
                import traceback

                traceback.print_exc()
                return 1

        c = _APPLY(cb)
        self._keep.append(c)
        return c

    def apply(self, x, y=None, remove_constant_pressure=False):
        """y = StokesPCApply{type}(x); remove_constant_pressure adds the projection KSPSetNullSpace applies on the outer KSP."""
        import torch

        if y is None:
            y = torch.empty_like(x)
        assert x.numel() == self.stokes.g and y.numel() == self.stokes.g
        fn = lib().sb200_apply_saddle if remove_constant_pressure else lib().sb200_saddle_apply
        _ck(fn(self._h, _ptr(x), _ptr(y), _stream()))
        return y

    def as_ksp_pc(self):
        """(function pointer, context) of sb200_apply_saddle for sb200_ksp_set_operators: no Python in the outer iteration."""
        return ctypes.cast(lib().sb200_apply_saddle, ctypes.c_void_p), self._h

    @property
    def inner_its(self):
        a, b = ctypes.c_longlong(), ctypes.c_longlong()
        _ck(lib().sb200_saddle_get_inner_its(self._h, ctypes.byref(a), ctypes.byref(b)))
        return {"velocity": a.value, "schur": b.value}

    def destroy(self):
        if self._h:
            _ck(lib().sb200_saddle_destroy(self._h))
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def vec_split(x, d):
    """scatterGV / scatterGP (stokes.C:867-877) on device tensors."""
    import torch

    nodes = x.numel() // (d + 1)
    v = torch.empty(nodes * d, dtype=torch.float64, device=x.device)
    p = torch.empty(nodes, dtype=torch.float64, device=x.device)
    _ck(lib().sb200_vec_split(ctypes.c_longlong(nodes), ctypes.c_int(d), _ptr(x), _ptr(v), _ptr(p), _stream()))
    return v, p


def vec_merge(v, p, d):
    import torch

    nodes = p.numel()
    x = torch.empty(nodes * (d + 1), dtype=torch.float64, device=p.device)
    _ck(lib().sb200_vec_merge(ctypes.c_longlong(nodes), ctypes.c_int(d), _ptr(v), _ptr(p), _ptr(x), _stream()))
    return x


def vec_axpby(a, x, b, y):
    """y <- a x + b y in place (VecAXPBY)."""
    _ck(lib().sb200_vec_axpby(ctypes.c_longlong(y.numel()), ctypes.c_double(a), None if x is None else _ptr(x), ctypes.c_double(b), _ptr(y), _stream()))
    return y


def vec_pointwise_divide(x, diag):
    import torch

    y = torch.empty_like(x)
    _ck(lib().sb200_vec_pointwise_divide(ctypes.c_longlong(x.numel()), _ptr(x), _ptr(diag), _ptr(y), _stream()))
    return y


def csr_diagonal(rowptr, colidx, vals):
    """MatGetDiagonal of a device CSR matrix (int32 rowptr / colidx, fp64 vals as returned by jacobian_csr / pc_velocity_csr)."""
    import torch

    n = rowptr.numel() - 1
    assert rowptr.dtype == torch.int32 and colidx.dtype == torch.int32 and rowptr.is_contiguous() and colidx.is_contiguous()
    diag = torch.empty(n, dtype=torch.float64, device=vals.device)
    _ck(lib().sb200_csr_diagonal(ctypes.c_longlong(n), ctypes.c_void_p(rowptr.data_ptr()), ctypes.c_void_p(colidx.data_ptr()), _ptr(vals), _ptr(diag), _stream()))
    return diag


def vec_remove_mean(x, stride=1, offset=0):
    """MatNullSpaceRemove with the constant vector on x[offset::stride], in place."""
    import torch

    scratch = torch.empty(1024, dtype=torch.float64, device=x.device)  # SB200_REDUCE_SCRATCH_DOUBLES
    n = (x.numel() - offset + stride - 1) // stride
    _ck(lib().sb200_vec_remove_mean(ctypes.c_longlong(n), ctypes.c_int(stride), ctypes.c_int(offset), _ptr(x), _ptr(scratch), _stream()))
    return x


def _wrap(ptr, n):
    """View n doubles of device memory owned by the library as a torch tensor (no copy)."""
    import torch

    class _Holder:
        pass

    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(ptr), False), "version": 3}
    return torch.as_tensor(h, device="cuda")
