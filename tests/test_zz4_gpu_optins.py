"""(Runs last of all in the GPU suite.)  The opt-in evaluation paths written after this round's GPU budget was spent, each against
the default path it replaces: sb200_stokes_set_trace_divergence / _set_fold_pressure (StokesMatMult / StokesFunction from one
gradient and one divergence) and path 4 of MatMult_Elliptic (the generic launches replayed from a CUDA graph).  All are OFF by
default; these tests are what has to be green before they are switched on."""
import numpy as np
import pytest
import torch

import spectral_petsc_b200 as sp
from test_zz1_gpu_saddle import _state

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dim,rheology", [([16, 16, 16], 1), ([32, 32, 32], 1), ([9, 7, 6], 1), ([8, 6], 0), ([20, 20, 20], 0)], ids=str)
def test_trace_divergence_option_gives_the_same_operator(cuda, dim, rheology):
    """Opt-in sb200_stokes_set_trace_divergence: the pressure rows of StokesMatMult / StokesFunction taken from the trace of the
    velocity gradient the viscous part computes, instead of a second StokesDivergence pass on the same input (stokes.C:509,746)."""
    S = _state(cuda, dim, rheology)
    S.set_trace_divergence(False)  # (both switches are on by default since round 2: start from the literal three-shell sequence)
    S.set_fold_pressure(False)
    d = len(dim)
    rng = np.random.default_rng(3)
    x = torch.from_numpy(rng.standard_normal(S.g)).to(cuda)
    xs = torch.from_numpy(0.3 * rng.standard_normal(S.g)).to(cuda)
    F0 = S.function(xs).clone()
    l0 = sp.launch_count()
    y0 = S.mat_mult(x).clone()
    n_off = sp.launch_count() - l0
    S.set_trace_divergence(True)
    F1 = S.function(xs).clone()
    l0 = sp.launch_count()
    y1 = S.mat_mult(x).clone()
    n_on = sp.launch_count() - l0
    S.set_trace_divergence(False)
    for a, b in ((y1, y0), (F1, F0)):
        assert float((a - b).abs().max()) <= 1e-13 * float(b.abs().max())
        assert torch.equal(a.reshape(-1, d + 1)[:, :d], b.reshape(-1, d + 1)[:, :d])  # the velocity rows do not change path
    if n_off:  # (0 over the CPU test double of the dry run)
        assert n_on < n_off
    print("trace-divergence %s: launches %d -> %d, pressure rows bitwise equal: %s" % (dim, n_off, n_on, torch.equal(y1, y0) and torch.equal(F1, F0)))
    # opt-in 2, alone and with the first: the pressure gradient out of the viscous divergence (flux = eta*eps - p I); the sum of the
    # two terms is rounded once instead of twice, so the bar is the parity bar of the operator (1e-12), not bit equality
    for trace in (False, True):
        S.set_trace_divergence(trace)
        S.set_fold_pressure(True)
        F2 = S.function(xs).clone()
        l0 = sp.launch_count()
        y2 = S.mat_mult(x).clone()
        n_fold = sp.launch_count() - l0
        S.set_fold_pressure(False)
        S.set_trace_divergence(False)
        for a, b in ((y2, y0), (F2, F0)):
            assert float((a - b).abs().max()) <= 1e-12 * float(b.abs().max())
        if n_off:
            assert n_fold < n_off
        print("fold-pressure (trace %s) %s: launches %d -> %d, max rel diff %.2e" % (trace, dim, n_off, n_fold, float((y2 - y0).abs().max() / y0.abs().max())))
    # the switches leave no state behind
    assert torch.equal(S.mat_mult(x), y0) and torch.equal(S.function(xs), F0)
    # the default (both on) is the combination tested last above
    S.set_trace_divergence(True)
    S.set_fold_pressure(True)
    assert torch.equal(S.mat_mult(x), y2)
    S.destroy()


@pytest.mark.parametrize("dim", [[16, 16, 16], [12, 12, 12, 12, 12], [20, 17, 9], [24, 24]], ids=str)
def test_graph_captured_generic_path_equals_generic_path(cuda, dim):
    """Opt-in path 4 of MatMult_Elliptic: the generic path's launches captured once into a CUDA graph and replayed - the same
    kernels on the same data, so the same bits; a new FormFunction state is seen without re-capture."""
    E = sp.Elliptic(dim, gamma=4.0, exponent=2.0)
    rng = np.random.default_rng(2)
    U = torch.from_numpy(rng.standard_normal(E.g)).to(cuda)
    for trial in range(3):
        E.form_function(torch.from_numpy(0.1 * rng.standard_normal(E.g)).to(cuda))  # new eta / deta / gradu each time
        E.set_path(1)
        l0 = sp.launch_count()
        ref = E.mat_mult(U).clone()
        n_generic = sp.launch_count() - l0
        E.set_path(4)
        out = E.mat_mult(U)
        assert torch.equal(out, ref)
        l0 = sp.launch_count()
        out2 = E.mat_mult(U)
        assert torch.equal(out2, ref)
        if n_generic:
            assert sp.launch_count() - l0 == n_generic  # the graph replays as many kernels as the generic path launches
    E.set_path(0)
    E.destroy()


@pytest.mark.parametrize("dim,rheology", [([20, 20, 20], 1), ([16, 16, 16], 1), ([8, 6], 0)], ids=str)
def test_graph_replayed_stokes_shells_equal_the_launched_ones(cuda, dim, rheology):
    """Opt-in sb200_stokes_set_graph: StokesMatMult / VV / PV / VP replayed from CUDA graphs (BASELINE config 4's 20^3 grid is
    launch-bound) - the same kernels on the same data, so the same bits, also after a new StokesFunction state and with the
    evaluation switches on; the saddle-point PC composed over graph-replayed shells gives the same vector."""
    S = _state(cuda, dim, rheology)
    rng = np.random.default_rng(4)
    x = torch.from_numpy(rng.standard_normal(S.g)).to(cuda)
    v = torch.from_numpy(rng.standard_normal(S.gv)).to(cuda)
    p = torch.from_numpy(rng.standard_normal(S.gp)).to(cuda)
    shells = lambda: [S.mat_mult(x).clone(), S.mat_mult_vv(v).clone(), S.mat_mult_pv(v).clone(), S.mat_mult_vp(p).clone()]
    for trial in range(2):
        S.function(torch.from_numpy(0.3 * rng.standard_normal(S.g)).to(cuda))  # a new eta / deta / strain: no re-capture needed
        for switches in (False, True):
            S.set_trace_divergence(switches)
            S.set_fold_pressure(switches)
            S.set_graph(False)
            ref = shells()
            S.set_graph(True)
            for a, b in zip(shells(), ref):
                assert torch.equal(a, b)
            for a, b in zip(shells(), ref):  # replay
                assert torch.equal(a, b)
    S.set_trace_divergence(True)
    S.set_fold_pressure(True)
    S.set_graph(False)
    pc = sp.StokesSaddle(S, 0, velocity_pc=None, vel_max_it=4, schur_max_it=3, svel_preonly=True)
    y_ref = pc.apply(x).clone()
    S.set_graph(True)
    assert torch.equal(pc.apply(x), y_ref)
    S.set_graph(False)
    pc.destroy()
    S.destroy()
