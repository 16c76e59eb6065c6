"""GPU parity of MatMult_Elliptic / FormFunction (C ABI) against the oracle restatement of
elliptic.C on identical inputs; bar 1e-12 max-norm relative on random inputs."""
import numpy as np
import pytest
import torch

import spectral_petsc_b200 as sp
from oracle.elliptic import MatElliptic
from conftest import rel_max

pytestmark = pytest.mark.gpu
TOL = 1e-12


def make_pair(dim, gamma, exponent, cuda, exact=2, cos_scale=None):
    O = MatElliptic(dim, gamma=gamma, exponent=exponent)
    u, u2 = O.create_exact_solution(exact, cos_scale=cos_scale)
    G = sp.Elliptic(dim, gamma=gamma, exponent=exponent)
    assert (G.m, G.g, G.nd) == (O.m, O.g, O.nd)
    G.set_dirichlet(torch.from_numpy(O.dirichlet).to(cuda))
    G.set_rhs(torch.from_numpy(O.b).to(cuda))
    return O, G, u, u2


CASES = [([8, 6], 0.0, 2.0), ([8, 6], 4.0, 2.0), ([7, 6, 5], 4.0, 2.0), ([16, 16, 16], 0.0, 2.0), ([16, 16, 16], 4.0, 2.0),
         ([20, 20, 20], 4.0, 3.0), ([12] * 5, 0.0, 2.0), ([12] * 5, 4.0, 2.0), ([5, 4, 3, 6], 1.5, 2.0), ([33, 9], 4.0, 2.0),
         ([3, 3, 3], 4.0, 2.0), ([64, 64, 64], 4.0, 2.0)]


@pytest.mark.parametrize("dim,gamma,exponent", CASES, ids=lambda v: str(v))
def test_function_and_matmult_match_oracle(cuda, dim, gamma, exponent):
    O, G, u, u2 = make_pair(dim, gamma, exponent, cuda)
    rng0, rng1 = np.random.default_rng(0), np.random.default_rng(1)
    # state: residual at a random positive-ish field (scaled so pow(u,p) stays tame)
    Us = 0.1 * rng1.standard_normal(O.g)
    Fo = O.form_function(Us)
    Fg = G.form_function(torch.from_numpy(Us).to(cuda))
    assert rel_max(Fg.cpu().numpy(), Fo) < TOL
    assert rel_max(G.get_state(0).cpu().numpy(), O.eta) < 1e-14
    assert rel_max(G.get_state(1).cpu().numpy(), O.deta) < 1e-14 or np.abs(O.deta).max() == 0
    for k in range(O.d):
        assert rel_max(G.get_state(2 + k).cpu().numpy(), O.gradu[k]) < TOL
    U = rng0.standard_normal(O.g)
    Vo = O.mat_mult(U)
    Vg = G.mat_mult(torch.from_numpy(U).to(cuda))
    assert rel_max(Vg.cpu().numpy(), Vo) < TOL
    assert np.array_equal(G.mat_mult_host(U), Vg.cpu().numpy())


def test_K3_exact2_residual(cuda):
    for dim in ([16, 16, 16], [12] * 5, [20, 20, 20]):
        O, G, u, u2 = make_pair(dim, 0.0, 2.0, cuda)
        r = G.form_function(torch.from_numpy(u).to(cuda)).cpu().numpy()
        assert np.abs(r).max() < 5e-11  # "Norm of exact residual" (elliptic.C:208)


def test_pad_crop_scatter_semantics(cuda):
    # TEST_SCATTER block of elliptic.C:436-456: G->L, D->L, L->G round trips
    O, G, u, u2 = make_pair([6, 5, 4], 0.0, 2.0, cuda)
    U = np.arange(1, O.g + 1, dtype=np.float64)
    L = G.pad(torch.from_numpy(U).to(cuda), with_dirichlet=True).cpu().numpy()
    ref = np.zeros(O.m)
    ref[O.ixG] = U
    ref[O.ixD] = O.dirichlet
    assert np.array_equal(L, ref)
    L0 = G.pad(torch.from_numpy(U).to(cuda), with_dirichlet=False).cpu().numpy()
    ref[O.ixD] = 0.0
    assert np.array_equal(L0, ref)
    assert np.array_equal(G.crop(torch.from_numpy(ref).to(cuda)).cpu().numpy(), U)


def test_full_size_128(cuda):
    dim = [128, 128, 128]
    O, G, u, u2 = make_pair(dim, 4.0, 2.0, cuda)
    assert (G.m, G.g, G.nd) == (2097152, 2000376, 96776)
    Us = 0.1 * np.random.default_rng(1).standard_normal(O.g)
    Fo = O.form_function(Us)
    Fg = G.form_function(torch.from_numpy(Us).to(cuda))
    assert rel_max(Fg.cpu().numpy(), Fo) < TOL
    U = np.random.default_rng(0).standard_normal(O.g)
    Vo = O.mat_mult(U)
    Vg = G.mat_mult(torch.from_numpy(U).to(cuda))
    assert rel_max(Vg.cpu().numpy(), Vo) < TOL
    # linearity at full size
    a = torch.from_numpy(U).to(cuda)
    b = torch.from_numpy(Us).to(cuda)
    lhs = G.mat_mult(3.0 * a - b)
    rhs = 3.0 * G.mat_mult(a) - G.mat_mult(b)
    assert ((lhs - rhs).abs().max() / rhs.abs().max()).item() < 1e-12


@pytest.mark.parametrize("dim", [[32, 32], [32, 32, 32], [64, 64, 64], [32, 32, 32, 32], [128, 128]], ids=lambda v: str(v))
def test_fused_chain_path_matches_generic_and_oracle(cuda, dim):
    # path 1 = generic per-axis kernels, 2 = one chain kernel per axis, 3 = single persistent chain kernel
    O, G, u, u2 = make_pair(dim, 4.0, 2.0, cuda)
    Us = 0.1 * np.random.default_rng(1).standard_normal(O.g)
    O.form_function(Us)
    G.form_function(torch.from_numpy(Us).to(cuda))
    U = np.random.default_rng(0).standard_normal(O.g)
    Vo = O.mat_mult(U)
    Ud = torch.from_numpy(U).to(cuda)
    G.set_path(1)
    V1 = G.mat_mult(Ud).cpu().numpy()
    G.set_path(2)
    V2 = G.mat_mult(Ud).cpu().numpy()
    G.set_path(3)
    V3 = G.mat_mult(Ud).cpu().numpy()
    V3b = G.mat_mult(Ud).cpu().numpy()  # counters re-armed by the previous launch
    assert rel_max(V1, Vo) < TOL
    assert rel_max(V2, Vo) < TOL
    assert rel_max(V3, Vo) < TOL
    assert rel_max(V2, V1) < 1e-13
    assert np.array_equal(V3, V2)  # same arithmetic, same order
    assert np.array_equal(V3, V3b)


@pytest.mark.parametrize("dim", [[96, 96], [96, 96, 96], [48, 48, 48], [80, 80, 80], [112, 112], [112, 112, 112], [144, 144], [160, 160]], ids=str)
def test_persistent_chain_at_96(cuda, dim):
    """Every extent P % 16 == 0 from 32 to 160 runs the persistent chain kernel by default (P = 96: 6 pair tiles; 48 / 80 / 112 / 144: an odd
    number of tiles; 144 / 160: 9 / 10 tiles); against the oracle and the generic path."""
    O, G, u, u2 = make_pair(dim, 4.0, 2.0, cuda)
    Us = 0.1 * np.random.default_rng(1).standard_normal(O.g)
    O.form_function(Us)
    G.form_function(torch.from_numpy(Us).to(cuda))
    U = np.random.default_rng(0).standard_normal(O.g)
    Vo = O.mat_mult(U)
    Ud = torch.from_numpy(U).to(cuda)
    V0 = G.mat_mult(Ud).cpu().numpy()
    assert "persist" in G.kernel_name()
    V0b = G.mat_mult(Ud).cpu().numpy()
    G.set_path(1)
    V1 = G.mat_mult(Ud).cpu().numpy()
    assert rel_max(V0, Vo) < TOL and rel_max(V1, Vo) < TOL and rel_max(V0, V1) < 1e-13
    assert np.array_equal(V0, V0b)


def test_fused_path_rejects_unsupported_extents(cuda):
    G = sp.Elliptic([16, 16, 16])
    G.set_path(2)
    U = torch.zeros(G.g, dtype=torch.float64, device=cuda)
    with pytest.raises(sp.SB200Error) as ei:
        G.mat_mult(U)
    assert ei.value.code == 56
