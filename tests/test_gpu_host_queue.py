"""Queued host-buffer MatMult_Elliptic (sb200_elliptic_matmult_host_submit / _wait): every application of a stream of
pinned host vectors equals the blocking sb200_elliptic_matmult_host call bit for bit, in order, whatever the queue depth;
the queue bounds are reported as errors (PETSC_ERR_USER convention), not by blocking or overwriting."""
import numpy as np
import pytest
import torch

import spectral_petsc_b200 as sp

pytestmark = pytest.mark.gpu


def pinned(n):
    return torch.empty(n, dtype=torch.float64).pin_memory()


@pytest.mark.parametrize("dim", [[16, 16, 16], [64, 64, 64], [12] * 5], ids=lambda v: str(v))
def test_stream_of_vectors_equals_blocking_calls(cuda, dim):
    G = sp.Elliptic(dim, gamma=4.0, exponent=2.0)
    rng = np.random.default_rng(3)
    # state written on the default stream right before the queue's first use: the queue must order itself behind it
    G.form_function(torch.from_numpy(0.1 * rng.standard_normal(G.g)).to(cuda))
    nvec = 11  # more than two rounds of the 4-deep queue, not a multiple of it
    keep = [(pinned(G.g), pinned(G.g)) for _ in range(nvec)]
    Us, Vs = [a.numpy() for a, _ in keep], [b.numpy() for _, b in keep]
    for U in Us:
        U[:] = rng.standard_normal(G.g)
    for V in Vs:
        V[:] = np.nan
    G.mat_mult_host_stream(Us, Vs)
    assert G.mat_mult_host_pending() == 0
    for U, V in zip(Us, Vs):
        assert np.array_equal(V, G.mat_mult_host(U))
    # a second pass through the same slots (device vectors reused) gives the same bits
    V2 = [np.full(G.g, np.nan) for _ in range(nvec)]
    G.mat_mult_host_stream(Us, V2)
    for a, b in zip(Vs, V2):
        assert np.array_equal(a, b)


def test_queue_bounds_are_errors(cuda):
    G = sp.Elliptic([8, 8, 8])
    with pytest.raises(sp.SB200Error) as ei:
        G.mat_mult_host_wait()  # nothing submitted
    assert ei.value.code == 83
    keep = [(pinned(G.g), pinned(G.g)) for _ in range(G.HOST_QUEUE_DEPTH + 1)]
    for a, _ in keep:
        a.zero_()
    for a, b in keep[:G.HOST_QUEUE_DEPTH]:
        G.mat_mult_host_submit(a.numpy(), b.numpy())
    assert G.mat_mult_host_pending() == G.HOST_QUEUE_DEPTH
    with pytest.raises(sp.SB200Error) as ei:
        G.mat_mult_host_submit(keep[-1][0].numpy(), keep[-1][1].numpy())  # queue full
    assert ei.value.code == 83
    for _ in range(G.HOST_QUEUE_DEPTH):
        G.mat_mult_host_wait()
    assert G.mat_mult_host_pending() == 0
    for _, b in keep[:G.HOST_QUEUE_DEPTH]:
        assert float(b.abs().max()) == 0.0  # A * 0 = 0 landed in every result buffer
    with pytest.raises(sp.SB200Error) as ei:
        G.mat_mult_host_submit(keep[0][0].numpy(), keep[0][0].numpy())  # x aliases y
    assert ei.value.code == 62
