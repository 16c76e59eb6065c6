"""spectral_petsc_b200 - B200-native matrix-free Chebyshev collocation operators.

Host-side mirror of the reference's operator interface (chebyshev.h, elliptic.C, stokes.C) over the
C-ABI library ``libspectral_b200.so`` (include/spectral_b200.h).  PyTorch is used only for device
memory, streams and torch.distributed plumbing.  There is no CPU fallback: importing works without
a GPU (so the ABI can be inspected), but every compute call raises if the CUDA library or a CUDA
device is missing.
"""
from .capi import (  # noqa: F401
    SB200Error,
    lib,
    lib_path,
    launch_count,
    Cheb,
    Elliptic,
    Stokes,
    KSP,
    cheb_matrix,
    elliptic_exact_solution,
    stokes_exact_solution,
    HostILU,
    StokesSaddle,
    vec_split,
    vec_merge,
    vec_axpby,
    vec_pointwise_divide,
    vec_remove_mean,
    csr_diagonal,
    cheb_even_odd,
)

__all__ = ["SB200Error", "lib", "lib_path", "launch_count", "Cheb", "Elliptic", "Stokes", "KSP", "cheb_matrix", "elliptic_exact_solution", "stokes_exact_solution", "HostILU", "StokesSaddle", "vec_split", "vec_merge", "vec_axpby", "vec_pointwise_divide", "vec_remove_mean", "csr_diagonal", "cheb_even_odd"]
