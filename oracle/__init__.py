"""CPU oracle for the matrix-free Chebyshev collocation path of spectral-petsc.

TEST INFRASTRUCTURE ONLY.  This package is a numpy/scipy restatement of the
reference's algorithm (chebyshev.c, elliptic.C, stokes.C, util.C).  It exists so
that ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs have something to check (and time) the CUDA path
against.  Nothing under ``spectral_petsc_b200/`` imports it; the product path
fails loudly when the CUDA library is missing and has no CPU fallback.

Parity status: the reference cannot be built here (no FFTW, no PETSc, no MPI,
see DESIGN.md), and its own tests hold no stored golden vectors - only analytic
known-answer checks (cheb.c: d/dx e^x = e^x; the ``-exact`` manufactured
solutions; the constant-pressure null space; util.C's polyInterp self test).
The oracle is pinned against every one of those (tests/test_oracle_*.py).
Beyond that analytic tolerance: PARITY UNPINNED (no bitwise reference output
exists to compare with).  FFTW's REDFT00/RODFT00 are restated through
``scipy.fft.dct/dst(type=1)`` (pocketfft), which implement the same
unnormalised definitions.
"""
