"""CPU oracle for the matrix-free Chebyshev collocation path of spectral-petsc.

TEST INFRASTRUCTURE ONLY.  This package is a numpy/scipy restatement of the
reference's algorithm (chebyshev.c, elliptic.C, stokes.C, util.C).  It exists so
that ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs have something to check (and time) the CUDA path
against.  Nothing under ``spectral_petsc_b200/`` imports it; the product path
fails loudly when the CUDA library is missing and has no CPU fallback.

Parity status: PINNED AGAINST THE REFERENCE'S OWN SOURCE, with stand-ins for its two
external libraries.  The reference's build (FFTW + PETSc ~3.0 + mpicxx + CppAD) is not
available here, but its source files compile UNMODIFIED, where they lie under
/root/reference, against minimal stand-in headers (oracle/ref_stubs/: sequential
Vec / IS / VecScatter / MatShell with PETSc's documented semantics, FFTW's REDFT00 /
RODFT00 from their definitions, inert solver objects, an inert CppAD); recipe
oracle/Makefile, outputs oracle/_ref/lib{cheb,elliptic,stokes}ref.so, wrapper
oracle/ref.py.  tests/test_oracle_ref*.py compare this restatement with the
reference source on identical inputs (ChebMult, MatMult_Elliptic, FormFunction,
SetupBC, CreateExactSolution, FormJacobian, StokesFunction, StokesMatMult{,VV,PV,VP,
Schur}, StokesPressureReduceOrder, StokesPCSetUp0: 1e-12 or better), check the committed
golden vectors against it, and compare the CUDA path with it directly.  What the
stand-ins cannot pin: FFTW's own rounding (its transforms are replaced by an O(n^2)
long-double evaluation of the same definitions) and PETSc's solvers (KSP/SNES, restated
in oracle/fgmres.py from the published algorithm).  The reference's analytic
known-answer checks (cheb.c, -exact 1/2, the constant-pressure null space, util.C's
polyInterp self test) are covered by tests/test_oracle_*.py as before.
"""
