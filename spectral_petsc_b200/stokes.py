"""`python -m spectral_petsc_b200.stokes -dim 20,20,20 -exact 2 ...`: the reference's ./stokes (stokes.C:114-255)."""
import sys

from .drivers import _run, stokes_main

if __name__ == "__main__":
    sys.exit(_run(stokes_main, sys.argv[1:]))
