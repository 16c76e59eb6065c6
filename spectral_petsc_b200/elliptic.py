"""`python -m spectral_petsc_b200.elliptic -dim 16,16,16 -exact 2 -ksp_rtol 1e-10`: the reference's ./elliptic (elliptic.C:116-247)."""
import sys

from .drivers import _run, elliptic_main

if __name__ == "__main__":
    sys.exit(_run(elliptic_main, sys.argv[1:]))
