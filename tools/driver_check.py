"""Quick GPU run of the two command-line drivers (no oracle side): prints what they report as one JSON line each."""
import json
import sys
import time

sys.path.insert(0, ".")
from spectral_petsc_b200 import drivers  # noqa: E402

t0 = time.time()
L = []
r = drivers.elliptic_main("-dim 16,16,16 -exact 2 -ksp_rtol 1e-10 -pc_type lu".split(), out=L.append)
print(json.dumps({"driver": "elliptic", "snes_its": r["snes_its"], "ksp_its": r["ksp_its"], "reason": r["reason"], "error_abs": r["error_abs"],
                  "exact_residual_abs": r["exact_residual_abs"], "s": time.time() - t0}), flush=True)
L = []
cmd = ("-exact 2 -cont 2 -rheology 1 -eps 1e-2 -exponent 3 -schur_ksp_max_it 3 -vel_ksp_max_it 4 -svel_ksp_type preonly -vel_pc_type lu -svel_pc_type lu -dim 8,8,8 "
       "-ksp_rtol 1e-6 -ksp_max_it 300 -output_vtk gpurun_out/stokes_8.vtk")
r = drivers.stokes_main(cmd.split(), out=L.append)
print(json.dumps({"driver": "stokes", "steps": [(s["snes_its"], s["ksp_its"], s["reason"], s["error"]) for s in r["steps"]], "null_space": r["null_space"],
                  "exact_residual": r["exact_residual"], "s": time.time() - t0}), flush=True)
