import os, sys
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import spectral_petsc_b200 as sp
from spectral_petsc_b200 import dist as spd
from oracle.elliptic import MatElliptic
from test_gpu_slab import setup, run_function, run_matmult
cuda = torch.device("cuda:0")
dim = [int(a) for a in sys.argv[1].split(",")]; nr = int(sys.argv[2])
O, R = setup(dim, nr, 4.0, 2.0, cuda)
Us = 0.1 * np.random.default_rng(1).standard_normal(O.g)
parts = spd.split_global(Us, dim, nr)
plane = O.m // dim[0]
w0 = np.zeros(O.m); w0[O.ixG] = Us; w0[O.ixD] = O.dirichlet
for r, c in enumerate(R.ctx):
    L = c.pad(torch.from_numpy(parts[r].copy()).to(cuda), with_dirichlet=True).cpu().numpy()
    sl = slice(c.i0 * plane, (c.i0 + c.nloc) * plane)
    print("rank", r, "pad err", np.abs(L - w0[sl]).max())
F = run_function(O, R, Us); Fo = O.form_function(Us)
print("F err", np.abs(F - Fo).max() / np.abs(Fo).max())
for r, c in enumerate(R.ctx):
    sl = slice(c.i0 * plane, (c.i0 + c.nloc) * plane)
    print("rank", r, "eta err", np.abs(c.get_state(0).cpu().numpy() - O.eta[sl]).max(), [float(np.abs(c.get_state(2 + k).cpu().numpy() - O.gradu[k][sl]).max()) for k in range(O.d)], "timeouts", c.slab_timeouts())
