import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import spectral_petsc_b200 as sp
from oracle.fgmres import fgmres
cuda = torch.device("cuda:0")
rng = np.random.default_rng(0)
n = 300
A = rng.standard_normal((n, n)) + 2.0 * n ** 0.5 * np.eye(n)
b = rng.standard_normal(n)
Ad = torch.from_numpy(A).to(cuda)
xo, its_o, hist_o, reason_o = fgmres(lambda v: A @ v, b, restart=30, rtol=1e-10)
print("oracle", its_o, reason_o, hist_o[:6], hist_o[-3:])
K = sp.KSP(n, restart=30)
calls = [0]
def op(v):
    calls[0] += 1
    return Ad @ v
K.set_operators(op)
K.set_tolerances(rtol=1e-10, maxits=100)
x = K.solve(torch.from_numpy(b).to(cuda)).cpu().numpy()
print("gpu", K.result, K.history[:6], K.history[-3:], "calls", calls[0])
print("true res", np.linalg.norm(A @ x - b) / np.linalg.norm(b))
