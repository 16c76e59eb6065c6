"""Debug: per-item phase breakdown of the persistent MatMult kernel (needs `make EXTRA=-DSB200_TRACE`)."""
import ctypes, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spectral_petsc_b200 as sp
P = 128
G = sp.Elliptic([P] * 3, gamma=4.0, exponent=2.0)
G.form_function(torch.from_numpy(0.1 * np.random.default_rng(1).standard_normal(G.g)).cuda())
U = torch.randn(G.g, dtype=torch.float64, device="cuda")
V = torch.empty_like(U)
nitems = 3 * (P * P // 8)
buf = torch.zeros(nitems * 8, dtype=torch.int64, device="cuda")
for _ in range(3):
    G.mat_mult(U, V)
sp.lib().sb200_elliptic_debug_trace(G._h, ctypes.c_void_p(buf.data_ptr()))
G.mat_mult(U, V)
torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(nitems, 8)
t0 = t[:, 0][t[:, 0] > 0].min()
names = ["wait_load", "gemm1", "flux", "gemm2", "epilogue"]
for ax in range(3):
    a = t[ax * (nitems // 3):(ax + 1) * (nitems // 3)]
    d = np.diff(a[:, :6], axis=1)
    print("axis", ax, "start min/max us", (a[:, 0].min() - t0) / 1965., (a[:, 0].max() - t0) / 1965., "end max us", (a[:, 5].max() - t0) / 1965.)
    for k, n in enumerate(names):
        print("   %-10s mean %7.0f clk  p10 %7.0f  p90 %7.0f" % (n, d[:, k].mean(), np.percentile(d[:, k], 10), np.percentile(d[:, k], 90)))
    print("   total      mean %7.0f clk" % (a[:, 5] - a[:, 0]).mean())
# one SM's timeline
sm0 = t[t[:, 6] == t[0, 6]]
order = np.argsort(sm0[:, 0])
print("items on SM", t[0, 6], len(sm0))
for r in sm0[order][:40]:
    print(" warp %2d  start %7.1f us  " % (r[7], (r[0] - t0) / 1965.) + " ".join("%6d" % x for x in np.diff(r[:6])))
