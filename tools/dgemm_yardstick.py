"""cuBLAS DGEMM yardstick for the FP64 roofline denominator (library call, not the product)."""
import json, torch
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda")
b = torch.randn(n, n, dtype=torch.float64, device="cuda")
for _ in range(2):
    c = a @ b
torch.cuda.synchronize()
best = 1e30
for _ in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(json.dumps({"bench": "cublas_dgemm_8192", "ms": best, "tflops": 2 * n ** 3 / best * 1e-9}))
