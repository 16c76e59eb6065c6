"""Per-rank timeline of one slab MatMult (torchrun, SB200_XFLAGS=64): prints milestone offsets in us."""
import ctypes, json, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import spectral_petsc_b200 as sp
from spectral_petsc_b200 import dist as spd

NAMES = ["A_cta_start", "stage_pushed", "stage_fenced", "pencil_ready_seen", "DONE_raised", "B_cta_start", "B_wait_done_begin", "B_wait_done_end", "A_exit", "B_exit"]
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
P = int(sys.argv[1]) if len(sys.argv) > 1 else 128
G = sp.Elliptic([P] * 3, gamma=4.0, exponent=2.0, rank=rank, nranks=world)
spd.attach_peers(G)
U = torch.from_numpy(np.random.default_rng(0).standard_normal(G.g)).to(dev); V = torch.empty_like(U)
G.form_function(0.1 * U)
for _ in range(20): G.mat_mult(U, V)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 20)()
sp.lib().sb200_elliptic_debug_timeline(G._h, buf, None)
TL = int(os.environ.get("SB200_TL_EPOCH", "35"))
for _ in range(20): G.mat_mult(U, V)   # epochs 21..40; the kernels stamp epoch TL only
torch.cuda.synchronize()
sp.lib().sb200_elliptic_debug_timeline(G._h, buf, None)
t = np.array(list(buf), dtype=np.float64).reshape(10, 2)
t0 = t[0, 0]
print(json.dumps({"rank": rank, "epoch": TL, "us_min_max": {n: [round((a - t0) / 1e3, 1), round((b - t0) / 1e3, 1)] for n, (a, b) in zip(NAMES, t)}}), flush=True)
dist.barrier()
dist.destroy_process_group()
