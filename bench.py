#!/usr/bin/env python
"""bench.py - spectral MatMult throughput of the B200-native Chebyshev collocation operator.

Contract (see DESIGN.md "Measurement"):
  python bench.py --gpus N --steps K --warmup W            -> ONE JSON line (this repo's CUDA path)
  python bench.py --impl reference --gpus N --steps K ...   -> ONE JSON line (CPU restatement of the
                                                               reference's FFT path, all host cores)
A "step" is one MatMult_Elliptic application (elliptic.C:297-339) on the 128^3 grid with a
variable-coefficient state (eta, deta, gradu populated by one FormFunction call, gamma=4, exponent=2),
on N(0,1) synthetic input (numpy default_rng(0)), SURVEY.md 8(d).
metric = "spectral MatMult GDOF/s (fp64)", N_dof = fields * prod(dim) = 2,097,152.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "spectral MatMult GDOF/s (fp64)"
UNIT = "GDOF/s"
DIM = [128, 128, 128]
GAMMA, EXPONENT = 4.0, 2.0
FP64_PEAK_TFLOPS = 37.1  # measured on this pool's B200 with tools/fp64_peak.cu (profiles/r01_fp64_peak.jsonl)


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def alg_flops(dim):
    # 2*P flop per node per scalar axis-derivative, 2 derivatives per axis
    m = int(np.prod(dim))
    return sum(2 * (2 * p) for p in dim) * m


def alg_bytes(dim):
    # variable-coefficient MatMult_Elliptic = 8*(2g + (2+d)m)
    m = int(np.prod(dim))
    g = int(np.prod([p - 2 for p in dim]))
    return 8 * (2 * g + (2 + len(dim)) * m)


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, reasons = [], [], set()
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        if sm:
            # median over the samples taken under load (upper half of the observed clocks)
            hi = sorted(sm)[len(sm) // 2:]
            out["sm_mhz"] = statistics.median(hi)
            out["sm_max_mhz"] = max(smax)
            out["samples"] = len(sm)
        out["reasons"] = sorted(reasons)
        return out


def build_state_oracle(workers):
    """CPU restatement with the benchmark's synthetic state (used by cpu_baseline / --impl reference)."""
    from oracle.elliptic import MatElliptic

    O = MatElliptic(DIM, gamma=GAMMA, exponent=EXPONENT, workers=workers)
    Us = 0.1 * np.random.default_rng(1).standard_normal(O.g)
    O.form_function(Us)
    U = np.random.default_rng(0).standard_normal(O.g)
    return O, U


def cpu_baseline(sample_steps=3):
    cores = os.cpu_count() or 1
    O, U = build_state_oracle(workers=cores)
    O.mat_mult(U)
    t0 = time.perf_counter()
    for _ in range(sample_steps):
        O.mat_mult(U)
    dt = time.perf_counter() - t0
    return {"value": O.m * sample_steps / dt / 1e9, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d full MatMult_Elliptic applications at 128^3 by the numpy/scipy(pocketfft) restatement of the reference's FFT path "
                      "(FFTW/PETSc unavailable), scipy.fft workers=%d" % (sample_steps, cores),
            "ms_per_step": dt / sample_steps * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    O, U = build_state_oracle(workers=cores)
    for _ in range(min(args.warmup, 3)):
        O.mat_mult(U)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.mat_mult(U)
    dt = time.perf_counter() - t0
    val = O.m * args.steps / dt / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 3),
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": "elliptic 3D -dim 128,128,128 MatMult_Elliptic, variable coefficients (gamma=4, exponent=2)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "each step = one full MatMult_Elliptic at 128^3 by the numpy/scipy(pocketfft) restatement of the reference's FFT path; "
                                   "the reference binary itself cannot be built here (no FFTW/PETSc/MPI)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def run_cuda(args):
    import torch
    import torch.distributed as dist

    import spectral_petsc_b200 as sp

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    G = sp.Elliptic(DIM, gamma=GAMMA, exponent=EXPONENT)
    if args.path is not None:
        G.set_path(args.path)
    Us = torch.from_numpy(0.1 * np.random.default_rng(1).standard_normal(G.g)).to(dev)
    G.form_function(Us)  # populates eta / deta / gradu
    U_host = torch.from_numpy(np.random.default_rng(0).standard_normal(G.g)).pin_memory()
    V_host = torch.empty(G.g, dtype=torch.float64).pin_memory()
    U = U_host.to(dev)
    V = torch.empty_like(U)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # 256 MiB > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        G.mat_mult(U, V)
    barrier()

    # ---- device-resident timing: per-step CUDA events, L2 flushed between steps ----------
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    l0 = sp.launch_count()
    barrier()
    for a, b in ev:
        flush.zero_()
        a.record()
        G.mat_mult(U, V)
        b.record()
    barrier()
    launches = sp.launch_count() - l0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = sum(step_ms)
    # back-to-back (L2-warm) figure for context
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        G.mat_mult(U, V)
    e1.record()
    barrier()
    hot_ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None

    # ---- end-to-end through the host-buffer C-ABI call (H2D + op + D2H every step) ----------
    Uh, Vh = U_host.numpy(), V_host.numpy()
    lib = sp.lib()
    import ctypes
    hp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    for _ in range(2):
        lib.sb200_elliptic_matmult_host(G._h, hp(Uh), hp(Vh))
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rc = lib.sb200_elliptic_matmult_host(G._h, hp(Uh), hp(Vh))
        assert rc == 0
    barrier()
    e2e_s = time.perf_counter() - t0

    t = torch.tensor([total_ms, hot_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, hot_ms, e2e_ms = t.tolist()

    if rank == 0:
        ndof = G.m * world  # replicas until the slab partition lands: every rank applies the full operator
        value = ndof * args.steps / (total_ms * 1e-3) / 1e9
        fl = alg_flops(DIM)
        achieved = fl * args.steps / (total_ms * 1e-3) / 1e12
        peaks = measured_peaks()
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak" if world > 1 else "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "elliptic 3D -dim 128,128,128 MatMult_Elliptic, variable coefficients (gamma=4, exponent=2)",
                       "n_dof_per_step": G.m, "global_vec_len": G.g, "l2": "256 MiB flush between timed steps (per-step CUDA events, flush untimed)",
                       "value_l2_warm": ndof * args.steps / (hot_ms * 1e-3) / 1e9,
                       "parallelism": "replicas" if world > 1 else "single"},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": achieved / FP64_PEAK_TFLOPS,
                         "traffic": None, "pipe": "fp64 DMMA", "peak_source": "tools/fp64_peak.cu on this pool (profiles/r01_fp64_peak.jsonl); MEASURED_PEAKS.json has no fp64 entry",
                         "algorithmic_flops_per_step": fl, "kernel": "whole MatMult step (all launches)",
                         "hbm": {"achieved": alg_bytes(DIM) * args.steps / (total_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": alg_bytes(DIM) * args.steps / (total_ms * 1e-3) / 1e9 / hbm_peak, "algorithmic_bytes_per_step": alg_bytes(DIM)}},
            "e2e": {"value": ndof * args.steps / (e2e_ms * 1e-3) / 1e9, "unit": UNIT, "h2d_bytes_per_step": G.g * 8, "d2h_bytes_per_step": G.g * 8,
                    "api": "sb200_elliptic_matmult_host (pinned host buffers)"},
            "gpu_launches": launches,
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--path", type=int, default=None, help="kernel path override (1 generic, 2 fused)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_cuda(args)


if __name__ == "__main__":
    sys.exit(main())
