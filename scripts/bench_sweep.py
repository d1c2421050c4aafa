"""BASELINE config 5: resampling-bound sweep of the core step, K = 1e3 ... 1e6 particles per row, small
batch (and the headline shape for comparison).  Prints one JSON line per (B, K, mode); `us_per_step` is the
T-step loop replayed as one CUDA graph, `us_per_step_stepwise` the same launches issued one by one from Python:

    python scripts/bench_sweep.py [--steps 20] > profiles/r1_sweep.jsonl

Under torchrun (one rank per GPU) every rank runs its own B rows -- rows are independent, the path has no
collective -- and rank 0 reports the whole job (N * B rows over the slowest rank's time):

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/bench_sweep.py
"""
import argparse
import json
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aesmc_b200 import _lib, _ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=20)   # T
args = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
torch.cuda.set_device(dev)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
peak = 6550.1
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
T = args.steps
for B, K in [(8, 1000), (64, 1000), (8, 10000), (64, 10000), (8, 100000), (64, 100000), (8, 1000000), (64, 1000000),
             (4096, 1024), (1024, 16384), (4096, 4096)]:
    gen = torch.Generator(device=dev).manual_seed(0)
    ring = [[torch.randn(B, K, device=dev, generator=gen) - 1.4 for _ in range(3)] for _ in range(2)]
    arena = [torch.randn(B, K, device=dev, generator=gen), torch.empty(B, K, device=dev)]
    u = torch.rand(T, B, dtype=torch.float64, device=dev, generator=gen)
    log_w = torch.empty(B, K, device=dev)
    lse = torch.empty(B, device=dev)
    idx = torch.empty(B, K, dtype=torch.int32, device=dev)
    flags = _ops.new_flags(dev)
    ws_bytes = int(_lib.load().aesmc_smc_step_workspace_bytes(B, K))
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
    for mode in ("exact", "fast"):
        code = _ops.mode_code(mode)

        def run():
            for t in range(T):
                a, b, c = ring[t & 1]
                _lib.call("aesmc_smc_step_ws_f32", a.data_ptr(), b.data_ptr(), c.data_ptr(), u[t].data_ptr(), B, K,
                          log_w.data_ptr(), lse.data_ptr(), idx.data_ptr(), arena[t & 1].data_ptr(),
                          arena[(t + 1) & 1].data_ptr(), 1, flags.data_ptr(), code, ws.data_ptr() if ws_bytes else None, ws_bytes)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        reps = 3

        def timed(fn):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if world > 1:
                dist.barrier()
                torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            t_ms = e0.elapsed_time(e1) / (reps * T)
            if world > 1:
                tt = torch.tensor([t_ms], device=dev)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                t_ms = float(tt.item())
            return t_ms

        ms_stepwise = timed(run)  # T launches issued from Python (ctypes), as infer() does
        # the same T steps captured once as a CUDA graph: what the kernels cost without the host in the loop
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            run()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            run()
        graph.replay()
        torch.cuda.synchronize()
        ms = timed(graph.replay)
        gbs = 28.0 * B * K / (ms * 1e-3) / 1e9  # per GPU
        line = {"n_gpus": world, "B": B * world, "B_per_gpu": B, "K": K, "T": T, "mode": mode, "us_per_step": round(ms * 1e3, 2),
                "us_per_step_stepwise": round(ms_stepwise * 1e3, 2),
                "particle_steps_per_s": world * B * K / (ms * 1e-3), "algorithmic_GBps_per_gpu": round(gbs, 1),
                "frac_of_measured_hbm": round(gbs / peak, 4), "path": "multi-CTA" if ws_bytes else "single-CTA",
                "flags": int(flags.item())}
        if rank == 0:
            print(json.dumps(line), flush=True)
    del ring, arena, log_w, idx, ws
    torch.cuda.empty_cache()
if world > 1:
    dist.destroy_process_group()
