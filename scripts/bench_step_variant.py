"""Time the fused step kernel of ONE build of the library (AESMC_B200_LIB selects it) at the BASELINE config-2
shape, the way bench.py does (ring of input sets larger than L2, CUDA events around each launch), and check
64 rows of the result against the CPU oracle.  One JSON line.

    AESMC_B200_LIB=build/variants/libaesmc_v3.so python scripts/bench_step_variant.py --label v3
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aesmc_b200 import _lib, _ops  # noqa: E402
from oracle import core as oracle  # noqa: E402  (checker only)

ap = argparse.ArgumentParser()
ap.add_argument("--label", default="")
ap.add_argument("--mode", default="exact")
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--particles", type=int, default=4096)
ap.add_argument("--launches", type=int, default=60)
ap.add_argument("--scales", default="1,8")
args = ap.parse_args()
dev = torch.device("cuda", 0)
B, K = args.batch, args.particles
gen = torch.Generator(device=dev).manual_seed(0)
out = {"label": args.label, "lib": os.path.basename(_lib.LIB_PATH), "mode": args.mode, "B": B, "K": K}
for scale in [float(s) for s in args.scales.split(",")]:
    ring = [[scale * torch.randn(B, K, device=dev, generator=gen) - 1.4 for _ in range(3)] for _ in range(4)]
    x = [torch.randn(B, K, device=dev, generator=gen), torch.empty(B, K, device=dev)]
    u = torch.rand(B, dtype=torch.float64, device=dev, generator=gen)
    log_w, lse = torch.empty(B, K, device=dev), torch.empty(B, device=dev)
    idx = torch.empty(B, K, dtype=torch.int32, device=dev)
    flags = _ops.new_flags(dev)
    code = _ops.mode_code(args.mode)

    def launch(i):
        a, b, c = ring[i % 4]
        _lib.call("aesmc_smc_step_f32", _lib.ptr(a), _lib.ptr(b), _lib.ptr(c), _lib.ptr(u), B, K, _lib.ptr(log_w),
                  _lib.ptr(lse), _lib.ptr(idx), _lib.ptr(x[i & 1]), _lib.ptr(x[(i + 1) & 1]), 1, _lib.ptr(flags), code)

    for i in range(8):
        launch(i)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.launches + 1)]
    for i in range(args.launches):
        ev[i].record()
        launch(i)
    ev[-1].record()
    torch.cuda.synchronize()
    d = sorted(ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(args.launches))
    out["us_scale_%g" % scale] = round(sum(d) / len(d), 2)
    out["us_min_scale_%g" % scale] = round(d[0], 2)
    # parity of the last launch on 64 rows
    i = args.launches - 1
    a, b, c = [t[:64].cpu().numpy() for t in ring[i % 4]]
    lw_ref = oracle.log_weight(a, b, c)
    idx_ref, st, lse_ref, _, _ = oracle.sample_ancestral_index(lw_ref, u[:64].cpu().numpy(), return_parts=True)
    got = idx[:64].cpu().numpy()
    out["mismatch_scale_%g" % scale] = int((got != np.minimum(idx_ref, K - 1)).sum())
    out["lse_bits_differ_scale_%g" % scale] = int((lse[:64].cpu().numpy().view(np.int32) != lse_ref.view(np.int32)).sum())
    out["gather_ok_scale_%g" % scale] = bool(torch.equal(x[(i + 1) & 1][:64], torch.gather(x[i & 1][:64], 1, idx[:64].long())))
    out["flags"] = int(flags.item())
    del ring, x, log_w, idx
    torch.cuda.empty_cache()
out["frac_of_6550_scale_1"] = round(28.0 * B * K / (out["us_scale_1"] * 1e-6) / 1e9 / 6550.1, 4) if "us_scale_1" in out else None
print(json.dumps(out), flush=True)
