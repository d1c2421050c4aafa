"""Launch the fused step kernel a few times at the BASELINE config-2 shape (for ncu captures).

    ncu --set full --clock-control none --import-source on -k regex:smc_step -s 4 -c 2 \
        -o gpurun_out/step_exact python scripts/profile_step.py --mode exact
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aesmc_b200 import _lib, _ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="exact")
ap.add_argument("--batch", type=int, default=4096)
ap.add_argument("--particles", type=int, default=4096)
ap.add_argument("--dim", type=int, default=1)
ap.add_argument("--launches", type=int, default=6)
ap.add_argument("--scale", type=float, default=1.0, help="spread of the synthetic log-weights (8: collapsed weights)")
args = ap.parse_args()
dev = torch.device("cuda", 0)
B, K, D = args.batch, args.particles, args.dim
gen = torch.Generator(device=dev).manual_seed(0)
sets = [[args.scale * torch.randn(B, K, device=dev, generator=gen) - 1.4 for _ in range(3)] for _ in range(2)]
x = [torch.randn(B, K, D, device=dev, generator=gen), torch.empty(B, K, D, device=dev)]
u = torch.rand(B, dtype=torch.float64, device=dev, generator=gen)
log_w = torch.empty(B, K, device=dev)
lse = torch.empty(B, device=dev)
idx = torch.empty(B, K, dtype=torch.int32, device=dev)
flags = _ops.new_flags(dev)
ws_bytes = int(_lib.load().aesmc_smc_step_workspace_bytes(B, K))  # > 0: rows larger than one CTA
ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
for i in range(args.launches):
    a, b, c = sets[i & 1]
    _lib.call("aesmc_smc_step_ws_f32", a.data_ptr(), b.data_ptr(), c.data_ptr(), u.data_ptr(), B, K,
              log_w.data_ptr(), lse.data_ptr(), idx.data_ptr(), x[i & 1].data_ptr(), x[(i + 1) & 1].data_ptr(), D,
              flags.data_ptr(), _ops.mode_code(args.mode), ws.data_ptr() if ws_bytes else None, ws_bytes)
torch.cuda.synchronize()
print("flags", int(flags.item()))
