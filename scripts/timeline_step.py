"""Per-warp clock stamps of the exact row kernel's phases (a build with -DAESMC_X_TIMELINE=1, selected by AESMC_B200_LIB):
cycles since the row's start, averaged over the rows CTA 0 processed, at B = K = 4096."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aesmc_b200 import _lib, _ops  # noqa: E402

dev = torch.device("cuda", 0)
lib = ctypes.CDLL(_lib.LIB_PATH)
B = K = 4096
gen = torch.Generator(device=dev).manual_seed(0)
for scale in (1.0, 8.0):
    a, b, c = [scale * torch.randn(B, K, device=dev, generator=gen) - 1.4 for _ in range(3)]
    x = torch.randn(B, K, device=dev, generator=gen)
    u = torch.rand(B, dtype=torch.float64, device=dev, generator=gen)
    flags = _ops.new_flags(dev)
    for it in range(3):
        if it == 2:
            lib.aesmc_debug_timeline(None, 1)
        _ops.smc_step(a, b, c, u, x, flags, "exact", True)
        torch.cuda.synchronize()
    buf = (ctypes.c_longlong * (32 * 16))()
    lib.aesmc_debug_timeline(buf, 0)
    names = ["P1 done", "bar1", "P2a done", "bar2", "P2b+P3a done", "bar4", "level1 done", "E known", "replay done", "total known",
             "P4 done", "bar8", "scan done", "row done"]
    print("scale", scale, "rows", buf[15], flush=True)
    print("%-14s" % "stage" + "".join("%8s" % ("w%d" % w) for w in range(8)))
    for s_, n in enumerate(names):
        print("%-14s" % n + "".join("%8d" % (buf[16 * w + s_] // max(buf[16 * w + 15], 1)) for w in range(8)))
    print("%-14s" % "general lanes" + "".join("%8d" % buf[16 * w + 14] for w in range(8)))
