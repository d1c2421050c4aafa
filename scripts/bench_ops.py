"""Throughput of the secondary kernels of the path (gather and its backward, step backward, row reductions,
Normal.log_prob) at the headline shape, as algorithmic GB/s against the measured HBM peak.  One JSON line
per op:

    python scripts/bench_ops.py > profiles/r1_ops.jsonl
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aesmc_b200 import _lib, _ops  # noqa: E402

dev = torch.device("cuda", 0)
peak = 6550.1
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
B = K = 4096
gen = torch.Generator(device=dev).manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > L2


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def clocks():
    """SM clock / max clock / throttle reasons right after the measurement (one nvidia-smi query per line)."""
    import subprocess
    try:
        out = subprocess.run(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
                              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                              "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=10).stdout.strip().split(",")
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]),
                "reasons": [n for n, v in zip(names, out[2:6]) if v.strip().lower().startswith("active")]}
    except Exception as exc:  # noqa: BLE001
        return {"error": str(exc)}


def timed_ring(fns, reps=5):
    """Kernels too short to time one launch at a time (a launch costs ~3 us on the device, ~6 us on the host):
    `fns` are the same op on different input tables that together exceed L2 (no flush needed); they are issued
    back to back `reps` times between one pair of events.  Returns us per call."""
    for fn in fns:
        fn()
    torch.cuda.synchronize()
    best = None
    for _ in range(3):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            for fn in fns:
                fn()
        e.record()
        torch.cuda.synchronize()
        t = s.elapsed_time(e) * 1e3 / (reps * len(fns))
        best = t if best is None else min(best, t)
    return best


def report(name, us, nbytes, **kw):
    gbs = nbytes / us / 1e3
    print(json.dumps(dict(op=name, B=B, K=K, us=round(us, 1), algorithmic_MB=round(nbytes / 1e6, 1),
                          algorithmic_GBps=round(gbs, 1), frac_of_measured_hbm=round(gbs / peak, 3), clocks=clocks(), **kw)),
          flush=True)


tiny = torch.empty(32, dtype=torch.int32, device=dev)
report("(host overhead of one C call: iota on 32 elements)", timed(lambda: _ops.iota_index(1, 32, dev)), 128)
lw = torch.randn(B, K, device=dev, generator=gen) - 1.4
u = torch.rand(B, dtype=torch.float64, device=dev, generator=gen)
flags = _ops.new_flags(dev)
_, lse, idx, _ = _ops.smc_step(lw, None, None, u, None, flags, "exact", True)
idx64 = idx.long()
n = B * K

for D in (1, 4, 10):
    x = torch.randn(B, K, D, device=dev, generator=gen) if D > 1 else torch.randn(B, K, device=dev, generator=gen)
    g = torch.randn_like(x)
    report("gather D=%d (int32 idx)" % D, timed(lambda: _ops.gather(x, idx, True)), n * (4 + 8 * D), D=D)
    report("gather D=%d (int64 idx)" % D, timed(lambda: _ops.gather(x, idx64, True)), n * (8 + 8 * D), D=D)
    gsrc = torch.empty_like(g)
    report("gather backward D=%d (sorted runs, C call)" % D,
           timed(lambda: _lib.call("aesmc_gather_bwd_f32", _lib.ptr(g), _lib.ptr(idx), 0, B, K, D, _lib.ptr(gsrc), 1)),
           n * (4 + 8 * D), D=D)

# collapsed weights (what training sees early on): a few parents own thousands of children each
_, _, idx_c, _ = _ops.smc_step(lw * 8.0, None, None, u, None, flags, "exact", True)
g1 = torch.randn(B, K, device=dev, generator=gen)
gs1 = torch.empty_like(g1)
report("gather backward D=1, collapsed weights (max run %d)" % int(torch.bincount(idx_c[0].long()).max()),
       timed(lambda: _lib.call("aesmc_gather_bwd_f32", _lib.ptr(g1), _lib.ptr(idx_c), 0, B, K, 1, _lib.ptr(gs1), 1)), n * 12)

a = lw.clone().requires_grad_(True)
b = torch.randn(B, K, device=dev, generator=gen).requires_grad_(True)
c = torch.randn(B, K, device=dev, generator=gen).requires_grad_(True)
lw_o, lse_o, _, _ = _ops.smc_step(a, b, c, u, None, flags, "exact", True)
g_lw, g_lse = torch.randn_like(lw_o), torch.randn_like(lse_o)
report("step backward (g_a, g_b, g_c from g_log_w, g_lse)",
       timed(lambda: torch.autograd.grad([lw_o, lse_o], [a, b, c], [g_lw, g_lse], retain_graph=True)), n * (4 + 4 + 8))
report("logsumexp rows (single launch, L2 flushed)", timed(lambda: _ops.logsumexp_rows(lw)), n * 4)
report("lognormexp rows (exp)", timed(lambda: _ops.lognormexp_rows(lw, True)), n * 8)
report("log ESS rows (single launch, L2 flushed)", timed(lambda: _ops.log_ess_rows(lw)), n * 4)
xs = torch.randn(B, K, device=dev, generator=gen)
report("weighted moments (x, x^2) (single launch, L2 flushed)", timed(lambda: _ops.weighted_moments(xs, lw)), n * 8)
# the same three on a ring of 4 input tables (268 MB > L2), launches back to back: the kernel itself
lws = [torch.randn(B, K, device=dev, generator=gen) - 1.4 for _ in range(4)]
xss = [torch.randn(B, K, device=dev, generator=gen) for _ in range(4)]
outs = [torch.empty(B, device=dev) for _ in range(4)]
outs2 = [torch.empty(B, device=dev) for _ in range(4)]
report("logsumexp rows (ring of 4 tables, back to back)",
       timed_ring([lambda i=i: _lib.call("aesmc_logsumexp_f32", _lib.ptr(lws[i]), B, K, _lib.ptr(outs[i]), _lib.ptr(flags))
                   for i in range(4)]), n * 4)
report("log ESS rows (ring of 4 tables, back to back)",
       timed_ring([lambda i=i: _lib.call("aesmc_log_ess_f32", _lib.ptr(lws[i]), B, K, _lib.ptr(outs[i])) for i in range(4)]), n * 4)
report("weighted moments (x, x^2) (ring of 4 tables, back to back)",
       timed_ring([lambda i=i: _lib.call("aesmc_weighted_moments_f32", _lib.ptr(xss[i]), _lib.ptr(lws[i]), B, K, 1,
                                         _lib.ptr(outs[i]), _lib.ptr(outs2[i])) for i in range(4)]), n * 8)
del lws, xss
acc = torch.zeros(B, K, device=dev)
report("IS accumulate", timed(lambda: _ops.is_accumulate(lw, lw, lw, acc, lw, False)), n * 20)
dist = torch.distributions.Normal(torch.randn(B, K, device=dev, generator=gen), 0.7, validate_args=False)
report("Normal.log_prob (one kernel)", timed(lambda: _ops.normal_log_prob(dist, xs)), n * 12)
report("compose genealogy index", timed(lambda: _ops.compose_index(idx, idx)), n * 12)
