#!/usr/bin/env python
"""bench.py -- SMC hot-path throughput on B200 (BASELINE.json metric: particle-steps/s = B*K*T / s).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's sm_100a path
    python bench.py --impl reference [--steps K] [--warmup W]       # the reference's CPU path (oracle port)
    torchrun --nproc-per-node N bench.py --gpus N ...               # one rank per GPU, weak scaling

Workload (BASELINE config 2): B = 4096 independent rows x K = 4096 particles x T = 100 time steps,
scalar latent (D = 1), float32.  One bench "step" = one full pass of the hot path over the T time steps:
    value  T launches of the fused step kernel (log-weights, logsumexp, systematic resampling, ancestral
           gather) + the log-evidence reduction, on synthetic log-prob tensors already resident in HBM
           (a ring of 4 x 192 MiB input sets, larger than the 126 MB L2, so nothing is served from cache)
    e2e    aesmc_b200.inference.infer('smc', ...) -- the public API a user calls -- on the bootstrap-filter
           LGSSM, with the observations [T,B] in pinned HOST memory copied H2D inside the timed region and
           the log-evidence [B] copied back D2H; `e2e` uses a model of the fused family (evaluated inside
           the step kernel), `e2e_eager` the same model written as plain torch callables, `e2e_eager_graph`
           those callables with the whole infer() call replayed as a CUDA graph (inference.GraphedInfer)
Rows are independent SMC problems, so ranks shard the batch axis with no data-path collective
("scaling": "weak": every rank processes B rows).
"""
import argparse
import json
import math
import os
import statistics as pystats
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_ROWS, K_PARTICLES, T_STEPS, D_LATENT = 4096, 4096, 100, 1
RING = 4
METRIC = "smc_particle_steps_per_sec"
UNIT = "particle-steps/s"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 6:
                self.rows.append(parts)

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
            except ValueError:
                continue
            for name, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": pystats.median(sm) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_setup(n_gpus):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    return rank, world, local


def barrier_sync(world):
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()


def max_over_ranks(seconds, world, dev):
    if world == 1:
        return seconds
    t = torch.tensor([seconds], dtype=torch.float64, device=dev)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return float(t.item())


# --------------------------------------------------------------------------------------------------
# this repo's arm
# --------------------------------------------------------------------------------------------------
def run_native(args):
    import __graft_entry__ as graft
    graft.build()
    from aesmc_b200 import _lib, _ops, inference
    from tests.models import lgssm

    rank, world, local = dist_setup(args.gpus)
    dev = torch.device("cuda", local)
    B, K, T, D = args.batch, args.particles, args.timesteps, D_LATENT
    mode = args.mode
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)

    # ---- synthetic inputs resident in HBM (SURVEY 8d core microbench) ----------------------------
    ring = [[torch.randn(B, K, device=dev, generator=gen) - 1.4 for _ in range(3)] for _ in range(RING)]
    arena = [torch.randn(B, K, device=dev, generator=gen), torch.empty(B, K, device=dev)]
    u = torch.rand(T - 1, B, dtype=torch.float64, device=dev, generator=gen)
    log_w = torch.empty(T, B, K, device=dev)
    idx = torch.empty(max(T - 1, 1), B, K, dtype=torch.int32, device=dev)
    lse = torch.empty(T, B, device=dev)
    flags = _ops.new_flags(dev)
    mode_code = _ops.mode_code(mode)
    stream = torch.cuda.current_stream()
    logK = math.log(K)

    def core_pass(events=None):
        for t in range(T):
            a, b, c = ring[t % RING]
            last = t == T - 1
            if events is not None:
                events[t].record(stream)
            _lib.call("aesmc_smc_step_f32", a.data_ptr(), b.data_ptr(), c.data_ptr(),
                      None if last else u[t].data_ptr(), B, K, log_w[t].data_ptr(), lse[t].data_ptr(),
                      None if last else idx[t].data_ptr(), None if last else arena[t & 1].data_ptr(),
                      None if last else arena[(t + 1) & 1].data_ptr(), D, flags.data_ptr(), mode_code)
        if events is not None:
            events[T].record(stream)
        return (lse - logK).sum(dim=0)  # inference.py:130-132

    # clocks / throttle reasons are sampled from the first warm-up pass to the end of the timed region (the
    # same kernels run back to back throughout; the timed region alone can be shorter than nvidia-smi's period)
    with ClockSampler(local) as clocks:
        for _ in range(args.warmup):
            lml = core_pass()
        barrier_sync(world)
        launches0 = _lib.launch_count()
        events = [[torch.cuda.Event(enable_timing=True) for _ in range(T + 1)] for _ in range(args.steps)]
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record(stream)
        for s in range(args.steps):
            lml = core_pass(events[s])
        stop.record(stream)
        barrier_sync(world)
        launches = _lib.launch_count() - launches0
        if len(clocks.rows) < 2:  # very short runs: keep the GPU on the same work until a sample lands
            t_end = time.time() + 1.0
            while len(clocks.rows) < 2 and time.time() < t_end:
                core_pass()
                torch.cuda.synchronize()
    seconds = max_over_ranks(start.elapsed_time(stop) / 1e3, world, dev)
    assert int(flags.item()) == 0 and bool(torch.isfinite(lml).all())
    ms_per_step = seconds * 1e3 / args.steps
    value = world * B * K * T / (seconds / args.steps)

    # dominant kernel: per-launch duration from the event pairs inside the timed region
    durs = [events[s][t].elapsed_time(events[s][t + 1]) for s in range(args.steps) for t in range(T - 1)]
    kernel_ms = sum(durs) / len(durs)
    bytes_per_launch = (20 + 8 * D) * B * K
    peak, peak_src = load_peaks()
    achieved = bytes_per_launch / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get("smc_step_kernel_%s" % mode)
    # (what aesmc_smc_step_f32 dispatches to at this shape: smc_step_x.cu's row kernel in exact mode for K = 1024 ... 16384
    # and a scalar latent, smc_step_reg.cu's otherwise)
    xk = mode == "exact" and D == 1 and K in (1024, 2048, 4096, 8192, 16384)
    kname = "smc_step_x_kernel<%d,1,0> (exact)" % (K // 16) if xk else "smc_step_reg_kernel (%s)" % mode
    roofline = {"bound": "hbm", "kernel": kname, "achieved": round(achieved, 1), "peak": peak,
                "unit": "GB/s", "frac": round(achieved / peak, 4), "traffic": traffic,
                "bytes_per_launch": bytes_per_launch, "launch_ms": round(kernel_ms, 4), "peak_source": peak_src}

    # ---- e2e: public infer() with host observation buffers -----------------------------------
    # Same model (BASELINE config 2 bootstrap-filter LGSSM), same public call, two kinds of user model:
    #   e2e        aesmc_b200.fused.ScalarLinearGaussianSSM -- infer() recognises it and evaluates the model
    #              inside the step kernel (one launch per time step)
    #   e2e_eager  plain torch callables (tests/models/lgssm.py) -- ~25 torch elementwise kernels per step
    e2e = e2e_eager = e2e_eager_graph = None
    if not args.no_e2e:
        from aesmc_b200 import fused
        ys = lgssm.simulate(T, B, seed=100 + rank)
        obs_host = torch.from_numpy(ys).pin_memory()
        u_host = torch.from_numpy(np.random.default_rng(7 + rank).random((T - 1, B))).pin_memory()
        out_host = torch.empty(B, dtype=torch.float32).pin_memory()
        # production setting for torch.distributions: argument/sample validation reads a device flag on
        # the host for every distribution built and every log_prob (7 synchronisations per time step)
        torch.distributions.Distribution.set_default_validate_args(False)
        del log_w, idx
        torch.cuda.empty_cache()

        def measure(models, reps, label):
            def one_pass():
                obs = obs_host.to(dev, non_blocking=True)
                uu = u_host.to(dev, non_blocking=True)
                with torch.no_grad():
                    res = inference.infer("smc", obs, *models, K, return_log_marginal_likelihood=True,
                                          return_latents=False, return_log_weight=False, uniforms=uu,
                                          resampling_mode=mode)
                out_host.copy_(res["log_marginal_likelihood"], non_blocking=True)
                torch.cuda.current_stream().synchronize()

            for _ in range(2):
                one_pass()
            barrier_sync(world)
            t0 = time.perf_counter()
            e_start, e_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e_start.record()
            for _ in range(reps):
                one_pass()
            e_stop.record()
            barrier_sync(world)
            sec = max(time.perf_counter() - t0, e_start.elapsed_time(e_stop) / 1e3)
            sec = max_over_ranks(sec, world, dev)
            assert bool(np.isfinite(out_host.numpy()).all())
            return {"value": world * B * K * T / (sec / reps), "unit": UNIT,
                    "h2d_bytes_per_step": obs_host.numel() * 4 + u_host.numel() * 8, "d2h_bytes_per_step": B * 4,
                    "ms_per_step": sec * 1e3 / reps, "steps": reps, "api": label}

        reps = max(1, min(args.steps, 5))
        fused_model = fused.ScalarLinearGaussianSSM(0.0, 1.0, 0.9, 0.0, 1.0, 1.0, 0.0, 0.5, device=dev)
        e2e = measure(fused_model.callables(), reps,
                      "aesmc_b200.inference.infer('smc') on fused.ScalarLinearGaussianSSM (model evaluated in the step kernel)")
        e2e_eager = measure(lgssm.bootstrap_filter(device=dev), reps,
                            "aesmc_b200.inference.infer('smc') on torch-eager user callables, "
                            "Distribution.set_default_validate_args(False)")

        # the same torch-eager user model with the whole infer() call captured once as a CUDA graph
        def measure_graphed(models, reps, label):
            obs_host = torch.from_numpy(lgssm.simulate(T, B, seed=7 + rank)).pin_memory()
            out_host = torch.empty(B, dtype=torch.float32).pin_memory()
            obs = obs_host.to(dev)
            g = inference.GraphedInfer("smc", [obs[t] for t in range(T)], *models, K, return_log_marginal_likelihood=True,
                                       return_latents=False, return_log_weight=False, resampling_mode=mode)

            def one_pass():
                dobs = obs_host.to(dev, non_blocking=True)
                res = g([dobs[t] for t in range(T)])
                out_host.copy_(res["log_marginal_likelihood"], non_blocking=True)
                torch.cuda.current_stream().synchronize()

            for _ in range(2):
                one_pass()
            barrier_sync(world)
            t0 = time.perf_counter()
            for _ in range(reps):
                one_pass()
            barrier_sync(world)
            sec = max_over_ranks(time.perf_counter() - t0, world, dev)
            assert bool(np.isfinite(out_host.numpy()).all())
            return {"value": world * B * K * T / (sec / reps), "unit": UNIT, "h2d_bytes_per_step": obs_host.numel() * 4,
                    "d2h_bytes_per_step": B * 4, "ms_per_step": sec * 1e3 / reps, "steps": reps, "api": label}

        e2e_eager_graph = measure_graphed(lgssm.bootstrap_filter(device=dev), reps,
                                          "aesmc_b200.inference.GraphedInfer('smc') on the same torch-eager callables: "
                                          "the whole infer() call replayed as one CUDA graph, uniforms drawn on the device")

    # ---- extra legs: parity counts, strong scaling, config-5 sweep, config-4 training ----
    extras = {}
    if not args.no_extras:
        torch.cuda.empty_cache()
        ctx = {"rank": rank, "world": world, "dev": dev, "mode": mode, "K": K, "T": T, "B": B}
        for name, fn in (("parity", leg_parity), ("strong", leg_strong), ("sweep_c5", leg_sweep_c5),
                         ("infer_c3", leg_infer_c3), ("train_c4", leg_train_c4), ("train_lgssm", leg_train_lgssm)):
            try:
                extras[name] = fn(ctx, ring, arena, u, value, ms_per_step)
            except Exception as exc:  # an extra leg must never take the headline line down with it
                extras[name] = {"error": "%s: %s" % (type(exc).__name__, exc)}
            torch.cuda.empty_cache()

    # ---- CPU baseline beside it (rank 0, N = 1 only): the oracle port on a bounded sample ---------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu_baseline = time_oracle_core(K, T, D, args.cpu_rows)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "BASELINE config 2: 1-D LGSSM-shaped SMC core, B=%d rows/GPU x K=%d particles x T=%d, D=%d"
                                       % (B, K, T, D), "resampling_mode": mode, "parallelism": "batch rows sharded, dp%d" % world,
                           "l2": "inputs ring %d x %d MiB > 126 MB L2; no flush needed" % (RING, 3 * B * K * 4 >> 20)},
                "roofline": roofline, "cpu_baseline": cpu_baseline, "clocks": clocks.summary(), "e2e": e2e,
                "e2e_eager": e2e_eager, "e2e_eager_graph": e2e_eager_graph, "gpu_launches": launches}
        line.update(extras)
        print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


# --------------------------------------------------------------------------------------------------
# extra legs: each returns a small dict that becomes a key of the JSON line (all ranks must call them:
# the multi-rank ones contain collectives; rank 0's dict is the one printed)
# --------------------------------------------------------------------------------------------------
def _event_time_ms(fn, reps, world, dev):
    """fn() x reps between two CUDA events on the current stream, barrier + sync either side, max over ranks."""
    barrier_sync(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    barrier_sync(world)
    return max_over_ranks(e0.elapsed_time(e1) / 1e3, world, dev) * 1e3 / reps


def _graphed(fn, dev):
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay()
    torch.cuda.synchronize()
    return g


def leg_parity(ctx, ring, arena, u, value, ms_per_step, rows=64):
    """the GPU path against the CPU oracle (oracle/smc_oracle.c, the reference's
    arithmetic: inference.py:250-264, math.py:6-51) on `rows` full-size rows of this very workload --
    ancestor-index mismatches, max |delta idx|, and the relative error of log-weights, per-step lse and the
    log-evidence.  Rank 0 only (the oracle is the checker here, never the thing measured)."""
    if ctx["rank"] != 0:
        return None
    from aesmc_b200 import _lib, _ops
    from oracle import core as oracle
    dev, K, T, mode = ctx["dev"], ctx["K"], ctx["T"], ctx["mode"]
    R = min(rows, ctx["B"])
    code = _ops.mode_code(mode)
    sub = [[t[:R].contiguous() for t in trio] for trio in ring]
    x = [arena[0][:R].clone(), torch.empty(R, K, device=dev)]
    x0 = x[0].cpu().numpy().copy()
    uu = u[:, :R].contiguous()
    lw = torch.empty(T, R, K, device=dev)
    idx = torch.empty(max(T - 1, 1), R, K, dtype=torch.int32, device=dev)
    lse = torch.empty(T, R, device=dev)
    flags = _ops.new_flags(dev)
    for t in range(T):
        a, b, c = sub[t % RING]
        last = t == T - 1
        _lib.call("aesmc_smc_step_f32", _lib.ptr(a), _lib.ptr(b), _lib.ptr(c), None if last else _lib.ptr(uu[t]), R, K,
                  _lib.ptr(lw[t]), _lib.ptr(lse[t]), None if last else _lib.ptr(idx[t]),
                  None if last else _lib.ptr(x[t & 1]), None if last else _lib.ptr(x[(t + 1) & 1]), 1, _lib.ptr(flags),
                  code)
    torch.cuda.synchronize()
    g_lw, g_idx, g_lse = lw.cpu().numpy(), idx.cpu().numpy(), lse.cpu().numpy()
    g_x = x[(T - 1) & 1].cpu().numpy()
    h = [[t.cpu().numpy() for t in trio] for trio in sub]
    hu = uu.cpu().numpy()
    t0 = time.perf_counter()
    mism = near = n_idx = 0
    max_d = 0
    lw_bits_differ = 0
    lse_rel = lw_rel = 0.0
    xr = x0
    ev_ref = np.zeros(R, np.float64)
    for t in range(T):
        a, b, c = h[t % RING]
        lw_ref = oracle.log_weight(a, b, c)
        lw_bits_differ += int((lw_ref.view(np.int32) != g_lw[t].view(np.int32)).sum())
        lw_rel = max(lw_rel, float(np.max(np.abs(lw_ref - g_lw[t]) / np.maximum(np.abs(lw_ref), 1e-30))))
        if t < T - 1:
            idx_ref, st, lse_ref, _, cdf_ref = oracle.sample_ancestral_index(lw_ref, hu[t], return_parts=True)
            assert st == 0
            d = idx_ref.astype(np.int64) - g_idx[t].astype(np.int64)
            bad = np.nonzero(d)
            n_idx += d.size
            mism += len(bad[0])
            if len(bad[0]):
                max_d = max(max_d, int(np.abs(d).max()))
                # a disagreement is "at a boundary" when the position (u + k)/K lies within 1 ulp of the
                # CDF entry that separates the two candidate ancestors (north_star's allowance)
                cn = (cdf_ref / cdf_ref[:, -1:]).astype(np.float32)
                for r_, k_ in zip(*bad):
                    pos = (hu[t][r_] + k_) / K
                    j = min(int(idx_ref[r_, k_]), int(g_idx[t][r_, k_]))
                    edge = np.float32(cn[r_, j])
                    near += int(abs(pos - float(edge)) <= float(np.spacing(edge)))
            xr = oracle.resample(xr.reshape(R, K, 1), idx_ref).reshape(R, K)
        else:
            lse_ref = oracle.logsumexp_rows(lw_ref)
        lse_rel = max(lse_rel, float(np.max(np.abs(lse_ref - g_lse[t]) / np.maximum(np.abs(lse_ref), 1e-30))))
        ev_ref += lse_ref.astype(np.float64) - math.log(K)
    ev_gpu = (g_lse.astype(np.float64) - math.log(K)).sum(axis=0)
    return {"against": "oracle/smc_oracle.c (CPU restatement of inference.py:250-264, math.py:6-51), same inputs and uniforms",
            "rows": R, "K": K, "T": T, "mode": mode, "indices_compared": n_idx, "index_mismatches": mism,
            "mismatches_within_1ulp_of_cdf_boundary": near, "max_abs_idx_diff": max_d,
            "log_w_bits_differing": lw_bits_differ, "log_w_max_rel_err": lw_rel, "lse_max_rel_err": lse_rel,
            "log_evidence_max_rel_err": float(np.max(np.abs(ev_gpu - ev_ref) / np.maximum(np.abs(ev_ref), 1e-30))),
            "resampled_latent_equal": bool(np.array_equal(g_x, xr)) if T > 1 else None,
            "flags": int(flags.item()), "tolerance": "indices bit-exact (exact mode); log-weights / log-evidence 1e-5 relative",
            "oracle_seconds": round(time.perf_counter() - t0, 2)}


def leg_strong(ctx, ring, arena, u, value, ms_per_step):
    """Strong scaling of BASELINE config 2: B = 4096 rows IN TOTAL, 4096 / N per rank (at 8 GPUs 512 rows per
    rank: less than one wave of the 592-CTA grid).  The T launches + the evidence reduction are replayed as
    one CUDA graph per rank (at 512 rows a launch is ~25 us: issuing from Python would time the host)."""
    from aesmc_b200 import _lib, _ops
    world, dev, K, T, mode = ctx["world"], ctx["dev"], ctx["K"], ctx["T"], ctx["mode"]
    total = ctx["B"]
    if world == 1:
        return {"value": value, "unit": UNIT, "rows_total": total, "rows_per_rank": total, "ms_per_step": ms_per_step,
                "note": "N = 1: identical to the headline `value`"}
    Bs = total // world
    code = _ops.mode_code(mode)
    sub = [[t[:Bs] for t in trio] for trio in ring]  # leading rows of a row-major table: contiguous
    x = [arena[0][:Bs], arena[1][:Bs]]
    lw = torch.empty(T, Bs, K, device=dev)
    idx = torch.empty(T, Bs, K, dtype=torch.int32, device=dev)
    lse = torch.empty(T, Bs, device=dev)
    lml = torch.empty(Bs, device=dev)
    flags = _ops.new_flags(dev)

    def one_pass():
        for t in range(T):
            a, b, c = sub[t % RING]
            last = t == T - 1
            _lib.call("aesmc_smc_step_f32", _lib.ptr(a), _lib.ptr(b), _lib.ptr(c), None if last else _lib.ptr(u[t]), Bs, K,
                      _lib.ptr(lw[t]), _lib.ptr(lse[t]), None if last else _lib.ptr(idx[t]),
                      None if last else _lib.ptr(x[t & 1]), None if last else _lib.ptr(x[(t + 1) & 1]), 1,
                      _lib.ptr(flags), code)
        lml.copy_((lse - math.log(K)).sum(dim=0))

    for _ in range(2):
        one_pass()
    ms_stepwise = _event_time_ms(one_pass, 5, world, dev)
    g = _graphed(one_pass, dev)
    ms = _event_time_ms(g.replay, 10, world, dev)
    assert int(flags.item()) == 0
    return {"value": Bs * world * K * T / (ms * 1e-3), "unit": UNIT, "rows_total": Bs * world, "rows_per_rank": Bs,
            "ms_per_step": ms, "ms_per_step_launched_from_python": ms_stepwise, "scaling": "strong",
            "note": "one CUDA-graph replay per rank per pass; device time, max over ranks; no data-path collective"}


def leg_sweep_c5(ctx, ring, arena, u, value, ms_per_step):
    """BASELINE config 5 (resampling-bound, multi-CTA path): K = 1e5 and 1e6 particles per row, B = 64 rows IN
    TOTAL (64 / N per rank), T = 20 steps replayed as one CUDA graph; exact and fast modes."""
    from aesmc_b200 import _lib, _ops
    world, dev = ctx["world"], ctx["dev"]
    peak, _ = load_peaks()
    Bs, T = max(64 // world, 1), 20
    out = []
    for K in (100000, 1000000):
        gen = torch.Generator(device=dev).manual_seed(5 + ctx["rank"])
        trio = [[torch.randn(Bs, K, device=dev, generator=gen) - 1.4 for _ in range(3)] for _ in range(2)]
        x = [torch.randn(Bs, K, device=dev, generator=gen), torch.empty(Bs, K, device=dev)]
        uu = torch.rand(T, Bs, dtype=torch.float64, device=dev, generator=gen)
        lw, lse = torch.empty(Bs, K, device=dev), torch.empty(Bs, device=dev)
        idx = torch.empty(Bs, K, dtype=torch.int32, device=dev)
        flags = _ops.new_flags(dev)
        ws_bytes = int(_lib.load().aesmc_smc_step_workspace_bytes(Bs, K))
        ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
        for mode in ("exact", "fast"):
            code = _ops.mode_code(mode)

            def run():
                for t in range(T):
                    a, b, c = trio[t & 1]
                    _lib.call("aesmc_smc_step_ws_f32", _lib.ptr(a), _lib.ptr(b), _lib.ptr(c), _lib.ptr(uu[t]), Bs, K,
                              _lib.ptr(lw), _lib.ptr(lse), _lib.ptr(idx), _lib.ptr(x[t & 1]), _lib.ptr(x[(t + 1) & 1]), 1,
                              _lib.ptr(flags), code, _lib.ptr(ws) if ws_bytes else None, ws_bytes)

            for _ in range(2):
                run()
            g = _graphed(run, dev)
            ms = _event_time_ms(g.replay, 3, world, dev) / T
            gbs = 28.0 * Bs * K / (ms * 1e-3) / 1e9
            out.append({"K": K, "rows_total": Bs * world, "rows_per_rank": Bs, "mode": mode, "us_per_step": round(ms * 1e3, 2),
                        "value": Bs * world * K / (ms * 1e-3), "unit": UNIT, "algorithmic_GBps_per_gpu": round(gbs, 1),
                        "frac_of_measured_hbm": round(gbs / peak, 4), "flags": int(flags.item())})
        del trio, x, lw, idx, ws
        torch.cuda.empty_cache()
    return out


def leg_train_c4(ctx, ring, arena, u, value, ms_per_step):
    """BASELINE config 4: AESMC training ('aesmc' = SMC ELBO, losses.py:5-65) of the nonlinear state-space
    model with an MLP proposal, batch rows sharded over the ranks, the flattened-gradient NCCL all-reduce of
    train.py:36->37 INSIDE the timed optimiser step; its own duration is timed with CUDA events."""
    from aesmc_b200 import distributed, losses, train
    from tests.models import nonlinear
    world, dev, rank = ctx["world"], ctx["dev"], ctx["rank"]
    Bl, K, T = 512, 1024, 20
    torch.distributions.Distribution.set_default_validate_args(False)
    torch.manual_seed(1000 + rank)
    np.random.seed(1000 + rank)
    init = nonlinear.Initial(dev)
    loader = train.get_synthetic_dataloader(init, nonlinear.Transition().to(dev), nonlinear.Emission().to(dev), T, Bl)
    batch = next(iter(loader))
    torch.manual_seed(0)  # identical replicas
    trans, emis, prop = nonlinear.Transition(scale=2.0).to(dev), nonlinear.Emission(mult=0.03).to(dev), nonlinear.Proposal().to(dev)
    params = list(train.get_chained_params(trans, emis, prop))
    opt = torch.optim.Adam(params, lr=1e-3)
    ar_events = []

    def step(record=False):
        opt.zero_grad()
        loss = losses.get_loss(batch, K, "aesmc", init, trans, emis, prop)
        loss.backward()
        if world > 1:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            distributed.all_reduce_gradients(params, Bl, Bl * world)
            e1.record()
            if record:
                ar_events.append((e0, e1))
        opt.step()
        return loss

    for _ in range(3):
        step()
    reps = 8
    ms = _event_time_ms(lambda: step(True), reps, world, dev)
    ar_us = None
    if ar_events:
        ar_us = max_over_ranks(sum(a.elapsed_time(b) for a, b in ar_events) / len(ar_events) * 1e-3, world, dev) * 1e6
    flat = torch.cat([p.detach().reshape(-1) for p in params])
    in_sync = True
    if world > 1:
        ref = flat.clone()
        torch.distributed.broadcast(ref, src=0)
        ok = torch.tensor([int(torch.equal(ref, flat))], device=dev)
        torch.distributed.all_reduce(ok, op=torch.distributed.ReduceOp.MIN)
        in_sync = bool(ok.item())
    out = {"model": "nonlinear SSM + MLP proposal (hidden 32), get_loss('aesmc') forward + backward + all-reduce + Adam, torch-eager",
           "rows_per_rank": Bl, "rows_total": Bl * world, "K": K, "T": T, "ms_per_optimizer_step": ms,
           "allreduce_us": ar_us, "allreduce_floats": int(flat.numel()), "replicas_in_sync": in_sync,
           "value": Bl * world * K * T / (ms * 1e-3), "unit": UNIT, "scaling": "weak"}
    # the same step replayed as ONE CUDA graph per rank: forward, backward, the NCCL all-reduce (captured inside the
    # graph) and Adam, fed by the graph-captured prior sampler (train.GraphedPriorSampler: data generated on the device)
    try:
        torch.manual_seed(2000 + rank)
        sampler = train.GraphedPriorSampler(init, nonlinear.Transition().to(dev), nonlinear.Emission().to(dev), T, Bl)
        gopt = torch.optim.Adam(params, lr=1e-3, capturable=True)
        gstep = train.GraphedTrainStep(sampler(clone=True), K, "aesmc", init, trans, emis, prop, gopt)
        for _ in range(2):
            gstep(sampler())
        gms = _event_time_ms(lambda: gstep(sampler()), reps, world, dev)
        flat = torch.cat([p.detach().reshape(-1) for p in params])
        g_sync = True
        if world > 1:
            ref = flat.clone()
            torch.distributed.broadcast(ref, src=0)
            ok = torch.tensor([int(torch.equal(ref, flat))], device=dev)
            torch.distributed.all_reduce(ok, op=torch.distributed.ReduceOp.MIN)
            g_sync = bool(ok.item())
        gstep.release()
        sampler.graph.reset()
        torch.cuda.synchronize(dev)
        out.update({"graph_replay_ms_per_optimizer_step": gms, "graph_replay_value": Bl * world * K * T / (gms * 1e-3),
                    "graph_replay_includes": "on-device prior sampling + forward + backward + captured all-reduce + Adam",
                    "graph_replicas_in_sync": g_sync})
    except Exception as exc:  # noqa: BLE001
        out["graph_replay_error"] = "%s: %s" % (type(exc).__name__, exc)
    return out


def leg_infer_c3(ctx, ring, arena, u, value, ms_per_step):
    """BASELINE config 3: infer('smc') on the 10-D LGSSM with dense transition / emission and a learned Gaussian proposal
    (tests/models/lgssm_dense.py), B = K = 1024 per rank, T = 20, observations from pinned host memory, log-evidence back
    to the host: the user's modules as torch-eager callables, and the same modules linked into the fused vector path
    (aesmc_b200.fused.link_dense: one model launch + one step launch per time step)."""
    from aesmc_b200 import fused, inference
    from tests.models import lgssm_dense
    world, dev, rank = ctx["world"], ctx["dev"], ctx["rank"]
    dx = dy = 10
    B, K, T = 1024, 1024, 20
    torch.distributions.Distribution.set_default_validate_args(False)
    A, C = lgssm_dense.make_system(dx, dy, seed=3, device=dev)
    ys = lgssm_dense.simulate(A, C, T, B, 1.0, 0.5, 0.5, seed=4 + rank).pin_memory()
    init = lgssm_dense.Initial(dx, 1.0, dev)
    trans, emis = lgssm_dense.Transition(A, 0.5).to(dev), lgssm_dense.Emission(C, 0.5).to(dev)
    torch.manual_seed(0)
    prop = lgssm_dense.Proposal(dx, dy).to(dev)

    def one():
        with torch.no_grad():
            obs = ys.to(dev, non_blocking=True)
            r = inference.infer("smc", obs, init, trans, emis, prop, K, return_log_marginal_likelihood=True, return_latents=False)
            return r["log_marginal_likelihood"].cpu()

    out = {"model": "10-D dense LGSSM, learned Gaussian proposal", "rows_per_rank": B, "rows_total": B * world, "K": K, "T": T,
           "D": dx, "h2d_bytes_per_step": int(ys.numel() * 4), "d2h_bytes_per_step": B * 4}
    for label in ("eager", "fused"):
        if label == "fused":
            fused.link_dense(init, trans, emis, prop)
        for _ in range(2):
            one()
        reps = 5 if label == "eager" else 20
        t0 = time.perf_counter()
        barrier_sync(world)
        t0 = time.perf_counter()
        for _ in range(reps):
            one()
        torch.cuda.synchronize(dev)
        ms = max_over_ranks((time.perf_counter() - t0) / reps, world, dev) * 1e3
        out[label + "_ms"] = ms
        out[label + "_value"] = B * world * K * T / (ms * 1e-3)
    out["speedup_fused_over_eager"] = out["eager_ms"] / out["fused_ms"]
    # fused: model kernel 8 D + 4 bytes, step kernel 8 + 8 D + 8 bytes per particle-step
    peak, _ = load_peaks()
    out["fused_frac_of_measured_hbm"] = round((16.0 * dx + 20.0) * B * K * T / (out["fused_ms"] * 1e-3) / 1e9 / peak, 4)
    out["unit"] = UNIT
    return out



def leg_train_lgssm(ctx, ring, arena, u, value, ms_per_step):
    """AESMC training of the reference's own trainable LGSSM (test/models/lgssm.py:19-72: learnable transition /
    emission multipliers, two-Linear proposal) at B = 1024 rows per rank, K = 4096, T = 50: one optimiser step =
    get_loss('aesmc') forward + backward (+ gradient all-reduce at N > 1) + Adam.  `fused`: the modules linked to the
    fused kernels (aesmc_b200.fused.link: T forward + T backward launches); `generic`: the same modules through
    torch-eager callables and torch autograd around the step kernel."""
    from aesmc_b200 import distributed, fused, losses
    from tests.models import lgssm
    world, dev, rank = ctx["world"], ctx["dev"], ctx["rank"]
    B, K, T = 1024, 4096, 50
    torch.distributions.Distribution.set_default_validate_args(False)
    obs = torch.from_numpy(lgssm.simulate(T, B, A=0.9, Q=1.0, C=1.0, R=0.25, seed=50 + rank)).to(dev)
    obs_list = [obs[t] for t in range(T)]
    out = {"model": "reference-style trainable LGSSM (learnable multipliers, Linear(1,1) / Linear(2,1) proposal), 'aesmc' loss + Adam",
           "rows_per_rank": B, "rows_total": B * world, "K": K, "T": T, "unit_ms": "ms per optimiser step"}
    for label, link in (("fused", True), ("generic", False)):
        torch.manual_seed(0)
        np.random.seed(0)
        init, trans, emis, prop = (lgssm.Initial(0.0, 1.0), lgssm.Transition(0.5, 1.0).to(dev), lgssm.Emission(0.5, 0.5).to(dev),
                                   lgssm.Proposal(0.9, 0.9).to(dev))
        if link:
            fused.link(init, trans, emis, prop)
        params = [q for m in (trans, emis, prop) for q in m.parameters()]
        opt = torch.optim.Adam(params, lr=1e-3)

        def step():
            opt.zero_grad()
            loss = losses.get_loss(obs_list, K, "aesmc", init, trans, emis, prop)
            loss.backward()
            if world > 1:
                distributed.all_reduce_gradients(params, B, B * world)
            opt.step()
            return loss

        for _ in range(2):
            step()
        reps = 5 if link else 3
        ms = _event_time_ms(step, reps, world, dev)
        out[label + "_ms"] = ms
        out[label + "_value"] = B * world * K * T / (ms * 1e-3)
        del init, trans, emis, prop, params, opt
        torch.cuda.empty_cache()
    out["speedup_fused_over_generic"] = out["generic_ms"] / out["fused_ms"]
    out["unit"] = UNIT
    # forward + backward of the fused path move 16 + 20 bytes per particle-step
    peak, _ = load_peaks()
    out["fused_frac_of_measured_hbm"] = round(36.0 * B * K * T / (out["fused_ms"] * 1e-3) / 1e9 / peak, 4)
    return out


def time_oracle_core(K, T, D, rows):
    """The C restatement of the reference's per-step arithmetic (oracle/smc_oracle.c), one host
    thread, on `rows` rows of the same workload."""
    from oracle import core as oracle
    rng = np.random.default_rng(0)
    a, b, c = [(rng.standard_normal((RING, rows, K)) - 1.4).astype(np.float32) for _ in range(3)]
    x = rng.standard_normal((rows, K, D)).astype(np.float32)
    u = rng.random((T - 1, rows))
    t0 = time.perf_counter()
    _, _, st = oracle.core_pass(a, b, c, u, x, T)
    dt = time.perf_counter() - t0
    assert st == 0
    return {"value": rows * K * T / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "oracle/smc_oracle.c core pass, %d rows x K=%d x T=%d (%.1f s)" % (rows, K, T, dt)}


# --------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU algorithm (oracle/reference_port.py) on the host cores
# --------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import reference_port as port
    from tests.models import lgssm
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    K, T = args.particles, args.timesteps
    rows = args.ref_rows
    ys = lgssm.simulate(T, rows, seed=100)
    obs = [torch.from_numpy(y) for y in ys]
    models = lgssm.bootstrap_filter()

    def one():
        with torch.no_grad():
            res = port.infer("smc", obs, *models, K, return_log_marginal_likelihood=True, return_latents=False,
                             return_log_weight=False)
        return res["log_marginal_likelihood"]

    for _ in range(min(args.warmup, 1)):
        one()
    times = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        one()
        times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    value = rows * K * T / sec
    sample = "oracle/reference_port.infer('smc') bootstrap LGSSM, %d rows x K=%d x T=%d per step" % (rows, K, T)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "BASELINE config 2 shape (K=%d, T=%d) on a bounded sample of %d rows" % (K, T, rows)},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--mode", default="exact", choices=["exact", "fast"])
    ap.add_argument("--batch", type=int, default=B_ROWS)
    ap.add_argument("--particles", type=int, default=K_PARTICLES)
    ap.add_argument("--timesteps", type=int, default=T_STEPS)
    ap.add_argument("--cpu-rows", type=int, default=256)
    ap.add_argument("--ref-rows", type=int, default=64)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the parity / strong / sweep_c5 / train_c4 legs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
