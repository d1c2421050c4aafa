"""Importance sampling / sequential Monte Carlo driver.

Keeps the call surface of the reference's aesmc/inference.py -- infer :8-193,
get_resampled_latents :196-231, sample_ancestral_index :234-269 -- so user-supplied callables
(initial / transition / emission / proposal returning torch.distributions) stay drop-in, while the
per-time-step arithmetic runs in libaesmc_b200's fused step kernel:

    reference, per step (host numpy + ~12 torch passes)      here (one launch)
    ------------------------------------------------------   ---------------------------------
    log_w = trans + emis - prop        inference.py:125-126   aesmc_smc_step_f32:
    sample_ancestral_index(log_w)      inference.py:234-269     log_w, lse, systematic ancestors,
    state.resample(latent, index)      inference.py:102-104     gather of the newest latent
    logsumexp(stack(log_weights))      inference.py:130

The reference gathers its entire latent history with the newest index at every step (O(T^2),
SURVEY Q1); here user callables receive a lazy sequence whose last entry is the fused gather result
and whose older entries are gathered on access with the same (reference) semantics.

Extra keyword-only arguments (all optional, defaults reproduce the reference's behaviour):
    uniforms         [T-1, B] float64 array/tensor of per-row resampling uniforms; default: drawn
                     from numpy's global RNG, np.random.uniform(size=[B, 1]) once per step, exactly
                     as inference.py:250 does (same seed -> same uniforms as the reference)
    resampling_mode  'exact' | 'fast' (default: module setting, see set_resampling_mode)
    check_finite     read the device NaN/degeneracy flag once at the end and raise
                     FloatingPointError like inference.py:244-245 (default True; the only host
                     synchronisation in infer).  As in the reference, only weights that are actually
                     RESAMPLED (SMC, steps 0..T-2) can raise; NaN / infinite weights in 'is' mode or
                     at the last step propagate into the result.  Stricter than the reference in one
                     case: a resampled row whose weights are all -inf (or contain +inf) raises too,
                     where the reference silently emits out-of-range ancestor indices
"""
import collections.abc
import contextlib
import math as _pymath

import numpy as np
import torch

from . import _ops
from . import fused
from . import state
from ._ops import get_resampling_mode, set_resampling_mode  # noqa: F401  (re-exported)


def _first_tensor(value):
    return next(iter(value.values())) if isinstance(value, dict) else value


def _map_tensors(fn, value):
    if isinstance(value, dict):
        return {name: fn(v) for name, v in value.items()}
    return fn(value)


class ResampledHistory(collections.abc.Sequence):
    """previous_latents as seen by proposal/transition at time t in SMC mode: element j is
    state.resample(latents[j], ancestral_indices[t-1]) (inference.py:102-104), computed on first
    access.  ``newest`` is the already-gathered last element produced by the fused step kernel."""

    def __init__(self, history, index32, newest, home):
        self._history = list(history)
        self._index = index32
        self._home = home
        self._cache = {}
        if newest is not None:
            self._cache[len(self._history) - 1] = newest

    def __len__(self):
        return len(self._history)

    def _one(self, j):
        if j not in self._cache:
            def gather(v):
                out = _ops.gather(_ops.to_device(v), self._index, sorted_rows=True)
                return out if v.is_cuda else out.to(self._home)
            self._cache[j] = _map_tensors(gather, self._history[j])
        return self._cache[j]

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self._one(j) for j in range(*i.indices(len(self)))]
        n = len(self)
        if i < 0:
            i += n
        if not 0 <= i < n:
            raise IndexError("previous_latents index out of range")
        return self._one(i)


def _as_f32_device(t):
    """Log-prob tensor -> contiguous float32 CUDA (autograd-aware; no copy if already so)."""
    if not t.is_cuda:
        t = t.to(_ops.device())
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _fusable(latent):
    return torch.is_tensor(latent) and latent.is_cuda and latent.dtype == torch.float32 and latent.dim() >= 2


def infer(inference_algorithm, observations, initial, transition, emission, proposal, num_particles,
          return_log_marginal_likelihood=False, return_latents=True, return_original_latents=False,
          return_log_weight=True, return_log_weights=False, return_ancestral_indices=False, *,
          uniforms=None, resampling_mode=None, check_finite=True, _allow_fused=True):
    """Importance sampling ('is') or sequential Monte Carlo ('smc') on a state-space model.

    Arguments, callable conventions and the returned dict (keys log_marginal_likelihood, latents,
    original_latents, log_weight, log_weights, ancestral_indices, last_latent; un-requested entries
    are None) are those of the reference's infer() (inference.py:13-70, 187-193).
    """
    if inference_algorithm not in ("is", "smc"):
        raise ValueError("inference_algorithm must be either is or smc. currently = {}".format(
            inference_algorithm))
    smc = inference_algorithm == "smc"
    T = len(observations)
    B = _first_tensor(observations[0]).size(0)
    K = num_particles
    if smc:
        # models of the fused family (aesmc_b200.fused) run with their sampling and log-densities inside
        # the step kernel; everything else takes the generic path below
        model = fused.model_of(initial, transition, emission, proposal) if _allow_fused else None
        evidence_only = return_log_marginal_likelihood and not (
            return_latents or return_original_latents or return_log_weight or return_log_weights or return_ancestral_indices)
        if model is not None and fused.applicable(model, observations, K, evidence_only):
            if isinstance(model, fused.VectorLinearGaussianSSM):  # vector latents: model launch + step launch per time step
                return fused.infer_fused_vector(model, observations, K, return_log_marginal_likelihood, return_latents,
                                                return_original_latents, return_log_weight, return_log_weights,
                                                return_ancestral_indices, uniforms=uniforms,
                                                resampling_mode=resampling_mode, check_finite=check_finite)
            if torch.is_grad_enabled() and model.requires_grad():
                # training (losses.get_loss): forward and backward on the fused kernels, one launch per time step each
                result = dict.fromkeys(("log_marginal_likelihood", "latents", "original_latents", "log_weight",
                                        "log_weights", "ancestral_indices", "last_latent"))
                result["log_marginal_likelihood"] = fused.evidence_with_grad(
                    model, observations, K, uniforms=uniforms, resampling_mode=resampling_mode, check_finite=check_finite)
                return result
            return fused.infer_fused(model, observations, K, return_log_marginal_likelihood, return_latents,
                                     return_original_latents, return_log_weight, return_log_weights,
                                     return_ancestral_indices, uniforms=uniforms, resampling_mode=resampling_mode,
                                     check_finite=check_finite)
    keep_originals = return_original_latents or return_latents
    grad_path = torch.is_grad_enabled()
    # log-weights of every step are retained only when somebody can read them: the caller asked for
    # them, or autograd saves them anyway; the last one serves return_log_weight
    keep_log_weights = return_log_weights
    keep_index = return_latents or return_ancestral_indices
    u_all = None  # [T-1, B] float64 on the device: uploaded once, sliced per step

    def draw_uniforms(t, dev):
        nonlocal u_all
        if u_all is None:
            if uniforms is None:
                # inference.py:250 draws np.random.uniform(size=[B, 1]) once per step from numpy's global RNG;
                # the T-1 draws are taken here in one call -- the same stream values in the same order (unless a
                # user callable itself consumes numpy's global RNG between steps)
                host = np.random.uniform(size=[T - 1, B])
            else:
                host = uniforms
            u_all = _ops.uniforms_table_to_device(host, T - 1, B, dev)
        return u_all[t - 1]

    # ---- t = 0 (inference.py:85-98) ------------------------------------------------------------
    q = proposal(time=0, observations=observations)
    latent = state.sample(q, B, K)
    history = [latent]
    lq = state.log_prob(q, latent)
    lp = state.log_prob(initial(), latent)
    le = state.log_prob(emission(latents=history, time=0), state.expand_observation(observations[0], K))
    home = lq.device
    pending = tuple(_as_f32_device(v) for v in (lp, le, lq))  # log_w = (lp + le) - lq
    dev = pending[0].device
    # flag word 0: steps whose weights are resampled -- the only place the reference raises
    # (sample_ancestral_index, inference.py:244-245); word 1: everything else ('is' mode, the last step),
    # where the reference lets NaN / infinite weights propagate into the result, and so does this
    flag_words = torch.zeros(2, dtype=torch.int32, device=dev)
    flags, quiet_flags = flag_words[0:1], flag_words[1:2]
    originals = [latent] if keep_originals else None
    log_weights, lses, ancestors = [], [], []
    total = None  # running sum of log-weights in 'is' mode

    def close_step(resample_u):
        """Fold the pending log-probs into a log-weight; in SMC mode with ``resample_u`` also draw
        the ancestors and gather the newest latent."""
        nonlocal total
        a, b, c = pending
        if smc:
            newest = history[-1]
            fuse = resample_u is not None and _fusable(newest)
            x = newest.contiguous() if fuse else None
            log_w, lse, idx, x_res = _ops.smc_step(a, b, c, resample_u, x,
                                                   flags if resample_u is not None else quiet_flags,
                                                   resampling_mode, resample=resample_u is not None)
            if keep_log_weights:
                log_weights.append(log_w)
            else:
                log_weights[:] = [log_w]
            lses.append(lse)
            return idx, x_res
        if grad_path and any(t.requires_grad for t in (a, b, c)):
            log_w, _, _, _ = _ops.smc_step(a, b, c, None, None, quiet_flags, resampling_mode, resample=False)
            # inference.py:156 (sequential over t).  `total` is never an alias of a per-step tensor: a
            # later step without gradients may accumulate into it in place
            total = log_w + 0 if total is None else total + log_w
        else:  # one pass: log_w = (a + b) - c and total += log_w (aesmc_is_accumulate_f32)
            log_w = torch.empty_like(a) if keep_log_weights else None
            first = total is None
            if first:
                total = torch.empty_like(a)
            elif total.requires_grad:  # earlier steps carried gradients: stay out of place
                term = torch.empty_like(a)
                _ops.is_accumulate(a, b, c, term, None, True)
                total = total + term
                if keep_log_weights:
                    log_weights.append(term)
                return None, None
            _ops.is_accumulate(a, b, c, total, log_w, first)
        if keep_log_weights:
            log_weights.append(log_w)
        return None, None

    # ---- t = 1 .. T-1 (inference.py:99-126) -------------------------------------------------
    for t in range(1, T):
        if smc:
            idx, x_res = close_step(draw_uniforms(t, dev))
            if keep_index:
                ancestors.append(idx)
            previous = ResampledHistory(history, idx, x_res, home)
        else:
            close_step(None)
            previous = history  # same list object as `history` (reference aliasing)
        q = proposal(previous_latents=previous, time=t, observations=observations)
        latent = state.sample(q, B, K)
        history += [latent]
        lq = state.log_prob(q, latent)
        lt = state.log_prob(
            transition(previous_latents=previous, time=t, previous_observations=observations[:t]), latent)
        le = state.log_prob(
            emission(latents=history, time=t, previous_observations=observations[:t]),
            state.expand_observation(observations[t], K))
        if keep_originals:
            originals.append(latent)
        pending = tuple(_as_f32_device(v) for v in (lt, le, lq))
    close_step(None)

    # ---- epilogue (inference.py:128-193) ----------------------------------------------------
    def back(t):
        return t if (t is None or t.device == home) else t.to(home)

    result = dict.fromkeys(("log_marginal_likelihood", "latents", "original_latents", "log_weight",
                            "log_weights", "ancestral_indices"))
    if smc:
        if return_log_marginal_likelihood:
            per_step = torch.stack(lses, dim=0) - _pymath.log(K)
            result["log_marginal_likelihood"] = back(per_step.sum(dim=0))
        if return_latents:
            result["latents"] = _trace_genealogy(originals, ancestors, home)
        if return_original_latents:
            result["original_latents"] = originals
        if return_log_weight:
            result["log_weight"] = back(log_weights[-1])
        if return_ancestral_indices:
            result["ancestral_indices"] = [back(_ops.widen_index(i)) for i in ancestors]
    else:
        if return_log_marginal_likelihood:
            result["log_marginal_likelihood"] = back(_ops.logsumexp_rows(total, quiet_flags) - _pymath.log(K))
        if return_latents:
            result["latents"] = originals
        if return_original_latents:
            raise RuntimeWarning("return_original_latents shouldn't be True for is")
        if return_log_weight:
            result["log_weight"] = back(total)
        if return_ancestral_indices:
            raise RuntimeWarning("return_ancestral_indices shouldn't be True for is")
    if return_log_weights:
        result["log_weights"] = [back(w) for w in log_weights]
    result["last_latent"] = latent
    if check_finite:
        _ops.raise_on_flags(flags)
    return result


@contextlib.contextmanager
def _scalars_by_fill_kernel():
    """torch.distributions turns Python-number parameters into tensors with torch.tensor(v, device=cuda) -- a
    pageable host-to-device copy, which a graph capture forbids.  While a GraphedInfer runs the user callables,
    torch.tensor(<python number>, device=<cuda>) is served by torch.full((), v, ...) instead: a fill kernel whose
    argument is baked into the graph.  Everything else goes to the real torch.tensor."""
    real = torch.tensor

    def tensor(data, *args, **kwargs):
        dev = kwargs.get("device")
        if isinstance(data, (bool, int, float)) and not args and dev is not None and torch.device(dev).type == "cuda":
            dtype = kwargs.get("dtype")
            if dtype is None:
                dtype = torch.bool if isinstance(data, bool) else (torch.int64 if isinstance(data, int) else torch.get_default_dtype())
            return torch.full((), data, dtype=dtype, device=dev)
        return real(data, *args, **kwargs)

    torch.tensor = tensor
    try:
        yield
    finally:
        torch.tensor = real


class GraphedInfer:
    """infer() on ANY user model, captured once as a CUDA graph and replayed per batch of observations:
    the T-step loop -- user callables, torch.distributions sampling and log-densities, the step kernel, the
    evidence reduction -- runs without per-step Python, launch or allocation cost (BASELINE config 1,
    B = 1, K = 100, T = 50: ~25 ms eager).

        g = GraphedInfer('smc', observations, initial, transition, emission, proposal, num_particles,
                         return_log_marginal_likelihood=True, return_latents=False)
        result = g()                      # the captured observations
        result = g(new_observations)      # same shapes; returns the same dict of (overwritten) tensors

    Requirements (those of CUDA graph capture): CUDA observations and model; callables free of host
    synchronisation -- construct distributions with validate_args=False (or
    torch.distributions.Distribution.set_default_validate_args(False)), no .item()/.cpu()/numpy inside
    (Python-number distribution parameters are fine: see _scalars_by_fill_kernel); shapes fixed; inference
    only (no autograd).  Resampling uniforms are drawn on the device with torch's
    graph-safe generator (torch.manual_seed controls them and the model's sampling), not from numpy's global
    RNG.  Non-finite weights are not raised inside a replay: call check() (one host sync) when you care."""

    def __init__(self, inference_algorithm, observations, initial, transition, emission, proposal, num_particles,
                 **infer_kwargs):
        for banned in ("uniforms", "check_finite", "_allow_fused"):
            if banned in infer_kwargs:
                raise ValueError("GraphedInfer sets %r itself" % banned)
        self._alg = inference_algorithm
        self._model = (initial, transition, emission, proposal)
        self._K = num_particles
        self._kwargs = infer_kwargs
        self.observations = [_map_tensors(lambda v: v.detach().clone(), o) for o in observations]
        probe = _first_tensor(self.observations[0])
        if not probe.is_cuda:
            raise ValueError("GraphedInfer needs CUDA observations")
        self._dev, self._T, self._B = probe.device, len(self.observations), probe.size(0)
        side = torch.cuda.Stream(device=self._dev)  # warm-up off the capture: lazy inits, function attributes
        side.wait_stream(torch.cuda.current_stream(self._dev))
        with torch.cuda.stream(side):
            self._run()
        torch.cuda.current_stream(self._dev).wait_stream(side)
        torch.cuda.synchronize(self._dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.result = self._run()

    def _run(self):
        u = None
        if self._alg == "smc" and self._T > 1:
            u = torch.rand(self._T - 1, self._B, dtype=torch.float64, device=self._dev)
        with torch.no_grad(), _scalars_by_fill_kernel():
            return infer(self._alg, self.observations, *self._model, self._K, uniforms=u, check_finite=False,
                         _allow_fused=False, **self._kwargs)

    def __call__(self, observations=None):
        if observations is not None:
            if len(observations) != self._T:
                raise ValueError("expected %d observations, got %d" % (self._T, len(observations)))
            for dst, src in zip(self.observations, observations):
                if isinstance(dst, dict):
                    for name in dst:
                        dst[name].copy_(src[name], non_blocking=True)
                else:
                    dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.result

    def check(self):
        """Raise FloatingPointError if the last replay produced non-finite log-weights / evidence."""
        for key in ("log_marginal_likelihood", "log_weight"):
            v = self.result.get(key)
            if v is not None and not bool(torch.isfinite(v).all()):
                raise FloatingPointError("log_weight contains nan element(s)")


def _trace_genealogy(latents, ancestors32, home, sorted_rows=True):
    """Back-trace with int32 device indices: latents[t] re-indexed by the composed ancestry of the
    final particles (inference.py:196-231).  sorted_rows: every row of every index table is
    non-decreasing (true for the systematic-resampling kernel's output, and then for compositions of
    such tables); selects the deterministic sorted gather backward.  Indices of unknown origin must pass
    False: the sorted backward is wrong for unsorted rows."""
    T = len(latents)
    probe = _first_tensor(latents[0])
    B, K = probe.shape[:2]
    dev = ancestors32[0].device if ancestors32 else _ops.device()
    out = [None] * T
    cursor = None  # None = identity
    for t in range(T - 1, -1, -1):
        if cursor is None:
            out[t] = _map_tensors(lambda v: v.clone(), latents[t])
        else:
            def gather(v, cursor=cursor):
                r = _ops.gather(_ops.to_device(v), cursor, sorted_rows=sorted_rows)
                return r if v.is_cuda else r.to(home)
            out[t] = _map_tensors(gather, latents[t])
        if t > 0:
            nxt = ancestors32[t - 1]
            cursor = nxt if cursor is None else _ops.compose_index(nxt, cursor)
    return out


def get_resampled_latents(latents, ancestral_indices):
    """Re-index every latents[t] by the ancestry of the final particles.

    latents: list (length T) of tensors [batch, particles, ...] or dicts thereof;
    ancestral_indices: list (length T-1, may be empty) of integer tensors [batch, particles].
    Returns a list of the same element type as latents."""
    assert len(ancestral_indices) == len(latents) - 1
    probe = _first_tensor(latents[0])
    home = probe.device
    idx32 = []
    for a in ancestral_indices:
        a = _ops.to_device(a)
        idx32.append(a if a.dtype == torch.int32 else _ops.narrow_index(a.long().contiguous()))
    # user-supplied indices may come from any resampler (multinomial, permuted ...): the gather backward
    # must not assume sorted rows (the reference's torch.gather backward is correct for any index)
    return _trace_genealogy(list(latents), idx32, home, sorted_rows=False)


def sample_ancestral_index(log_weight, *, uniforms=None, resampling_mode=None):
    """Systematic resampling: log_weight [batch, particles] -> LongTensor [batch, particles] of
    zero-based ancestor indices on log_weight's device.  One uniform per row, drawn from numpy's
    global RNG like the reference (inference.py:250) unless ``uniforms`` ([batch] float64) is given.
    Raises FloatingPointError if log_weight contains NaN (inference.py:244-245)."""
    B, K = log_weight.size()
    if uniforms is None:
        uniforms = np.random.uniform(size=[B, 1])
    lw = _ops.to_device(log_weight.detach(), torch.float32)
    flags = _ops.new_flags(lw.device)
    u = _ops.uniforms_to_device(uniforms, B, lw.device)
    _, _, idx, _ = _ops.smc_step(lw, None, None, u, None, flags, resampling_mode, resample=True)
    _ops.raise_on_flags(flags)
    out = _ops.widen_index(idx)
    return out if log_weight.is_cuda else out.to(log_weight.device)
