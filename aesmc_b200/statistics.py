"""Post-hoc estimators on weighted particle sets, and the prior sampler.

Mirrors aesmc/statistics.py of the reference: empirical_expectation :7-44, empirical_mean :47-60,
empirical_variance :63-76, log_ess :79-91, ess :94-104, sample_from_prior :108-162.

log_ess/ess and the moments behind empirical_mean/empirical_variance run in single-pass row kernels
(aesmc_log_ess_f32/f64, aesmc_weighted_moments_f32); the general empirical_expectation keeps the
reference's particle-by-particle accumulation because ``f`` is arbitrary user code.
"""
import torch

from . import _ops
from . import math
from . import state


def empirical_expectation(value, log_weight, f):
    """sum_k w_k f(value[:, k]) with w = softmax(log_weight, dim=1).

    value [batch, particles, ...]; log_weight [batch, particles]; f maps [batch, ...] -> [batch, ...].
    Accumulated in particle order exactly as the reference does (statistics.py:27-42)."""
    assert value.size()[:2] == log_weight.size()
    weights = math.exponentiate_and_normalize(log_weight, dim=1)
    total = None
    for k in range(weights.size(1)):
        fk = f(value[:, k])
        wk = weights[:, k].reshape((-1,) + (1,) * (fk.dim() - 1))
        term = wk.expand_as(fk) * fk
        total = term if total is None else total + term
    return total


def _moments(value, log_weight, want_second):
    home = value.device
    x = _ops.to_device(value.detach(), torch.float32)
    lw = _ops.to_device(log_weight.detach(), torch.float32)
    B, K = lw.shape
    mean, second = _ops.weighted_moments(x.reshape(B, K, -1), lw, want_second)
    shape = (B,) + tuple(value.shape[2:])
    mean = mean.reshape(shape).to(value.dtype)
    second = None if second is None else second.reshape(shape).to(value.dtype)
    if not value.is_cuda:
        mean = mean.to(home)
        second = None if second is None else second.to(home)
    return mean, second


def _fast_path(value, log_weight):
    differentiable = torch.is_grad_enabled() and (value.requires_grad or log_weight.requires_grad)
    return (not differentiable) and value.is_floating_point() and value.numel() > 0


def empirical_mean(value, log_weight):
    """Weighted particle mean, [batch, ...]."""
    assert value.size()[:2] == log_weight.size()
    if _fast_path(value, log_weight):
        return _moments(value, log_weight, False)[0]
    return empirical_expectation(value, log_weight, lambda x: x)


def empirical_variance(value, log_weight):
    """Weighted particle variance E[x^2] - E[x]^2, [batch, ...]."""
    assert value.size()[:2] == log_weight.size()
    if _fast_path(value, log_weight):
        mean, second = _moments(value, log_weight, True)
        return second - mean ** 2
    return empirical_expectation(value, log_weight, lambda x: x ** 2) - empirical_mean(value, log_weight) ** 2


def log_ess(log_weight):
    """log effective sample size 2*lse(lw) - lse(2*lw): [batch, particles] -> [batch]; [particles] -> 0-d."""
    squeeze = log_weight.dim() == 1
    lw = log_weight.unsqueeze(0) if squeeze else log_weight
    if torch.is_grad_enabled() and lw.requires_grad:
        dev_lw = _ops.to_device(lw)
        if dev_lw.dtype not in (torch.float32, torch.float64):
            dev_lw = dev_lw.float()
        out = 2 * _ops.logsumexp_rows(dev_lw) - _ops.logsumexp_rows(2 * dev_lw)
    else:
        dev_lw = _ops.to_device(lw.detach())
        if dev_lw.dtype not in (torch.float32, torch.float64):
            dev_lw = dev_lw.float()
        out = _ops.log_ess_rows(dev_lw)
    out = out.to(log_weight.dtype) if log_weight.is_floating_point() else out
    if not log_weight.is_cuda:
        out = out.to(log_weight.device)
    return out.squeeze(0) if squeeze else out


def ess(log_weight):
    """Effective sample size exp(log_ess)."""
    return torch.exp(log_ess(log_weight))


def sample_from_prior(initial, transition, emission, num_timesteps, batch_size):
    """Ancestral sampling of (latents, observations) from the generative model; lists of length
    num_timesteps of tensors [batch_size, ...] (or dicts).  Runs the model callables only (torch)."""
    latents, observations = [], []
    for t in range(num_timesteps):
        if t == 0:
            prior = initial()
        else:
            prior = transition(previous_latents=latents, time=t, previous_observations=observations[:t])
        latents.append(state.sample(prior, batch_size, 1))
        observations.append(state.sample(
            emission(latents=latents, time=t, previous_observations=observations[:t]), batch_size, 1))

    def drop_particle_axis(v):
        if isinstance(v, dict):
            return {name: drop_particle_axis(x) for name, x in v.items()}
        return v.squeeze(1)

    return [drop_particle_axis(v) for v in latents], [drop_particle_axis(v) for v in observations]
