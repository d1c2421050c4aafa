"""Particle-state helpers: the boundary between user-supplied torch.distributions and the
[batch, particle, ...] layout of the SMC core.

Mirrors the public surface of the reference's aesmc/state.py (BatchShapeMode :6-9,
set_/get_batch_shape_mode :12-58, sample :61-111, log_prob :114-155, resample :158-183,
expand_observation :186-203).  sample / log_prob / expand_observation stay in torch -- they ARE the
user-model boundary -- while ``resample`` runs the library's ancestral-gather kernel.
"""
import enum
import warnings

import torch

from . import _ops


class BatchShapeMode(enum.Enum):
    """How a distribution's batch_shape relates to [batch_size, num_particles]."""
    NOT_EXPANDED = 0    # batch_shape == [...]
    BATCH_EXPANDED = 1  # batch_shape == [batch_size, ...]
    FULLY_EXPANDED = 2  # batch_shape == [batch_size, num_particles, ...]


def _canonical(mode):
    """Accept BatchShapeMode members of this package or of the reference package (user models written
    against `aesmc.state.BatchShapeMode` stay drop-in): members are matched by name."""
    if isinstance(mode, BatchShapeMode):
        return mode
    name = getattr(mode, "name", None)
    if name in BatchShapeMode.__members__:
        return BatchShapeMode[name]
    raise ValueError("batch_shape_mode {} not supported".format(mode))


def set_batch_shape_mode(distribution, batch_shape_mode):
    """Tag ``distribution`` with an explicit BatchShapeMode and return it."""
    distribution.batch_shape_mode = batch_shape_mode
    return distribution


def get_batch_shape_mode(distribution, batch_size=None, num_particles=None):
    """Explicit tag if present, else inferred from batch_shape; ambiguous inferences (a leading
    dimension that happens to equal batch_size) emit a RuntimeWarning, as in the reference."""
    tag = getattr(distribution, "batch_shape_mode", None)
    if tag is not None:
        return tag
    shape = tuple(distribution.batch_shape)
    guess = BatchShapeMode.NOT_EXPANDED
    ambiguous = False
    if len(shape) >= 1 and shape[0] == batch_size:
        ambiguous = True
        guess = BatchShapeMode.BATCH_EXPANDED
        if len(shape) >= 2 and shape[1] == num_particles:
            guess = BatchShapeMode.FULLY_EXPANDED
    if ambiguous:
        warnings.warn(
            "Inferred batch_shape_mode ({}) of distribution ({}) might be wrong given its batch_shape "
            "({}), batch_size ({}) and num_particles ({}). Consider specifying the batch_shape_mode "
            "explicitly.".format(guess, distribution, distribution.batch_shape, batch_size, num_particles),
            RuntimeWarning)
    return guess


_SAMPLE_SHAPE = {
    BatchShapeMode.NOT_EXPANDED: lambda b, k: (b, k),
    BatchShapeMode.BATCH_EXPANDED: lambda b, k: (k,),
    BatchShapeMode.FULLY_EXPANDED: lambda b, k: (),
}


def sample(distribution, batch_size, num_particles):
    """Reparameterised draw of shape [batch_size, num_particles, ...] (dicts map over values;
    tensors pass through)."""
    if isinstance(distribution, dict):
        return {name: sample(d, batch_size, num_particles) for name, d in distribution.items()}
    if isinstance(distribution, torch.Tensor):
        return distribution
    if not isinstance(distribution, torch.distributions.Distribution):
        raise AttributeError(
            "distribution must be a dict or a torch.distributions.Distribution. Got: {}".format(distribution))
    mode = _canonical(get_batch_shape_mode(distribution, batch_size, num_particles))
    if not distribution.has_rsample:
        raise ValueError("distribution not reparameterizable")
    drawn = distribution.rsample(sample_shape=_SAMPLE_SHAPE[mode](batch_size, num_particles))
    # BATCH_EXPANDED draws come out as [num_particles, batch_size, ...]
    return drawn.transpose(0, 1) if mode is BatchShapeMode.BATCH_EXPANDED else drawn


def log_prob(distribution, value):
    """log density of ``value`` [batch_size, num_particles, ...] reduced to [batch_size, num_particles].
    Dicts of distributions sum their members' log-probs (the reference's dict branch is unreachable:
    it raises NameError)."""
    if isinstance(distribution, dict):
        return torch.stack([log_prob(d, value[name]) for name, d in distribution.items()], dim=0).sum(dim=0)
    if not isinstance(distribution, torch.distributions.Distribution):
        raise AttributeError(
            "distribution must be a dict or a torch.distributions.Distribution. Got: {}".format(distribution))
    lead = value.dim() - len(distribution.event_shape)  # number of batch-like dims of value
    have = len(distribution.batch_shape)
    if lead == 2 and have in (0, 1, 2) and type(distribution) is torch.distributions.Normal and value.is_cuda:
        # scalar Normal over a [batch, particle] table: one kernel, bit-identical to torch's six
        if getattr(distribution, "_validate_args", True):
            distribution._validate_sample(value if have != 1 else value.transpose(0, 1))
        fast = _ops.normal_log_prob(distribution, value)
        if fast is not None:
            return fast
    if lead == 2 and have == 2 and value.is_cuda and type(distribution) is torch.distributions.Independent:
        # Independent(Normal) with vector latents (BASELINE config 3): the elementwise part in one kernel
        if getattr(distribution, "_validate_args", True):
            distribution._validate_sample(value)
        fast = _ops.independent_normal_log_prob(distribution, value)
        if fast is not None:
            return fast
    if lead == have or lead == have + 2:
        # The reference validates the sample unconditionally (state.py:142), which costs a host
        # synchronisation per call on CUDA tensors; here a distribution built with
        # validate_args=False (or under Distribution.set_default_validate_args(False)) is trusted.
        if getattr(distribution, "_validate_args", True):
            distribution._validate_sample(value)
        lp = distribution.log_prob(value)
    elif lead == have + 1:  # batch-expanded distribution: particles must lead for broadcasting
        lp = distribution.log_prob(value.transpose(0, 1)).transpose(0, 1)
    else:
        raise RuntimeError("Incompatible distribution.batch_shape ({}) and value.shape ({}).".format(
            distribution.batch_shape, value.shape))
    if lp.dim() == 2:  # no event / extra dims to reduce: skip the [B, K, 1] sum (a full copy of the table)
        return lp
    return lp.reshape(value.size(0), value.size(1), -1).sum(dim=2)


def resample(value, ancestral_index, *, check_range=True):
    """Ancestral gather without side effects: out[b, k, ...] = value[b, ancestral_index[b, k], ...].

    value: tensor [batch_size, num_particles, ...] or dict thereof; ancestral_index: integer tensor
    [batch_size, num_particles] (arbitrary order).  Runs aesmc_gather_bytes on the GPU; CPU inputs are
    staged through the device and returned on the CPU.  An index outside [0, num_particles) raises
    IndexError (torch.gather raises on the CPU and device-asserts on CUDA); the check reads a device flag
    word, one host synchronisation -- callers inside a CUDA-graph capture or a latency-critical loop pass
    check_range=False (out-of-range indices are then clamped)."""
    if isinstance(value, dict):
        return {name: resample(v, ancestral_index, check_range=check_range) for name, v in value.items()}
    if not torch.is_tensor(value):
        raise AttributeError("value must be a dict or a torch.Tensor. Got: {}".format(value))
    assert ancestral_index.size() == value.size()[:2]
    home = value.device
    v = _ops.to_device(value)
    idx = _ops.to_device(ancestral_index)
    if idx.dtype not in (torch.int32, torch.int64):
        idx = idx.long()
    flags = _ops.new_flags(v.device)
    out = _ops.gather(v, idx, sorted_rows=False, flags=flags)
    if check_range and not torch.cuda.is_current_stream_capturing():
        _ops.raise_on_flags(flags)
    if not value.is_cuda:
        out = out.to(home)
    return out


def expand_observation(observation, num_particles):
    """[batch_size, ...] -> broadcast view [batch_size, num_particles, ...] (dicts map over values)."""
    if isinstance(observation, dict):
        return {name: expand_observation(o, num_particles) for name, o in observation.items()}
    return observation.unsqueeze(1).expand(observation.size(0), num_particles, *observation.shape[1:])
