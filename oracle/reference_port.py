"""oracle/reference_port.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A CPU restatement (torch-CPU + numpy, like the reference itself) of aesmc's inference driver, used
as the end-to-end checker for aesmc_b200.inference.infer / losses.get_loss and as the timed
"reference" arm of bench.py.  It deliberately reproduces the reference's observable behaviour,
including its quirks (SURVEY.md Q1 O(T^2) history gather, Q2 importance-sampling list aliasing),
and its cost profile (per-row np.digitize loop, host-side numpy resampling).

Restates (paths relative to /root/reference):
    infer                      aesmc/inference.py:8-193
    get_resampled_latents      aesmc/inference.py:196-231
    sample_ancestral_index     aesmc/inference.py:234-269
    lognormexp / exponentiate_and_normalize (numpy branch)   aesmc/math.py:6-51
    state.sample/log_prob/resample/expand_observation        aesmc/state.py:61-203
    get_loss                   aesmc/losses.py:5-65
    log_ess                    aesmc/statistics.py:79-91

The only difference from the reference's call surface is the optional ``uniforms`` argument
(list/array [T-1, B]): when given, step t uses uniforms[t-1] instead of drawing
np.random.uniform(size=[B, 1]) (inference.py:250), so both sides of a parity test see identical
randomness.  When omitted, the numpy global RNG is consumed exactly like the reference does.

Parity pinning: tests/test_oracle_pinning.py compares this port with tests/golden/*.npz, which
tests/golden/make_golden.py recorded from the unmodified reference.
"""
import numpy as np
import torch

_MODE_NAMES = ("NOT_EXPANDED", "BATCH_EXPANDED", "FULLY_EXPANDED")


def _mode_name(dist, batch_size, num_particles):
    """state.py:20-58 without the warnings; compares enum members by name so that distributions
    tagged by either package's BatchShapeMode work."""
    tag = getattr(dist, "batch_shape_mode", None)
    if tag is not None:
        return tag.name
    shape = tuple(dist.batch_shape)
    if len(shape) == 0:
        return "NOT_EXPANDED"
    if len(shape) == 1:
        return "BATCH_EXPANDED" if shape[0] == batch_size else "NOT_EXPANDED"
    if shape[0] != batch_size:
        return "NOT_EXPANDED"
    return "FULLY_EXPANDED" if shape[1] == num_particles else "BATCH_EXPANDED"


def draw(dist, batch_size, num_particles):
    """state.py:61-111"""
    if isinstance(dist, dict):
        return {k: draw(v, batch_size, num_particles) for k, v in dist.items()}
    if torch.is_tensor(dist):
        return dist
    mode = _mode_name(dist, batch_size, num_particles)
    shape = {"NOT_EXPANDED": (batch_size, num_particles), "BATCH_EXPANDED": (num_particles,),
             "FULLY_EXPANDED": ()}[mode]
    if not dist.has_rsample:
        raise ValueError("distribution not reparameterizable")
    out = dist.rsample(sample_shape=shape)
    return out.transpose(0, 1) if mode == "BATCH_EXPANDED" else out


def logp(dist, value):
    """state.py:114-155 (tensor distributions only; the dict branch of the reference raises
    NameError)."""
    nb = value.dim() - len(dist.event_shape)
    db = len(dist.batch_shape)
    if nb == db or nb - 2 == db:
        dist._validate_sample(value)
        lp = dist.log_prob(value)
    elif nb - 1 == db:
        lp = dist.log_prob(value.transpose(0, 1)).transpose(0, 1)
    else:
        raise RuntimeError("Incompatible distribution.batch_shape ({}) and value.shape ({}).".format(
            dist.batch_shape, value.shape))
    return lp.reshape(value.size(0), value.size(1), -1).sum(dim=2)


def widen_obs(obs, num_particles):
    """state.py:186-203"""
    if isinstance(obs, dict):
        return {k: widen_obs(v, num_particles) for k, v in obs.items()}
    return obs.unsqueeze(1).expand(obs.size(0), num_particles, *obs.shape[1:])


def gather_particles(value, index):
    """state.py:158-183"""
    if isinstance(value, dict):
        return {k: gather_particles(v, index) for k, v in value.items()}
    assert index.size() == value.size()[:2]
    ix = index.reshape(index.shape + (1,) * (value.dim() - 2)).expand_as(value)
    return torch.gather(value, 1, ix)


def np_logsumexp_rows(a):
    """scipy 1.18.1 special/_logsumexp.py:201-249 for a real 2-D array reduced over axis 1 with
    keepdims (the only way math.py:22 is reached from the hot path), written with the same numpy
    primitives scipy uses so that the bits follow the installed numpy."""
    with np.errstate(all="ignore"):
        a = np.asarray(a)
        amax = np.max(a, axis=1, keepdims=True)
        is_max = a == amax
        rest = np.where(is_max, -np.inf, a).astype(a.dtype)
        m = np.sum(is_max.astype(a.dtype), axis=1, keepdims=True, dtype=a.dtype)
        e = np.exp(rest - amax)
        s = np.sum(e, axis=1, keepdims=True, dtype=e.dtype)
        s = np.where(s == 0, s, s / m)
        out = np.log1p(s) + np.log(m) + amax
        direct = np.log(np.sum(np.exp(a), axis=1, keepdims=True))
        return np.where(np.isfinite(out), out, direct)


def systematic_ancestors(log_weight, uniforms=None):
    """inference.py:234-269.  log_weight: torch [B, K]; uniforms: optional [B] float64."""
    if int(torch.sum(log_weight != log_weight)) != 0:
        raise FloatingPointError("log_weight contains nan element(s)")
    B, K = log_weight.shape
    if uniforms is None:
        uniforms = np.random.uniform(size=[B, 1])
    u = np.asarray(uniforms, dtype=np.float64).reshape(B, 1)
    pos = (u + np.arange(0, K)) / K
    lw = log_weight.detach().cpu().numpy()
    with np.errstate(all="ignore"):
        w = np.exp(lw - np_logsumexp_rows(lw))
        cdf = np.cumsum(w, axis=1)
        cdf = cdf / np.max(cdf, axis=1, keepdims=True)
    out = np.zeros([B, K])
    for b in range(B):
        out[b] = np.digitize(pos[b], cdf[b])
    return torch.from_numpy(out).long()


def trace_genealogy(latents, ancestral_indices):
    """inference.py:196-231"""
    assert len(ancestral_indices) == len(latents) - 1
    probe = next(iter(latents[0].values())) if isinstance(latents[0], dict) else latents[0]
    B, K = probe.shape[:2]
    cursor = torch.arange(K).long().unsqueeze(0).expand(B, K)
    out = [None] * len(latents)
    for t in range(len(latents) - 1, -1, -1):
        out[t] = gather_particles(latents[t], cursor)
        if t > 0:
            cursor = torch.gather(ancestral_indices[t - 1], 1, cursor)
    return out


def infer(inference_algorithm, observations, initial, transition, emission, proposal, num_particles,
          return_log_marginal_likelihood=False, return_latents=True, return_original_latents=False,
          return_log_weight=True, return_log_weights=False, return_ancestral_indices=False,
          uniforms=None):
    """inference.py:8-193 (same arguments, same result dict)."""
    if inference_algorithm not in ("is", "smc"):
        raise ValueError("inference_algorithm must be either is or smc. currently = {}".format(
            inference_algorithm))
    smc = inference_algorithm == "smc"
    first = observations[0]
    B = next(iter(first.values())).size(0) if isinstance(first, dict) else first.size(0)
    K = num_particles
    keep_latents = return_original_latents or return_latents
    originals, ancestors, log_weights = [], [], []

    q = proposal(time=0, observations=observations)
    x = draw(q, B, K)
    history = [x]
    lq = logp(q, x)
    lp0 = logp(initial(), x)
    le = logp(emission(latents=history, time=0), widen_obs(observations[0], K))
    if keep_latents:
        originals.append(x)
    log_weights.append(lp0 + le - lq)

    for t in range(1, len(observations)):
        if smc:
            ut = None if uniforms is None else uniforms[t - 1]
            ancestors.append(systematic_ancestors(log_weights[-1], ut))
            # Q1: every stored latent is gathered with the newest index only
            previous = [gather_particles(h, ancestors[-1]) for h in history]
        else:
            previous = history  # Q2: alias, mutated by the += below
        q = proposal(previous_latents=previous, time=t, observations=observations)
        x = draw(q, B, K)
        history += [x]
        lq = logp(q, x)
        lt = logp(transition(previous_latents=previous, time=t,
                             previous_observations=observations[:t]), x)
        le = logp(emission(latents=history, time=t, previous_observations=observations[:t]),
                  widen_obs(observations[t], K))
        if keep_latents:
            originals.append(x)
        log_weights.append(lt + le - lq)

    result = dict.fromkeys(("log_marginal_likelihood", "latents", "original_latents", "log_weight",
                            "log_weights", "ancestral_indices"))
    if smc:
        if return_log_marginal_likelihood:
            per_step = torch.logsumexp(torch.stack(log_weights, dim=0), dim=2) - np.log(K)
            result["log_marginal_likelihood"] = torch.sum(per_step, dim=0)
        if return_latents:
            result["latents"] = trace_genealogy(originals, ancestors)
        if return_original_latents:
            result["original_latents"] = originals
        if return_log_weight:
            result["log_weight"] = log_weights[-1]
        if return_ancestral_indices:
            result["ancestral_indices"] = ancestors
    else:
        total = None
        if return_log_marginal_likelihood or return_log_weight:
            total = torch.sum(torch.stack(log_weights, dim=0), dim=0)
        if return_log_marginal_likelihood:
            result["log_marginal_likelihood"] = torch.logsumexp(total, dim=1) - np.log(K)
        if return_latents:
            result["latents"] = originals
        if return_original_latents:
            raise RuntimeWarning("return_original_latents shouldn't be True for is")
        if return_log_weight:
            result["log_weight"] = total
        if return_ancestral_indices:
            raise RuntimeWarning("return_ancestral_indices shouldn't be True for is")
    if return_log_weights:
        result["log_weights"] = log_weights
    result["last_latent"] = x
    return result


def get_loss(observations, num_particles, algorithm, initial, transition, emission, proposal,
             uniforms=None):
    """losses.py:5-65"""
    algo = {"iwae": "is", "aesmc": "smc"}[algorithm]
    res = infer(algo, observations, initial, transition, emission, proposal, num_particles,
                return_log_marginal_likelihood=True, return_latents=False,
                return_original_latents=False, return_log_weight=False, return_log_weights=False,
                return_ancestral_indices=False, uniforms=uniforms)
    return -torch.mean(res["log_marginal_likelihood"])


def log_ess(log_weight):
    """statistics.py:79-91"""
    dim = 1 if log_weight.dim() == 2 else 0
    return 2 * torch.logsumexp(log_weight, dim=dim) - torch.logsumexp(2 * log_weight, dim=dim)


def weighted_expectation(value, log_weight, f):
    """statistics.py:7-44: sum_k w_k f(x_k), accumulated particle by particle in index order."""
    w = torch.exp(log_weight - torch.logsumexp(log_weight, dim=1, keepdim=True))
    acc = None
    for p in range(w.size(1)):
        fx = f(value[:, p])
        wp = w[:, p].reshape((-1,) + (1,) * (fx.dim() - 1))
        term = wp.expand_as(fx) * fx
        acc = term if acc is None else acc + term
    return acc
