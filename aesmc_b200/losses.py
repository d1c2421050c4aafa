"""Training objectives (mirrors aesmc/losses.py:5-65 of the reference)."""
import torch

from . import inference

_ALGORITHMS = {"iwae": "is", "aesmc": "smc"}


def get_loss(observations, num_particles, algorithm, initial, transition, emission, proposal, **infer_kwargs):
    """Negative batch-mean evidence lower bound, differentiable w.r.t. the parameters of the four
    callables.  algorithm: 'iwae' (importance sampling) or 'aesmc' (SMC); any other string raises
    UnboundLocalError like the reference (losses.py:45-48 has no else branch).  Extra keyword
    arguments (uniforms=, resampling_mode=, check_finite=) are forwarded to inference.infer."""
    if algorithm not in _ALGORITHMS:
        raise UnboundLocalError(
            "cannot access local variable 'inference_algorithm' where it is not associated with a value "
            "(algorithm must be 'iwae' or 'aesmc', got {!r})".format(algorithm))
    result = inference.infer(inference_algorithm=_ALGORITHMS[algorithm], observations=observations,
                             initial=initial, transition=transition, emission=emission, proposal=proposal,
                             num_particles=num_particles, return_log_marginal_likelihood=True,
                             return_latents=False, return_original_latents=False, return_log_weight=False,
                             return_log_weights=False, return_ancestral_indices=False, **infer_kwargs)
    return -torch.mean(result["log_marginal_likelihood"])
