"""Normalisation of log-weights over one axis.

Mirrors aesmc/math.py of the reference (lognormexp :6-30, exponentiate_and_normalize :33-51): accepts
a torch.Tensor or a numpy.ndarray of any rank, normalises over ``dim`` and returns the same type.
The arithmetic runs in the library's row-reduction kernel (aesmc_lognormexp_f32 /
aesmc_logsumexp_f64); numpy arrays and CPU tensors are staged through the GPU.
"""
import numpy as np
import torch

from . import _ops


def _rows_view(t, dim):
    """Move ``dim`` last and flatten the rest: returns ([rows, n] contiguous, restore())."""
    dim = dim % t.dim()
    moved = t.movedim(dim, -1)
    shape = moved.shape
    flat = moved.reshape(-1, shape[-1]).contiguous()

    def restore(out):
        return out.reshape(shape).movedim(-1, dim)

    return flat, restore


def _normalise(values, dim, exponentiate):
    as_numpy = isinstance(values, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(values)) if as_numpy else values
    if t.dim() == 0:
        raise ValueError("values must have at least one dimension")
    if not t.is_floating_point():
        # integer input (the reference's tests pass np.array([1, 2, 3])): scipy / numpy promote to float64
        t = t.double()
    home, dtype = t.device, t.dtype
    work = _ops.to_device(t.detach() if as_numpy else t)
    flat, restore = _rows_view(work, dim)
    if flat.dtype == torch.float32 and not (torch.is_grad_enabled() and flat.requires_grad):
        out = _ops.lognormexp_rows(flat, exponentiate)
    else:
        # float64 inputs and differentiable calls: kernel logsumexp (with its autograd), torch glue
        if flat.dtype not in (torch.float32, torch.float64):
            flat = flat.float()
        out = flat - _ops.logsumexp_rows(flat).unsqueeze(1)
        if exponentiate:
            out = torch.exp(out)
    out = restore(out).to(dtype)
    if as_numpy:
        return out.cpu().numpy()
    return out if values.is_cuda else out.to(home)


def lognormexp(values, dim=0):
    """values - logsumexp(values, dim, keepdim=True); torch.Tensor or np.ndarray in, same type out."""
    return _normalise(values, dim, exponentiate=False)


def exponentiate_and_normalize(values, dim=0):
    """softmax of ``values`` over ``dim``; torch.Tensor or np.ndarray in, same type out."""
    return _normalise(values, dim, exponentiate=True)
