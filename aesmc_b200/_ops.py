"""Tensor-level wrappers and autograd Functions over the C ABI (include/aesmc_b200.h).

Everything here enqueues kernels of libaesmc_b200.so on torch's current CUDA stream; torch is used
for device memory, streams and autograd bookkeeping only.  Tensors that arrive on the CPU (or numpy
arrays, where the reference API accepts them) are staged through the GPU and the result is returned
on the caller's device: there is no host implementation of any of these operations.
"""
import numpy as np
import torch

from . import _lib

_RESAMPLING_MODES = {"exact": _lib.MODE_EXACT, "fast": _lib.MODE_FAST}
_default_mode = "exact"


def set_resampling_mode(mode):
    """'exact' (default): reference-order arithmetic, ancestor indices bit-identical to the
    reference's numpy path.  'fast': parallel scan / approximate exp; statistically equivalent."""
    global _default_mode
    if mode not in _RESAMPLING_MODES:
        raise ValueError("resampling mode must be 'exact' or 'fast', got {!r}".format(mode))
    _default_mode = mode


def get_resampling_mode():
    return _default_mode


def mode_code(mode=None):
    return _RESAMPLING_MODES[_default_mode if mode is None else mode]


def device():
    _lib.require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def to_device(t, dtype=None):
    """Contiguous CUDA view/copy of a tensor (staging CPU inputs onto the GPU)."""
    if not t.is_cuda:
        t = t.to(device(), non_blocking=True)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def new_flags(dev=None):
    return torch.zeros(1, dtype=torch.int32, device=dev or device())


def raise_on_flags(flags, where="log_weight"):
    """One host read of the device flag word; maps to the reference's exceptions."""
    bits = int(flags.item())
    if bits & _lib.FLAG_NAN:
        raise FloatingPointError("log_weight contains nan element(s)")  # inference.py:244-245
    if bits & _lib.FLAG_DEGENERATE:
        raise FloatingPointError(
            "{}: a row has no finite positive normaliser (all -inf or +inf); the reference would "
            "silently emit out-of-range ancestor indices here".format(where))
    if bits & _lib.FLAG_INDEX_RANGE:
        raise IndexError("ancestral_index out of range [0, num_particles)")


# ------------------------------------------------------------------------------------------------
# fused SMC step
# ------------------------------------------------------------------------------------------------
class _SMCStep(torch.autograd.Function):
    """log_w = (a + b) - c ; lse = logsumexp_k log_w ; idx = systematic ancestors(log_w, u) ;
    x_out = x[idx]   (inference.py:97-104,125-126,130,234-269; state.py:158-183)."""

    @staticmethod
    def forward(ctx, a, b, c, x, u, flags, mode, resample):
        B, K = a.shape
        log_w = torch.empty_like(a)
        lse = torch.empty(B, dtype=torch.float32, device=a.device)
        idx = torch.empty((B, K), dtype=torch.int32, device=a.device) if resample else None
        x_out = None
        D = 1
        if resample and x is not None:
            D = x[0, 0].numel()
            x_out = torch.empty_like(x)
        ws_bytes = int(_lib.load().aesmc_smc_step_workspace_bytes(B, K))  # > 0: multi-CTA path (large K)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=a.device) if ws_bytes else None
        _lib.call("aesmc_smc_step_ws_f32", _lib.ptr(a), _lib.ptr(b), _lib.ptr(c), _lib.ptr(u), B, K,
                  _lib.ptr(log_w), _lib.ptr(lse), _lib.ptr(idx), _lib.ptr(x if x_out is not None else None),
                  _lib.ptr(x_out), D, _lib.ptr(flags), mode, _lib.ptr(ws), ws_bytes)
        ctx.save_for_backward(log_w, lse, idx)
        ctx.has = (b is not None, c is not None, x_out is not None)
        ctx.x_shape = None if x is None else x.shape
        if idx is not None:
            ctx.mark_non_differentiable(idx)
        outs = (log_w, lse, idx if idx is not None else torch.empty(0, device=a.device),
                x_out if x_out is not None else torch.empty(0, device=a.device))
        return outs

    @staticmethod
    def backward(ctx, g_log_w, g_lse, _g_idx, g_x):
        log_w, lse, idx = ctx.saved_tensors
        has_b, has_c, has_x = ctx.has
        B, K = log_w.shape
        g_pos = g_neg = None
        if g_log_w is not None or g_lse is not None:
            g_log_w = None if g_log_w is None else g_log_w.contiguous()
            g_lse = None if g_lse is None else g_lse.contiguous()
            g_pos = torch.empty_like(log_w)
            g_neg = torch.empty_like(log_w) if has_c else None
            _lib.call("aesmc_step_bwd_f32", _lib.ptr(log_w), _lib.ptr(lse), _lib.ptr(g_log_w),
                      _lib.ptr(g_lse), B, K, _lib.ptr(g_pos), _lib.ptr(g_neg))
        gx = None
        if has_x and g_x is not None and ctx.needs_input_grad[3]:
            g_x = g_x.contiguous()
            gx = torch.empty_like(g_x)
            D = g_x[0, 0].numel()
            _lib.call("aesmc_gather_bwd_f32", _lib.ptr(g_x), _lib.ptr(idx), 0, B, K, D, _lib.ptr(gx), 1)
        return g_pos, (g_pos if has_b else None), g_neg, gx, None, None, None, None


def smc_step(a, b, c, u, x, flags, mode=None, resample=True):
    """Run one fused step.  a [B,K] float32 CUDA; b, c same or None; u [B] float64 (needed iff
    resample); x [B,K,...] float32 or None.  Returns (log_w, lse, idx_int32 | None, x_resampled | None)."""
    log_w, lse, idx, x_out = _SMCStep.apply(a, b, c, x, u, flags, mode_code(mode), resample)
    return log_w, lse, (idx if resample else None), (x_out if (resample and x is not None) else None)


def resample_from_weights(w, u, flags, mode=None):
    """Systematic ancestors from normalised weights [B,K] (cumulative sum onwards); int32 [B,K]."""
    B, K = w.shape
    idx = torch.empty((B, K), dtype=torch.int32, device=w.device)
    _lib.call("aesmc_resample_from_weights_f32", _lib.ptr(w), _lib.ptr(u), B, K, _lib.ptr(idx), _lib.ptr(flags),
              mode_code(mode))
    return idx


def resample_from_cdf(cdf, u, flags):
    """Systematic ancestors from a normalised CDF [B,K] (search only); int32 [B,K]."""
    B, K = cdf.shape
    idx = torch.empty((B, K), dtype=torch.int32, device=cdf.device)
    _lib.call("aesmc_resample_from_cdf_f32", _lib.ptr(cdf), _lib.ptr(u), B, K, _lib.ptr(idx), _lib.ptr(flags))
    return idx


# ------------------------------------------------------------------------------------------------
# ancestral gather
# ------------------------------------------------------------------------------------------------
class _Gather(torch.autograd.Function):
    """state.resample for one tensor: out[b,k,...] = value[b, idx[b,k], ...] (state.py:158-183)."""

    @staticmethod
    def forward(ctx, value, idx, sorted_rows, flags):
        B, K = value.shape[:2]
        out = torch.empty_like(value)
        row_bytes = value[0, 0].numel() * value.element_size()
        _lib.call("aesmc_gather_bytes", _lib.ptr(value), _lib.ptr(idx), int(idx.dtype == torch.int64), B, K,
                  row_bytes, _lib.ptr(out), _lib.ptr(flags))
        ctx.save_for_backward(idx)
        ctx.sorted_rows = sorted_rows
        return out

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        B, K = g.shape[:2]
        D = g[0, 0].numel()
        dt = g.dtype
        if dt not in (torch.float32, torch.float64):
            g = g.float()
        g = g.contiguous()
        gsrc = torch.empty_like(g)
        fn = "aesmc_gather_bwd_f32" if g.dtype == torch.float32 else "aesmc_gather_bwd_f64"
        _lib.call(fn, _lib.ptr(g), _lib.ptr(idx), int(idx.dtype == torch.int64), B, K, D, _lib.ptr(gsrc),
                  int(ctx.sorted_rows))
        return gsrc.to(dt), None, None, None


def gather(value, idx, sorted_rows=False, flags=None):
    """value [B,K,...] CUDA contiguous, idx [B,K] int32/int64 CUDA contiguous."""
    if value.numel() == 0:
        return value.clone()
    return _Gather.apply(value, idx, sorted_rows, flags)


# ------------------------------------------------------------------------------------------------
# reductions / elementwise companions
# ------------------------------------------------------------------------------------------------
class _LogSumExp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, lw, flags):
        B, K = lw.shape
        out = torch.empty(B, dtype=lw.dtype, device=lw.device)
        fn = "aesmc_logsumexp_f32" if lw.dtype == torch.float32 else "aesmc_logsumexp_f64"
        _lib.call(fn, _lib.ptr(lw), B, K, _lib.ptr(out), _lib.ptr(flags))
        ctx.save_for_backward(lw, out)
        return out

    @staticmethod
    def backward(ctx, g):
        lw, lse = ctx.saved_tensors
        if lw.dtype != torch.float32:
            return g.unsqueeze(1) * torch.exp(lw - lse.unsqueeze(1)), None
        B, K = lw.shape
        out = torch.empty_like(lw)
        _lib.call("aesmc_step_bwd_f32", _lib.ptr(lw), _lib.ptr(lse), None, _lib.ptr(g.contiguous()), B, K,
                  _lib.ptr(out), None)
        return out, None


def logsumexp_rows(lw, flags=None):
    """[B,K] float32/float64 CUDA contiguous -> [B]; differentiable."""
    return _LogSumExp.apply(lw, flags)


def is_accumulate(a, b, c, acc, log_w, first):
    """acc (+)= (a + b) - c elementwise, log_w receives the per-step term (importance sampling)."""
    _lib.call("aesmc_is_accumulate_f32", _lib.ptr(a), _lib.ptr(b), _lib.ptr(c), _lib.ptr(acc), _lib.ptr(log_w),
              a.numel(), int(first))


def lognormexp_rows(lw, exponentiate):
    B, K = lw.shape
    out = torch.empty_like(lw)
    _lib.call("aesmc_lognormexp_f32", _lib.ptr(lw), B, K, _lib.ptr(out), int(exponentiate))
    return out


def log_ess_rows(lw):
    B, K = lw.shape
    out = torch.empty(B, dtype=lw.dtype, device=lw.device)
    fn = "aesmc_log_ess_f32" if lw.dtype == torch.float32 else "aesmc_log_ess_f64"
    _lib.call(fn, _lib.ptr(lw), B, K, _lib.ptr(out))
    return out


def weighted_moments(x, lw, want_second=True):
    """x [B,K,D] float32, lw [B,K] float32 -> (mean [B,D], second [B,D] | None)."""
    B, K = lw.shape
    D = x[0, 0].numel()
    mean = torch.empty((B, D), dtype=torch.float32, device=x.device)
    second = torch.empty((B, D), dtype=torch.float32, device=x.device) if want_second else None
    _lib.call("aesmc_weighted_moments_f32", _lib.ptr(x), _lib.ptr(lw), B, K, D, _lib.ptr(mean), _lib.ptr(second))
    return mean, second


# ------------------------------------------------------------------------------------------------
# torch.distributions.Normal.log_prob in one kernel (bit-identical to torch's elementwise sequence)
# ------------------------------------------------------------------------------------------------
_HALF_LOG_2PI = float(np.float32(np.log(np.sqrt(2 * np.pi))))


def _compact(t, B, K):
    """(compact differentiable view, kind) of a tensor that broadcasts against a [B, K] table:
    kind 0 = [B, K] contiguous, 1 = one value per row [B], 2 = a single element; None if neither."""
    if t.dim() == 0:
        return t, 2
    if t.dim() == 1 and t.shape[0] == B:
        return (t[0], 2) if (t.stride(0) == 0 and B > 1) else (t, 1)
    if t.dim() == 2 and tuple(t.shape) == (B, K):
        s0, s1 = t.stride()
        if s0 == 0 and s1 == 0:
            return t[0, 0], 2
        if s1 == 0:
            return t[:, 0], 1
        return t, 0
    if t.dim() == 2 and tuple(t.shape) == (B, 1):
        return t[:, 0], 1
    return None, None


class _NormalLogProb(torch.autograd.Function):
    @staticmethod
    def forward(ctx, value, value_kind, loc, loc_kind, scale, B, K):
        dev = value.device
        value = value.contiguous()
        loc_dev = loc is not None and loc.is_cuda
        scale_dev = scale.is_cuda
        loc_c = loc.contiguous() if loc_dev else None
        loc_host = 0.0 if loc_dev else float(loc)
        if scale_dev:
            scale_c, inv_two_var, log_scale = scale.contiguous(), 0.0, 0.0
        else:  # CPU scalar: torch multiplies by the float32 reciprocal of (2 * var) and subtracts scale.log()
            two_var = np.float32(2.0) * (np.float32(float(scale)) * np.float32(float(scale)))
            scale_c, inv_two_var, log_scale = None, float(np.float32(1.0) / two_var), float(scale.log())
        out = torch.empty((B, K), dtype=torch.float32, device=dev)
        _lib.call("aesmc_normal_log_prob_f32", _lib.ptr(value), value_kind, _lib.ptr(loc_c), loc_kind, loc_host,
                  _lib.ptr(scale_c), inv_two_var, log_scale, _HALF_LOG_2PI, B, K, _lib.ptr(out))
        ctx.save_for_backward(value, loc_c, scale_c)
        ctx.meta = (value_kind, loc_kind, loc_host, float(scale) if not scale_dev else 0.0, B, K)
        return out

    @staticmethod
    def backward(ctx, g):
        value, loc_c, scale_c = ctx.saved_tensors
        value_kind, loc_kind, loc_host, scale_host, B, K = ctx.meta
        need_v, need_l, need_s = ctx.needs_input_grad[0], ctx.needs_input_grad[2], ctx.needs_input_grad[4]
        g = g.contiguous()
        new = lambda: torch.empty((B, K), dtype=torch.float32, device=g.device)  # noqa: E731
        gv = new() if need_v else None
        gl = new() if need_l else None
        gs = new() if need_s else None
        _lib.call("aesmc_normal_log_prob_bwd_f32", _lib.ptr(value), value_kind, _lib.ptr(loc_c), loc_kind, loc_host,
                  _lib.ptr(scale_c), scale_host, _lib.ptr(g), B, K, _lib.ptr(gv), _lib.ptr(gl), _lib.ptr(gs))

        def fold(t, kind):
            if t is None:
                return None
            return t if kind == 0 else (t.sum(dim=1) if kind == 1 else t.sum())

        return fold(gv, value_kind), None, fold(gl, loc_kind), None, fold(gs, 2), None, None


def normal_log_prob(distribution, value):
    """log N(value; loc, scale) as a [B, K] table in one kernel, or None when the operands do not fit the
    fast path (then the caller uses distribution.log_prob).  Handles value [B, K] (dense or broadcast from
    [B]), loc dense / per-row / scalar (CUDA or CPU), scalar scale (CUDA or CPU)."""
    if type(distribution) is not torch.distributions.Normal or value.dim() != 2:
        return None
    if not (value.is_cuda and value.dtype == torch.float32):
        return None
    B, K = value.shape
    loc, scale = distribution.loc, distribution.scale
    if loc.dtype != torch.float32 or scale.dtype != torch.float32:
        return None
    v_c, v_kind = _compact(value, B, K)
    l_c, l_kind = _compact(loc, B, K)
    s_c, s_kind = _compact(scale, B, K)
    if v_c is None or l_c is None or s_kind != 2 or v_kind == 2:
        return None
    for t in (l_c, s_c):
        if t.is_cuda and t.device != value.device:
            return None
        if not t.is_cuda and (l_kind != 2 if t is l_c else False):
            return None
        if not t.is_cuda and t.requires_grad:  # CPU scalar parameters: leave the autograd graph to torch
            return None
    return _NormalLogProb.apply(v_c, v_kind, l_c, l_kind, s_c, B, K)


def independent_normal_log_prob(distribution, value):
    """log-density of Independent(Normal(loc [B, K, D], scalar scale), 1) at value [B, K, D], summed over D, or
    None when the operands do not fit.  The [B, K, D] tables are viewed as [B, K*D] and go through the
    one-kernel Normal.log_prob above (bit-identical to torch's elementwise sequence); the sum over D is the
    same torch reduction the generic path uses, so the result equals distribution.log_prob(value) bit for bit
    -- one elementwise pass instead of six."""
    if type(distribution) is not torch.distributions.Independent or distribution.reinterpreted_batch_ndims != 1:
        return None
    base = distribution.base_dist
    if type(base) is not torch.distributions.Normal or value.dim() != 3:
        return None
    if not (value.is_cuda and value.dtype == torch.float32):
        return None
    B, K, D = value.shape
    loc, scale = base.loc, base.scale
    if loc.dtype != torch.float32 or scale.dtype != torch.float32 or tuple(loc.shape) != (B, K, D):
        return None
    if tuple(scale.shape) != (B, K, D) or any(scale.stride()) or scale.device != value.device:
        return None  # only a scalar scale (expanded from one element)
    if loc.device != value.device or not loc.is_contiguous():
        return None
    flat = _NormalLogProb.apply(value.contiguous().view(B, K * D), 0, loc.view(B, K * D), 0, scale[0, 0, 0], B, K * D)
    return flat.view(B, K, D).sum(dim=2)


def compose_index(prev, cur):
    B, K = prev.shape
    out = torch.empty_like(prev)
    _lib.call("aesmc_compose_index_i32", _lib.ptr(prev), _lib.ptr(cur), B, K, _lib.ptr(out))
    return out


def iota_index(B, K, dev):
    out = torch.empty((B, K), dtype=torch.int32, device=dev)
    _lib.call("aesmc_iota_index_i32", B, K, _lib.ptr(out))
    return out


def widen_index(idx32):
    out = torch.empty(idx32.shape, dtype=torch.int64, device=idx32.device)
    _lib.call("aesmc_index_widen", _lib.ptr(idx32), _lib.ptr(out), idx32.numel())
    return out


def narrow_index(idx64):
    out = torch.empty(idx64.shape, dtype=torch.int32, device=idx64.device)
    _lib.call("aesmc_index_narrow", _lib.ptr(idx64), _lib.ptr(out), idx64.numel())
    return out


def uniforms_to_device(u, B, dev):
    """Per-row float64 uniforms as a contiguous CUDA tensor [B]."""
    if isinstance(u, np.ndarray):
        u = torch.from_numpy(np.ascontiguousarray(u, dtype=np.float64))
    u = u.reshape(B)
    if u.dtype != torch.float64:
        u = u.double()
    return u.to(dev, non_blocking=True).contiguous()


def uniforms_table_to_device(u, steps, B, dev):
    """All per-step, per-row resampling uniforms as ONE contiguous float64 CUDA tensor [steps, B]: one
    upload per infer() call instead of a pageable host-to-device copy per time step.  Accepts a numpy
    array / tensor of shape [steps, B] or [steps, B, 1] (host or device), or a sequence of per-step rows."""
    if steps == 0:
        return torch.empty((0, B), dtype=torch.float64, device=dev)
    if isinstance(u, (list, tuple)):
        if all(torch.is_tensor(r) for r in u):
            u = torch.stack([r.reshape(B) for r in u[:steps]])
        else:
            u = np.stack([np.asarray(r, dtype=np.float64).reshape(B) for r in u[:steps]])
    if isinstance(u, np.ndarray):
        u = torch.from_numpy(np.ascontiguousarray(u, dtype=np.float64))
    if u.shape[0] < steps:
        raise ValueError("uniforms: need %d rows of per-step uniforms, got %d" % (steps, u.shape[0]))
    u = u[:steps].reshape(steps, B)
    if u.dtype != torch.float64:
        u = u.double()
    return u.to(dev, non_blocking=True).contiguous()
