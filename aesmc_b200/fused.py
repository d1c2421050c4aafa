"""Fused-model fast path: scalar linear-Gaussian state-space models whose sampling and
log-densities are evaluated INSIDE the step kernel (aesmc_smc_step_lg_f32) instead of through ~25 torch
elementwise kernels per time step.

    x_0 ~ N(m0, s0^2)      x_t | x_{t-1} ~ N(a x_{t-1} + b, sx^2)      y_t | x_t ~ N(c x_t + d, sy^2)
    proposal: 'bootstrap' (the prior dynamics) or affine-Gaussian
              q(x_0 | y_0) = N(p0_y y_0 + p0_off, p0_scale^2),  q(x_t | x_{t-1}, y_t) = N(pt_x x_{t-1} + pt_y y_t + pt_off, pt_scale^2)

``ScalarLinearGaussianSSM`` is an ordinary user model: ``.initial / .transition / .emission / .proposal``
are callables with the reference's conventions returning torch.distributions, so the same object runs
through the generic path, through the oracle port and through the reference itself.  ``inference.infer``
recognises the four bound methods of one instance (or four reference-style modules tied together with
``link``) and dispatches to ``infer_fused``: same arguments, same result dict.  When gradients are required and only
the log-evidence is requested -- exactly what ``losses.get_loss`` asks for -- the step kernels run under one
autograd node whose backward is ONE launch per time step (aesmc_lg_step_bwd_f32: logsumexp gradient, the
ancestral gather's scatter-add and the analytic Normal gradients); anything else that needs gradients takes the
differentiable generic path.  The kernel reproduces torch.distributions.Normal's
float32 arithmetic operation for operation, so with injected noise the fused and the eager path agree
bit for bit (tests/test_fused_gpu.py); without, noise comes from an in-kernel Philox4x32-10 stream
seeded from torch's global generator.
"""
import math

import numpy as np
import torch

from . import _lib, _ops, state

Normal = torch.distributions.Normal
_FULL = state.BatchShapeMode.FULLY_EXPANDED
_BATCH = state.BatchShapeMode.BATCH_EXPANDED


class ScalarLinearGaussianSSM:
    def __init__(self, initial_loc=0.0, initial_scale=1.0, transition_mult=0.9, transition_offset=0.0,
                 transition_scale=1.0, emission_mult=1.0, emission_offset=0.0, emission_scale=0.5,
                 proposal="bootstrap", device=None):
        f = lambda v: torch.tensor(float(v), dtype=torch.float32, device=device)  # noqa: E731
        self.device = device
        self.m0, self.s0 = f(initial_loc), f(initial_scale)
        self.a, self.b, self.sx = f(transition_mult), f(transition_offset), f(transition_scale)
        self.c, self.d, self.sy = f(emission_mult), f(emission_offset), f(emission_scale)
        if proposal == "bootstrap":
            self.prop = None
        else:
            self.prop = {k: f(proposal[k]) for k in ("p0_y", "p0_off", "p0_scale", "pt_x", "pt_y", "pt_off", "pt_scale")}

    # ---- the reference's callable conventions (eager torch path) ---------------------------------
    def initial(self):
        return Normal(self.m0, self.s0)

    def transition(self, previous_latents=None, time=None, previous_observations=None):
        return state.set_batch_shape_mode(Normal(previous_latents[-1] * self.a + self.b, self.sx), _FULL)

    def emission(self, latents=None, time=None, previous_observations=None):
        return state.set_batch_shape_mode(Normal(latents[-1] * self.c + self.d, self.sy), _FULL)

    def proposal(self, previous_latents=None, time=None, observations=None):
        if self.prop is None:
            return self.initial() if time == 0 else self.transition(previous_latents=previous_latents, time=time)
        if time == 0:
            loc = observations[0] * self.prop["p0_y"] + self.prop["p0_off"]
            return state.set_batch_shape_mode(Normal(loc, self.prop["p0_scale"]), _BATCH)
        row = observations[time] * self.prop["pt_y"] + self.prop["pt_off"]
        loc = previous_latents[-1] * self.prop["pt_x"] + row.unsqueeze(1)
        return state.set_batch_shape_mode(Normal(loc, self.prop["pt_scale"]), _FULL)

    def callables(self):
        return self.initial, self.transition, self.emission, self.proposal

    def tensors(self):
        base = [self.m0, self.s0, self.a, self.b, self.sx, self.c, self.d, self.sy]
        return base + (list(self.prop.values()) if self.prop is not None else [])

    def requires_grad(self):
        return any(t.requires_grad for t in self.tensors())

    def scales_require_grad(self):
        scales = [self.s0, self.sx, self.sy] + ([self.prop["p0_scale"], self.prop["pt_scale"]] if self.prop is not None else [])
        return any(t.requires_grad for t in scales)

    def differentiable_parameters(self):
        """The ten multipliers / offsets the fused backward differentiates, in the order _FusedEvidence expects
        (proposal entries are placeholders for a bootstrap model)."""
        z = torch.zeros((), dtype=torch.float32, device=self.m0.device)
        pr = self.prop or {}
        return [self.m0, self.a, self.b, self.c, self.d, pr.get("p0_y", z), pr.get("p0_off", z), pr.get("pt_x", z),
                pr.get("pt_y", z), pr.get("pt_off", z)]

    def kernel_params_device(self):
        """float32 [2, 15] ON THE DEVICE (no host round trip; row 0 for t = 0, row 1 for t >= 1; each row t | e | q)."""
        with torch.no_grad():
            zero = torch.zeros_like(self.m0)
            init = self._affine(zero, self.m0, self.s0)
            trans = self._affine(self.a, self.b, self.sx)
            emis = self._affine(self.c, self.d, self.sy)
            if self.prop is None:
                q0, qt = init, trans
            else:
                q0 = self._affine(zero, self.prop["p0_off"], self.prop["p0_scale"])
                qt = self._affine(self.prop["pt_x"], self.prop["pt_off"], self.prop["pt_scale"])
            return torch.stack([torch.cat([init, emis, q0]), torch.cat([trans, emis, qt])]).float().contiguous()

    # ---- parameters for the fused kernel -------------------------------------------------------------
    def _affine(self, mult, off, scale):
        # (mult, off, scale, 2*var, log scale) computed with torch on the model's device, exactly as
        # Normal.log_prob computes them (float32)
        return torch.stack([mult, off, scale, 2 * (scale ** 2), scale.log()])

    def kernel_params(self):
        """float32 [2, 15] on the host: row 0 for t = 0, row 1 for t >= 1; each row = t | e | q."""
        zero = torch.zeros_like(self.m0)
        init = self._affine(zero, self.m0, self.s0)
        trans = self._affine(self.a, self.b, self.sx)
        emis = self._affine(self.c, self.d, self.sy)
        if self.prop is None:
            q0, qt = init, trans
        else:
            q0 = self._affine(zero, self.prop["p0_off"], self.prop["p0_scale"])
            qt = self._affine(self.prop["pt_x"], self.prop["pt_off"], self.prop["pt_scale"])
        return torch.stack([torch.cat([init, emis, q0]), torch.cat([trans, emis, qt])]).cpu().numpy().astype(np.float32)

    def proposal_row_offset(self, observation, time):
        """[B] per-row offset of the proposal mean (None when it does not depend on the observation)."""
        if self.prop is None:
            return None
        if time == 0:
            return (observation * self.prop["p0_y"] + self.prop["p0_off"]).contiguous()
        return (observation * self.prop["pt_y"] + self.prop["pt_off"]).contiguous()


class LinkedLGSSM(ScalarLinearGaussianSSM):
    """A ScalarLinearGaussianSSM VIEW of four reference-style callables (the reference's test/models/lgssm.py:10-72
    and this repo's tests/models/lgssm.py): Initial(loc, scale); Transition / Emission modules with a scalar
    Parameter ``mult`` and a constant ``scale``; a Proposal module with ``lin_0 = Linear(1, 1)``,
    ``lin_t = Linear(2, 1)`` on (x_prev, y_t) and constant ``scale_0`` / ``scale_t``.  Parameters are read from the
    modules at every call, so optimiser updates are seen and gradients land on the modules' own Parameters."""

    def __init__(self, initial, transition, emission, proposal, proposal_scale_t=None):
        self._mods = (initial, transition, emission, proposal)
        self.device = transition.mult.device
        self._scale_t_attr = proposal_scale_t

    def _f(self, v):
        if torch.is_tensor(v):
            return v.to(device=self.device, dtype=torch.float32).reshape(())
        return torch.tensor(float(v), dtype=torch.float32, device=self.device)

    m0 = property(lambda s: s._f(s._mods[0].loc))
    s0 = property(lambda s: s._f(s._mods[0].scale))
    a = property(lambda s: s._mods[1].mult.reshape(()))
    b = property(lambda s: s._f(0.0))
    sx = property(lambda s: s._f(s._mods[1].scale))
    c = property(lambda s: s._mods[2].mult.reshape(()))
    d = property(lambda s: s._f(0.0))
    sy = property(lambda s: s._f(s._mods[2].scale))

    @property
    def prop(self):
        q = self._mods[3]
        scale_t = q.scale_t if self._scale_t_attr is None else self._scale_t_attr
        return {"p0_y": q.lin_0.weight[0, 0], "p0_off": q.lin_0.bias[0], "p0_scale": self._f(q.scale_0),
                "pt_x": q.lin_t.weight[0, 0], "pt_y": q.lin_t.weight[0, 1], "pt_off": q.lin_t.bias[0],
                "pt_scale": self._f(scale_t)}


def link(initial, transition, emission, proposal, proposal_scale_t=None):
    """Opt reference-style LGSSM modules into the fused kernels: checks their structure and ties them together, so
    that ``infer`` / ``get_loss`` called with exactly these four objects run (and train) through
    aesmc_smc_step_lg_* instead of torch elementwise kernels.  Returns the LinkedLGSSM view.
    proposal_scale_t: the scale the proposal's forward() really uses for t >= 1 (the reference's own
    test/models/lgssm.py:63 passes scale_0 there); default proposal.scale_t."""
    import torch.nn as nn
    ok = (hasattr(initial, "loc") and hasattr(initial, "scale")
          and isinstance(getattr(transition, "mult", None), torch.Tensor) and hasattr(transition, "scale")
          and isinstance(getattr(emission, "mult", None), torch.Tensor) and hasattr(emission, "scale")
          and isinstance(getattr(proposal, "lin_0", None), nn.Linear) and isinstance(getattr(proposal, "lin_t", None), nn.Linear)
          and hasattr(proposal, "scale_0") and hasattr(proposal, "scale_t"))
    if not ok or tuple(proposal.lin_0.weight.shape) != (1, 1) or tuple(proposal.lin_t.weight.shape) != (1, 2) \
            or transition.mult.numel() != 1 or emission.mult.numel() != 1:
        raise ValueError("link() needs Initial(loc, scale), Transition/Emission modules with a scalar `mult` Parameter and a "
                         "`scale`, and a Proposal with lin_0 = Linear(1, 1), lin_t = Linear(2, 1), scale_0, scale_t")
    view = LinkedLGSSM(initial, transition, emission, proposal, proposal_scale_t)
    proposal._aesmc_b200_fused = view
    return view


def model_of(initial, transition, emission, proposal):
    """The ScalarLinearGaussianSSM whose bound methods these four callables are (or the LinkedLGSSM view that
    ``link`` tied to exactly these four objects), else None."""
    view = getattr(proposal, "_aesmc_b200_fused", None)
    if isinstance(view, (LinkedLGSSM, LinkedDenseLGSSM)):
        return view if all(a is b for a, b in zip(view._mods, (initial, transition, emission, proposal))) else None
    owner = getattr(proposal, "__self__", None)
    if not isinstance(owner, (ScalarLinearGaussianSSM, VectorLinearGaussianSSM)):
        return None
    for fn, name in ((initial, "initial"), (transition, "transition"), (emission, "emission"), (proposal, "proposal")):
        if getattr(fn, "__self__", None) is not owner or getattr(fn, "__name__", None) != name:
            return None
    return owner


def applicable(model, observations, num_particles, evidence_only=False):
    """Can this call run on the fused kernels?  evidence_only: the caller wants nothing but the log-evidence
    (get_loss) -- the one request the fused path can differentiate."""
    if model is None or not torch.cuda.is_available():
        return False
    if isinstance(model, VectorLinearGaussianSSM):
        return vector_applicable(model, observations, num_particles)
    first = observations[0]
    if isinstance(first, dict) or not torch.is_tensor(first) or first.dim() != 1 or not first.is_cuda:
        return False
    if first.dtype != torch.float32 or model.m0.device != first.device:
        return False
    if not (64 <= num_particles <= 16384 and num_particles % 4 == 0):
        return False
    if torch.is_grad_enabled() and model.requires_grad():
        # trainable model: fused only for the evidence, and only with constant scales (the backward kernel
        # differentiates multipliers and offsets); everything else takes the differentiable generic path
        return evidence_only and not model.scales_require_grad()
    return True


_HALF_LOG_2PI = float(np.float32(math.log(math.sqrt(2 * math.pi))))


def infer_fused(model, observations, num_particles, return_log_marginal_likelihood=False, return_latents=True,
                return_original_latents=False, return_log_weight=True, return_log_weights=False,
                return_ancestral_indices=False, uniforms=None, resampling_mode=None, check_finite=True, noise=None):
    """SMC ('smc' only) with the model fused into the step kernel; arguments and result as inference.infer.
    noise: optional [T, B, K] float32 standard normals to use instead of the in-kernel Philox stream."""
    from . import inference  # late import: inference imports this module
    T = len(observations)
    B = observations[0].size(0)
    K = num_particles
    dev = observations[0].device
    params = model.kernel_params()
    mode = _ops.mode_code(resampling_mode)
    flags = _ops.new_flags(dev)
    seed = int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())  # torch.manual_seed controls the stream
    keep_originals = return_original_latents or return_latents
    keep_index = return_ancestral_indices or return_latents
    originals, log_weights, ancestors = [], [], []
    lses = torch.empty(T, B, dtype=torch.float32, device=dev)
    arena = [torch.empty(B, K, dtype=torch.float32, device=dev) for _ in range(2)]
    x_prev = None
    x_last = None
    log_w = None
    u_all = None
    if T > 1:  # inference.py:250: one np.random.uniform(size=[B, 1]) per step, here drawn and uploaded once
        u_all = _ops.uniforms_table_to_device(np.random.uniform(size=[T - 1, B]) if uniforms is None else uniforms,
                                              T - 1, B, dev)
    for t in range(T):
        last = t == T - 1
        y = observations[t].contiguous()
        q_off = model.proposal_row_offset(y, t)
        x_new = torch.empty(B, K, dtype=torch.float32, device=dev) if (keep_originals or last) else None
        want_lw = return_log_weights or (last and return_log_weight)
        log_w = torch.empty(B, K, dtype=torch.float32, device=dev) if want_lw else None
        if last:
            u_dev = idx = x_out = None
        else:
            u_dev = u_all[t]
            idx = torch.empty(B, K, dtype=torch.int32, device=dev) if keep_index else None  # ancestors not stored
            x_out = arena[t & 1]
        nz = None if noise is None else noise[t].contiguous()
        _lib.call("aesmc_smc_step_lg_f32", _lib.ptr(x_prev), _lib.ptr(y), _lib.ptr(nz), _lib.ptr(q_off),
                  params[0 if t == 0 else 1].ctypes.data, _HALF_LOG_2PI, seed, None, t, B, K, _lib.ptr(u_dev),
                  _lib.ptr(x_new), _lib.ptr(log_w), _lib.ptr(lses[t]), _lib.ptr(idx), _lib.ptr(x_out),
                  _lib.ptr(flags), mode)
        if keep_originals:
            originals.append(x_new)
        if return_log_weights:
            log_weights.append(log_w)
        if not last and keep_index:
            ancestors.append(idx)
        x_prev = x_out
        x_last = x_new
    result = dict.fromkeys(("log_marginal_likelihood", "latents", "original_latents", "log_weight", "log_weights",
                            "ancestral_indices"))
    if return_log_marginal_likelihood:
        result["log_marginal_likelihood"] = (lses - math.log(K)).sum(dim=0)  # inference.py:130-132
    if return_latents:
        result["latents"] = inference._trace_genealogy(originals, ancestors, dev)
    if return_original_latents:
        result["original_latents"] = originals
    if return_log_weight:
        result["log_weight"] = log_w
    if return_log_weights:
        result["log_weights"] = log_weights
    if return_ancestral_indices:
        result["ancestral_indices"] = [_ops.widen_index(i) for i in ancestors]
    result["last_latent"] = x_last
    if check_finite:
        _ops.raise_on_flags(flags)
    return result


class _FusedEvidence(torch.autograd.Function):
    """log-evidence [B] of the SMC filter on a scalar linear-Gaussian model, differentiable w.r.t. the model's ten
    multipliers / offsets: T fused-step launches forward, T aesmc_lg_step_bwd_f32 launches backward.  Replaces the
    autograd graph that losses.get_loss builds over inference.py:85-134 (~60 torch kernels per time step)."""

    @staticmethod
    def forward(ctx, obs, q_off, params_dev, u_all, noise, num_particles, mode, seed, flags, bootstrap, *theta):
        T, B = obs.shape
        K, dev = num_particles, obs.device
        X = torch.empty(T, B, K, dtype=torch.float32, device=dev)                 # proposed latents x_t
        XP = torch.empty(max(T - 1, 1), B, K, dtype=torch.float32, device=dev)    # XP[t] = x_t[idx_t]: inputs of step t + 1
        IDX = torch.empty(max(T - 1, 1), B, K, dtype=torch.int32, device=dev)
        lses = torch.empty(T, B, dtype=torch.float32, device=dev)
        for t in range(T):
            last = t == T - 1
            _lib.call("aesmc_smc_step_lg_dev_f32", _lib.ptr(XP[t - 1]) if t else None, _lib.ptr(obs[t]),
                      _lib.ptr(noise[t]) if noise is not None else None, _lib.ptr(q_off[t]) if q_off is not None else None,
                      _lib.ptr(params_dev[0 if t == 0 else 1]), _HALF_LOG_2PI, seed, None, t, B, K,
                      None if last else _lib.ptr(u_all[t]), _lib.ptr(X[t]), None, _lib.ptr(lses[t]),
                      None if last else _lib.ptr(IDX[t]), None if last else _lib.ptr(XP[t]), _lib.ptr(flags), mode)
        ctx.save_for_backward(obs, q_off, params_dev, X, XP, IDX, lses)
        ctx.bootstrap = bootstrap
        return (lses - math.log(K)).sum(dim=0)  # inference.py:130-132

    @staticmethod
    def backward(ctx, g_lml):
        obs, q_off, params_dev, X, XP, IDX, lses = ctx.saved_tensors
        T, B, K = X.shape
        dev = X.device
        g_lse = g_lml.contiguous().float()  # d lml / d lse_t = 1 for every t
        gpar = torch.empty(T, B, 6, dtype=torch.float32, device=dev)
        G = None
        for t in range(T - 1, -1, -1):
            g_xp = torch.empty(B, K, dtype=torch.float32, device=dev) if t else None
            _lib.call("aesmc_lg_step_bwd_f32", _lib.ptr(X[t]), _lib.ptr(XP[t - 1]) if t else None, _lib.ptr(obs[t]),
                      _lib.ptr(q_off[t]) if q_off is not None else None, _lib.ptr(params_dev[0 if t == 0 else 1]),
                      _lib.ptr(lses[t]), _lib.ptr(g_lse), _lib.ptr(G) if G is not None else None,
                      _lib.ptr(IDX[t]) if t < T - 1 else None, B, K, _lib.ptr(g_xp), _lib.ptr(gpar[t]))
            G = g_xp
        P = gpar.sum(dim=1)                       # [T, 6]: t.mult, t.off, e.mult, e.off, q.mult, q_off
        Pt = P[1:].sum(dim=0) if T > 1 else torch.zeros(6, device=dev)
        gq = gpar[:, :, 5]                        # [T, B] dL/dq_off[t, b]
        g_m0, g_a, g_b = P[0, 1], Pt[0], Pt[1]
        g_c, g_d = P[:, 2].sum(), P[:, 3].sum()
        zero = torch.zeros((), device=dev)
        if ctx.bootstrap:  # the proposal IS the prior: its reparameterised path lands on the prior's parameters
            g_m0 = g_m0 + gq[0].sum()
            g_a, g_b = g_a + Pt[4], g_b + Pt[5]
            g_p0y = g_p0o = g_ptx = g_pty = g_pto = zero
        else:
            g_p0y, g_p0o = (gq[0] * obs[0]).sum(), gq[0].sum()
            g_ptx = Pt[4]
            g_pty, g_pto = ((gq[1:] * obs[1:]).sum(), gq[1:].sum()) if T > 1 else (zero, zero)
        return (None,) * 10 + (g_m0, g_a, g_b, g_c, g_d, g_p0y, g_p0o, g_ptx, g_pty, g_pto)


def evidence_with_grad(model, observations, num_particles, uniforms=None, resampling_mode=None, check_finite=True,
                       noise=None):
    """Differentiable SMC log-evidence [B] of a (trainable) scalar linear-Gaussian model on the fused kernels."""
    T = len(observations)
    obs = observations if torch.is_tensor(observations) else torch.stack(list(observations))
    obs = obs.contiguous()
    B, K, dev = obs.shape[1], num_particles, obs.device
    flags = _ops.new_flags(dev)
    u_all = None
    if T > 1:
        u_all = _ops.uniforms_table_to_device(np.random.uniform(size=[T - 1, B]) if uniforms is None else uniforms, T - 1, B, dev)
    seed = int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())
    q_off = None
    if model.prop is not None:
        with torch.no_grad():
            pr = model.prop
            q_off = torch.cat([(obs[:1] * pr["p0_y"] + pr["p0_off"]), (obs[1:] * pr["pt_y"] + pr["pt_off"])]).contiguous()
    nz = None if noise is None else (noise if torch.is_tensor(noise) else torch.stack(list(noise))).contiguous()
    lml = _FusedEvidence.apply(obs, q_off, model.kernel_params_device(), u_all, nz, K, _ops.mode_code(resampling_mode),
                               seed, flags, model.prop is None, *model.differentiable_parameters())
    if check_finite:
        _ops.raise_on_flags(flags)
    return lml


class GraphedFilter:
    """The T-step bootstrap / guided particle filter of a ScalarLinearGaussianSSM captured ONCE as a CUDA
    graph (T fused-step launches, device-side uniforms, the evidence reduction) and replayed per batch of
    observation sequences: no per-step Python, launch or allocation cost -- what makes small problems
    (BASELINE config 1: B = 1, K = 100, T = 50) launch-latency-free.

        f = GraphedFilter(model, num_timesteps=T, batch_size=B, num_particles=K)
        log_evidence = f(observations)          # [T, B] float32 tensor (host or device) -> [B]

    Every replay advances the Philox key (a device-resident counter) and draws fresh resampling uniforms
    with torch's graph-safe generator.  `noise` / `uniforms` buffers can be filled for tests
    (inject_noise=True).  Degenerate-weight flags accumulate in `self.flags` (check with `check()`)."""

    def __init__(self, model, num_timesteps, batch_size, num_particles, resampling_mode=None, inject_noise=False):
        first = torch.empty(batch_size, dtype=torch.float32, device=model.m0.device)
        if not applicable(model, [first], num_particles):
            raise ValueError("GraphedFilter needs a CUDA ScalarLinearGaussianSSM and 64 <= K <= 16384, K % 4 == 0")
        T, B, K = num_timesteps, batch_size, num_particles
        dev = model.m0.device
        self.model, self.T, self.B, self.K = model, T, B, K
        self.obs = torch.zeros(T, B, dtype=torch.float32, device=dev)
        self.uniforms = torch.zeros(max(T - 1, 1), B, dtype=torch.float64, device=dev)
        self.noise = torch.zeros(T, B, K, dtype=torch.float32, device=dev) if inject_noise else None
        self.lses = torch.zeros(T, B, dtype=torch.float32, device=dev)
        self.log_weight = torch.empty(B, K, dtype=torch.float32, device=dev)
        self.last_latent = torch.empty(B, K, dtype=torch.float32, device=dev)
        self.flags = _ops.new_flags(dev)
        self.seed = torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).to(dev)
        self._arena = [torch.empty(B, K, dtype=torch.float32, device=dev) for _ in range(2)]
        self._q_off = None if model.prop is None else torch.empty(T, B, dtype=torch.float32, device=dev)
        self._params = model.kernel_params()
        self._mode = _ops.mode_code(resampling_mode)
        self._inject = inject_noise
        self.log_evidence = torch.empty(B, dtype=torch.float32, device=dev)
        # warm-up on a side stream (loads the module, sets function attributes), then capture
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            self._body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.flags.zero_()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._body()

    def _body(self):
        model, T, B, K = self.model, self.T, self.B, self.K
        if not self._inject:
            self.seed.add_(1)
            self.uniforms.copy_(torch.rand(self.uniforms.shape, dtype=torch.float64, device=self.uniforms.device))
        if self._q_off is not None:
            self._q_off[0].copy_(self.obs[0] * model.prop["p0_y"] + model.prop["p0_off"])
            if T > 1:
                self._q_off[1:].copy_(self.obs[1:] * model.prop["pt_y"] + model.prop["pt_off"])
        x_prev = None
        for t in range(T):
            last = t == T - 1
            x_out = None if last else self._arena[t & 1]
            _lib.call("aesmc_smc_step_lg_f32", _lib.ptr(x_prev), _lib.ptr(self.obs[t]),
                      _lib.ptr(None if self.noise is None else self.noise[t]),
                      _lib.ptr(None if self._q_off is None else self._q_off[t]),
                      self._params[0 if t == 0 else 1].ctypes.data, _HALF_LOG_2PI, 0,
                      None if self._inject else _lib.ptr(self.seed), t, B, K,
                      None if last else _lib.ptr(self.uniforms[t]), _lib.ptr(self.last_latent) if last else None,
                      _lib.ptr(self.log_weight) if last else None, _lib.ptr(self.lses[t]),
                      None, _lib.ptr(x_out), _lib.ptr(self.flags), self._mode)  # ancestors are not stored
            x_prev = x_out
        self.log_evidence.copy_((self.lses - math.log(K)).sum(dim=0))

    def __call__(self, observations, clone=True):
        """observations: [T, B] float32 tensor (any device) or list of T tensors [B]."""
        if not torch.is_tensor(observations):
            observations = torch.stack(list(observations))
        self.obs.copy_(observations, non_blocking=True)
        self.graph.replay()
        return self.log_evidence.clone() if clone else self.log_evidence

    def check(self):
        """Host read of the accumulated NaN / degenerate-row flags (raises like inference.infer)."""
        _ops.raise_on_flags(self.flags)


# ------------------------------------------------------------------------------------------------------
# vector latents: D-dimensional linear-Gaussian models with diagonal noise (BASELINE config 3)
# ------------------------------------------------------------------------------------------------------
Independent = torch.distributions.Independent


class VectorLinearGaussianSSM:
    """x_0 ~ N(m0, diag s0^2),  x_t = A x_{t-1} + b + N(0, diag sx^2),  y_t = C x_t + d + N(0, diag sy^2), with a
    bootstrap proposal or q(x_0 | y_0) = N(W0 y_0 + b0, diag sq0^2), q(x_t | x_{t-1}, y_t) = N(Wx x_{t-1} + Wy y_t + bt,
    diag sqt^2).  Latents are [B, K, D] tensors (event_shape [D]), observations [B, Dy].

    Like the scalar family it is an ordinary user model (``.initial / .transition / .emission / .proposal`` follow the
    reference's callable conventions and run through the generic path, the oracle port and the reference itself);
    ``inference.infer`` recognises its bound methods and, when no gradient is needed, evaluates the model with ONE
    launch per time step (aesmc_lgv_propose_f32) followed by the step kernel, instead of ~40 torch kernels.
    Scales may be floats or [D] / [Dy] tensors.  1 <= D, Dy <= 16."""

    def __init__(self, initial_loc, initial_scale, A, transition_scale, C, emission_scale, b=None, d=None,
                 proposal="bootstrap", device=None):
        dev = device if device is not None else A.device
        f = lambda v, n: (v.to(dev, torch.float32) if torch.is_tensor(v) else torch.full((n,), float(v), device=dev)).reshape(n)  # noqa: E731
        self.device = dev
        self.A = A.to(dev, torch.float32)
        self.C = C.to(dev, torch.float32)
        self.D, self.Dy = self.A.shape[0], self.C.shape[0]
        D, Dy = self.D, self.Dy
        self.m0, self.s0 = f(initial_loc, D), f(initial_scale, D)
        self.b = f(0.0 if b is None else b, D)
        self.d = f(0.0 if d is None else d, Dy)
        self.sx, self.sy = f(transition_scale, D), f(emission_scale, Dy)
        if proposal == "bootstrap":
            self.prop = None
        else:
            self.prop = {"W0": proposal["W0"].to(dev, torch.float32), "b0": f(proposal["b0"], D), "s0": f(proposal["s0"], D),
                         "Wx": proposal["Wx"].to(dev, torch.float32), "Wy": proposal["Wy"].to(dev, torch.float32),
                         "bt": f(proposal["bt"], D), "st": f(proposal["st"], D)}

    # ---- the reference's callable conventions (eager torch path) ---------------------------------
    def initial(self):
        return Independent(Normal(self.m0, self.s0), 1)

    def transition(self, previous_latents=None, time=None, previous_observations=None):
        return state.set_batch_shape_mode(Independent(Normal(previous_latents[-1] @ self.A.T + self.b, self.sx), 1), _FULL)

    def emission(self, latents=None, time=None, previous_observations=None):
        return state.set_batch_shape_mode(Independent(Normal(latents[-1] @ self.C.T + self.d, self.sy), 1), _FULL)

    def proposal(self, previous_latents=None, time=None, observations=None):
        if self.prop is None:
            return self.initial() if time == 0 else self.transition(previous_latents=previous_latents, time=time)
        pr = self.prop
        if time == 0:
            return state.set_batch_shape_mode(Independent(Normal(observations[0] @ pr["W0"].T + pr["b0"], pr["s0"]), 1), _BATCH)
        loc = previous_latents[-1] @ pr["Wx"].T + (observations[time] @ pr["Wy"].T + pr["bt"]).unsqueeze(1)
        return state.set_batch_shape_mode(Independent(Normal(loc, pr["st"]), 1), _FULL)

    def callables(self):
        return self.initial, self.transition, self.emission, self.proposal

    def tensors(self):
        base = [self.m0, self.s0, self.A, self.b, self.sx, self.C, self.d, self.sy]
        return base + (list(self.prop.values()) if self.prop is not None else [])

    def requires_grad(self):
        return any(t.requires_grad for t in self.tensors())

    # ---- parameters for aesmc_lgv_propose_f32 ---------------------------------------------------------
    def kernel_params(self):
        """(params_t0, params_t) float32 numpy blocks on the host: A | b | sx | C | d | sy | Wx | sq."""
        D = self.D
        with torch.no_grad():
            zero = torch.zeros(D, D, device=self.device)
            tail0 = [zero, self.s0] if self.prop is None else [zero, self.prop["s0"]]
            tailt = [zero, self.sx] if self.prop is None else [self.prop["Wx"], self.prop["st"]]
            emis = [self.C, self.d, self.sy]
            first = torch.cat([t.reshape(-1) for t in [zero, self.m0, self.s0] + emis + tail0])
            later = torch.cat([t.reshape(-1) for t in [self.A, self.b, self.sx] + emis + tailt])
            both = torch.stack([first, later]).float().cpu().numpy()
        return np.ascontiguousarray(both[0]), np.ascontiguousarray(both[1])

    def proposal_row_means(self, obs):
        """[T, B, D]: the part of the proposal mean that depends on the observation only (None for a bootstrap model)."""
        if self.prop is None:
            return None
        with torch.no_grad():
            pr = self.prop
            first = obs[:1] @ pr["W0"].T + pr["b0"]
            return torch.cat([first, obs[1:] @ pr["Wy"].T + pr["bt"]]).contiguous()


class LinkedDenseLGSSM(VectorLinearGaussianSSM):
    """A VectorLinearGaussianSSM VIEW of four modules shaped like tests/models/lgssm_dense.py (BASELINE config 3's user
    model): Initial(loc [D], scale), Transition(A, scale), Emission(C, scale), Proposal(lin_0 = Linear(Dy, D),
    lin_t = Linear(D + Dy, D) on [x_prev, y_t], log_scale_0, log_scale_t).  Parameters are read at every call."""

    def __init__(self, initial, transition, emission, proposal, bootstrap=False):
        self._mods = (initial, transition, emission, proposal)
        self._boot = bootstrap
        self.device = transition.A.device
        self.D, self.Dy = transition.A.shape[0], emission.C.shape[0]

    def _v(self, v, n):
        if torch.is_tensor(v):
            return v.detach().to(self.device, torch.float32).reshape(-1).expand(n) if v.numel() == 1 else v.detach().to(self.device, torch.float32).reshape(n)
        return torch.full((n,), float(v), device=self.device)

    m0 = property(lambda s: s._v(s._mods[0].loc, s.D))
    s0 = property(lambda s: s._v(s._mods[0].scale, s.D))
    A = property(lambda s: s._mods[1].A.detach())
    b = property(lambda s: torch.zeros(s.D, device=s.device))
    sx = property(lambda s: s._v(s._mods[1].scale, s.D))
    C = property(lambda s: s._mods[2].C.detach())
    d = property(lambda s: torch.zeros(s.Dy, device=s.device))
    sy = property(lambda s: s._v(s._mods[2].scale, s.Dy))

    @property
    def prop(self):
        q, D = self._mods[3], self.D
        if self._boot:
            return None
        Wt = q.lin_t.weight.detach()
        return {"W0": q.lin_0.weight.detach(), "b0": q.lin_0.bias.detach(), "s0": q.log_scale_0.detach().exp(),
                "Wx": Wt[:, :D].contiguous(), "Wy": Wt[:, D:].contiguous(), "bt": q.lin_t.bias.detach(),
                "st": q.log_scale_t.detach().exp()}

    def requires_grad(self):
        mods = self._mods
        ts = [mods[1].A, mods[2].C] + ([] if self._boot else list(mods[3].parameters()))
        return any(torch.is_tensor(t) and t.requires_grad for t in ts)


def link_dense(initial, transition, emission, proposal):
    """Opt modules shaped like tests/models/lgssm_dense.py into the fused vector path (see LinkedDenseLGSSM); returns the
    view.  A proposal object with attributes ``initial`` and ``transition`` equal to the first two arguments (the prior
    as proposal) is taken as a bootstrap proposal."""
    import torch.nn as nn
    boot = getattr(proposal, "initial", None) is initial and getattr(proposal, "transition", None) is transition
    ok = (torch.is_tensor(getattr(initial, "loc", None)) and hasattr(initial, "scale")
          and torch.is_tensor(getattr(transition, "A", None)) and transition.A.dim() == 2 and hasattr(transition, "scale")
          and torch.is_tensor(getattr(emission, "C", None)) and emission.C.dim() == 2 and hasattr(emission, "scale"))
    if ok and not boot:
        D, Dy = transition.A.shape[0], emission.C.shape[0]
        ok = (isinstance(getattr(proposal, "lin_0", None), nn.Linear) and isinstance(getattr(proposal, "lin_t", None), nn.Linear)
              and tuple(proposal.lin_0.weight.shape) == (D, Dy) and tuple(proposal.lin_t.weight.shape) == (D, D + Dy)
              and hasattr(proposal, "log_scale_0") and hasattr(proposal, "log_scale_t"))
    if not ok:
        raise ValueError("link_dense() needs Initial(loc [D], scale), Transition(A [D,D], scale), Emission(C [Dy,D], scale) and "
                         "a Proposal with lin_0 = Linear(Dy, D), lin_t = Linear(D + Dy, D), log_scale_0, log_scale_t (or the "
                         "prior as proposal)")
    view = LinkedDenseLGSSM(initial, transition, emission, proposal, bootstrap=boot)
    try:
        proposal._aesmc_b200_fused = view
    except AttributeError:
        raise ValueError("link_dense(): cannot tag the proposal object")
    return view


def vector_applicable(model, observations, num_particles):
    """Can this call run through aesmc_lgv_propose_f32 + the step kernel?  (Forward only: anything that needs
    gradients takes the differentiable generic path.)"""
    if model is None or not torch.cuda.is_available():
        return False
    first = observations[0]
    if isinstance(first, dict) or not torch.is_tensor(first) or first.dim() != 2 or not first.is_cuda:
        return False
    if first.dtype != torch.float32 or first.device != model.A.device or first.size(1) != model.Dy:
        return False
    if not (1 <= model.D <= 16 and 1 <= model.Dy <= 16 and num_particles >= 1):
        return False
    return not (torch.is_grad_enabled() and model.requires_grad())


def infer_fused_vector(model, observations, num_particles, return_log_marginal_likelihood=False, return_latents=True,
                       return_original_latents=False, return_log_weight=True, return_log_weights=False,
                       return_ancestral_indices=False, uniforms=None, resampling_mode=None, check_finite=True, noise=None):
    """SMC with a VectorLinearGaussianSSM: per time step one model launch (sampling + three log-densities) and one
    step launch (lse, ancestors, D-float gather); arguments and result as inference.infer.
    noise: optional [T, B, K, D] float32 standard normals instead of the in-kernel Philox stream."""
    from . import inference
    T, K, D, Dy = len(observations), num_particles, model.D, model.Dy
    obs = observations if torch.is_tensor(observations) else torch.stack(list(observations))
    obs = obs.contiguous()
    B, dev = obs.shape[1], obs.device
    p0, pt = model.kernel_params()
    q_rows = model.proposal_row_means(obs)
    boot = 1 if q_rows is None else 0
    flags = _ops.new_flags(dev)
    seed = int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())
    keep_originals = return_original_latents or return_latents
    keep_index = return_ancestral_indices or return_latents
    originals, log_weights, ancestors = [], [], []
    lses = torch.empty(T, B, dtype=torch.float32, device=dev)
    u_all = None
    if T > 1:
        u_all = _ops.uniforms_table_to_device(np.random.uniform(size=[T - 1, B]) if uniforms is None else uniforms, T - 1, B, dev)
    scratch = [torch.empty(B, K, D, dtype=torch.float32, device=dev) for _ in range(2)]
    x_prev = x_new = log_w = None
    with torch.no_grad():
        for t in range(T):
            last = t == T - 1
            x_new = torch.empty(B, K, D, dtype=torch.float32, device=dev) if (keep_originals or last) else scratch[t & 1]
            lw_raw = torch.empty(B, K, dtype=torch.float32, device=dev)
            nz = None if noise is None else noise[t].contiguous()
            _lib.call("aesmc_lgv_propose_f32", _lib.ptr(x_prev), _lib.ptr(obs[t]), _lib.ptr(nz),
                      _lib.ptr(None if q_rows is None else q_rows[t]), (p0 if t == 0 else pt).ctypes.data, D, Dy, boot, seed, t,
                      B, K, _lib.ptr(x_new), _lib.ptr(lw_raw))
            log_w, lse, idx, x_res = _ops.smc_step(lw_raw, None, None, None if last else u_all[t], None if last else x_new,
                                                   flags, resampling_mode, resample=not last)
            lses[t] = lse
            if keep_originals:
                originals.append(x_new)
            if return_log_weights:
                log_weights.append(log_w)
            if not last and keep_index:
                ancestors.append(idx)
            x_prev = x_res
    result = dict.fromkeys(("log_marginal_likelihood", "latents", "original_latents", "log_weight", "log_weights",
                            "ancestral_indices"))
    if return_log_marginal_likelihood:
        result["log_marginal_likelihood"] = (lses - math.log(K)).sum(dim=0)  # inference.py:130-132
    if return_latents:
        result["latents"] = inference._trace_genealogy(originals, ancestors, dev)
    if return_original_latents:
        result["original_latents"] = originals
    if return_log_weight:
        result["log_weight"] = log_w
    if return_log_weights:
        result["log_weights"] = log_weights
    if return_ancestral_indices:
        result["ancestral_indices"] = [_ops.widen_index(i) for i in ancestors]
    result["last_latent"] = x_new
    if check_finite:
        _ops.raise_on_flags(flags)
    return result
