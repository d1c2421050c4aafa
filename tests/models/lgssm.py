"""User models written against the aesmc callable conventions: scalar linear-Gaussian SSM

    x_0 ~ N(m0, s0^2),  x_t = a x_{t-1} + N(0, s_x^2),  y_t = c x_t + N(0, s_y^2)

`Initial/Transition/Emission/Proposal` play the role of the reference's test/models/lgssm.py (learnable
multipliers, two-Linear proposal); `Bootstrap*` are parameter-free callables for the bootstrap
particle filter of BASELINE config 2 (proposal == transition), as in test/test_inference.py:87-143.
All of them work with either `aesmc_b200` or the reference package, since distributions are tagged
through whichever `state` module is passed in (default: aesmc_b200.state).
"""
import math

import torch
import torch.nn as nn

import aesmc_b200.state as _state

Normal = torch.distributions.Normal


class Initial:
    def __init__(self, loc, scale):
        self.loc, self.scale = loc, scale

    def __call__(self):
        return Normal(self.loc, self.scale)


class Transition(nn.Module):
    def __init__(self, init_mult, scale, state=_state):
        super().__init__()
        self.mult = nn.Parameter(torch.tensor(float(init_mult)))
        self.scale = scale
        self._state = state

    def forward(self, previous_latents=None, time=None, previous_observations=None):
        dist = Normal(self.mult * previous_latents[-1], self.scale)
        return self._state.set_batch_shape_mode(dist, self._state.BatchShapeMode.FULLY_EXPANDED)


class Emission(nn.Module):
    def __init__(self, init_mult, scale, state=_state):
        super().__init__()
        self.mult = nn.Parameter(torch.tensor(float(init_mult)))
        self.scale = scale
        self._state = state

    def forward(self, latents=None, time=None, previous_observations=None):
        dist = Normal(self.mult * latents[-1], self.scale)
        return self._state.set_batch_shape_mode(dist, self._state.BatchShapeMode.FULLY_EXPANDED)


class Proposal(nn.Module):
    """q(x_0 | y_0) = N(lin_0(y_0), scale_0^2);  q(x_t | x_{t-1}, y_t) = N(lin_t([x_{t-1}, y_t]), scale_t^2)."""

    def __init__(self, scale_0, scale_t, state=_state):
        super().__init__()
        self.scale_0, self.scale_t = scale_0, scale_t
        self.lin_0 = nn.Linear(1, 1)
        self.lin_t = nn.Linear(2, 1)
        self._state = state

    def forward(self, previous_latents=None, time=None, observations=None):
        modes = self._state.BatchShapeMode
        if time == 0:
            loc = self.lin_0(observations[0].unsqueeze(-1)).squeeze(-1)
            return self._state.set_batch_shape_mode(Normal(loc=loc, scale=self.scale_0), modes.BATCH_EXPANDED)
        prev = previous_latents[-1]
        num_particles = prev.shape[1]
        feats = torch.cat([prev.unsqueeze(-1),
                           observations[time].view(-1, 1, 1).expand(-1, num_particles, 1)], dim=2)
        loc = self.lin_t(feats.view(-1, 2)).squeeze(-1).view(-1, num_particles)
        return self._state.set_batch_shape_mode(Normal(loc=loc, scale=self.scale_t), modes.FULLY_EXPANDED)


# ---- bootstrap particle filter on a fixed scalar LGSSM (no learnable parameters) ------------------
def _scalar(v, device):
    return torch.tensor(float(v), device=device)


class BootstrapInitial:
    def __init__(self, mean, variance, device=None):
        self.mean, self.std = _scalar(mean, device), _scalar(math.sqrt(variance), device)

    def __call__(self):
        return Normal(loc=self.mean, scale=self.std)


class BootstrapTransition:
    def __init__(self, matrix, variance, offset=0.0):
        self.matrix, self.std, self.offset = matrix, math.sqrt(variance), offset

    def __call__(self, previous_latents=None, time=None, previous_observations=None):
        return Normal(loc=previous_latents[-1] * self.matrix + self.offset, scale=self.std)


class BootstrapEmission:
    def __init__(self, matrix, variance, offset=0.0):
        self.matrix, self.std, self.offset = matrix, math.sqrt(variance), offset

    def __call__(self, latents=None, time=None, previous_observations=None):
        return Normal(loc=latents[-1] * self.matrix + self.offset, scale=self.std)


class BootstrapProposal:
    """Proposal == prior dynamics, so log-weights reduce to the emission log-density."""

    def __init__(self, initial_mean, initial_variance, matrix, variance, offset=0.0, device=None):
        self.m0, self.s0 = _scalar(initial_mean, device), _scalar(math.sqrt(initial_variance), device)
        self.matrix, self.std, self.offset = matrix, math.sqrt(variance), offset

    def __call__(self, previous_latents=None, time=None, observations=None):
        if time == 0:
            return Normal(loc=self.m0, scale=self.s0)
        return Normal(loc=previous_latents[-1] * self.matrix + self.offset, scale=self.std)


def bootstrap_filter(m0=0.0, P0=1.0, A=0.9, Q=1.0, C=1.0, R=0.25, device=None):
    """(initial, transition, emission, proposal) of the BASELINE config-2 bootstrap filter.  ``device``
    is where the time-0 proposal draws its particles (everything downstream follows the latents)."""
    return (BootstrapInitial(m0, P0, device), BootstrapTransition(A, Q), BootstrapEmission(C, R),
            BootstrapProposal(m0, P0, A, Q, device=device))


def simulate(T, B, m0=0.0, P0=1.0, A=0.9, Q=1.0, C=1.0, R=0.25, seed=0):
    """Observations [T, B] float32 (numpy) simulated from the model with numpy's Generator(seed)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    x = m0 + math.sqrt(P0) * rng.standard_normal(B)
    ys = np.empty((T, B), np.float32)
    for t in range(T):
        if t > 0:
            x = A * x + math.sqrt(Q) * rng.standard_normal(B)
        ys[t] = C * x + math.sqrt(R) * rng.standard_normal(B)
    return ys
