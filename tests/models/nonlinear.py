"""BASELINE config 4: the classic nonlinear state-space benchmark with an MLP proposal (not in the
reference; written against the aesmc callable conventions).

    x_0 ~ N(0, 5),  x_t = x/2 + 25 x/(1+x^2) + 8 cos(1.2 t) + N(0, s_x^2),  y_t = a x_t^2 + N(0, s_y^2)
"""
import math

import torch
import torch.nn as nn

import aesmc_b200.state as state

Normal = torch.distributions.Normal
FULL, BATCH = state.BatchShapeMode.FULLY_EXPANDED, state.BatchShapeMode.BATCH_EXPANDED


def drift(x, t):
    return x / 2 + 25 * x / (1 + x * x) + 8 * math.cos(1.2 * t)


class Initial:
    def __init__(self, device=None):
        self.loc = torch.zeros((), device=device)
        self.scale = math.sqrt(5.0)

    def __call__(self):
        return Normal(self.loc, self.scale)


class Transition(nn.Module):
    def __init__(self, scale=math.sqrt(10.0)):
        super().__init__()
        self.log_scale = nn.Parameter(torch.tensor(math.log(scale)))

    def forward(self, previous_latents=None, time=None, previous_observations=None):
        d = Normal(drift(previous_latents[-1], time), self.log_scale.exp())
        return state.set_batch_shape_mode(d, FULL)


class Emission(nn.Module):
    def __init__(self, mult=0.05, scale=1.0):
        super().__init__()
        self.mult = nn.Parameter(torch.tensor(float(mult)))
        self.scale = scale

    def forward(self, latents=None, time=None, previous_observations=None):
        x = latents[-1]
        return state.set_batch_shape_mode(Normal(self.mult * x * x, self.scale), FULL)


class Proposal(nn.Module):
    """MLP on (drift(x_{t-1}), y_t) -> (mean, log-scale); at t = 0 on y_0 alone."""

    def __init__(self, hidden=32):
        super().__init__()
        self.net_0 = nn.Sequential(nn.Linear(1, hidden), nn.Tanh(), nn.Linear(hidden, 2))
        self.net_t = nn.Sequential(nn.Linear(2, hidden), nn.Tanh(), nn.Linear(hidden, 2))

    def forward(self, previous_latents=None, time=None, observations=None):
        if time == 0:
            out = self.net_0(observations[0].unsqueeze(-1))
            return state.set_batch_shape_mode(Normal(out[..., 0], out[..., 1].clamp(-5, 3).exp()), BATCH)
        prev = previous_latents[-1]
        y = observations[time].unsqueeze(1).expand_as(prev)
        feats = torch.stack([drift(prev, time) / 10, y / 10], dim=-1)
        out = self.net_t(feats)
        loc = drift(prev, time) + out[..., 0]
        return state.set_batch_shape_mode(Normal(loc, out[..., 1].clamp(-5, 3).exp()), FULL)
