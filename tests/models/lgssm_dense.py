"""BASELINE config 3: D-dimensional linear-Gaussian SSM with dense transition/emission matrices and a
learned Gaussian proposal (not in the reference; written against the aesmc callable conventions).

    x_0 ~ N(0, s0^2 I),  x_t = A x_{t-1} + N(0, q^2 I),  y_t = C x_t + N(0, r^2 I)

Latents are [batch, particles, Dx] tensors (event_shape [Dx]), so the ancestral gather moves Dx floats
per particle."""
import torch
import torch.nn as nn

import aesmc_b200.state as state

Normal, Independent = torch.distributions.Normal, torch.distributions.Independent
FULL, BATCH = state.BatchShapeMode.FULLY_EXPANDED, state.BatchShapeMode.BATCH_EXPANDED


def make_system(dx=10, dy=10, seed=0, device=None):
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(dx, dx, generator=g)
    A = 0.9 * A / torch.linalg.eigvals(A).abs().max()  # spectral radius 0.9
    C = torch.randn(dy, dx, generator=g) / dx ** 0.5
    return A.to(device), C.to(device)


class Initial:
    def __init__(self, dx, scale, device=None):
        self.loc = torch.zeros(dx, device=device)
        self.scale = scale

    def __call__(self):
        return Independent(Normal(self.loc, self.scale), 1)


class Transition(nn.Module):
    def __init__(self, A, scale, learn=False):
        super().__init__()
        self.A = nn.Parameter(A.clone()) if learn else A
        self.scale = scale

    def forward(self, previous_latents=None, time=None, previous_observations=None):
        d = Independent(Normal(previous_latents[-1] @ self.A.T, self.scale), 1)
        return state.set_batch_shape_mode(d, FULL)


class Emission(nn.Module):
    def __init__(self, C, scale, learn=False):
        super().__init__()
        self.C = nn.Parameter(C.clone()) if learn else C
        self.scale = scale

    def forward(self, latents=None, time=None, previous_observations=None):
        d = Independent(Normal(latents[-1] @ self.C.T, self.scale), 1)
        return state.set_batch_shape_mode(d, FULL)


class Proposal(nn.Module):
    """q(x_0 | y_0) = N(W0 y_0, .), q(x_t | x_{t-1}, y_t) = N(W [x_{t-1}, y_t], .) with learned scales."""

    def __init__(self, dx, dy):
        super().__init__()
        self.lin_0 = nn.Linear(dy, dx)
        self.lin_t = nn.Linear(dx + dy, dx)
        self.log_scale_0 = nn.Parameter(torch.zeros(dx))
        self.log_scale_t = nn.Parameter(torch.zeros(dx))

    def forward(self, previous_latents=None, time=None, observations=None):
        if time == 0:
            d = Independent(Normal(self.lin_0(observations[0]), self.log_scale_0.exp()), 1)
            return state.set_batch_shape_mode(d, BATCH)
        prev = previous_latents[-1]
        y = observations[time].unsqueeze(1).expand(-1, prev.shape[1], -1)
        d = Independent(Normal(self.lin_t(torch.cat([prev, y], dim=-1)), self.log_scale_t.exp()), 1)
        return state.set_batch_shape_mode(d, FULL)


class PriorProposal:
    """Bootstrap proposal (prior dynamics) for checking the evidence against the Kalman filter."""

    def __init__(self, initial, transition):
        self.initial, self.transition = initial, transition

    def __call__(self, previous_latents=None, time=None, observations=None):
        if time == 0:
            return self.initial()
        return self.transition(previous_latents=previous_latents, time=time)


def simulate(A, C, T, B, s0, q, r, seed=0):
    g = torch.Generator().manual_seed(seed)
    A, C = A.cpu(), C.cpu()
    x = s0 * torch.randn(B, A.shape[0], generator=g)
    ys = []
    for t in range(T):
        if t:
            x = x @ A.T + q * torch.randn(B, A.shape[0], generator=g)
        ys.append(x @ C.T + r * torch.randn(B, C.shape[0], generator=g))
    return torch.stack(ys)  # [T, B, Dy]
