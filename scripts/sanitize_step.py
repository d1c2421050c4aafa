"""Small driver for compute-sanitizer (memcheck / racecheck / synccheck) over every step-kernel variant:
register-blocked (regular and irregular K, D = 1 and D > 1), generic, multi-CTA (both span sizes of the chained
exact scan, vector and scalar tails), fused, both modes; then the gather / backward / row-reduction kernels."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from aesmc_b200 import _ops, fused  # noqa: E402

dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(0)
for mode in ("exact", "fast"):
    for B, K, D in [(5, 4096, 1), (3, 1000, 1), (3, 2048, 3), (2, 1023, 1), (2, 30000, 1), (1, 70000, 2), (25, 33000, 1), (2, 40003, 1)]:
        spread = 8.0 if K == 4096 else 1.0  # collapsed weights on one shape: absorbed blocks joining pure runs
        a, b, c = [spread * torch.randn(B, K, device=dev, generator=gen) for _ in range(3)]
        x = torch.randn(B, K, D, device=dev, generator=gen) if D > 1 else torch.randn(B, K, device=dev, generator=gen)
        u = torch.rand(B, dtype=torch.float64, device=dev, generator=gen)
        flags = _ops.new_flags(dev)
        out = _ops.smc_step(a, b, c, u, x, flags, mode, True)
        out2 = _ops.smc_step(a, b, c, None, None, flags, mode, False)
        torch.cuda.synchronize()
        assert int(flags.item()) == 0 and int(out[2].max()) < K
    model = fused.ScalarLinearGaussianSSM(device=dev)
    obs = torch.randn(4, 3, device=dev, generator=gen)
    with torch.no_grad():
        r = fused.infer_fused(model, obs, 512, return_log_marginal_likelihood=True, resampling_mode=mode)
    torch.cuda.synchronize()
    assert torch.isfinite(r["log_marginal_likelihood"]).all()
# round 2: the second-generation exact row kernel at every CTA size (with / without latents), its fused-model instance
# with the training backward (aesmc_lg_step_bwd_f32), the vector model kernel, the per-output-tile resampling of the
# multi-CTA path on collapsed weights, the parent-centric gather backward and the multi-warp row statistics
for K in (1024, 2048, 4096, 8192, 16384):
    for spread in (1.0, 12.0):
        B = 3
        a, b, c = [spread * torch.randn(B, K, device=dev, generator=gen) for _ in range(3)]
        xs = torch.randn(B, K, device=dev, generator=gen)
        u = torch.rand(B, dtype=torch.float64, device=dev, generator=gen)
        flags = _ops.new_flags(dev)
        o1 = _ops.smc_step(a, b, c, u, xs, flags, "exact", True)
        o2 = _ops.smc_step(a, None, None, u, None, flags, "exact", True)
        torch.cuda.synchronize()
        assert int(flags.item()) == 0 and int(o1[2].max()) < K and int(o2[2].max()) < K
from tests.models import lgssm  # noqa: E402
from aesmc_b200 import losses  # noqa: E402
init, trans, emis, prop = lgssm.Initial(0.0, 1.0), lgssm.Transition(0.5, 1.0).to(dev), lgssm.Emission(0.5, 0.5).to(dev), lgssm.Proposal(0.9, 0.9).to(dev)
fused.link(init, trans, emis, prop)
obs = [torch.randn(3, device=dev, generator=gen) for _ in range(4)]
for K in (1024, 600):
    loss = losses.get_loss(obs, K, "aesmc", init, trans, emis, prop)
    loss.backward()
torch.cuda.synchronize()
A, C = torch.randn(10, 10, device=dev, generator=gen) * 0.2, torch.randn(7, 10, device=dev, generator=gen) * 0.3
vm = fused.VectorLinearGaussianSSM(torch.zeros(10), 1.0, A, 0.5, C, 0.7, device=dev)
with torch.no_grad():
    rv = fused.infer_fused_vector(vm, torch.randn(3, 2, 7, device=dev, generator=gen), 1000, return_log_marginal_likelihood=True)
assert torch.isfinite(rv["log_marginal_likelihood"]).all()
for B, K in [(2, 70000), (1, 300000)]:
    a = 20.0 * torch.randn(B, K, device=dev, generator=gen)          # collapsed: whole input tiles without offspring
    xs = torch.randn(B, K, device=dev, generator=gen)
    u = torch.rand(B, dtype=torch.float64, device=dev, generator=gen)
    flags = _ops.new_flags(dev)
    for mode in ("exact", "fast"):
        o = _ops.smc_step(a, None, None, u, xs, flags, mode, True)
        torch.cuda.synchronize()
        assert int(o[2].max()) < K and bool((o[2][:, 1:] >= o[2][:, :-1]).all())
for K in (1024, 4096):
    B = 3
    xg = torch.randn(B, K, device=dev, generator=gen).requires_grad_()
    ix = torch.sort(torch.randint(0, K, (B, K), device=dev, generator=gen), dim=1).values
    ix[0, 10:900] = ix[0, 10]
    ix[1] = ix[1, 5]
    for t in (ix.int(), ix):
        _ops.gather(xg, t, True).sum().backward()
big = torch.randn(160, 4096, device=dev, generator=gen)
_ops.logsumexp_rows(big); _ops.log_ess_rows(big); _ops.weighted_moments(torch.randn(160, 4096, device=dev, generator=gen), big)
big = torch.randn(600, 1024, device=dev, generator=gen)
_ops.logsumexp_rows(big); _ops.log_ess_rows(big); _ops.weighted_moments(torch.randn(600, 1024, device=dev, generator=gen), big)
torch.cuda.synchronize()
from aesmc_b200 import statistics  # noqa: E402
for D in (1, 2, 4, 10):
    B, K = 3, 1001 if D == 1 else 1000
    x = (torch.randn(B, K, D, device=dev, generator=gen) if D > 1 else torch.randn(B, K, device=dev, generator=gen)).requires_grad_()
    idx = torch.sort(torch.randint(0, K, (B, K), device=dev, generator=gen), dim=1).values
    collapsed = idx.clone()
    collapsed[0, 10:900] = collapsed[0, 10]     # a run of 890 children: the warp-cooperative completion
    collapsed[1] = collapsed[1, 0]              # everything under one parent, long childless stretch after it
    for ix, srt in ((idx.int(), True), (idx, True), (collapsed.int(), True), (idx.flip(1).contiguous(), False)):
        out = _ops.gather(x, ix, srt)
        out.backward(torch.ones_like(out))
lw = torch.randn(5, 4096, device=dev, generator=gen)
_ops.logsumexp_rows(lw); _ops.lognormexp_rows(lw, True); _ops.log_ess_rows(lw)
_ops.weighted_moments(torch.randn(5, 4096, device=dev, generator=gen), lw)
_ops.weighted_moments(torch.randn(5, 4096, 3, device=dev, generator=gen), lw)
_ops.logsumexp_rows(lw[:, :1001].contiguous()); _ops.log_ess_rows(lw[:, :1001].contiguous())
for K in (4096, 1001):
    v = torch.randn(5, K, device=dev, generator=gen).requires_grad_()
    mu = torch.randn(5, K, device=dev, generator=gen).requires_grad_()
    lp = _ops.normal_log_prob(torch.distributions.Normal(mu, 0.7, validate_args=False), v)
    lp.sum().backward()
_ops.compose_index(idx.int(), idx.int())
torch.cuda.synchronize()
print("sanitize driver ok")
