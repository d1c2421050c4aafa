"""Host-side logic that needs no GPU: the torch adapters between user distributions and the
[batch, particle] layout (they mirror the reference's test/test_state.py), argument validation, the
lazy history sequence, the train loop plumbing and the aesmc alias."""
import itertools
import sys
import warnings

import numpy as np
import pytest
import torch

import aesmc_b200
from aesmc_b200 import inference, losses, state, train
from aesmc_b200.state import BatchShapeMode

Normal = torch.distributions.Normal


def test_batch_shape_mode_explicit_and_inferred():
    # test/test_state.py:7-52
    B, K = 2, 3
    d = state.set_batch_shape_mode(Normal(torch.zeros(7), 1.0), BatchShapeMode.BATCH_EXPANDED)
    assert state.get_batch_shape_mode(d) is BatchShapeMode.BATCH_EXPANDED
    assert state.get_batch_shape_mode(Normal(0.0, 1.0), B, K) is BatchShapeMode.NOT_EXPANDED
    assert state.get_batch_shape_mode(Normal(torch.zeros(5), 1.0), B, K) is BatchShapeMode.NOT_EXPANDED
    assert state.get_batch_shape_mode(Normal(torch.zeros(5, 6), 1.0), B, K) is BatchShapeMode.NOT_EXPANDED
    with pytest.warns(RuntimeWarning):
        assert state.get_batch_shape_mode(Normal(torch.zeros(B), 1.0), B, K) is BatchShapeMode.BATCH_EXPANDED
    with pytest.warns(RuntimeWarning):
        assert state.get_batch_shape_mode(Normal(torch.zeros(B, 5), 1.0), B, K) is BatchShapeMode.BATCH_EXPANDED
    with pytest.warns(RuntimeWarning):
        assert state.get_batch_shape_mode(Normal(torch.zeros(B, K, 4), 1.0), B, K) is BatchShapeMode.FULLY_EXPANDED


def test_foreign_enum_members_are_matched_by_name():
    import enum

    class Foreign(enum.Enum):  # stands for the reference package's aesmc.state.BatchShapeMode
        NOT_EXPANDED = 0
        BATCH_EXPANDED = 1
        FULLY_EXPANDED = 2

    d = state.set_batch_shape_mode(Normal(torch.zeros(4, 5), 1.0), Foreign.FULLY_EXPANDED)
    assert state.sample(d, 4, 5).shape == (4, 5)
    with pytest.raises(ValueError):
        state.sample(state.set_batch_shape_mode(Normal(0.0, 1.0), "nonsense"), 4, 5)


@pytest.mark.parametrize("event", [(), (4,), (4, 5)])
def test_sample_shapes_all_modes(event):
    # test/test_state.py:86-163
    B, K = 2, 3
    modes = {
        BatchShapeMode.NOT_EXPANDED: (),
        BatchShapeMode.BATCH_EXPANDED: (B,),
        BatchShapeMode.FULLY_EXPANDED: (B, K),
    }
    for mode, lead in modes.items():
        d = state.set_batch_shape_mode(Normal(torch.zeros(lead + event), 1.0), mode)
        assert state.sample(d, B, K).shape == (B, K) + event
        out = state.sample({"a": d, "b": d}, B, K)
        assert set(out) == {"a", "b"} and out["a"].shape == (B, K) + event
    t = torch.zeros(B, K, 7)
    assert state.sample(t, B, K) is t
    with pytest.raises(AttributeError):
        state.sample(3.0, B, K)
    with pytest.raises(ValueError):
        state.sample(torch.distributions.Categorical(torch.ones(3)), B, K)


def test_sample_means_track_parameters():
    # test/test_state.py:165-193 (10-sigma bound)
    B, K = 3, 20000
    loc = torch.tensor([-2.0, 0.5, 4.0])
    d = state.set_batch_shape_mode(Normal(loc, 1.0), BatchShapeMode.BATCH_EXPANDED)
    torch.manual_seed(0)
    m = state.sample(d, B, K).mean(dim=1)
    assert torch.all((m - loc).abs() < 10 / np.sqrt(K))


def test_log_prob_shapes_and_values():
    # test/test_state.py:196-268
    B, K = 2, 3
    for event in [(), (4,), (4, 5)]:
        for lead in [(), (B,), (B, K)]:
            d = Normal(torch.randn(lead + event), 1.0)
            v = torch.randn((B, K) + event)
            lp = state.log_prob(d, v)
            assert lp.shape == (B, K)
            loc = d.loc
            if lead == (B,):
                loc = loc.unsqueeze(1)
            want = Normal(loc.expand((B, K) + event) if lead != () else loc, 1.0).log_prob(v).reshape(B, K, -1).sum(-1)
            torch.testing.assert_close(lp, want)
    oh = torch.distributions.OneHotCategorical(probs=torch.ones(B, K, 5) / 5)
    assert state.log_prob(oh, oh.sample()).shape == (B, K)
    both = state.log_prob({"a": Normal(0.0, 1.0), "b": Normal(1.0, 2.0)}, {"a": torch.zeros(B, K), "b": torch.ones(B, K)})
    torch.testing.assert_close(both, Normal(0.0, 1.0).log_prob(torch.zeros(B, K)) + Normal(1.0, 2.0).log_prob(torch.ones(B, K)))
    with pytest.raises(RuntimeError):
        state.log_prob(Normal(torch.zeros(2, 3, 4, 5), 1.0), torch.zeros(2, 3))
    with pytest.raises(AttributeError):
        state.log_prob("nope", torch.zeros(2, 3))
    with pytest.raises(ValueError):  # sample validation is on unless the distribution opted out
        state.log_prob(torch.distributions.Exponential(torch.ones(())), -torch.ones(2, 3))
    lax = torch.distributions.Exponential(torch.ones(()), validate_args=False)
    assert state.log_prob(lax, -torch.ones(2, 3)).shape == (2, 3)


def test_expand_observation():
    # test/test_state.py:306-334
    o = torch.rand(2, 4, 5)
    e = state.expand_observation(o, 3)
    assert e.shape == (2, 3, 4, 5) and torch.equal(e[:, 1], o)
    d = state.expand_observation({"a": torch.rand(2), "b": torch.rand(2, 7)}, 3)
    assert d["a"].shape == (2, 3) and d["b"].shape == (2, 3, 7)


def test_argument_validation_precedes_any_gpu_work():
    obs = [torch.zeros(2)]
    with pytest.raises(ValueError):
        inference.infer("pf", obs, None, None, None, None, 4)
    with pytest.raises(UnboundLocalError):
        losses.get_loss(obs, 4, "smc", None, None, None, None)
    with pytest.raises(ValueError):
        aesmc_b200.set_resampling_mode("approximate")
    assert aesmc_b200.get_resampling_mode() in ("exact", "fast")


def test_hot_path_fails_loudly_without_a_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        inference.sample_ancestral_index(torch.zeros(2, 3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        state.resample(torch.zeros(2, 3), torch.zeros(2, 3, dtype=torch.long))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        aesmc_b200.statistics.log_ess(torch.zeros(2, 3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        aesmc_b200.math.lognormexp(np.zeros((2, 3)))


def test_get_chained_params_and_dataset_plumbing():
    lin = torch.nn.Linear(2, 1)
    assert train.get_chained_params(None, lambda: 0) is None
    assert len(list(train.get_chained_params(lin, None, torch.nn.Linear(1, 1)))) == 4

    class Init:
        def __call__(self):
            return Normal(0.0, 1.0)

    class Trans:
        def __call__(self, previous_latents=None, time=None, previous_observations=None):
            return Normal(0.5 * previous_latents[-1], 1.0)

    class Emis:
        def __call__(self, latents=None, time=None, previous_observations=None):
            return Normal(latents[-1], 0.1)

    lat, obs = aesmc_b200.statistics.sample_from_prior(Init(), Trans(), Emis(), 5, 7)
    assert len(lat) == len(obs) == 5 and lat[0].shape == (7,) and obs[-1].shape == (7,)
    loader = train.get_synthetic_dataloader(Init(), Trans(), Emis(), 4, 3)
    batch = next(iter(loader))
    assert len(batch) == 4 and batch[0].shape == (3,)
    assert len(train.SyntheticDataset(Init(), Trans(), Emis(), 4, 3)) == sys.maxsize


def test_install_as_aesmc_alias():
    saved = {k: v for k, v in sys.modules.items() if k == "aesmc" or k.startswith("aesmc.")}
    try:
        for k in saved:
            del sys.modules[k]
        mod = aesmc_b200.install_as_aesmc()
        import aesmc
        import aesmc.state as st
        assert aesmc is mod and st is aesmc_b200.state and aesmc.inference is inference
        assert aesmc.__version__ == "0.1.0"
    finally:
        for k in [k for k in sys.modules if k == "aesmc" or k.startswith("aesmc.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_graph_wrappers_validate_arguments_before_touching_the_gpu():
    """inference.GraphedInfer / train.GraphedTrainStep: the argument checks that need no device."""
    from aesmc_b200 import inference, train
    obs = [torch.zeros(3) for _ in range(4)]
    with pytest.raises(ValueError, match="uniforms"):
        inference.GraphedInfer("smc", obs, None, None, None, None, 8, uniforms=None)
    with pytest.raises(ValueError, match="check_finite"):
        inference.GraphedInfer("smc", obs, None, None, None, None, 8, check_finite=False)
    with pytest.raises(ValueError, match="CUDA"):
        inference.GraphedInfer("smc", obs, None, None, None, None, 8)
    lin = torch.nn.Linear(1, 1)
    with pytest.raises(ValueError, match="capturable"):
        train.GraphedTrainStep(obs, 8, "aesmc", None, None, None, None, torch.optim.Adam(lin.parameters()))
    with pytest.raises(ValueError, match="CUDA"):
        train.GraphedTrainStep(obs, 8, "aesmc", None, None, None, None, torch.optim.SGD(lin.parameters(), lr=0.1))


def test_vendored_reference_tests_are_unmodified():
    """tests/golden/ref_tests/test/ must stay byte-identical to the reference's test/ package (checked where
    the reference checkout exists: the build container, not the GPU box)."""
    import filecmp
    import glob
    import os
    import pytest
    ref = "/root/reference/test"
    if not os.path.isdir(ref):
        pytest.skip("no reference checkout here")
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_tests", "test")
    files = sorted(glob.glob(os.path.join(ref, "**", "*.py"), recursive=True))
    assert len(files) >= 8
    for f in files:
        assert filecmp.cmp(f, os.path.join(here, os.path.relpath(f, ref)), shallow=False), f
