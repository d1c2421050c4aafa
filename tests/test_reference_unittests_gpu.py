"""The reference's OWN unittest files, unmodified (tests/golden/ref_tests/test/, byte-for-byte copies of
tuananhle7/aesmc test/), executed against this repository through aesmc_b200.install_as_aesmc():
`import aesmc.inference as inference` etc. inside those files resolve to the sm_100a path.

    test_math.py, test_state.py, test_statistics.py           whole files
    test_inference.py                                          whole file: TestGetResampledLatentStates,
                                                               TestSampleAncestralIndex (:13-84) and TestInfer
                                                               (IS / SMC against a Kalman smoother)
    test_losses.py                                             both training tests (iwae + aesmc), opt-in long
"""
import importlib
import os
import sys
import unittest

import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
REF_TESTS = os.path.join(HERE, "golden", "ref_tests")


@pytest.fixture()
def reference_tests(tmp_path, monkeypatch):
    """`test.*` importable from the vendored copy, `aesmc` = this package, matplotlib / pykalman stubbed,
    cwd = a scratch directory with the plot folders the tests write to."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import aesmc_b200
    saved = {k: v for k, v in sys.modules.items()
             if k == "aesmc" or k.startswith("aesmc.") or k == "test" or k.startswith("test.")
             or k in ("matplotlib", "matplotlib.pyplot", "pykalman")}
    for k in list(saved):
        del sys.modules[k]
    sys.path.insert(0, REF_TESTS)
    sys.path.insert(0, os.path.join(REF_TESTS))
    from tests.golden.ref_tests import stubs
    added = stubs.install()
    aesmc_b200.install_as_aesmc(force=True)
    (tmp_path / "test" / "test_inference_plots").mkdir(parents=True)
    (tmp_path / "test" / "test_autoencoder_plots").mkdir(parents=True)
    monkeypatch.chdir(tmp_path)
    try:
        yield
    finally:
        sys.path.remove(REF_TESTS)
        if REF_TESTS in sys.path:
            sys.path.remove(REF_TESTS)
        for k in [k for k in sys.modules if k == "aesmc" or k.startswith("aesmc.") or k == "test"
                  or k.startswith("test.")] + added:
            sys.modules.pop(k, None)
        sys.modules.update(saved)


def _run(module_name, pattern=None):
    mod = importlib.import_module(module_name)
    suite = unittest.defaultTestLoader.loadTestsFromModule(mod)
    if pattern is not None:
        keep = unittest.TestSuite()
        for group in suite:
            for case in group:
                if pattern(case.id()):
                    keep.addTest(case)
        suite = keep
    n = suite.countTestCases()
    assert n > 0, "no tests collected from %s" % module_name
    result = unittest.TextTestRunner(verbosity=0).run(suite)
    problems = ["%s\n%s" % (case.id(), tb) for case, tb in result.failures + result.errors]
    assert not problems, "\n\n".join(problems)
    assert result.testsRun == n
    return n


@pytest.mark.parametrize("name,expected", [("test.test_math", 6), ("test.test_state", None),
                                           ("test.test_statistics", None)])
def test_reference_unittest_file(reference_tests, name, expected):
    n = _run(name)
    if expected is not None:
        assert n >= expected


def test_reference_test_inference_hot_path_classes(reference_tests):
    """test/test_inference.py:13-84 -- TestGetResampledLatentStates and TestSampleAncestralIndex."""
    _run("test.test_inference", lambda tid: "TestInfer" not in tid)


def test_reference_test_inference_kalman(reference_tests):
    """test/test_inference.py:146-375 -- IS and SMC (K = 1000, T = 100) against the Kalman smoother."""
    _run("test.test_inference", lambda tid: "TestInfer" in tid)


@pytest.mark.skipif(os.environ.get("AESMC_REF_TESTS_LONG", "0") != "1",
                    reason="2 x 500 + 2000 optimiser steps of tiny launch-bound batches; set AESMC_REF_TESTS_LONG=1")
def test_reference_test_losses(reference_tests):
    _run("test.test_losses")
