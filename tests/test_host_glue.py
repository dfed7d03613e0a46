"""Host-side glue of the sampler plugin that needs no GPU."""
import ctypes

import numpy as np
import pytest

from optuna_b200 import _lib, mini
from optuna_b200.sampler import B200TPESampler, default_gamma, default_weights, hyperopt_default_gamma
from tests._util import draw_uniforms


def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()
    assert lib.tpe_abi_version() == _lib.ABI_VERSION == 2
    import os, re
    hdr = open(os.path.join(os.path.dirname(_lib.LIB_PATH), "..", "include", "optuna_b200_tpe.h")).read()
    declared = set(re.findall(r"\b(tpe_[a-z0-9_]+)\s*\(", hdr)) - {"tpe_ctx"}
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name)


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from optuna_b200 import TPEEngine
    with pytest.raises(RuntimeError):
        TPEEngine(0)


def test_single_random_sample_call_is_the_reference_stream():
    C = 24
    for ncat, nnum in ((0, 1), (2, 3), (12, 52), (1, 0)):
        a, b = np.random.RandomState(9), np.random.RandomState(9)
        assert np.array_equal(draw_uniforms(a, C, ncat, nnum), b.random_sample(C * (1 + ncat + nnum)))


def test_gamma_and_weights_mirror_reference_formulas():
    assert [default_gamma(n) for n in (0, 1, 10, 11, 250, 100000)] == [0, 1, 1, 2, 25, 25]
    assert [hyperopt_default_gamma(n) for n in (1, 16, 100, 10 ** 6)] == [1, 1, 3, 25]
    assert default_weights(0).size == 0 and np.array_equal(default_weights(7), np.ones(7))
    w = default_weights(100)
    assert w.shape == (100,) and w[0] == 0.01 and np.all(w[-25:] == 1) and np.all(np.diff(w[:75]) > 0)


def test_sampler_constructor_contract():
    with pytest.raises(ValueError):
        B200TPESampler(group=True)  # needs multivariate
    s = B200TPESampler(seed=1, multivariate=True, group=True, constant_liar=True)
    assert set(B200TPESampler.hyperopt_parameters()) >= {"gamma", "weights", "n_startup_trials"}
    # startup trials never touch the device
    study = mini.create_study(sampler=B200TPESampler(seed=0, n_startup_trials=5))
    study.optimize(lambda t: t.suggest_float("x", 0, 1) + t.suggest_int("k", 1, 3), n_trials=5)
    assert len(study.trials) == 5 and study.sampler._engine is None


def test_device_synced_rng_flushes_on_access_and_pickle():
    """sampler._DeviceSyncedRng: while the newer MT19937 state lives on the device, any access to
    `.rng` (and pickling) first copies it back -- here with a stand-in engine."""
    import pickle
    from optuna_b200.mini import LazyRandomState
    from optuna_b200.sampler import _DeviceSyncedRng

    class FakeEngine:
        def __init__(self, end_state):
            self.end_state, self.calls = end_state, 0

        def finish_rng(self, rng):
            self.calls += 1
            rng.set_state(self.end_state)

    ahead = np.random.RandomState(3)
    ahead.random_sample(1000)                      # what the device would have drawn
    proxy = _DeviceSyncedRng(LazyRandomState(3))
    eng = FakeEngine(ahead.get_state())
    assert not proxy.on_device(eng)
    proxy.mark_device(eng)
    assert proxy.on_device(eng) and not proxy.on_device(object())
    clone = pickle.loads(pickle.dumps(proxy))      # pickling flushes
    assert eng.calls == 1 and not proxy.on_device(eng)
    want = ahead.random_sample(4)
    assert np.array_equal(clone.rng.random_sample(4), want)
    assert np.array_equal(proxy.rng.random_sample(4), want) and eng.calls == 1
