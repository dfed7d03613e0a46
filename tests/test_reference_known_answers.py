"""Known answers held by the reference's OWN tests for this path, applied to
the oracle -- independent of the golden vectors generated from the live reference.

* tests/samplers_tests/tpe_tests/test_parzen_estimator.py:54-181  (mu / sigma / weights of all kernel kinds)
* tests/samplers_tests/tpe_tests/test_parzen_estimator.py:250-323 (sigmas for the endpoints / magic-clip matrix)
* tests/hypervolume_tests/test_wfg.py:17-66                       (exact integer hypervolumes, 2 .. 9 objectives)
* tests/samplers_tests/tpe_tests/test_sampler.py:711-805           (split order incl. +-inf values)
"""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import motpe as mo
from oracle import tpe_oracle as orc

SPACE = [orc.Param("float", 1.0, 100.0), orc.Param("float", 1.0, 100.0, None, True),
         orc.Param("float", 1.0, 100.0, 3.0), orc.Param("int", 1.0, 100.0, 1.0), orc.Param("int", 1.0, 100.0, 1.0, True),
         orc.Param("cat", n_choices=3), orc.Param("cat", n_choices=4)]


@pytest.mark.parametrize("multivariate", [True, False])
def test_init_parzen_estimator_known_answers(multivariate):
    cfg = orc.Config(prior_weight=1.0, magic_clip=False, endpoints=False, multivariate=multivariate,
                     weights=lambda n: np.arange(n) + 1.0)
    mix = orc.build_mixture(np.ones((1, 7)), SPACE, cfg)
    np.testing.assert_allclose(mix.weights, [0.5, 0.5])
    L = np.log
    if multivariate:
        s0 = 0.2
        want = [([1.0, 50.5], [s0 * 99.0, 99.0]), ([0.0, L(100) / 2], [s0 * L(100), L(100)]),
                ([1.0, 50.5], [s0 * 102.0, 102.0]), ([1.0, 50.5], [s0 * 100.0, 100.0]),
                ([0.0, (L(100.5) + L(0.5)) / 2], [s0 * (L(100.5) - L(0.5)), L(100.5) - L(0.5)])]
    else:
        want = [([1.0, 50.5], [49.5, 99.0]), ([0.0, L(100) / 2], [L(100) / 2, L(100)]),
                ([1.0, 50.5], [49.5, 102.0]), ([1.0, 50.5], [49.5, 100.0]),
                ([0.0, (L(100.5) + L(0.5)) / 2], [(L(100.5) + L(0.5)) / 2, L(100.5) - L(0.5)])]
    for j, (mu, sg) in enumerate(want):
        np.testing.assert_allclose(mix.mu[j], mu, rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(mix.sigma[j], sg, rtol=1e-12)
    np.testing.assert_allclose(mix.cat_w[5], [[0.2, 0.6, 0.2], [1 / 3, 1 / 3, 1 / 3]])
    np.testing.assert_allclose(mix.cat_w[6], [[1 / 6, 0.5, 1 / 6, 1 / 6], [0.25] * 4])


@pytest.mark.parametrize("mus, clip, endp, sigmas", [
    ([], False, True, [2.0]), ([0.4], False, True, [0.6, 2.0]), ([-0.4], False, True, [0.6, 2.0]),
    ([-0.4, 0.4], False, True, [0.6, 0.6, 2.0]), ([-0.4, 0.4], False, False, [0.4, 0.4, 2.0]),
    ([-0.4, 0.4, 0.41, 0.42], False, True, [0.6, 0.4, 0.01, 0.58, 2.0]),
    ([-0.4, 0.4, 0.41, 0.42], True, True, [0.6, 0.4, 1.0 / 3, 0.58, 2.0]),
])
def test_calculate_known_answers(mus, clip, endp, sigmas):
    cfg = orc.Config(prior_weight=1.0, magic_clip=clip, endpoints=endp, multivariate=False)
    mix = orc.build_mixture(np.asarray(mus, dtype=float).reshape(-1, 1), [orc.Param("float", -1.0, 1.0)], cfg)
    np.testing.assert_allclose(mix.sigma[0], sigmas, rtol=1e-12)
    np.testing.assert_allclose(mix.mu[0], list(mus) + [0.0], atol=1e-15)
    np.testing.assert_allclose(mix.weights, np.full(len(mus) + 1, 1.0 / (len(mus) + 1)))


def _shuffle_filter(s, assume_pareto, rng):
    rng.shuffle(s)
    return s[mo.is_pareto_front(s, False)] if assume_pareto else s


def _device_hv():
    here = os.path.dirname(os.path.abspath(__file__))
    lib = C.CDLL(os.path.join(here, "csrc", "libmath_shim.so"))
    lib.shim_hypervolume.restype = C.c_double

    def hv(v, ref, assume_pareto):
        v = np.ascontiguousarray(v, dtype=np.float64)
        ref = np.ascontiguousarray(ref, dtype=np.float64)
        return lib.shim_hypervolume(v.ctypes.data_as(C.c_void_p), C.c_int(v.shape[0]), C.c_int(v.shape[1]),
                                    ref.ctypes.data_as(C.c_void_p), C.c_int(int(assume_pareto)))
    return hv


@pytest.mark.parametrize("assume_pareto", (True, False))
def test_exact_integer_hypervolumes(assume_pareto):
    dev = _device_hv() if os.path.exists(os.path.join(os.path.dirname(__file__), "csrc", "libmath_shim.so")) else None
    for n in range(2, 30):  # test_wfg.py:17-27
        rng = np.random.RandomState(42)
        s = np.empty((2 * n + 1, 2), dtype=float)
        s[:n] = np.stack([np.arange(n), np.arange(n)[::-1]], axis=-1)
        s[n:] = np.stack([np.arange(n + 1), np.arange(n + 1)[::-1]], axis=-1)
        s = _shuffle_filter(s, assume_pareto, rng)
        want = n * n - n * (n - 1) // 2
        assert mo.hypervolume(s, n * np.ones(2), assume_pareto) == want
        if dev:
            assert dev(s, n * np.ones(2), assume_pareto) == want
    for n in range(2, 10):  # test_wfg.py:30-45
        rng = np.random.RandomState(42)
        s = np.array([[x, y, n - 1 - x - y] for x in range(n) for y in range(n - x)], dtype=float)
        s = _shuffle_filter(s, assume_pareto, rng)
        want = n ** 3 - (n - 1) * n * (n + 1) // 6
        assert mo.hypervolume(s, n * np.ones(3), assume_pareto) == want
        if dev:
            assert dev(s, n * np.ones(3), assume_pareto) == want
    for m in range(2, 10):  # test_wfg.py:59-66
        rng = np.random.RandomState(42)
        s = np.vstack([np.identity(m), rng.randint(1, 10, size=(10, m))]).astype(float)
        s = _shuffle_filter(s, assume_pareto, rng)
        assert mo.hypervolume(s, 10 * np.ones(m), assume_pareto) == 10 ** m - 1
        if dev and m <= 8:
            assert dev(s, 10 * np.ones(m), assume_pareto) == 10 ** m - 1


def test_split_known_order_with_infinite_values():
    """test_sampler.py:711-805: minimise, values incl. -inf / +inf, pruned after complete, infeasible last."""
    cat = np.array([0, 0, 0, 0, 1, 1, 2, 3], dtype=np.int8)
    key = np.zeros((8, 2))
    key[:4, 0] = [float("inf"), 1.0, -float("inf"), 0.5]
    key[4] = (-3, 0.2)   # pruned at step 3
    key[5] = (-1, 0.1)   # pruned at step 1
    key[6, 0] = 2.0      # infeasible
    for n_below, want in [(0, []), (1, [2]), (2, [2, 3]), (3, [1, 2, 3]), (4, [0, 1, 2, 3]), (5, [0, 1, 2, 3, 4]),
                          (6, [0, 1, 2, 3, 4, 5]), (7, [0, 1, 2, 3, 4, 5, 6]), (8, [0, 1, 2, 3, 4, 5, 6])]:
        below, above = orc.split_trials(cat, key, n_below)
        assert below.tolist() == want
        assert sorted(below.tolist() + above.tolist()) == list(range(8))
