"""MOTPE on the CUDA path vs live-reference goldens (tests/golden/motpe.npz):
split by non-domination rank + greedy HSSP, hypervolume weights, full suggestions."""
import numpy as np
import pytest

from oracle import motpe as mo
from tests._util import draw_uniforms, load

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from optuna_b200 import TPEEngine
    e = TPEEngine(0)
    yield e
    e.close()


def _setup(eng, X, vals, cat=None):
    from optuna_b200.engine import ParamSpec
    n, P = X.shape
    eng.set_space([ParamSpec(kind=0, low=0.0, high=1.0) for _ in range(P)])
    eng.set_history(X, np.zeros(n, np.int8) if cat is None else cat, np.zeros((n, 2)))
    eng.set_values(vals, 0)


def test_mo_split_and_weights_match_reference(eng):
    g = load("motpe.npz")
    rs = np.random.RandomState(0)
    for ci in range(int(g["mo_n"])):
        t = f"mo{ci}/"
        v, nb, want = g[t + "v"], int(g[t + "nb"]), g[t + "below"]
        if not np.isfinite(v).all():
            v = v.copy()  # +-inf objective values are legal (test_sampler.py:811-861)
        n = v.shape[0]
        _setup(eng, rs.uniform(0, 1, (n, 2)), v)
        info = eng.prepare([0, 1], n_below=nb, n_candidates=8, multivariate=True)
        below, above = eng.get_split()
        assert np.array_equal(below, want), (ci, below, want)
        assert info[0] == want.size
        if 0 < want.size:
            eng.build()
            w = eng.get_mo_weights()
            ref = g[t + "w"]
            np.testing.assert_allclose(w, ref, rtol=1e-9, atol=1e-15, err_msg=str(ci))


def test_motpe_suggestions_match_reference(eng):
    g = load("motpe.npz")
    for ci in range(int(g["mosg_n"])):
        t = f"mosg{ci}/"
        mv, C, seed, n_below, m = g[t + "cfg"]
        mv, C, n_below = bool(mv), int(C), int(n_below)
        X, vals = g[t + "X"], g[t + "values"]
        P = X.shape[1]
        _setup(eng, X, vals)
        rng = np.random.RandomState(int(seed))
        calls = [list(range(P))] if mv else [[j] for j in range(P)]
        ret = []
        for q, cols in enumerate(calls):
            u = draw_uniforms(rng, C, 0, len(cols))
            x, acq, best = eng.suggest(cols, u, 1, n_below=n_below, n_candidates=C, multivariate=mv)
            smp, ll, lg = eng.get_candidates()
            wb = eng.get_mixture(0)[0]
            # leave-one-out contributions are differences of hypervolumes (cancellation ~ hv / contrib) and
            # the reference sums its 2-D / 3-D hypervolumes inside BLAS dot products (order unspecified)
            np.testing.assert_allclose(wb, g[f"{t}c{q}/wb"], rtol=1e-9, atol=1e-15)
            np.testing.assert_allclose(smp, g[f"{t}c{q}/samples"], rtol=1e-11, atol=1e-12)
            np.testing.assert_allclose(ll, g[f"{t}c{q}/ll"], rtol=0, atol=1e-11)
            np.testing.assert_allclose(lg, g[f"{t}c{q}/lg"], rtol=0, atol=1e-12)
            assert int(best[0]) == int(np.argmax(g[f"{t}c{q}/ll"] - g[f"{t}c{q}/lg"]))
            ret.extend(x[0].tolist())
        np.testing.assert_allclose(ret, g[t + "ret"], rtol=1e-11, atol=1e-12)


def test_mo_rank_properties_at_scale(eng):
    """Config-4 shape (N = 20 000, 4 objectives): the selected below set is exactly what the oracle
    selects (ranks by peeling + HSSP on the tie rank)."""
    rs = np.random.RandomState(3)
    n = 20000
    X = rs.uniform(0, 1, (n, 8))
    cs = np.array([0.2, 0.4, 0.6, 0.8])
    vals = np.stack([((X - c) ** 2).sum(1) for c in cs], 1)
    _setup(eng, X, vals)
    eng.prepare(list(range(8)), n_below=25, n_candidates=24, multivariate=True)
    below, above = eng.get_split()
    want = mo.split_complete_mo(vals, 25)
    assert np.array_equal(below, want)
    assert below.size + above.size == n
    eng.build()
    np.testing.assert_allclose(eng.get_mo_weights(), mo.weights_below_mo(vals[want]), rtol=1e-9, atol=1e-15)


@pytest.mark.parametrize("M", [2, 3, 4])
def test_large_below_sets_gamma_ten_percent(eng, M):
    """SURVEY.md 8d, config 4 with gamma = ceil(0.1 n): the below set has hundreds to thousands of trials -- no cap on
    the device (the reference has none: sampler.py:745-779, :824-863, hssp.py:143-176).  Split and weights against the
    oracle; rows lacking a selected parameter pick their weight by position (weights_below[param_mask_below])."""
    rs = np.random.RandomState(M)
    N, P = (20000, 8) if M == 4 else (6000, 4)
    X = rs.uniform(0, 1, (N, P))
    X[rs.uniform(size=N) < 0.1, 1] = np.nan           # conditional parameter: absent in 10 % of the trials
    cs = np.array([0.2, 0.4, 0.6, 0.8])[:M]
    vals = ((np.nan_to_num(X, nan=0.5)[:, None, :] - cs[None, :, None]) ** 2).sum(2)
    cat = np.zeros(N, np.int8)
    cat[rs.uniform(size=N) < 0.02] = 2                # a few infeasible trials (EPS weight when they end up below)
    _setup(eng, X, vals, cat)
    nb = int(np.ceil(0.1 * N))
    info = eng.prepare([0, 1], n_below=nb, n_candidates=8, multivariate=True)
    below, above = eng.get_split()
    comp = np.flatnonzero(cat == 0)
    want = comp[mo.split_complete_mo(vals[comp], min(nb, comp.size))]
    assert info[0] == nb
    rest = nb - want.size                              # filled from the infeasible group, by violation then trial order
    infe = np.flatnonzero(cat == 2)[:rest]
    all_below = np.sort(np.concatenate([want, infe]))
    has = ~np.isnan(X[all_below][:, [0, 1]]).any(1)
    assert np.array_equal(below, all_below[has])
    eng.build()
    w = eng.get_mo_weights()
    ref = mo.weights_below_mo(vals[all_below], cat[all_below] != 2)
    np.testing.assert_allclose(w, ref, rtol=1e-9, atol=1e-15)
    wb = eng.get_mixture(0)[0]                         # mixture weights: the rows holding both parameters + prior
    raw = np.append(ref[has], 1.0)
    np.testing.assert_allclose(wb, raw / raw.sum(), rtol=1e-9, atol=1e-18)
