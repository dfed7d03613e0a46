"""B200TPESampler driven the way the reference drives its sampler (study.optimize / ask+tell),
checked against live-reference golden trajectories (tests/golden/branin.npz, suggest.npz)."""
import math
import pickle

import numpy as np
import pytest

from tests._util import load

pytestmark = pytest.mark.gpu


def branin(t):
    x = t.suggest_float("x", -5, 10)
    y = t.suggest_float("y", 0, 15)
    return ((y - 5.1 / (4 * math.pi**2) * x * x + 5 / math.pi * x - 6) ** 2
            + 10 * (1 - 1 / (8 * math.pi)) * math.cos(x) + 10)


@pytest.mark.parametrize("mv", [False, True])
def test_branin_200_trials_matches_reference_trajectory(mv):
    """BASELINE config 1: same seed => same 200-trial trajectory as optuna.samplers.TPESampler."""
    from optuna_b200 import B200TPESampler, mini
    g = load("branin.npz")
    tag = "mv" if mv else "uni"
    study = mini.create_study(sampler=B200TPESampler(seed=0, multivariate=mv))
    study.optimize(branin, n_trials=200)
    xy = np.asarray([[t.params["x"], t.params["y"]] for t in study.trials])
    ref = g[f"branin_{tag}/xy"]
    # startup trials (RandomSampler stream) are bit-identical
    assert np.array_equal(xy[:10], ref[:10])
    np.testing.assert_allclose(xy, ref, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose([t.value for t in study.trials], g[f"branin_{tag}/values"], rtol=1e-8, atol=1e-8)
    if not mv:
        assert abs(study.best_value - 0.4069652013131506) < 1e-9


def test_sampler_contract_types_and_ranges():
    """optuna/testing/pytest_samplers.py:81-491 in miniature: python-native return types, values in
    range for every distribution kind, uni- and multivariate, n_startup_trials=0."""
    from optuna_b200 import B200TPESampler, mini

    def obj(t):
        a = t.suggest_float("a", -1.0, 1.0)
        b = t.suggest_float("b", 1e-3, 10.0, log=True)
        c = t.suggest_float("c", 0.0, 2.0, step=0.25)
        d = t.suggest_int("d", -3, 7)
        e = t.suggest_int("e", 1, 64, log=True)
        f = t.suggest_int("f", 0, 30, step=5)
        g = t.suggest_categorical("g", ["p", "q", None, 3])
        assert type(a) is float and type(b) is float and type(c) is float
        assert type(d) is int and type(e) is int and type(f) is int
        assert -1 <= a <= 1 and 1e-3 <= b <= 10 and c in np.arange(0, 2.01, 0.25)
        assert -3 <= d <= 7 and 1 <= e <= 64 and f in range(0, 31, 5) and g in ["p", "q", None, 3]
        return a * a + math.log(b) ** 2 + c + d * 0.1 + (g == "p")

    for mv in (False, True):
        s = mini.create_study(sampler=B200TPESampler(seed=3, multivariate=mv, n_startup_trials=0))
        s.optimize(obj, n_trials=25)
        assert len(s.trials) == 25


def test_reproducible_and_picklable():
    from optuna_b200 import B200TPESampler, mini

    def run(sampler):
        s = mini.create_study(sampler=sampler, direction="maximize")
        s.optimize(lambda t: -(t.suggest_float("x", 0, 1) - 0.3) ** 2 + t.suggest_int("k", 0, 5) * 0.0, n_trials=30)
        return [t.params for t in s.trials]

    a = run(B200TPESampler(seed=11))
    b = run(pickle.loads(pickle.dumps(B200TPESampler(seed=11))))
    assert a == b
    assert run(B200TPESampler(seed=12)) != a


def test_custom_weights_gamma_pruned_constraints_run():
    from optuna_b200 import B200TPESampler, mini

    def obj(t):
        x = t.suggest_float("x", -3, 3)
        t.set_user_attr("c", x - 1.0)
        if t.number % 5 == 4:
            t.report(abs(x), 1)
            raise mini.TrialPruned()
        return x * x

    sampler = B200TPESampler(seed=1, gamma=lambda n: max(1, n // 4), weights=lambda n: np.arange(1, n + 1) ** 0.5,
                             constraints_func=lambda tr: (tr.user_attrs["c"],), n_startup_trials=5)
    s = mini.create_study(sampler=sampler)
    s.optimize(obj, n_trials=40)
    assert len(s.trials) == 40
    with pytest.raises(ValueError):
        bad = mini.create_study(sampler=B200TPESampler(seed=1, weights=lambda n: -np.ones(n), n_startup_trials=2))
        bad.optimize(lambda t: t.suggest_float("x", 0, 1), n_trials=5)


def test_constant_liar_and_categorical_distance_through_plugin():
    from optuna_b200 import B200TPESampler, mini

    def obj(t):
        a = t.suggest_categorical("a", [0, 1, 2, 3])
        x = t.suggest_float("x", -2, 2)
        return (a - 2) ** 2 + x * x

    sampler = B200TPESampler(seed=2, multivariate=True, constant_liar=True, n_startup_trials=4,
                             categorical_distance_func={"a": lambda p, q: abs(p - q)})
    s = mini.create_study(sampler=sampler)
    # a batch of asks before any tell: later asks see earlier RUNNING trials in g(x) (sampler.py:526-535)
    s.optimize(obj, n_trials=8)
    batch = [s.ask() for _ in range(4)]
    vals = [obj(t) for t in batch]
    for t, v in zip(batch, vals):
        s.tell(t, v)
    assert len(s.trials) == 12 and all(t.state == mini.TrialState.COMPLETE for t in s.trials)
    assert any("tpe:relative_params:0" in t.system_attrs for t in s.trials[8:])


def test_batched_ask_equals_sequential_asks():
    """BASELINE config 5 semantics: ask_batch(n) == n sequential study.ask() with no tell between."""
    from optuna_b200 import B200TPESampler, mini
    from optuna_b200.batch import ask_batch

    def obj(t):
        return sum((t.suggest_float(f"x{j}", 0, 1) - 0.3) ** 2 for j in range(5)) + t.suggest_int("k", 0, 9) * 0.01

    def warm(seed):
        s = mini.create_study(sampler=B200TPESampler(seed=seed, multivariate=True, n_ei_candidates=32))
        s.optimize(obj, n_trials=30)
        return s

    a, b = warm(7), warm(7)
    batch = ask_batch(a, 50)
    seq = [b.ask() for _ in range(50)]
    pa = [[t.suggest_float(f"x{j}", 0, 1) for j in range(5)] + [t.suggest_int("k", 0, 9)] for t in batch]
    pb = [[t.suggest_float(f"x{j}", 0, 1) for j in range(5)] + [t.suggest_int("k", 0, 9)] for t in seq]
    assert pa == pb
    assert len({tuple(p) for p in pa}) > 40  # different uniforms per ask


def test_group_decomposed_search_space():
    """group=True: conditional parameters are sampled jointly within the groups that co-occur
    (sampler.py:394-405, :417-431)."""
    from optuna_b200 import B200TPESampler, mini

    def obj(t):
        kind = t.suggest_categorical("kind", ["a", "b"])
        x = t.suggest_float("x", -1, 1)
        if kind == "a":
            return x * x + t.suggest_float("ya", 0, 2)
        return x * x + (t.suggest_int("yb", 0, 5) - 2) ** 2 + t.suggest_float("zb", 1e-2, 1, log=True)

    s = mini.create_study(sampler=B200TPESampler(seed=4, multivariate=True, group=True, n_startup_trials=6))
    s.optimize(obj, n_trials=40)
    assert len(s.trials) == 40
    groups = s.sampler._hist.groups
    names = sorted(sorted(g) for g in groups)
    assert names == [["kind", "x"], ["ya"], ["yb", "zb"]]
    with pytest.raises(ValueError):
        B200TPESampler(group=True)


def test_device_generated_uniforms_give_the_same_suggestions():
    """Asks large enough for the device MT19937 (>= DEVICE_RNG_MIN uniforms): consecutive
    sample_relative calls, a foreign draw in between, and close() must all reproduce the suggestions
    computed from host-drawn uniforms of a plain same-seeded RandomState."""
    from optuna_b200 import B200TPESampler, TPEEngine, mini
    from optuna_b200.sampler import _spec_of
    P, C, n = 16, 1024, 300
    assert C * (1 + P) >= B200TPESampler.DEVICE_RNG_MIN
    rs = np.random.RandomState(2)
    space = {f"x{j:02d}": mini.FloatDistribution(0.0, 1.0) for j in range(P)}
    names = list(space)
    X = rs.uniform(0, 1, (n, P))
    loss = ((X - 0.4) ** 2).sum(1)
    sampler = B200TPESampler(seed=11, n_ei_candidates=C, multivariate=True)
    study = mini.create_study(sampler=sampler)
    study._storage.trials = [mini.FrozenTrial(i, mini.TrialState.COMPLETE, value=float(loss[i]),
                                              params=dict(zip(names, X[i].tolist())), distributions=space)
                             for i in range(n)]
    frozen = mini.FrozenTrial(n, mini.TrialState.RUNNING)
    got = []
    for it in range(5):
        got.append(sampler.sample_relative(study, frozen, space))
        if it == 2:
            sampler._rng.rng.random_sample(3)  # someone else consumes from the generator
    sampler.close()
    got.append(sampler.sample_relative(study, frozen, space))  # engine re-created
    eng = TPEEngine(0)
    eng.set_space([_spec_of(nm, space[nm], {}) for nm in names])
    eng.set_history(X, np.zeros(n, np.int8), np.stack([loss, np.zeros(n)], 1))
    ref_rng = np.random.RandomState(11)
    n_below = int(sampler._gamma(n))
    for it in range(6):
        u = ref_rng.random_sample(C * (1 + P))
        x, _, _ = eng.suggest(list(range(P)), u, 1, n_below=n_below, n_candidates=C, multivariate=True)
        assert [got[it][nm] for nm in names] == x[0].tolist(), it
        if it == 2:
            ref_rng.random_sample(3)
    assert np.array_equal(sampler._rng.rng.random_sample(4), ref_rng.random_sample(4))
    eng.close()
    sampler.close()


def test_constant_liar_incremental_rows_equal_a_full_rebuild():
    """constant_liar=True keeps RUNNING trials as rows that are refreshed in place (tpe_history_update)
    and promoted when they finish; the suggestions must equal those of a sampler that re-uploads the
    whole history at every ask -- including out-of-order tells and a failed trial (positions shift)."""
    from optuna_b200 import B200TPESampler, mini

    def obj(t):
        x = t.suggest_float("x", -2, 2)
        k = t.suggest_int("k", 0, 6)
        return x * x + (k - 3) ** 2

    def scenario(force_rebuild):
        sampler = B200TPESampler(seed=5, multivariate=True, constant_liar=True, n_startup_trials=5)
        if force_rebuild:
            orig = sampler._sync

            def sync(study, trial, space):
                sampler._hist.token = None
                return orig(study, trial, space)
            sampler._sync = sync
        s = mini.create_study(sampler=sampler)
        s.optimize(obj, n_trials=10)
        out = []
        pending = [s.ask() for _ in range(5)]
        vals = [obj(t) for t in pending]
        out += [dict(t.params) for t in pending]
        s.tell(pending[3], vals[3])                          # out of order
        s.tell(pending[1], state=mini.TrialState.FAIL)       # disappears from the list
        more = [s.ask() for _ in range(3)]
        out += [dict(t.params) for t in more if obj(t) is not None]
        s.tell(pending[0], vals[0])
        last = [s.ask() for _ in range(2)]
        out += [dict(t.params) for t in last if obj(t) is not None]
        sampler.close()
        return out

    a, b = scenario(False), scenario(True)
    assert a == b
    assert len({tuple(sorted(p.items())) for p in a}) > 5
