"""Generate golden vectors from the LIVE reference (runs only in the build container).

    PYTHONPATH=/root/reference:/tmp/stubs PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden.py

Needs /root/reference (read-only checkout of optuna @ 4df4b72) and a 6-line ``colorlog`` stub
on the path (optuna/logging.py:14 imports it); writes tests/golden/*.npz.  The fixtures are what
travels to the GPU box; nothing at test time reads /root/reference.

Every array stored here is produced by reference code (optuna.samplers._tpe.*), never by the
oracle -- the oracle and the CUDA path are both checked against them.
"""
from __future__ import annotations

import math
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden")

import optuna  # noqa: E402  (the reference)
from optuna.distributions import CategoricalDistribution, FloatDistribution, IntDistribution  # noqa: E402
from optuna.samplers import TPESampler  # noqa: E402
from optuna.samplers._tpe import _truncnorm  # noqa: E402
from optuna.samplers._tpe._erf import erf as ref_erf  # noqa: E402
from optuna.samplers._tpe.parzen_estimator import _ParzenEstimator, _ParzenEstimatorParameters  # noqa: E402
from optuna.samplers._tpe.sampler import _split_trials, default_weights  # noqa: E402
from optuna.trial import TrialState, create_trial  # noqa: E402

optuna.logging.set_verbosity(optuna.logging.ERROR)
warnings.filterwarnings("ignore")


def encode_space(space: dict) -> np.ndarray:
    """[P, 6] = kind(0 float, 1 int, 2 cat), low, high, step(nan=None), log, n_choices."""
    rows = []
    for d in space.values():
        if isinstance(d, CategoricalDistribution):
            rows.append([2, 0, 0, np.nan, 0, len(d.choices)])
        elif isinstance(d, IntDistribution):
            rows.append([1, d.low, d.high, d.step, float(d.log), 0])
        else:
            rows.append([0, d.low, d.high, np.nan if d.step is None else d.step, float(d.log), 0])
    return np.asarray(rows, dtype=float)


MIXED_SPACE = {
    "a_cat": CategoricalDistribution([10, 20, 30, 40]),
    "b_float": FloatDistribution(-3.0, 5.0),
    "c_logf": FloatDistribution(1e-4, 10.0, log=True),
    "d_stepf": FloatDistribution(0.0, 10.0, step=0.5),
    "e_int": IntDistribution(0, 50),
    "f_logint": IntDistribution(1, 512, log=True),
    "g_cat": CategoricalDistribution(["x", "y", "z", "w", "v", "u"]),
    "h_intstep": IntDistribution(-10, 20, step=3),
}


def random_external(space: dict, rng: np.random.RandomState) -> dict:
    out = {}
    for name, d in space.items():
        if isinstance(d, CategoricalDistribution):
            out[name] = d.choices[rng.randint(len(d.choices))]
        elif isinstance(d, IntDistribution):
            if d.log:
                v = int(np.clip(round(math.exp(rng.uniform(math.log(d.low), math.log(d.high)))),
                                d.low, d.high))
            else:
                v = int(d.low + d.step * rng.randint((d.high - d.low) // d.step + 1))
            out[name] = v
        else:
            if d.log:
                out[name] = float(math.exp(rng.uniform(math.log(d.low), math.log(d.high))))
            elif d.step is not None:
                out[name] = float(d.low + d.step * rng.randint(int(round((d.high - d.low) / d.step)) + 1))
            else:
                out[name] = float(rng.uniform(d.low, d.high))
    return out


def internal_matrix(trials, space: dict) -> np.ndarray:
    X = np.full((len(trials), len(space)), np.nan)
    for i, t in enumerate(trials):
        for j, (name, d) in enumerate(space.items()):
            if name in t.params:
                X[i, j] = d.to_internal_repr(t.params[name])
    return X


def mixture_arrays(mpe: _ParzenEstimator, prefix: str, out: dict) -> None:
    mix = mpe._mixture_distribution
    out[prefix + "w"] = np.asarray(mix.weights)
    for j, d in enumerate(mix.distributions):
        if hasattr(d, "weights"):
            out[f"{prefix}cat{j}"] = np.asarray(d.weights)
        else:
            out[f"{prefix}mu{j}"] = np.asarray(d.mu)
            out[f"{prefix}sigma{j}"] = np.asarray(d.sigma)


# ------------------------------------------------------------------------------------------
def gold_math() -> dict:
    out = {}
    rs = np.random.RandomState(7)
    x = np.concatenate([np.linspace(-7, 7, 2801), rs.normal(0, 2, 400), [0.0, -0.0, 2.0**-30, 0.84375,
                        1.25, 1 / 0.35, 6.0, -6.0, 5.999999, 28.0, np.inf, -np.inf]])
    out["erf_x"] = x
    out["erf_big"] = ref_erf(x)  # polynomial path (size >= 2000)
    out["erf_small"] = ref_erf(x[:500])  # libm path
    t = np.concatenate([np.linspace(-40, 10, 1001), rs.normal(0, 3, 300), [-20.0, 6.0, -19.999, 6.001]])
    out["lndtr_t"] = t
    out["lndtr"] = _truncnorm._log_ndtr(t)
    out["ndtr_big"] = _truncnorm._ndtr(np.linspace(-9, 9, 2401))
    a = rs.uniform(-30, 30, 3000)
    w = np.abs(rs.normal(0, 3, 3000)) + 10 ** rs.uniform(-9, 0, 3000)
    b = a + w
    out["lgm_a"], out["lgm_b"] = a, b
    out["lgm"] = _truncnorm._log_gauss_mass(a, b)
    out["lgm_small"] = _truncnorm._log_gauss_mass(a[:700], b[:700])
    q = np.concatenate([rs.uniform(0, 1, 1490), [0.0, 1.0, 1e-300, 1 - 1e-16, 0.5, 1e-9, 1e-17, 0.999,
                        2.0**-53, 0.25]])
    pa = rs.uniform(-12, 8, 1500)
    pb = pa + np.abs(rs.normal(0, 4, 1500)) + 1e-3
    out["ppf_q"], out["ppf_a"], out["ppf_b"] = q, pa, pb
    out["ppf"] = _truncnorm.ppf(q, pa, pb)
    y = -(10 ** rs.uniform(-12, 3, 800))
    out["ndtri_y"] = y
    out["ndtri"] = _truncnorm._ndtri_exp(y.copy())
    xs = rs.uniform(-4, 4, (50, 1, 3))
    loc = rs.uniform(-2, 2, (40, 3))
    sc = rs.uniform(0.05, 3, (40, 3))
    lo, hi = np.array([-4.0, -5.0, -4.5]), np.array([4.0, 6.0, 4.5])
    out["lpdf_x"], out["lpdf_loc"], out["lpdf_scale"] = xs, loc, sc
    out["lpdf_lo"], out["lpdf_hi"] = lo, hi
    out["lpdf"] = _truncnorm.logpdf(xs, (lo - loc) / sc, (hi - loc) / sc, loc, sc)
    return out


def gold_parzen(space: dict, n_obs: int, seed: int, multivariate: bool, magic_clip: bool,
                endpoints: bool, prior_weight: float, C: int, tag: str) -> dict:
    """One _ParzenEstimator build + sample + log_pdf, straight from reference classes."""
    rng = np.random.RandomState(seed)
    ext = [random_external(space, rng) for _ in range(n_obs)]
    obs = {name: np.asarray([d.to_internal_repr(e[name]) for e in ext], dtype=float)
           for name, d in space.items()}
    pars = _ParzenEstimatorParameters(prior_weight=prior_weight, consider_magic_clip=magic_clip,
                                      consider_endpoints=endpoints, weights=default_weights,
                                      multivariate=multivariate, categorical_distance_func={})
    mpe = _ParzenEstimator(obs, space, pars)
    out = {"space": encode_space(space),
           "obs": np.asarray([obs[k] for k in space]).T.reshape(n_obs, len(space)),
           "flags": np.asarray([multivariate, magic_clip, endpoints, prior_weight, C, seed], dtype=float)}
    mixture_arrays(mpe, "", out)
    srng = np.random.RandomState(seed + 1000)
    smp = mpe.sample(srng, C)
    out["samples"] = np.asarray([smp[k] for k in space]).T
    out["logpdf"] = mpe.log_pdf(smp)
    # second evaluation point set: samples from a different estimator (covers far-from-kernel x)
    other = _ParzenEstimator({k: v[: max(1, n_obs // 7)] for k, v in obs.items()}, space, pars)
    smp2 = other.sample(np.random.RandomState(seed + 2000), C)
    out["samples2"] = np.asarray([smp2[k] for k in space]).T
    out["logpdf2"] = mpe.log_pdf(smp2)
    return {f"{tag}/{k}": v for k, v in out.items()}


def make_trials(space: dict, n: int, seed: int, n_pruned: int = 0, n_missing: int = 0,
                directions=("minimize",)):
    rng = np.random.RandomState(seed)
    trials = []
    names = list(space)
    for i in range(n):
        params = random_external(space, rng)
        dists = dict(space)
        if i < n_missing:  # drop one param -> conditional search space
            drop = names[rng.randint(len(names))]
            params.pop(drop)
            dists = {k: v for k, v in dists.items() if k != drop}
        if len(directions) == 1:
            val = float(rng.normal())
            if i >= n - n_pruned:
                iv = {int(s): float(rng.normal()) for s in range(rng.randint(0, 4))}
                t = create_trial(state=TrialState.PRUNED, params=params, distributions=dists,
                                 intermediate_values=iv, value=None)
            else:
                t = create_trial(value=val, params=params, distributions=dists)
        else:
            t = create_trial(values=[float(v) for v in rng.normal(size=len(directions))],
                             params=params, distributions=dists)
        t.number = i
        t._trial_id = i
        trials.append(t)
    return trials


def history_arrays(trials, study) -> tuple[np.ndarray, np.ndarray]:
    """category / key arrays as the C-ABI takes them (single objective)."""
    from optuna.samplers._tpe.sampler import _get_pruned_trial_score
    from optuna.study import StudyDirection
    cat = np.zeros(len(trials), dtype=np.int8)
    key = np.zeros((len(trials), 2))
    sign = -1.0 if study.direction == StudyDirection.MAXIMIZE else 1.0
    for i, t in enumerate(trials):
        if t.state == TrialState.COMPLETE:
            cat[i] = 0
            key[i, 0] = sign * t.value
        elif t.state == TrialState.PRUNED:
            cat[i] = 1
            key[i] = _get_pruned_trial_score(t, study)
        elif t.state == TrialState.RUNNING:
            cat[i] = 3
    return cat, key


def gold_suggest(space: dict, n: int, seed: int, multivariate: bool, C: int, direction: str,
                 n_pruned: int, n_missing: int, tag: str, gamma=None) -> dict:
    """Full TPESampler._sample on an injected history: every stage captured from the reference."""
    from unittest.mock import patch
    kw = {} if gamma is None else {"gamma": gamma}
    sampler = TPESampler(seed=seed, n_ei_candidates=C, multivariate=multivariate, n_startup_trials=0, **kw)
    study = optuna.create_study(direction=direction, sampler=sampler)
    trials = make_trials(space, n, seed + 1, n_pruned=n_pruned, n_missing=n_missing)
    out = {"space": encode_space(space), "X": internal_matrix(trials, space)}
    cat, key = history_arrays(trials, study)
    out["category"], out["key"] = cat, key
    n_fin = len(trials)
    n_below = sampler._gamma(n_fin)
    below, above = _split_trials(study, trials, n_below, False)
    out["below"] = np.asarray([t.number for t in below], dtype=np.int64)
    out["above"] = np.asarray([t.number for t in above], dtype=np.int64)
    out["cfg"] = np.asarray([multivariate, C, seed, n_below], dtype=float)
    captured = {}
    orig_acq = TPESampler._compute_acquisition_func

    def spy(self, samples, mpe_below, mpe_above):
        captured["samples"] = {k: v.copy() for k, v in samples.items()}
        captured["mb"], captured["ma"] = mpe_below, mpe_above
        captured["ll"] = mpe_below.log_pdf(samples)
        captured["lg"] = mpe_above.log_pdf(samples)
        return orig_acq(self, samples, mpe_below, mpe_above)

    frozen = create_trial(state=TrialState.RUNNING, params={}, distributions={})
    frozen.number = n
    frozen._trial_id = n
    with patch.object(study._storage, "get_all_trials", return_value=trials), \
            patch.object(TPESampler, "_compute_acquisition_func", spy):
        if multivariate:
            res = sampler._sample(study, frozen, dict(space))
            results = [res]
            calls = [list(space)]
        else:
            results, calls = [], []
            for name, d in space.items():
                results.append(sampler._sample(study, frozen, {name: d}))
                calls.append([name])
                # capture per call
                j = list(space).index(name)
                mixture_arrays(captured["mb"], f"u{j}/b_", out)
                mixture_arrays(captured["ma"], f"u{j}/a_", out)
                out[f"u{j}/samples"] = np.asarray(captured["samples"][name])[:, None]
                out[f"u{j}/ll"], out[f"u{j}/lg"] = captured["ll"], captured["lg"]
    if multivariate:
        mixture_arrays(captured["mb"], "b_", out)
        mixture_arrays(captured["ma"], "a_", out)
        out["samples"] = np.asarray([captured["samples"][k] for k in space]).T
        out["ll"], out["lg"] = captured["ll"], captured["lg"]
    # returned external values -> internal repr for storage in npz
    ret = np.full(len(space), np.nan)
    for res in results:
        for name, v in res.items():
            ret[list(space).index(name)] = space[name].to_internal_repr(v)
    out["ret_internal"] = ret
    return {f"{tag}/{k}": v for k, v in out.items()}


def gold_branin() -> dict:
    def branin(t):
        x = t.suggest_float("x", -5, 10)
        y = t.suggest_float("y", 0, 15)
        return ((y - 5.1 / (4 * math.pi**2) * x * x + 5 / math.pi * x - 6) ** 2
                + 10 * (1 - 1 / (8 * math.pi)) * math.cos(x) + 10)
    out = {}
    for mv in (False, True):
        s = optuna.create_study(sampler=TPESampler(seed=0, multivariate=mv))
        s.optimize(branin, n_trials=200)
        tag = "mv" if mv else "uni"
        out[f"branin_{tag}/xy"] = np.asarray([[t.params["x"], t.params["y"]] for t in s.trials])
        out[f"branin_{tag}/values"] = np.asarray([t.value for t in s.trials])
    return out


def gold_mo() -> dict:
    """Multi-objective pieces straight from the reference: ranks, hypervolumes, HSSP, the MOTPE
    split, the hypervolume weights, and full MOTPE _sample traces."""
    from unittest.mock import patch
    from optuna._hypervolume import compute_hypervolume
    from optuna._hypervolume.hssp import _solve_hssp
    from optuna.samplers._tpe.sampler import (_calculate_weights_below_for_multi_objective,
                                              _get_reference_point, _split_complete_trials_multi_objective)
    from optuna.study._multi_objective import _fast_non_domination_rank
    out = {}
    rs = np.random.RandomState(21)
    # hypervolume / rank / hssp on raw point sets
    ci = 0
    for m in (2, 3, 4, 5):
        for n in (1, 2, 3, 7, 40):
            v = rs.uniform(0, 1, (n, m))
            if n >= 7:
                v[1] = v[0]  # duplicate
                v = np.round(v, 1) if (n == 40 and m == 3) else v  # many ties
            ref = _get_reference_point(v)
            out[f"hv{ci}/v"], out[f"hv{ci}/ref"] = v, ref
            out[f"hv{ci}/hv"] = np.asarray(compute_hypervolume(v, ref))
            out[f"hv{ci}/rank"] = _fast_non_domination_rank(v)
            nb = max(1, n // 3)
            out[f"hv{ci}/rank_nb"] = _fast_non_domination_rank(v, n_below=nb)
            k = max(1, n // 2)
            out[f"hv{ci}/hssp"] = _solve_hssp(v, np.arange(n) * 3, k, ref)
            ci += 1
    out["hv_n"] = np.asarray(ci)

    class _S:  # the split / weights helpers only read directions
        def __init__(self, m):
            self.directions = [optuna.study.StudyDirection.MINIMIZE] * m
    ci = 0
    for m, n, nb in ((2, 60, 7), (3, 200, 25), (4, 300, 25), (4, 300, 60), (2, 50, 20), (3, 40, 39), (4, 30, 0)):
        v = rs.normal(size=(n, m))
        if ci == 4:
            v = np.round(v, 0)  # heavy duplication
        if ci == 1:
            v[5, 0] = np.inf
            v[9, 1] = -np.inf
        trials = [create_trial(values=list(map(float, v[i])), params={}, distributions={}) for i in range(n)]
        for i, t in enumerate(trials):
            t.number = i
        below, above = _split_complete_trials_multi_objective(trials, _S(m), nb)
        bidx = np.asarray([t.number for t in below], dtype=np.int64)
        out[f"mo{ci}/v"], out[f"mo{ci}/nb"], out[f"mo{ci}/below"] = v, np.asarray(nb), bidx
        out[f"mo{ci}/w"] = _calculate_weights_below_for_multi_objective(_S(m), below, None)
        ci += 1
    out["mo_n"] = np.asarray(ci)

    # full MOTPE suggestions (multivariate and univariate)
    ci = 0
    for m, n, mv, C, seed, gamma in ((4, 400, True, 24, 31, None), (2, 300, True, 32, 32, None),
                                     (3, 300, False, 24, 33, None),
                                     (4, 500, True, 24, 34, lambda x: math.ceil(0.1 * x))):
        space = {f"x{j}": FloatDistribution(0, 1) for j in range(8 if mv else 3)}
        kw = {} if gamma is None else {"gamma": gamma}
        sampler = TPESampler(seed=seed, n_ei_candidates=C, multivariate=mv, n_startup_trials=0, **kw)
        study = optuna.create_study(directions=["minimize"] * m, sampler=sampler)
        r2 = np.random.RandomState(seed + 1)
        X = r2.uniform(0, 1, (n, len(space)))
        cs = np.linspace(0.2, 0.8, m)
        vals = np.stack([((X - c) ** 2).sum(1) for c in cs], 1)
        trials = []
        for i in range(n):
            t = create_trial(values=list(map(float, vals[i])), params=dict(zip(space, X[i].tolist())),
                             distributions=space)
            t.number = i
            t._trial_id = i
            trials.append(t)
        captured = {}
        orig_acq = TPESampler._compute_acquisition_func

        def spy(self, samples, mpe_below, mpe_above):
            captured.setdefault("calls", []).append(
                ({k: v.copy() for k, v in samples.items()}, mpe_below.log_pdf(samples), mpe_above.log_pdf(samples),
                 np.asarray(mpe_below._mixture_distribution.weights)))
            return orig_acq(self, samples, mpe_below, mpe_above)

        frozen = create_trial(state=TrialState.RUNNING, params={}, distributions={})
        frozen.number = n
        frozen._trial_id = n
        t = f"mosg{ci}/"
        out[t + "X"], out[t + "values"] = X, vals
        out[t + "cfg"] = np.asarray([mv, C, seed, sampler._gamma(n), m], dtype=float)
        with patch.object(study._storage, "get_all_trials", return_value=trials), \
                patch.object(TPESampler, "_compute_acquisition_func", spy):
            if mv:
                res = sampler._sample(study, frozen, dict(space))
                out[t + "ret"] = np.asarray([res[k] for k in space])
            else:
                ret = []
                for name, d in space.items():
                    ret.append(sampler._sample(study, frozen, {name: d})[name])
                out[t + "ret"] = np.asarray(ret)
        for q, (smp, ll, lg, wb) in enumerate(captured["calls"]):
            out[f"{t}c{q}/samples"] = np.asarray([smp[k] for k in smp]).T
            out[f"{t}c{q}/ll"], out[f"{t}c{q}/lg"], out[f"{t}c{q}/wb"] = ll, lg, wb
        out[t + "ncalls"] = np.asarray(len(captured["calls"]))
        ci += 1
    out["mosg_n"] = np.asarray(ci)
    return out


def main() -> None:
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "motpe.npz"), **gold_mo())
    np.savez_compressed(os.path.join(OUT, "math.npz"), **gold_math())

    pz = {}
    i = 0
    for mv in (False, True):
        for clip in (True, False):
            for endp in (False, True):
                for n_obs, pw in ((0, 1.0), (1, 1.0), (7, 1.0), (40, 0.5), (300, 1.0)):
                    pz.update(gold_parzen(MIXED_SPACE, n_obs, 11 + i, mv, clip, endp, pw, 24, f"pz{i}"))
                    i += 1
    pz["n_cases"] = np.asarray(i)
    np.savez_compressed(os.path.join(OUT, "parzen.npz"), **pz)

    sg = {}
    cases = [
        (MIXED_SPACE, 60, 3, True, 24, "minimize", 0, 0),
        (MIXED_SPACE, 60, 4, False, 24, "minimize", 0, 0),
        (MIXED_SPACE, 400, 5, True, 64, "maximize", 30, 0),
        (MIXED_SPACE, 400, 6, False, 32, "maximize", 30, 0),
        (MIXED_SPACE, 250, 7, True, 24, "minimize", 10, 40),
        ({f"x{j:02d}": FloatDistribution(0, 1) for j in range(32)}, 3000, 8, True, 128, "minimize", 0, 0),
        ({f"x{j:02d}": FloatDistribution(0, 1) for j in range(4)}, 3000, 9, False, 96, "minimize", 0, 0),
        ({f"x{j:02d}": FloatDistribution(-2, 3) for j in range(6)}, 2000, 10, True, 48, "minimize", 0, 0),
    ]
    for ci, (space, n, seed, mv, C, direction, npr, nmiss) in enumerate(cases):
        sg.update(gold_suggest(space, n, seed, mv, C, direction, npr, nmiss, f"sg{ci}"))
    # custom gamma -> large below set
    sg.update(gold_suggest({f"x{j:02d}": FloatDistribution(0, 1) for j in range(8)}, 2000, 12, True, 32,
                           "minimize", 0, 0, f"sg{len(cases)}", gamma=lambda n: math.ceil(0.1 * n)))
    sg["n_cases"] = np.asarray(len(cases) + 1)
    np.savez_compressed(os.path.join(OUT, "suggest.npz"), **sg)

    np.savez_compressed(os.path.join(OUT, "branin.npz"), **gold_branin())
    print("golden vectors written to", os.path.abspath(OUT))


if __name__ == "__main__":
    sys.exit(main())
