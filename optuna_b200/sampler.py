"""B200TPESampler -- drop-in for optuna.samplers.TPESampler backed by libtpe_b200.so.

Mirrors the reference's plugin surface (optuna/samplers/_base.py:31-228) and constructor
(optuna/samplers/_tpe/sampler.py:305-385).  What stays on the host is exactly what the reference
can only do in Python: walking FrozenTrial objects once (incrementally) into arrays, evaluating the
user's callables (gamma, weights, constraints_func, categorical_distance_func), drawing the
uniforms from the sampler's own numpy RandomState in the reference's order
(probability_distributions.py:87,100,138-144) and converting the winner back with
``to_external_repr``.  Split, estimator build, candidate sampling, the log-density grid and the
argmax run on the GPU through the C ABI (include/optuna_b200_tpe.h).  There is no CPU fallback.
"""
from __future__ import annotations

import json
import math
import threading
import warnings
from typing import Any, Callable, Sequence

import numpy as np

from . import _lib
from ._compat import (CONSTRAINTS_KEY, RELATIVE_PARAMS_KEY, SYSTEM_ATTR_MAX_LENGTH, BaseDistribution,
                      BaseSampler, CategoricalDistribution, FloatDistribution, IntDistribution,
                      LazyRandomState, StudyDirection, TrialState, random_independent)
from .engine import ParamSpec, TPEEngine

EPS = 1e-12


def default_gamma(x: int) -> int:
    """sampler.py:53-54"""
    return min(math.ceil(0.1 * x), 25)


def hyperopt_default_gamma(x: int) -> int:
    """sampler.py:57-58"""
    return min(math.ceil(0.25 * x**0.5), 25)


def default_weights(x: int) -> np.ndarray:
    """sampler.py:61-69.  When this very function is the sampler's ``weights`` the library
    evaluates it on the device (k_weights); the host version serves custom compositions."""
    if x == 0:
        return np.asarray([])
    if x < 25:
        return np.ones(x)
    return np.concatenate([np.linspace(1.0 / x, 1.0, num=x - 25), np.ones(25)], axis=0)


def _checked_weights(func: Callable[[int], np.ndarray], n: int) -> np.ndarray:
    """parzen_estimator.py:88-109"""
    w = np.array(func(n))[:n]
    if np.any(w < 0):
        raise ValueError(f"The `weights` function is not allowed to return negative values {w}. "
                         f"The argument of the `weights` function is {n}.")
    if len(w) > 0 and np.sum(w) <= 0:
        raise ValueError(f"The `weight` function is not allowed to return all-zero values {w}."
                         f" The argument of the `weights` function is {n}.")
    if not np.all(np.isfinite(w)):
        raise ValueError("The `weights`function is not allowed to return infinite or NaN values "
                         f"{w}. The argument of the `weights` function is {n}.")
    return np.asarray(w, dtype=np.float64)


def _pruned_key(trial, sign: float) -> tuple[float, float]:
    """sampler.py:782-792"""
    if len(trial.intermediate_values) > 0:
        step, v = max(trial.intermediate_values.items())
        if math.isnan(v):
            return -step, float("inf")
        return -step, sign * v
    return 1, 0.0


def _infeasible_score(trial) -> float:
    """sampler.py:803-813"""
    con = trial.system_attrs.get(CONSTRAINTS_KEY)
    if con is None:
        warnings.warn(f"Trial {trial.number} does not have constraint values."
                      " It will be treated as a lower priority than other trials.")
        return float("inf")
    return sum(v for v in con if v > 0)


def _spec_of(name: str, d: BaseDistribution, dist_funcs: dict) -> ParamSpec:
    if isinstance(d, CategoricalDistribution):
        table = None
        if name in dist_funcs:
            f = dist_funcs[name]
            table = np.asarray([[f(a, b) for b in d.choices] for a in d.choices], dtype=np.float64)
        return ParamSpec(kind=_lib.KIND_CAT, n_choices=len(d.choices), dist_table=table)
    if isinstance(d, IntDistribution):
        return ParamSpec(kind=_lib.KIND_INT, low=float(d.low), high=float(d.high), step=float(d.step), log=bool(d.log))
    assert isinstance(d, FloatDistribution), d
    return ParamSpec(kind=_lib.KIND_FLOAT, low=d.low, high=d.high, step=d.step, log=bool(d.log))


class _History:
    """Device-resident trial history kept in step with the study (replaces the per-call
    FrozenTrial walks of sampler.py:511-521 and :686-722)."""

    def __init__(self) -> None:
        self.columns: dict[str, int] = {}
        self.dists: list[BaseDistribution] = []
        self.n = 0
        self.last_number = -1
        self.sign = 1.0
        self.token: tuple | None = None
        self.n_finished = 0
        # constant_liar: position -> trial number of the rows uploaded while the trial was RUNNING
        self.running: dict[int, int] = {}
        # incremental intersection search space (optuna/search_space/intersection.py:14-55)
        self.inter: dict[str, BaseDistribution] | None = None
        self.inter_n = 0
        self.inter_last = -1
        # incremental group decomposition (optuna/search_space/group_decomposed.py:14-68)
        self.groups: list[dict[str, BaseDistribution]] = []
        self.groups_n = 0
        self.groups_last = -1


class _DeviceSyncedRng:
    """The sampler's ``LazyRandomState`` whose MT19937 state may temporarily be newer on the device.

    Large asks draw their uniforms on the GPU; instead of copying the generator state back into the host
    ``RandomState`` after every ask (get_state + set_state cost 80 us), the state stays on the device
    and the next ask continues from it.  Any access to ``.rng`` -- an ask drawn on the host, reseeding,
    pickling, somebody reading ``sampler._rng.rng`` -- first brings the host generator up to date."""

    def __init__(self, inner) -> None:
        self._inner = inner
        self._engine = None  # the engine holding a newer state, if any

    @property
    def rng(self) -> np.random.RandomState:
        eng, self._engine = self._engine, None
        if eng is not None:
            eng.finish_rng(self._inner.rng)
        return self._inner.rng

    def on_device(self, eng) -> bool:
        return self._engine is not None and self._engine is eng

    def mark_device(self, eng) -> None:
        self._engine = eng

    def __getstate__(self) -> dict:
        self.rng  # flush
        return {"_inner": self._inner, "_engine": None}

    def __setstate__(self, state: dict) -> None:
        self.__dict__.update(state)


class B200TPESampler(BaseSampler):
    def __init__(
        self,
        *,
        consider_prior: bool = True,
        prior_weight: float = 1.0,
        consider_magic_clip: bool = True,
        consider_endpoints: bool = False,
        n_startup_trials: int = 10,
        n_ei_candidates: int = 24,
        gamma: Callable[[int], int] = default_gamma,
        weights: Callable[[int], np.ndarray] = default_weights,
        seed: int | None = None,
        multivariate: bool = False,
        group: bool = False,
        warn_independent_sampling: bool = True,
        constant_liar: bool = False,
        constraints_func: Callable[[Any], Sequence[float]] | None = None,
        categorical_distance_func: dict[str, Callable[[Any, Any], float]] | None = None,
        device: int = 0,
    ) -> None:
        if not consider_prior:
            warnings.warn("`consider_prior` is deprecated; it falls back to `True`.", FutureWarning)
        if group and not multivariate:
            raise ValueError("``group`` option can only be enabled when ``multivariate`` is enabled.")
        self._prior_weight = prior_weight
        self._magic_clip = consider_magic_clip
        self._endpoints = consider_endpoints
        self._n_startup_trials = n_startup_trials
        self._n_ei_candidates = n_ei_candidates
        self._gamma = gamma
        self._weights = weights
        self._multivariate = multivariate
        self._group = group
        self._warn_independent_sampling = warn_independent_sampling
        self._constant_liar = constant_liar
        self._constraints_func = constraints_func
        self._cat_dist_funcs = categorical_distance_func or {}
        self._rng = _DeviceSyncedRng(LazyRandomState(seed))
        self._startup_rng = LazyRandomState(seed)  # the embedded RandomSampler's own state (sampler.py:348-349)
        self._device = device
        self._engine: TPEEngine | None = None
        self._hist = _History()
        self._lock = threading.RLock()

    # -- pickling: device state is a cache re-creatable from the study ----------
    def __getstate__(self) -> dict:
        state = self.__dict__.copy()
        state["_engine"] = None
        state["_hist"] = _History()
        del state["_lock"]
        return state

    def __setstate__(self, state: dict) -> None:
        self.__dict__.update(state)
        self._lock = threading.RLock()

    def close(self) -> None:
        """Release the device context (re-created on demand)."""
        eng, self._engine = getattr(self, "_engine", None), None
        if eng is not None:
            rng = getattr(self, "_rng", None)
            if rng is not None:
                rng.rng  # bring the host generator up to date before the device state goes away
            eng.close()
        self._hist = _History()

    def __del__(self) -> None:  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def hyperopt_parameters() -> dict[str, Any]:
        return {"consider_prior": True, "prior_weight": 1.0, "consider_magic_clip": True,
                "consider_endpoints": False, "n_startup_trials": 20, "n_ei_candidates": 24,
                "gamma": hyperopt_default_gamma, "weights": default_weights}

    def reseed_rng(self) -> None:
        self._rng.rng.seed()
        self._startup_rng.rng.seed()

    # -- plugin surface -------------------------------------------------------------------------------
    def infer_relative_search_space(self, study, trial) -> dict[str, BaseDistribution]:
        if not self._multivariate:
            return {}
        if self._group:
            with self._lock:
                groups = self._group_spaces(study)
            out: dict[str, BaseDistribution] = {}
            for sub in groups:
                for name, d in sorted(sub.items()):
                    if not d.single():
                        out[name] = d
            return out
        with self._lock:
            space = self._intersection(study)
        return {k: d for k, d in space.items() if not d.single()}

    def sample_relative(self, study, trial, search_space: dict[str, BaseDistribution]) -> dict[str, Any]:
        if self._group:
            # one joint suggestion per group of parameters that always appear together (sampler.py:417-431)
            with self._lock:
                groups = [dict(g) for g in self._hist.groups]
            params: dict[str, Any] = {}
            for sub in groups:
                part = {name: d for name, d in sorted(sub.items()) if not d.single() and name in search_space}
                params.update(self._sample_relative(study, trial, part))
        else:
            params = self._sample_relative(study, trial, search_space)
        if params != {} and self._constant_liar:
            text = json.dumps(params)
            for i in range(0, len(text), SYSTEM_ATTR_MAX_LENGTH):
                study._storage.set_trial_system_attr(trial._trial_id,
                                                     f"{RELATIVE_PARAMS_KEY}:{i // SYSTEM_ATTR_MAX_LENGTH}",
                                                     text[i: i + SYSTEM_ATTR_MAX_LENGTH])
        return params

    def _sample_relative(self, study, trial, search_space) -> dict[str, Any]:
        if search_space == {}:
            return {}
        trials = study._get_trials(deepcopy=False, states=(TrialState.COMPLETE, TrialState.PRUNED), use_cache=True)
        if len(trials) < self._n_startup_trials:
            return {}
        return self._sample(study, trial, search_space)

    def sample_independent(self, study, trial, param_name: str, param_distribution: BaseDistribution) -> Any:
        trials = study._get_trials(deepcopy=False, states=(TrialState.COMPLETE, TrialState.PRUNED), use_cache=True)
        if len(trials) < self._n_startup_trials:
            return random_independent(self._startup_rng.rng, param_distribution)
        if self._warn_independent_sampling and self._multivariate:
            if any(param_name in t.params for t in trials):
                warnings.warn(f"The parameter '{param_name}' in trial#{trial.number} is sampled independently "
                              "instead of being sampled by multivariate TPE sampler (dynamic search space is "
                              "not supported for `multivariate=True`).")
        return self._sample(study, trial, {param_name: param_distribution})[param_name]

    def before_trial(self, study, trial) -> None:
        pass

    def after_trial(self, study, trial, state, values) -> None:
        assert state in (TrialState.COMPLETE, TrialState.FAIL, TrialState.PRUNED)
        if self._constraints_func is None or state not in (TrialState.COMPLETE, TrialState.PRUNED):
            return
        con = None
        try:
            out = self._constraints_func(trial)
            if not isinstance(out, (tuple, list)):
                warnings.warn("Constraints should be a sequence of floats.")
            con = tuple(out)
        finally:
            study._storage.set_trial_system_attr(trial._trial_id, CONSTRAINTS_KEY, con)

    # -- host glue -------------------------------------------------------------------------------------
    def _eng(self) -> TPEEngine:
        if self._engine is None:
            self._engine = TPEEngine(self._device)
        return self._engine

    def _get_params(self, trial) -> dict[str, Any]:
        """sampler.py:493-509"""
        if trial.state.is_finished() or not self._multivariate:
            return trial.params
        chunks = []
        i = 0
        while (c := trial.system_attrs.get(f"{RELATIVE_PARAMS_KEY}:{i}")):
            chunks.append(c)
            i += 1
        if not chunks:
            return trial.params
        params = json.loads("".join(chunks))
        params.update(trial.params)
        return params

    def _intersection(self, study) -> dict[str, BaseDistribution]:
        h = self._hist
        trials = study._get_trials(deepcopy=False, states=(TrialState.COMPLETE, TrialState.PRUNED), use_cache=True)
        fresh = not (h.inter_n <= len(trials) and (h.inter_n == 0 or trials[h.inter_n - 1].number == h.inter_last))
        if fresh:
            h.inter, h.inter_n = None, 0
        for t in trials[h.inter_n:]:
            if h.inter is None:
                h.inter = dict(t.distributions)
            else:
                h.inter = {k: d for k, d in h.inter.items() if k in t.distributions and d == t.distributions[k]}
        h.inter_n = len(trials)
        h.inter_last = trials[-1].number if trials else -1
        return dict(sorted((h.inter or {}).items(), key=lambda kv: kv[0]))

    def _group_spaces(self, study) -> list[dict[str, BaseDistribution]]:
        h = self._hist
        trials = study._get_trials(deepcopy=False, states=(TrialState.COMPLETE, TrialState.PRUNED), use_cache=True)
        if not (h.groups_n <= len(trials) and (h.groups_n == 0 or trials[h.groups_n - 1].number == h.groups_last)):
            h.groups, h.groups_n = [], 0
        for t in trials[h.groups_n:]:
            left = set(t.distributions)
            nxt: list[dict[str, BaseDistribution]] = []
            for sub in h.groups:
                keys = set(sub)
                nxt.append({name: sub[name] for name in keys & left})
                nxt.append({name: sub[name] for name in keys - left})
                left -= keys
            nxt.append({name: t.distributions[name] for name in left})
            h.groups = [g for g in nxt if g]
        h.groups_n = len(trials)
        h.groups_last = trials[-1].number if trials else -1
        return [dict(g) for g in h.groups]

    def sample_relative_batch(self, study, search_space: dict[str, BaseDistribution], n_asks: int) -> list[dict]:
        """`n_asks` joint suggestions against the current (frozen) history in ONE device call.

        Equivalent to calling ``sample_relative`` n_asks times without a ``tell`` in between
        (the reference idiom is a Python loop of ``study.ask()``; with
        ``constant_liar=False`` every such ask sees the same split and the same two mixtures and
        only the RNG position differs).  The uniforms are drawn ask by ask from the sampler's own
        RandomState, so the results are those of the sequential loop."""
        if self._group or self._constant_liar:
            raise ValueError("sample_relative_batch needs group=False and constant_liar=False "
                             "(constant-liar asks depend on each other)")
        if search_space == {} or n_asks <= 0:
            return [{} for _ in range(max(n_asks, 0))]
        trials = study._get_trials(deepcopy=False, states=(TrialState.COMPLETE, TrialState.PRUNED), use_cache=True)
        if len(trials) < self._n_startup_trials:
            return [{} for _ in range(n_asks)]
        with self._lock:
            n_finished, cols = self._sync(study, None, search_space)
            cfg = dict(n_below=int(self._gamma(n_finished)), n_candidates=self._n_ei_candidates,
                       multivariate=self._multivariate, prior_weight=self._prior_weight,
                       magic_clip=self._magic_clip, endpoints=self._endpoints)
            eng = self._eng()
            multi = study._is_multi_objective()
            _, nb, na = eng.prepare(cols, **cfg)
            if self._weights is default_weights:
                build = eng.build
            else:
                wb = None if multi else _checked_weights(self._weights, nb)
                wa = _checked_weights(self._weights, na)
                build = lambda: eng.build(wb, wa)  # noqa: E731
            # ask-by-ask draws are consecutive stretches of one stream: one generation yields the same numbers
            x = self._sample_and_select(eng, search_space, n_asks, build)
        # column-wise conversion (FloatDistribution.to_external_repr is the identity): 8192 x 32 values in
        # a few ms instead of one Python call per value
        names = list(search_space)
        columns = []
        for j, name in enumerate(names):
            d = search_space[name]
            col = x[:, j].tolist()
            columns.append(col if isinstance(d, FloatDistribution) else [d.to_external_repr(v) for v in col])
        return [dict(zip(names, row)) for row in zip(*columns)]

    def _rows(self, study, trials, names: list[str], dists: list[BaseDistribution]):
        sign = -1.0 if (not study._is_multi_objective() and study.direction == StudyDirection.MAXIMIZE) else 1.0
        n, p = len(trials), len(names)
        X = np.full((n, p), np.nan)
        cat = np.zeros(n, dtype=np.int8)
        key = np.zeros((n, 2))
        multi = study._is_multi_objective()
        signs = np.asarray([-1.0 if d == StudyDirection.MAXIMIZE else 1.0 for d in study.directions])
        vals = np.full((n, len(signs)), np.inf) if multi else None
        for i, t in enumerate(trials):
            params = self._get_params(t)
            for j, name in enumerate(names):
                if name in params:
                    X[i, j] = dists[j].to_internal_repr(params[name])
            if t.state == TrialState.RUNNING:
                cat[i] = _lib.CAT_RUNNING
            elif self._constraints_func is not None and (score := _infeasible_score(t)) > 0:
                cat[i] = _lib.CAT_INFEASIBLE
                key[i, 0] = score
            elif t.state == TrialState.COMPLETE:
                cat[i] = _lib.CAT_COMPLETE
                key[i, 0] = sign * t.value if not multi else 0.0
            else:
                cat[i] = _lib.CAT_PRUNED
                key[i] = _pruned_key(t, sign) if not multi else (1, 0.0)
            if multi and t.values is not None:
                vals[i] = signs * np.asarray(t.values, dtype=float)
        return X, cat, key, vals

    def _sync(self, study, trial, search_space: dict[str, BaseDistribution]) -> tuple[int, list[int]]:
        """Bring the device history up to date; returns (#finished trials, device columns).

        Rows are kept in trial-number order.  Finished trials are appended as they appear.  With
        ``constant_liar`` the RUNNING trials (other than the one being sampled) are rows too
        (sampler.py:526-535); a RUNNING row is re-uploaded in place at every ask (its parameters may still
        grow) until the trial has finished, so an ask costs O(#running), not O(#trials)."""
        h = self._hist
        eng = self._eng()
        if self._constant_liar:
            states = (TrialState.COMPLETE, TrialState.PRUNED, TrialState.RUNNING)
            trials = study._get_trials(deepcopy=False, states=states, use_cache=False)
            for i in range(len(trials) - 1, max(len(trials) - 4097, -1), -1):  # the current trial is recent
                if trials[i].number == trial.number:
                    trials = trials[:i] + trials[i + 1:]
                    break
            else:
                trials = [t for t in trials if t.number != trial.number]
        else:
            trials = study._get_trials(deepcopy=False, states=(TrialState.COMPLETE, TrialState.PRUNED), use_cache=True)
        token = (getattr(study, "_study_id", None), id(getattr(study, "_storage", None)),
                 tuple(getattr(study, "directions", ())))
        rebuild = h.token != token
        for name, d in search_space.items():
            j = h.columns.get(name)
            if j is None or h.dists[j] != d:
                rebuild = True
        if not rebuild:
            if not (h.n <= len(trials) and (h.n == 0 or trials[h.n - 1].number == h.last_number)):
                rebuild = True
            elif any(trials[pos].number != num for pos, num in h.running.items()):
                rebuild = True  # a RUNNING trial disappeared (FAIL): the positions have shifted
        if rebuild:
            names = list(h.columns) if h.token == token else []
            dists = list(h.dists) if h.token == token else []
            for name, d in search_space.items():
                if name in names:
                    dists[names.index(name)] = d
                else:
                    names.append(name)
                    dists.append(d)
            h.columns = {name: j for j, name in enumerate(names)}
            h.dists = dists
            h.token = token
            eng.set_space([_spec_of(nm, d, self._cat_dist_funcs) for nm, d in zip(names, dists)])
            X, cat, key, vals = self._rows(study, trials, names, dists)
            eng.set_history(X, cat, key)
            if vals is not None:
                eng.set_values(vals, 0)
            h.running = {i: t.number for i, t in enumerate(trials) if t.state == TrialState.RUNNING}
            h.n_finished = len(trials) - len(h.running)
        else:
            names = list(h.columns)
            for pos in sorted(h.running):  # refresh the rows of trials that were RUNNING at the last ask
                t = trials[pos]
                X, cat, key, vals = self._rows(study, [t], names, h.dists)
                eng.update_history(X, cat, key, pos)
                if vals is not None:
                    eng.set_values(vals, pos)
                if t.state != TrialState.RUNNING:
                    del h.running[pos]
                    h.n_finished += 1
            if len(trials) > h.n:
                fresh = trials[h.n:]
                X, cat, key, vals = self._rows(study, fresh, names, h.dists)
                eng.append_history(X, cat, key)
                if vals is not None:
                    eng.set_values(vals, h.n)
                for i, t in enumerate(fresh):
                    if t.state == TrialState.RUNNING:
                        h.running[h.n + i] = t.number
                    else:
                        h.n_finished += 1
        h.n = len(trials)
        h.last_number = trials[-1].number if trials else -1
        return h.n_finished, [h.columns[name] for name in search_space]

    #: asks needing at least this many uniforms have them generated on the device
    DEVICE_RNG_MIN = 4096

    def _draw_uniforms(self, search_space: dict[str, BaseDistribution]) -> np.ndarray:
        """The uniforms one reference `_sample` consumes, in its order: C for `rng.choice`, C per
        categorical column, then an (n_numeric, C) block (probability_distributions.py:87,100,138-144).
        `rand`, `choice` and `uniform(0, 1)` all take consecutive `random_sample` outputs unchanged, so
        ONE call yields the identical stream (checked in tests/test_host_glue.py) at half the cost."""
        return self._rng.rng.random_sample(self._n_ei_candidates * (1 + len(search_space)))

    def _sample_and_select(self, eng: TPEEngine, search_space: dict[str, BaseDistribution], n_asks: int,
                           build) -> np.ndarray:
        """Uniforms + stages 3-4.  Large asks: the library generates the generator's next outputs on the
        GPU (k_mt19937_uniform, the same MT19937 stream bit for bit) while `build()` queues the
        estimator builds, and the host generator is then moved to the state after the draws; small asks
        draw on the host."""
        n = n_asks * self._n_ei_candidates * (1 + len(search_space))
        if n >= self.DEVICE_RNG_MIN:
            if self._rng.on_device(eng):
                eng.stage_rng(None, n)            # continue from the state the previous ask ended in
            else:
                eng.stage_rng(self._rng.rng, n)
            build()
            x, _, _ = eng.sample_and_select(None, n_asks)
            self._rng.mark_device(eng)            # the host generator is brought up to date on demand
        else:
            build()
            x, _, _ = eng.sample_and_select(self._rng.rng.random_sample(n), n_asks)
        return x

    def _sample(self, study, trial, search_space: dict[str, BaseDistribution]) -> dict[str, Any]:
        """TPESampler._sample (sampler.py:523-560)."""
        with self._lock:
            n_finished, cols = self._sync(study, trial, search_space)
            n_below = self._gamma(n_finished)
            cfg = dict(n_below=int(n_below), n_candidates=self._n_ei_candidates, multivariate=self._multivariate,
                       prior_weight=self._prior_weight, magic_clip=self._magic_clip, endpoints=self._endpoints)
            if self._prior_weight < 0:
                raise ValueError("A non-negative value must be specified for prior_weight,"
                                 f" but got {self._prior_weight}.")
            eng = self._eng()
            if self._weights is default_weights:
                eng.prepare(cols, **cfg)
                x = self._sample_and_select(eng, search_space, 1, eng.build)
            else:
                _, nb, na = eng.prepare(cols, **cfg)
                # multi-objective studies weight l(x) by hypervolume contributions (computed by the
                # library); the user's weights function then only shapes g(x) (sampler.py:570-584)
                wb = None if study._is_multi_objective() else _checked_weights(self._weights, nb)
                wa = _checked_weights(self._weights, na)
                x = self._sample_and_select(eng, search_space, 1, lambda: eng.build(wb, wa))
        out = {}
        for j, (name, d) in enumerate(search_space.items()):
            out[name] = d.to_external_repr(float(x[0, j]))
        return out
