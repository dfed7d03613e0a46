"""Minimal host-side mirror of the reference objects that sit either side of the sampler plugin.

``B200TPESampler`` is a drop-in for ``optuna.samplers.TPESampler``: when ``optuna`` is importable
it subclasses optuna's ``BaseSampler`` and consumes optuna's own ``Study`` / ``FrozenTrial`` /
distributions unchanged (optuna_b200/_compat.py).  The GPU box has no optuna, so the parity tests,
``bench.py`` and ``smoke()`` need *something* that calls the plugin the way the reference does.
This module is that caller -- same names, argument meaning and error behaviour as the reference
for the slice the TPE path touches, and nothing more (no storages, pruners, CLI, ...):

* distributions     optuna/distributions.py:109 (Float), :310 (Int), :470 (Categorical)
* TrialState        optuna/trial/_state.py:4-27
* FrozenTrial       optuna/trial/_frozen.py:145-177 ; create_trial :483
* Trial             optuna/trial/_trial.py:56-82 (relative_params), :624-653 (_suggest)
* Study             optuna/study/study.py:269-287 (_get_trials), :388 (optimize), :502 (ask), :588 (tell)
* BaseSampler       optuna/samplers/_base.py:31-228
* RandomSampler     optuna/samplers/_random.py:61-72 (+ _transform.py bounds / untransform)
"""
from __future__ import annotations

import copy
import datetime
import decimal
import enum
import math
import threading
import warnings
from numbers import Real
from typing import Any, Callable, Sequence

import numpy as np


# ------------------------------------------------------------------------------------------------
# exceptions / enums
# ------------------------------------------------------------------------------------------------
class TrialPruned(Exception):
    pass


class ExperimentalWarning(Warning):
    pass


class TrialState(enum.IntEnum):
    RUNNING = 0
    COMPLETE = 1
    PRUNED = 2
    FAIL = 3
    WAITING = 4

    def is_finished(self) -> bool:
        return self != TrialState.RUNNING and self != TrialState.WAITING


class StudyDirection(enum.IntEnum):
    NOT_SET = 0
    MINIMIZE = 1
    MAXIMIZE = 2


# ------------------------------------------------------------------------------------------------
# distributions
# ------------------------------------------------------------------------------------------------
class BaseDistribution:
    def to_external_repr(self, v: float) -> Any:
        return v

    def to_internal_repr(self, v: Any) -> float:  # pragma: no cover - abstract
        raise NotImplementedError

    def single(self) -> bool:  # pragma: no cover - abstract
        raise NotImplementedError

    def _contains(self, v: float) -> bool:  # pragma: no cover - abstract
        raise NotImplementedError

    def _asdict(self) -> dict:
        return self.__dict__

    def __eq__(self, other: Any) -> bool:
        if not isinstance(other, BaseDistribution):
            return NotImplemented
        if type(self) is not type(other):
            return False
        return self.__dict__ == other.__dict__

    def __hash__(self) -> int:
        return hash((self.__class__,) + tuple(sorted(self.__dict__.items())))

    def __repr__(self) -> str:
        kw = ", ".join(f"{k}={v!r}" for k, v in sorted(self._asdict().items()))
        return f"{self.__class__.__name__}({kw})"


def _numeric_internal(v: Any, log: bool) -> float:
    try:
        out = float(v)
    except (ValueError, TypeError) as e:
        raise ValueError(f"'{v}' is not a valid type. float-castable value is expected.") from e
    if math.isnan(out):
        raise ValueError(f"`{v}` is invalid value.")
    if log and out <= 0.0:
        raise ValueError(f"`{v}` is invalid value for the case log=True.")
    return out


class FloatDistribution(BaseDistribution):
    def __init__(self, low: float, high: float, log: bool = False, step: float | None = None) -> None:
        if log and step is not None:
            raise ValueError("The parameter `step` is not supported when `log` is true.")
        if low > high:
            raise ValueError(f"`low <= high` must hold, but got ({low=}, {high=}).")
        if log and low <= 0.0:
            raise ValueError(f"`low > 0` must hold for `log=True`, but got ({low=}, {high=}).")
        if step is not None and step <= 0:
            raise ValueError(f"`step > 0` must hold, but got {step=}.")
        self.step = None
        if step is not None:
            rng_ = decimal.Decimal(str(high)) - decimal.Decimal(str(low))
            d_step = decimal.Decimal(str(step))
            if rng_ % d_step != decimal.Decimal("0"):
                new_high = float((rng_ // d_step) * d_step + decimal.Decimal(str(low)))
                warnings.warn(f"The range [{low}, {high}] is not divisible by {step=}; high becomes {new_high}.")
                high = new_high
            self.step = float(step)
        self.low = float(low)
        self.high = float(high)
        self.log = log

    def single(self) -> bool:
        if self.step is None or self.low == self.high:
            return self.low == self.high
        span = decimal.Decimal(str(self.high)) - decimal.Decimal(str(self.low))
        return span < decimal.Decimal(str(self.step))

    def _contains(self, v: float) -> bool:
        if self.step is None:
            return self.low <= v <= self.high
        k = (v - self.low) / self.step
        return self.low <= v <= self.high and abs(k - round(k)) < 1.0e-8

    def to_internal_repr(self, v: float) -> float:
        return _numeric_internal(v, self.log)


class IntDistribution(BaseDistribution):
    def __init__(self, low: int, high: int, log: bool = False, step: int = 1) -> None:
        if log and step != 1:
            raise ValueError("Samplers and other components only accept step is 1 when `log` argument is True.")
        if low > high:
            raise ValueError(f"`low <= high` must hold, but got ({low=}, {high=}).")
        if log and low < 1:
            raise ValueError(f"`low >= 1` must hold for `log=True`, but got ({low=}, {high=}).")
        if step <= 0:
            raise ValueError(f"`step > 0` must hold, but got {step=}.")
        self.log = log
        self.step = int(step)
        self.low = int(low)
        high = int(high)
        span = high - self.low
        if span % self.step != 0:
            new_high = span // self.step * self.step + self.low
            warnings.warn(f"The range [{low}, {high}] is not divisible by {step=}; high becomes {new_high}.")
            high = new_high
        self.high = high

    def to_external_repr(self, v: float) -> int:
        return int(v)

    def to_internal_repr(self, v: int) -> float:
        return _numeric_internal(v, self.log)

    def single(self) -> bool:
        if self.log or self.low == self.high:
            return self.low == self.high
        return (self.high - self.low) < self.step

    def _contains(self, v: float) -> bool:
        return self.low <= v <= self.high and (v - self.low) % self.step == 0


def _choice_equal(a: Any, b: Any) -> bool:
    a_nan = isinstance(a, Real) and math.isnan(float(a))
    b_nan = isinstance(b, Real) and math.isnan(float(b))
    return (a == b) or (a_nan and b_nan)


class CategoricalDistribution(BaseDistribution):
    def __init__(self, choices: Sequence[Any]) -> None:
        if len(choices) == 0:
            raise ValueError("The `choices` must contain one or more elements.")
        self.choices = tuple(choices)

    def to_external_repr(self, v: float) -> Any:
        return self.choices[int(v)]

    def to_internal_repr(self, v: Any) -> float:
        try:
            return self.choices.index(v)
        except ValueError:
            for i, c in enumerate(self.choices):
                if _choice_equal(v, c):
                    return i
        raise ValueError(f"'{v}' not in {self.choices}.")

    def single(self) -> bool:
        return len(self.choices) == 1

    def _contains(self, v: float) -> bool:
        return 0 <= int(v) < len(self.choices)

    def __eq__(self, other: Any) -> bool:
        if not isinstance(other, BaseDistribution):
            return NotImplemented
        if not isinstance(other, CategoricalDistribution) or len(self.choices) != len(other.choices):
            return False
        return all(_choice_equal(a, b) for a, b in zip(self.choices, other.choices))

    __hash__ = BaseDistribution.__hash__


def check_distribution_compatibility(old: BaseDistribution, new: BaseDistribution) -> None:
    if old.__class__ != new.__class__:
        raise ValueError("Cannot set different distribution kind to the same parameter name.")
    if isinstance(old, (FloatDistribution, IntDistribution)):
        if old.log != new.log:  # type: ignore[union-attr]
            raise ValueError("Cannot set different log configuration to the same parameter name.")
    elif old != new:
        raise ValueError("CategoricalDistribution does not support dynamic value space.")


def _single_value(d: BaseDistribution) -> Any:
    return d.choices[0] if isinstance(d, CategoricalDistribution) else d.low  # type: ignore[union-attr]


# ------------------------------------------------------------------------------------------------
# trials
# ------------------------------------------------------------------------------------------------
class FrozenTrial:
    def __init__(self, number: int, state: TrialState, value: float | None = None,
                 values: Sequence[float] | None = None, params: dict | None = None,
                 distributions: dict | None = None, intermediate_values: dict | None = None,
                 system_attrs: dict | None = None, user_attrs: dict | None = None, trial_id: int | None = None,
                 datetime_start=None, datetime_complete=None) -> None:
        if value is not None and values is not None:
            raise ValueError("Specify only one of `value` and `values`.")
        self.number = number
        self.state = state
        self.values = [value] if value is not None else (list(values) if values is not None else None)
        self.params = dict(params or {})
        self.distributions = dict(distributions or {})
        self.intermediate_values = dict(intermediate_values or {})
        self.system_attrs = dict(system_attrs or {})
        self.user_attrs = dict(user_attrs or {})
        self._trial_id = number if trial_id is None else trial_id
        self.datetime_start = datetime_start
        self.datetime_complete = datetime_complete

    @property
    def value(self) -> float | None:
        if self.values is None:
            return None
        if len(self.values) > 1:
            raise RuntimeError("This attribute is not available during multi-objective optimization.")
        return self.values[0]

    @property
    def last_step(self) -> int | None:
        return max(self.intermediate_values) if self.intermediate_values else None


def create_trial(*, state: TrialState = TrialState.COMPLETE, value: float | None = None,
                 values: Sequence[float] | None = None, params: dict | None = None,
                 distributions: dict | None = None, intermediate_values: dict | None = None,
                 system_attrs: dict | None = None, user_attrs: dict | None = None) -> FrozenTrial:
    params = params or {}
    distributions = distributions or {}
    if set(params) != set(distributions):
        raise ValueError("Inconsistent parameters and distributions.")
    for name, v in params.items():
        d = distributions[name]
        if not d._contains(d.to_internal_repr(v)):
            raise ValueError(f"The value {v} of parameter '{name}' isn't contained in {d}.")
    if state == TrialState.COMPLETE and value is None and values is None:
        raise ValueError("values should be specified for a complete trial.")
    return FrozenTrial(-1, state, value, values, params, distributions, intermediate_values, system_attrs,
                       user_attrs, datetime_start=datetime.datetime.now())


class _Storage:
    """The single in-memory trial table (what the reference's InMemoryStorage gives the sampler)."""

    def __init__(self) -> None:
        self.trials: list[FrozenTrial] = []
        self.lock = threading.RLock()
        self.version = 0  # bumped whenever a trial is added or changes state
        self._cache: dict = {}

    def touch(self) -> None:
        self.version += 1

    def get_all_trials(self, study_id: int = 0, deepcopy: bool = True, states=None) -> list[FrozenTrial]:
        with self.lock:
            key = None if states is None else tuple(sorted(int(s) for s in states))
            hit = self._cache.get(key)
            if hit is not None and hit[0] == (self.version, len(self.trials), id(self.trials)):
                out = hit[1]
            else:
                out = [t for t in self.trials if states is None or t.state in states]
                self._cache[key] = ((self.version, len(self.trials), id(self.trials)), out)
        return copy.deepcopy(out) if deepcopy else out

    def set_trial_system_attr(self, trial_id: int, key: str, value: Any) -> None:
        with self.lock:
            self.trials[trial_id].system_attrs[key] = value

    def get_trial(self, trial_id: int) -> FrozenTrial:
        return self.trials[trial_id]


class Trial:
    def __init__(self, study: "Study", trial_id: int) -> None:
        self.study = study
        self._trial_id = trial_id
        self._frozen = study._storage.get_trial(trial_id)
        self.study.sampler.before_trial(study, self._frozen)
        self.relative_search_space: dict | None = None
        self._relative_params: dict | None = None

    @property
    def number(self) -> int:
        return self._frozen.number

    @property
    def params(self) -> dict:
        return dict(self._frozen.params)

    @property
    def relative_params(self) -> dict:
        if self._relative_params is None:
            s = self.study.sampler
            self.relative_search_space = s.infer_relative_search_space(self.study, self._frozen)
            self._relative_params = s.sample_relative(self.study, self._frozen, self.relative_search_space)
        return self._relative_params

    def _suggest(self, name: str, d: BaseDistribution) -> Any:
        t = self._frozen
        if name in t.distributions:
            check_distribution_compatibility(t.distributions[name], d)
            return t.params[name]
        if d.single():
            v = _single_value(d)
        elif self._is_relative(name, d):
            v = self.relative_params[name]
        else:
            v = self.study.sampler.sample_independent(self.study, t, name, d)
        d.to_internal_repr(v)  # validates
        with self.study._storage.lock:
            t.distributions[name] = d
            t.params[name] = v
        return v

    def _is_relative(self, name: str, d: BaseDistribution) -> bool:
        if name not in self.relative_params:
            return False
        assert self.relative_search_space is not None
        if name not in self.relative_search_space:
            raise ValueError(f"The parameter {name} was sampled by `sample_relative` method but it is not "
                             "contained in the relative search space.")
        check_distribution_compatibility(self.relative_search_space[name], d)
        return d._contains(d.to_internal_repr(self.relative_params[name]))

    def suggest_float(self, name: str, low: float, high: float, *, step: float | None = None, log: bool = False):
        return self._suggest(name, FloatDistribution(low, high, log=log, step=step))

    def suggest_int(self, name: str, low: int, high: int, *, step: int = 1, log: bool = False) -> int:
        return self._suggest(name, IntDistribution(low, high, log=log, step=step))

    def suggest_categorical(self, name: str, choices: Sequence[Any]):
        return self._suggest(name, CategoricalDistribution(choices))

    def report(self, value: float, step: int) -> None:
        self._frozen.intermediate_values[int(step)] = float(value)

    def should_prune(self) -> bool:
        return False

    def set_user_attr(self, key: str, value: Any) -> None:
        self._frozen.user_attrs[key] = value


# ------------------------------------------------------------------------------------------------
# samplers
# ------------------------------------------------------------------------------------------------
_CONSTRAINTS_KEY = "constraints"


class BaseSampler:
    def infer_relative_search_space(self, study: "Study", trial: FrozenTrial) -> dict:
        raise NotImplementedError

    def sample_relative(self, study: "Study", trial: FrozenTrial, search_space: dict) -> dict:
        raise NotImplementedError

    def sample_independent(self, study: "Study", trial: FrozenTrial, param_name: str,
                           param_distribution: BaseDistribution) -> Any:
        raise NotImplementedError

    def before_trial(self, study: "Study", trial: FrozenTrial) -> None:
        pass

    def after_trial(self, study: "Study", trial: FrozenTrial, state: TrialState,
                    values: Sequence[float] | None) -> None:
        pass

    def reseed_rng(self) -> None:
        pass


class LazyRandomState:
    def __init__(self, seed: int | None = None) -> None:
        self._rng: np.random.RandomState | None = None
        if seed is not None:
            self.rng.seed(seed=seed)

    @property
    def rng(self) -> np.random.RandomState:
        if self._rng is None:
            self._rng = np.random.RandomState()
        return self._rng


def random_independent(rng: np.random.RandomState, d: BaseDistribution) -> Any:
    """RandomSampler.sample_independent: uniform draw on the transformed bounds, then untransform
    (optuna/samplers/_random.py:61-72, optuna/_transform.py:170-330)."""
    if isinstance(d, CategoricalDistribution):
        n = len(d.choices)
        u = rng.uniform(np.zeros(n), np.ones(n))
        return d.to_external_repr(int(u.argmax()))
    half = 0.0
    if isinstance(d, IntDistribution) or d.step is not None:  # type: ignore[union-attr]
        half = 0.5 * d.step  # type: ignore[union-attr,operator]
    if d.log:  # type: ignore[union-attr]
        lo, hi = math.log(d.low - half), math.log(d.high + half)  # type: ignore[union-attr]
    else:
        lo, hi = float(d.low) - half, float(d.high) + half  # type: ignore[union-attr]
    u = rng.uniform(np.asarray([lo]), np.asarray([hi]))
    x = u.item()
    if isinstance(d, FloatDistribution):
        if d.log:
            v = math.exp(x)
            return v if d.single() else min(v, np.nextafter(d.high, d.high - 1))
        if d.step is not None:
            return float(np.clip(np.round((x - d.low) / d.step) * d.step + d.low, d.low, d.high))
        return x if d.single() else min(x, np.nextafter(d.high, d.high - 1))
    assert isinstance(d, IntDistribution)
    if d.log:
        return int(np.clip(np.round(math.exp(x)), d.low, d.high))
    return int(np.clip(np.round((x - d.low) / d.step) * d.step + d.low, d.low, d.high))


class RandomSampler(BaseSampler):
    def __init__(self, seed: int | None = None) -> None:
        self._rng = LazyRandomState(seed)

    def reseed_rng(self) -> None:
        self._rng.rng.seed()

    def infer_relative_search_space(self, study, trial) -> dict:
        return {}

    def sample_relative(self, study, trial, search_space) -> dict:
        return {}

    def sample_independent(self, study, trial, param_name, param_distribution):
        return random_independent(self._rng.rng, param_distribution)


# ------------------------------------------------------------------------------------------------
# study
# ------------------------------------------------------------------------------------------------
class Study:
    def __init__(self, sampler: BaseSampler, directions: Sequence[StudyDirection]) -> None:
        self.sampler = sampler
        self._directions = list(directions)
        self._storage = _Storage()
        self._study_id = 0
        self._thread_local = threading.local()

    @property
    def directions(self) -> list[StudyDirection]:
        return list(self._directions)

    @property
    def direction(self) -> StudyDirection:
        if len(self._directions) > 1:
            raise RuntimeError("A single direction cannot be retrieved from a multi-objective study.")
        return self._directions[0]

    def _is_multi_objective(self) -> bool:
        return len(self._directions) > 1

    def _get_trials(self, deepcopy: bool = True, states=None, use_cache: bool = False) -> list[FrozenTrial]:
        return self._storage.get_all_trials(self._study_id, deepcopy=deepcopy, states=states)

    def get_trials(self, deepcopy: bool = True, states=None) -> list[FrozenTrial]:
        return self._get_trials(deepcopy, states)

    @property
    def trials(self) -> list[FrozenTrial]:
        return self.get_trials(deepcopy=True)

    @property
    def best_trial(self) -> FrozenTrial:
        done = [t for t in self._get_trials(False, (TrialState.COMPLETE,))]
        if not done:
            raise ValueError("No trials are completed yet.")
        pick = min if self.direction == StudyDirection.MINIMIZE else max
        return copy.deepcopy(pick(done, key=lambda t: t.value))

    @property
    def best_value(self) -> float:
        return self.best_trial.value  # type: ignore[return-value]

    @property
    def best_params(self) -> dict:
        return self.best_trial.params

    def add_trial(self, trial: FrozenTrial) -> None:
        with self._storage.lock:
            t = copy.copy(trial)
            t.number = t._trial_id = len(self._storage.trials)
            self._storage.trials.append(t)
            self._storage.touch()

    def add_trials(self, trials: Sequence[FrozenTrial]) -> None:
        for t in trials:
            self.add_trial(t)

    def ask(self, fixed_distributions: dict | None = None) -> Trial:
        with self._storage.lock:
            n = len(self._storage.trials)
            self._storage.trials.append(FrozenTrial(n, TrialState.RUNNING, datetime_start=datetime.datetime.now()))
            self._storage.touch()
        trial = Trial(self, n)
        for name, d in (fixed_distributions or {}).items():
            trial._suggest(name, d)
        return trial

    def tell(self, trial: Trial | int, values: float | Sequence[float] | None = None,
             state: TrialState | None = None) -> FrozenTrial:
        number = trial.number if isinstance(trial, Trial) else int(trial)
        frozen = self._storage.get_trial(number)
        if state is None:
            state = TrialState.COMPLETE
        vals: list[float] | None = None
        if state == TrialState.COMPLETE:
            if values is None:
                raise ValueError("No values were told.")
            vals = [float(values)] if isinstance(values, Real) else [float(v) for v in values]
            if len(vals) != len(self._directions):
                raise ValueError("The number of the values and the number of the objectives are mismatched.")
            if any(math.isnan(v) for v in vals):
                state, vals = TrialState.FAIL, None
        elif state == TrialState.PRUNED and frozen.intermediate_values:
            last = frozen.intermediate_values[max(frozen.intermediate_values)]
            if not math.isnan(last):
                vals = [last]
        try:
            self.sampler.after_trial(self, frozen, state, vals)
        finally:
            with self._storage.lock:
                frozen.values = vals
                frozen.state = state
                frozen.datetime_complete = datetime.datetime.now()
                self._storage.touch()
        return copy.deepcopy(frozen)

    def optimize(self, func: Callable[[Trial], Any], n_trials: int) -> None:
        for _ in range(n_trials):
            trial = self.ask()
            try:
                out = func(trial)
            except TrialPruned:
                self.tell(trial, state=TrialState.PRUNED)
                continue
            self.tell(trial, out)


def create_study(*, sampler: BaseSampler, direction: str | None = None,
                 directions: Sequence[str] | None = None) -> Study:
    if direction is not None and directions is not None:
        raise ValueError("Specify only one of `direction` and `directions`.")
    names = list(directions) if directions is not None else [direction or "minimize"]
    dirs = []
    for d in names:
        if d not in ("minimize", "maximize"):
            raise ValueError("Please set either 'minimize' or 'maximize' to direction.")
        dirs.append(StudyDirection.MINIMIZE if d == "minimize" else StudyDirection.MAXIMIZE)
    return Study(sampler, dirs)
