"""The objects the sampler plugs into: optuna's own.

``B200TPESampler`` is a plugin for optuna (``optuna.samplers.BaseSampler``, optuna/samplers/_base.py:31-228): it
needs the optuna package the study lives in -- its distributions, ``TrialState``, ``RandomSampler`` for the
startup trials (sampler.py:348-349, :471-474), the constraint bookkeeping -- and uses exactly those classes, never
private look-alikes.  The array-level engine (``optuna_b200.TPEEngine``) needs none of this and imports without
optuna.
"""
from __future__ import annotations

try:
    import optuna  # noqa: F401
except ImportError as e:  # pragma: no cover - depends on the environment
    raise ImportError(
        "optuna_b200.B200TPESampler is a sampler plugin for optuna and needs the `optuna` package "
        "(the array-level optuna_b200.TPEEngine does not)") from e

from optuna._experimental import warn_experimental_argument  # noqa: E402,F401
from optuna._warnings import optuna_warn  # noqa: E402,F401
from optuna.distributions import (BaseDistribution, CategoricalDistribution, FloatDistribution,  # noqa: E402,F401
                                  IntDistribution)
from optuna.logging import get_logger  # noqa: E402,F401
from optuna.samplers import BaseSampler, RandomSampler  # noqa: E402,F401
from optuna.samplers._base import (_CONSTRAINTS_KEY as CONSTRAINTS_KEY,  # noqa: E402,F401
                                   _INDEPENDENT_SAMPLING_WARNING_TEMPLATE, _process_constraints_after_trial)
from optuna.samplers._lazy_random_state import LazyRandomState  # noqa: E402,F401
from optuna.storages import InMemoryStorage  # noqa: E402,F401
from optuna.study import StudyDirection  # noqa: E402,F401
from optuna.trial import FrozenTrial, TrialState  # noqa: E402,F401

RELATIVE_PARAMS_KEY = "tpe:relative_params"  # optuna/samplers/_tpe/sampler.py:48
SYSTEM_ATTR_MAX_LENGTH = 2045  # sampler.py:50
