"""Pick the objects the sampler plugs into: optuna's own when optuna is importable (drop-in use),
otherwise the minimal mirror in optuna_b200/mini.py (GPU box, tests, bench)."""
from __future__ import annotations

try:  # pragma: no cover - depends on the environment
    import optuna  # noqa: F401
    from optuna.distributions import (BaseDistribution, CategoricalDistribution, FloatDistribution,
                                      IntDistribution)
    from optuna.samplers import BaseSampler
    from optuna.study import StudyDirection
    from optuna.trial import FrozenTrial, TrialState
    HAVE_OPTUNA = True
except Exception:  # ImportError, or a broken partial install
    from .mini import (BaseDistribution, BaseSampler, CategoricalDistribution, FloatDistribution,  # noqa: F401
                       FrozenTrial, IntDistribution, StudyDirection, TrialState)
    HAVE_OPTUNA = False

from .mini import LazyRandomState, random_independent  # noqa: E402,F401  (pure numpy helpers)

CONSTRAINTS_KEY = "constraints"  # optuna/samplers/_base.py:23
RELATIVE_PARAMS_KEY = "tpe:relative_params"  # optuna/samplers/_tpe/sampler.py:48
SYSTEM_ATTR_MAX_LENGTH = 2045  # sampler.py:50
