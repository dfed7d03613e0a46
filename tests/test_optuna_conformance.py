"""optuna's own sampler conformance suite (optuna/testing/pytest_samplers.py:81-540 -- the classes optuna's
tests/samplers_tests/test_samplers.py:92-140 instantiates for TPESampler) run unmodified against B200TPESampler.

``sampler`` is the fixture the suite asks for: a zero-argument factory.  Engines: CPU oracle (anywhere) and the CUDA
library (``gpu``)."""
import warnings

import pytest

optuna = pytest.importorskip("optuna")
from optuna.testing.pytest_samplers import (BasicSamplerTestCase, MultiObjectiveSamplerTestCase,  # noqa: E402
                                            RelativeSamplerTestCase)

warnings.filterwarnings("ignore", category=optuna.exceptions.ExperimentalWarning)
pytestmark = pytest.mark.filterwarnings("ignore::optuna.exceptions.ExperimentalWarning")

# the parameter sets optuna uses for TPESampler (tests/samplers_tests/test_samplers.py:96-97, :114, :130-131)
BASIC = [dict(n_startup_trials=0), dict(n_startup_trials=0, multivariate=True)]
RELATIVE = [dict(n_startup_trials=0, multivariate=True)]
MULTI = [dict(n_startup_trials=0), dict(n_startup_trials=0, multivariate=True)]


def _ids(ps):
    return ["mv" if p.get("multivariate") else "uni" for p in ps]


class TestBasicSampler(BasicSamplerTestCase):
    @pytest.fixture(params=BASIC, ids=_ids(BASIC))
    def sampler(self, request, make_sampler):
        return lambda: make_sampler(**request.param)


class TestRelativeSampler(RelativeSamplerTestCase):
    @pytest.fixture(params=RELATIVE, ids=_ids(RELATIVE))
    def sampler(self, request, make_sampler):
        return lambda: make_sampler(**request.param)


class TestMultiObjectiveSampler(MultiObjectiveSamplerTestCase):
    @pytest.fixture(params=MULTI, ids=_ids(MULTI))
    def sampler(self, request, make_sampler):
        return lambda: make_sampler(**request.param)
