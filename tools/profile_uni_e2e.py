"""Host-side profile of univariate trials through optuna's Study (cProfile over 30 trials, N = 20 000)."""
import cProfile
import os
import pstats
import sys
import warnings

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref  # noqa: E402

ref.enable()
import optuna  # noqa: E402

from optuna_b200 import B200TPESampler  # noqa: E402

warnings.filterwarnings("ignore")
optuna.logging.set_verbosity(optuna.logging.ERROR)
N, P, C = 20_000, 32, 4096
names = [f"x{j}" for j in range(P)]
dist = {n: optuna.distributions.FloatDistribution(0.0, 1.0) for n in names}
rs = np.random.RandomState(0)
X = rs.uniform(0, 1, (N, P))
study = optuna.create_study(sampler=B200TPESampler(seed=1, n_ei_candidates=C, multivariate=False))
study.add_trials([optuna.trial.create_trial(params=dict(zip(names, row.tolist())), distributions=dist,
                                            value=float(((row - 0.5) ** 2).sum())) for row in X])


def one():
    t = study.ask()
    x = [t.suggest_float(n, 0.0, 1.0) for n in names]
    study.tell(t, sum((v - 0.5) ** 2 for v in x))


for _ in range(5):
    one()
pr = cProfile.Profile()
pr.enable()
for _ in range(30):
    one()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(28)
