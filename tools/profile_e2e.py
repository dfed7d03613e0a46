"""Host-side time of one B200TPESampler.sample_relative call at config 2, by phase (wall clock)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optuna_b200 import B200TPESampler, mini  # noqa: E402
from optuna_b200.engine import TPEEngine  # noqa: E402

N, P, C = 100_000, 32, 4096
rs = np.random.RandomState(0)
X = rs.uniform(0, 1, (N, P))
loss = ((X - 0.5) ** 2).sum(1)
space = {f"x{j:02d}": mini.FloatDistribution(0.0, 1.0) for j in range(P)}
sampler = B200TPESampler(seed=1, n_ei_candidates=C, multivariate=True)
study = mini.create_study(sampler=sampler)
names = list(space)
study._storage.trials = [mini.FrozenTrial(i, mini.TrialState.COMPLETE, value=float(loss[i]),
                                          params=dict(zip(names, X[i].tolist())), distributions=space)
                         for i in range(N)]
frozen = mini.FrozenTrial(N, mini.TrialState.RUNNING)

acc: dict[str, float] = {}


def timed(obj, name):
    f = getattr(obj, name)

    def g(*a, **k):
        t0 = time.perf_counter()
        r = f(*a, **k)
        acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0
        return r
    setattr(obj, name, g)


for _ in range(3):
    sampler.sample_relative(study, frozen, space)
for nm in ("prepare", "build", "sample_and_select"):
    timed(sampler._engine, nm)
for nm in ("_sync", "_draw_uniforms"):
    timed(sampler, nm)
steps = 30
t0 = time.perf_counter()
for _ in range(steps):
    sampler.sample_relative(study, frozen, space)
tot = time.perf_counter() - t0
print(f"per call {1e3 * tot / steps:.3f} ms: " + ", ".join(f"{k} {1e3 * v / steps:.3f}" for k, v in acc.items())
      + f", other {1e3 * (tot - sum(acc.values())) / steps:.3f}")
