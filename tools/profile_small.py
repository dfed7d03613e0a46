"""Latency of one sample_independent call on a small study (N = 150, univariate): host phases (wall clock)
and device stages (CUDA events)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optuna_b200 import B200TPESampler, mini  # noqa: E402

N = 150
rs = np.random.RandomState(0)
space = {"x": mini.FloatDistribution(-5.0, 10.0), "y": mini.FloatDistribution(0.0, 15.0)}
sampler = B200TPESampler(seed=1)
study = mini.create_study(sampler=sampler)
study._storage.trials = [mini.FrozenTrial(i, mini.TrialState.COMPLETE, value=float(rs.normal()),
                                          params={"x": float(rs.uniform(-5, 10)), "y": float(rs.uniform(0, 15))},
                                          distributions=space) for i in range(N)]
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


for _ in range(5):
    sampler.sample_independent(study, frozen, "x", space["x"])
for nm in ("prepare", "build", "sample_and_select", "suggest"):
    timed(sampler._engine, nm)
for nm in ("_sync", "_draw_uniforms"):
    timed(sampler, nm)
steps = 200
t0 = time.perf_counter()
for i in range(steps):
    sampler.sample_independent(study, frozen, "xy"[i & 1], space["xy"[i & 1]])
tot = time.perf_counter() - t0
ms, nl = sampler._engine.last_timing()
print(f"per call {1e3 * tot / steps:.3f} ms: " + ", ".join(f"{k} {1e3 * v / steps:.3f}" for k, v in acc.items())
      + f", other {1e3 * (tot - sum(acc.values())) / steps:.3f}")
print("device stage ms", np.round(ms, 4).tolist(), "launches", nl)
