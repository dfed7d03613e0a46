"""BASELINE config 1: 200-trial Branin study through the sampler plugin -- wall time (launch-latency bound)."""
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optuna_b200 import B200TPESampler, mini  # noqa: E402


def branin(t):
    x = t.suggest_float("x", -5.0, 10.0)
    y = t.suggest_float("y", 0.0, 15.0)
    a, b, c, r, s, tt = 1.0, 5.1 / (4 * math.pi ** 2), 5 / math.pi, 6.0, 10.0, 1 / (8 * math.pi)
    return a * (y - b * x * x + c * x - r) ** 2 + s * (1 - tt) * math.cos(x) + s


for mv in (False, True):
    for rep in range(2):
        study = mini.create_study(sampler=B200TPESampler(seed=0, multivariate=mv))
        t0 = time.perf_counter()
        study.optimize(branin, n_trials=200)
        dt = time.perf_counter() - t0
    print(f"multivariate={mv}: 200 trials in {1e3 * dt:.1f} ms = {1e3 * dt / 190:.3f} ms per TPE trial, best {study.best_value:.6f}")
