"""Run a few config-2 suggestions through the C ABI (target for ncu; see profiles/README.md)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optuna_b200 import ParamSpec, TPEEngine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--n", type=int, default=100_000)
ap.add_argument("--p", type=int, default=32)
ap.add_argument("--c", type=int, default=4096)
ap.add_argument("--asks", type=int, default=1)
ap.add_argument("--univariate", action="store_true")
a = ap.parse_args()
rs = np.random.RandomState(0)
X = rs.uniform(0, 1, (a.n, a.p))
key = np.stack([((X - 0.5) ** 2).sum(1), np.zeros(a.n)], 1)
eng = TPEEngine(0)
eng.set_space([ParamSpec(kind=0, low=0.0, high=1.0) for _ in range(a.p)])
eng.set_history(X, np.zeros(a.n, np.int8), key)
rng = np.random.RandomState(1)
for s in range(a.steps):
    if a.univariate:
        for j in range(a.p):
            u = rng.random_sample(a.asks * a.c * 2)
            eng.suggest([j], u, a.asks, n_below=25, n_candidates=a.c, multivariate=False)
    else:
        u = rng.random_sample(a.asks * a.c * (1 + a.p))
        x, acq, best = eng.suggest(list(range(a.p)), u, a.asks, n_below=25, n_candidates=a.c, multivariate=True)
    ms, nl = eng.last_timing()
    print("step", s, "stage ms", np.round(ms, 4).tolist(), "launches", nl, eng.last_logpdf_kernel())
