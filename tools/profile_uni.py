"""One batched univariate trial at config 2 (target for an ncu launch list)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optuna_b200 import ParamSpec, TPEEngine  # noqa: E402

N, P, C = 100_000, 32, 4096
rs = np.random.RandomState(0)
X = rs.uniform(0, 1, (N, P))
eng = TPEEngine(0)
eng.set_space([ParamSpec(kind=0, low=0.0, high=1.0) for _ in range(P)])
eng.set_history(X, np.zeros(N, np.int8), np.stack([((X - 0.5) ** 2).sum(1), np.zeros(N)], 1))
cfg = dict(n_below=25, n_candidates=C, multivariate=False)
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    u = rs.random_sample(P * 2 * C)
    t0 = time.perf_counter()
    eng.suggest_univariate_batch(list(range(P)), u, **cfg)
    print("batch ms", (time.perf_counter() - t0) * 1e3, "device span", eng.last_timing()[0][8])
