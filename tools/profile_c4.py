"""BASELINE config 4: MOTPE, 4 objectives, N = 20 000, P = 8 floats, multivariate -- wall time per stage."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optuna_b200 import ParamSpec, TPEEngine  # noqa: E402

n, P, C = 20000, 8, 24
rs = np.random.RandomState(3)
X = rs.uniform(0, 1, (n, P))
cs = np.array([0.2, 0.4, 0.6, 0.8])
vals = np.stack([((X - c) ** 2).sum(1) for c in cs], 1)
eng = TPEEngine(0)
eng.set_space([ParamSpec(kind=0, low=0.0, high=1.0) for _ in range(P)])
eng.set_history(X, np.zeros(n, np.int8), np.zeros((n, 2)))
eng.set_values(vals, 0)
rng = np.random.RandomState(1)
for s in range(4):
    u = rng.random_sample(C * (1 + P))
    t0 = time.perf_counter()
    eng.prepare(list(range(P)), n_below=25, n_candidates=C, multivariate=True)
    t1 = time.perf_counter()
    eng.build()
    x, acq, best = eng.sample_and_select(u, 1)
    t2 = time.perf_counter()
    print(f"step {s}: prepare (ranks + HSSP + split) {1e3 * (t1 - t0):.3f} ms, build+sample+select {1e3 * (t2 - t1):.3f} ms,"
          f" below {eng.split_info()}")
