"""BASELINE config 3: 64 mixed params (24 float, 8 log-float, 8 step-float, 8 int, 4 log-int, 12 categorical),
N = 50 000, multivariate -- stage timings through the C ABI."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optuna_b200 import ParamSpec, TPEEngine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=50_000)
ap.add_argument("--c", type=int, default=24)
ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
rs = np.random.RandomState(0)
specs, cols = [], []
for _ in range(24):
    specs.append(ParamSpec(kind=0, low=0.0, high=1.0)); cols.append(rs.uniform(0, 1, a.n))
for _ in range(8):
    specs.append(ParamSpec(kind=0, low=1e-5, high=1.0, log=True)); cols.append(np.exp(rs.uniform(np.log(1e-5), 0, a.n)))
for _ in range(8):
    specs.append(ParamSpec(kind=0, low=0.0, high=10.0, step=0.5)); cols.append(rs.randint(0, 21, a.n) * 0.5)
for _ in range(8):
    specs.append(ParamSpec(kind=1, low=0, high=100, step=1)); cols.append(rs.randint(0, 101, a.n).astype(float))
for _ in range(4):
    specs.append(ParamSpec(kind=1, low=1, high=1024, step=1, log=True)); cols.append(np.round(np.exp(rs.uniform(0, np.log(1024), a.n))))
for k in range(12):
    nch = 4 + k % 5
    specs.append(ParamSpec(kind=2, n_choices=nch)); cols.append(rs.randint(0, nch, a.n).astype(float))
X = np.stack(cols, 1)
key = np.stack([rs.normal(size=a.n), np.zeros(a.n)], 1)
eng = TPEEngine(0)
eng.set_space(specs)
eng.set_history(X, np.zeros(a.n, np.int8), key)
ncat, nnum = 12, 52
rng = np.random.RandomState(1)
for s in range(a.steps):
    u = rng.random_sample(a.c * (1 + ncat + nnum))
    x, acq, best = eng.suggest(list(range(64)), u, 1, n_below=25, n_candidates=a.c, multivariate=True)
    ms, nl = eng.last_timing()
    print("step", s, "stage ms", np.round(ms, 4).tolist(), "launches", nl, eng.last_logpdf_kernel())
