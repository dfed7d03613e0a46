"""cProfile of the end-to-end step of bench.py (optuna's Study -> B200TPESampler) at N = 100k."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

optuna = bench.import_optuna()
from optuna_b200 import B200TPESampler  # noqa: E402

X, loss = bench.synthetic_history()
s = B200TPESampler(seed=1, n_ei_candidates=bench.N_CAND, multivariate=True)
study, space = bench.build_study(optuna, s, X, loss)


def one():
    t = study.ask()
    x = [t.suggest_float(n, 0.0, 1.0) for n in bench.NAMES]
    study.tell(t, sum((v - 0.5) ** 2 for v in x))


for _ in range(4):
    one()
t0 = time.perf_counter()
for _ in range(20):
    one()
print("per trial ms", (time.perf_counter() - t0) / 20 * 1e3, "last sync/device s", s.last_ask_s)
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    one()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(30)
