"""Where the time of an end-to-end step goes (optuna's Study -> B200TPESampler -> libtpe_b200.so) at N = 100k:
wall time of every engine call, of the sampler's sync, and a cProfile of 20 steps."""
import cProfile
import collections
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

optuna = bench.import_optuna()
from optuna_b200 import B200TPESampler, TPEEngine  # noqa: E402

acc = collections.defaultdict(float)
cnt = collections.Counter()
for name in ("update_history", "prepare", "build", "stage_rng", "sample_and_select", "set_history", "set_space"):
    orig = getattr(TPEEngine, name)

    def wrap(self, *a, _o=orig, _n=name, **k):
        t0 = time.perf_counter()
        try:
            return _o(self, *a, **k)
        finally:
            acc[_n] += time.perf_counter() - t0
            cnt[_n] += 1
    setattr(TPEEngine, name, wrap)

X, loss = bench.synthetic_history()
s = B200TPESampler(seed=1, n_ei_candidates=bench.N_CAND, multivariate=True)
study, space = bench.build_study(optuna, s, X, loss)


def one():
    t = study.ask()
    x = [t.suggest_float(n, 0.0, 1.0) for n in bench.NAMES]
    study.tell(t, sum((v - 0.5) ** 2 for v in x))


for _ in range(4):
    one()
acc.clear()
cnt.clear()
sync = dev = 0.0
t0 = time.perf_counter()
for _ in range(40):
    one()
    sync += s.last_ask_s[0]
    dev += s.last_ask_s[1]
wall = time.perf_counter() - t0
print("per trial ms %.3f | sampler sync %.3f | device calls %.3f" % (wall / 40 * 1e3, sync / 40 * 1e3, dev / 40 * 1e3))
for k, v in sorted(acc.items(), key=lambda kv: -kv[1]):
    print("  engine.%-20s %.3f ms per call x %d" % (k, v / max(cnt[k], 1) * 1e3, cnt[k]))
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    one()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
