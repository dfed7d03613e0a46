"""Where the host time of the univariate look-ahead goes (cProfile, N = 30 000)."""
import cProfile, os, pstats, sys, time, warnings
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref
ref.enable()
import optuna
from optuna_b200 import B200TPESampler
warnings.filterwarnings("ignore"); optuna.logging.set_verbosity(optuna.logging.ERROR)
N, P, C = int(os.environ.get("N", 30_000)), 32, 4096
names = [f"x{j}" for j in range(P)]
dist = {n: optuna.distributions.FloatDistribution(0.0, 1.0) for n in names}
rs = np.random.RandomState(0); X = rs.uniform(0, 1, (N, P))
smp = B200TPESampler(seed=1, n_ei_candidates=C, multivariate=False)
study = optuna.create_study(sampler=smp)
study.add_trials([optuna.trial.create_trial(params=dict(zip(names, row.tolist())), distributions=dist, value=float(((row - 0.5) ** 2).sum())) for row in X])
def one():
    t = study.ask(); x = [t.suggest_float(n, 0.0, 1.0) for n in names]; study.tell(t, sum((v - 0.5) ** 2 for v in x))
for _ in range(5): one()
pr = cProfile.Profile(); pr.enable()
for _ in range(30): one()
pr.disable()
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(14)
print(smp.ahead_stats)
