"""Debug helper (GPU box): first divergence of the CUDA engine from the stable-sort oracle engine in the
univariate mixed-space scenario of tests/test_plugin_optuna.py, with the inputs of the diverging call."""
import sys, os, warnings, cProfile, pstats, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref; ref.enable()
import numpy as np, optuna
warnings.simplefilter("ignore")
from optuna_b200 import B200TPESampler, TPEEngine
from tests._oracle_engine import StableOracleEngine
from tests.test_plugin_optuna import mixed

kw = dict(seed=3, multivariate=False, n_startup_trials=5)
a = optuna.create_study(sampler=B200TPESampler(**kw), direction="minimize")
sb = B200TPESampler(**kw); sb._engine_cls = StableOracleEngine
b = optuna.create_study(sampler=sb, direction="minimize")

# record every engine-level call of the oracle arm so the diverging one can be replayed on the CUDA engine
log = []
orig_prepare, orig_sas = StableOracleEngine.prepare, StableOracleEngine.sample_and_select
def prepare(self, cols, **cfg):
    log.append(dict(cols=list(cols), cfg=dict(cfg), X=self.X.copy(), cat=self.cat.copy(), key=self.key.copy(), specs=list(self.specs)))
    return orig_prepare(self, cols, **cfg)
def sas(self, uniforms, n_asks=1):
    out = orig_sas(self, uniforms, n_asks)
    log[-1].update(u=np.array(uniforms), x=out[0].copy(), best=out[2].copy(),
                   below=self._below.copy(), above=self._above.copy())
    return out
StableOracleEngine.prepare, StableOracleEngine.sample_and_select = prepare, sas

for i in range(45):
    a.optimize(mixed, n_trials=1); b.optimize(mixed, n_trials=1)
    pa, pb = a.trials[-1].params, b.trials[-1].params
    bad = [k for k in pb if pa.get(k) != pb[k] and not (isinstance(pb[k], float) and abs(pa[k]-pb[k]) < 1e-9)]
    if bad:
        print("trial", i, "differs in", bad, pa, pb)
        names = list(b.sampler._hist.columns)
        k = bad[0]
        # the oracle arm's call for that parameter
        call = [c for c in log[-7:] if [names[j] for j in c["cols"]] == [k]][-1]
        eng = TPEEngine(0)
        eng.set_space(call["specs"]); eng.set_history(call["X"], call["cat"], call["key"])
        x, acq, best = eng.suggest(call["cols"], call["u"], 1, **call["cfg"])
        below, above = eng.get_split()
        print("replay on CUDA: x", x, "oracle x", call["x"], "best", best, call["best"])
        print("split equal:", np.array_equal(below, call["below"]), np.array_equal(above, call["above"]))
        print("n rows", len(call["cat"]), "cats", np.bincount(call["cat"], minlength=5), "cfg", call["cfg"])
        smp, ll, lg = eng.get_candidates()
        from oracle import tpe_oracle as orc
        o = StableOracleEngine(); o.set_space(call["specs"]); o.set_history(call["X"], call["cat"], call["key"])
        o.prepare(call["cols"], **call["cfg"]); o.build()
        from tests._oracle_engine import _ReplayRng
        cand = orc.mixture_sample(o._mix_b, _ReplayRng(call["u"]), call["cfg"]["n_candidates"])
        oll, olg = orc.mixture_log_pdf(o._mix_b, cand), orc.mixture_log_pdf(o._mix_a, cand)
        print("cand equal", np.array_equal(cand, smp), "max|dll|", np.abs(ll-oll).max(), "max|dlg|", np.abs(lg-olg).max())
        print("acq cuda", (ll-lg)[:24]); print("acq orc ", (oll-olg)[:24]); print("cand", cand.ravel()[:24])
        w, mu, sg = eng.get_mixture(1)
        print("above weights diff", np.abs(w - o._mix_a.weights).max())
        break
else:
    print("no divergence in 45 trials")
