import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests._util import load, specs_from_space, decode_space
from tests.test_gpu_parity import _run_case
from optuna_b200 import TPEEngine
eng = TPEEngine(0)
g = load("suggest.npz")
for ci in range(int(g["n_cases"])):
    t = f"sg{ci}/"
    mv, C, seed, n_below = g[t + "cfg"]
    mv, C, n_below = bool(mv), int(C), int(n_below)
    if mv: continue
    X, cat, key = g[t + "X"], g[t + "category"], g[t + "key"]
    P = X.shape[1]
    eng.set_space(specs_from_space(g[t + "space"]))
    eng.set_history(X, cat, key)
    rng = np.random.RandomState(int(seed))
    for j in range(P):
        params, x, acq, best, smp, ll, lg = _run_case(eng, g, t, [j], rng, False, C, n_below)
        p = params[0]
        if p.is_cat or p.step is not None: continue
        for nm, mine in (("ll", ll), ("lg", lg)):
            ref = g[f"{t}u{j}/{nm}"]
            bad = ~(np.abs(mine - ref) <= 1e-12) & ~(np.isinf(mine) & np.isinf(ref) & (np.sign(mine) == np.sign(ref)))
            if bad.any():
                idx = np.flatnonzero(bad)
                print("case", ci, "col", j, nm, "C", C, "param", p, "kernel", eng.last_logpdf_kernel())
                print("  idx", idx, "mine", mine[idx], "ref", ref[idx], "x", smp[idx].ravel())
                below, above = eng.get_split()
                print("  n_below", len(below), "n_above", len(above))
                w, mu, sg = eng.get_mixture(0 if nm == "ll" else 1)
                print("  K", len(w), "mu range", mu.min(), mu.max(), "sigma range", sg.min(), sg.max())
