"""CUDA path (through the C ABI) vs the oracle and the live-reference golden vectors.

Tolerances: split indices and categorical / integer draws exact;
mu / sigma / weights <= 1e-15 relative (sigma of multivariate kernels goes through device pow():
<= 2 ulp); sampled floats <= 1e-12 relative; log_pdf <= 1e-12 absolute for continuous and
categorical columns; discrete columns carry the reference's own ill-conditioning
 so they get 1e-9.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import tpe_oracle as orc
from tests._util import decode_space, draw_uniforms, kinds_of, load, specs_from_space

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from optuna_b200 import TPEEngine
    e = TPEEngine(0)
    yield e
    e.close()


def close(a, b, rtol, atol=0.0):
    a, b = np.asarray(a, float), np.asarray(b, float)
    assert a.shape == b.shape, (a.shape, b.shape)
    both_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    both_nan = np.isnan(a) & np.isnan(b)
    skip = both_inf | both_nan
    err = np.abs(a - b)
    lim = atol + rtol * np.abs(b)
    bad = ~skip & ~(err <= lim)
    assert not bad.any(), f"max err {np.nanmax(np.where(skip, 0, err))} at {np.argwhere(bad)[:5].tolist()}"


def has_discrete(params):
    return any((not p.is_cat) and p.step is not None for p in params)


def test_parzen_build_sample_logpdf(eng):
    g = load("parzen.npz")
    for i in range(int(g["n_cases"])):
        t = f"pz{i}/"
        params = decode_space(g[t + "space"])
        mv, clip, endp, pw, C, seed = g[t + "flags"]
        C = int(C)
        obs = g[t + "obs"]
        n = obs.shape[0]
        eng.set_space(specs_from_space(g[t + "space"]))
        eng.set_history(obs, np.zeros(n, np.int8), np.stack([np.arange(n, dtype=float), np.zeros(n)], 1))
        cfg = dict(n_candidates=C, multivariate=bool(mv), prior_weight=float(pw), magic_clip=bool(clip),
                   endpoints=bool(endp))
        cols = list(range(len(params)))
        # everything "below" -> est[0] is the estimator of the golden case
        info = eng.prepare(cols, n_below=n, **cfg)
        assert info == (n, n, 0)
        eng.build()
        w, mu, sg = eng.get_mixture(0)
        close(w, g[t + "w"], 1e-15)
        for j, p in enumerate(params):
            if p.is_cat:
                continue
            close(mu[:, j], g[f"{t}mu{j}"], 1e-15, 1e-300)
            # ties in np.argsort(unstable) only matter for duplicate observations (discrete columns)
            if p.step is None or bool(mv):
                close(sg[:, j], g[f"{t}sigma{j}"], 1e-15)
        ncat, nnum = kinds_of(params)
        u = draw_uniforms(np.random.RandomState(int(seed) + 1000), C, ncat, nnum)
        eng.sample_and_select(u, 1)
        smp, _, _ = eng.get_candidates()
        ref = g[t + "samples"]
        for j, p in enumerate(params):
            if p.is_cat or p.step is not None:
                if p.step is None or bool(mv) or n <= 1:
                    assert np.array_equal(smp[:, j], ref[:, j]), (i, j)
            else:
                close(smp[:, j], ref[:, j], 1e-12, 1e-12)
        tol = 1e-9 if has_discrete(params) else 1e-12
        if bool(mv) or n <= 1:
            close(eng.logpdf(0, ref), g[t + "logpdf"], 0, tol)
            close(eng.logpdf(0, g[t + "samples2"]), g[t + "logpdf2"], 0, tol)


def _run_case(eng, g, t, cols, rng, mv, C, n_below):
    params_all = decode_space(g[t + "space"])
    params = [params_all[c] for c in cols]
    ncat, nnum = kinds_of(params)
    u = draw_uniforms(rng, C, ncat, nnum)
    x, acq, best = eng.suggest(cols, u, 1, n_below=n_below, n_candidates=C, multivariate=mv)
    smp, ll, lg = eng.get_candidates()
    return params, x[0], acq[0], int(best[0]), smp, ll, lg


def test_suggest_against_reference_goldens(eng):
    g = load("suggest.npz")
    for ci in range(int(g["n_cases"])):
        t = f"sg{ci}/"
        mv, C, seed, n_below = g[t + "cfg"]
        mv, C, n_below = bool(mv), int(C), int(n_below)
        X, cat, key = g[t + "X"], g[t + "category"], g[t + "key"]
        P = X.shape[1]
        eng.set_space(specs_from_space(g[t + "space"]))
        eng.set_history(X, cat, key)
        rng = np.random.RandomState(int(seed))
        if mv:
            cols = list(range(P))
            params, x, acq, best, smp, ll, lg = _run_case(eng, g, t, cols, rng, True, C, n_below)
            below, above = eng.get_split()
            ob, keep_b = orc.observations(X, g[t + "below"], cols)
            oa, keep_a = orc.observations(X, g[t + "above"], cols)
            assert np.array_equal(below, g[t + "below"][keep_b])
            assert np.array_equal(above, g[t + "above"][keep_a])
            ref = g[t + "samples"]
            for j, p in enumerate(params):
                if p.is_cat or p.step is not None:
                    assert np.array_equal(smp[:, j], ref[:, j]), (ci, j)
                else:
                    close(smp[:, j], ref[:, j], 1e-12, 1e-12)
            tol = 1e-9 if has_discrete(params) else 1e-12
            close(ll, g[t + "ll"], 0, tol)
            close(lg, g[t + "lg"], 0, tol)
            assert best == int(np.argmax(g[t + "ll"] - g[t + "lg"]))
            want = g[t + "ret_internal"]
            for j, p in enumerate(params):
                if p.is_cat or p.step is not None:
                    assert x[j] == want[j]
                else:
                    close(x[j], want[j], 1e-12, 1e-12)
        else:
            for j in range(P):
                params, x, acq, best, smp, ll, lg = _run_case(eng, g, t, [j], rng, False, C, n_below)
                p = params[0]
                ref = g[f"{t}u{j}/samples"]
                dup = (not p.is_cat) and p.step is not None  # duplicate observations: argsort tie order
                if p.is_cat:
                    assert np.array_equal(smp, ref)
                    close(ll, g[f"{t}u{j}/ll"], 0, 1e-12)
                    close(lg, g[f"{t}u{j}/lg"], 0, 1e-12)
                elif not dup:
                    close(smp, ref, 1e-12, 1e-12)
                    close(ll, g[f"{t}u{j}/ll"], 0, 1e-12)
                    close(lg, g[f"{t}u{j}/lg"], 0, 1e-12)
                    close(x[0], g[t + "ret_internal"][j], 1e-12, 1e-12)


def test_univariate_duplicates_match_stable_oracle(eng):
    """Discrete columns in univariate mode: the reference's bandwidths depend on np.argsort's
    (unstable) tie order; the CUDA path uses the stable order, so compare with the oracle run with
    stable_sort=True on the same inputs."""
    g = load("suggest.npz")
    t = "sg3/"
    mv, C, seed, n_below = g[t + "cfg"]
    C, n_below = int(C), int(n_below)
    X, cat, key = g[t + "X"], g[t + "category"], g[t + "key"]
    params_all = decode_space(g[t + "space"])
    eng.set_space(specs_from_space(g[t + "space"]))
    eng.set_history(X, cat, key)
    cfg = orc.Config(multivariate=False, stable_sort=True)
    for j, p in enumerate(params_all):
        if p.is_cat or p.step is None:
            continue
        rng_o = np.random.RandomState(100 + j)
        s = orc.suggest(X, cat, key, params_all, [j], cfg, n_below, C, rng_o)
        u = draw_uniforms(np.random.RandomState(100 + j), C, 0, 1)
        x, acq, best = eng.suggest([j], u, 1, n_below=n_below, n_candidates=C, multivariate=False)
        smp, ll, lg = eng.get_candidates()
        wb, mub, sgb = eng.get_mixture(0)
        wa, mua, sga = eng.get_mixture(1)
        close(sgb[:, 0], s.mix_below.sigma[0], 1e-15)
        close(sga[:, 0], s.mix_above.sigma[0], 1e-15)
        assert np.array_equal(smp[:, 0], s.samples[:, 0])
        close(ll, s.logl, 0, 1e-9)
        close(lg, s.logg, 0, 1e-9)
        assert int(best[0]) == s.best
        assert x[0, 0] == s.x[0]


def test_split_edge_cases(eng):
    """+-inf values, ties (stable), pruned ordering, infeasible, running (sampler.py:686-821)."""
    from optuna_b200.engine import ParamSpec
    rs = np.random.RandomState(5)
    eng.set_space([ParamSpec(kind=0, low=0.0, high=1.0)])
    for trial in range(30):
        n = int(rs.randint(1, 400))
        cat = rs.choice([0, 0, 0, 1, 2, 3], size=n).astype(np.int8)
        key = np.zeros((n, 2))
        key[:, 0] = rs.choice([-np.inf, np.inf, 0.0, -0.0, 1.0, 2.0, 3.5], size=n) if trial % 2 else rs.normal(size=n)
        pr = cat == 1
        key[pr, 0] = -rs.randint(0, 3, size=pr.sum())
        key[pr, 1] = rs.choice([0.5, 1.5, np.inf], size=pr.sum())
        X = rs.uniform(size=(n, 1))
        eng.set_history(X, cat, key)
        for n_below in (0, 1, n // 3, n, n + 5):
            eng.prepare([0], n_below=n_below, n_candidates=4, multivariate=True)
            below, above = eng.get_split()
            ob, oa = orc.split_trials(cat, key, n_below)
            assert np.array_equal(below, ob), (trial, n_below)
            assert np.array_equal(above, oa), (trial, n_below)


def test_config2_shape_against_chunked_oracle(eng):
    """BASELINE config 2 at full size (N=100k, P=32, C=4096, multivariate): every stage vs the
    oracle, the log-density on a slice of candidates (the oracle needs ~1 s per 4 candidates)."""
    from optuna_b200.engine import ParamSpec
    N, P, C = 100_000, 32, 4096
    rs = np.random.RandomState(0)
    X = rs.uniform(0, 1, (N, P))
    loss = ((X - 0.5) ** 2).sum(1)
    cat = np.zeros(N, np.int8)
    key = np.stack([loss, np.zeros(N)], 1)
    eng.set_space([ParamSpec(kind=0, low=0.0, high=1.0) for _ in range(P)])
    eng.set_history(X, cat, key)
    n_below = orc.default_gamma(N)
    rng = np.random.RandomState(1)
    u = draw_uniforms(rng, C, 0, P)
    x, acq, best = eng.suggest(list(range(P)), u, 1, n_below=n_below, n_candidates=C, multivariate=True)
    assert eng.last_logpdf_kernel().startswith(("k_tcs", "k_logpdf_mma", "k_logpdf_fast", "k_logpdf_screen"))
    smp, ll, lg = eng.get_candidates()
    params = [orc.Param("float", 0.0, 1.0) for _ in range(P)]
    cfg = orc.Config(multivariate=True)
    ob, oa = orc.split_trials(cat, key, n_below)
    below, above = eng.get_split()
    assert np.array_equal(below, ob) and np.array_equal(above, oa)
    mb = orc.build_mixture(X[ob], params, cfg)
    ma = orc.build_mixture(X[oa], params, cfg)
    osmp = orc.mixture_sample(mb, np.random.RandomState(1), C)
    close(smp, osmp, 1e-12, 1e-12)
    wa, mua, sga = eng.get_mixture(1)
    close(wa, ma.weights, 1e-15)
    close(sga[:, 0], ma.sigma[0], 1e-15)
    pick = np.concatenate([np.arange(8), [int(best[0])], rs.randint(0, C, 7)])
    close(ll[pick], orc.mixture_log_pdf(mb, osmp[pick]), 0, 1e-12)
    close(lg[pick], orc.mixture_log_pdf_chunked(ma, osmp[pick], rows=4), 0, 1e-12)
    # size-independent properties on the whole batch
    assert np.all((smp >= 0) & (smp <= 1))
    assert int(best[0]) == int(np.argmax(ll - lg))
    assert np.array_equal(x[0], smp[int(best[0])])
    assert np.isfinite(ll).all() and np.isfinite(lg).all()
    # tpe_logpdf entry point == values produced inside sample_and_select
    close(eng.logpdf(1, smp[:64]), lg[:64], 0, 1e-13)


def test_batched_asks_equal_sequential(eng):
    """n_asks suggestions in one call == the same asks one by one (config 5 semantics)."""
    g = load("suggest.npz")
    t = "sg7/"
    X, cat, key = g[t + "X"], g[t + "category"], g[t + "key"]
    P = X.shape[1]
    eng.set_space(specs_from_space(g[t + "space"]))
    eng.set_history(X, cat, key)
    C, n_asks = 24, 37
    rng = np.random.RandomState(3)
    us = [draw_uniforms(rng, C, 0, P) for _ in range(n_asks)]
    cols = list(range(P))
    xb, ab, bb = eng.suggest(cols, np.concatenate(us), n_asks, n_below=25, n_candidates=C, multivariate=True)
    for a in range(n_asks):
        x1, a1, b1 = eng.suggest(cols, us[a], 1, n_below=25, n_candidates=C, multivariate=True)
        assert np.array_equal(x1[0], xb[a]) and b1[0] == bb[a]
        close(a1[0], ab[a], 0, 1e-13)


def test_error_contract(eng):
    from optuna_b200.engine import ParamSpec
    eng.set_space([ParamSpec(kind=0, low=0.0, high=1.0)])
    eng.set_history(np.random.RandomState(0).uniform(size=(20, 1)), np.zeros(20, np.int8), np.zeros((20, 2)))
    with pytest.raises(ValueError):
        eng.prepare([0], n_below=3, n_candidates=8, multivariate=True, prior_weight=-1.0)
    eng.prepare([0], n_below=3, n_candidates=8, multivariate=True)
    with pytest.raises(ValueError):
        eng.build(w_above=-np.ones(17))
    with pytest.raises(ValueError):
        eng.build(w_above=np.zeros(17))
    with pytest.raises(ValueError):
        eng.set_space([ParamSpec(kind=0, low=2.0, high=1.0)])


@pytest.mark.parametrize("n,C", [(4000, 24), (50_000, 24), (4000, 160), (50_000, 96), (50_000, 4096)])
def test_config3_shape_mixed_64_params_against_oracle(eng, n, C):
    """BASELINE config 3 layout (24 float + 8 log-float + 8 step-float + 8 int + 4 log-int +
    12 categorical, multivariate) at N = 4000 and at the full N = 50 000 (chunked oracle).  From 64 candidates on
    g(x) runs through k_logpdf_mixed (kernel-minor tables, table rows in shared memory); at C = 4096 the oracle is
    replaced by the elementwise kernel k_logpdf_pairs on the same candidates (tpe_logpdf), which the smaller cases
    pin to the oracle."""
    from optuna_b200.engine import ParamSpec
    rs = np.random.RandomState(0)
    specs, params, cols = [], [], []
    for _ in range(24):
        specs.append(ParamSpec(kind=0, low=0.0, high=1.0)); params.append(orc.Param("float", 0.0, 1.0))
        cols.append(rs.uniform(0, 1, n))
    for _ in range(8):
        specs.append(ParamSpec(kind=0, low=1e-5, high=1.0, log=True)); params.append(orc.Param("float", 1e-5, 1.0, None, True))
        cols.append(np.exp(rs.uniform(np.log(1e-5), 0, n)))
    for _ in range(8):
        specs.append(ParamSpec(kind=0, low=0.0, high=10.0, step=0.5)); params.append(orc.Param("float", 0.0, 10.0, 0.5))
        cols.append(rs.randint(0, 21, n) * 0.5)
    for _ in range(8):
        specs.append(ParamSpec(kind=1, low=0, high=100, step=1)); params.append(orc.Param("int", 0.0, 100.0, 1.0))
        cols.append(rs.randint(0, 101, n).astype(float))
    for _ in range(4):
        specs.append(ParamSpec(kind=1, low=1, high=1024, step=1, log=True)); params.append(orc.Param("int", 1.0, 1024.0, 1.0, True))
        cols.append(np.round(np.exp(rs.uniform(0, np.log(1024), n))))
    for k in range(12):
        nch = 4 + k % 5
        specs.append(ParamSpec(kind=2, n_choices=nch)); params.append(orc.Param("cat", n_choices=nch))
        cols.append(rs.randint(0, nch, n).astype(float))
    X = np.stack(cols, 1)
    cat = np.zeros(n, np.int8)
    key = np.stack([rs.normal(size=n), np.zeros(n)], 1)
    eng.set_space(specs)
    eng.set_history(X, cat, key)
    u = draw_uniforms(np.random.RandomState(9), C, 12, 52)
    x, acq, best = eng.suggest(list(range(64)), u, 1, n_below=25, n_candidates=C, multivariate=True)
    assert eng.last_logpdf_kernel() == ("k_logpdf_mixed" if C >= 64 else "k_logpdf_pairs")
    smp, ll, lg = eng.get_candidates()
    if C > 1000:
        lg_pairs = np.concatenate([eng.logpdf(1, smp[i: i + 1024]) for i in range(0, C, 1024)])
        assert eng.last_logpdf_kernel() == "k_logpdf_pairs"
        close(lg, lg_pairs, 0, 1e-9)
        assert int(best[0]) == int(np.argmax(ll - lg))
        return
    s = orc.suggest(X, cat, key, params, list(range(64)), orc.Config(multivariate=True), 25, C,
                    np.random.RandomState(9), chunk_rows=2 if n > 10_000 else None)
    for j, p in enumerate(params):
        if p.is_cat or p.step is not None:
            assert np.array_equal(smp[:, j], s.samples[:, j]), j
        else:
            close(smp[:, j], s.samples[:, j], 1e-12, 1e-12)
    # integer ranges up to 1024 steps: the reference's own _log_diff conditioning
    close(ll, s.logl, 0, 1e-8)
    close(lg, s.logg, 0, 1e-8)
    assert int(best[0]) == s.best
    for j, p in enumerate(params):
        if p.is_cat or p.step is not None:
            assert x[0, j] == s.x[j]


@pytest.mark.parametrize("C,offgrid", [(8, False), (64, False), (64, True)])
def test_tabulated_discrete_columns(eng, C, offgrid):
    """Discrete columns evaluated through the cell-mass tables (candidate-indexed when C < grid size,
    grid-indexed otherwise) and through the direct formula when an observation is off the grid."""
    from optuna_b200.engine import ParamSpec
    rs = np.random.RandomState(5)
    n = 700
    specs = [ParamSpec(kind=0, low=-1.0, high=2.0, step=0.25), ParamSpec(kind=1, low=2, high=40, step=2),
             ParamSpec(kind=1, low=1, high=300, step=1, log=True), ParamSpec(kind=0, low=0.0, high=1.0)]
    params = [orc.Param("float", -1.0, 2.0, 0.25), orc.Param("int", 2.0, 40.0, 2.0),
              orc.Param("int", 1.0, 300.0, 1.0, True), orc.Param("float", 0.0, 1.0)]
    X = np.stack([rs.randint(0, 13, n) * 0.25 - 1.0, rs.randint(1, 21, n) * 2.0,
                  np.round(np.exp(rs.uniform(0, np.log(300), n))), rs.uniform(0, 1, n)], 1)
    if offgrid:
        X[17, 0] = 0.3
        X[400, 1] = 7.0
    cat = np.zeros(n, np.int8)
    key = np.stack([rs.normal(size=n), np.zeros(n)], 1)
    eng.set_space(specs)
    eng.set_history(X, cat, key)
    u = draw_uniforms(np.random.RandomState(3), C, 0, 4)
    x, acq, best = eng.suggest([0, 1, 2, 3], u, 1, n_below=40, n_candidates=C, multivariate=True)
    smp, ll, lg = eng.get_candidates()
    s = orc.suggest(X, cat, key, params, [0, 1, 2, 3], orc.Config(multivariate=True), 40, C,
                    np.random.RandomState(3))
    for j in range(3):
        assert np.array_equal(smp[:, j], s.samples[:, j]), j
    close(smp[:, 3], s.samples[:, 3], 1e-12, 1e-12)
    close(ll, s.logl, 0, 1e-9)
    close(lg, s.logg, 0, 1e-9)
    assert int(best[0]) == s.best
    # user points off the grid: direct formula (grid-indexed tables) or their own row (candidate-indexed)
    pts = s.samples.copy()
    pts[::3, 0] += 0.01
    close(eng.logpdf(1, pts), orc.mixture_log_pdf(s.mix_above, pts), 0, 1e-9)
    close(eng.logpdf(0, pts), orc.mixture_log_pdf(s.mix_below, pts), 0, 1e-9)


@pytest.mark.parametrize("P,n,C,n_below", [(5, 3, 8, 1), (8, 40, 100, 10), (12, 700, 24, 25), (20, 2500, 70, 40),
                                            (33, 900, 130, 25), (64, 300, 24, 12), (9, 9, 300, 8),
                                            # 256 candidates and more (with TPE_TCS=1: the bf16-screened grid k_tcs, P padded to 16 / 32 / 64)
                                            (12, 3000, 300, 25), (16, 129, 256, 10), (27, 6000, 700, 25),
                                            (32, 1500, 1030, 30), (40, 2100, 520, 25), (64, 5000, 384, 25)])
def test_tensor_core_kernel_shapes(eng, P, n, C, n_below):
    """Multivariate, all-continuous spaces go through k_logpdf_mma (P >= 5): ragged P (padding of the
    k-steps), fewer kernels than one group of 8, padded candidate tiles, log-scaled columns, and the
    magic-clip off so that sigma is not clamped.  log_pdf against the oracle at 1e-12."""
    from optuna_b200.engine import ParamSpec
    rs = np.random.RandomState(100 + P)
    specs, params, cols = [], [], []
    for j in range(P):
        if j % 4 == 3:
            specs.append(ParamSpec(kind=0, low=1e-3, high=50.0, log=True))
            params.append(orc.Param("float", 1e-3, 50.0, None, True))
            cols.append(np.exp(rs.uniform(np.log(1e-3), np.log(50.0), n)))
        else:
            lo, hi = -2.0 - j, 3.0 + 0.5 * j
            specs.append(ParamSpec(kind=0, low=lo, high=hi))
            params.append(orc.Param("float", lo, hi))
            cols.append(rs.uniform(lo, hi, n))
    X = np.stack(cols, 1)
    cat = np.zeros(n, np.int8)
    key = np.stack([rs.normal(size=n), np.zeros(n)], 1)
    eng.set_space(specs)
    eng.set_history(X, cat, key)
    for clip in (True, False):
        u = draw_uniforms(np.random.RandomState(7), C, 0, P)
        x, acq, best = eng.suggest(list(range(P)), u, 1, n_below=n_below, n_candidates=C, multivariate=True,
                                   magic_clip=clip)
        if P <= 64:
            assert eng.last_logpdf_kernel().startswith(("k_tcs", "k_logpdf_mma", "k_logpdf_fast")), eng.last_logpdf_kernel()
        if C >= 256 and P >= 9 and clip and n >= 1100 and os.environ.get("TPE_TCS") == "1":
            assert eng.last_logpdf_kernel().startswith("k_tcs"), eng.last_logpdf_kernel()
        smp, ll, lg = eng.get_candidates()
        s = orc.suggest(X, cat, key, params, list(range(P)), orc.Config(multivariate=True, magic_clip=clip),
                        n_below, C, np.random.RandomState(7))
        close(smp, s.samples, 1e-12, 1e-12)
        close(ll, s.logl, 1e-14, 1e-12)
        close(lg, s.logg, 1e-14, 1e-12)
        assert int(best[0]) == s.best


def test_bf16_screened_grid_variant_passes_the_same_parity_tests():
    """TPE_TCS=1 routes the multivariate grid of large all-continuous estimators through k_tcs (bf16 tensor-core
    screen with a rigorous rounding bound + exact fp64 evaluation of the survivors, tpe_tcscreen.cuh) -- slower than
    the fp64 tensor-core kernel and therefore off by default, but kept correct: the shape sweep, the full-size
    config-2 fixture and the massive-ties case run again in a process with the switch on."""
    if os.environ.get("TPE_TCS") == "1":
        pytest.skip("already inside the TPE_TCS=1 run")
    env = dict(os.environ, TPE_TCS="1")
    r = subprocess.run([sys.executable, "-m", "pytest", __file__, "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider", "-k",
                        "tensor_core_kernel_shapes or precomputed_oracle_fixture or massive_ties or config2_shape"],
                       env=env, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout


def test_categorical_distance_func_tables(eng):
    """categorical_distance_func (parzen_estimator.py:152-160): rows exp(-(d / max d)^2 * coef)."""
    from optuna_b200.engine import ParamSpec
    rs = np.random.RandomState(4)
    nch, n, C = 5, 300, 32
    table = np.abs(np.subtract.outer(np.arange(nch), np.arange(nch))).astype(float) + 0.25 * (1 - np.eye(nch))
    eng.set_space([ParamSpec(kind=2, n_choices=nch, dist_table=table), ParamSpec(kind=0, low=-1.0, high=1.0)])
    X = np.stack([rs.randint(0, nch, n).astype(float), rs.uniform(-1, 1, n)], 1)
    cat = np.zeros(n, np.int8)
    key = np.stack([rs.normal(size=n), np.zeros(n)], 1)
    eng.set_history(X, cat, key)
    params = [orc.Param("cat", n_choices=nch, dist_table=table), orc.Param("float", -1.0, 1.0)]
    for mv in (True, False):
        calls = [[0, 1]] if mv else [[0], [1]]
        for cols in calls:
            sub = [params[c] for c in cols]
            ncat, nnum = kinds_of(sub)
            u = draw_uniforms(np.random.RandomState(6), C, ncat, nnum)
            x, acq, best = eng.suggest(cols, u, 1, n_below=20, n_candidates=C, multivariate=mv)
            smp, ll, lg = eng.get_candidates()
            s = orc.suggest(X, cat, key, params, cols, orc.Config(multivariate=mv), 20, C, np.random.RandomState(6))
            for jj, c in enumerate(cols):
                if params[c].is_cat:
                    assert np.array_equal(smp[:, jj], s.samples[:, jj])
                else:
                    close(smp[:, jj], s.samples[:, jj], 1e-12, 1e-12)
            close(ll, s.logl, 0, 1e-12)
            close(lg, s.logg, 0, 1e-12)
            assert int(best[0]) == s.best


def test_univariate_large_history_radix_sorted_bandwidths(eng):
    """Univariate bandwidths above 4096 kernels go through the cooperative radix sort; compare
    sigma / samples / log_pdf with the oracle (stable order, so duplicates are well defined)."""
    from optuna_b200.engine import ParamSpec
    rs = np.random.RandomState(8)
    n, C = 9000, 64
    specs = [ParamSpec(kind=0, low=-2.0, high=3.0), ParamSpec(kind=0, low=1e-3, high=10.0, log=True),
             ParamSpec(kind=1, low=0, high=40, step=1)]
    params = [orc.Param("float", -2.0, 3.0), orc.Param("float", 1e-3, 10.0, None, True), orc.Param("int", 0.0, 40.0, 1.0)]
    X = np.stack([rs.uniform(-2, 3, n), np.exp(rs.uniform(np.log(1e-3), np.log(10), n)),
                  rs.randint(0, 41, n).astype(float)], 1)
    X[:50, 0] = X[50:100, 0]  # exact duplicates in a continuous column too
    cat = np.zeros(n, np.int8)
    key = np.stack([rs.normal(size=n), np.zeros(n)], 1)
    eng.set_space(specs)
    eng.set_history(X, cat, key)
    for magic_clip in (True, False):
        cfg = orc.Config(multivariate=False, stable_sort=True, magic_clip=magic_clip)
        for j in range(3):
            u = draw_uniforms(np.random.RandomState(40 + j), C, 0, 1)
            x, acq, best = eng.suggest([j], u, 1, n_below=25, n_candidates=C, multivariate=False,
                                       magic_clip=magic_clip)
            smp, ll, lg = eng.get_candidates()
            s = orc.suggest(X, cat, key, params, [j], cfg, 25, C, np.random.RandomState(40 + j))
            wa, mua, sga = eng.get_mixture(1)
            # gaps of log-transformed values inherit the 1-ulp difference between device and NumPy log()
            close(sga[:, 0], s.mix_above.sigma[0], 1e-15, 2e-15 if params[j].log else 1e-300)
            if params[j].step is None:
                close(smp[:, 0], s.samples[:, 0], 1e-12, 1e-12)
                tol = 1e-12 if magic_clip else 1e-9  # without the clip sigma reaches 1e-12 * range
            else:
                assert np.array_equal(smp[:, 0], s.samples[:, 0])
                tol = 1e-9
            close(ll, s.logl, 0, tol)
            close(lg, s.logg, 0, tol)
            assert int(best[0]) == s.best


@pytest.mark.parametrize("seed,pre,skip,count", [(0, 0, 0, 1000), (1, 7, 0, 311), (2, 623, 5, 313), (3, 100, 4096, 70000),
                                                 (4, 1249, 0, 1), (5, 311, 311, 312), (6, 0, 0, 135168)])
def test_device_mt19937_is_numpys_stream(eng, seed, pre, skip, count):
    """k_mt19937_uniform reproduces RandomState.random_sample bit for bit (state blocks, odd positions,
    doubles straddling a block boundary, dropped prefix) and returns the generator's end state."""
    a, b = np.random.RandomState(seed), np.random.RandomState(seed)
    if pre:
        a.randint(0, 2 ** 31 - 1, size=pre)  # one 32-bit word each: odd positions inside a block
        b.randint(0, 2 ** 31 - 1, size=pre)
    assert np.array_equal(a.get_state()[1], b.get_state()[1]) and a.get_state()[2] == b.get_state()[2]
    eng.stage_rng(a, count, skip)
    got = eng.get_uniforms(count)
    want = b.random_sample(skip + count)[skip:]
    assert np.array_equal(got, want)
    eng.finish_rng(a)
    sa, sb = a.get_state(), b.get_state()
    assert sa[2] == sb[2] and np.array_equal(sa[1], sb[1])
    assert np.array_equal(a.random_sample(5), b.random_sample(5))


def test_sharded_asks_with_device_uniforms_equal_the_full_batch(eng):
    """Config-5 semantics with device-generated uniforms: every "rank" starts from the same generator
    state, drops the uniforms of the earlier ranks' asks (skip) and evaluates its block; the blocks
    concatenate to the result of one batch drawn on the host."""
    from optuna_b200.dist import shard_asks
    from optuna_b200.engine import ParamSpec
    rs = np.random.RandomState(4)
    n, P, C, n_asks = 500, 6, 700, 7          # 700 * 7 = 4900 uniforms per ask
    X = rs.uniform(0, 1, (n, P))
    key = np.stack([((X - 0.3) ** 2).sum(1), np.zeros(n)], 1)
    eng.set_space([ParamSpec(kind=0, low=0.0, high=1.0) for _ in range(P)])
    eng.set_history(X, np.zeros(n, np.int8), key)
    cfg = dict(n_below=30, n_candidates=C, multivariate=True)
    per_ask = C * (1 + P)
    full_rng = np.random.RandomState(77)
    want, _, _ = eng.suggest(list(range(P)), full_rng.random_sample(n_asks * per_ask), n_asks, **cfg)
    got = []
    for world in (3,):
        for rank in range(world):
            start, count = shard_asks(n_asks, world, rank)
            r = np.random.RandomState(77)
            eng.prepare(list(range(P)), **cfg)
            eng.build()
            eng.stage_rng(r, count * per_ask, skip=start * per_ask)
            x, _, _ = eng.sample_and_select(None, count)
            eng.finish_rng(r)
            got.append(x)
            tail = (n_asks - start - count) * per_ask
            if tail:
                eng.stage_rng(r, 1, skip=tail - 1)
                eng.finish_rng(r)
            assert np.array_equal(r.random_sample(3), np.random.RandomState(77).random_sample(n_asks * per_ask + 3)[-3:])
    assert np.array_equal(np.concatenate(got), want)


def test_staged_and_pinned_uniforms_change_nothing(eng):
    """tpe_stage_uniforms / tpe_host_alloc are latency hints: same suggestion as the plain call."""
    from optuna_b200.engine import ParamSpec
    rs = np.random.RandomState(8)
    n, P, C = 400, 5, 64
    X = rs.uniform(-1, 1, (n, P))
    key = np.stack([(X ** 2).sum(1), np.zeros(n)], 1)
    eng.set_space([ParamSpec(kind=0, low=-1.0, high=1.0) for _ in range(P)])
    eng.set_history(X, np.zeros(n, np.int8), key)
    cfg = dict(n_below=20, n_candidates=C, multivariate=True)
    u = np.random.RandomState(1).random_sample(C * (1 + P))
    want, wacq, wbest = eng.suggest(list(range(P)), u, 1, **cfg)
    pinned = eng.pinned_empty(u.size)
    pinned[:] = u
    eng.prepare(list(range(P)), **cfg)
    staged = eng.stage_uniforms(pinned)
    eng.build()
    x, acq, best = eng.sample_and_select(staged, 1)
    assert np.array_equal(x, want) and np.array_equal(acq, wacq) and np.array_equal(best, wbest)
    # staging one buffer and passing another must fall back to a fresh upload
    eng.prepare(list(range(P)), **cfg)
    eng.stage_uniforms(np.zeros(u.size))
    eng.build()
    x2, _, _ = eng.sample_and_select(u, 1)
    assert np.array_equal(x2, want)


def _random_space(rs, P, kinds):
    from optuna_b200.engine import ParamSpec
    specs, params, draw = [], [], []
    for j in range(P):
        kind = kinds[rs.randint(len(kinds))]
        if kind == "float":
            lo = float(rs.uniform(-5, 0)); hi = lo + float(rs.uniform(0.5, 8))
            specs.append(ParamSpec(kind=0, low=lo, high=hi)); params.append(orc.Param("float", lo, hi))
            draw.append(lambda n, lo=lo, hi=hi: rs.uniform(lo, hi, n))
        elif kind == "logfloat":
            specs.append(ParamSpec(kind=0, low=1e-4, high=10.0, log=True)); params.append(orc.Param("float", 1e-4, 10.0, None, True))
            draw.append(lambda n: np.exp(rs.uniform(np.log(1e-4), np.log(10.0), n)))
        elif kind == "stepfloat":
            specs.append(ParamSpec(kind=0, low=0.0, high=3.0, step=0.25)); params.append(orc.Param("float", 0.0, 3.0, 0.25))
            draw.append(lambda n: rs.randint(0, 13, n) * 0.25)
        elif kind == "int":
            specs.append(ParamSpec(kind=1, low=-3, high=12, step=1)); params.append(orc.Param("int", -3.0, 12.0, 1.0))
            draw.append(lambda n: rs.randint(-3, 13, n).astype(float))
        elif kind == "logint":
            specs.append(ParamSpec(kind=1, low=1, high=200, step=1, log=True)); params.append(orc.Param("int", 1.0, 200.0, 1.0, True))
            draw.append(lambda n: np.round(np.exp(rs.uniform(0, np.log(200), n))))
        else:
            nch = int(rs.randint(2, 7))
            specs.append(ParamSpec(kind=2, n_choices=nch)); params.append(orc.Param("cat", n_choices=nch))
            draw.append(lambda n, nch=nch: rs.randint(0, nch, n).astype(float))
    return specs, params, draw


@pytest.mark.parametrize("case", range(16))
def test_random_spaces_against_the_oracle(eng, case):
    """Differential test on random spaces: mixed parameter kinds, COMPLETE / PRUNED / infeasible / RUNNING
    trials, rows lacking a parameter, small and large estimators (the large ones take the 8-candidate
    tiling of the pair kernel, all-continuous ones the grid kernels), uni- and multivariate."""
    rs = np.random.RandomState(1000 + case)
    mv = case % 4 != 3
    P = int(rs.randint(1, 9)) if mv else 1
    kinds = [["float", "logfloat"], ["float", "int", "cat"], ["float", "logfloat", "stepfloat", "int", "logint", "cat"],
             ["float"]][case % 4]
    n = [60, 900, 5000, 300][case % 4]
    C = int(rs.choice([8, 24, 100]))
    specs, params, draw = _random_space(rs, P, kinds)
    X = np.stack([d(n) for d in draw], 1)
    if case % 5 == 1 and P > 1:     # some rows lack a parameter
        X[rs.uniform(size=n) < 0.1, rs.randint(P)] = np.nan
    cat = np.zeros(n, np.int8)
    if case % 3 == 0:
        cat = rs.choice([0, 0, 0, 1, 2, 3], size=n).astype(np.int8)
    key = np.stack([rs.normal(size=n), rs.normal(size=n) * (cat == 1)], 1)
    key[cat == 2, 0] = np.abs(key[cat == 2, 0]) + 0.1
    n_below = int(rs.randint(1, min(40, n)))
    eng.set_space(specs)
    eng.set_history(X, cat, key)
    cols = list(range(P))
    ncat = sum(p.is_cat for p in params)
    u = draw_uniforms(np.random.RandomState(case), C, ncat, P - ncat)
    cfg = dict(n_below=n_below, n_candidates=C, multivariate=mv)
    x, acq, best = eng.suggest(cols, u, 1, **cfg)
    smp, ll, lg = eng.get_candidates()
    s = orc.suggest(X, cat, key, params, cols, orc.Config(multivariate=mv, stable_sort=True), n_below, C,
                    np.random.RandomState(case))
    below, above = eng.get_split()
    ob, keep_b = orc.observations(X, s.below, cols)
    oa, keep_a = orc.observations(X, s.above, cols)
    assert np.array_equal(below, s.below[keep_b]) and np.array_equal(above, s.above[keep_a])
    for j, p in enumerate(params):
        if p.is_cat or p.step is not None:
            assert np.array_equal(smp[:, j], s.samples[:, j]), j
        else:
            close(smp[:, j], s.samples[:, j], 1e-12, 1e-12)
    tol = 1e-9 if has_discrete(params) else 1e-12
    close(ll, s.logl, 1e-14, tol)
    close(lg, s.logg, 1e-14, tol)
    assert int(best[0]) == s.best


def test_tensor_core_kernel_with_massive_ties(eng):
    """Thousands of identical observations: every kernel term of g(x) is within a few units of the max
    (the exact "near" tier sees all of them, the fp32 "far" tier none), plus a cluster far away that is
    dropped entirely.  log_pdf must still match the oracle at 1e-12."""
    from optuna_b200.engine import ParamSpec
    rs = np.random.RandomState(12)
    P, C = 8, 96
    base = rs.uniform(0.2, 0.8, (5, P))
    X = np.concatenate([np.repeat(base, 600, axis=0), np.repeat(rs.uniform(0.0, 0.05, (1, P)), 200, axis=0)])
    n = X.shape[0]
    key = np.stack([((X - 0.5) ** 2).sum(1) + 1e-9 * np.arange(n), np.zeros(n)], 1)
    specs = [ParamSpec(kind=0, low=0.0, high=1.0) for _ in range(P)]
    params = [orc.Param("float", 0.0, 1.0) for _ in range(P)]
    eng.set_space(specs)
    eng.set_history(X, np.zeros(n, np.int8), key)
    u = draw_uniforms(np.random.RandomState(2), C, 0, P)
    x, acq, best = eng.suggest(list(range(P)), u, 1, n_below=20, n_candidates=C, multivariate=True)
    assert eng.last_logpdf_kernel().startswith(("k_tcs", "k_logpdf_mma"))
    smp, ll, lg = eng.get_candidates()
    s = orc.suggest(X, np.zeros(n, np.int8), key, params, list(range(P)), orc.Config(multivariate=True), 20, C,
                    np.random.RandomState(2))
    close(smp, s.samples, 1e-12, 1e-12)
    close(ll, s.logl, 1e-14, 1e-12)
    close(lg, s.logg, 1e-14, 1e-12)
    assert int(best[0]) == s.best


def test_univariate_batch_equals_the_per_parameter_calls(eng):
    """tpe_suggest_univariate_batch (the P sample_independent calls of one trial on P concurrent column contexts)
    against P sequential single-column tpe_suggest calls on the same stretch of uniforms: bit-identical, at a small
    size with every numeric kind + a categorical, and at config 2's full size (N = 100 000 x P = 32, C = 4096)."""
    from optuna_b200.engine import ParamSpec
    rs = np.random.RandomState(12)
    specs = [ParamSpec(kind=0, low=-1.0, high=2.0), ParamSpec(kind=0, low=1e-3, high=10.0, log=True),
             ParamSpec(kind=0, low=0.0, high=3.0, step=0.25), ParamSpec(kind=1, low=0, high=20, step=1),
             ParamSpec(kind=2, n_choices=4), ParamSpec(kind=1, low=1, high=512, step=1, log=True)]
    n = 3000
    X = np.stack([rs.uniform(-1, 2, n), np.exp(rs.uniform(np.log(1e-3), np.log(10), n)),
                  rs.randint(0, 13, n) * 0.25, rs.randint(0, 21, n).astype(float), rs.randint(0, 4, n).astype(float),
                  np.round(np.exp(rs.uniform(0, np.log(512), n)))], 1)
    key = np.stack([rs.normal(size=n), np.zeros(n)], 1)
    for specs_i, X_i, key_i, C in ((specs, X, key, 24), (None, None, None, 4096)):
        if specs_i is None:
            N, P = 100_000, 32
            r2 = np.random.RandomState(0)
            X_i = r2.uniform(0, 1, (N, P))
            key_i = np.stack([((X_i - 0.5) ** 2).sum(1), np.zeros(N)], 1)
            specs_i = [ParamSpec(kind=0, low=0.0, high=1.0) for _ in range(P)]
        P = len(specs_i)
        eng.set_space(specs_i)
        eng.set_history(X_i, np.zeros(len(X_i), np.int8), key_i)
        cfg = dict(n_below=25, n_candidates=C, multivariate=False)
        u = np.random.RandomState(5).random_sample(P * 2 * C)
        xb, ab, bb = eng.suggest_univariate_batch(list(range(P)), u, **cfg)
        for j in range(P):
            x, acq, best = eng.suggest([j], u[j * 2 * C: (j + 1) * 2 * C], 1, **cfg)
            assert x[0, 0] == xb[j] and best[0] == bb[j] and acq[0] == ab[j], (C, j)
        # device-generated uniforms: the same stream
        r = np.random.RandomState(5)
        eng.stage_rng(r, P * 2 * C)
        xd, _, _ = eng.suggest_univariate_batch(list(range(P)), None, **cfg)
        assert np.array_equal(xd, xb)
        eng.finish_rng(r)
        assert np.array_equal(r.random_sample(3), np.random.RandomState(5).random_sample(P * 2 * C + 3)[-3:])
    # a parameter absent from some trials: the columns cannot share a split
    Xm = X.copy()
    Xm[5, 2] = np.nan
    eng.set_space(specs)
    eng.set_history(Xm, np.zeros(n, np.int8), key)
    with pytest.raises(RuntimeError, match="not batchable"):
        eng.suggest_univariate_batch([0, 2], np.zeros(2 * 2 * 24), n_below=25, n_candidates=24, multivariate=False)


@pytest.mark.parametrize("all_continuous", [False, True])
def test_univariate_batch_incremental_orders_equal_fresh_sorts(eng, all_continuous):
    """Between two univariate batch calls the column contexts keep each column's sorted order and bring it up to date
    (identical rows: reuse; one trial appended: insert) instead of sorting again.  A long-lived engine fed one trial
    at a time must answer bit-identically to a fresh engine that sorts from scratch, through every kind of change:
    plain appends (continuous columns and step columns full of ties), a trial that enters the below set and pushes
    another one into the above set, repeated calls without a change, a changed column list, a RUNNING row that
    completes later, and an overwritten visible row."""
    from optuna_b200 import TPEEngine
    from optuna_b200.engine import ParamSpec
    rs = np.random.RandomState(77)
    # all_continuous: the staged path (one launch per stage over all columns, tpe_unib.cuh); otherwise one column
    # context per column (discrete columns take the general kernels)
    if all_continuous:
        specs = [ParamSpec(kind=0, low=-1.0, high=2.0), ParamSpec(kind=0, low=0.0, high=3.0),
                 ParamSpec(kind=0, low=0.0, high=20.0), ParamSpec(kind=0, low=1e-3, high=10.0, log=True)]
    else:
        specs = [ParamSpec(kind=0, low=-1.0, high=2.0), ParamSpec(kind=0, low=0.0, high=3.0, step=0.25),
                 ParamSpec(kind=1, low=0, high=20, step=1), ParamSpec(kind=0, low=1e-3, high=10.0, log=True)]
    P, n0, C = len(specs), 6000, 64

    def draw(n):
        if all_continuous:   # (the third column repeats values: ties in a continuous column)
            return np.stack([rs.uniform(-1, 2, n), rs.uniform(0, 3, n), rs.randint(0, 21, n).astype(float),
                             np.exp(rs.uniform(np.log(1e-3), np.log(10), n))], 1)
        return np.stack([rs.uniform(-1, 2, n), rs.randint(0, 13, n) * 0.25, rs.randint(0, 21, n).astype(float),
                         np.exp(rs.uniform(np.log(1e-3), np.log(10), n))], 1)

    X = draw(n0)
    cat = np.zeros(n0, np.int8)
    key = np.stack([rs.uniform(1, 2, n0), np.zeros(n0)], 1)
    cfg = dict(n_below=25, n_candidates=C, multivariate=False)
    eng.set_space(specs)
    eng.set_history(X, cat, key)
    fresh = TPEEngine(0)
    step = [0]

    def check(cols):
        step[0] += 1
        u = np.random.RandomState(1000 + step[0]).random_sample(len(cols) * 2 * C)
        got = eng.suggest_univariate_batch(cols, u, **cfg)
        fresh.set_space(specs)
        fresh.set_history(X, cat, key)
        want = fresh.suggest_univariate_batch(cols, u, **cfg)
        for g, w in zip(got, want):
            assert np.array_equal(g, w), (step[0], g, w)

    try:
        allc = list(range(P))
        check(allc)
        check(allc)                                     # nothing changed: the orders are reused
        for i in range(12):                             # plain appends (worse than the below set)
            x1 = draw(1)
            k1 = np.array([[rs.uniform(1, 2), 0.0]])
            eng.append_history(x1, np.zeros(1, np.int8), k1)
            X, cat, key = np.concatenate([X, x1]), np.concatenate([cat, np.zeros(1, np.int8)]), np.concatenate([key, k1])
            check(allc)
        x1, k1 = draw(1), np.array([[0.5, 0.0]])          # enters the below set; another trial moves to the above set
        eng.append_history(x1, np.zeros(1, np.int8), k1)
        X, cat, key = np.concatenate([X, x1]), np.concatenate([cat, np.zeros(1, np.int8)]), np.concatenate([key, k1])
        check(allc)
        check([0, 2])                                   # fewer columns ...
        x1, k1 = draw(1), np.array([[1.5, 0.0]])
        eng.append_history(x1, np.zeros(1, np.int8), k1)
        X, cat, key = np.concatenate([X, x1]), np.concatenate([cat, np.zeros(1, np.int8)]), np.concatenate([key, k1])
        check([2, 0, 3])                                # ... then other columns in other slots, one of them stale
        check(allc)
        # a RUNNING placeholder (excluded), a later trial finishing before it, then the placeholder completing
        r_run = len(cat)
        eng.append_history(np.full((1, P), np.nan), np.full(1, 4, np.int8), np.zeros((1, 2)))
        X, cat, key = np.concatenate([X, np.full((1, P), np.nan)]), np.concatenate([cat, np.full(1, 4, np.int8)]), \
            np.concatenate([key, np.zeros((1, 2))])
        check(allc)
        x1, k1 = draw(1), np.array([[1.7, 0.0]])
        eng.append_history(x1, np.zeros(1, np.int8), k1)
        X, cat, key = np.concatenate([X, x1]), np.concatenate([cat, np.zeros(1, np.int8)]), np.concatenate([key, k1])
        check(allc)
        x1, k1 = draw(1), np.array([[1.2, 0.0]])          # the placeholder completes: a row in the MIDDLE appears
        eng.update_history(x1, np.zeros(1, np.int8), k1, r_run)
        X[r_run], cat[r_run], key[r_run] = x1[0], 0, k1[0]
        check(allc)
        x1, k1 = draw(1), np.array([[1.9, 0.0]])          # a visible row is overwritten: same sizes, other values
        eng.update_history(x1, np.zeros(1, np.int8), k1, 100)
        X[100], key[100] = x1[0], k1[0]
        check(allc)
        check(allc)
    finally:
        fresh.close()


@pytest.mark.parametrize("layout", ["dense", "clusters", "corner", "log"])
def test_univariate_fast_gauss_transform_of_the_floor_bandwidth_kernels(eng, layout):
    """Large 1-D estimators: the kernels whose bandwidth is the clip floor are summed by a fast Gauss transform
    (k_fgt_coeff / k_fgt_eval), the others pair by pair (k_uni_grid).  log g at points all over the support --
    between the observations, in gaps, far from every observation -- against the oracle: 1e-12, the float
    tolerance of BASELINE.md section 3.  Layouts: every kernel at the floor; clusters with gaps (both kinds of
    kernels, boxes without kernels); all observations in one corner (far boxes that still count, the prior
    kernel taking over); a log-scaled column."""
    from optuna_b200.engine import ParamSpec
    rs = np.random.RandomState({"dense": 1, "clusters": 2, "corner": 3, "log": 4}[layout])
    n = 12_000
    lo, hi, log = (-3.0, 5.0, False) if layout != "log" else (1e-4, 50.0, True)
    if layout == "dense":
        x = rs.uniform(lo, hi, n)
    elif layout == "clusters":
        x = np.concatenate([rs.normal(-2.0, 0.05, n // 2), rs.normal(1.0, 0.3, n // 2 - 40), rs.uniform(2.5, 5.0, 40)])
        x = np.clip(x, lo, hi)
    elif layout == "corner":
        x = lo + (hi - lo) * 0.04 * rs.beta(2, 5, n)
    else:
        x = np.exp(rs.uniform(np.log(lo), np.log(hi), n))
    key = np.stack([rs.uniform(size=n), np.zeros(n)], 1)
    cat = np.zeros(n, np.int8)
    eng.set_space([ParamSpec(kind=0, low=lo, high=hi, log=log)])
    eng.set_history(x[:, None], cat, key)
    prm = [orc.Param("float", lo, hi, None, log)]
    pts = np.linspace(lo, hi, 1500) if not log else np.exp(np.linspace(np.log(lo), np.log(hi), 1500))
    pts = np.concatenate([pts, x[:300], [lo, hi]])[:, None]
    for prior_weight in (1.0, 1e-9):
        cfg = orc.Config(multivariate=False, stable_sort=True, prior_weight=prior_weight)
        eng.prepare([0], n_below=25, n_candidates=24, multivariate=False, prior_weight=prior_weight)
        eng.build()
        _, above = eng.get_split()
        mix = orc.build_mixture(x[above][:, None], prm, cfg, None)
        want = orc.mixture_log_pdf(mix, pts)
        for lo_i in range(0, len(pts), 4096):
            got = eng.logpdf(1, pts[lo_i: lo_i + 4096])
            close(got, want[lo_i: lo_i + 4096], 0, 1e-12)
        assert "k_fgt_eval" in eng.last_logpdf_kernel()
        _, _, sg = eng.get_mixture(1)
        floor = (np.log(hi) - np.log(lo) if log else hi - lo) / 100.0
        at_floor = np.mean(np.isclose(sg[:-1, 0], floor, rtol=1e-14, atol=0))
        assert (at_floor > 0.99) if layout in ("dense", "corner", "log") else (0.5 < at_floor < 0.999), at_floor


def test_asynchronous_entry_points_equal_the_blocking_ones(eng):
    """tpe_sample_and_select_async / tpe_collect and tpe_suggest_univariate_batch_async / tpe_collect_univariate queue
    the work and return; the results are those of the blocking calls bit for bit -- also when the caller uploads
    history rows while the suggestion is still queued (a few rows travel through page-locked staging without
    waiting; they land behind the queued kernels and change nothing those read)."""
    from optuna_b200 import TPEEngine
    from optuna_b200.engine import ParamSpec
    rs = np.random.RandomState(21)
    P, n, C = 6, 5000, 512
    specs = [ParamSpec(kind=0, low=0.0, high=1.0) for _ in range(P - 1)] + [ParamSpec(kind=0, low=1e-3, high=10.0, log=True)]
    X = np.concatenate([rs.uniform(0, 1, (n, P - 1)), np.exp(rs.uniform(np.log(1e-3), np.log(10), (n, 1)))], 1)
    key = np.stack([rs.uniform(size=n), np.zeros(n)], 1)
    one = np.zeros(1, np.int8)
    other = TPEEngine(0)
    try:
        for e in (eng, other):
            e.set_space(specs)
            e.set_history(X, np.zeros(n, np.int8), key)
        cols = list(range(P))
        for step, mv in enumerate((True, False)):
            cfg = dict(n_below=25, n_candidates=C, multivariate=mv)
            u = rs.random_sample(C * (1 + P))
            want = other.suggest(cols, u, 1, **cfg)
            eng.prepare(cols, **cfg)
            eng.build()
            eng.sample_and_select_async(u, 1)
            # rows uploaded while the suggestion is queued: a new trial, and row 3 rewritten with its own content
            row = np.concatenate([rs.uniform(0.1, 0.9, (1, P - 1)), [[0.5]]], 1)
            k1 = np.array([[2.0 + step, 0.0]])
            eng.append_history(row, one, k1)
            eng.update_history(X[3:4], one, key[3:4], 3)
            got = eng.collect()
            for g, w in zip(got, want):
                assert np.array_equal(g, w)
            other.append_history(row, one, k1)
        # univariate batch (all parameters continuous: the path that runs stage by stage)
        cfg = dict(n_below=25, n_candidates=C, multivariate=False)
        for rep in range(3):
            u = rs.random_sample(P * 2 * C)
            want = other.suggest_univariate_batch(cols, u, **cfg)
            eng.suggest_univariate_batch_async(cols, u, **cfg)
            row = np.concatenate([rs.uniform(0.1, 0.9, (1, P - 1)), [[0.7]]], 1)
            k1 = np.array([[5.0 + rep, 0.0]])
            eng.append_history(row, one, k1)
            got = eng.collect_univariate()
            for g, w in zip(got, want):
                assert np.array_equal(g, w)
            other.append_history(row, one, k1)
        with pytest.raises(RuntimeError, match="pending"):
            eng.collect()
        with pytest.raises(RuntimeError, match="pending"):
            eng.collect_univariate()
    finally:
        other.close()


@pytest.mark.parametrize("world,n,P,C", [(2, 6000, 12, 512), (3, 5000, 32, 300), (4, 300, 9, 64), (2, 100_000, 32, 4096)])
def test_kernel_sharded_suggestion_equals_the_single_context_one(eng, world, n, P, C):
    """tpe_set_kernel_shard / tpe_sample_and_partial / tpe_finish_from_partials: `world` contexts (here on one device;
    across GPUs the gather is one ncclAllGather, optuna_b200/dist.py) each evaluate g(x) over their slice of the above
    kernels, the per-candidate (max, sum) partials are gathered and every context finishes with the suggestion the
    single-context call computes: same candidate, acquisition value within 1e-12.  Uneven slices, slices without
    kernels (fewer tiles than ranks) and the full config-2 size."""
    import torch
    from optuna_b200 import TPEEngine
    from optuna_b200.dist import device_view
    from optuna_b200.engine import ParamSpec
    rs = np.random.RandomState(world * 7 + P)
    specs = [ParamSpec(kind=0, low=0.0, high=1.0) for _ in range(P)]
    X = rs.uniform(0, 1, (n, P))
    key = np.stack([((X - 0.5) ** 2).sum(1), np.zeros(n)], 1)
    cols = list(range(P))
    cfg = dict(n_below=25, n_candidates=C, multivariate=True)
    u = rs.random_sample(C * (1 + P))
    eng.set_space(specs)
    eng.set_history(X, np.zeros(n, np.int8), key)
    want_x, want_acq, want_best = eng.suggest(cols, u, 1, **cfg)
    _, ll, lg = eng.get_candidates()
    ranks = [TPEEngine(0) for _ in range(world)]
    try:
        parts = []
        for r, e in enumerate(ranks):
            e.set_space(specs)
            e.set_history(X, np.zeros(n, np.int8), key)
            e.set_kernel_shard(r, world)
            e.prepare(cols, **cfg)
            e.build()
            ptr, stride = e.sample_and_partial(u, 1)
            parts.append(device_view(ptr, (stride, 2), "<f8", torch.device("cuda", 0)))
        gathered = torch.stack(parts).contiguous()
        torch.cuda.synchronize()
        for e in ranks:
            x, acq, best = e.finish_from_partials(gathered.data_ptr(), world)
            assert np.array_equal(x, want_x) and best[0] == want_best[0]
            close(acq, want_acq, 0, 1e-12)
            _, ll_r, lg_r = e.get_candidates()
            close(lg_r, lg, 0, 1e-12)
            assert np.array_equal(ll_r, ll)
        with pytest.raises(RuntimeError, match="tpe_sample_and_partial must precede"):
            ranks[0].finish_from_partials(gathered.data_ptr(), world)
        ranks[0].set_kernel_shard(0, 1)                      # off again: the ordinary call
        x, acq, best = ranks[0].suggest(cols, u, 1, **cfg)
        assert np.array_equal(x, want_x) and acq[0] == want_acq[0]
    finally:
        for e in ranks:
            e.close()


def test_config2_full_size_against_the_precomputed_oracle_fixture(eng):
    """BASELINE config 2 at full size: log l(x) and log g(x) of 256 points -- the first 256 candidates the oracle draws
    -- against tests/golden/c2_logpdf.npz (oracle/gen_c2_fixture.py: the chunked oracle, ~6 min of CPU, so it is
    committed rather than recomputed).  tpe_logpdf evaluates the SAME points (bit for bit) with the same grid
    kernels a suggestion uses; 1e-12 absolute."""
    from optuna_b200.engine import ParamSpec
    g = load("c2_logpdf.npz")
    N, P = 100_000, 32
    rs = np.random.RandomState(0)
    X = rs.uniform(0, 1, (N, P))
    key = np.stack([((X - 0.5) ** 2).sum(1), np.zeros(N)], 1)
    eng.set_space([ParamSpec(kind=0, low=0.0, high=1.0) for _ in range(P)])
    eng.set_history(X, np.zeros(N, np.int8), key)
    eng.prepare(list(range(P)), n_below=orc.default_gamma(N), n_candidates=4096, multivariate=True)
    below, _ = eng.get_split()
    assert np.array_equal(below, g["below"])
    eng.build()
    lg = eng.logpdf(1, g["x"])
    assert eng.last_logpdf_kernel().startswith(("k_tcs", "k_logpdf_mma"))
    ll = eng.logpdf(0, g["x"])
    assert g["x"].shape == (256, P)
    close(lg, g["logg"], 0, 1e-12)
    close(ll, g["logl"], 0, 1e-12)
    # ... and inside a suggestion (same seed => the candidates are these points to 1e-12): the 256 values again
    u = draw_uniforms(np.random.RandomState(1), 4096, 0, P)
    eng.suggest(list(range(P)), u, 1, n_below=orc.default_gamma(N), n_candidates=4096, multivariate=True)
    smp, ll2, lg2 = eng.get_candidates()
    close(smp[:256], g["x"], 1e-12, 1e-12)
    close(lg2[:256], g["logg"], 0, 2e-12)   # the points differ by <= 1e-12 relative: |d log g / dx| <~ 1e2 each
    close(ll2[:256], g["logl"], 0, 2e-12)


def test_far_tier_worst_case_every_term_just_below_the_exact_window(eng):
    """The fp32 ("far") tier of the tensor-core kernel is bounded for the case that ALL K terms sit just below the
    exact window (ln K + 17.5 under the max): build exactly that -- one kernel at the query point, every other kernel
    on a sphere around it whose radius puts its term 0.01 nats outside the window -- plus the mirror case 0.01 nats
    inside (every term exact), and a sphere just inside / outside the drop line (ln K + 30).  1e-12 absolute."""
    from optuna_b200.engine import ParamSpec
    P, n = 8, 20_000
    rs = np.random.RandomState(5)
    specs = [ParamSpec(kind=0, low=0.0, high=1.0) for _ in range(P)]
    params = [orc.Param("float", 0.0, 1.0) for _ in range(P)]
    cfg = orc.Config(multivariate=True)
    x0 = np.full(P, 0.5)
    sigma = max(0.2 * n ** (-1.0 / (P + 4)), 1.0 / min(100.0, 1.0 + n + 1))     # parzen_estimator.py:187-228, range 1
    K = n + 1
    for gap in (np.log(K) + 17.5 + 0.01, np.log(K) + 17.5 - 0.01, np.log(K) + 30.0 - 0.01, np.log(K) + 30.0 + 0.01):
        r = sigma * np.sqrt(2.0 * gap)                                           # |x0 - mu|^2 / (2 sigma^2) = gap
        d = rs.choice([-1.0, 1.0], size=(n - 1, P)) * (r / np.sqrt(P))        # on the sphere, every coordinate inside (0, 1)
        X = np.concatenate([x0[None, :], x0[None, :] + d])
        assert X.min() > 0 and X.max() < 1
        # all in the above set: 25 dummy best trials far away take the below slots
        Xall = np.concatenate([rs.uniform(0.0, 0.02, (25, P)), X])
        keyv = np.concatenate([np.full(25, -1.0), np.ones(n)])
        key = np.stack([keyv, np.zeros(n + 25)], 1)
        eng.set_space(specs)
        eng.set_history(Xall, np.zeros(n + 25, np.int8), key)
        eng.prepare(list(range(P)), n_below=25, n_candidates=64, multivariate=True)
        eng.build(None, np.ones(n))                                              # equal weights: equal terms
        pts = np.stack([x0, x0 + 1e-3, x0 - 2e-3 * np.arange(P) / P])
        got = eng.logpdf(1, pts)
        assert eng.last_logpdf_kernel().startswith(("k_tcs", "k_logpdf_mma"))
        cfg_w = orc.Config(multivariate=True, weights=lambda k: np.ones(k))
        ma = orc.build_mixture(X, params, cfg_w)
        close(got, orc.mixture_log_pdf(ma, pts), 0, 1e-12)


def test_nan_cells_and_nan_wins_argmax(eng):
    """_truncnorm.logpdf returns nan where a == b (a parameter whose low == high, _truncnorm.py:286-297); the
    acquisition is then nan for every candidate and np.argmax returns the first index (sampler.py:603-618: NaN
    wins).  Same on the device: k_select."""
    from optuna_b200.engine import ParamSpec
    rs = np.random.RandomState(9)
    n, C = 300, 32
    X = np.stack([rs.uniform(-1, 1, n), np.full(n, 2.5), rs.uniform(0, 4, n)], 1)
    key = np.stack([(X[:, 0] ** 2 + X[:, 2]), np.zeros(n)], 1)
    specs = [ParamSpec(kind=0, low=-1.0, high=1.0), ParamSpec(kind=0, low=2.5, high=2.5), ParamSpec(kind=0, low=0.0, high=4.0)]
    params = [orc.Param("float", -1.0, 1.0), orc.Param("float", 2.5, 2.5), orc.Param("float", 0.0, 4.0)]
    eng.set_space(specs)
    eng.set_history(X, np.zeros(n, np.int8), key)
    for mv, cols in ((True, [0, 1, 2]), (False, [1]), (True, [0, 2])):
        u = draw_uniforms(np.random.RandomState(4), C, 0, len(cols))
        with np.errstate(all="ignore"):
            s = orc.suggest(X, np.zeros(n, np.int8), key, params, cols, orc.Config(multivariate=mv), 25, C,
                            np.random.RandomState(4))
        x, acq, best = eng.suggest(cols, u, 1, n_below=25, n_candidates=C, multivariate=mv)
        smp, ll, lg = eng.get_candidates()
        assert np.array_equal(np.isnan(ll - lg), np.isnan(s.acq)), (mv, cols)
        assert int(best[0]) == s.best, (mv, cols, best, s.best)
        if 1 in cols:
            assert np.isnan(s.acq).all() and s.best == 0 and np.isnan(acq[0])
