"""Pin the oracle (oracle/*.py) against golden vectors produced by the live reference.

The fixtures under tests/golden/ were written by oracle/gen_golden.py importing
/root/reference (optuna @ 4df4b72); see that script for the exact reference calls.
"""
import numpy as np
import pytest

from oracle import tpe_math as tm
from oracle import tpe_oracle as orc
from tests._util import decode_space, load, mixture_from_gold


def same(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    assert a.shape == b.shape
    assert np.array_equal(a, b, equal_nan=True), f"max abs diff {np.nanmax(np.abs(a - b))}"


def test_math_bit_exact():
    g = load("math.npz")
    same(tm.erf(g["erf_x"]), g["erf_big"])
    same(tm.erf(g["erf_x"][:500]), g["erf_small"])
    same(tm.log_ndtr(g["lndtr_t"]), g["lndtr"])
    same(tm.ndtr(np.linspace(-9, 9, 2401)), g["ndtr_big"])
    same(tm.log_gauss_mass(g["lgm_a"], g["lgm_b"]), g["lgm"])
    same(tm.log_gauss_mass(g["lgm_a"][:700], g["lgm_b"][:700]), g["lgm_small"])
    same(tm.ppf(g["ppf_q"], g["ppf_a"], g["ppf_b"]), g["ppf"])
    same(tm.ndtri_exp(g["ndtri_y"].copy()), g["ndtri"])
    lo, hi, loc, sc = g["lpdf_lo"], g["lpdf_hi"], g["lpdf_loc"], g["lpdf_scale"]
    same(tm.logpdf(g["lpdf_x"], (lo - loc) / sc, (hi - loc) / sc, loc, sc), g["lpdf"])


def test_parzen_build_sample_logpdf_bit_exact():
    g = load("parzen.npz")
    for i in range(int(g["n_cases"])):
        t = f"pz{i}/"
        params = decode_space(g[t + "space"])
        mv, clip, endp, pw, C, seed = g[t + "flags"]
        cfg = orc.Config(prior_weight=float(pw), magic_clip=bool(clip), endpoints=bool(endp),
                         multivariate=bool(mv))
        mix = orc.build_mixture(g[t + "obs"], params, cfg)
        same(mix.weights, g[t + "w"])
        for j, p in enumerate(params):
            if p.is_cat:
                same(mix.cat_w[j], g[f"{t}cat{j}"])
            else:
                same(mix.mu[j], g[f"{t}mu{j}"])
                same(mix.sigma[j], g[f"{t}sigma{j}"])
        smp = orc.mixture_sample(mix, np.random.RandomState(int(seed) + 1000), int(C))
        same(smp, g[t + "samples"])
        same(orc.mixture_log_pdf(mix, smp), g[t + "logpdf"])
        same(orc.mixture_log_pdf(mix, g[t + "samples2"]), g[t + "logpdf2"])


def test_suggest_bit_exact():
    g = load("suggest.npz")
    for ci in range(int(g["n_cases"])):
        t = f"sg{ci}/"
        params = decode_space(g[t + "space"])
        mv, C, seed, n_below = g[t + "cfg"]
        X, cat, key = g[t + "X"], g[t + "category"], g[t + "key"]
        below, above = orc.split_trials(cat, key, int(n_below))
        assert np.array_equal(below, g[t + "below"])
        assert np.array_equal(above, g[t + "above"])
        cfg = orc.Config(multivariate=bool(mv))
        rng = np.random.RandomState(int(seed))
        if mv:
            s = orc.suggest(X, cat, key, params, list(range(len(params))), cfg, int(n_below), int(C), rng)
            same(s.samples, g[t + "samples"])
            same(s.logl, g[t + "ll"])
            same(s.logg, g[t + "lg"])
            same(s.mix_above.weights, g[t + "a_w"])
            same(s.x, g[t + "ret_internal"])
        else:
            for j in range(len(params)):
                s = orc.suggest(X, cat, key, params, [j], cfg, int(n_below), int(C), rng)
                same(s.samples, g[f"{t}u{j}/samples"])
                same(s.logl, g[f"{t}u{j}/ll"])
                same(s.logg, g[f"{t}u{j}/lg"])
                same(s.x[0], g[t + "ret_internal"][j])


def test_chunked_logpdf_matches_unchunked_for_continuous():
    g = load("suggest.npz")
    t = "sg5/"
    params = decode_space(g[t + "space"])
    mix = mixture_from_gold(g, t + "a_", params)
    full = orc.mixture_log_pdf(mix, g[t + "samples"][:32])
    same(orc.mixture_log_pdf_chunked(mix, g[t + "samples"][:32], rows=5), full)
    same(full, g[t + "lg"][:32])


def test_motpe_primitives_bit_exact():
    from oracle import motpe as mo
    g = load("motpe.npz")
    for ci in range(int(g["hv_n"])):
        t = f"hv{ci}/"
        v, ref = g[t + "v"], g[t + "ref"]
        same(mo.reference_point(v), ref)
        same(mo.hypervolume(v, ref), g[t + "hv"])
        assert np.array_equal(mo.nondomination_rank(v), g[t + "rank"])
        n = v.shape[0]
        assert np.array_equal(mo.nondomination_rank(v, n_below=max(1, n // 3)), g[t + "rank_nb"])
        assert np.array_equal(mo.solve_hssp(v, np.arange(n) * 3, max(1, n // 2), ref), g[t + "hssp"])


def test_motpe_split_and_weights_bit_exact():
    from oracle import motpe as mo
    g = load("motpe.npz")
    for ci in range(int(g["mo_n"])):
        t = f"mo{ci}/"
        v, nb = g[t + "v"], int(g[t + "nb"])
        below = mo.split_complete_mo(v, nb)
        assert np.array_equal(below, g[t + "below"]), ci
        same(mo.weights_below_mo(v[below]), g[t + "w"])


def test_motpe_suggest_bit_exact():
    from oracle import motpe as mo
    g = load("motpe.npz")
    for ci in range(int(g["mosg_n"])):
        t = f"mosg{ci}/"
        mv, C, seed, n_below, m = g[t + "cfg"]
        X, vals = g[t + "X"], g[t + "values"]
        n, P = X.shape
        params = [orc.Param("float", 0.0, 1.0) for _ in range(P)]
        cat, key = np.zeros(n, np.int8), np.zeros((n, 2))
        cfg = orc.Config(multivariate=bool(mv))
        rng = np.random.RandomState(int(seed))
        sel = lambda comp, k: comp[mo.split_complete_mo(vals[comp], k)]  # noqa: E731
        below, _ = orc.split_trials(cat, key, int(n_below), sel)
        wb = mo.weights_below_mo(vals[below])
        calls = [list(range(P))] if mv else [[j] for j in range(P)]
        ret = []
        for q, cols in enumerate(calls):
            s = orc.suggest(X, cat, key, params, cols, cfg, int(n_below), int(C), rng, weights_below=wb,
                            complete_selector=sel)
            same(s.mix_below.weights, g[f"{t}c{q}/wb"])
            same(s.samples, g[f"{t}c{q}/samples"])
            same(s.logl, g[f"{t}c{q}/ll"])
            same(s.logg, g[f"{t}c{q}/lg"])
            ret.extend(s.x.tolist())
        same(np.asarray(ret), g[t + "ret"])
