"""Oracle for the TPE suggestion path (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

Array-in / array-out NumPy restatement of what the reference does per suggestion:

* gamma / weights          <- optuna/samplers/_tpe/sampler.py:53-69
* split_trials             <- sampler.py:686-722, :735-742, :782-821   (below / above sets)
* build_mixture            <- optuna/samplers/_tpe/parzen_estimator.py:39-78, :132-251
* mixture_sample           <- optuna/samplers/_tpe/probability_distributions.py:86-152
* mixture_log_pdf          <- probability_distributions.py:154-223
* suggest                  <- sampler.py:523-560, :591-618

The history is handed in as arrays (the same arrays the C-ABI takes), not FrozenTrial lists:
``X[N, P]`` internal representation with NaN = parameter absent, ``category[N]`` (0 COMPLETE,
1 PRUNED, 2 infeasible, 3 RUNNING) and ``key[N, 2]`` = the sort key the reference uses inside
each category (COMPLETE: (signed value, 0); PRUNED: (-last_step, signed value), sampler.py:782-792;
infeasible: (violation, 0), sampler.py:803-813).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable

import numpy as np

from . import tpe_math as tm

EPS = 1e-12  # parzen_estimator.py:24, sampler.py:45

CAT_COMPLETE, CAT_PRUNED, CAT_INFEASIBLE, CAT_RUNNING = 0, 1, 2, 3


# ----------------------------------------------------------------------------------------------
# search space description
# ----------------------------------------------------------------------------------------------
@dataclass
class Param:
    """One search-space dimension.  kind: "float" | "int" | "cat"."""

    kind: str
    low: float = 0.0
    high: float = 0.0
    step: float | None = None  # None => continuous; ints always carry a step (distributions.py:340-360)
    log: bool = False
    n_choices: int = 0
    # optional [n_choices, n_choices] table of categorical_distance_func values
    dist_table: np.ndarray | None = None

    @property
    def is_cat(self) -> bool:
        return self.kind == "cat"


@dataclass
class Config:
    prior_weight: float = 1.0
    magic_clip: bool = True
    endpoints: bool = False
    multivariate: bool = False
    weights: Callable[[int], np.ndarray] | None = None  # default_weights if None
    stable_sort: bool = False  # reference uses np.argsort default (unstable) in compute_sigmas


@dataclass
class Mixture:
    weights: np.ndarray  # [K]
    params: list[Param]
    # per param: categorical -> W[K, n_choices]; numeric -> (mu[K], sigma[K])
    cat_w: dict[int, np.ndarray] = field(default_factory=dict)
    mu: dict[int, np.ndarray] = field(default_factory=dict)
    sigma: dict[int, np.ndarray] = field(default_factory=dict)


def default_gamma(n: int) -> int:
    return min(math.ceil(0.1 * n), 25)


def hyperopt_default_gamma(n: int) -> int:
    return min(math.ceil(0.25 * n**0.5), 25)


def default_weights(n: int) -> np.ndarray:
    if n == 0:
        return np.asarray([])
    if n < 25:
        return np.ones(n)
    return np.concatenate([np.linspace(1.0 / n, 1.0, num=n - 25), np.ones(25)], axis=0)


# ----------------------------------------------------------------------------------------------
# split
# ----------------------------------------------------------------------------------------------
def _take_best(idx: np.ndarray, key: np.ndarray, m: int) -> tuple[np.ndarray, np.ndarray]:
    """Stable selection of the m lexicographically smallest (key0, key1) rows among idx."""
    m = min(m, idx.size)
    if idx.size == 0:
        return idx, idx
    # np.lexsort is stable; last key is the primary one.
    order = np.lexsort((key[idx, 1], key[idx, 0]))
    return idx[order[:m]], idx[order[m:]]


def split_trials(category: np.ndarray, key: np.ndarray, n_below: int,
                 complete_selector: Callable[[np.ndarray, int], np.ndarray] | None = None
                 ) -> tuple[np.ndarray, np.ndarray]:
    """Index sets (ascending trial order) of the below / above groups.

    ``complete_selector(idx_complete, m)`` overrides the single-objective value sort for the
    COMPLETE group (used by the multi-objective split, oracle/motpe.py).
    """
    category = np.asarray(category)
    all_idx = np.arange(category.size)
    comp = all_idx[category == CAT_COMPLETE]
    prun = all_idx[category == CAT_PRUNED]
    infe = all_idx[category == CAT_INFEASIBLE]
    runn = all_idx[category == CAT_RUNNING]

    if complete_selector is None:
        b0, a0 = _take_best(comp, key, n_below)
    else:
        m = min(n_below, comp.size)
        b0 = np.asarray(complete_selector(comp, m), dtype=np.int64)
        a0 = np.setdiff1d(comp, b0)
    n_below = max(0, n_below - b0.size)
    b1, a1 = _take_best(prun, key, n_below)
    n_below = max(0, n_below - b1.size)
    b2, a2 = _take_best(infe, key, n_below)
    below = np.sort(np.concatenate([b0, b1, b2]).astype(np.int64))
    above = np.sort(np.concatenate([a0, a1, a2, runn]).astype(np.int64))
    return below, above


# ----------------------------------------------------------------------------------------------
# Parzen estimator build
# ----------------------------------------------------------------------------------------------
def _call_weights(cfg: Config, n: int) -> np.ndarray:
    func = cfg.weights or default_weights
    w = np.array(func(n))[:n]
    if np.any(w < 0):
        raise ValueError("The `weights` function is not allowed to return negative values.")
    if len(w) > 0 and np.sum(w) <= 0:
        raise ValueError("The `weight` function is not allowed to return all-zero values.")
    if not np.all(np.isfinite(w)):
        raise ValueError("The `weights`function is not allowed to return infinite or NaN values.")
    return w


def numeric_domain(p: Param) -> tuple[float, float]:
    """Kernel support in the (possibly log) kernel space (parzen_estimator.py:174-182)."""
    lo, hi = p.low, p.high
    if p.step is not None:
        lo = lo - p.step / 2
        hi = hi + p.step / 2
    if p.log:
        lo = np.log(lo)
        hi = np.log(hi)
    return lo, hi


def _bandwidths(mus: np.ndarray, lo: float, hi: float, d: int, cfg: Config) -> np.ndarray:
    n = mus.size
    if cfg.multivariate:
        s = 0.2 * max(n, 1) ** (-1.0 / (d + 4)) * (hi - lo)
        sig = np.full((n,), s)
    else:
        centre = 0.5 * (lo + hi)
        ext = np.append(mus, centre)
        order = np.argsort(ext, kind="stable") if cfg.stable_sort else np.argsort(ext)
        padded = np.empty(ext.size + 2, dtype=float)
        padded[0] = lo
        padded[1:-1] = ext[order]
        padded[-1] = hi
        gaps = np.maximum(padded[1:-1] - padded[:-2], padded[2:] - padded[1:-1])
        if not cfg.endpoints and padded.size >= 4:
            gaps[0] = padded[2] - padded[1]
            gaps[-1] = padded[-2] - padded[-3]
        sig = gaps[np.argsort(order)][:n]
    top = hi - lo
    if cfg.magic_clip:
        bottom = (hi - lo) / min(100.0, 1.0 + (n + 1))
    else:
        bottom = EPS
    return np.asarray(np.clip(sig, bottom, top))


def _categorical_rows(obs: np.ndarray, p: Param, cfg: Config) -> np.ndarray:
    c = p.n_choices
    if obs.size == 0:
        return np.full((1, c), 1.0 / c)
    k = obs.size + 1
    rows = np.full((k, c), cfg.prior_weight / k)
    seen = obs.astype(int)
    if p.dist_table is not None:
        uniq, back = np.unique(seen, return_inverse=True)
        d = np.asarray(p.dist_table, dtype=float)[uniq]
        coef = np.log(k / cfg.prior_weight) * np.log(c) / np.log(6)
        rows[: seen.size] = np.exp(-((d / np.max(d, axis=1)[:, np.newaxis]) ** 2) * coef)[back]
    else:
        rows[np.arange(seen.size), seen] += 1
    tot = rows.sum(axis=1, keepdims=True)
    rows /= np.where(tot == 0, 1, tot)
    return rows


def build_mixture(obs: np.ndarray, params: list[Param], cfg: Config,
                  predetermined_weights: np.ndarray | None = None) -> Mixture:
    """obs: [n, P] observations (rows already restricted to trials holding every param)."""
    if cfg.prior_weight < 0:
        raise ValueError("A non-negative value must be specified for prior_weight.")
    obs = np.asarray(obs, dtype=float).reshape(-1, len(params))
    n = obs.shape[0]
    if predetermined_weights is not None:
        assert len(predetermined_weights) == n
        w = np.asarray(predetermined_weights, dtype=float)
    else:
        w = _call_weights(cfg, n)
    if n == 0:
        w = np.array([1.0])
    else:
        w = np.append(w, [cfg.prior_weight])
    w = w / w.sum()
    mix = Mixture(weights=w, params=list(params))
    d = len(params)
    for j, p in enumerate(params):
        col = obs[:, j]
        if p.is_cat:
            mix.cat_w[j] = _categorical_rows(col, p, cfg)
            continue
        lo, hi = numeric_domain(p)
        mus = np.log(col) if p.log else col
        sig = _bandwidths(mus, lo, hi, d, cfg)
        mix.mu[j] = np.append(mus, [0.5 * (lo + hi)])
        mix.sigma[j] = np.append(sig, [hi - lo])
    return mix


# ----------------------------------------------------------------------------------------------
# sampling from l(x)
# ----------------------------------------------------------------------------------------------
def mixture_sample(mix: Mixture, rng: np.random.RandomState, size: int) -> np.ndarray:
    """[size, P] candidates in the *external-numeric* space (log params exponentiated)."""
    active = rng.choice(len(mix.weights), p=mix.weights, size=size)
    out = np.empty((size, len(mix.params)), dtype=float)
    num_idx = [j for j, p in enumerate(mix.params) if not p.is_cat]
    for j, p in enumerate(mix.params):
        if not p.is_cat:
            continue
        rows = mix.cat_w[j][active, :]
        u = rng.rand(size)
        cdf = np.cumsum(rows, axis=-1)
        assert np.isclose(cdf[:, -1], 1).all()
        cdf[:, -1] = 1
        out[:, j] = np.sum(cdf < u[:, np.newaxis], axis=-1)
    if num_idx:
        mu_a = np.asarray([mix.mu[j][active] for j in num_idx])
        sg_a = np.asarray([mix.sigma[j][active] for j in num_idx])
        doms = [numeric_domain(mix.params[j]) for j in num_idx]
        lo_k = np.asarray([d[0] for d in doms])
        hi_k = np.asarray([d[1] for d in doms])
        a = (lo_k[:, np.newaxis] - mu_a) / sg_a
        b = (hi_k[:, np.newaxis] - mu_a) / sg_a
        shape = np.broadcast(a, b, mu_a, sg_a).shape
        u = rng.uniform(low=0, high=1, size=shape)
        draws = (tm.ppf(u, a, b) * sg_a + mu_a).T
        out[:, num_idx] = draws
        log_idx = [j for j in num_idx if mix.params[j].log]
        out[:, log_idx] = np.exp(out[:, log_idx])
        disc = [j for j in num_idx if mix.params[j].step is not None]
        if disc:
            lo_d = np.asarray([mix.params[j].low for j in disc], dtype=float)
            hi_d = np.asarray([mix.params[j].high for j in disc], dtype=float)
            st_d = np.asarray([mix.params[j].step for j in disc], dtype=float)
            out[:, disc] = np.clip(lo_d + np.round((out[:, disc] - lo_d) / st_d) * st_d, lo_d, hi_d)
    return out


# ----------------------------------------------------------------------------------------------
# log density of the mixture
# ----------------------------------------------------------------------------------------------
def _pair_unique(u: np.ndarray, v: np.ndarray) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    order = np.argsort(v)
    order = order[np.argsort(u[order], kind="stable")]
    us, vs = u[order], v[order]
    first = np.empty(u.shape, dtype=bool)
    first[0] = True
    first[1:] = (us[1:] != us[:-1]) | (vs[1:] != vs[:-1])
    inv = np.empty(u.size, dtype=int)
    inv[order] = np.cumsum(first) - 1
    return us[first], vs[first], inv


def _mass_dedup(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    ua, ub, inv = _pair_unique(a.ravel(), b.ravel())
    return tm.log_gauss_mass(ua, ub)[inv].reshape(a.shape)


def mixture_log_pdf(mix: Mixture, x: np.ndarray) -> np.ndarray:
    """x: [C, P] in the external-numeric space.  Returns [C]."""
    x = np.asarray(x, dtype=float).reshape(-1, len(mix.params))
    acc = np.zeros((x.shape[0], mix.weights.size), dtype=np.float64)
    cont: list[int] = []
    for j, p in enumerate(mix.params):
        if p.is_cat:
            pick = x[:, j, np.newaxis, np.newaxis].astype(np.int64)
            with np.errstate(divide="ignore"):
                acc += np.log(np.take_along_axis(mix.cat_w[j][np.newaxis], pick, axis=-1))[..., 0]
        elif p.step is None:
            cont.append(j)
        else:
            half = p.step / 2
            xu, xinv = np.unique(x[:, j], return_inverse=True)
            mu_u, sg_u, kinv = _pair_unique(mix.mu[j], mix.sigma[j])
            if p.log:
                lo_edge, hi_edge = np.log(xu - half), np.log(xu + half)
                dom_lo, dom_hi = np.log(p.low - half), np.log(p.high + half)
            else:
                lo_edge, hi_edge = xu - half, xu + half
                dom_lo, dom_hi = p.low - half, p.high + half
            acc += _mass_dedup((lo_edge[:, np.newaxis] - mu_u) / sg_u,
                               (hi_edge[:, np.newaxis] - mu_u) / sg_u)[np.ix_(xinv, kinv)]
            acc -= tm.log_gauss_mass((dom_lo - mu_u) / sg_u, (dom_hi - mu_u) / sg_u)[kinv]
    if cont:
        xs = np.asarray([np.log(x[:, j]) if mix.params[j].log else x[:, j] for j in cont]).T
        lo = np.asarray([np.log(mix.params[j].low) if mix.params[j].log else mix.params[j].low
                         for j in cont])
        hi = np.asarray([np.log(mix.params[j].high) if mix.params[j].log else mix.params[j].high
                         for j in cont])
        mu = np.asarray([mix.mu[j] for j in cont]).T
        sg = np.asarray([mix.sigma[j] for j in cont]).T
        acc += tm.logpdf(xs[:, np.newaxis, :], (lo - mu) / sg, (hi - mu) / sg, mu, sg).sum(axis=-1)
    with np.errstate(divide="ignore"):
        acc += np.log(mix.weights[np.newaxis])
    top = acc.max(axis=1)
    top[np.isneginf(top)] = 0
    with np.errstate(divide="ignore"):
        return np.log(np.exp(acc - top[:, None]).sum(axis=1)) + top


def mixture_log_pdf_chunked(mix: Mixture, x: np.ndarray, rows: int = 16) -> np.ndarray:
    """Same values as mixture_log_pdf, evaluated on slices of the candidate axis.

    Rows of the log-density are independent (verified bit-identical on the live
    reference for continuous params); needed because one (C, K, P) temporary at config 2 is 105 GB.
    NOTE: for *discrete* params np.unique over a slice changes which erf code path runs
    (_erf.py:134 size switch), so slices agree only to ~1e-16 there.
    """
    x = np.asarray(x, dtype=float).reshape(-1, len(mix.params))
    parts = [mixture_log_pdf(mix, x[s: s + rows]) for s in range(0, x.shape[0], rows)]
    return np.concatenate(parts) if parts else np.empty((0,))


# ----------------------------------------------------------------------------------------------
# one suggestion
# ----------------------------------------------------------------------------------------------
def observations(X: np.ndarray, rows: np.ndarray, cols: list[int]) -> tuple[np.ndarray, np.ndarray]:
    """Rows of the chosen trials that hold every selected param (sampler.py:511-521).

    Returns (obs[n, len(cols)], mask over ``rows``)."""
    sub = X[np.asarray(rows, dtype=np.int64)][:, cols]
    keep = ~np.isnan(sub).any(axis=1)
    return sub[keep], keep


@dataclass
class Suggestion:
    x: np.ndarray  # [P] chosen candidate, external-numeric (internal repr of the distribution)
    best: int
    acq: np.ndarray
    samples: np.ndarray
    logl: np.ndarray
    logg: np.ndarray
    below: np.ndarray
    above: np.ndarray
    mix_below: Mixture
    mix_above: Mixture


def suggest(X: np.ndarray, category: np.ndarray, key: np.ndarray, params: list[Param],
            cols: list[int], cfg: Config, n_below: int, n_candidates: int,
            rng: np.random.RandomState, weights_below: np.ndarray | None = None,
            complete_selector=None, chunk_rows: int | None = None) -> Suggestion:
    below, above = split_trials(category, key, n_below, complete_selector)
    sub = [params[c] for c in cols]
    obs_b, keep_b = observations(X, below, cols)
    obs_a, _ = observations(X, above, cols)
    wb = None if weights_below is None else np.asarray(weights_below)[keep_b]
    mix_b = build_mixture(obs_b, sub, cfg, wb)
    mix_a = build_mixture(obs_a, sub, cfg)
    cand = mixture_sample(mix_b, rng, n_candidates)
    if chunk_rows:
        ll = mixture_log_pdf_chunked(mix_b, cand, chunk_rows)
        lg = mixture_log_pdf_chunked(mix_a, cand, chunk_rows)
    else:
        ll = mixture_log_pdf(mix_b, cand)
        lg = mixture_log_pdf(mix_a, cand)
    acq = ll - lg
    if cand.shape[0] == 0:
        raise ValueError("The size of `samples` must be positive, but got 0.")
    best = int(np.argmax(acq))
    return Suggestion(cand[best].copy(), best, acq, cand, ll, lg, below, above, mix_b, mix_a)
