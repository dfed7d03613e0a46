"""Oracle for the multi-objective part of TPE (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

NumPy restatement of

* nondomination_rank   <- optuna/study/_multi_objective.py:127-219  (Pareto peeling on the unique,
                          lexsorted loss vectors; duplicates share a rank)
* hypervolume          <- optuna/_hypervolume/wfg.py:8-181          (2-D sweep, 3-D cummax table,
                          N-D WFG recursion)
* solve_hssp           <- optuna/_hypervolume/hssp.py:10-176        (greedy subset selection with
                          lazily updated submodular upper bounds)
* reference_point      <- optuna/samplers/_tpe/sampler.py:679-683
* split_complete_mo    <- sampler.py:745-779
* weights_below_mo     <- sampler.py:824-863

Loss vectors are sign-normalised (every objective minimised).  Arithmetic order follows the
reference so that results agree bit-for-bit (pinned by tests/test_oracle_golden.py).
"""
from __future__ import annotations

import math

import numpy as np

EPS = 1e-12


# ------------------------------------------------------------------------------------------------
# Pareto fronts / ranks
# ------------------------------------------------------------------------------------------------
def _front_unique_sorted(v: np.ndarray) -> np.ndarray:
    """Mask of the non-dominated rows of a unique, lexsorted [n, m] array."""
    n, m = v.shape
    if m == 1:
        out = np.zeros(n, dtype=bool)
        out[0] = True
        return out
    if m == 2:
        run = np.minimum.accumulate(v[:, 1])
        out = np.ones(n, dtype=bool)
        out[1:] = run[1:] < run[:-1]
        return out
    rest = v[:, 1:]
    out = np.zeros(n, dtype=bool)
    alive = np.arange(n)
    while len(alive):
        head = alive[0]
        out[head] = True
        keep = np.any(rest[alive] < rest[head], axis=1)
        alive = alive[keep]
    return out


def is_pareto_front(v: np.ndarray, assume_unique_lexsorted: bool) -> np.ndarray:
    if assume_unique_lexsorted:
        return _front_unique_sorted(v)
    u, inv = np.unique(v, axis=0, return_inverse=True)
    return _front_unique_sorted(u)[inv.reshape(-1)]


def nondomination_rank(v: np.ndarray, n_below: int | None = None) -> np.ndarray:
    if len(v) == 0 or (n_below is not None and n_below <= 0):
        return np.zeros(len(v), dtype=int)
    n, m = v.shape
    if m == 1:
        return np.unique(v[:, 0], return_inverse=True)[1]
    u, inv = np.unique(v, return_inverse=True, axis=0)
    nu = u.shape[0]
    n_below = min(n_below or nu, nu)
    ranks = np.zeros(nu, dtype=int)
    idx = np.arange(nu)
    r = 0
    while nu - idx.size < n_below:
        front = _front_unique_sorted(u)
        ranks[idx[front]] = r
        idx = idx[~front]
        u = u[~front]
        r += 1
    ranks[idx] = r
    return ranks[inv.reshape(-1)]


# ------------------------------------------------------------------------------------------------
# hypervolume
# ------------------------------------------------------------------------------------------------
def _hv_2d(s: np.ndarray, ref: np.ndarray) -> float:
    ys = np.concatenate([ref[1:], s[:-1, 1]])
    return (ref[0] - s[:, 0]) @ (ys - s[:, 1])


def _hv_3d(s: np.ndarray, ref: np.ndarray) -> float:
    n = s.shape[0]
    yo = np.argsort(s[:, 1])
    z = np.zeros((n, n), dtype=float)
    z[yo, np.arange(n)] = ref[2] - s[yo, 2]
    z = np.maximum.accumulate(np.maximum.accumulate(z, axis=0), axis=1)
    xv = s[:, 0]
    yv = s[yo, 1]
    dx = np.concatenate([xv[1:], ref[:1]]) - xv
    dy = np.concatenate([yv[1:], ref[1:2]]) - yv
    return np.dot(np.dot(z, dy), dx)


def _hv_nd(s: np.ndarray, ref: np.ndarray) -> float:
    if s.shape[0] == 1:
        out = 1.0
        for r, v in zip(ref, s[0]):
            out *= r - v
        return float(out)
    if s.shape[0] == 2:
        h1 = h2 = cap = 1.0
        for r, a, b in zip(ref, s[0], s[1]):
            h1 *= r - a
            h2 *= r - b
            cap *= r - max(a, b)
        return h1 + h2 - cap
    incl = (ref - s).prod(axis=-1)
    lim = np.maximum(s[:, np.newaxis], s)
    return incl[-1] + sum(_excl(lim[i, i + 1:], incl[i], ref) for i in range(incl.size - 1))


def _excl(lim: np.ndarray, incl: float, ref: np.ndarray) -> float:
    if lim.shape[0] <= 3:
        return incl - _hv_nd(lim, ref)
    front = _front_unique_sorted(lim)
    return incl - _hv_nd(lim[front], ref)


def hypervolume(v: np.ndarray, ref: np.ndarray, assume_pareto: bool = False) -> float:
    if not np.all(v <= ref):
        raise ValueError("All points must dominate or equal the reference point.")
    if not np.all(np.isfinite(ref)):
        return float("inf")
    if v.size == 0:
        return 0.0
    if not assume_pareto:
        u = np.unique(v, axis=0)
        s = u[_front_unique_sorted(u)]
    else:
        s = v[v[:, 0].argsort()]
    if ref.shape[0] == 2:
        hv = _hv_2d(s, ref)
    elif ref.shape[0] == 3:
        hv = _hv_3d(s, ref)
    else:
        hv = _hv_nd(s, ref)
    return hv if np.isfinite(hv) else float("inf")


# ------------------------------------------------------------------------------------------------
# greedy hypervolume subset selection
# ------------------------------------------------------------------------------------------------
def _hssp_2d(v: np.ndarray, idx: np.ndarray, k: int, ref: np.ndarray) -> np.ndarray:
    n = v.shape[0]
    order = np.arange(n)
    pts = v.copy()
    diag = np.repeat(ref[np.newaxis, :], n, axis=0)
    out = np.zeros(k, dtype=int)
    for i in range(k):
        contrib = np.prod(diag - pts, axis=-1)
        j = np.argmax(contrib)
        out[i] = idx[order[j]]
        chosen = pts[j].copy()
        keep = np.ones(n - i, dtype=bool)
        keep[j] = False
        order, diag, pts = order[keep], diag[keep], pts[keep]
        diag[:j, 0] = np.minimum(chosen[0], diag[:j, 0])
        diag[j:, 1] = np.minimum(chosen[1], diag[j:, 1])
    return out


def _lazy_update(contrib: np.ndarray, pts: np.ndarray, sel: np.ndarray, ref: np.ndarray, hv_sel: float) -> np.ndarray:
    if math.isinf(hv_sel):
        return np.full_like(contrib, np.inf)
    cap = np.maximum(pts[:, np.newaxis], sel[:-1])
    incl = np.prod(ref - pts, axis=1)
    inf_mask = np.isinf(incl)
    contrib = np.minimum(contrib, incl - np.prod(ref - cap[:, -1], axis=1))
    best = 0.0
    fast = pts.shape[1] <= 3
    for i in np.argsort(-contrib):
        if inf_mask[i]:
            best = contrib[i] = np.inf
            continue
        if contrib[i] < best:
            continue
        if fast:
            sel[-1] = pts[i].copy()
            contrib[i] = hypervolume(sel, ref, assume_pareto=True) - hv_sel
        else:
            contrib[i] = incl[i] - hypervolume(cap[i], ref)
        best = max(contrib[i], best)
    return contrib


def _hssp_unique(v: np.ndarray, idx: np.ndarray, k: int, ref: np.ndarray) -> np.ndarray:
    if not np.isfinite(ref).all():
        return idx[:k]
    if idx.size == k:
        return idx
    if v.shape[-1] == 2:
        return _hssp_2d(v, idx, k, ref)
    n, m = v.shape
    contrib = np.prod(ref - v, axis=-1)
    picks = np.zeros(k, dtype=int)
    sel = np.empty((k, m))
    pos = np.arange(n)
    hv = 0
    for t in range(k):
        j = int(np.argmax(contrib))
        hv += contrib[j]
        picks[t] = pos[j]
        sel[t] = v[j].copy()
        keep = np.ones(contrib.size, dtype=bool)
        keep[j] = False
        contrib, pos, v = contrib[keep], pos[keep], v[keep]
        if t == k - 1:
            break
        contrib = _lazy_update(contrib, v, sel[: t + 2], ref, hv)
    return idx[picks]


def solve_hssp(v: np.ndarray, idx: np.ndarray, k: int, ref: np.ndarray) -> np.ndarray:
    if k == idx.size:
        return idx
    u, first = np.unique(v, return_index=True, axis=0)
    nu = first.size
    if nu < k:
        chosen = np.zeros(idx.size, dtype=bool)
        chosen[first] = True
        dup = np.arange(idx.size)[~chosen]
        chosen[dup[: k - nu]] = True
        return idx[chosen]
    return idx[_hssp_unique(u, first, k, ref)]


# ------------------------------------------------------------------------------------------------
# MOTPE split and weights
# ------------------------------------------------------------------------------------------------
def reference_point(v: np.ndarray) -> np.ndarray:
    worst = np.max(v, axis=0)
    ref = np.maximum(1.1 * worst, 0.9 * worst)
    ref[ref == 0] = EPS
    return ref


def split_complete_mo(lvals: np.ndarray, n_below: int) -> np.ndarray:
    """Positions (into the COMPLETE-trial list) selected for the below set, ascending."""
    n = lvals.shape[0]
    n_below = min(n_below, n)
    if n_below == 0:
        return np.zeros(0, dtype=np.int64)
    if n_below == n:
        return np.arange(n, dtype=np.int64)
    ranks = nondomination_rank(lvals, n_below=n_below)
    uniq, counts = np.unique(ranks, return_counts=True)
    last = int(np.max(uniq[np.cumsum(counts) <= n_below], initial=-1))
    pos = np.arange(n)
    chosen = pos[ranks <= last]
    if chosen.size < n_below:
        tie = ranks == last + 1
        tv = lvals[tie]
        extra = solve_hssp(tv, pos[tie], n_below - chosen.size, reference_point(tv))
        chosen = np.append(chosen, extra)
    return np.sort(chosen).astype(np.int64)


def weights_below_mo(lvals: np.ndarray, feasible: np.ndarray | None = None) -> np.ndarray:
    """lvals: [n_below, M] loss vectors of the below trials in trial order."""
    n = lvals.shape[0]
    feas = np.ones(n, dtype=bool) if feasible is None else np.asarray(feasible, dtype=bool)
    w = np.where(feas, 1.0, EPS)
    nf = np.count_nonzero(feas)
    if nf <= 1:
        return w
    v = np.asarray(lvals, dtype=float)[feas]
    ref = reference_point(v)
    on_front = is_pareto_front(v, assume_unique_lexsorted=False)
    ps = v[on_front]
    hv = hypervolume(ps, ref, assume_pareto=True)
    if math.isinf(hv):
        return w
    loo = ~np.eye(ps.shape[0], dtype=bool)
    contrib = np.zeros(nf, dtype=float)
    if v.shape[1] <= 3:
        contrib[on_front] = [hv - hypervolume(ps[m], ref, assume_pareto=True) for m in loo]
    else:
        contrib[on_front] = np.prod(ref - ps, axis=-1)
        lim = np.maximum(ps, ps[:, np.newaxis])
        contrib[on_front] -= [hypervolume(lim[i, m], ref) for i, m in enumerate(loo)]
    w[feas] = np.maximum(contrib / max(np.max(contrib), EPS), EPS)
    return w
