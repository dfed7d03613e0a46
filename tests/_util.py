"""Shared helpers for the parity tests (decode golden fixtures into oracle inputs)."""
from __future__ import annotations

import os

import numpy as np

from oracle import tpe_oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name: str):
    return np.load(os.path.join(GOLD, name))


def decode_space(arr: np.ndarray) -> list[orc.Param]:
    out = []
    for kind, low, high, step, log, nch in arr:
        if kind == 2:
            out.append(orc.Param("cat", n_choices=int(nch)))
        else:
            out.append(orc.Param("int" if kind == 1 else "float", float(low), float(high),
                                 None if np.isnan(step) else float(step), bool(log)))
    return out


def mixture_from_gold(g, prefix: str, params) -> orc.Mixture:
    mix = orc.Mixture(weights=g[prefix + "w"], params=list(params))
    for j, p in enumerate(params):
        if p.is_cat:
            mix.cat_w[j] = g[f"{prefix}cat{j}"]
        else:
            mix.mu[j] = g[f"{prefix}mu{j}"]
            mix.sigma[j] = g[f"{prefix}sigma{j}"]
    return mix


# ---- helpers for the CUDA-path tests ---------------------------------------------------------
def specs_from_space(arr: np.ndarray):
    from optuna_b200.engine import ParamSpec
    out = []
    for kind, low, high, step, log, nch in arr:
        if kind == 2:
            out.append(ParamSpec(kind=2, n_choices=int(nch)))
        else:
            out.append(ParamSpec(kind=int(kind), low=float(low), high=float(high),
                                 step=None if np.isnan(step) else float(step), log=bool(log)))
    return out


def draw_uniforms(rng: np.random.RandomState, C: int, ncat: int, nnum: int) -> np.ndarray:
    """Uniforms of one ask in the reference's consumption order
    (probability_distributions.py:87,100,138-144; _truncnorm.py:282)."""
    parts = [rng.rand(C)]
    for _ in range(ncat):
        parts.append(rng.rand(C))
    if nnum:
        parts.append(rng.uniform(low=0, high=1, size=(nnum, C)).ravel())
    return np.concatenate(parts)


def kinds_of(params) -> tuple[int, int]:
    ncat = sum(1 for p in params if p.is_cat)
    return ncat, len(params) - ncat
