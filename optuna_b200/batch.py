"""Batched ``study.ask()`` (BASELINE config 5): n trials whose relative parameters come from ONE
device call, optionally sharded over the GPUs of a box (optuna_b200/dist.py)."""
from __future__ import annotations


def _frozen_of(trial):
    f = getattr(trial, "_cached_frozen_trial", None)  # optuna.trial.Trial
    return f if f is not None else getattr(trial, "_frozen")  # optuna_b200.mini.Trial


def ask_batch(study, n_asks: int) -> list:
    """``[study.ask() for _ in range(n_asks)]`` with the TPE suggestions of all asks computed together.

    Works with optuna's Study and with optuna_b200.mini.Study; the study's sampler must be a
    B200TPESampler(multivariate=True, constant_liar=False).  Parameters outside the relative search
    space still go through ``sample_independent`` when the objective asks for them."""
    sampler = study.sampler
    trials = [study.ask() for _ in range(n_asks)]
    if not trials:
        return trials
    space = sampler.infer_relative_search_space(study, _frozen_of(trials[0]))
    params = sampler.sample_relative_batch(study, space, n_asks)
    for t, p in zip(trials, params):
        t.relative_search_space = space
        t._relative_params = p
    return trials
