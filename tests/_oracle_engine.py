"""``OracleEngine`` -- the ``TPEEngine`` interface answered by the CPU oracle (TEST INFRASTRUCTURE).

``B200TPESampler``'s host glue (trial log, device mirror, RNG hand-over, search spaces) is plain Python and can be
checked without a GPU: plugged into this engine, the sampler behind optuna's own ``Study`` must reproduce the
reference ``TPESampler`` trajectory.  On the GPU box the very same tests run against the CUDA engine.  The product
never imports this module (``B200TPESampler._engine_cls`` is ``TPEEngine``).
"""
from __future__ import annotations

import numpy as np

from oracle import motpe as mo
from oracle import tpe_oracle as orc


class _ReplayRng:
    """numpy RandomState look-alike that replays a given array of ``random_sample`` outputs in the order
    ``oracle.tpe_oracle.mixture_sample`` consumes them (choice / rand / uniform all map to random_sample)."""

    def __init__(self, u: np.ndarray) -> None:
        self.u, self.at = np.asarray(u, dtype=np.float64).ravel(), 0

    def random_sample(self, size):
        n = int(np.prod(size))
        out = self.u[self.at: self.at + n]
        assert out.size == n, "not enough uniforms staged"
        self.at += n
        return out.reshape(size)

    def rand(self, n):
        return self.random_sample((n,))

    def uniform(self, low=0, high=1, size=None):
        assert low == 0 and high == 1
        return self.random_sample(size)

    def choice(self, n, p, size):  # numpy/random/mtrand.pyx: cdf.searchsorted(random_sample, side="right")
        cdf = np.cumsum(p)
        cdf /= cdf[-1]
        return cdf.searchsorted(self.random_sample((size,)), side="right")


class OracleEngine:
    #: np.argsort's default kind, as the reference calls it (parzen_estimator.py:200): on the same machine the
    #: same tie order as the live reference.  The CUDA path sorts stably: StableOracleEngine.
    stable_sort = False

    def __init__(self, device: int = 0) -> None:
        self.device = device
        self.specs, self.params = [], []
        self.X = np.zeros((0, 0))
        self.cat = np.zeros(0, np.int8)
        self.key = np.zeros((0, 2))
        self.vals = None
        self.calls: list[tuple] = []  # (name, rows) of every history call, for the O(#changes) assertions
        self._staged = None
        self._rng_src = None

    # -- space / history ----------------------------------------------------------------------------
    def set_space(self, specs) -> None:
        self.specs = list(specs)
        self.params = []
        for s in specs:
            if s.kind == 2:
                self.params.append(orc.Param("cat", n_choices=s.n_choices, dist_table=s.dist_table))
            else:
                self.params.append(orc.Param("int" if s.kind == 1 else "float", s.low, s.high, s.step, s.log))
        self.n_params = len(specs)
        self.X = np.zeros((0, self.n_params))
        self.cat, self.key, self.vals = np.zeros(0, np.int8), np.zeros((0, 2)), None
        self.calls.append(("set_space", self.n_params))

    def set_history(self, X, category, key) -> None:
        self.X = np.array(X, dtype=np.float64).reshape(-1, self.n_params)
        self.cat = np.array(category, dtype=np.int8)
        self.key = np.array(key, dtype=np.float64).reshape(-1, 2)
        self.vals = None
        self.calls.append(("set_history", len(self.cat)))

    def update_history(self, X, category, key, at_row: int) -> None:
        X = np.asarray(X, dtype=np.float64).reshape(-1, self.n_params)
        n = X.shape[0]
        assert 0 <= at_row <= len(self.cat), "a write must continue the history"
        grow = at_row + n - len(self.cat)
        if grow > 0:
            self.X = np.concatenate([self.X, np.full((grow, self.n_params), np.nan)])
            self.cat = np.concatenate([self.cat, np.full(grow, 4, np.int8)])
            self.key = np.concatenate([self.key, np.zeros((grow, 2))])
            if self.vals is not None:
                self.vals = np.concatenate([self.vals, np.full((grow, self.vals.shape[1]), np.inf)])
        self.X[at_row: at_row + n] = X
        self.cat[at_row: at_row + n] = category
        self.key[at_row: at_row + n] = np.asarray(key).reshape(-1, 2)
        self.calls.append(("update_history", n))

    def append_history(self, X, category, key) -> None:
        self.update_history(X, category, key, len(self.cat))

    def set_values(self, values, at_row: int = 0, n_objectives=None) -> None:
        v = np.asarray(values, dtype=np.float64)
        m = int(n_objectives) if n_objectives is not None else v.shape[1]
        v = v.reshape(-1, m)
        if self.vals is None or self.vals.shape[1] != m:
            self.vals = np.full((len(self.cat), m), np.inf)
        assert at_row + v.shape[0] <= len(self.cat)
        self.vals[at_row: at_row + v.shape[0]] = v

    @property
    def history_size(self) -> int:
        return len(self.cat)

    # -- stages -----------------------------------------------------------------------------------------
    def prepare(self, cols, *, n_below, n_candidates, multivariate, prior_weight=1.0, magic_clip=True,
                endpoints=False):
        self._cols = [int(c) for c in cols]
        self._pc = len(self._cols)
        self._C = int(n_candidates)
        self._cfg = orc.Config(prior_weight=prior_weight, magic_clip=magic_clip, endpoints=endpoints,
                               multivariate=bool(multivariate), stable_sort=self.stable_sort)
        selector = None
        multi = self.vals is not None and self.vals.shape[1] >= 2
        if multi:
            selector = lambda idx, m: idx[mo.split_complete_mo(self.vals[idx], m)]  # noqa: E731
        self._below, self._above = orc.split_trials(self.cat, self.key, int(n_below), selector)
        self._sub = [self.params[c] for c in self._cols]
        self._obs_b, self._keep_b = orc.observations(self.X, self._below, self._cols)
        self._obs_a, _ = orc.observations(self.X, self._above, self._cols)
        self._multi = multi
        self._info = (len(self._below), self._obs_b.shape[0], self._obs_a.shape[0])
        return self._info

    def build(self, w_below=None, w_above=None) -> None:
        wb = None if w_below is None else np.asarray(w_below, dtype=np.float64)
        if wb is None and self._multi and len(self._below):
            feas = self.cat[self._below] != 2
            wb = mo.weights_below_mo(self.vals[self._below], feas)[self._keep_b]
        self._mix_b = orc.build_mixture(self._obs_b, self._sub, self._cfg, wb)
        self._mix_a = orc.build_mixture(self._obs_a, self._sub, self._cfg,
                                        None if w_above is None else np.asarray(w_above, dtype=np.float64))

    def stage_rng(self, rng, count: int, skip: int = 0, state=None) -> None:
        if rng is not None:
            self._rng_src = np.random.RandomState()
            self._rng_src.set_state(rng.get_state() if state is None else state)
        if skip:
            self._rng_src.random_sample(skip)
        self._staged = self._rng_src.random_sample(count)

    def finish_rng(self, rng) -> None:
        rng.set_state(self._rng_src.get_state())

    def sample_and_select(self, uniforms, n_asks: int = 1):
        u = self._staged if uniforms is None else np.asarray(uniforms, dtype=np.float64).ravel()
        per = u.size // n_asks
        x = np.empty((n_asks, self._pc))
        acq = np.empty(n_asks)
        best = np.empty(n_asks, dtype=np.int64)
        for a in range(n_asks):
            rng = _ReplayRng(u[a * per: (a + 1) * per])
            cand = orc.mixture_sample(self._mix_b, rng, self._C)
            assert rng.at == per
            ll, lg = orc.mixture_log_pdf(self._mix_b, cand), orc.mixture_log_pdf(self._mix_a, cand)
            self._last = (cand, ll, lg)
            score = ll - lg
            best[a] = int(np.argmax(score))
            acq[a] = score[best[a]]
            x[a] = cand[best[a]]
        return x, acq, best

    def stage_uniforms(self, uniforms):
        return np.asarray(uniforms, dtype=np.float64).ravel()

    def sample_and_select_async(self, uniforms, n_asks: int = 1) -> None:
        self._deferred = self.sample_and_select(uniforms, n_asks)

    def collect(self):
        out, self._deferred = self._deferred, None
        assert out is not None, "collect without sample_and_select_async"
        return out

    def rng_snapshot(self):
        return self._rng_src.get_state()

    def get_candidates(self):
        return self._last

    def get_uniforms(self, count):
        return np.asarray(self._staged[:count])

    def suggest(self, cols, uniforms, n_asks=1, w_below=None, w_above=None, **cfg):
        self.prepare(cols, **cfg)
        self.build(w_below, w_above)
        return self.sample_and_select(uniforms, n_asks)

    def suggest_univariate_batch(self, cols, uniforms, w_below=None, w_above=None, **cfg):
        """The per-parameter calls of one trial, one after the other (what the batched CUDA entry must equal)."""
        if self.vals is not None and self.vals.shape[1] >= 2:
            raise RuntimeError("not batchable: multi-objective history")
        if np.isnan(self.X[np.isin(self.cat, (0, 1, 2, 3))][:, list(cols)]).any():
            raise RuntimeError("not batchable: a selected parameter is absent from some trials")
        u = self._staged if uniforms is None else np.asarray(uniforms, dtype=np.float64).ravel()
        per = 2 * int(cfg["n_candidates"])
        x, acq, best = np.empty(len(cols)), np.empty(len(cols)), np.empty(len(cols), dtype=np.int64)
        self.calls.append(("univariate_batch", len(cols)))
        for j, c in enumerate(cols):
            self.prepare([c], **cfg)
            self.build(w_below, w_above)
            xj, aj, bj = self.sample_and_select(u[j * per: (j + 1) * per], 1)
            x[j], acq[j], best[j] = xj[0, 0], aj[0], bj[0]
        return x, acq, best

    # -- one suggestion over several ranks (tpe_set_kernel_shard / tpe_sample_and_partial / tpe_finish_from_partials) --
    def set_kernel_shard(self, rank: int, world: int) -> None:
        self._ks = (int(rank), int(world))

    def sample_and_partial_host(self, uniforms, n_asks: int = 1) -> np.ndarray:
        """(max, sum) per candidate of g(x) over this rank's slice of the above kernels; the prior kernel belongs to
        rank 0 (as in the library)."""
        assert n_asks == 1
        rank, world = getattr(self, "_ks", (0, 1))
        u = np.asarray(uniforms, dtype=np.float64).ravel()
        cand = orc.mixture_sample(self._mix_b, _ReplayRng(u), self._C)
        self._ks_cand, self._ks_ll = cand, orc.mixture_log_pdf(self._mix_b, cand)
        K = self._mix_a.weights.size
        chunk = -(-(K - 1) // world)
        idx = list(range(min(K - 1, rank * chunk), min(K - 1, (rank + 1) * chunk))) + ([K - 1] if rank == 0 else [])
        out = np.stack([np.full(self._C, -np.inf), np.zeros(self._C)], 1)
        if idx:
            m = self._mix_a
            sub = orc.Mixture(m.weights[idx], m.params, {j: w[idx] for j, w in m.cat_w.items()},
                              {j: v[idx] for j, v in m.mu.items()}, {j: v[idx] for j, v in m.sigma.items()})
            out[:, 0], out[:, 1] = orc.mixture_log_pdf(sub, cand), 1.0
        return out

    def finish_from_partials_host(self, allp: np.ndarray):
        m, s = allp[:, :, 0], allp[:, :, 1]
        top = m.max(axis=0)
        with np.errstate(divide="ignore", invalid="ignore"):
            lg = np.log((s * np.exp(m - top)).sum(axis=0)) + top
        score = self._ks_ll - lg
        self._last = (self._ks_cand, self._ks_ll, lg)
        best = int(np.argmax(score))
        return self._ks_cand[best][None], np.array([score[best]]), np.array([best])

    def suggest_univariate_batch_async(self, cols, uniforms, w_below=None, w_above=None, **cfg) -> None:
        self._uni_deferred = self.suggest_univariate_batch(cols, uniforms, w_below, w_above, **cfg)

    def collect_univariate(self):
        out, self._uni_deferred = self._uni_deferred, None
        assert out is not None, "collect_univariate without suggest_univariate_batch_async"
        return out

    def close(self) -> None:
        pass


class StableOracleEngine(OracleEngine):
    stable_sort = True
