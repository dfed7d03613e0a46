"""Host-side multi-GPU plumbing (optuna_b200/dist.py) on CPU: gloo backend, world_size 2."""
import os
import socket

import numpy as np
import pytest

from optuna_b200.dist import shard_asks


def test_shard_asks_partitions_exactly():
    for n in (0, 1, 7, 8, 8192, 8193):
        for world in (1, 2, 3, 8):
            spans = [shard_asks(n, world, r) for r in range(world)]
            assert sum(c for _, c in spans) == n
            at = 0
            for s, c in spans:
                assert s == at
                at += c
            sizes = [c for _, c in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank: int, world: int, port: int, out_dir: str) -> None:
    import torch.distributed as dist
    from optuna_b200.dist import broadcast_history, sharded_asks
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rs = np.random.RandomState(0)
    n, p, m = 1000, 5, 3
    X = rs.uniform(size=(n, p))
    X[3, 2] = np.nan
    cat = rs.randint(0, 3, size=n).astype(np.int8)
    key = rs.normal(size=(n, 2))
    vals = rs.normal(size=(n, m))
    if rank == 0:
        tX, tc, tk, tv = broadcast_history(X, cat, key, vals)
    else:
        tX, tc, tk, tv = broadcast_history(None, None, None, None)
    assert np.array_equal(tX.numpy(), X, equal_nan=True) and np.array_equal(tc.numpy(), cat)
    assert np.array_equal(tk.numpy(), key) and np.array_equal(tv.numpy(), vals)
    n_asks, per_ask = 37, 11
    U = np.random.RandomState(1).uniform(size=(n_asks, per_ask))

    def compute(u, count):  # stand-in for engine.sample_and_select: one row per ask
        assert u.shape == (count, per_ask)
        return np.stack([u.sum(1), u[:, 0]], 1)

    got = sharded_asks(n_asks, per_ask, U, compute)
    assert np.allclose(got, np.stack([U.sum(1), U[:, 0]], 1))
    np.save(os.path.join(out_dir, f"r{rank}.npy"), got)
    dist.destroy_process_group()


def test_broadcast_and_sharded_asks_gloo_world2(tmp_path):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    assert np.array_equal(a, b) and a.shape == (37, 2)


def _rng_worker(rank: int, world: int, port: int, out_dir: str) -> None:
    """sharded_asks_device_rng with the CPU oracle standing in for the CUDA engine (gloo): the ranks' blocks of
    asks concatenate to the batch one process computes, and every rank's generator ends in the same state."""
    import torch.distributed as dist
    from optuna_b200.dist import sharded_asks_device_rng
    from optuna_b200.engine import ParamSpec
    from tests._oracle_engine import OracleEngine
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rs = np.random.RandomState(3)
    n, P, C, n_asks = 60, 3, 8, 7
    X = rs.uniform(0, 1, (n, P))
    eng = OracleEngine()
    eng.set_space([ParamSpec(kind=0, low=0.0, high=1.0) for _ in range(P)])
    eng.set_history(X, np.zeros(n, np.int8), np.stack([((X - 0.4) ** 2).sum(1), np.zeros(n)], 1))
    eng.prepare(list(range(P)), n_below=6, n_candidates=C, multivariate=True)
    eng.build()
    per_ask = C * (1 + P)
    rng = np.random.RandomState(21)
    got = sharded_asks_device_rng(eng, rng, n_asks, per_ask, gather=True)
    ref_rng = np.random.RandomState(21)
    want, _, _ = eng.sample_and_select(ref_rng.random_sample(n_asks * per_ask), n_asks)
    assert np.array_equal(got, want)
    assert np.array_equal(rng.random_sample(5), ref_rng.random_sample(5))  # one generator that drew everything
    np.save(os.path.join(out_dir, f"g{rank}.npy"), got)
    dist.destroy_process_group()


def test_sharded_asks_with_engine_generated_uniforms_gloo_world2(tmp_path):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_rng_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert np.array_equal(np.load(tmp_path / "g0.npy"), np.load(tmp_path / "g1.npy"))


def _kshard_worker(rank: int, world: int, port: int, out_dir: str) -> None:
    """kernel_sharded_suggest with the CPU oracle standing in for the CUDA engine (gloo): g(x) summed over the ranks'
    slices of the above kernels equals the single-process suggestion on every rank."""
    import torch.distributed as dist
    from optuna_b200.dist import kernel_sharded_suggest
    from optuna_b200.engine import ParamSpec
    from tests._oracle_engine import OracleEngine
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rs = np.random.RandomState(5)
    n, P, C = 90, 4, 16
    X = rs.uniform(0, 1, (n, P))
    X[:, 3] = rs.randint(0, 3, n)
    eng = OracleEngine()
    eng.set_space([ParamSpec(kind=0, low=0.0, high=1.0) for _ in range(3)] + [ParamSpec(kind=2, n_choices=3)])
    eng.set_history(X, np.zeros(n, np.int8), np.stack([((X[:, :3] - 0.4) ** 2).sum(1), np.zeros(n)], 1))
    cfg = dict(n_below=8, n_candidates=C, multivariate=True)
    u = np.random.RandomState(9).random_sample(C * (1 + 1 + 3))
    x, acq, best = kernel_sharded_suggest(eng, list(range(P)), u, 1, **cfg)
    want = OracleEngine()
    want.set_space(eng.specs)
    want.set_history(X, np.zeros(n, np.int8), eng.key)
    wx, wacq, wbest = want.suggest(list(range(P)), u, 1, **cfg)
    assert np.array_equal(x, wx) and best[0] == wbest[0] and abs(acq[0] - wacq[0]) < 1e-12
    np.save(os.path.join(out_dir, f"k{rank}.npy"), np.concatenate([x.ravel(), acq]))
    dist.destroy_process_group()


def test_kernel_sharded_suggestion_gloo_world2(tmp_path):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_kshard_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = np.load(tmp_path / "k0.npy"), np.load(tmp_path / "k1.npy")
    assert np.array_equal(a, b)
