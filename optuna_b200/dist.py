"""Multi-GPU plumbing for a batch of concurrent asks (BASELINE config 5).

One process per GPU (torchrun), one TPEEngine per process.  Concurrent asks against a frozen history
are independent -- same split, same mixtures, different uniforms -- so the batch is block-partitioned
over ranks with NO collective on the data path:

  1. rank 0 holds the history; ONE broadcast replicates it (NCCL over NVLink for CUDA tensors);
  2. every rank builds the two estimators redundantly (cheaper than shipping mu / sigma);
  3. rank r evaluates asks [start_r, start_r + count_r) with its slice of the host-drawn uniforms;
  4. results are gathered (all_gather of [count, P] doubles) or simply read per rank.

torch.distributed is used for the plumbing only; the library never sees a torch type.
"""
from __future__ import annotations

from typing import Callable

import numpy as np


def shard_asks(n_asks: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block partition: (start, count) of rank's share; sizes differ by at most one."""
    base, extra = divmod(int(n_asks), int(world))
    count = base + (1 if rank < extra else 0)
    start = rank * base + min(rank, extra)
    return start, count


def broadcast_history(X, category, key, values=None, *, src: int = 0, device=None):
    """Replicate the history arrays from `src` to every rank with one broadcast per dtype.

    On `src` the arguments are numpy arrays; on the other ranks they may be None.  Returns torch
    tensors (on `device` when given, else CPU) as (X, category, key, values-or-None).  The float
    arrays travel as ONE flat fp64 tensor so that a 100k x 32 history is a single 26 MB collective.
    """
    import torch
    import torch.distributed as dist

    rank = dist.get_rank()
    dev = torch.device("cpu") if device is None else device
    meta = torch.zeros(4, dtype=torch.int64, device=dev)
    if rank == src:
        X = np.ascontiguousarray(X, dtype=np.float64)
        m = 0 if values is None else np.asarray(values).reshape(X.shape[0], -1).shape[1]
        meta = torch.tensor([X.shape[0], X.shape[1], m, 0], dtype=torch.int64, device=dev)
    dist.broadcast(meta, src)
    n, p, m = int(meta[0]), int(meta[1]), int(meta[2])
    flat = torch.empty(n * (p + 2 + m), dtype=torch.float64, device=dev)
    cat = torch.empty(n, dtype=torch.int8, device=dev)
    if rank == src:
        parts = [X.ravel(), np.ascontiguousarray(key, dtype=np.float64).ravel()]
        if m:
            parts.append(np.ascontiguousarray(values, dtype=np.float64).ravel())
        flat.copy_(torch.from_numpy(np.concatenate(parts)))
        cat.copy_(torch.from_numpy(np.ascontiguousarray(category, dtype=np.int8)))
    dist.broadcast(flat, src)
    dist.broadcast(cat, src)
    tX = flat[: n * p].view(n, p)
    tk = flat[n * p: n * (p + 2)].view(n, 2)
    tv = flat[n * (p + 2):].view(n, m) if m else None
    return tX, cat, tk, tv


def adopt_history(engine, tX, tcat, tkey, tvals=None) -> None:
    """Hand broadcast CUDA tensors to the engine (device-to-device copy inside the library)."""
    import torch
    assert tX.is_cuda and tX.is_contiguous() and tkey.is_contiguous() and tcat.is_contiguous()
    torch.cuda.synchronize(tX.device)
    engine.set_history_device(tX.data_ptr(), tcat.data_ptr(), tkey.data_ptr(), tX.shape[0],
                              np.isnan(tX.sum(dim=0).cpu().numpy()).astype(np.uint8))
    if tvals is not None:
        engine.set_values(tvals.cpu().numpy(), 0)


def sharded_asks(n_asks: int, per_ask: int, uniforms, compute: Callable[[np.ndarray, int], np.ndarray],
                 gather: bool = True):
    """Run `compute(uniforms_slice, count) -> [count, P]` on this rank's block of asks and (optionally)
    all_gather the results in ask order.  `uniforms` is the full [n_asks, per_ask] host array (every
    rank draws the same stream from the shared seed, or rank 0 broadcasts it)."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(), dist.get_rank()
    start, count = shard_asks(n_asks, world, rank)
    u = np.asarray(uniforms, dtype=np.float64).reshape(n_asks, per_ask)[start:start + count]
    mine = np.asarray(compute(u, count), dtype=np.float64).reshape(count, -1)
    if not gather:
        return mine
    width = mine.shape[1]
    most = shard_asks(n_asks, world, 0)[1]
    pad = torch.zeros((most, width), dtype=torch.float64)
    pad[:count] = torch.from_numpy(mine)
    backend = dist.get_backend()
    if backend == "nccl":
        pad = pad.cuda()
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad)
    rows = [outs[r][: shard_asks(n_asks, world, r)[1]].cpu().numpy() for r in range(world)]
    return np.concatenate(rows, axis=0)


class _DeviceArray:
    """A raw device pointer as something ``torch.as_tensor`` wraps without a copy (__cuda_array_interface__)."""

    def __init__(self, ptr: int, shape: tuple, typestr: str) -> None:
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (int(ptr), False), "version": 3}


def device_view(ptr: int, shape: tuple, typestr: str, device):
    import torch
    return torch.as_tensor(_DeviceArray(ptr, shape, typestr), device=device)


def sharded_asks_device_rng(engine, rng: np.random.RandomState, n_asks: int, per_ask: int, gather: bool = True,
                            timing: dict | None = None):
    """Like `sharded_asks`, but nothing touches the host between the generator state going in and the gathered
    suggestions coming out:

      * every rank starts from the same generator state (shared seed) and the library JUMPS to the uniforms of its
        own asks (multi-CTA MT19937 with jump-ahead polynomials: no walk over the earlier ranks' prefix);
      * the results stay in the context's device buffer (tpe_sample_and_select with out_x = NULL) and are gathered
        by NCCL straight from it (all_gather_into_tensor on [most, P] blocks);
      * the rank owning the last block ends in the state after ALL n_asks draws and broadcasts it device to device
        into every rank's generator state (2.5 KB), so that every rank continues as if one process had drawn
        everything; `rng` is brought up to date from there.

    The engine must have been prepared and built; returns [n_asks, P] (gather) or this rank's [count, P]."""
    import time

    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(), dist.get_rank()
    start, count = shard_asks(n_asks, world, rank)
    nccl = dist.get_backend() == "nccl"
    pc = int(engine._pc)
    t0 = time.perf_counter()
    if not nccl:  # gloo (CPU tests of the plumbing): same partition, host tensors
        if count > 0:
            engine.stage_rng(rng, count * per_ask, skip=start * per_ask)
            mine, _, _ = engine.sample_and_select(None, count)
            engine.finish_rng(rng)
        else:
            mine = np.zeros((0, pc))
        owner = max(r for r in range(world) if shard_asks(n_asks, world, r)[1] > 0) if n_asks > 0 else 0
        st = rng.get_state()
        pack = torch.empty(625, dtype=torch.int64)
        if rank == owner:
            pack[:624] = torch.from_numpy(np.asarray(st[1], dtype=np.int64))
            pack[624] = int(st[2])
        dist.broadcast(pack, src=owner)
        pack = pack.numpy()
        rng.set_state((st[0], pack[:624].astype(np.uint32), int(pack[624]), st[3], st[4]))
        mine = np.asarray(mine, dtype=np.float64).reshape(count, -1)
        if not gather:
            return mine
        most = shard_asks(n_asks, world, 0)[1]
        pad = torch.zeros((most, pc), dtype=torch.float64)
        pad[:count] = torch.from_numpy(mine)
        outs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(outs, pad)
        return np.concatenate([outs[r][: shard_asks(n_asks, world, r)[1]].numpy() for r in range(world)], axis=0)

    dev = torch.device("cuda", engine.device)
    assert count > 0, "more ranks than asks"
    engine.stage_rng(rng, count * per_ask, skip=start * per_ask)
    px = engine.sample_and_select_device(count)          # returns when this rank's asks are done (stream sync)
    t1 = time.perf_counter()
    owner = max(r for r in range(world) if shard_asks(n_asks, world, r)[1] > 0)
    state = device_view(engine.rng_state_device(), (625,), "<u4", dev).view(torch.int32)
    dist.broadcast(state, src=owner)
    mine = device_view(px, (count, pc), "<f8", dev)
    out = None
    if gather:
        most = shard_asks(n_asks, world, 0)[1]
        if count == most:
            block = mine
        else:
            block = torch.zeros((most, pc), dtype=torch.float64, device=dev)
            block[:count] = mine
        allb = torch.empty((world, most, pc), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allb.view(-1), block.reshape(-1))
        if n_asks == world * most:
            out = allb.view(n_asks, pc).cpu().numpy()
        else:
            out = np.concatenate([allb[r, : shard_asks(n_asks, world, r)[1]].cpu().numpy() for r in range(world)], axis=0)
    else:
        out = mine.cpu().numpy()
    torch.cuda.current_stream(dev).synchronize()      # the broadcast has landed in the engine's generator state
    engine.finish_rng(rng)
    if timing is not None:
        timing["compute_s"] = t1 - t0
        timing["collectives_s"] = time.perf_counter() - t1
    return out


def kernel_sharded_suggest(engine, cols, uniforms, n_asks: int = 1, gather=None, **cfg):
    """ONE suggestion (or batch of asks) evaluated by all ranks together: every rank holds the whole history and
    builds both estimators, evaluates g(x) over its slice of the above kernels (``tpe_set_kernel_shard``), the
    per-candidate (max, sum) partials are all-gathered device to device (16 bytes per candidate and rank) and every
    rank finishes with the same argmax (``tpe_finish_from_partials``).  SURVEY.md section 8e, the alternative for a
    single ask: the grid of one config-2 suggestion shrinks by the number of GPUs.

    ``gather(local: torch.Tensor) -> torch.Tensor [world, ...]`` defaults to ``dist.all_gather_into_tensor`` (NCCL);
    tests pass their own to emulate several ranks on one device."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    if gather is None and world == 1:
        return engine.suggest(cols, uniforms, n_asks, **cfg)
    if gather is None:
        engine.set_kernel_shard(rank, world)
    engine.prepare(cols, **cfg)
    engine.build()
    if hasattr(engine, "sample_and_partial_host"):       # an engine without a device (tests: the CPU oracle, gloo)
        mine_h = torch.from_numpy(np.ascontiguousarray(engine.sample_and_partial_host(uniforms, n_asks)))
        if gather is None:
            blocks = [torch.empty_like(mine_h) for _ in range(world)]
            dist.all_gather(blocks, mine_h)
            allh = torch.stack(blocks)
        else:
            allh = gather(mine_h)
        return engine.finish_from_partials_host(allh.numpy())
    ptr, stride = engine.sample_and_partial(uniforms, n_asks)
    dev = torch.device("cuda", engine.device)
    mine = device_view(ptr, (stride, 2), "<f8", dev)
    if gather is None:
        allp = torch.empty((world, stride, 2), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allp.view(-1), mine.reshape(-1))
    else:
        allp = gather(mine)
    torch.cuda.synchronize(dev)
    return engine.finish_from_partials(allp.data_ptr(), allp.shape[0])
