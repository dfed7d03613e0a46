#!/usr/bin/env python
"""bench.py -- TPE suggestions/sec at a 100k-trial history (BASELINE.json metric).

Workload (config.workload "c2"): N = 100 000 COMPLETE trials x P = 32 FloatDistribution(0, 1)
params, X = RandomState(0).uniform, loss = sum((x - 0.5)^2), n_ei_candidates C = 4096,
multivariate TPE, default gamma (n_below = 25) -- SURVEY.md section 8d.  One *step* = one
suggestion = split + both Parzen-estimator builds + C candidate draws + log l(x) and log g(x) over
the C x K x P grid + argmax.  Nothing is cached between steps.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

``value``  device-resident throughput: the K timed steps go through the C ABI (tpe_suggest) with the
           history already in HBM; time = wall clock between device synchronisations around the K
           steps (each call ends with a stream sync), max over ranks; the CUDA-event time of the same
           steps is reported as ``device_ms_per_step``.
``e2e``    the same metric through the plugin a user calls -- B200TPESampler.sample_relative(study,
           trial, search_space) on a 100k-trial study: host uniforms (numpy RandomState, the
           reference's RNG order) -> H2D, result D2H, to_external_repr.
``--impl reference``  the reference algorithm on the host cores (oracle port of the NumPy path; the
           reference is pure Python and does not travel to the GPU box), bounded sample per step.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_TRIALS, N_PARAMS, N_CAND = 100_000, 32, 4096
METRIC = "TPE suggestions/sec at 100k-trial history, 32 params"
UNIT = "suggestions/s"


def synthetic_history(n=N_TRIALS, p=N_PARAMS):
    rs = np.random.RandomState(0)
    X = rs.uniform(0, 1, (n, p))
    loss = ((X - 0.5) ** 2).sum(1)
    return X, loss


def algorithmic_bytes(n=N_TRIALS, p=N_PARAMS, c=N_CAND, m=1) -> int:
    """every input read once, every output written once (multivariate)."""
    return 8 * (n * p + n * m + n) + 16 * c * p + 16 * c


def algorithmic_flops(n=N_TRIALS, p=N_PARAMS, c=N_CAND, n_below=25) -> float:
    """C * (K_below + K_above) * (4 P + 25)."""
    return float(c) * (n + 2) * (4 * p + 25)


def executed_flops(n=N_TRIALS, p=N_PARAMS, c=N_CAND) -> float:
    """fp64 flops the g(x) grid kernel issues: C x K_above x P fma on the tensor-core path."""
    return 2.0 * c * (n - 25) * p


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (NVML, ~1 kHz; nvidia-smi
    subprocesses are too slow for a 50 ms region and remain only as a fallback)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int) -> None:
        self.index = index
        self.sm: list[float] = []
        self.max_sm = None
        self.reasons: set[str] = set()
        self._stop = threading.Event()
        self._t: threading.Thread | None = None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None

    def _poll_nvml(self) -> None:
        n = self._nvml
        bits = {"hw_slowdown": getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)))
                r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.002)

    def _poll_smi(self) -> None:
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    c = [v.strip() for v in line.split(",")]
                    self.sm.append(float(c[0]))
                    self.max_sm = float(c[1])
                    for i, nm in enumerate(names):
                        if len(c) > 3 + i and c[3 + i].lower().startswith("active"):
                            self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self) -> None:
        self._t = threading.Thread(target=self._poll_nvml if self._nvml else self._poll_smi, daemon=True)
        self._t.start()

    def stop(self) -> dict:
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_sm,
                "reasons": sorted(self.reasons), "samples": len(self.sm),
                "source": "nvml" if self._nvml else "nvidia-smi"}


def measured_peaks() -> tuple[float, str]:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "MEASURED_PEAKS.json"
        except Exception:
            pass
    return 6650.0, "fallback"


# -------------------------------------------------------------------------------------------------
# CPU legs (oracle port of the reference's NumPy path)
# -------------------------------------------------------------------------------------------------
def _pool_eval(ma, mb, rows):
    from oracle import tpe_oracle as orc
    return orc.mixture_log_pdf(mb, rows).sum() + orc.mixture_log_pdf(ma, rows).sum()


def cpu_step(X, loss, procs: int, pool=None, ra: int = 1, rb: int = 3) -> dict:
    """One bounded sample of a reference suggestion (oracle port of the NumPy path).

    split + both estimator builds + all 4096 candidate draws are executed in full.  log_pdf is
    evaluated on `ra` and then on `rb` candidates per process: the difference gives the
    per-candidate slope, the remainder the per-call fixed cost (the (K, P) normalisers), and
    full_s = fixed + intercept + slope * 4096 is what one un-chunked, all-cores evaluation would
    take if its 105 GB temporaries fitted in memory -- the most favourable reading for the CPU.
    """
    from oracle import tpe_oracle as orc
    n, p = X.shape
    cat = np.zeros(n, np.int8)
    key = np.stack([loss, np.zeros(n)], 1)
    params = [orc.Param("float", 0.0, 1.0) for _ in range(p)]
    cfg = orc.Config(multivariate=True)
    t0 = time.perf_counter()
    below, above = orc.split_trials(cat, key, orc.default_gamma(n))
    mb = orc.build_mixture(X[below], params, cfg)
    ma = orc.build_mixture(X[above], params, cfg)
    t1 = time.perf_counter()
    cand = orc.mixture_sample(mb, np.random.RandomState(1), N_CAND)
    t2 = time.perf_counter()

    def timed(rows_per_proc: int) -> float:
        s = time.perf_counter()
        if pool is None:
            _pool_eval(ma, mb, cand[:rows_per_proc])
        else:
            chunks = [cand[i * rows_per_proc:(i + 1) * rows_per_proc] for i in range(procs)]
            pool.starmap(_pool_eval, [(ma, mb, c) for c in chunks])
        return time.perf_counter() - s

    ta, tb = timed(ra), timed(rb)
    slope = max(tb - ta, 1e-9) / ((rb - ra) * procs)  # seconds per candidate with all processes busy
    intercept = max(ta - slope * ra * procs, 0.0)
    fixed = t2 - t0
    return {"build_s": t1 - t0, "sample_s": t2 - t1, "logpdf_s": ta + tb, "slope_s": slope,
            "intercept_s": intercept, "full_s": fixed + intercept + slope * N_CAND}


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    X, loss = synthetic_history()
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 32))
    ctx = mp.get_context("fork")
    times = []
    with ctx.Pool(procs) as pool:
        for i in range(args.warmup + args.steps):
            r = cpu_step(X, loss, procs, pool)
            if i >= args.warmup:
                times.append(r)
    full = float(np.mean([r["full_s"] for r in times]))
    val = 1.0 / full
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": full * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "c2: N=100000 trials x P=32 float params, C=4096 candidates, multivariate TPE",
                   "n_trials": N_TRIALS, "n_params": N_PARAMS, "n_ei_candidates": N_CAND},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": procs, "kind": "port",
                         "sample": (f"oracle port of the NumPy path, {procs} processes: split + both estimator builds + "
                                    f"4096 candidate draws in full; log_pdf timed on 1 and on 3 candidates per process, "
                                    "per-candidate slope x 4096 + per-call fixed cost")},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "stages_s": {k: float(np.mean([r[k] for r in times]))
                     for k in ("build_s", "sample_s", "logpdf_s", "slope_s", "intercept_s")},
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_leg(X, loss) -> dict:
    """Single-process oracle timing on a bounded sample (what a user of the reference gets: NumPy
    elementwise kernels are single-threaded)."""
    r = cpu_step(X, loss, procs=1, pool=None, ra=2, rb=6)
    return {"value": 1.0 / r["full_s"], "unit": UNIT, "cores": 1, "kind": "port",
            "sample": ("oracle port, 1 process (NumPy elementwise kernels are single-threaded): split + builds + "
                       f"4096 draws in full ({r['build_s']:.1f}s + {r['sample_s']:.1f}s); log_pdf timed on 2 and 6 "
                       f"candidates ({r['logpdf_s']:.1f}s): {r['slope_s']:.2f}s per candidate x 4096 + "
                       f"{r['intercept_s']:.1f}s per call -> {r['full_s']:.0f}s per suggestion")}


# -------------------------------------------------------------------------------------------------
# GPU arm
# -------------------------------------------------------------------------------------------------
def run_b200(args) -> None:
    import torch
    import torch.distributed as dist

    from optuna_b200 import B200TPESampler, ParamSpec, TPEEngine, mini

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    X, loss = synthetic_history()
    cat = np.zeros(N_TRIALS, np.int8)
    key = np.stack([loss, np.zeros(N_TRIALS)], 1)
    eng = TPEEngine(local)
    eng.set_space([ParamSpec(kind=0, low=0.0, high=1.0) for _ in range(N_PARAMS)])
    if world > 1:
        # frozen history: rank 0 uploads, ONE NCCL broadcast over NVLink, every rank adopts the buffer
        dev = torch.device("cuda", local)
        tX = torch.empty((N_TRIALS, N_PARAMS), dtype=torch.float64, device=dev)
        tk = torch.empty((N_TRIALS, 2), dtype=torch.float64, device=dev)
        tc = torch.empty((N_TRIALS,), dtype=torch.int8, device=dev)
        if rank == 0:
            tX.copy_(torch.from_numpy(X))
            tk.copy_(torch.from_numpy(key))
            tc.copy_(torch.from_numpy(cat))
        flat = torch.cat([tX.view(-1), tk.view(-1)])
        dist.broadcast(flat, 0)
        dist.broadcast(tc, 0)
        tX = flat[: N_TRIALS * N_PARAMS].view(N_TRIALS, N_PARAMS).contiguous()
        tk = flat[N_TRIALS * N_PARAMS:].view(N_TRIALS, 2).contiguous()
        torch.cuda.synchronize()
        eng.set_history_device(tX.data_ptr(), tc.data_ptr(), tk.data_ptr(), N_TRIALS, np.zeros(N_PARAMS, np.uint8))
    else:
        eng.set_history(X, cat, key)

    cols = list(range(N_PARAMS))
    cfg = dict(n_below=min(math.ceil(0.1 * N_TRIALS), 25), n_candidates=N_CAND, multivariate=True)
    per_ask = N_CAND * (1 + N_PARAMS)
    rng = np.random.RandomState(1000 + rank)
    total = args.warmup + args.steps
    U = torch.empty((total, per_ask), dtype=torch.float64).pin_memory()
    U.numpy()[:] = rng.random_sample((total, per_ask))
    Unp = U.numpy()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        eng.suggest(cols, Unp[i], 1, **cfg)
    clocks = ClockSampler(local)
    barrier()
    if rank == 0:
        clocks.start()
    stage = np.zeros(9)
    launches = 0
    t0 = time.perf_counter()
    for i in range(args.warmup, total):
        eng.suggest(cols, Unp[i], 1, **cfg)
        ms, nl = eng.last_timing()
        stage += ms
        launches += nl
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    barrier()
    clk = clocks.stop() if rank == 0 else {}
    kernel_name = eng.last_logpdf_kernel()
    wall = t1 - t0
    dev_ms = float(stage[8])
    if world > 1:
        t = torch.tensor([wall, dev_ms], dtype=torch.float64, device=torch.device("cuda", local))
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall, dev_ms = float(t[0]), float(t[1])
    value = world * args.steps / wall

    # ---- end to end through the sampler plugin (host buffers, copies inside the timed region) ----
    space = {f"x{j:02d}": mini.FloatDistribution(0.0, 1.0) for j in range(N_PARAMS)}
    sampler = B200TPESampler(seed=1 + rank, n_ei_candidates=N_CAND, multivariate=True, device=local)
    study = mini.create_study(sampler=sampler)
    names = list(space)
    trials = []
    for i in range(N_TRIALS):
        t = mini.FrozenTrial(i, mini.TrialState.COMPLETE, value=float(loss[i]),
                             params=dict(zip(names, X[i].tolist())), distributions=space)
        trials.append(t)
    study._storage.trials = trials
    frozen = mini.FrozenTrial(N_TRIALS, mini.TrialState.RUNNING)
    e2e_steps = max(3, min(args.steps, 20))
    for _ in range(max(2, min(args.warmup, 3))):
        sampler.sample_relative(study, frozen, space)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        out = sampler.sample_relative(study, frozen, space)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    barrier()
    e2e_wall = t1 - t0
    if world > 1:
        t = torch.tensor([e2e_wall], dtype=torch.float64, device=torch.device("cuda", local))
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_wall = float(t[0])
    e2e_value = world * e2e_steps / e2e_wall
    assert len(out) == N_PARAMS and all(0.0 <= v <= 1.0 for v in out.values())

    # ---- extras (rank 0, N=1 only): other shapes of the same path, not the headline ------------------
    extras = {}
    if rank == 0 and world == 1 and not args.no_extras:
        # cold suggestion: history upload (25.6 MB H2D) + everything else
        t0 = time.perf_counter()
        eng.set_history(X, cat, key)
        eng.suggest(cols, Unp[0], 1, **cfg)
        extras["cold_suggestion_ms"] = (time.perf_counter() - t0) * 1e3
        # univariate TPE (the reference default): 32 sample_independent-style calls per trial
        ucfg = dict(cfg, multivariate=False)
        ru = np.random.RandomState(5)
        for rep in range(2):
            # a new trial = a new history version (zero-row append): the split is shared by the 32 calls
            # of one trial, never across trials
            eng.append_history(np.zeros((0, N_PARAMS)), np.zeros(0, np.int8), np.zeros((0, 2)))
            t0 = time.perf_counter()
            for j in range(N_PARAMS):
                eng.suggest([j], ru.random_sample(N_CAND * 2), 1, **ucfg)
            uni = time.perf_counter() - t0
        extras["univariate_trial_ms"] = uni * 1e3
        extras["univariate_suggestions_per_s"] = 1.0 / uni
        # config-5 shape: 8192 concurrent asks with the default n_ei_candidates = 24, one device call
        bcfg = dict(cfg, n_candidates=24)
        n_asks = 8192
        ub = np.random.RandomState(6).random_sample(n_asks * 24 * (1 + N_PARAMS))
        for rep in range(2):
            t0 = time.perf_counter()
            eng.suggest(cols, ub, n_asks, **bcfg)
            bt = time.perf_counter() - t0
        extras["batched_asks"] = {"n_asks": n_asks, "n_ei_candidates": 24, "ms": bt * 1e3,
                                  "suggestions_per_s": n_asks / bt,
                                  "note": "tpe_suggest with 52 MB of host-drawn uniforms uploaded inside the call"}
        # the same batch end to end through the plugin: uniforms generated on the device (MT19937 stream
        # of the sampler's RandomState), results converted to parameter dicts
        bs = B200TPESampler(seed=3, n_ei_candidates=24, multivariate=True, device=local)
        bstudy = mini.create_study(sampler=bs)
        bstudy._storage.trials = trials
        for rep in range(2):
            t0 = time.perf_counter()
            res = bs.sample_relative_batch(bstudy, space, n_asks)
            be = time.perf_counter() - t0
        assert len(res) == n_asks and len(res[0]) == N_PARAMS
        extras["batched_asks_e2e"] = {"n_asks": n_asks, "n_ei_candidates": 24, "ms": be * 1e3,
                                      "suggestions_per_s": n_asks / be,
                                      "path": "B200TPESampler.sample_relative_batch (device MT19937)"}
        bs.close()
    # ---- config 5: 8192 concurrent asks (default n_ei_candidates = 24) sharded over the ranks -----------
    # every rank evaluates its block of asks on uniforms its GPU generates from the shared generator state
    # (MT19937 stream of one sampler consumed sequentially, the reference's semantics); results all-gathered
    config5 = None
    if not args.no_extras:
        from optuna_b200.dist import shard_asks, sharded_asks_device_rng
        n_asks5, c5 = 8192, 24
        per_ask5 = c5 * (1 + N_PARAMS)
        cfg5 = dict(cfg, n_candidates=c5)
        for rep in range(2):
            rng5 = np.random.RandomState(6)
            barrier()
            t0 = time.perf_counter()
            eng.prepare(cols, **cfg5)
            eng.build()
            if world > 1:
                res5 = sharded_asks_device_rng(eng, rng5, n_asks5, per_ask5, gather=True)
            else:
                eng.stage_rng(rng5, n_asks5 * per_ask5)
                res5, _, _ = eng.sample_and_select(None, n_asks5)
                eng.finish_rng(rng5)
            torch.cuda.synchronize()
            barrier()
            dt5 = time.perf_counter() - t0
        assert res5.shape == (n_asks5, N_PARAMS)
        if world > 1:
            t = torch.tensor([dt5], dtype=torch.float64, device=torch.device("cuda", local))
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt5 = float(t[0])
        config5 = {"n_asks": n_asks5, "n_ei_candidates": c5, "ms": dt5 * 1e3, "suggestions_per_s": n_asks5 / dt5,
                   "asks_per_rank": shard_asks(n_asks5, world, 0)[1],
                   "path": "tpe_prepare / tpe_build / tpe_stage_uniforms_mt19937(skip) / tpe_sample_and_select per "
                           "rank + all_gather of the [8192, 32] results"}
    if rank == 0:
        peak, peak_src = measured_peaks()
        k_ms = float(stage[5]) / args.steps  # main log-density kernel under g(x)
        ach = algorithmic_bytes() / (k_ms * 1e-3) / 1e9
        try:
            fp64_peak = eng.probe_fp64_tflops()
        except Exception:
            fp64_peak = None
        fl = algorithmic_flops()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "c2: N=100000 trials x P=32 float params, C=4096 candidates, multivariate TPE",
                       "n_trials": N_TRIALS, "n_params": N_PARAMS, "n_ei_candidates": N_CAND,
                       "parallelism": f"{world} independent asks per step, history replicated by one NCCL broadcast",
                       "l2": "inputs larger than L2: the g(x) kernel table is 51 MB/launch and every step rebuilds it "
                             "(126 MB L2; split/build kernels in between evict it)"},
            "device_ms_per_step": dev_ms / args.steps,
            "stage_ms": {k: float(v) / args.steps for k, v in zip(
                ["split", "build", "h2d", "sample", "logpdf_below", "logpdf_above_main", "logpdf_fixup", "select",
                 "span"], stage)},
            "gpu_launches": int(launches),
            "clocks": clk,
            # inputs of one ask: the generator state (624 words + position; the uniforms themselves are
            # produced on the device by the same MT19937), the column list and the config struct
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 625 * 4 + N_PARAMS * 4 + 32,
                    "d2h_bytes_per_step": N_PARAMS * 8 + 16 + 24 + 625 * 4, "steps": e2e_steps,
                    "path": "B200TPESampler.sample_relative -> ctypes -> tpe_prepare / tpe_stage_uniforms_mt19937 / "
                            "tpe_build / tpe_sample_and_select / tpe_rng_state"},
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": 27546880, "traffic_source": "ncu --set full, profiles/r1_ncu_raw_logpdf_mma_final.txt "
                         "(dram__bytes_read.sum + dram__bytes_write.sum of this launch)",
                         "peak_source": peak_src, "kernel": kernel_name,
                         "kernel_ms": k_ms, "algorithmic_bytes": algorithmic_bytes(),
                         "note": "the C x K x P grid reuses every history byte C=4096 times from shared memory, so "
                                 "this kernel sits on the fp64 (DMMA) pipe, not on HBM (SURVEY.md section 8d); see fp64"},
            # algorithmic = the reference's operation count (4P + 25 flops per cell);
            # executed = what the tensor-core kernel issues after expanding the square: one fma per
            # cell on mma.m8n8k4.f64 (DMMA and DFMA share the 37 TFLOP/s fp64 peak on B200)
            "fp64": {"achieved_tflops": fl / (k_ms * 1e-3) / 1e12, "peak_tflops": fp64_peak,
                     "frac": (fl / (k_ms * 1e-3) / 1e12 / fp64_peak) if fp64_peak else None,
                     "algorithmic_flops": fl,
                     "executed_flops": executed_flops(),
                     "executed_tflops": executed_flops() / (k_ms * 1e-3) / 1e12,
                     "executed_frac": (executed_flops() / (k_ms * 1e-3) / 1e12 / fp64_peak) if fp64_peak else None,
                     "kernel": eng.last_logpdf_kernel(),
                     "peak_source": "tpe_probe_fp64_tflops (DFMA microbenchmark, this run; tools/probe_dmma.cu "
                                    "measures the same 37.1 TFLOP/s for DMMA.8x8x4)"},
        }
        if extras:
            line["extras"] = extras
        if config5:
            line["config5"] = config5
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline_leg(X, loss)
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the univariate / batched / cold extras")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
