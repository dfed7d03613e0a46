"""Host instantiation of csrc/tpe_math.cuh vs the oracle (logic check without a GPU).

The shim (tests/csrc/math_shim.cu) is test-only; the product library never runs these on the
host.  Device-side parity of the same functions is covered by the `gpu` tests.
"""
import ctypes
import math
import os
import subprocess

import numpy as np
import pytest

from tests._util import load

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "math_shim.cu")
LIB = os.path.join(HERE, "csrc", "libmath_shim.so")


@pytest.fixture(scope="module")
def shim():
    hdr = os.path.join(HERE, "..", "optuna_b200", "csrc", "tpe_math.cuh")
    if (not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(hdr))):
        subprocess.check_call(["nvcc", "-O2", "-shared", "-Xcompiler", "-fPIC,-ffp-contract=off",
                               "-Wno-deprecated-gpu-targets", "-o", LIB, SRC])
    return ctypes.CDLL(LIB)


def _map1(lib, name, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    getattr(lib, name)(x.ctypes.data_as(ctypes.c_void_p), y.ctypes.data_as(ctypes.c_void_p),
                       ctypes.c_long(x.size))
    return y


def _mapn(lib, name, *arrs):
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in arrs]
    y = np.empty_like(arrs[0])
    getattr(lib, name)(*[a.ctypes.data_as(ctypes.c_void_p) for a in arrs],
                       y.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(y.size))
    return y


def test_erf_np_bit_exact_vs_reference_polynomial_path(shim):
    g = load("math.npz")
    got = _map1(shim, "shim_erf_np", g["erf_x"])
    ref = g["erf_big"]
    # identical rounding sequence; only exp() (glibc on both sides here) could differ
    bad = ~((got == ref) | (np.isnan(got) & np.isnan(ref)))
    assert np.abs(got - ref)[~np.isnan(ref)].max() <= 1.2e-16, np.abs(got - ref)[~np.isnan(ref)].max()
    assert bad.mean() < 0.02


def test_erf_c_erfc_c_match_libm(shim):
    x = np.concatenate([np.linspace(-7, 7, 5001), np.random.RandomState(0).normal(0, 1.5, 3000),
                        [0.0, 1e-320, 1e-10, 0.25, 0.84375, 1.25, 1 / 0.35, 6.0, 27.9, 28.0, 30.0, -30.0]])
    ref_erf = np.asarray([math.erf(v) for v in x])
    ref_erfc = np.asarray([math.erfc(v) for v in x])
    got_erf = _map1(shim, "shim_erf_c", x)
    got_erfc = _map1(shim, "shim_erfc_c", x)
    assert np.max(np.abs(got_erf - ref_erf)) <= 2.3e-16
    rel = np.abs(got_erfc - ref_erfc) / np.maximum(np.abs(ref_erfc), 1e-300)
    assert np.max(rel) <= 5e-16


def test_log_ndtr_and_mass(shim):
    g = load("math.npz")
    got = _map1(shim, "shim_log_ndtr", g["lndtr_t"])
    ref = g["lndtr"]
    assert np.max(np.abs(got - ref) / np.maximum(1.0, np.abs(ref))) < 1e-15
    got = _mapn(shim, "shim_log_gauss_mass", g["lgm_a"], g["lgm_b"])
    ref = g["lgm"]
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(got), fin)
    # _log_diff cancellation amplifies 1-ulp libm differences by 1/(1 - Phi(a)/Phi(b))
    err = np.abs(got - ref)[fin]
    assert np.percentile(err, 99) < 1e-13
    assert np.max(err) < 1e-9


def test_ppf_and_ndtri(shim):
    g = load("math.npz")
    got = _map1(shim, "shim_ndtri_exp", g["ndtri_y"])
    ref = g["ndtri"]
    assert np.max(np.abs(got - ref) / np.maximum(1.0, np.abs(ref))) < 1e-12
    got = _mapn(shim, "shim_trunc_ppf", g["ppf_q"], g["ppf_a"], g["ppf_b"])
    ref = g["ppf"]
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref)
    assert np.max(np.abs(got - ref)[ok] / np.maximum(1.0, np.abs(ref[ok]))) < 1e-10


def test_pairwise_sum_matches_numpy(shim):
    rs = np.random.RandomState(3)
    shim.shim_pairwise.restype = ctypes.c_double
    for n in (0, 1, 5, 7, 8, 9, 31, 64, 127, 128, 129, 500, 4097):
        x = rs.normal(size=n) * 10 ** rs.uniform(-3, 3, size=n)
        got = shim.shim_pairwise(x.ctypes.data_as(ctypes.c_void_p), ctypes.c_long(n))
        assert got == float(np.sum(x)), n


def test_hypervolume_host_instantiation_vs_reference_goldens(shim):
    """csrc/tpe_motpe.cuh (WFG / 2-D / 3-D hypervolume) against live-reference values."""
    import ctypes as C
    from oracle import motpe as mo
    shim.shim_hypervolume.restype = C.c_double
    g = load("motpe.npz")

    def hv(v, ref, assume_pareto=False):
        v = np.ascontiguousarray(v, dtype=np.float64)
        ref = np.ascontiguousarray(ref, dtype=np.float64)
        return shim.shim_hypervolume(v.ctypes.data_as(C.c_void_p), C.c_int(v.shape[0]), C.c_int(v.shape[1]),
                                     ref.ctypes.data_as(C.c_void_p), C.c_int(int(assume_pareto)))

    for ci in range(int(g["hv_n"])):
        v, ref, want = g[f"hv{ci}/v"], g[f"hv{ci}/ref"], float(g[f"hv{ci}/hv"])
        got = hv(v, ref)
        assert abs(got - want) <= 4e-16 * abs(want), (ci, got, want)
    # random Pareto subsets in 2..6 dims vs the oracle (itself pinned to the reference)
    rs = np.random.RandomState(5)
    for m in (2, 3, 4, 5, 6):
        for n in (1, 2, 3, 4, 9, 25):
            v = rs.uniform(0, 1, (n, m))
            ref = mo.reference_point(v)
            want = mo.hypervolume(v, ref)
            assert abs(hv(v, ref) - want) <= 1e-15 * abs(want), (m, n)
            ps = v[mo.is_pareto_front(v, False)]
            want = mo.hypervolume(ps, ref, assume_pareto=True)
            assert abs(hv(ps, ref, True) - want) <= 1e-15 * abs(want), (m, n)
    inf_ref = np.array([1.0, np.inf, 1.0, 1.0])
    assert hv(rs.uniform(0, 1, (5, 4)), inf_ref) == np.inf
