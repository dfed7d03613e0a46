"""The truncation bound of the fast Gauss transform used for the floor-bandwidth kernels of 1-D estimators
(optuna_b200/csrc/tpe_uni.cuh: k_fgt_coeff / k_fgt_eval), restated in NumPy and checked against the direct sum.

For the sources of a box (|t_j| <= 1/8 in units of sigma sqrt 2 around the box centre) and a target y,
    sum_j w_j exp(-(y - t_j)^2) = exp(-y^2) sum_n A_n H_n(y),   A_n = sum_j w_j t_j^n / n!,
truncated after kFgtTerms = 24 terms; DESIGN.md section 3 claims a remainder <= 4e-14 of the box's own sum up to
|y| = kFgtYmax = 9 (boxes further out are summed directly by the kernel).  No GPU involved."""
import math

import numpy as np
import pytest

TERMS, YMAX = 24, 9.0   # kFgtTerms, kFgtYmax


def box_sum_expansion(w, t, y):
    """What k_fgt_coeff + k_fgt_eval compute for one box, in their operation order (fp64)."""
    a = np.zeros(TERMS)
    pw = w.astype(float).copy()
    for n in range(TERMS):
        a[n] = pw.sum()
        pw = pw * (t / (n + 1))
    y2 = 2.0 * y
    hm, h = 1.0, y2
    acc = a[1] * h + a[0]
    for n in range(1, TERMS - 1):
        hm, h = h, y2 * h - 2.0 * n * hm
        acc = a[n + 1] * h + acc
    return acc * math.exp(-y * y)


@pytest.mark.parametrize("seed", range(6))
def test_truncation_after_24_terms_stays_below_the_documented_bound(seed):
    rs = np.random.RandomState(seed)
    n = int(rs.choice([1, 3, 40, 400]))
    t = rs.uniform(-0.125, 0.125, n)
    if seed % 3 == 0:
        t = np.abs(t) * (1 if seed % 2 else -1)   # one-sided boxes: the alternating series is the hard case
    w = np.exp(rs.uniform(-12, 0, n))              # mixture weights spread like the ramp 1/n .. 1
    worst = 0.0
    for y in np.concatenate([np.linspace(-YMAX, YMAX, 181), rs.uniform(-YMAX, YMAX, 50)]):
        direct = float(np.sum(w * np.exp(-(y - t) ** 2)))
        got = box_sum_expansion(w, t, float(y))
        rel = abs(got - direct) / direct
        worst = max(worst, rel)
        assert rel <= (2e-14 if abs(y) <= 5 else 4e-13), (y, rel)
    assert worst < 4e-13


def test_one_more_box_width_breaks_the_bound():
    """The box width is not arbitrary: sources at |t| = 1/4 (boxes twice as wide) lose the accuracy at |y| = 9."""
    w, t = np.array([1.0]), np.array([-0.25])          # the source on the far side of the centre
    y = 9.0
    direct = float(np.exp(-(y - t[0]) ** 2))
    assert abs(box_sum_expansion(w, t, y) - direct) / direct > 1e-8
