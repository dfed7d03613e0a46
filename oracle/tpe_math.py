"""Oracle special functions (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

NumPy restatement of the truncated-normal arithmetic of the reference:

* erf            <- optuna/samplers/_tpe/_erf.py:112-142   (msun rational approximations,
                    scalar libm erf for arrays smaller than 2000 elements, :134)
* ndtr           <- optuna/samplers/_tpe/_truncnorm.py:73-75
* ndtr_scalar    <- _truncnorm.py:59-70
* log_ndtr       <- _truncnorm.py:79-106  (three regimes: t>6, -20<t<=6, asymptotic series)
* log_gauss_mass <- _truncnorm.py:113-149 (left / right / central cases)
* ndtri_exp      <- _truncnorm.py:152-221 (Newton on log_ndtr, batch-global stop)
* ppf            <- _truncnorm.py:224-266
* logpdf         <- _truncnorm.py:286-297

The arithmetic order of every expression follows the reference so that results agree
bit-for-bit with it on the same libm/NumPy (pinned by tests/test_oracle_golden.py).
"""
from __future__ import annotations

import math
import sys

import numpy as np

SQRT2 = 2**0.5
SQRT_2PI = math.sqrt(2 * math.pi)
LOG_SQRT_2PI = math.log(SQRT_2PI)  # _truncnorm.py:45-46
LOGISTIC_C = math.sqrt(3) / math.pi  # _truncnorm.py:47
ERF_SMALL_ARRAY = 2000  # _erf.py:134

# msun s_erf.c constants (FreeBSD libm; also listed in _erf.py:30-109). Lowest order first.
ERX = 8.45062911510467529297e-01
EFX = 1.28379167095512586316e-01
PP = (1.28379167095512558561e-01, -3.25042107247001499370e-01, -2.84817495755985104766e-02,
      -5.77027029648944159157e-03, -2.37630166566501626084e-05)
QQ = (1.0, 3.97917223959155352819e-01, 6.50222499887672944485e-02, 5.08130628187576562776e-03,
      1.32494738004321644526e-04, -3.96022827877536812320e-06)
PA = (-2.36211856075265944077e-03, 4.14856118683748331666e-01, -3.72207876035701323847e-01,
      3.18346619901161753674e-01, -1.10894694282396677476e-01, 3.54783043256182359371e-02,
      -2.16637559486879084300e-03)
QA = (1.0, 1.06420880400844228286e-01, 5.40397917702171048937e-01, 7.18286544141962662868e-02,
      1.26171219808761642112e-01, 1.36370839120290507362e-02, 1.19844998467991074170e-02)
RA = (-9.86494403484714822705e-03, -6.93858572707181764372e-01, -1.05586262253232909814e01,
      -6.23753324503260060396e01, -1.62396669462573470355e02, -1.84605092906711035994e02,
      -8.12874355063065934246e01, -9.81432934416914548592e00)
SA = (1.0, 1.96512716674392571292e01, 1.37657754143519042600e02, 4.34565877475229228821e02,
      6.45387271733267880336e02, 4.29008140027567833386e02, 1.08635005541779435134e02,
      6.57024977031928170135e00, -6.04244152148580987438e-02)
RB = (-9.86494292470009928597e-03, -7.99283237680523006574e-01, -1.77579549177547519889e01,
      -1.60636384855821916062e02, -6.37566443368389627722e02, -1.02509513161107724954e03,
      -4.83519191608651397019e02)
SB = (1.0, 3.03380607434824582924e01, 3.25792512996573918826e02, 1.53672958608443695994e03,
      3.19985821950859553908e03, 2.55305040643316442583e03, 4.74528541206955367215e02,
      -2.24409524465858183362e01)


def _horner(coefs: tuple[float, ...], t: np.ndarray) -> np.ndarray:
    # numpy.polynomial evaluates c[-1], then repeatedly c[-i] + acc * t (separate mul / add).
    acc = coefs[-1] + t * 0.0
    for c in reversed(coefs[:-1]):
        acc = c + acc * t
    return acc


def _erf_abs_below_6(v: np.ndarray) -> np.ndarray:
    """erf on a 1-D array of values in [0, 6) -- the five msun intervals (_erf.py:112-130)."""
    res = np.empty_like(v)
    edges = (2.0**-28, 0.84375, 1.25, 1 / 0.35)
    which = np.zeros(v.shape, dtype=np.int64)
    for e in edges:
        which += v >= e
    sel = which == 0
    if sel.any():
        res[sel] = (1 + EFX) * v[sel]
    sel = which == 1
    if sel.any():
        t = v[sel]
        z = t * t
        res[sel] = t * (1 + _horner(PP, z) / _horner(QQ, z))
    sel = which == 2
    if sel.any():
        s = v[sel] - 1
        res[sel] = ERX + _horner(PA, s) / _horner(QA, s)
    for code, (num, den) in ((3, (RA, SA)), (4, (RB, SB))):
        sel = which == code
        if sel.any():
            t = v[sel]
            z = t * t
            s = 1 / z
            res[sel] = 1 - np.exp(-z - 0.5625 + _horner(num, s) / _horner(den, s)) / t
    return res


def erf(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=float)
    if x.size < ERF_SMALL_ARRAY:
        flat = [math.erf(v) for v in x.ravel()]
        return np.asarray(flat, dtype=float).reshape(x.shape)
    mag = np.abs(x).ravel()
    ok = ~np.isnan(mag)
    res = np.where(ok, 1.0, np.nan)
    idx = np.nonzero(ok & (mag < 6))[0]
    res[idx] = _erf_abs_below_6(mag[idx])
    return np.sign(x) * res.reshape(x.shape)


def ndtr(t: np.ndarray) -> np.ndarray:
    return 0.5 + 0.5 * erf(t / SQRT2)


def ndtr_scalar(t: float) -> float:
    u = t / SQRT2
    if u < -1 / SQRT2:
        return 0.5 * math.erfc(-u)
    if u < 1 / SQRT2:
        return 0.5 + 0.5 * math.erf(u)
    return 1.0 - 0.5 * math.erfc(u)


def log_ndtr_scalar(t: float) -> float:
    if t > 6:
        return -ndtr_scalar(-t)
    if t > -20:
        return math.log(ndtr_scalar(t))
    head = -0.5 * t**2 - math.log(-t) - 0.5 * math.log(2 * math.pi)
    prev = 0.0
    total = 1.0
    num = 1.0
    den = 1.0
    inv_t2 = 1 / t**2
    sgn = 1
    i = 0
    while abs(prev - total) > sys.float_info.epsilon:
        i += 1
        prev = total
        sgn = -sgn
        den *= inv_t2
        num *= 2 * i - 1
        total += sgn * num * den
    return head + math.log(total)


def log_ndtr(t: np.ndarray) -> np.ndarray:
    t = np.asarray(t, dtype=float)
    out = np.empty(t.shape, dtype=float)
    flat_in = t.ravel()
    flat_out = out.reshape(-1)
    cache: dict[float, float] = {}
    for i, v in enumerate(flat_in.tolist()):
        r = cache.get(v)
        if r is None:
            r = log_ndtr_scalar(v)
            cache[v] = r
        flat_out[i] = r
    return out


def _log_diff(lp: np.ndarray, lq: np.ndarray) -> np.ndarray:
    return lp + np.log1p(-np.exp(lq - lp))


def log_gauss_mass(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """ln(Phi(b) - Phi(a)); a, b of identical shape."""
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    left = b <= 0
    right = a > 0
    mid = ~(left | right)
    out = np.full(a.shape, np.nan, dtype=float)
    if left.any():
        out[left] = _log_diff(log_ndtr(b[left]), log_ndtr(a[left]))
    if right.any():
        out[right] = _log_diff(log_ndtr(-a[right]), log_ndtr(-b[right]))
    if mid.any():
        with np.errstate(divide="ignore", invalid="ignore"):
            out[mid] = np.log1p(-ndtr(a[mid]) - ndtr(-b[mid]))
    return out


def ndtri_exp(y: np.ndarray) -> np.ndarray:
    """x with log_ndtr(x) == y, by Newton iterations stopped batch-globally (rtol 1e-8)."""
    y = np.asarray(y, dtype=float)
    flip = y > -1e-2
    z = y.copy()
    with np.errstate(divide="ignore"):
        z[flip] = np.log(-np.expm1(y[flip]))
    x = np.empty_like(y)
    tail = z < -5
    if tail.any():
        x[tail] = -np.sqrt(-2.0 * (z[tail] + LOG_SQRT_2PI))
    body = ~tail
    if body.any():
        x[body] = -LOGISTIC_C * np.log(np.expm1(-z[body]))
    for _ in range(100):
        lphi = log_ndtr(x)
        lpdf = -0.5 * x**2 - LOG_SQRT_2PI
        dx = (lphi - z) * np.exp(lphi - lpdf)
        x -= dx
        if np.all(np.abs(dx) < 1e-8 * np.abs(x)):
            break
    x[flip] *= -1
    return x


def ppf(q: np.ndarray, a: np.ndarray, b: np.ndarray) -> np.ndarray:
    q, a, b = np.broadcast_arrays(*np.atleast_1d(q, a, b))
    neg = a < 0
    pos = ~neg
    lm = log_gauss_mass(a, b)
    out = np.empty(q.shape, dtype=float)
    with np.errstate(divide="ignore"):
        if neg.any():
            out[neg] = ndtri_exp(np.logaddexp(log_ndtr(a[neg]), np.log(q[neg]) + lm[neg]))
        if pos.any():
            out[pos] = -ndtri_exp(np.logaddexp(log_ndtr(-b[pos]), np.log1p(-q[pos]) + lm[pos]))
    out[q == 0] = a[q == 0]
    out[q == 1] = b[q == 1]
    out[a == b] = math.nan
    return out


def logpdf(x: np.ndarray, a: np.ndarray, b: np.ndarray, loc: np.ndarray,
           scale: np.ndarray) -> np.ndarray:
    z = (x - loc) / scale
    z, a, b = np.atleast_1d(z, a, b)
    val = (-(z**2) / 2.0 - LOG_SQRT_2PI) - log_gauss_mass(a, b) - np.log(scale)
    z, a, b = np.broadcast_arrays(z, a, b)
    return np.select([a == b, (z < a) | (z > b)], [np.nan, -np.inf], default=val)
