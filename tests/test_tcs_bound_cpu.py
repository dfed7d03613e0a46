"""The rounding bound of the opt-in bf16-screened grid (optuna_b200/csrc/tpe_tcscreen.cuh, TPE_TCS=1): the host
computes delta = P rho^2 (2^-8 + 2^-16 + (P + 2) 5.97e-8) 1.02 + 2e-3 and keeps every cell whose bf16 value lies
within skip + 2 delta of the row maximum.  Here: bf16 rounding and an fp32 inner product emulated in NumPy, against
the fp64 value, for coordinates up to rho.  No GPU involved."""
import numpy as np
import pytest


def to_bf16(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even to bfloat16, returned as float32."""
    u = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) >> 16 << 16
    return u.astype(np.uint32).view(np.float32)


@pytest.mark.parametrize("P,rho", [(16, 3.0), (32, 3.45), (32, 6.0), (64, 2.5)])
def test_bf16_inner_product_stays_within_delta(P, rho):
    rs = np.random.RandomState(P)
    delta = P * rho * rho * (1.0 / 256 + 1.0 / 65536 + (P + 2) * 5.97e-8) * 1.02 + 2e-3
    worst = 0.0
    for _ in range(200):
        x = rs.uniform(-rho, rho, (64, P))
        m = rs.uniform(-rho, rho, (64, P))
        if _ % 4 == 0:                       # the extreme corners
            x, m = np.sign(x) * rho * (1 - 1e-3 * rs.uniform(size=x.shape)), np.sign(m) * rho
        exact = x @ m.T
        xb, mb = to_bf16(x), to_bf16(m)
        acc = np.zeros((64, 64), dtype=np.float32)
        for j in range(P):                   # fp32 accumulation, one product at a time (exact products)
            acc = (acc + (xb[:, j:j + 1] * mb[:, j][None, :]).astype(np.float32)).astype(np.float32)
        worst = max(worst, float(np.abs(acc.astype(np.float64) - exact).max()))
    assert worst <= delta, (worst, delta)
    assert worst >= 0.02 * delta            # (the bound is within two orders of what actually happens)
