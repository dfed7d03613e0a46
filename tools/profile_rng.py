"""Device MT19937 (k_mt19937_uniform): host-side cost of stage / finish and the kernel's duration."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optuna_b200 import TPEEngine  # noqa: E402

e = TPEEngine(0)
r = np.random.RandomState(0)
for n in (135168, 135168, 135168, 8192 * 24 * 33):
    t0 = time.perf_counter()
    e.stage_rng(r, n)
    t1 = time.perf_counter()
    e.get_uniforms(8)
    t2 = time.perf_counter()
    e.finish_rng(r)
    t3 = time.perf_counter()
    print(f"n={n}: stage {1e3 * (t1 - t0):.3f} ms, generate+read {1e3 * (t2 - t1):.3f} ms, finish {1e3 * (t3 - t2):.3f} ms")
