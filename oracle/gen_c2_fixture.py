"""Full-size parity fixture for BASELINE config 2 (TEST INFRASTRUCTURE): log l(x) and log g(x) of 256 points under the
two estimators of the synthetic 100 000 x 32 history, by the oracle's chunked log_pdf (bit-identical to the un-chunked
call).  The points are the first 256 of the 4096 candidates the oracle draws with seed 1 --
what the reference itself would evaluate.  ~6 min of CPU; writes tests/golden/c2_logpdf.npz (16 KB).

    python oracle/gen_c2_fixture.py
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import tpe_oracle as orc  # noqa: E402

N, P, C, M = 100_000, 32, 4096, 256
rs = np.random.RandomState(0)
X = rs.uniform(0, 1, (N, P))
loss = ((X - 0.5) ** 2).sum(1)
cat = np.zeros(N, np.int8)
key = np.stack([loss, np.zeros(N)], 1)
params = [orc.Param("float", 0.0, 1.0) for _ in range(P)]
cfg = orc.Config(multivariate=True)
below, above = orc.split_trials(cat, key, orc.default_gamma(N))
mb = orc.build_mixture(X[below], params, cfg)
ma = orc.build_mixture(X[above], params, cfg)
cand = orc.mixture_sample(mb, np.random.RandomState(1), C)[:M]
t0 = time.time()
ll = orc.mixture_log_pdf_chunked(mb, cand, 64)
lg = orc.mixture_log_pdf_chunked(ma, cand, 8)
print("oracle log_pdf of", M, "points:", time.time() - t0, "s")
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "c2_logpdf.npz")
np.savez_compressed(out, x=cand, logl=ll, logg=lg, below=below)
print("wrote", out, os.path.getsize(out), "bytes")
