"""CPU oracle for the TPE suggestion path -- TEST INFRASTRUCTURE ONLY.

This package is a NumPy restatement of the reference algorithm
(optuna/samplers/_tpe/*.py, optuna/_hypervolume/*.py,
optuna/study/_multi_objective.py @ 4df4b72).  It exists to *check* the CUDA
path.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU
baseline / ``--impl reference`` legs may import it.  Nothing under
``optuna_b200/`` imports it, and the product path raises if the CUDA library
is missing -- there is no CPU fallback.

Parity pinning: the oracle is checked bit-for-bit (or to 1e-15) against golden
vectors produced by importing the live reference in the build container
(``oracle/gen_golden.py`` -> ``tests/golden/*.npz``); see tests/test_oracle_golden.py.
"""
