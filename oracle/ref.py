"""``oracle.ref.enable()`` -- put the unmodified reference (oracle/_ref, see oracle/build_ref.py) on sys.path.

Test / baseline infrastructure only.  If a real optuna is already importable it is left alone."""
from __future__ import annotations

import os
import sys

REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_DIR, "optuna"))


def enable() -> bool:
    """Make ``import optuna`` resolve (to an installed optuna, else to oracle/_ref).  Returns success."""
    try:
        import optuna  # noqa: F401
        return True
    except ImportError:
        pass
    if not available():
        return False
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import optuna  # noqa: F401
    optuna.logging.set_verbosity(optuna.logging.WARNING)
    return True
