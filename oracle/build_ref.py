"""Recipe for ``oracle/_ref`` -- the UNMODIFIED reference (optuna @ /root/reference), importable on the GPU box.

TEST / BASELINE INFRASTRUCTURE, not product: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` put
``oracle/_ref`` on ``sys.path`` (``oracle.ref.enable()``).  ``optuna_b200`` itself never does; when a user has
optuna installed it plugs into that installation.

The reference is pure Python (no build system to run): the package directory is copied byte for byte where it
lies under /root/reference into ``oracle/_ref/optuna`` (git-ignored, so no reference source enters the history;
not gpurun-ignored, so it travels with the snapshot), plus a stub for its one missing import, ``colorlog``
(optuna/logging.py:14,38 -- a coloured log formatter).  ``sqlalchemy`` / ``alembic`` are only
imported by the RDB storage, which this path never touches.

    python oracle/build_ref.py            # no-op when /root/reference is absent (GPU box: uses the shipped copy)
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/optuna"
REF_DST = os.path.join(HERE, "_ref")

COLORLOG_STUB = '''"""Stub of the third-party ``colorlog`` package (absent from this image): optuna/logging.py:38 only needs
``TTYColoredFormatter(fmt, stream=...)``."""
import logging


class ColoredFormatter(logging.Formatter):
    def __init__(self, fmt=None, datefmt=None, style="%", *args, **kwargs):
        kwargs.pop("stream", None)
        fmt = (fmt or "%(message)s").replace("%(log_color)s", "").replace("%(reset)s", "")
        super().__init__(fmt, datefmt, style)


TTYColoredFormatter = ColoredFormatter
'''


def build(force: bool = False) -> bool:
    """Copy the reference package; returns True when ``oracle/_ref/optuna`` exists afterwards."""
    dst = os.path.join(REF_DST, "optuna")
    if os.path.isdir(REF_SRC):
        stamp = os.path.join(REF_DST, ".stamp")
        src_mtime = max(os.path.getmtime(os.path.join(d, f)) for d, _, fs in os.walk(REF_SRC) for f in fs
                        if f.endswith(".py"))
        if force or not os.path.exists(stamp) or os.path.getmtime(stamp) < src_mtime or not os.path.isdir(dst):
            shutil.rmtree(REF_DST, ignore_errors=True)
            os.makedirs(REF_DST)
            shutil.copytree(REF_SRC, dst, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
            os.makedirs(os.path.join(REF_DST, "colorlog"))
            with open(os.path.join(REF_DST, "colorlog", "__init__.py"), "w") as f:
                f.write(COLORLOG_STUB)
            with open(stamp, "w") as f:
                f.write("copied from /root/reference/optuna by oracle/build_ref.py\n")
    return os.path.isdir(dst)


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref:", "ready" if ok else "absent (no /root/reference and no shipped copy)")
