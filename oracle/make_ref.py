"""Recipe for ``oracle/_ref/``: a git-ignored copy of the few reference source files the hot path lives in.

    python oracle/make_ref.py

Run in the build container (where ``/root/reference`` exists); ``__graft_entry__.build()`` calls it.  The copy is
listed in ``.gitignore`` (never committed) but not in ``.gpurunignore``, so it travels to the GPU box, where
``bench.py --impl reference`` and the ``cpu_baseline`` leg execute it through ``oracle/ref_loader.py``.
TEST INFRASTRUCTURE ONLY.
"""

from __future__ import annotations

import shutil
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))

from oracle.ref_loader import FILES, UPSTREAM, VENDORED  # noqa: E402


def make() -> bool:
    if not all((UPSTREAM / f).exists() for f in FILES):
        return VENDORED.exists()
    for f in FILES:
        dst = VENDORED / f
        dst.parent.mkdir(parents=True, exist_ok=True)
        if not dst.exists() or dst.read_bytes() != (UPSTREAM / f).read_bytes():
            shutil.copyfile(UPSTREAM / f, dst)
    for extra in ("LICENSE.txt",):
        src = UPSTREAM.parent / extra
        if src.exists():
            shutil.copyfile(src, VENDORED.parent / extra)
    return True


if __name__ == "__main__":
    print("oracle/_ref ready" if make() else "reference sources not available here")
