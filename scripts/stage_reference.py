#!/usr/bin/env python
"""Stage an UNMODIFIED checkout of the reference into the git-ignored ``baseline/_ref/``.

    python scripts/stage_reference.py [--reference /root/reference] [--force]

``/root/reference`` exists only in the build container; ``baseline/_ref/`` is git-ignored (no
reference source enters the history) but NOT gpurun-ignored, so it travels to the GPU box with the
snapshot.  Two users:

* ``tests/test_reference_harness_gpu.py`` drives the reference's own ``BenchmarkRunner`` /
  ``scripts/run_full_benchmark.py`` through ``vectordb_retrieval_b200.plugin.install()`` on a B200;
* ``bench.py``'s CPU legs time the reference's own NumPy ``LinearSearcher`` (kind "reference-numpy").

Only the code the harness executes is copied (``src/``, ``scripts/``, ``configs/``, ``tests/`` and the
top-level files): the archived ``benchmark_results/`` and the docs stay behind.  A ``STAGED_FROM`` note
records where the copy came from; nothing in it is edited."""
from __future__ import annotations

import argparse
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")
PARTS = ("src", "scripts", "configs", "tests")
FILES = ("main.py", "pytest.ini", "requirements.txt", "LICENSE")


def stage(reference: str = "/root/reference", force: bool = False) -> str:
    """Returns the staged path ('' when the reference tree is not available on this machine)."""
    marker = os.path.join(DEST, "STAGED_FROM")
    if os.path.exists(marker) and not force:
        return DEST
    if not os.path.isdir(os.path.join(reference, "src", "algorithms")):
        return ""
    if os.path.isdir(DEST):
        shutil.rmtree(DEST)
    os.makedirs(DEST)
    ignore = shutil.ignore_patterns("__pycache__", "*.pyc", ".pytest_cache")
    for part in PARTS:
        src = os.path.join(reference, part)
        if os.path.isdir(src):
            shutil.copytree(src, os.path.join(DEST, part), ignore=ignore)
    for name in FILES:
        src = os.path.join(reference, name)
        if os.path.isfile(src):
            shutil.copy2(src, os.path.join(DEST, name))
    with open(marker, "w") as f:
        f.write(f"unmodified copy of {os.path.abspath(reference)} ({', '.join(PARTS + FILES)})\n")
    return DEST


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    out = stage(a.reference, a.force)
    print(out or f"reference tree not found at {a.reference}")
    sys.exit(0 if out else 1)
