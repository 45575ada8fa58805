"""Installs the UNMODIFIED reference package next to the repo for the benchmark's reference arm.

The reference (gcross/Carcassonne) is pure Python with no setup.py, so `pip install --target baseline/_ref` has nothing
to build; what an install of it amounts to is its package directory on the import path.  This script copies
`/root/reference/carcassonne` (library modules only: no tests, no byte code) to `baseline/_ref/carcassonne`.
`baseline/_ref/` is git-ignored (no reference source enters the history) but not gpurun-ignored, so the copy travels
to the GPU box, where `/root/reference` does not exist.  `bench.py --impl reference` imports it from there and times
the reference's own `formExpectationStage3` multiplier (tensors/_2d/sparse.py:100-161) on the box's host cores; when
the copy is absent it falls back to the oracle port and says so (`cpu_baseline.kind`).

Run by `__graft_entry__.build()` whenever `/root/reference` is present.  Test infrastructure: nothing in
`carcassonne_b200/` imports it.
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOURCE = "/root/reference/carcassonne"
TARGET = os.path.join(ROOT, "baseline", "_ref", "carcassonne")


def install(source=SOURCE, target=TARGET):
    """-> True when the reference package is in place under baseline/_ref (copied now or earlier)."""
    if not os.path.isdir(source):
        return os.path.isdir(target)
    if os.path.isdir(target):
        shutil.rmtree(target)
    shutil.copytree(source, target, ignore=shutil.ignore_patterns("tests", "__pycache__", "*.pyc"))
    return True


if __name__ == "__main__":
    ok = install()
    print("reference package %s at %s" % ("installed" if ok else "NOT available", TARGET))
    sys.exit(0 if ok else 1)
