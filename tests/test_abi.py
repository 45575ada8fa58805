"""CPU-side checks of the drop-in boundary: the shared library loads and exports every symbol that
include/carc_b200.h declares (no compute calls -- there is no GPU here)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "carc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(carc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from carcassonne_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(_lib.lib, name), "libcarc_b200.so does not export " + name
        assert name in _lib.SIGNATURES, "ctypes binding lacks a signature for " + name
    assert _lib.lib.carc_version() >= 100


def test_no_oracle_import_in_product():
    """The product must never route through the CPU oracle."""
    pkg = os.path.join(ROOT, "carcassonne_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
