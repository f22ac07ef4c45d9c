"""CPU-side checks of the C ABI: the library loads and exports every symbol that
include/dsin_b200.h declares; the ctypes table covers the header; no compute calls."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "dsin_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dsin_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    import __graft_entry__ as g
    g.build()
    import ctypes
    from dsin_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), "libdsin_b200.so does not export %s" % s
    assert lib.dsin_version() >= 100


def test_ctypes_table_matches_header():
    from dsin_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _header_symbols()


def test_no_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from dsin_b200 import _lib, ops
    with pytest.raises(_lib.DsinLibraryError):
        ops.handle()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "dsin_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn
