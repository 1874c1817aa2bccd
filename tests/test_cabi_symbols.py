"""The C-ABI library loads and exports every symbol include/druggen_b200.h declares (no GPU work)."""
import os
import re

import pytest

from druggen_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "druggen_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dg_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip("library not built (run __graft_entry__.build())")
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 16
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(_lib.SIGNATURES) | set(_lib.INFO_SYMBOLS)
    assert lib.dg_abi_version() == _lib.ABI_VERSION
