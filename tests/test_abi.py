"""The C-ABI library loads and exports every symbol include/mapperatorinator_b200.h declares (no compute, no GPU)."""
import os
import re

from mapperatorinator_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "mapperatorinator_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mb200_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert _header_symbols() == sorted(_lib.ABI_SYMBOLS)


def test_library_exports_every_symbol():
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    lib = _lib.load()
    assert lib.mb200_abi_version() == 2
    for sym in _header_symbols():
        assert hasattr(lib, sym), sym


def test_no_product_import_of_oracle():
    """The product package must never import the oracle (parity claims depend on it)."""
    pkg = os.path.join(ROOT, "mapperatorinator_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), os.path.join(dirpath, f)
