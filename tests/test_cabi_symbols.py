"""The C-ABI shared library loads and exports every symbol include/nd_b200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "nd_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nd_b200_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(nd):
    nd._cabi.build()
    L = ctypes.CDLL(nd._cabi.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 15
    for s in declared:
        assert hasattr(L, s), f"{s} declared in include/nd_b200.h but not exported"
    assert sorted(nd._cabi.EXPORTED_SYMBOLS) == declared
    assert L.nd_b200_abi_version() == nd._cabi.ABI_VERSION == 5


def test_ctypes_struct_sizes_match_header(nd):
    # field-by-field mirror of the header structs (LP64): catches drift between nd_b200.h and _cabi.py
    assert ctypes.sizeof(nd._cabi.VBatch) == 4 * 4 + 8 + 8 + 4 * 8 + 2 * 4 + 8
    assert ctypes.sizeof(nd._cabi.EBatch) == 6 * 4 + 8 + 8 + 4 * 8 + 2 * 4 + 2 * 4 + 8
    assert ctypes.sizeof(nd._cabi.Desc) == 8 + 16 + 16 + 8 + 8 + 16 + 32 + 16 + 8 + 16 + 16
    assert ctypes.sizeof(nd._cabi.CustomKind) == 6 * 4 + 2 * 8 + 2 * 4


def test_missing_library_fails_loudly(nd, monkeypatch):
    monkeypatch.setattr(nd._cabi, "_lib", None)
    monkeypatch.setattr(nd._cabi, "LIB_PATH", "/nonexistent/libnd_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        nd._cabi.lib()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "networkdynamics.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower() or f.endswith((".cu", ".cuh")) and "import" not in text, (dirpath, f)


def test_product_never_loads_the_emulated_library():
    """tests/cusim (the CPU emulation of the kernels used by the CPU suite) is test infrastructure: no Python module of
    the package mentions it, and the only trace in the CUDA sources is the ND_CUSIM compile-time switch, which the
    product build (nvcc, _cabi.build) never defines."""
    pkg = os.path.join(ROOT, "networkdynamics.jl_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            text = open(os.path.join(pkg, f)).read().lower()
            assert "cusim" not in text and "_sim.so" not in text, f
    build_src = open(os.path.join(pkg, "_cabi.py")).read()
    assert "ND_CUSIM" not in build_src
