"""CPU: the C-ABI shared library loads and exports every symbol include/h2b200.h declares (no compute)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "h2b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(h2b_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    from halo2_scaffold_b200 import exported_symbols
    assert sorted(exported_symbols()) == declared_symbols()


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    lib_path = os.path.join(ROOT, "halo2_scaffold_b200", "lib", "libh2b200.so")
    if not os.path.exists(lib_path):
        ge.build()
    L = ctypes.CDLL(lib_path)
    for sym in declared_symbols():
        assert hasattr(L, sym), "libh2b200.so does not export %s" % sym
    L.h2b_version.restype = ctypes.c_char_p
    assert b"sm_100a" in L.h2b_version()
    assert L.h2b_is_emulator() == 0


def test_no_cpu_fallback_without_init():
    # compute entry points refuse to run before h2b_init; nothing routes to a CPU implementation
    from halo2_scaffold_b200 import load, H2BError
    import numpy as np
    L = load()
    if L.device_count() != 0:
        pytest.skip("library already initialised in this process")
    with pytest.raises(H2BError):
        L.ntt(np.zeros((2, 4), dtype=np.uint64), np.zeros(4, dtype=np.uint64), 1)
    with pytest.raises(H2BError):
        L.msm(np.zeros((1, 4), dtype=np.uint64), np.zeros((1, 8), dtype=np.uint64))


def test_product_loader_refuses_the_emulator(emu):
    from halo2_scaffold_b200._lib import Lib
    with pytest.raises(RuntimeError):
        Lib(emu.path)


def test_product_sources_do_not_reference_the_oracle():
    pkg = os.path.join(ROOT, "halo2_scaffold_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", ".inc")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle_c" not in text and "libh2oracle" not in text and "import bn254" not in text, f
