import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oc():
    """The C++ CPU oracle (test infrastructure)."""
    import oracle_c
    oracle_c.build()
    return oracle_c


@pytest.fixture(scope="session")
def emu():
    """The kernel sources compiled for the CPU kernel-logic emulator (tools/emu). Test infrastructure only."""
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "halo2_scaffold_b200", "csrc"), "-s", "emu"])
    from halo2_scaffold_b200._lib import Lib
    L = Lib(os.path.join(ROOT, "tools", "emu", "libh2b200_emu.so"), allow_emulator=True)
    L.init(1)
    return L


@pytest.fixture(scope="session")
def gpu():
    """The product library on EVERY visible B200 (host-pointer MSMs from 2^18 points on are split by point range across them,
    columns go round-robin; the *_dev tests address device 0).  Fails loudly (no skip, no fallback) if it cannot run."""
    from halo2_scaffold_b200 import load
    L = load()
    assert not L.is_emulator
    L.init(0)
    return L


@pytest.fixture(scope="session")
def golden():
    d = os.path.join(ROOT, "tests", "golden")
    return {name: np.load(os.path.join(d, name + "_golden.npz")) for name in ("ntt", "msm", "domain", "prover")}
