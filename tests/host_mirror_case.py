"""Builds tests/cpp/host_mirror_test.cpp against a given libh2b200 build, runs it on seeded inputs and checks every
output file against the oracle (shared by the CPU emulator test and the GPU test)."""
import os
import subprocess

import numpy as np

import bn254 as o
import parity_cases as pc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_host_mirror(oc, lib_path, tmp_path, k=6, j=4):
    n = 1 << k
    exe = str(tmp_path / "host_mirror_test")
    libdir, libfile = os.path.split(lib_path)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-o", exe, os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp"),
                           "-L" + libdir, "-l:" + libfile, "-Wl,-rpath," + libdir])
    s, P, lag = oc.random_fr(0xC0 + k, n), oc.gen_points(0xC1 + k, n), oc.random_fr(0xC2 + k, n)
    s[1] = 0
    P[2] = 0
    for name, arr in (("scalars", s), ("bases", P), ("lagrange", lag)):
        np.ascontiguousarray(arr).tofile(str(tmp_path / (name + ".bin")))
    out = subprocess.run([exe, str(tmp_path), str(j), str(k)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "HOST_MIRROR_OK" in out.stdout, (out.stdout, out.stderr)

    def rd(name, cols):
        return np.fromfile(str(tmp_path / (name + ".bin")), dtype=np.uint64).reshape(-1, cols)

    assert (pc.affine_of(oc, rd("msm", 12)[0]) == pc.affine_of(oc, oc.best_multiexp(s, P))).all()
    dom = o.EvaluationDomain(j, k)
    lag_i = [o.from_mont(v, o.R_MOD) for v in oc.words_to_ints(lag)]
    coeff = dom.lagrange_to_coeff(list(lag_i))
    ext = dom.coeff_to_extended(list(coeff))
    back = dom.extended_to_coeff(list(ext))

    def words(vals):
        return oc.ints_to_words([o.to_mont(v, o.R_MOD) for v in vals])

    assert (rd("coeff", 4) == words(coeff)).all()
    assert (rd("extended", 4) == words(ext)).all()
    assert (rd("back", 4) == words(back)).all()
    assert (rd("fft", 4) == oc.best_fft(lag, pc.omega_words(oc, k), k)).all()
    c = rd("commit", 12)
    assert (pc.affine_of(oc, c[0]) == pc.affine_of(oc, oc.best_multiexp(s, P))).all()
    h = n // 2
    assert (pc.affine_of(oc, c[1]) == pc.affine_of(oc, oc.best_multiexp(s[:h], P[:h]))).all()
