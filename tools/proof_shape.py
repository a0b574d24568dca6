#!/usr/bin/env python3
"""
Hot-path time of ONE create_proof, as the patched prover would spend it: the sequence of best_multiexp / best_fft calls
of halo2_proofs' prover issued through the host-pointer drop-ins with ordinary (pageable) host
arrays, next to the CPU restatement (oracle, all host threads) timed per call kind and multiplied by the call counts.

The reference's own create_proof cannot be run here (no Rust toolchain), and the column counts are chosen at run time by
halo2-base; the shapes below are the estimates of SURVEY.md (section 3.3 table + appendix) and are printed with the result.
  A advice columns (incl. lookup advice), I instance columns, L lookup arguments, P equality-enabled columns,
  d = cs.degree(), sets = ceil(P / (d - 2)), n = 2^k, extended domain 2^ek.
  MSMs of n points: A + 3L + sets + (d - 1) + 3      iNTTs of n: I + A + 3L + sets      coset NTTs of 2^ek: A + I + 3L + sets + 1
One JSON line per configuration.   usage: python tools/proof_shape.py [cfg ...]
"""
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np

import halo2_scaffold_b200 as h2
from bench import omega_words

SHAPES = {
    "standard_plonk_k5": dict(k=5, A=3, I=1, L=0, P=3, d=3, ref="examples/standard_plonk.rs:26 (k = 5), src/circuits/standard_plonk.rs:29-48"),
    "halo2_lib_k16": dict(k=16, A=1, I=1, L=0, P=2, d=3, ref="examples/halo2_lib.rs via scaffold::prove at DEGREE=16"),
    "linear_regression_k20": dict(k=20, A=3, I=1, L=1, P=4, d=4, ref="examples/linear_regression.rs at DEGREE=20, LOOKUP_BITS=19"),
    "logistic_regression_k22": dict(k=22, A=8, I=1, L=2, P=9, d=4, ref="examples/logistic_regression.rs at DEGREE=22"),
}


def counts(s):
    sets = math.ceil(s["P"] / (s["d"] - 2))
    ek = s["k"]
    while (1 << ek) < (1 << s["k"]) * (s["d"] - 1):
        ek += 1
    return dict(sets=sets, ek=ek, msm=s["A"] + 3 * s["L"] + sets + (s["d"] - 1) + 3, intt=s["I"] + s["A"] + 3 * s["L"] + sets,
                coset=s["A"] + s["I"] + 3 * s["L"] + sets + 1)


def main():
    import oracle_c as oc
    oc.build()
    cores = oc.hardware_threads()
    L = h2.load()
    L.init_device(0)
    names = sys.argv[1:] or list(SHAPES)
    for name in names:
        s = SHAPES[name]
        c = counts(s)
        k, ek, n = s["k"], c["ek"], 1 << s["k"]
        scal = L.gen_scalars(0xC000 + k, n, 1)                # witness-like column
        rand = L.gen_scalars(0xC100 + k, n, 0)                # quotient / product polynomials are uniform
        g = L.gen_points(0xC200 + k, n)
        ext = L.gen_scalars(0xC300 + k, 1 << ek, 0)
        w_n, w_e = omega_words(k), omega_words(ek)
        # the prover holds two SRS vectors; both get their tables on first use (timed separately: once per prover)
        t0 = time.perf_counter()
        for _ in range(2):
            L.msm(scal, g)
        # first NTT of each size: kernel load + twiddle tables for (omega, log n)
        L.ntt(rand, w_n, k)
        L.ntt(ext, w_e, ek)
        setup_ms = (time.perf_counter() - t0) * 1e3
        t0 = time.perf_counter()
        n_wit = s["A"]                                            # advice commits see witness-like scalars
        for i in range(c["msm"]):
            L.msm(scal if i < n_wit else rand, g)
        t1 = time.perf_counter()
        for _ in range(c["intt"]):
            L.ntt(rand, w_n, k)
        t2 = time.perf_counter()
        for _ in range(c["coset"]):
            L.ntt(ext, w_e, ek)
        t3 = time.perf_counter()
        gpu_ms = (t3 - t0) * 1e3
        per_call = {"msm_ms": round((t1 - t0) * 1e3 / c["msm"], 3), "intt_ms": round((t2 - t1) * 1e3 / max(1, c["intt"]), 3),
                    "coset_ntt_ms": round((t3 - t2) * 1e3 / c["coset"], 3)}
        # One advice column, two ways: (a) the three drop-in calls the unchanged prover makes -- commit_lagrange,
        # lagrange_to_coeff, coeff_to_extended, each with its own PCIe round trip (the host-side scaling / padding passes of
        # the reference are NOT counted); (b) the device-resident pipeline: one upload, MSM + both conversions on the device,
        # downloads of the coefficient and extended forms the host evaluator needs.
        from halo2_scaffold_b200.domain import EvaluationDomain, fr_to_words
        dom = EvaluationDomain(s["d"], k, lib=L)
        h = L.register_bases(g)
        en = 1 << ek
        col = scal.copy()
        padded = np.zeros((en, 4), dtype=np.uint64)
        t0 = time.perf_counter()
        for it in range(4):
            if it == 1:
                t0 = time.perf_counter()
            L.msm_registered(col, h)
            L.ntt(col, fr_to_words(dom.omega_inv), k)
            padded[:n] = col
            L.ntt(padded, fr_to_words(dom.extended_omega), ek)
        dropin_ms = (time.perf_counter() - t0) / 3 * 1e3
        d_col, d_ext, d_out = L.dev_alloc(0, n * 32), L.dev_alloc(0, en * 32), L.dev_alloc(0, 224)
        coeff, extended = np.empty((n, 4), dtype=np.uint64), np.empty((en, 4), dtype=np.uint64)
        zs = np.stack([fr_to_words(1), fr_to_words(dom.g_coset), fr_to_words(dom.g_coset_inv)])
        t0 = time.perf_counter()
        for it in range(4):
            if it == 1:
                t0 = time.perf_counter()          # iteration 0 warms scratch buffers and twiddle tables
            L.h2d(0, d_col, scal)
            L.msm_dev_registered(0, d_col, h, 0, n, d_out)
            L.lagrange_to_coeff_dev(0, d_col, k, fr_to_words(dom.omega_inv), fr_to_words(dom.ifft_divisor))
            L.d2h(0, coeff, d_col)
            L.check(L.L.h2b_memcpy_h2d(0, d_ext, coeff.ctypes.data, n * 32))      # stays on the device in a real pipeline; here: same bytes
            L.coeff_to_extended_dev(0, d_ext, k, ek, fr_to_words(dom.extended_omega), zs)
            L.d2h(0, extended, d_ext)
        pipeline_ms = (time.perf_counter() - t0) / 3 * 1e3
        for d in (d_col, d_ext, d_out):
            L.dev_free(0, d)
        L.unregister_bases(h)
        # CPU restatement: one call of each kind, scaled by the counts
        t0 = time.perf_counter(); oc.best_multiexp(scal, g, cores); cpu_msm_wit = time.perf_counter() - t0
        t0 = time.perf_counter(); oc.best_multiexp(rand, g, cores); cpu_msm = time.perf_counter() - t0
        t0 = time.perf_counter(); oc.best_fft(rand, w_n, k, cores); cpu_ntt = time.perf_counter() - t0
        t0 = time.perf_counter(); oc.best_fft(ext, w_e, ek, cores); cpu_ext = time.perf_counter() - t0
        cpu_ms = (n_wit * cpu_msm_wit + (c["msm"] - n_wit) * cpu_msm + c["intt"] * cpu_ntt + c["coset"] * cpu_ext) * 1e3
        print(json.dumps({"config": name, "shape": s, "calls": c, "gpu_hot_path_ms": round(gpu_ms, 2), "gpu_per_call": per_call, "advice_column_ms": {"three_drop_in_calls": round(dropin_ms, 2), "device_resident_pipeline": round(pipeline_ms, 2)}, "gpu_first_use_ms": round(setup_ms, 1),
                          "cpu_hot_path_ms": round(cpu_ms, 1), "cpu_threads": cores, "speedup": round(cpu_ms / gpu_ms, 1),
                          "note": "drop-in host-pointer calls with pageable arrays on 1 x B200; CPU = C++ restatement of halo2_proofs v2023_02_02, per-call times x call counts; column counts are estimates"}), flush=True)


if __name__ == "__main__":
    main()
