#!/usr/bin/env python3
"""Batched multi-column MSM / NTT on ONE device against the same columns issued one by one (the shipped
examples run k = 16 .. 20: /root/reference/examples/linear_regression.rs:130-131, logistic_regression.rs:152-153).
One JSON line per (k, columns).  usage: python tools/batch_bench.py [k ...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from halo2_scaffold_b200._lib import Lib
from halo2_scaffold_b200 import verify as V
from bench import omega_words

L = Lib()
L.init_device(0)
st = torch.cuda.current_stream().cuda_stream
ks = [int(a) for a in sys.argv[1:]] or [16, 18, 20]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for k in ks:
    n = 1 << k
    seed_p = 0xB2001000 + k
    P = L.gen_points(seed_p, n)
    h = L.register_bases(P)
    info = L.base_set_info(h)
    for C in (1, 4, 8, 16, 32):
        if C * n > (1 << 23):
            continue
        cols = [torch.empty(n * 4, dtype=torch.int64, device="cuda") for _ in range(C)]
        for j, c in enumerate(cols):
            L.gen_scalars_dev(0, 0xB2000000 + k + 17 * j, n, j % 2, c.data_ptr(), st)
        out = torch.empty(28 * C, dtype=torch.int64, device="cuda")
        ptrs = [c.data_ptr() for c in cols]

        def serial():
            for j in range(C):
                L.msm_dev_registered(0, ptrs[j], h, 0, n, out.data_ptr() + 224 * j, st)

        def batched():
            L.msm_dev_batch_registered(0, ptrs, [n] * C, h, out.data_ptr(), st)
        reps = 20 if k <= 18 else 5
        ms_serial = timed(serial, reps)
        want = out.cpu().numpy().view(np.uint64).reshape(C, 28).copy()
        out.zero_()
        ms_batch = timed(batched, reps)
        got = out.cpu().numpy().view(np.uint64).reshape(C, 28)
        same = all(V.jacobian_words_to_affine(got[j]) == V.jacobian_words_to_affine(want[j]) for j in range(C))
        # checksum of column 0 (uniform) against [sum s z] G
        d_c = torch.empty(4, dtype=torch.int64, device="cuda")
        L.msm_checksum_dev(0, ptrs[0], seed_p, n, d_c.data_ptr(), stream=st)
        ok0 = V.jacobian_words_to_affine(got[0]) == V.scalar_mul_generator(V.words_to_int(d_c.cpu().numpy().view(np.uint64)))
        # NTT: C polynomials
        w = omega_words(k)

        def ntt_serial():
            for j in range(C):
                L.ntt_dev(0, ptrs[j], w, k, st)

        def ntt_batched():
            L.ntt_dev_batch(0, ptrs, w, k, st)
        ntt_s = timed(ntt_serial, reps)
        ntt_b = timed(ntt_batched, reps)
        print(json.dumps({"k": k, "columns": C, "msm_serial_ms": round(ms_serial, 4), "msm_batched_ms": round(ms_batch, 4),
                          "msm_speedup": round(ms_serial / ms_batch, 2), "msm_points_per_s_batched": C * n / ms_batch * 1e3,
                          "batched_equals_serial": bool(same), "column0_equals_checksum": bool(ok0),
                          "ntt_serial_ms": round(ntt_s, 4), "ntt_batched_ms": round(ntt_b, 4), "ntt_speedup": round(ntt_s / ntt_b, 2),
                          "ntt_elements_per_s_batched": C * n / ntt_b * 1e3, "tables": info["n_tables"], "spacing": info["spacing"]}), flush=True)
        del cols, out
    L.unregister_bases(h)
