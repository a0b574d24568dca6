#!/usr/bin/env python3
"""Randomised soak of the two drop-ins against the CPU oracle on one B200: MSM shapes across the thresholds of the implementation
(table spacings 8 / 16 / 20, chunked uploads from 2^20 points, block-scan reduction levels, heavy buckets, implicit cache vs registered
sets, sub-ranges) and NTT sizes 1..2^22 in both directions.  usage: python tools/soak.py SECONDS [seed]   -> one JSON line"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np

import oracle_c as oc
import parity_cases as pc
from halo2_scaffold_b200._lib import Lib

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
oc.build()
L = Lib()
L.init(1)
t_end = time.time() + budget
counts = {"msm": 0, "msm_registered": 0, "ntt": 0, "points": 0}
sizes = [1, 2, 31, 32, 33, 100, 1000, 4095, 4096, 5000, 1 << 14, (1 << 15) + 7, 1 << 16, 1 << 17, (1 << 18) - 3, 1 << 18, 1 << 19, 1 << 20, (1 << 20) + 1024, 1 << 21]
P_big = L.gen_points(77, 1 << 21)
handles = {}
while time.time() < t_end:
    n = int(rng.choice(sizes)) if rng.random() < 0.8 else int(rng.integers(1, 1 << 19))
    kind = int(rng.integers(0, 2))
    s = L.gen_scalars(int(rng.integers(0, 1 << 30)), n, kind)
    if n > 10 and rng.random() < 0.5:           # edge scalars: 0, 1, r - 1, repeated values
        idx = rng.integers(0, n, size=4)
        s[idx[0]] = 0
        s[idx[1]] = s[idx[2]]
    if rng.random() < 0.5 or n > (1 << 20):
        # registered set (tables): random sub-range of a set of n_set points
        n_set = max(n, int(rng.choice([n, min(1 << 21, 2 * n), 1 << 21])))
        if n_set not in handles:
            if len(handles) >= 3:
                k0 = next(iter(handles))
                L.unregister_bases(handles.pop(k0))
            handles[n_set] = L.register_bases(P_big[:n_set])
        off = int(rng.integers(0, n_set - n + 1))
        got = L.msm_registered(s, handles[n_set], off)
        want = oc.best_multiexp(s, P_big[off:off + n])
        counts["msm_registered"] += 1
    else:
        P = P_big[:n].copy()
        if n > 8:
            P[3] = 0
            P[5] = P[4]
        got = L.msm(s, P)
        want = oc.best_multiexp(s, P)
        counts["msm"] += 1
    assert (pc.affine_of(oc, got) == pc.affine_of(oc, want)).all(), ("MSM mismatch", n, kind)
    counts["points"] += n
    k = int(rng.integers(1, 23)) if rng.random() < 0.7 else int(rng.integers(1, 15))
    a = oc.random_fr(int(rng.integers(0, 1 << 30)), 1 << k)
    w = pc.omega_words(oc, k, bool(rng.integers(0, 2)))
    assert (L.ntt(a.copy(), w, k) == oc.best_fft(a, w, k)).all(), ("NTT mismatch", k)
    counts["ntt"] += 1
print(json.dumps({"soak_seconds": budget, "seed": seed, "ok": True, **counts, "implicit_cache": L.implicit_cache_stats()}), flush=True)
