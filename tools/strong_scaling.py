#!/usr/bin/env python3
"""Strong scaling of ONE MSM across the GPUs of a box, in one process (h2b_init(D): point-range sharding inside the
library, partial sums folded on device 0 -- SURVEY.md 8e row 1).  Host-pointer entry point with pinned scalars, so the
numbers are end to end.  One JSON line per (k, D).   usage: python tools/strong_scaling.py D k [k ...]
(one process per D: the device set of the library is fixed at h2b_init)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from halo2_scaffold_b200._lib import Lib

D = int(sys.argv[1])
L = Lib()
L.init(D)
assert L.device_count() == D
for k in [int(a) for a in sys.argv[2:]]:
    n = 1 << k
    s = torch.empty(n * 4, dtype=torch.int64).pin_memory()
    s_np = s.numpy().view(np.uint64).reshape(n, 4)
    s_np[:] = L.gen_scalars(0xB2000000 + k, n, 0)
    P = L.gen_points(0xB2001000 + k, n)
    t0 = time.perf_counter()
    h = L.register_bases(P)
    reg_ms = (time.perf_counter() - t0) * 1e3
    del P
    for _ in range(2):
        r = L.msm_registered(s_np, h)
    steps = 5 if k <= 24 else 3
    t0 = time.perf_counter()
    for _ in range(steps):
        r = L.msm_registered(s_np, h)
    ms = (time.perf_counter() - t0) / steps * 1e3
    L.unregister_bases(h)
    print(json.dumps({"devices": D, "k": k, "msm_e2e_ms": round(ms, 3), "points_per_s": n / ms * 1e3, "registration_ms": round(reg_ms, 1),
                      "result_x0": int(r[0])}), flush=True)
    del s, s_np
