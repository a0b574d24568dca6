#!/usr/bin/env python3
"""Times the device-resident NTT at the given sizes and prints one JSON line per size with per-pass milliseconds.
Usage: python tools/ntt_tune.py K [K ...]   (environment: H2B_NTT_BMAX, H2B_NTT_SINGLE)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import halo2_scaffold_b200 as h2
from halo2_scaffold_b200.domain import FR_MODULUS, FR_ROOT_OF_UNITY, FR_S, fr_to_words

L = h2.load()
L.init_device(0)
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
for k in [int(a) for a in sys.argv[1:]]:
    n = 1 << k
    w = FR_ROOT_OF_UNITY
    for _ in range(k, FR_S):
        w = w * w % FR_MODULUS
    ww = fr_to_words(w)
    d = torch.empty(n * 4, dtype=torch.int64, device=dev)
    L.gen_scalars_dev(0, 0xA000 + k, n, 0, d.data_ptr(), st)
    for _ in range(3):
        L.ntt_dev(0, d.data_ptr(), ww, k, st)
    torch.cuda.synchronize()
    steps = 10 if k <= 24 else 4
    L.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        L.ntt_dev(0, d.data_ptr(), ww, k, st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    ph = {}
    for tag, t in L.profile_read():
        ph[tag] = ph.get(tag, 0.0) + t / steps
    L.profile_enable(False)
    print(json.dumps({"k": k, "ms": round(ms, 4), "gelem_s": round(n / ms / 1e6, 3), "passes_ms": {str(a): round(b, 4) for a, b in sorted(ph.items())},
                      "env": {e: os.environ.get(e) for e in ("H2B_NTT_BMAX", "H2B_NTT_SINGLE")}}), flush=True)
    del d
