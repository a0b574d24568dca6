#!/usr/bin/env python3
"""k = 16..26 sweep of the hot path on one B200 (BASELINE.json configs[3]): MSM over a registered SRS vector and NTT,
device-resident (CUDA events) and end to end through the host-pointer drop-ins (pinned host buffers), uniform scalars.
One JSON line per k on stdout.   usage: python tools/sweep.py [kmin kmax]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import halo2_scaffold_b200 as h2
from bench import IMADS_PER_POINT, omega_words

kmin = int(sys.argv[1]) if len(sys.argv) > 1 else 16
kmax = int(sys.argv[2]) if len(sys.argv) > 2 else 26
L = h2.load()
L.init_device(0)
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
imad_ms, imad_ops = L.imad_bench(0, 4096)
imad_peak = imad_ops / imad_ms * 1e3
hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed(fn, steps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def wall(fn, steps, warm=1):
    for _ in range(warm):
        fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    return (time.perf_counter() - t0) * 1e3 / steps


for k in range(kmin, kmax + 1):
    n = 1 << k
    steps = 20 if k <= 20 else (5 if k <= 24 else 2)
    d_scal = torch.empty(n * 4, dtype=torch.int64, device=dev)
    d_base = torch.empty(n * 8, dtype=torch.int64, device=dev)
    d_block = torch.empty(28, dtype=torch.int64, device=dev)
    L.gen_scalars_dev(0, 0xB2000000 + k, n, 0, d_scal.data_ptr(), st)
    L.gen_points_dev(0, 0xB2001000 + k, n, d_base.data_ptr(), st)
    torch.cuda.synchronize()
    hb = d_base.cpu()
    del d_base
    t0 = time.perf_counter()
    handle = L.register_bases(hb.numpy().view(np.uint64))
    reg_ms = (time.perf_counter() - t0) * 1e3
    del hb
    info = L.base_set_info(handle)
    msm_ms = timed(lambda: L.msm_dev_registered(0, d_scal.data_ptr(), handle, 0, n, d_block.data_ptr(), st), steps)
    h_scal = torch.empty(n * 4, dtype=torch.int64).pin_memory()
    h_scal.copy_(d_scal)
    s_np = h_scal.numpy().view(np.uint64).reshape(n, 4)
    msm_e2e_ms = wall(lambda: L.msm_registered(s_np, handle), steps)
    L.unregister_bases(handle)
    w = omega_words(k)
    ntt_ms = timed(lambda: L.ntt_dev(0, d_scal.data_ptr(), w, k, st), steps)
    ntt_e2e_ms = wall(lambda: L.ntt(s_np, w, k), steps)
    print(json.dumps({
        "k": k, "msm_ms": round(msm_ms, 4), "msm_points_per_s": n / msm_ms * 1e3, "msm_e2e_ms": round(msm_e2e_ms, 4), "msm_e2e_points_per_s": n / msm_e2e_ms * 1e3,
        "msm_imad_frac": IMADS_PER_POINT * n / (msm_ms / 1e3) / imad_peak, "srs_tables": info["n_tables"], "srs_spacing": info["spacing"], "srs_registration_ms": round(reg_ms, 1),
        "ntt_ms": round(ntt_ms, 4), "ntt_elements_per_s": n / ntt_ms * 1e3, "ntt_e2e_ms": round(ntt_e2e_ms, 4), "ntt_e2e_elements_per_s": n / ntt_e2e_ms * 1e3,
        "ntt_hbm_frac": 64.0 * n / (ntt_ms / 1e3) / 1e9 / hbm_peak, "ntt_imad_frac": 136 * (n / 2) * k / (ntt_ms / 1e3) / imad_peak}), flush=True)
    del d_scal, h_scal, s_np
    torch.cuda.empty_cache()
