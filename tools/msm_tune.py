#!/usr/bin/env python3
"""Times one device-resident MSM configuration (table spacing / window / slice length come from the environment:
H2B_MSM_PRECOMP, H2B_MSM_C, H2B_MSM_SLICE) and prints one JSON line with the per-phase milliseconds.
Usage: python tools/msm_tune.py K [uniform|witness] [steps]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import halo2_scaffold_b200 as h2

k = int(sys.argv[1])
kind = 1 if (len(sys.argv) > 2 and sys.argv[2] == "witness") else 0
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
n = 1 << k
L = h2.load()
L.init_device(0)
dev = torch.device("cuda", 0)
d_scal = torch.empty(n * 4, dtype=torch.int64, device=dev)
d_base = torch.empty(n * 8, dtype=torch.int64, device=dev)
d_block = torch.empty(28, dtype=torch.int64, device=dev)
st = torch.cuda.current_stream().cuda_stream
L.gen_scalars_dev(0, 0xB2000000 + k, n, kind, d_scal.data_ptr(), st)
L.gen_points_dev(0, 0xB2001000 + k, n, d_base.data_ptr(), st)
torch.cuda.synchronize()
plain = os.environ.get("H2B_MSM_PRECOMP", "") == "0"
t0 = time.perf_counter()
handle = None
if not plain:
    hb = d_base.cpu()
    handle = L.register_bases(hb.numpy().view(np.uint64))
    del hb
reg_ms = (time.perf_counter() - t0) * 1e3
info = L.base_set_info(handle) if handle else {}


def step():
    if handle:
        L.msm_dev_registered(0, d_scal.data_ptr(), handle, 0, n, d_block.data_ptr(), st)
    else:
        L.msm_dev_partial(0, d_scal.data_ptr(), d_base.data_ptr(), n, d_block.data_ptr(), st)


for _ in range(2):
    step()
torch.cuda.synchronize()
L.profile_enable(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
ph = {}
for tag, t in L.profile_read():
    ph[tag] = ph.get(tag, 0.0) + t / steps
print(json.dumps({"k": k, "kind": kind, "ms": round(ms, 3), "mpts_s": round(n / ms / 1e3, 1), "phases": {str(a): round(b, 3) for a, b in sorted(ph.items())},
                  "srs": info, "reg_ms": round(reg_ms, 1), "env": {e: os.environ.get(e) for e in ("H2B_MSM_PRECOMP", "H2B_MSM_C", "H2B_MSM_SLICE")},
                  "x0": int(d_block.cpu().numpy().view(np.uint64)[0])}), flush=True)
