#!/usr/bin/env python3
"""e2e of the host-pointer drop-ins with PAGEABLE host buffers (what a Rust Vec is) vs pinned ones, 2^k (default 24)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import halo2_scaffold_b200 as h2
from bench import omega_words

k = int(sys.argv[1]) if len(sys.argv) > 1 else 24
n = 1 << k
L = h2.load(); L.init_device(0)
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
d_scal = torch.empty(n * 4, dtype=torch.int64, device=dev)
d_base = torch.empty(n * 8, dtype=torch.int64, device=dev)
L.gen_scalars_dev(0, 1, n, 0, d_scal.data_ptr(), st)
L.gen_points_dev(0, 2, n, d_base.data_ptr(), st)
torch.cuda.synchronize()
hb = d_base.cpu(); del d_base
handle = L.register_bases(hb.numpy().view(np.uint64)); del hb
pinned = torch.empty(n * 4, dtype=torch.int64).pin_memory(); pinned.copy_(d_scal)
pageable = np.array(pinned.numpy().view(np.uint64).reshape(n, 4), copy=True)
w = omega_words(k)
out = {"k": k}
for name, arr in (("pinned", pinned.numpy().view(np.uint64).reshape(n, 4)), ("pageable", pageable)):
    for _ in range(2):
        L.msm_registered(arr, handle)
    t0 = time.perf_counter()
    for _ in range(3):
        L.msm_registered(arr, handle)
    out["msm_e2e_ms_" + name] = round((time.perf_counter() - t0) / 3 * 1e3, 2)
    a = np.array(arr, copy=True) if name == "pageable" else arr
    L.ntt(a, w, k)
    t0 = time.perf_counter()
    for _ in range(3):
        L.ntt(a, w, k)
    out["ntt_e2e_ms_" + name] = round((time.perf_counter() - t0) / 3 * 1e3, 2)
print(json.dumps(out))
