#!/usr/bin/env python3
"""Minimal device-resident workload for ncu captures: `reps` MSMs over a registered SRS vector (window tables) and
`reps` NTTs at 2^k (default 24) through the C ABI."""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np
import torch
import halo2_scaffold_b200 as h2
from bench import omega_words

k = int(sys.argv[1]) if len(sys.argv) > 1 else 24
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n = 1 << k
torch.cuda.set_device(0)
L = h2.load(); L.init_device(0)
st = torch.cuda.current_stream().cuda_stream
d_s = torch.empty(n * 4, dtype=torch.int64, device="cuda")
d_b = torch.empty(n * 8, dtype=torch.int64, device="cuda")
d_o = torch.empty(28, dtype=torch.int64, device="cuda")
L.gen_scalars_dev(0, 1, n, 0, d_s.data_ptr(), st)
L.gen_points_dev(0, 2, n, d_b.data_ptr(), st)
torch.cuda.synchronize()
hb = d_b.cpu()
handle = L.register_bases(hb.numpy().view(np.uint64))
del hb, d_b
for _ in range(reps):
    L.msm_dev_registered(0, d_s.data_ptr(), handle, 0, n, d_o.data_ptr(), st)
w = omega_words(k)
for _ in range(reps):
    L.ntt_dev(0, d_s.data_ptr(), w, k, st)
torch.cuda.synchronize()
print("ncu_target done", k, reps, L.launch_count())
