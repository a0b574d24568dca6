#!/usr/bin/env python3
"""Minimal workload for ncu captures of the widened rows: one launch each of the point decompression, the custom-gate
interpreter, the permutation grand product kernels, then one lookup permutation (bitonic passes, matching, fill) at 2^k rows."""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools")]
import numpy as np
import torch
import halo2_scaffold_b200 as h2
import evaluate_h_bench

k = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << k
L = h2.load(); L.init_device(0)
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
pts = torch.empty(n * 8, dtype=torch.int64, device=dev)
enc = torch.empty(n * 4, dtype=torch.int64, device=dev)
L.gen_points_dev(0, 3, n, pts.data_ptr(), st)
L.check(L.L.h2b_g1_encode_dev(0, pts.data_ptr(), n, enc.data_ptr(), st))
L.check(L.L.h2b_g1_decode_dev(0, enc.data_ptr(), n, 0, pts.data_ptr(), None, st))
del pts, enc
evaluate_h_bench.measure(L, k, k + 2, 8, 2, reps=1, cpu_rows=0)
cols = [torch.empty(n * 4, dtype=torch.int64, device=dev) for _ in range(8)]
for j, c in enumerate(cols):
    L.gen_scalars_dev(0, 100 + j, n, 0, c.data_ptr(), st)
sc = L.gen_scalars(5, 6)
L.permutation_product_dev(0, [c.data_ptr() for c in cols[:3]], [c.data_ptr() for c in cols[3:6]], n, sc[0], sc[1], sc[2], sc[3], sc[4], sc[5], cols[6].data_ptr(), st)
a = cols[0].clone()
a.view(-1, 4)[: n - 6] = cols[0].view(-1, 4)[: n - 6].flip(0)
L.lookup_permute_dev(0, a.data_ptr(), cols[0].data_ptr(), n - 6, cols[6].data_ptr(), cols[7].data_ptr(), st)
torch.cuda.synchronize()
print("ncu_target_prover done", k, L.launch_count())
