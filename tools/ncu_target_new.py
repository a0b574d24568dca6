#!/usr/bin/env python3
"""ncu target for the kernels added late in round 2: block-scan bucket reduction (a 2^20-point MSM), TMA transpose (4096 x 4096), three-distance bitonic sweep (2^20-row lookup permutation)."""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np
import torch
import halo2_scaffold_b200 as h2

torch.cuda.set_device(0)
L = h2.load(); L.init_device(0)
st = torch.cuda.current_stream().cuda_stream
n = 1 << 20
d_s = torch.empty(n * 4, dtype=torch.int64, device="cuda")
d_b = torch.empty(n * 8, dtype=torch.int64, device="cuda")
d_o = torch.empty(28, dtype=torch.int64, device="cuda")
L.gen_scalars_dev(0, 1, n, 0, d_s.data_ptr(), st)
L.gen_points_dev(0, 2, n, d_b.data_ptr(), st)
torch.cuda.synchronize()
handle = L.register_bases(d_b.cpu().numpy().view(np.uint64))
for _ in range(2):
    L.msm_dev_registered(0, d_s.data_ptr(), handle, 0, n, d_o.data_ptr(), st)
a = torch.randint(0, 2 ** 62, (4096 * 4096 * 4,), dtype=torch.int64, device="cuda")
b = torch.empty_like(a)
for _ in range(2):
    L.check(L.L.h2b_fr_transpose_dev(0, a.data_ptr(), b.data_ptr(), 4096, 4096, st))
del a, b
cols = [torch.empty(n * 4, dtype=torch.int64, device="cuda") for _ in range(3)]
L.gen_scalars_dev(0, 9, n, 0, cols[0].data_ptr(), st)
inp = cols[0].clone()
inp.view(-1, 4)[: n - 6] = cols[0].view(-1, 4)[: n - 6].flip(0)
for _ in range(2):
    L.lookup_permute_dev(0, inp.data_ptr(), cols[0].data_ptr(), n - 6, cols[1].data_ptr(), cols[2].data_ptr(), st)
torch.cuda.synchronize()
print("ncu_target_new done", L.launch_count())
