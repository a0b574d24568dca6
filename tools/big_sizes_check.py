#!/usr/bin/env python3
"""One-off checks at the maximum sizes of the C ABI on one B200 (not part of the test-suite: tens of GB):
  * NTT at 2^27 and 2^28 (the two-adicity limit of Fr): iNTT(NTT(a)) == n * a on samples + a sparse-input spot check;
  * MSM at 2^27 points (more than one 2^26 chunk, table-less mode): the O(n) checksum MSM(s, [z_i]G) == [sum s_i z_i]G.
Prints one JSON line per check."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np
import torch

import bn254 as o
import oracle_c as oc
import parity_cases as pc
from halo2_scaffold_b200._lib import Lib

oc.build()
L = Lib()
L.init_device(0)
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream


def fr_words(x):
    return oc.ints_to_words([o.to_mont(x % o.R_MOD, o.R_MOD)])[0]


for k in [int(a) for a in (sys.argv[1:] or ["27", "28"])]:
    n = 1 << k
    w_int = o.omega_for(k)
    w, wi = fr_words(w_int), fr_words(pow(w_int, -1, o.R_MOD))
    d = torch.empty(n * 4, dtype=torch.int64, device=dev)
    L.gen_scalars_dev(0, 0xA000 + k, n, 0, d.data_ptr(), st)
    torch.cuda.synchronize()
    idx = torch.arange(0, n, 65537, device=dev)
    before = d.view(n, 4)[idx].cpu().numpy().view(np.uint64)
    t0 = time.perf_counter()
    L.ntt_dev(0, d.data_ptr(), w, k, st)
    L.ntt_dev(0, d.data_ptr(), wi, k, st)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3
    after = d.view(n, 4)[idx].cpu().numpy().view(np.uint64)
    ninv = fr_words(pow(n, -1, o.R_MOD))
    ok_rt = bool((oc.fr_scale(np.ascontiguousarray(after), ninv) == before).all())
    # sparse input: out[i] = sum_j a[j] w^(ij)
    d.zero_()
    pos = [0, 1, 12345, n // 2 + 7, n - 1]
    vals = o.random_fr(5, 5)
    dv = d.view(n, 4)
    for pp, v in zip(pos, vals):
        dv[pp] = torch.from_numpy(fr_words(v).view(np.int64)).to(dev)
    L.ntt_dev(0, d.data_ptr(), w, k, st)
    torch.cuda.synchronize()
    ok_sp = True
    for i in (0, 1, 2, 77777, n // 2, n - 1):
        want = sum(v * pow(w_int, (i * j) % n, o.R_MOD) for v, j in zip(vals, pos)) % o.R_MOD
        got = oc.words_to_ints(dv[i:i + 1].cpu().numpy().view(np.uint64))[0]
        ok_sp = ok_sp and got == o.to_mont(want, o.R_MOD)
    print(json.dumps({"check": "ntt", "k": k, "round_trip_ok": ok_rt, "sparse_spot_ok": bool(ok_sp), "two_transforms_ms_incl_tables": round(ms, 1)}), flush=True)
    del d, dv
    torch.cuda.empty_cache()

# MSM at 2^27 (two 2^26 chunks into shared buckets), scalars < 2^64 so that the checksum is cheap on the host
k = 27
n = 1 << k
seed_p = 0xB2001000 + k
d_p = torch.empty(n * 8, dtype=torch.int64, device=dev)
L.gen_points_dev(0, seed_p, n, d_p.data_ptr(), st)
rng = np.random.default_rng(7)
small = rng.integers(0, 1 << 62, size=n, dtype=np.uint64)
s = np.zeros((n, 4), dtype=np.uint64)
s[:, 0] = small
s = oc.fr_to_mont(s)
d_s = torch.from_numpy(s.view(np.int64)).to(dev)
d_o = torch.empty(12, dtype=torch.int64, device=dev)
torch.cuda.synchronize()
t0 = time.perf_counter()
L.msm_dev(0, d_s.data_ptr(), d_p.data_ptr(), n, d_o.data_ptr(), st)
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) * 1e3
out = d_o.cpu().numpy().view(np.uint64)
idx = np.arange(1, n + 1, dtype=np.uint64)
with np.errstate(over="ignore"):
    z = np.uint64(seed_p) + idx * np.uint64(0x9E3779B97F4A7C15)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    z = z ^ (z >> np.uint64(31))
# sum s_i z_i mod r with Python ints in blocks (n = 2^27: vectorised 64x64 -> 128 via two halves)
lo = (small & np.uint64(0xFFFFFFFF)).astype(object)
hi = (small >> np.uint64(32)).astype(object)
zo = z.astype(object)
t = 0
B = 1 << 20
for a in range(0, n, B):
    t += int((lo[a:a + B] * zo[a:a + B]).sum()) + (int((hi[a:a + B] * zo[a:a + B]).sum()) << 32)
t %= o.R_MOD
want = o.g1_mul(o.G1_GEN, t)
gw = oc.words_to_ints(pc.affine_of(oc, out).reshape(2, 4))
got = (o.from_mont(gw[0], o.P_MOD), o.from_mont(gw[1], o.P_MOD))
print(json.dumps({"check": "msm", "k": k, "mode": "plain, 2 chunks of 2^26 merged into shared buckets", "checksum_ok": bool(got == want), "ms": round(ms, 1)}), flush=True)
