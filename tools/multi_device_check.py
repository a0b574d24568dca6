#!/usr/bin/env python3
"""In-process multi-GPU paths of libh2b200, checked against the oracle on every visible B200:
  * one host-pointer MSM split by point range across the devices (partials folded on device 0),
  * one host-pointer NTT split over the devices (four-step: column blocks, one exchange of 2-D peer copies, row blocks),
  * independent NTTs / MSMs issued from concurrent host threads, spread round-robin over the devices.
Prints MULTI_DEVICE_OK <n_devices>."""
import os
import sys
import threading

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np

import oracle_c as oc
import parity_cases as pc
from halo2_scaffold_b200._lib import Lib

oc.build()
L = Lib()
L.init(0)                                  # every visible device
nd = L.device_count()
n = 1 << 20                                # >= 2^18 and a whole number of digest blocks: cached, sharded across the devices
s = L.gen_scalars(5, n, 0)
P = L.gen_points(6, n)
want = pc.affine_of(oc, oc.best_multiexp(s, P))
for _ in range(3):                         # 1st call plain upload, 2nd builds the window tables on every device, 3rd reuses them
    assert (pc.affine_of(oc, L.msm(s, P)) == want).all()
st = L.implicit_cache_stats()
assert st["uploads"] == 1 and st["hits"] == 2 and st["sets_with_tables"] == 1, st
hs = L.register_bases_sharded(P)           # explicit point-range residency: device d keeps rows [d n/D, (d+1) n/D) and their tables
info = L.base_set_info(hs)
assert info["device_bytes"] <= info["n_tables"] * ((n + nd - 1) // nd + 1) * 64, info
assert (pc.affine_of(oc, L.msm_registered(s, hs)) == want).all()
assert (pc.affine_of(oc, L.msm_registered(s[:300001], hs, 77)) == pc.affine_of(oc, oc.best_multiexp(s[:300001], P[77:300078]))).all()
L.unregister_bases(hs)
h = L.register_bases(P)
assert (pc.affine_of(oc, L.msm_registered(s, h)) == want).all()
off, m = 12345, 1 << 19
assert (pc.affine_of(oc, L.msm_registered(s[:m], h, off)) == pc.affine_of(oc, oc.best_multiexp(s[:m], P[off:off + m]))).all()
# batched columns, round-robin over the devices (results identical to single calls)
cols = [L.gen_scalars(400 + j, (1 << 18) - 1000 * j, j % 2) for j in range(2 * nd + 1)]
got = L.msm_batch_registered(cols, h)
for j, c in enumerate(cols):
    assert (pc.affine_of(oc, got[j]) == pc.affine_of(oc, oc.best_multiexp(c, P[:c.shape[0]]))).all(), j
L.unregister_bases(h)
polys = [oc.random_fr(500 + j, 1 << 16) for j in range(2 * nd + 1)]
wantp = [oc.best_fft(a, pc.omega_words(oc, 16), 16) for a in polys]
L.ntt_batch(polys, pc.omega_words(oc, 16), 16)
for a, w_ in zip(polys, wantp):
    assert (a == w_).all()

# ONE NTT across the devices (four-step split of h2b_ntt_bn254_fr; a power-of-two device count): forward, inverse, odd log_n
if nd >= 2 and nd & (nd - 1) == 0:
    for k in (22, 23):
        a = oc.random_fr(700 + k, 1 << k)
        for inverse in (False, True):
            w = pc.omega_words(oc, k, inverse)
            assert (L.ntt(a.copy(), w, k) == oc.best_fft(a, w, k)).all(), ("one NTT across devices", k, inverse)

# concurrent callers: 2 * nd threads, each an NTT round trip and a small MSM
errs = []


def worker(t):
    try:
        k = 14 + (t % 3)
        a = oc.random_fr(100 + t, 1 << k)
        w = pc.omega_words(oc, k)
        got = L.ntt(a.copy(), w, k)
        assert (got == oc.best_fft(a, w, k)).all()
        ss, PP = oc.random_fr(200 + t, 3000), oc.gen_points(300 + t, 3000)
        assert (pc.affine_of(oc, L.msm(ss, PP)) == pc.affine_of(oc, oc.best_multiexp(ss, PP))).all()
    except Exception as e:      # noqa: BLE001
        errs.append((t, repr(e)))


ths = [threading.Thread(target=worker, args=(t,)) for t in range(2 * nd)]
for t in ths:
    t.start()
for t in ths:
    t.join()
assert not errs, errs

# SRS file -> base sets resident on EVERY device (decoded on device 0, peer-copied, tables built per device); a sharded MSM over them
import tempfile
from halo2_scaffold_b200 import evaluation as ev
k = 18
g, gl = P[: 1 << k], P[1 << k: 2 << k] if n >= (2 << k) else P[: 1 << k][::-1].copy()
path = os.path.join(tempfile.mkdtemp(prefix="h2b_md_"), "kzg_bn254_%d.srs" % k)
L.srs_write(path, 0, k, g, gl, bytes(128))
r = L.srs_read(path, 0)
assert (r["g"] == g).all() and (r["g_lagrange"] == gl).all()
sk = s[: 1 << k]
assert (pc.affine_of(oc, L.msm_registered(sk, r["handle_g"])) == pc.affine_of(oc, oc.best_multiexp(sk, g))).all()
assert (pc.affine_of(oc, L.msm_registered(sk, r["handle_g_lagrange"])) == pc.affine_of(oc, oc.best_multiexp(sk, gl))).all()
L.unregister_bases(r["handle_g"]); L.unregister_bases(r["handle_g_lagrange"])
L.srs_cache_clear()
# the widened rows on the LAST device of the process: quotient evaluation, grand product, lookup permutation
dev = nd - 1
rng = np.random.default_rng(7)
graph, n_const = pc.random_graph(rng, 2, 3, 1, 1, 120)
gq, ga = pc._graph_pair(oc, graph, n_const, 7)
rows = 5000
cols = [oc.random_fr(900 + j, rows) for j in range(6)]
scq = oc.random_fr(950, 5)
vals = oc.random_fr(951, rows)
want_v = oc.evaluate_graph(gq, cols[:2], cols[2:5], cols[5:], scq[:1], scq[1], scq[2], scq[3], scq[4], vals, 2)
assert (ev.evaluate_graph(L, ga, cols[:2], cols[2:5], cols[5:], scq[:1], scq[1], scq[2], scq[3], scq[4], vals, 2, device=dev) == want_v).all()
w13 = pc.omega_words(oc, 13)
assert (L.permutation_product(cols[:2], cols[2:4], scq[0], scq[1], scq[2], scq[3], w13, scq[4], device=dev) ==
        oc.permutation_product(cols[:2], cols[2:4], scq[0], scq[1], scq[2], scq[3], w13, scq[4])).all()
tab = cols[0]
inp = tab[:rows - 6][rng.integers(0, rows - 6, size=rows)]
got_a, got_t = L.lookup_permute(inp, tab, rows - 6, device=dev)
want_a, want_t = oc.lookup_permute(inp, tab, rows - 6)
assert (got_a == want_a).all() and (got_t == want_t).all()
print("MULTI_DEVICE_OK", nd)
