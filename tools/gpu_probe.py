#!/usr/bin/env python3
"""
Development probe run on a B200 through gpurun: parity of every layer against the CPU oracle, then
kernel timings (CUDA events on torch's stream) for NTT / MSM sweeps and the IMAD micro-benchmarks.
Prints one JSON object per line.  Usage: python tools/gpu_probe.py [parity] [ntt] [msm] [imad]
"""
import json
import os
import sys
import time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import numpy as np
import torch

import bn254 as o
import oracle_c as c
from halo2_scaffold_b200._lib import load

L = load()
L.init_device(0)
torch.cuda.set_device(0)
what = set(sys.argv[1:]) or {"parity", "ntt", "msm", "imad"}


def emit(**kw):
    print(json.dumps(kw), flush=True)


def omega(k, inv=False):
    w = o.omega_for(k)
    if inv:
        w = pow(w, -1, o.R_MOD)
    return c.ints_to_words([o.to_mont(w, o.R_MOD)])[0]


def timed(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sorted(ts)[len(ts) // 2]


emit(version=L.version(), gpu=torch.cuda.get_device_name(0), cores=c.hardware_threads())

if "parity" in what:
    a, b = c.random_fr(1, 4096), c.random_fr(2, 4096)
    a[0] = 0; b[1] = 0
    res = {}
    for f in ("fr", "fq"):
        for op in ("add", "sub", "mul"):
            res["%s_%s" % (f, op)] = bool((L.field_op(f, op, a, b) == c.field_op(f, op, a, b)).all())
        res["%s_sqr" % f] = bool((L.field_op(f, "sqr", a) == c.field_op(f, "mul", a, a)).all())
    res["fr_from_mont"] = bool((L.field_op("fr", "from_mont", a) == c.fr_from_mont(a)).all())
    res["fr_to_mont"] = bool((L.field_op("fr", "to_mont", a) == c.fr_to_mont(a)).all())
    inv = L.field_op("fq", "inv", a[2:200])
    one = L.field_op("fq", "mul", inv, a[2:200])
    res["fq_inv"] = bool((one == one[0]).all())
    emit(test="field", **res)
    P, Pc = L.gen_points(7, 3000), c.gen_points(7, 3000)
    emit(test="gen_points", ok=bool((P == Pc).all()))
    S, Sc = L.gen_scalars(11, 3000, 0), c.random_fr(11, 3000)
    emit(test="gen_scalars", ok=bool((S == Sc).all()))
    for k in list(range(1, 21)) + [22]:
        x = c.random_fr(100 + k, 1 << k)
        ok = True
        for inv_ in (False, True):
            w = omega(k, inv_)
            got = L.ntt(x.copy(), w, k)
            ok = ok and bool((got == c.best_fft(x, w, k)).all())
        emit(test="ntt", k=k, ok=ok)
    for n, kind in [(1, 0), (2, 0), (3, 0), (33, 0), (1000, 0), (1 << 12, 0), (1 << 14, 1), (1 << 16, 0), (1 << 16, 1), (1 << 18, 0), (1 << 18, 1), ((1 << 17) + 12345, 0)]:
        s = L.gen_scalars(n, n, kind)
        Pn = L.gen_points(n + 1, n)
        if n > 40:
            s[1] = 0; Pn[3] = 0; Pn[5] = Pn[4]; s[5] = s[4]; Pn[9] = Pn[8]
        want = c.g1_to_affine(c.best_multiexp(s, Pn))
        oks = {}
        for cw in ([0] if n < 1000 else [0, 8, 13]):
            L.set_msm_window(cw)
            got = c.g1_to_affine(L.msm(s, Pn))
            oks["c%d" % cw] = bool((got == want).all())
        L.set_msm_window(0)
        emit(test="msm", n=n, kind=kind, **oks)

if "imad" in what:
    for kind, name in [(0, "imad32"), (1, "imad_wide"), (2, "fq_mul_chain"), (3, "xyzz_madd_chain")]:
        ms, ops = L.imad_bench(kind, 2048 if kind < 2 else (4096 if kind == 2 else 512))
        emit(test="imad", kind=name, ms=ms, ops=ops, gops_per_s=ops / ms / 1e6)

if "ntt" in what:
    for k in [16, 18, 20, 22, 24, 26]:
        n = 1 << k
        buf = torch.empty(n * 4, dtype=torch.int64, device="cuda")
        L.gen_scalars_dev(0, 5, n, 0, buf.data_ptr(), torch.cuda.current_stream().cuda_stream)
        w = omega(k)
        st = torch.cuda.current_stream().cuda_stream
        best, med = timed(lambda: L.ntt_dev(0, buf.data_ptr(), w, k, st))
        emit(bench="ntt", k=k, ms_best=best, ms_med=med, elems_per_s=n / best * 1e3, hbm_gbs_alg=64 * n / best / 1e6)
        del buf

if "msm" in what:
    for k in [16, 18, 20, 22, 24]:
        n = 1 << k
        sc = torch.empty(n * 4, dtype=torch.int64, device="cuda")
        pts = torch.empty(n * 8, dtype=torch.int64, device="cuda")
        out = torch.empty(12, dtype=torch.int64, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        L.gen_scalars_dev(0, 5, n, 0, sc.data_ptr(), st)
        t0 = time.time()
        L.gen_points_dev(0, 9, n, pts.data_ptr(), st)
        torch.cuda.synchronize()
        tgen = time.time() - t0
        cands = {16: [0, 9, 10, 11, 12, 13], 18: [0, 11, 12, 13, 14], 20: [0, 12, 13, 14, 15, 16], 22: [0, 14, 15, 16, 17], 24: [0, 15, 16, 17, 18, 20]}[k]
        for cw in cands:
            L.set_msm_window(cw)
            best, med = timed(lambda: L.msm_dev(0, sc.data_ptr(), pts.data_ptr(), n, out.data_ptr(), st), iters=3, warm=1)
            emit(bench="msm", k=k, c=cw, ms_best=best, ms_med=med, points_per_s=n / best * 1e3, gen_s=tgen)
        L.set_msm_window(0)
        # witness-like scalars
        L.gen_scalars_dev(0, 6, n, 1, sc.data_ptr(), st)
        best, med = timed(lambda: L.msm_dev(0, sc.data_ptr(), pts.data_ptr(), n, out.data_ptr(), st), iters=3, warm=1)
        emit(bench="msm_skew", k=k, c=0, ms_best=best, ms_med=med, points_per_s=n / best * 1e3)
        del sc, pts
emit(done=True)
