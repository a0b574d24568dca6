#!/usr/bin/env python3
"""The multi-GPU modes a single prover process uses (`h2b_init(D)`), measured and checked in ONE process:

  parity : one host-pointer MSM split by point range over D devices (implicit cache -> sharded resident copy; sharded
           registered set), batched columns round-robin over the devices, batched NTTs, ONE NTT split over the devices -- each
           against the CPU oracle
           (2^18 .. 2^20 points, sizes the oracle finishes in seconds);
  strong : ONE MSM of fixed total size 2^k over D devices through the host-pointer entry point with pageable scalars
           (what a Rust Vec is): sharded registration time, end-to-end ms, and the O(n) checksum [sum s_i z_i] G of the result.

Prints one JSON object.  usage: python tools/multi_gpu_inprocess.py D [k ...]        (bench.py runs it at N > 1, rank 0)
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np

from halo2_scaffold_b200._lib import Lib
from halo2_scaffold_b200 import verify as V


def main():
    D = int(sys.argv[1])
    ks = [int(a) for a in sys.argv[2:]]
    L = Lib()
    L.init(D)
    assert L.device_count() == D
    out = {"devices": D}

    # ---- parity against the oracle ----------------------------------------------------------------------------------
    import oracle_c as oc
    import parity_cases as pc
    oc.build()
    t0 = time.perf_counter()
    n = 1 << 19
    s = L.gen_scalars(5, n, 0)
    P = L.gen_points(6, n)
    want = pc.affine_of(oc, oc.best_multiexp(s, P))
    checks = {}
    ok = all((pc.affine_of(oc, L.msm(s, P)) == want).all() for _ in range(3))      # upload (sharded), tables, reuse
    st = L.implicit_cache_stats()
    checks["implicit_sharded_msm"] = bool(ok and st["uploads"] == 1 and st["hits"] == 2)
    hs = L.register_bases_sharded(P)
    info = L.base_set_info(hs)
    ok = (pc.affine_of(oc, L.msm_registered(s, hs)) == want).all()
    m, off = 200001, 77
    ok = ok and (pc.affine_of(oc, L.msm_registered(s[:m], hs, off)) == pc.affine_of(oc, oc.best_multiexp(s[:m], P[off:off + m]))).all()
    checks["registered_sharded_msm"] = bool(ok and info["device_bytes"] <= info["n_tables"] * ((n + D - 1) // D + 1) * 64)
    L.unregister_bases(hs)
    h = L.register_bases(P)
    cols = [L.gen_scalars(400 + j, (1 << 16) - 100 * j, j % 2) for j in range(2 * D + 1)]
    got = L.msm_batch_registered(cols, h)
    checks["batched_columns_round_robin"] = bool(all((pc.affine_of(oc, got[j]) == pc.affine_of(oc, oc.best_multiexp(c, P[:c.shape[0]]))).all()
                                                     for j, c in enumerate(cols)))
    L.unregister_bases(h)
    polys = [oc.random_fr(500 + j, 1 << 16) for j in range(2 * D + 1)]
    wantp = [oc.best_fft(a, pc.omega_words(oc, 16), 16) for a in polys]
    L.ntt_batch(polys, pc.omega_words(oc, 16), 16)
    checks["batched_ntts_round_robin"] = bool(all((a == w_).all() for a, w_ in zip(polys, wantp)))
    # one NTT split over the D devices (four-step: column blocks, one exchange, row blocks), forward and inverse, pageable array
    if D & (D - 1) == 0:
        a = oc.random_fr(900, 1 << 22)
        ok = True
        for inverse in (False, True):
            w = pc.omega_words(oc, 22, inverse)
            ok = ok and bool((L.ntt(a.copy(), w, 22) == oc.best_fft(a, w, 22)).all())
        checks["one_ntt_across_devices"] = ok
        del a
    out["parity_vs_oracle"] = checks
    out["parity_ok"] = all(checks.values())
    out["parity_s"] = round(time.perf_counter() - t0, 2)
    del s, P, cols, polys

    # ---- strong scaling of one MSM -------------------------------------------------------------------------------------
    strong = []
    for k in ks:
        n = 1 << k
        seed_s, seed_p = 0xB2000000 + k, 0xB2001000 + k
        d_s = L.dev_alloc(0, n * 32)
        d_c = L.dev_alloc(0, 32)
        L.gen_scalars_dev(0, seed_s, n, 0, d_s)
        L.msm_checksum_dev(0, d_s, seed_p, n, d_c)
        L.dev_sync(0)
        s = np.empty((n, 4), dtype=np.uint64)          # pageable
        L.d2h(0, s, d_s)
        c = np.zeros(4, dtype=np.uint64)
        L.d2h(0, c, d_c)
        L.dev_free(0, d_s)
        L.dev_free(0, d_c)
        P = L.gen_points(seed_p, n)
        t0 = time.perf_counter()
        h = L.register_bases_sharded(P)
        reg_ms = (time.perf_counter() - t0) * 1e3
        info = L.base_set_info(h)
        del P
        steps = 5 if k <= 24 else 3

        def timed(arr):
            for _ in range(2):
                r_ = L.msm_registered(arr, h)
            t0_ = time.perf_counter()
            for _ in range(steps):
                r_ = L.msm_registered(arr, h)
            return (time.perf_counter() - t0_) / steps * 1e3, r_
        ms, r = timed(s)                                   # pageable scalars: what a Rust Vec is (host threads stage them through pinned slots)
        ms_pinned, r_pinned = None, r
        try:
            import torch
            sp = torch.empty(n * 4, dtype=torch.int64).pin_memory()
            sp_np = sp.numpy().view(np.uint64).reshape(n, 4)
            sp_np[:] = s
            ms_pinned, r_pinned = timed(sp_np)
            del sp, sp_np
        except Exception:       # noqa: BLE001
            pass
        L.unregister_bases(h)
        want = V.scalar_mul_generator(V.words_to_int(c))
        verified = V.jacobian_words_to_affine(r) == want and V.jacobian_words_to_affine(r_pinned) == want
        strong.append({"k": k, "devices": D, "msm_e2e_ms": round(ms, 3), "points_per_s": n / ms * 1e3,
                       "msm_e2e_ms_pinned_scalars": None if ms_pinned is None else round(ms_pinned, 3),
                       "points_per_s_pinned_scalars": None if ms_pinned is None else n / ms_pinned * 1e3, "sharded_registration_ms": round(reg_ms, 1),
                       "tables": info["n_tables"], "spacing": info["spacing"], "device_bytes": info["device_bytes"], "verified": bool(verified),
                       "scalars": "pageable host memory (msm_e2e_ms) and pinned host memory (msm_e2e_ms_pinned_scalars)"})
        del s
    out["strong"] = strong

    # ---- one NTT across the devices: end to end through h2b_ntt_bn254_fr, pageable and pinned host arrays ----------------------------
    ntt = []
    if D & (D - 1) == 0:
        for k in ks:
            n = 1 << k
            w = pc.omega_words(oc, k)
            a = L.gen_scalars(0xB2000000 + k, n, 0)         # pageable
            row = {"k": k, "devices": D}
            for name in ("pageable", "pinned"):
                arr = a
                if name == "pinned":
                    try:
                        import torch
                        tp = torch.empty(n * 4, dtype=torch.int64).pin_memory()
                        arr = tp.numpy().view(np.uint64).reshape(n, 4)
                        arr[:] = a
                    except Exception:       # noqa: BLE001
                        continue
                L.ntt(arr, w, k)
                steps = 5 if k <= 24 else 3
                t0 = time.perf_counter()
                for _ in range(steps):
                    L.ntt(arr, w, k)
                ms = (time.perf_counter() - t0) / steps * 1e3
                row["ntt_e2e_ms_" + name] = round(ms, 3)
                row["elements_per_s_" + name] = n / ms * 1e3
            ntt.append(row)
            del a
    out["one_ntt_across_devices"] = ntt
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
