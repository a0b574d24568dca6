#!/usr/bin/env python3
"""SRS file -> device-resident base sets (h2b_srs_read): time of the whole call per format, of the decode kernel alone
(device-resident bytes, CUDA events) and of the CPU restatement's decompression on a bounded sample (all host threads,
as upstream's `parallelize` uses).  usage: python tools/srs_bench.py [k ...]"""
import json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np
import torch
import halo2_scaffold_b200 as h2
from halo2_scaffold_b200.kzg import SerdeFormat

L = h2.load(); L.init_device(0)
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
tmp = tempfile.mkdtemp(prefix="h2b_srs_")
for k in [int(a) for a in (sys.argv[1:] or ["16", "20", "22"])]:
    n = 1 << k
    out = {"k": k}
    g, gl = L.gen_points(11 + k, n), L.gen_points(12 + k, n)
    # decode kernel alone
    d_pts = torch.from_numpy(g.view(np.int64).reshape(-1)).to(dev)
    d_enc = torch.empty(n * 4, dtype=torch.int64, device=dev)
    d_dec = torch.empty(n * 8, dtype=torch.int64, device=dev)
    L.check(L.L.h2b_g1_encode_dev(0, d_pts.data_ptr(), n, d_enc.data_ptr(), st))
    for name, src, fmt in (("decompress", d_enc, 0), ("check_raw", d_pts, 1)):
        for _ in range(2):
            L.check(L.L.h2b_g1_decode_dev(0, src.data_ptr(), n, fmt, d_dec.data_ptr(), None, st))
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            L.check(L.L.h2b_g1_decode_dev(0, src.data_ptr(), n, fmt, d_dec.data_ptr(), None, st))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        out[name + "_kernel_ms"] = round(ms, 3)
        out[name + "_points_per_s"] = n / ms * 1e3
    assert torch.equal(d_dec, d_pts)
    # square root = a^((p + 1) / 4): bit_length - 1 squarings (74.75 G/s measured) + popcount - 1 multiplications (68.06 G/s), + 6
    e = (0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47 + 1) // 4
    per_point_s = (e.bit_length() - 1 + 2) / 74.75e9 + (bin(e).count("1") - 1 + 4) / 68.06e9
    out["decompress_frac_of_fq_rate"] = round(out["decompress_points_per_s"] * per_point_s, 3)
    del d_pts, d_enc, d_dec
    params = h2.ParamsKZG(k, g, gl, lib=L, g2_bytes=bytes(256))
    for fmt, name in ((SerdeFormat.Processed, "processed"), (SerdeFormat.RawBytes, "raw_bytes"), (SerdeFormat.RawBytesUnchecked, "raw_bytes_unchecked")):
        path = os.path.join(tmp, "kzg_bn254_%d.srs" % k)
        params.g2_bytes = bytes(128 if fmt == 0 else 256)
        params.write_custom(path, fmt)
        out[name + "_file_bytes"] = os.path.getsize(path)
        L.srs_cache_clear()
        t0 = time.perf_counter()
        r = L.srs_read(path, fmt, want_host=True, register=True)            # decode + window tables + host copies
        out[name + "_srs_read_s"] = round(time.perf_counter() - t0, 4)
        assert (r["g"] == g).all() and (r["g_lagrange"] == gl).all()
        L.unregister_bases(r["handle_g"]); L.unregister_bases(r["handle_g_lagrange"])
        t0 = time.perf_counter()
        r = L.srs_read(path, fmt, want_host=True, register=True)            # unchanged file: resident sets handed out again
        out[name + "_srs_reread_s"] = round(time.perf_counter() - t0, 4)
        assert (r["g"] == g).all()
        L.unregister_bases(r["handle_g"]); L.unregister_bases(r["handle_g_lagrange"])
        t0 = time.perf_counter()
        r = L.srs_read(path, fmt, want_host=False, register=True)
        out[name + "_srs_reread_device_only_s"] = round(time.perf_counter() - t0, 6)
        L.unregister_bases(r["handle_g"]); L.unregister_bases(r["handle_g_lagrange"])
        L.srs_cache_clear()
        os.unlink(path)
    params.close()
    # CPU restatement: decompression of a sample with every host thread
    import oracle_c as oc
    m = min(n, 1 << 16)
    enc = oc.g1_to_bytes(g[:m])
    t0 = time.perf_counter()
    dec, first = oc.g1_from_bytes(enc)
    cpu = time.perf_counter() - t0
    assert first == m and (dec == g[:m]).all()
    out["cpu_restatement_decompress_points_per_s"] = m / cpu
    out["cpu_threads"] = oc.hardware_threads()
    print(json.dumps(out), flush=True)
