#!/usr/bin/env python3
"""Where does the end-to-end MSM lose time against the device-resident one?  2^k points (default 24), wall-clock ms of the four
host-pointer variants: {registered SRS, implicit content-addressed cache} x {pinned, pageable scalars}.  One JSON line."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import halo2_scaffold_b200 as h2

k = int(sys.argv[1]) if len(sys.argv) > 1 else 24
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
n = 1 << k
L = h2.load(); L.init_device(0)
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
d_scal = torch.empty(n * 4, dtype=torch.int64, device=dev)
d_base = torch.empty(n * 8, dtype=torch.int64, device=dev)
d_block = torch.empty(28, dtype=torch.int64, device=dev)
L.gen_scalars_dev(0, 1, n, 0, d_scal.data_ptr(), st)
L.gen_points_dev(0, 2, n, d_base.data_ptr(), st)
torch.cuda.synchronize()
bases = d_base.cpu().numpy().view(np.uint64).reshape(n, 8).copy(); del d_base
handle = L.register_bases(bases)
pinned = torch.empty(n * 4, dtype=torch.int64).pin_memory(); pinned.copy_(d_scal)
s_pinned = pinned.numpy().view(np.uint64).reshape(n, 4)
s_pageable = np.array(s_pinned, copy=True)


def wall(fn):
    for _ in range(2):
        fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    return round((time.perf_counter() - t0) / steps * 1e3, 3)


out = {"k": k, "env": {e: os.environ.get(e) for e in ("H2B_DIGEST_THREADS", "H2B_STAGE_THREADS", "H2B_MSM_UPLOAD_CHUNK_LOG", "H2B_MSM_UPLOAD_SCHED", "H2B_STAGE_PIECE_LOG") if os.environ.get(e)}, "host_threads": os.cpu_count()}
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(2):
    L.msm_dev_registered(0, d_scal.data_ptr(), handle, 0, n, d_block.data_ptr(), st)
torch.cuda.synchronize()
e0.record()
for _ in range(steps):
    L.msm_dev_registered(0, d_scal.data_ptr(), handle, 0, n, d_block.data_ptr(), st)
e1.record(); torch.cuda.synchronize()
out["device_resident_ms"] = round(e0.elapsed_time(e1) / steps, 3)
out["registered_pinned_ms"] = wall(lambda: L.msm_registered(s_pinned, handle))
out["registered_pageable_ms"] = wall(lambda: L.msm_registered(s_pageable, handle))
L.unregister_bases(handle)
L.msm(s_pinned, bases)      # upload + digests
L.msm(s_pinned, bases)      # tables are built on the second use
out["implicit_pinned_ms"] = wall(lambda: L.msm(s_pinned, bases))
out["implicit_pageable_ms"] = wall(lambda: L.msm(s_pageable, bases))
out["implicit_cache"] = L.implicit_cache_stats()
print(json.dumps(out), flush=True)
