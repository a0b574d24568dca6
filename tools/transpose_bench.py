import sys, os, time, json
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import halo2_scaffold_b200 as h2
L = h2.load(); L.init_device(0)
st = torch.cuda.current_stream().cuda_stream
rows, cols = 4096, 4096
a = torch.randint(0, 2**62, (rows * cols * 4,), dtype=torch.int64, device="cuda")
b = torch.empty_like(a)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    L.check(L.L.h2b_fr_transpose_dev(0, a.data_ptr(), b.data_ptr(), rows, cols, st))
torch.cuda.synchronize()
e0.record()
for _ in range(20):
    L.check(L.L.h2b_fr_transpose_dev(0, a.data_ptr(), b.data_ptr(), rows, cols, st))
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
ok = bool((b.view(cols, rows, 4) == a.view(rows, cols, 4).transpose(0, 1)).all())
print(json.dumps({"op": "fr_transpose 4096 x 4096 (512 MiB)", "bulk": os.environ.get("H2B_TRANSPOSE_BULK", "1"), "ms": round(ms, 4), "gb_s": round(2 * rows * cols * 32 / ms / 1e6, 1), "ok": ok}))
