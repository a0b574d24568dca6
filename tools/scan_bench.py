#!/usr/bin/env python3
"""Device-resident batch inversion and exclusive prefix product of an Fr column (grand-product building blocks):
ms and GB/s of algorithmic traffic (64 B per element: read + write once).  usage: python tools/scan_bench.py [k ...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import halo2_scaffold_b200 as h2
L = h2.load(); L.init_device(0)
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for k in [int(a) for a in (sys.argv[1:] or ["20", "22", "24"])]:
    n = 1 << k
    d = torch.empty(n * 4, dtype=torch.int64, device=dev)
    L.gen_scalars_dev(0, 7 + k, n, 0, d.data_ptr(), st)
    out = {"k": k}
    import numpy as np
    x = np.array([0x1234567, 0x89abcdef, 0x55, 0x1], dtype=np.uint64)
    d_q = torch.empty(n * 4, dtype=torch.int64, device=dev)
    d_o = torch.empty(4, dtype=torch.int64, device=dev)
    for name, fn in (("batch_invert", lambda: L.fr_batch_invert_dev(0, d.data_ptr(), n, st)), ("prefix_product", lambda: L.fr_prefix_product_dev(0, d.data_ptr(), d.data_ptr(), n, st)),
                     ("eval_polynomial", lambda: L.check(L.L.h2b_fr_eval_polynomial_dev(0, d.data_ptr(), n, x.ctypes.data, d_o.data_ptr(), st))),
                     ("kate_division", lambda: L.check(L.L.h2b_fr_kate_division_dev(0, d.data_ptr(), n, x.ctypes.data, d_q.data_ptr(), st)))):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        out[name + "_ms"] = round(ms, 4)
        out[name + "_elements_per_s"] = n / ms * 1e3
        out[name + "_algorithmic_gb_s"] = round(64.0 * n / ms / 1e6, 1)
    # the grand products themselves: one permutation set of 3 columns (cs.degree() = 5), one lookup argument
    cols = [torch.empty(n * 4, dtype=torch.int64, device=dev) for _ in range(6)]
    for j, c in enumerate(cols):
        L.gen_scalars_dev(0, 100 + j, n, 0, c.data_ptr(), st)
    sc = L.gen_scalars(5, 6)
    for name, fn in (("permutation_product_3_columns", lambda: L.permutation_product_dev(0, [c.data_ptr() for c in cols[:3]], [c.data_ptr() for c in cols[3:]], n, sc[0], sc[1], sc[2], sc[3], sc[4], sc[5], d_q.data_ptr(), st)),
                     ("lookup_product", lambda: L.lookup_product_dev(0, cols[0].data_ptr(), cols[1].data_ptr(), cols[2].data_ptr(), cols[3].data_ptr(), n, sc[0], sc[1], d_q.data_ptr(), st))):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        out[name + "_ms"] = round(ms, 4)
        out[name + "_rows_per_s"] = n / ms * 1e3
    # permute_expression_pair: range-table lookup (values < 2^16, as LOOKUP_BITS-sized tables) and full-width random values
    usable = n - 6
    for name, bits in (("lookup_permute_range_table", 16), ("lookup_permute_full_width", 0)):
        if bits:
            import numpy as _np
            rng_ = min(1 << bits, usable)
            tab = torch.arange(n, device=dev, dtype=torch.int64) % rng_
            inp = (torch.arange(n, device=dev, dtype=torch.int64) * 2654435761) % rng_
            def to_cols(v):
                c = torch.zeros(n * 4, dtype=torch.int64, device=dev)
                c[0::4] = v
                # canonical small integers -> Montgomery form on the device: a Montgomery product with the raw limbs of R^2 mod r
                r2 = L.field_op("fr", "to_mont", L.field_op("fr", "to_mont", _np.array([[1, 0, 0, 0]], dtype=_np.uint64)))
                L.fr_scale_dev(0, c.data_ptr(), n, r2, st)
                return c
            a_col, t_col = to_cols(inp), to_cols(tab)
        else:
            t_col = cols[0]
            a_col = cols[0].clone()
            a_col.view(-1, 4)[:usable] = cols[0].view(-1, 4)[:usable].flip(0)     # a permutation of the table's usable rows
        for _ in range(2):
            L.lookup_permute_dev(0, a_col.data_ptr(), t_col.data_ptr(), usable, cols[4].data_ptr(), cols[5].data_ptr(), st)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(3):
            L.lookup_permute_dev(0, a_col.data_ptr(), t_col.data_ptr(), usable, cols[4].data_ptr(), cols[5].data_ptr(), st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        out[name + "_ms"] = round(ms, 4)
        out[name + "_rows_per_s"] = n / ms * 1e3
    del cols
    print(json.dumps(out), flush=True)
