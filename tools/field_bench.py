#!/usr/bin/env python3
"""Throughput of the field / group primitives on the GPU (dependent chains, every SM full): one JSON line."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo2_scaffold_b200 as h2
L = h2.load(); L.init_device(0)
out = {}
for kind, name in ((0, "imad32_T/s"), (2, "fq_mul_G/s"), (4, "fq_sqr_G/s"), (5, "fq_mul2_G/s"), (3, "xyzz_madd_G/s")):
    ms, ops = L.imad_bench(kind, 4096 if kind else 4096)
    out[name] = round(ops / ms / (1e9 if kind == 0 else 1e6), 2)
print(json.dumps(out))
