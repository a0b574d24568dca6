#!/usr/bin/env python3
"""Small workload that touches every kernel once, for compute-sanitizer (memcheck / racecheck) runs:
   compute-sanitizer --tool memcheck python tools/sanitizer_target.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
os.environ.setdefault("H2B_MSM_UPLOAD_CHUNK_LOG", "11")      # exercise the chunked upload + merge path at small n
import numpy as np

import oracle_c as oc
import parity_cases as pc
from halo2_scaffold_b200._lib import Lib

oc.build()
L = Lib()
L.init_device(0)
for k in (3, 9, 12, 14):
    pc.check_ntt(L, oc, k)
pc.check_msm(L, oc, 5000, kind=0, windows=(0, 6))
pc.check_msm(L, oc, 5000, kind=1)
pc.check_msm_tables(L, oc, 6000, 8, kind=0, windows=(0, 4), ranges=[(0, 6000), (100, 4000)])
pc.check_msm_tables(L, oc, 5000, 10, kind=1)
pc.check_msm_single_bucket(L, oc, 40000, scalar=1, tables=True)
pc.check_field(L, oc, 256)
pc.check_group(L, oc, 32)
print("SANITIZER_TARGET_OK")
