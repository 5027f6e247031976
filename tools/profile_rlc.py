#!/usr/bin/env python
"""One RLC verification of 65,536 BatchableProof README-4 presentations (for `ncu --metrics gpu__time_duration.sum`)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from aeonflux_b200 import Issuer, PresentationBatch, compact_to_batchable  # noqa: E402

B = 65536
sp, ip, sk, items = bench.load_fixture(B)
iss = Issuer(sp, ip, sk, device=0, max_batch=B)
comp = PresentationBatch.from_items(bench.KINDS_README4, items)
_, dbg = iss.verify_batch(comp, debug=True)
bb = PresentationBatch(bench.KINDS_README4, compact_to_batchable(bench.KINDS_README4, comp.fields, dbg["commitments"]))
for _ in range(2):
    v, fb = iss.verify_batchable_rlc(bb, bytes(range(32)))
print("rejected", int(v.sum()), "fallback chunks", fb)
