#!/usr/bin/env python
"""Soak run on one GPU: many back-to-back verify calls with random batch sizes and random corrupted items through every entry
form (SoA, wire, multi-chunk pipelined, device-resident), checking each verdict vector against the expected one -- a guard for
the aMAC -> "Z" constraint flag protocol and the stream pipelines.  python tests/tools/soak.py [iterations]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from aeonflux_b200 import Issuer, PresentationBatch  # noqa: E402
from oracle import coracle as C  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
sp, ip, sk = C.make_issuer(4)
orc = C.Issuer(sp, ip, sk)
kinds, base, _ = orc.synth(b"SSPE", [0, 3], b"soak", 0, 512, want_issuances=False)
rng = np.random.default_rng(99)
iss = Issuer(sp, ip, sk, device=0, max_batch=8192)
stream = torch.cuda.Stream()
checked = 0
for it in range(iters):
    count = int(rng.choice([1, 31, 255, 256, 257, 1000, 4097, 8192, 8193, 20000, 33333]))
    items = base[rng.integers(0, len(base), count)].copy()
    bad = np.unique(rng.integers(0, count, max(1, count // 300)))
    for i in bad:
        items[i, rng.integers(0, 28), rng.integers(0, 32)] ^= 1 << rng.integers(0, 8)
    expect = np.zeros(count, np.uint8); expect[bad] = 1
    mode = it % 3
    if mode == 0:
        v = iss.verify_batch(PresentationBatch.from_items(kinds, items))
    elif mode == 1:
        v = iss.verify_wire(kinds, items)
    else:
        m = min(count, 8192)
        f = torch.from_numpy(np.ascontiguousarray(items[:m].transpose(1, 0, 2))).cuda()
        out = torch.full((m,), 9, dtype=torch.uint8, device="cuda")
        with torch.cuda.stream(stream):
            iss.verify_batch_device(kinds, m, f.data_ptr(), out.data_ptr(), stream.cuda_stream)
        stream.synchronize()
        v, expect = out.cpu().numpy(), expect[:m]
    assert (v == expect).all(), (it, count, mode, np.where(v != expect)[0][:10])
    checked += len(v)
print("soak ok:", iters, "calls,", checked, "verdicts")
