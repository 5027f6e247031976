#!/usr/bin/env python
"""Experiment: streamed Issuer::verify through afx_verify_presentations_submit / afx_wait with ONE context (two submissions in
flight, kernels serialised on its execute stream) versus TWO contexts on the same GPU used alternately (each has its own
workspace and streams, so the point jobs of one batch fill the grid tail of the other batch's ladders).
    python tools/stream_two_contexts.py [steps]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from aeonflux_b200 import Issuer, PresentationBatch  # noqa: E402
from bench import KINDS_README4, load_fixture  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 16
B = 65536
sp, ip, sk, items = load_fixture(B)
torch.cuda.set_device(0)


def run(n_ctx, depth):
    ctxs = [Issuer(sp, ip, sk, device=0, max_batch=B) for _ in range(n_ctx)]
    batches = [PresentationBatch.from_items(KINDS_README4, items, host_array=c.host_array) for c in ctxs]
    for c, b in zip(ctxs, batches):
        assert not c.submit(b).wait().any()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pend = []
    for k in range(steps):
        pend.append(ctxs[k % n_ctx].submit(batches[k % n_ctx]))
        if len(pend) == depth:
            assert not pend.pop(0).wait().any()
    while pend:
        assert not pend.pop(0).wait().any()
    dt = time.perf_counter() - t0
    for c in ctxs:
        c.close()
    return steps * B / dt


out = {"one_context_two_in_flight": run(1, 2), "two_contexts_one_in_flight_each": run(2, 2), "two_contexts_two_in_flight_each": run(2, 4)}
print(json.dumps(out))
