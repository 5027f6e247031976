#!/usr/bin/env python
"""Single-process form of the multi-GPU path: one process, one context and one host thread per visible B200 (MultiGpuIssuer),
end to end through the host-buffer C ABI (H2D + kernels + D2H per device pass).  The driver's contract measures one process per
GPU (bench.py under torchrun); this is the figure for a caller that is one process, like the reference's library users.
    python tools/bench_threads.py [steps] [items_per_gpu]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from aeonflux_b200 import PresentationBatch  # noqa: E402
from aeonflux_b200.shard import MultiGpuIssuer  # noqa: E402
from bench import KINDS_README4, load_fixture  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
per = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
G = torch.cuda.device_count()
sp, ip, sk, items = load_fixture(per)
m = MultiGpuIssuer(sp, ip, sk, devices=list(range(G)), max_batch=per)
fields = m.host_array((28, G * per, 32))
for k in range(G):
    fields[:, k * per:(k + 1) * per] = np.roll(items, k * 131, axis=0).transpose(1, 0, 2)
batch = PresentationBatch(KINDS_README4, fields)
for _ in range(3):
    assert not m.verify_batch(batch).any()
t0 = time.perf_counter()
for _ in range(steps):
    v = m.verify_batch(batch)
dt = time.perf_counter() - t0
assert not v.any()
print(json.dumps({"workload": "Issuer::verify of %d README-4 presentations per call, one process, %d host threads / contexts / GPUs, host buffers in afx_host_alloc memory" % (G * per, G),
                  "n_gpus": G, "steps": steps, "ms_per_call": 1e3 * dt / steps, "presentations_per_s": steps * G * per / dt}))
m.close()
