#!/usr/bin/env python
"""Single-process form of the multi-GPU path: ONE process drives every visible B200 through the C ABI's multi-device handle
(afx_multi_*: replicated context, one host thread + stream set per device inside the library, contiguous item slices), end to end
through host buffers (H2D + kernels + D2H per device).  The driver's contract measures one process per GPU (bench.py under torchrun);
this is the figure for a caller that is one process, like the reference's library users.
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
wire = m.host_array((G * per, 28, 32))
for k in range(G):
    wire[k * per:(k + 1) * per] = np.roll(items, k * 131, axis=0)
wire[7, 1, 0] ^= 1; wire[G * per - 3, 9, 31] ^= 0x80
expect = np.zeros(G * per, np.uint8); expect[[7, G * per - 3]] = 1
for _ in range(3):
    assert (m.verify_wire(KINDS_README4, wire) == expect).all()
t0 = time.perf_counter()
for _ in range(steps):
    v = m.verify_wire(KINDS_README4, wire)
dt = time.perf_counter() - t0
assert (v == expect).all()
fields = m.host_array((28, G * per, 32))
fields[:] = wire.transpose(1, 0, 2)
batch = PresentationBatch(KINDS_README4, fields)
assert (m.verify_batch(batch) == expect).all()
t0 = time.perf_counter()
for _ in range(steps):
    v = m.verify_batch(batch)
dt_soa = time.perf_counter() - t0
print(json.dumps({"workload": "Issuer::verify of %d README-4 presentations per call through afx_multi_verify_presentations_wire: one process, the library's %d device threads / contexts / GPUs, "
                              "host buffers in afx_host_alloc memory, 2 corrupted items" % (G * per, G),
                  "n_gpus": G, "steps": steps, "ms_per_call": 1e3 * dt / steps, "presentations_per_s": steps * G * per / dt,
                  "soa_presentations_per_s": steps * G * per / dt_soa}))
m.close()
