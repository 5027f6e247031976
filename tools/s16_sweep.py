#!/usr/bin/env python
"""S16 (BASELINE configs[3]) pass-size sweep: 65,536 distinct S16 presentations through afx_verify_presentations_wire with the
context's max_batch = 65,536 (one pass: the 300 MB copy is exposed) and 32,768 / 16,384 / 8,192 (pipelined passes: the copy of pass
i+1 under the kernels of pass i), beside the device-resident rate of one pass of that size and the workspace it needs.
    python tools/s16_sweep.py > gpurun_out/s16_sweep.json"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from aeonflux_b200 import Issuer  # noqa: E402
from bench import KINDS_S16, WORDS_S16, load_issuer, synthesize_on_device, time_device, work_model  # noqa: E402

B = 65536
sp, ip, sk = load_issuer("issuer16.bin")
stream = torch.cuda.current_stream()
gen = Issuer(sp, ip, sk, device=0, max_batch=B)
f16 = synthesize_on_device(torch, gen, B, 4242, stream, KINDS_S16, "keypair16.bin")
wire = torch.empty((B, WORDS_S16, 32), dtype=torch.uint8).pin_memory()
wire.copy_(f16.permute(1, 0, 2))
gen.close()
torch.cuda.empty_cache()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
wm = work_model(KINDS_S16)
peak = 148 * 64 * 1965e6
rows = []
for mb in (65536, 32768, 16384, 8192):
    free0 = torch.cuda.mem_get_info()[0]
    iss = Issuer(sp, ip, sk, device=0, max_batch=mb)
    v = iss.verify_wire(KINDS_S16, wire.numpy())
    assert not v.any()
    used = free0 - torch.cuda.mem_get_info()[0]
    t0 = time.perf_counter()
    for _ in range(3):
        iss.verify_wire(KINDS_S16, wire.numpy())
    e2e = (time.perf_counter() - t0) / 3
    sub = f16[:, :mb].contiguous()
    vd = torch.empty(mb, dtype=torch.uint8, device="cuda")
    iss.set_stage_timing(True)
    ms = time_device(torch, stream, flush, lambda: iss.verify_batch_device(KINDS_S16, mb, sub.data_ptr(), vd.data_ptr(), stream.cuda_stream), 3)
    st = iss.stage_times_ms()
    rows.append({"max_batch": mb, "e2e_wire_65536_items_per_s": B / e2e, "e2e_ms": e2e * 1e3, "device_resident_items_per_s": mb / (ms * 1e-3), "device_ms_per_pass": ms,
                 "frac_of_imad_peak": 2 * wm["total"] * mb / (ms * 1e-3) / peak, "stage_ms": st, "context_device_bytes": int(used)})
    iss.close()
    del sub, vd
    torch.cuda.empty_cache()
print(json.dumps({"workload": "65,536 distinct S16 presentations, afx_verify_presentations_wire from pinned memory", "rows": rows}, indent=1))
