#!/usr/bin/env python
"""Call latency against batch size (a serving front end verifies what has arrived, not 65,536 items): one synchronous
afx_verify_presentations_wire call from page-locked memory -- copy in, every kernel, verdicts out -- for batches of 1 ... 65,536
distinct README-4 and S16 presentations, wall clock, median of several calls; beside it the per-stage device times of the same pass.
    python tools/latency_sweep.py > gpurun_out/latency_sweep.json"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from aeonflux_b200 import Issuer  # noqa: E402
from bench import KINDS_README4, KINDS_S16, load_issuer, synthesize_on_device  # noqa: E402

B = 65536
COUNTS = [int(x) for x in sys.argv[sys.argv.index("--counts") + 1].split(",")] if "--counts" in sys.argv else [1, 32, 256, 1024, 2048, 4096, 8192, 16384, 32768, 65536]
SPLITS = sys.argv[sys.argv.index("--splits") + 1].split(",") if "--splits" in sys.argv else ["0", "1", "2", "3", "4", "6", "8", "auto"]
SPLITS = [None if x == "auto" else x for x in SPLITS]
stream = torch.cuda.current_stream()
out = {"workload": "one synchronous afx_verify_presentations_wire call per batch (H2D + kernels + D2H), pinned host memory, median wall clock", "rows": []}
for name, kinds, issuer_file, kp in (("README-4", KINDS_README4, "issuer4.bin", "keypair4.bin"), ("S16", KINDS_S16, "issuer16.bin", "keypair16.bin")):
    sp, ip, sk = load_issuer(issuer_file)
    iss = Issuer(sp, ip, sk, device=0, max_batch=B)
    dev = synthesize_on_device(torch, iss, B, 99, stream, kinds, kp)
    wire = torch.empty((B, dev.shape[0], 32), dtype=torch.uint8).pin_memory()
    wire.copy_(dev.permute(1, 0, 2))
    del dev
    w = wire.numpy()
    for count in COUNTS:
        row = {"shape": name, "items": count}
        # AFX_AMAC_SPLIT: 0 = one aMAC ladder per item (the large-batch form), g = the ladder cut into parts of g terms, None = the library's choice
        for split in SPLITS:
            if split is None:
                os.environ.pop("AFX_AMAC_SPLIT", None)
            else:
                os.environ["AFX_AMAC_SPLIT"] = split
            if split not in (None, "0") and (count > 32768 or (count > 8192 and int(split) < 4)):
                continue
            reps = 15 if count <= 4096 else 5
            assert not iss.verify_wire(kinds, w[:count]).any()
            ts = []
            for _ in range(reps):
                t0 = time.perf_counter()
                iss.verify_wire(kinds, w[:count])
                ts.append(time.perf_counter() - t0)
            ms = 1e3 * float(np.median(ts))
            key = "auto" if split is None else "unsplit" if split == "0" else "split_%s" % split
            row["ms_" + key] = round(ms, 3)
            if split is None:
                row["items_per_s"] = round(count / (ms * 1e-3))
        out["rows"].append(row)
    iss.close()
    del wire
    torch.cuda.empty_cache()
# configs[2] in small: Issuer::issue (item-major requests in, issuances out) and CredentialIssuance::verify of the result
if "--no-issuance" not in sys.argv:
    sp, ip, sk = load_issuer("issuer4.bin")
    iss = Issuer(sp, ip, sk, device=0, max_batch=8192)
    kinds, n = bytes([0, 0, 2, 2]), 4
    rng = np.random.default_rng(7)
    pts, _ = iss.selftest_primitive("from_uniform", rng.integers(0, 256, (2 * 8192, 64), dtype=np.uint8))
    req = np.empty((8192, 3 * n + 14, 32), np.uint8)
    sc = rng.integers(0, 256, (8192, 2, 32), dtype=np.uint8); sc[:, :, 31] &= 0x0f
    req[:, 0:2] = sc; req[:, 2:4] = pts.reshape(8192, 2, 32)
    req[:, 4:] = rng.integers(0, 256, (8192, 3 * n + 10, 32), dtype=np.uint8)
    for count in (1, 1024, 8192):
        r = np.ascontiguousarray(req[:count])
        issued, st = iss.issue_wire(kinds, r)
        assert not st.any() and not iss.verify_wire(kinds, issued, issuance=True).any()
        ti, tv = [], []
        for _ in range(9):
            t0 = time.perf_counter(); iss.issue_wire(kinds, r); ti.append(time.perf_counter() - t0)
            t0 = time.perf_counter(); iss.verify_wire(kinds, issued, issuance=True); tv.append(time.perf_counter() - t0)
        out["rows"].append({"shape": "issuance n=4", "items": count, "ms_issue_wire": round(1e3 * float(np.median(ti)), 3),
                            "ms_verify_issuances_wire": round(1e3 * float(np.median(tv)), 3)})
    iss.close()
print(json.dumps(out, indent=1))
