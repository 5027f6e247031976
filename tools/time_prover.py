#!/usr/bin/env python
"""Device-resident timing of the prover paths on cuda:0 (CUDA events, L2 flush between passes): Issuer::issue of 65,536 revealed
4-attribute requests, the CredentialIssuance::verify of what it made (parity guard) and AnonymousCredential::show via the bench's
device-side synthesis.  Prints one JSON line; used to compare library variants of k_msm_ct (tools/gpu_variants_ct.sh)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    from aeonflux_b200 import Issuer
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    steps = 4
    stream = torch.cuda.current_stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    sp, ip, sk, items = bench.load_fixture(B)
    issuer = Issuer(sp, ip, sk, device=0, max_batch=B)
    kinds, n = bytes([0, 0, 2, 2]), 4
    rng = np.random.default_rng(1234)
    req = np.empty((3 * n + 14, B, 32), np.uint8)
    sc = rng.integers(0, 256, (2, B, 32), dtype=np.uint8); sc[:, :, 31] &= 0x0f
    req[0:2] = sc; req[2] = items[:B, 5]; req[3] = items[:B, 6]
    req[4:] = rng.integers(0, 256, (3 * n + 10, B, 32), dtype=np.uint8)
    req_dev = torch.from_numpy(req).cuda()
    iss = torch.empty((2 * n + 9, B, 32), dtype=torch.uint8, device="cuda")
    iss[:n] = req_dev[:n]
    st = torch.empty(B, dtype=torch.uint8, device="cuda")
    issuer.set_stage_timing(True)
    ms = bench.time_device(torch, stream, flush, lambda: issuer.issue_batch_device(kinds, B, req_dev.data_ptr(), iss[n:].data_ptr(), st.data_ptr(), stream.cuda_stream), steps)
    stages = issuer.stage_times_ms()
    issuer.set_stage_timing(False)
    assert int(st.sum().item()) == 0
    issuer.verify_issuance_batch_device(kinds, B, iss.data_ptr(), st.data_ptr(), stream.cuda_stream)
    torch.cuda.synchronize()
    assert int(st.sum().item()) == 0, "issued credentials failed CredentialIssuance::verify"
    digest = int(iss.to(torch.int64).sum().item())          # same rng bytes -> same issuance bytes, whatever the variant
    out = {"issue_ms": ms, "issue_per_s": B / (ms * 1e-3), "issue_stage_ms": stages, "issue_digest": digest}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
