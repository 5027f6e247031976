#!/usr/bin/env python
"""Runs one secondary workload on cuda:0 (device-resident, a few passes) so that `ncu -k regex:... -s N -c M` can capture its kernels:
    python tools/profile_secondary.py s16     Issuer::verify of 65,536 distinct S16 presentations (k_points, k_ladders)
    python tools/profile_secondary.py issue   Issuer::issue of 65,536 4-attribute requests (k_points, k_msm_ct), then
                                              CredentialIssuance::verify of the results (k_points, k_ladders)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    from aeonflux_b200 import Issuer
    what = sys.argv[1] if len(sys.argv) > 1 else "s16"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
    stream = torch.cuda.current_stream()
    if what == "s16":
        sp, ip, sk = bench.load_issuer("issuer16.bin")
        issuer = Issuer(sp, ip, sk, device=0, max_batch=B)
        f = bench.synthesize_on_device(torch, issuer, B, 4242, stream, bench.KINDS_S16, "keypair16.bin")
        v = torch.empty(B, dtype=torch.uint8, device="cuda")
        for _ in range(3):
            issuer.verify_batch_device(bench.KINDS_S16, B, f.data_ptr(), v.data_ptr(), stream.cuda_stream)
        torch.cuda.synchronize()
        assert int(v.sum().item()) == 0
    else:
        sp, ip, sk, items = bench.load_fixture(B)
        issuer = Issuer(sp, ip, sk, device=0, max_batch=B)
        kinds, n = bytes([0, 0, 2, 2]), 4
        rng = np.random.default_rng(1)
        req = np.empty((3 * n + 14, B, 32), np.uint8)
        sc = rng.integers(0, 256, (2, B, 32), dtype=np.uint8); sc[:, :, 31] &= 0x0f
        req[0:2] = sc; req[2] = items[:B, 5]; req[3] = items[:B, 6]
        req[4:] = rng.integers(0, 256, (3 * n + 10, B, 32), dtype=np.uint8)
        req_dev = torch.from_numpy(req).cuda()
        iss = torch.empty((2 * n + 9, B, 32), dtype=torch.uint8, device="cuda")
        iss[:n] = req_dev[:n]
        st = torch.empty(B, dtype=torch.uint8, device="cuda")
        for _ in range(3):
            issuer.issue_batch_device(kinds, B, req_dev.data_ptr(), iss[n:].data_ptr(), st.data_ptr(), stream.cuda_stream)
        for _ in range(3):
            issuer.verify_issuance_batch_device(kinds, B, iss.data_ptr(), st.data_ptr(), stream.cuda_stream)
        torch.cuda.synchronize()
        assert int(st.sum().item()) == 0
    print("ok", what, B)


if __name__ == "__main__":
    main()
