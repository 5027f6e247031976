#!/usr/bin/env python
"""Runs the secondary workloads once each on cuda:0 (device-resident, after two warm-up passes) so that
`ncu -k regex:'k_msm_ct|k_ladders|k_points' ...` can capture their ladder kernels:
  Issuer::issue of 65,536 4-attribute requests (k_points, k_msm_ct), CredentialIssuance::verify of the results (k_ladders),
  Issuer::verify of 16,384 S16 presentations (k_points, k_ladders)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    from aeonflux_b200 import Issuer
    B = 65536
    sp, ip, sk, items = bench.load_fixture(B)
    issuer = Issuer(sp, ip, sk, device=0, max_batch=B)
    stream = torch.cuda.current_stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    out = bench.secondary_measurements(torch, issuer, items, 0, stream, flush, B, 1)
    for k, v in out.items():
        print(k, round(v["value"]), "per s", round(v["ms_per_step"], 2), "ms")


if __name__ == "__main__":
    main()
