#!/usr/bin/env python
"""Small run of every batch operation for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python tests/tools/sanitize_small.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from aeonflux_b200 import Issuer, PresentationBatch, RequestBatch  # noqa: E402
from oracle import coracle as C  # noqa: E402

for n, rk, hide, count in ((4, b"SSPE", [0, 3], 300), (3, b"SES", [1], 70), (16, b"SSSSSSPP" + b"E" * 8, [0, 1] + list(range(8, 16)), 40)):
    sp, ip, sk = C.make_issuer(n)
    orc = C.Issuer(sp, ip, sk)
    kinds, pres, issu, showin = orc.synth(rk, hide, b"sanitize", 0, count, want_show_inputs=True)
    iss = Issuer(sp, ip, sk, device=0, max_batch=128)
    v = iss.verify_batch(PresentationBatch.from_items(kinds, pres))
    ov, _ = orc.verify_presentations(kinds, pres)
    assert (v == ov).all()
    assert (iss.verify_wire(kinds, pres) == ov).all()
    res, st = iss.show_batch(kinds, np.ascontiguousarray(showin.transpose(1, 0, 2)))
    assert (res.fields.transpose(1, 0, 2) == pres).all()
    ik = bytes(0 if c == ord("S") else 2 for c in rk)
    assert not iss.verify_issuance_batch(PresentationBatch.from_items(ik, issu)).any()
    R = np.random.default_rng(1).integers(0, 256, (count, n + 7, 64), dtype=np.uint8)
    out, status, _ = orc.issue(ik, np.ascontiguousarray(issu[:, :n]), R)
    ires, ist = iss.issue_batch(RequestBatch.from_request(ik, np.ascontiguousarray(issu[:, :n]), R))
    assert (ires.fields.transpose(1, 0, 2)[:, n:] == out).all()
    # BatchableProof mode: exact and random-linear-combination paths (incl. a fallback chunk)
    from aeonflux_b200 import compact_to_batchable
    comp = PresentationBatch.from_items(kinds, pres)
    _, dbg = iss.verify_batch(comp, debug=True)
    bf = compact_to_batchable(kinds, comp.fields, dbg["commitments"])
    assert not iss.verify_batchable(PresentationBatch(kinds, bf)).any()
    vr, fb = iss.verify_batchable_rlc(PresentationBatch(kinds, bf), bytes(range(32)))
    assert not vr.any() and fb == 0
    bf[1, 3, 2] ^= 1
    vr, fb = iss.verify_batchable_rlc(PresentationBatch(kinds, bf), bytes(range(32)))
    assert vr[3] == 1 and vr.sum() == 1 and fb == 1
    iss.close()
    print("ok", n, rk)
