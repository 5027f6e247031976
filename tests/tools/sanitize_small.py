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
    os.environ["AFX_AMAC_SPLIT"] = "0"      # small batches cut the aMAC ladder into parts (k_ladders_parts); this is the single fused ladder with its ticket and flags
    assert (iss.verify_wire(kinds, pres) == ov).all()
    os.environ["AFX_AMAC_SPLIT"] = "1"
    assert (iss.verify_wire(kinds, pres) == ov).all()
    del os.environ["AFX_AMAC_SPLIT"]
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
    # round 2: item-major prover calls, BatchableProof issuances, RLC bisection, non-canonical wire scalars
    reqs = np.ascontiguousarray(RequestBatch.from_request(ik, np.ascontiguousarray(issu[:, :n]), R).fields.transpose(1, 0, 2))
    wout, wst = iss.issue_wire(ik, reqs)
    assert (wout[:, n:] == out).all() and not iss.verify_wire(ik, wout, issuance=True).any()
    wp, wps = iss.show_wire(kinds, showin)
    assert (wp == pres).all()
    oi, _, tri = orc.verify_issuances(ik, issu, trace=True)
    bi = np.ascontiguousarray(np.concatenate([issu[:, :n + 3], tri["commitments"], issu[:, n + 4:]], axis=1))
    assert not iss.verify_issuance_batchable(PresentationBatch.from_items(ik, bi)).any()
    bi[2, n + 3, 1] ^= 4
    vr, fb = iss.verify_batchable_rlc(PresentationBatch.from_items(ik, bi), bytes(32), issuance=True)
    assert vr[2] == 1 and vr.sum() == 1
    os.environ["AFX_RLC_LEAF"] = "8"
    vr, fb = iss.verify_batchable_rlc(PresentationBatch(kinds, bf), bytes(32))
    assert vr[3] == 1 and vr.sum() == 1
    del os.environ["AFX_RLC_LEAF"]
    bad = pres[:16].copy()
    for j in range(16):
        bad[j, 1 + j % 3, 28:] = np.frombuffer((0xffff0000 + j).to_bytes(4, "little"), np.uint8)      # responses far above l
    assert iss.verify_wire(kinds, bad).all()
    iss.close()
    print("ok", n, rk)

# the mixed-shape stream object and the multi-device handle (two contexts on device 0)
from aeonflux_b200.shard import MixedStream, MultiGpuIssuer, interleave_records  # noqa: E402
sp, ip, sk = C.make_issuer(4)
orc = C.Issuer(sp, ip, sk)
k4, p4, i4 = orc.synth(b"SSPE", [0, 3], b"sanitize-stream", 0, 200)
k2, p2, _ = orc.synth(b"SSPP", [], b"sanitize-stream2", 0, 90, want_issuances=False)
p4[5, 2, 31] ^= 0x80; p2[7, 1, 0] ^= 1
order = np.random.default_rng(2).permutation(np.concatenate([np.zeros(200, np.uint8), np.ones(90, np.uint8)]))
blob, offsets = interleave_records([p4, p2], order)
a, b = Issuer(sp, ip, sk, device=0, max_batch=64), Issuer(sp, ip, sk, device=0, max_batch=32)
st = MixedStream()
st.add_shape(a, k4); st.add_shape(b, k2)
v = np.full(290, 9, np.uint8)
st.push(blob, offsets, order, v); st.flush()
e4, _ = orc.verify_presentations(k4, p4); e2, _ = orc.verify_presentations(k2, p2)
exp = np.empty(290, np.uint8); exp[order == 0] = e4; exp[order == 1] = e2
assert (v == exp).all() and exp.sum() == 2
st.close(); a.close(); b.close()
m = MultiGpuIssuer(sp, ip, sk, devices=[0, 0], max_batch=64)
assert (m.verify_wire(k4, p4) == e4).all()
m.close()
print("ok stream + multi")
