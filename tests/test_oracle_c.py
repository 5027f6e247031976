"""The C oracle (reference CPU schedule, oracle/c) against the committed golden vectors and the Python oracle."""
import ctypes
import glob
import hashlib
import json
import os

import numpy as np
import pytest

from oracle.pyoracle import aeonflux as A, ristretto as R

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _words(hexes):
    return np.frombuffer(b"".join(bytes.fromhex(h) for h in hexes), np.uint8).reshape(len(hexes), 32).copy()


def test_c_primitives(coracle):
    L = coracle.lib()
    out = (ctypes.c_uint8 * 32)()
    o64 = (ctypes.c_uint8 * 64)()
    L.afxo_merlin_kat(out)
    assert bytes(out).hex() == "d5a21972d0d5fe320c0d263fac7fffb8145aa640af6e9bca177c03c7efcf0615"
    for m in (b"abc", b"x" * 200, b"y" * 112, b"z" * 111, b""):
        L.afxo_sha512(m, len(m), o64)
        assert bytes(o64) == hashlib.sha512(m).digest()
        L.afxo_shake256(m, len(m), o64, 64)
        assert bytes(o64) == hashlib.shake_256(m).digest(64)
    g = json.load(open(os.path.join(GOLD, "prims.json")))
    for e in g["decompress"]:
        b = bytes.fromhex(e["in"])
        assert bool(L.afxo_decompress_compress(b, out)) == e["valid"]
        if e["valid"]:
            assert bytes(out) == b
    for e in g["from_uniform"]:
        L.afxo_from_uniform(bytes.fromhex(e["in"]), out)
        assert bytes(out).hex() == e["out"]
    for e in g["scalarmult"]:
        for vt in (0, 1):
            assert L.afxo_scalarmult(bytes.fromhex(e["s"]), bytes.fromhex(e["p"]), out, vt) == 1
            assert bytes(out).hex() == e["out"]
    for e in g["wide_reduce"]:
        L.afxo_sc_from_wide(bytes.fromhex(e["in"]), out)
        assert bytes(out).hex() == e["out"]
    rfc = json.load(open(os.path.join(GOLD, "rfc9496.json")))          # RFC 9496 Appendix A in full
    B = bytes.fromhex(rfc["multiples"][1])
    for k, h in enumerate(rfc["multiples"]):
        for vt in (0, 1):
            assert L.afxo_scalarmult(k.to_bytes(32, "little"), B, out, vt) == 1 and bytes(out).hex() == h
        assert L.afxo_decompress_compress(bytes.fromhex(h), out) == 1 and bytes(out).hex() == h
    for group, encs in rfc["invalid"].items():
        for h in encs:
            assert L.afxo_decompress_compress(bytes.fromhex(h), out) == 0, (group, h)
    for e in rfc["hash_to_group"]:
        L.afxo_from_uniform(bytes.fromhex(e["input"]), out)
        assert bytes(out).hex() == e["output"]
    rng = A.ShakeRng(b"c-sc")
    for _ in range(200):
        a, b, c = rng.scalar(), rng.scalar(), rng.scalar()
        L.afxo_sc_muladd(R.sc_to_bytes(a), R.sc_to_bytes(b), R.sc_to_bytes(c), out)
        assert int.from_bytes(bytes(out), "little") == (a * b + c) % R.L
    for edge in (b"\xff" * 64, bytes(64), (R.L - 1).to_bytes(64, "little"), R.L.to_bytes(64, "little"), (R.L * R.L).to_bytes(64, "little")):
        L.afxo_sc_from_wide(edge, out)
        assert bytes(out) == R.sc_to_bytes(R.sc_from_wide(edge))


@pytest.mark.parametrize("name", ["readme4", "s16", "revealed10", "plain1_hidden", "scalar1", "quirk_sp_first", "quirk_sp_middle"])
def test_c_oracle_matches_golden(coracle, name):
    g = json.load(open(os.path.join(GOLD, name + ".json")))
    sp, ip, sk = coracle.make_issuer(g["n"])
    assert sp.hex() == g["sysparams"] and ip.hex() == g["issuer_pub"] and sk.hex() == g["secret"]
    iss = coracle.Issuer(sp, ip, sk)
    rk = bytes({"PS": ord("S"), "PP": ord("P"), "EP": ord("E")}[k] for k in g["request"])
    for e in g["items"]:
        kinds, pres, issu = iss.synth(rk, g["hide"], g["config"].encode(), e["item"], 1, threads=1)
        assert list(kinds) == e["kinds"]
        assert pres[0].tobytes().hex() == "".join(e["words"])          # generator parity (prover side)
        assert issu[0].tobytes().hex() == "".join(e["issuance_words"])
        v, _, tr = iss.verify_presentations(kinds, pres, threads=1, trace=True)
        assert v[0] == e["verdict"]
        assert tr["Z"][0].tobytes().hex() == e["Z"]
        ncm = len(e["commitments"])
        assert tr["commitments"][0, :ncm].tobytes().hex() == "".join(e["commitments"])
        assert tr["challenges"][0, :len(e["challenges"])].tobytes().hex() == "".join(e["challenges"])
        vi, _, tri = iss.verify_issuances(bytes(e["issuance_kinds"]), issu, threads=1, trace=True)
        assert vi[0] == e["issuance_verdict"] == 0
        assert tri["commitments"][0].tobytes().hex() == "".join(e["issuance_commitments"])
        for c in e["corrupted"]:
            w = _words(c["words"])[None]
            v, _, tr = iss.verify_presentations(kinds, w, threads=1, trace=True)
            assert v[0] == c["verdict"] == 1, c["class"]
            if c["Z"]:
                assert tr["Z"][0].tobytes().hex() == c["Z"]
            ncm = len(c["commitments"])
            assert tr["commitments"][0, :ncm].tobytes().hex() == "".join(c["commitments"]), c["class"]


def test_c_issue_with_supplied_randomness_verifies(coracle):
    sp, ip, sk = coracle.make_issuer(4)
    iss = coracle.Issuer(sp, ip, sk)
    kinds, pres, issu = iss.synth(b"SSPE", [0, 3], b"readme4", 100, 16)
    attrs = issu[:, :4, :].copy()
    rnd = np.frombuffer(hashlib.shake_256(b"issue-rand").digest(16 * 11 * 64), np.uint8).reshape(16, 11, 64).copy()
    out, status, _ = iss.issue(bytes([0, 0, 2, 2]), attrs, rnd)
    assert not status.any()
    items = np.concatenate([attrs, out], axis=1)
    v, _ = iss.verify_issuances(bytes([0, 0, 2, 2]), np.ascontiguousarray(items))
    assert not v.any()
    # user-side verification needs no secret key
    user = coracle.Issuer(sp, ip, None)
    v, _ = user.verify_issuances(bytes([0, 0, 2, 2]), np.ascontiguousarray(items))
    assert not v.any()
    items[3, 5, 0] ^= 1  # flip a bit of U
    v, _ = user.verify_issuances(bytes([0, 0, 2, 2]), np.ascontiguousarray(items))
    assert v[3] == 1 and v.sum() == 1


def test_c_oracle_batch_threads_agree(coracle):
    sp, ip, sk = coracle.make_issuer(4)
    iss = coracle.Issuer(sp, ip, sk)
    kinds, pres, _ = iss.synth(b"SSPE", [0, 3], b"readme4", 0, 64, want_issuances=False)
    pres[5, 2, 0] ^= 1
    pres[17, 10, 3] ^= 4
    v1, _ = iss.verify_presentations(kinds, pres, threads=1)
    v8, _ = iss.verify_presentations(kinds, pres, threads=8)
    assert (v1 == v8).all() and v1[5] == 1 and v1[17] == 1 and v1.sum() == 2
