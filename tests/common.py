"""Shared helpers for the parity tests: golden fixtures, oracle batches, trace comparison."""
import json
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_SHAPES = ["readme4", "s16", "revealed10", "plain1_hidden", "scalar1", "quirk_sp_first", "quirk_sp_middle"]
REQ = {"PS": ord("S"), "PP": ord("P"), "EP": ord("E")}


def load_golden(name):
    return json.load(open(os.path.join(GOLD, name + ".json")))


def words(hexes):
    return np.frombuffer(b"".join(bytes.fromhex(h) for h in hexes), np.uint8).reshape(len(hexes), 32).copy()


def compare_with_oracle_trace(verdicts, dbg, overdicts, otrace):
    """Engine debug dump vs the oracle's trace.  The reference returns at the first failing `?`, so the oracle trace is
    only filled that far (all-zero rows beyond); everything it did compute must match bit for bit."""
    assert (np.asarray(verdicts) == np.asarray(overdicts)).all(), (verdicts, overdicts)
    count = len(verdicts)
    cm = dbg["commitments"].transpose(1, 0, 2)
    ch = dbg["challenges"].transpose(1, 0, 2)
    for i in range(count):
        if "Z" in otrace and otrace["Z"][i].any():
            assert (dbg["Z"][i] == otrace["Z"][i]).all(), ("Z", i)
        for k in range(otrace["commitments"].shape[1]):
            if otrace["commitments"][i, k].any():
                assert (cm[i, k] == otrace["commitments"][i, k]).all(), ("commitment", i, k)
        for k in range(otrace["challenges"].shape[1]):
            if otrace["challenges"][i, k].any():
                assert (ch[i, k] == otrace["challenges"][i, k]).all(), ("challenge", i, k)


CORRUPTIONS = ("response+1", "challenge+1", "C_x_0+B", "C_V+B", "revealed_scalar", "enc_E2+B", "enc_response+1",
               "undecodable_point", "noncanonical_scalar", "identity_point")


def corrupt_batch(pres, kinds, rng, fraction, oracle_points):
    """Corrupt ~fraction of the items of pres [count][W][32] in place with the SURVEY 8d classes (byte-level versions:
    adding B to a point is replaced by swapping in another valid point).  Returns the indices touched."""
    from oracle import coracle as C
    kinds = list(kinds)
    n = len(kinds)
    h_s = sum(k == C.KIND_SS for k in kinds)
    enc0 = 7 + h_s + n + sum(k in (C.KIND_PS, C.KIND_PP) for k in kinds)
    has_enc = any(k == C.KIND_SP for k in kinds)
    ps = [i for i, k in enumerate(kinds) if k == C.KIND_PS]
    count = pres.shape[0]
    L = 2**252 + 27742317777372353535851937790883648493
    idx = np.where(rng.random(count) < fraction)[0]
    for j, i in enumerate(idx):
        cls = CORRUPTIONS[j % len(CORRUPTIONS)]
        def sc_plus1(w):
            v = (int.from_bytes(pres[i, w].tobytes(), "little") + 1) % L
            pres[i, w] = np.frombuffer(v.to_bytes(32, "little"), np.uint8)
        other = oracle_points[rng.integers(len(oracle_points))]
        if cls == "response+1":
            sc_plus1(1)
        elif cls == "challenge+1":
            sc_plus1(0)
        elif cls == "C_x_0+B":
            pres[i, 4 + h_s] = other
        elif cls == "C_V+B":
            pres[i, 6 + h_s] = other
        elif cls == "revealed_scalar" and ps:
            rev = 7 + h_s + n + sum(1 for k in kinds[:ps[0]] if k in (C.KIND_PS, C.KIND_PP))
            sc_plus1(rev)
        elif cls == "enc_E2+B" and has_enc:
            pres[i, enc0 + 9] = other
        elif cls == "enc_response+1" and has_enc:
            sc_plus1(enc0 + 1)
        elif cls == "undecodable_point":
            pres[i, 5 + h_s] = np.frombuffer(bytes([3]) + bytes(31), np.uint8)   # odd => negative => invalid encoding
        elif cls == "noncanonical_scalar":
            v = int.from_bytes(pres[i, 2].tobytes(), "little") + L
            pres[i, 2] = np.frombuffer(v.to_bytes(32, "little"), np.uint8)
        elif cls == "identity_point":
            pres[i, 7 + h_s] = 0                                                   # C_y[0] := identity encoding
        else:
            sc_plus1(1)
    return idx


def golden_issue_request(g, entry):
    """Rebuild the CredentialRequest and the rng output behind a golden item's issuance: oracle/pyoracle/synth.py draws
    the attributes, then t, U and the n+5 blindings (64 bytes each) from SHAKE-256(prefix || config || item)."""
    from oracle.pyoracle import aeonflux as A, synth as S
    rng = S.item_rng(g["config"].encode(), entry["item"])
    for k in g["request"]:
        rng.fill(30 if k == "EP" else 64)
    n = g["n"]
    rnd = np.frombuffer(rng.fill(64 * (n + 7)), np.uint8).reshape(n + 7, 64).copy()
    attrs = words(entry["issuance_words"][:n])
    return attrs, rnd


def to_batchable(kinds, pres, commitments):
    """Compact presentations [count][W][32] + the commitments their (valid) proofs recompute [count][n_commit][32] -> the
    BatchableProof layout [count][Wb][32]: every challenge word replaced by that proof's commitments
    (oracle/pyoracle/flat.py:compact_to_batchable_words gives the word map)."""
    from oracle.pyoracle import flat as F
    W, nc = pres.shape[1], commitments.shape[1]
    order = F.compact_to_batchable_words(list(kinds), [("w", i) for i in range(W)], [("c", j) for j in range(nc)])
    out = np.empty((pres.shape[0], len(order), 32), np.uint8)
    for k, (src, idx) in enumerate(order):
        out[:, k] = pres[:, idx] if src == "w" else commitments[:, idx]
    return out


def presentation_scalar_words(kinds):
    """Indices of the words of the flat presentation layout (include/aeonflux_b200.h) that hold scalars."""
    kinds = list(kinds)
    n = len(kinds)
    h_s = sum(k == 1 for k in kinds)
    idx = list(range(0, 4 + h_s))                                  # challenge, responses[3 + h_s]
    pos = 7 + h_s + n
    for k in kinds:
        if k in (0, 2):
            if k == 0:
                idx.append(pos)                                    # a revealed scalar
            pos += 1
    for k in kinds:
        if k == 3:
            idx += list(range(pos, pos + 7))                       # enc_challenge, enc_responses[6]
            pos += 14
    return idx


def issuance_scalar_words(kinds):
    n = len(kinds)
    return [i for i, k in enumerate(kinds) if k == 0] + [n] + list(range(n + 3, 2 * n + 9))   # scalar attributes, t, challenge, responses


NONCANONICAL_TOP_WORDS = (0x10000001, 0x1fffffff, 0x20010000, 0x2001ffff, 0x3fff8000, 0x40000000, 0x7fff0000, 0x7fffffff, 0x80000000,
                          0x8000ffff, 0xc0007fff, 0xfffe8000, 0xffff0000, 0xffff7fff, 0xffff8000, 0xffffffff)


def noncanonical_scalar_batch(items, scalar_words, rng):
    """One copy of an honest item per (scalar word, top-word pattern): that word's top 32 bits are replaced by a pattern that
    makes the 256-bit value >= l (the unbiased top digit of the radix-2^16 / radix-4096 recodings would index far past a
    constant table), the low 224 bits are kept, zeroed or saturated.  Returns items [len(scalar_words) * 16][W][32]."""
    out = []
    for w in scalar_words:
        for j, top in enumerate(NONCANONICAL_TOP_WORDS):
            it = items[int(rng.integers(len(items)))].copy()
            if j % 3 == 1:
                it[w, :28] = 0
            elif j % 3 == 2:
                it[w, :28] = 0xff
            it[w, 28:] = np.frombuffer(int(top).to_bytes(4, "little"), np.uint8)
            out.append(it)
    return np.stack(out)


_L = 2**252 + 27742317777372353535851937790883648493
_P = 2**255 - 19


def _w(v):
    return np.frombuffer(int(v % 2**256).to_bytes(32, "little"), np.uint8)


# 32-byte values that sit on the edges of the wire rules (RFC 9496 decode: s < p, s non-negative, a square root exists, the result's
# T non-negative and Y non-zero; scalars: < l) -- each is a legal scalar or point encoding in some word and an illegal one in another
EDGE_WORDS = [_w(v) for v in (0, 1, 2, 3, 4, _L - 1, _L, _L + 1, 2 * _L, 8 * _L - 1, 2**252, 2**252 - 1, _P - 1, _P, _P + 1, _P + 2, 2**255 - 20,
                              2**255 - 1, 2**255, 2**255 + 1, 2**256 - 1, 2**256 - 38, (_P - 1) // 2, (_P + 1) // 2)] + [
    np.frombuffer(bytes.fromhex(h), np.uint8) for h in (
        "e2f2ae0a6abc4e71a884a961c500515f58e30b6aa582dd8db6a65945e08d2d76",      # the ristretto255 basepoint
        "6a493210f7499cd17fecb510ae0cea23a110e8d5b901f8acadd3095c73a3b919",      # 2B
        "ecffffffffffffffffffffffffffffffffffffffffffffffffffffffffffff7f",      # p - 1: s = -1, non-canonical sign
        "0100000000000000000000000000000000000000000000000000000000000080")]     # bit 255 set


def fuzz_mutations(items, rng, count, max_mut=3):
    """count mutated copies of honest items [m][W][32] for differential runs (engine vs oracle): 1..max_mut words of each copy are
    replaced by an edge value, random bytes, a word from another position of the same item (a valid encoding in the wrong
    place), the same word of another item, or get one bit flipped (any bit, the top byte's included)."""
    m, W, _ = items.shape
    out = np.empty((count, W, 32), np.uint8)
    for i in range(count):
        it = items[int(rng.integers(m))].copy()
        for _ in range(int(rng.integers(1, max_mut + 1))):
            w, how = int(rng.integers(W)), int(rng.integers(6))
            if how == 0:
                it[w] = EDGE_WORDS[int(rng.integers(len(EDGE_WORDS)))]
            elif how == 1:
                it[w] = rng.integers(0, 256, 32, dtype=np.uint8)
            elif how == 2:
                it[w] = it[int(rng.integers(W))]
            elif how == 3:
                it[w] = items[int(rng.integers(m)), w]
            elif how == 4:
                it[w, int(rng.integers(32))] ^= 1 << int(rng.integers(8))
            else:
                it[w, 31] ^= 1 << int(rng.integers(4, 8))          # the bits that decide s < p / scalar < l
        out[i] = it
    return out


def differential_fuzz(make_issuer, coracle, seed, count, shapes=((4, b"SSPE", [0, 3]), (3, b"SPE", [2])), with_trace=True):
    """Mutated presentations and issuances through Issuer::verify / CredentialIssuance::verify of the engine and of the C oracle:
    identical verdicts and (where the oracle's early return got that far) identical Z, commitments and challenges; then the same
    context still accepts the honest items.  Returns (accepted, rejected) over everything that was run."""
    from aeonflux_b200 import PresentationBatch
    rng = np.random.default_rng(seed)
    acc = rej = 0
    for n, rk, hide in shapes:
        sp, ip, sk = coracle.make_issuer(n)
        orc = coracle.Issuer(sp, ip, sk)
        kinds, pres, issu = orc.synth(rk, hide, b"fuzz-%d" % seed, 0, 16)
        iss = make_issuer(sp, ip, sk)
        bad = fuzz_mutations(pres, rng, count)
        ov, _, tr = orc.verify_presentations(kinds, bad, trace=True)
        if with_trace:
            v, dbg = iss.verify_batch(PresentationBatch.from_items(kinds, bad), debug=True)
            compare_with_oracle_trace(v, dbg, ov, tr)
        else:
            v = iss.verify_batch(PresentationBatch.from_items(kinds, bad))
            assert (v == ov).all(), np.where(v != ov)[0][:10]
        assert (iss.verify_wire(kinds, bad) == ov).all()
        acc += int((ov == 0).sum()); rej += int((ov == 1).sum())
        ik = bytes(0 if c == ord("S") else 2 for c in rk)
        ibad = fuzz_mutations(issu, rng, count)
        oi, _, tri = orc.verify_issuances(ik, ibad, trace=True)
        if with_trace:
            vi, dbgi = iss.verify_issuance_batch(PresentationBatch.from_items(ik, ibad), debug=True)
            compare_with_oracle_trace(vi, dbgi, oi, tri)
        else:
            vi = iss.verify_issuance_batch(PresentationBatch.from_items(ik, ibad))
            assert (vi == oi).all(), np.where(vi != oi)[0][:10]
        acc += int((oi == 0).sum()); rej += int((oi == 1).sum())
        assert not iss.verify_batch(PresentationBatch.from_items(kinds, pres)).any()
        assert not iss.verify_issuance_batch(PresentationBatch.from_items(ik, issu)).any()
        iss.close()
    return acc, rej
