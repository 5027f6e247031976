"""ORACLE (test infrastructure, NOT the product): the flat SoA wire layouts of the
batch boundary (SURVEY 8b / include/aeonflux_b200.h) <-> the in-memory structures
the reference holds, plus `verify_flat`, the oracle's answer for one flat item.

Flat-wire rule for byte strings the Rust types could never hold: an
undecodable point or a non-canonical scalar => verdict 1 (VerificationFailure), the
same outcome `decompress() -> None` / `from_canonical_bytes -> None` produce upstream.
"""
from . import ristretto as R
from .aeonflux import (Amac, Presentation, ProofOfEncryption, issuance_verify,
                       presentation_verify)
from .zkp import VerificationFailure

KIND_PS, KIND_SS, KIND_PP, KIND_SP = 0, 1, 2, 3
_KNAME = {0: "PS", 1: "SS", 2: "PP", 3: "SP"}
_KCODE = {v: k for k, v in _KNAME.items()}
ENC_WORDS = 14  # challenge, 6 responses, pk, E1, E2, C_y_1, C_y_2, C_y_3, C_y_2'


def presentation_kinds(p: Presentation):
    return bytes(_KCODE[k] for k, _ in p.encrypted_attributes)


def presentation_num_words(kinds) -> int:
    n = len(kinds)
    h_s = sum(1 for k in kinds if k == KIND_SS)
    r = sum(1 for k in kinds if k in (KIND_PS, KIND_PP))
    h_p = sum(1 for k in kinds if k == KIND_SP)
    return 1 + (3 + h_s) + 3 + n + r + ENC_WORDS * h_p


def presentation_to_words(p: Presentation):
    """Field order: challenge, responses[3+h_s], C_x_0, C_x_1, C_V, C_y[n],
    revealed[i] for each PS/PP in index order, then per SP in index order: enc_challenge,
    enc_responses[6], pk, E1, E2, C_y_1, C_y_2, C_y_3, C_y_2'."""
    c, resp = p.proof
    w = [R.sc_to_bytes(c)] + [R.sc_to_bytes(s) for s in resp]
    w += [p.C_x_0.compress(), p.C_x_1.compress(), p.C_V.compress()]
    w += [q.compress() for q in p.C_y]
    for k, v in p.encrypted_attributes:
        if k == "PS":
            w.append(R.sc_to_bytes(v))
        elif k == "PP":
            w.append(v.compress())
    for _i, pe in p.proofs_of_encryption:
        ec, er = pe.proof
        w += [R.sc_to_bytes(ec)] + [R.sc_to_bytes(s) for s in er]
        w += [pe.pk.compress(), pe.E1.compress(), pe.E2.compress(), pe.C_y_1.compress(),
              pe.C_y_2.compress(), pe.C_y_3.compress(), pe.C_y_2_prime.compress()]
    return w


def _sc(b):
    s = R.sc_from_canonical(b)
    if s is None:
        raise VerificationFailure("non-canonical scalar")
    return s


def _pt(b):
    p = R.decompress(b)
    if p is None:
        raise VerificationFailure("undecodable point")
    return p


def presentation_num_constraints(kinds) -> int:
    """Constraints of the main proof: Z, C_x_1 and one per compacted index i whose kinds[i] is not SecretPoint (A.6.1)."""
    kinds = list(kinds)
    nsp = sum(1 for k in kinds if k != KIND_SP)
    return 2 + sum(1 for i in range(nsp) if kinds[i] != KIND_SP)


def batchable_num_words(kinds) -> int:
    h_p = sum(1 for k in kinds if k == KIND_SP)
    return presentation_num_words(kinds) - 1 + presentation_num_constraints(kinds) + 4 * h_p


def compact_to_batchable_words(kinds, words, commitments):
    """The BatchableProof form of a (valid) compact presentation: each challenge word is replaced by that proof's blinding
    commitments -- which for a valid proof are exactly what verify_compact recomputes (`commitments`, main proof first)."""
    kinds = list(kinds)
    nc = presentation_num_constraints(kinds)
    h_s = sum(1 for k in kinds if k == KIND_SS)
    r = sum(1 for k in kinds if k in (KIND_PS, KIND_PP))
    head = 1 + 3 + h_s + 3 + len(kinds) + r
    out = list(commitments[:nc]) + list(words[1:head])
    pos, cpos = head, nc
    for k in kinds:
        if k == KIND_SP:
            out += list(commitments[cpos:cpos + 5]) + list(words[pos + 1:pos + ENC_WORDS])
            pos += ENC_WORDS; cpos += 5
    return out


def words_to_presentation(kinds, words, batchable=False) -> Presentation:
    kinds = list(kinds)
    if len(words) != (batchable_num_words(kinds) if batchable else presentation_num_words(kinds)):
        raise ValueError("malformed: word count")
    it = iter(words)
    n = len(kinds)
    hidden = [i for i, k in enumerate(kinds) if k == KIND_SS]
    c = [next(it) for _ in range(presentation_num_constraints(kinds))] if batchable else _sc(next(it))
    resp = [_sc(next(it)) for _ in range(3 + len(hidden))]
    C_x_0, C_x_1, C_V = _pt(next(it)), _pt(next(it)), _pt(next(it))
    C_y = [_pt(next(it)) for _ in range(n)]
    enc_attrs = []
    for k in kinds:
        if k == KIND_PS:
            enc_attrs.append(("PS", _sc(next(it))))
        elif k == KIND_PP:
            enc_attrs.append(("PP", _pt(next(it))))
        else:
            enc_attrs.append((_KNAME[k], None))
    poes = []
    for i, k in enumerate(kinds):
        if k != KIND_SP:
            continue
        ec = [next(it) for _ in range(5)] if batchable else _sc(next(it))
        er = [_sc(next(it)) for _ in range(6)]
        pk, E1, E2, C1, C2, C3, C2p = [_pt(next(it)) for _ in range(7)]
        poes.append((i, ProofOfEncryption((ec, er), pk, E1, E2, i, C1, C2, C3, C2p)))
    return Presentation((c, resp), poes, enc_attrs, hidden, C_x_0, C_x_1, C_V, C_y)


def verify_flat(issuer, kinds, words, batchable=False, linked=False):
    """-> (verdict, trace).  verdict 0 = Ok, 1 = VerificationFailure.
    trace: {'Z': bytes, 'commitments': [bytes...] (main proof's then each enc proof's, in
    constraint order), 'challenges': [32-byte recomputed challenge per proof]} -- filled as far
    as the reference's control flow gets (an early `?` return leaves the rest absent)."""
    trace = {}
    out = {"Z": None, "commitments": [], "challenges": []}
    verdict = 0
    try:
        p = words_to_presentation(kinds, words, batchable)
        presentation_verify(p, issuer, trace, linked=linked)
    except VerificationFailure:
        verdict = 1
    out["Z"] = trace.get("Z")
    vfs = ([trace["verifier"]] if "verifier" in trace else []) + trace.get("enc_verifiers", [])
    for vf in vfs:
        out["commitments"] += vf.trace["commitments"]
        if vf.trace["challenge"] is not None:
            out["challenges"].append(R.sc_to_bytes(vf.trace["challenge"]))
    return verdict, out


# ---- issuance ---------------------------------------------------------------
# request kinds: 0 = scalar attribute (Attribute::PublicScalar), 2 = point attribute
# (Attribute::PublicPoint, or the M1 of an EitherPoint plaintext) -- amacs.rs:224-244.

def issuance_num_words(n) -> int:
    return n + 3 + 1 + (n + 5)


def issuance_to_words(attrs, amac, proof):
    """attrs[n] (scalar / point / plaintext M1), t, U, V, challenge, responses[n+5]."""
    w = []
    for k, v in attrs:
        if k in ("PS", "SS"):
            w.append(R.sc_to_bytes(v))
        elif k == "PP":
            w.append(v.compress())
        else:
            w.append(v.M1.compress())
    w += [R.sc_to_bytes(amac.t), amac.U.compress(), amac.V.compress()]
    c, resp = proof
    return w + [R.sc_to_bytes(c)] + [R.sc_to_bytes(s) for s in resp]


def request_kinds(attrs):
    return bytes(KIND_PS if k in ("PS", "SS") else KIND_PP for k, _ in attrs)


def verify_issuance_flat(sp, ip, kinds, words, batchable=False):
    """-> (verdict, trace) for CredentialIssuance::verify (issuer.rs:48-57).  batchable: the challenge word is replaced by the three
    blinding commitments (C_W, I, V)."""
    kinds = list(kinds)
    n = len(kinds)
    out = {"commitments": [], "challenges": []}
    trace = {}
    verdict = 0
    try:
        if len(words) != issuance_num_words(n) + (2 if batchable else 0):
            raise ValueError("malformed: word count")
        attrs = []
        for k, b in zip(kinds, words[:n]):
            attrs.append(("PS", _sc(b)) if k == KIND_PS else ("PP", _pt(b)))
        t, U, V = _sc(words[n]), _pt(words[n + 1]), _pt(words[n + 2])
        c = list(words[n + 3:n + 6]) if batchable else _sc(words[n + 3])
        resp = [_sc(b) for b in words[n + (6 if batchable else 4):]]
        issuance_verify((c, resp), sp, ip, Amac(t, U, V), attrs, trace, batchable=batchable)
    except VerificationFailure:
        verdict = 1
    if "verifier" in trace:
        vf = trace["verifier"]
        out["commitments"] = vf.trace["commitments"]
        if vf.trace["challenge"] is not None:
            out["challenges"].append(R.sc_to_bytes(vf.trace["challenge"]))
    return verdict, out
