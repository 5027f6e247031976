"""Replay of the reference-vector dumper (oracle/_ref_recipe/append_presentation.rs) with the Python oracle.

TEST INFRASTRUCTURE.  The dumper runs the reference crate's own flow under a deterministic byte-stream rng and writes
tests/golden/ref_<case>.json.  This module performs the same flow -- same seed, same order of rng draws -- with the oracle's
restatement, so that every word the stream determines (parameters, keys, attributes, t, U, V, ciphertexts and every commitment
point of a presentation) can be compared byte for byte with what the reference produced.  The challenge and response words are
NOT determined by the stream: zkp's prove_compact finalises its TranscriptRng with rand::thread_rng() (zkp 0.7 prover.rs), so
they are compared through the verifier instead (a reference-made proof is accepted iff every commitment is recomputed to the
byte, since the challenge is a hash of all of them).
"""
import hashlib

from . import aeonflux as A, flat as F, ristretto as R

SEED_PREFIX = b"aeonflux-b200/reference-vectors/"

# (name, n, request, hide, items) -- the dump_case() calls of the `b200_vectors` test, in order
CASES = [
    ("readme4", 4, ("PS", "PS", "PP", "EP"), (0, 3), 4),
    ("s16", 16, ("PS",) * 6 + ("PP", "PP") + ("EP",) * 8, (0, 1) + tuple(range(8, 16)), 2),
    ("revealed10", 10, ("PP", "PP", "PS", "PS", "PP", "PS", "PP", "PS", "PS", "PP"), (), 1),
    ("plain10_hidden_scalar", 10, ("EP", "PP", "PS", "PS", "PP", "PS", "PP", "PS", "PS", "PP"), (2,), 1),
    ("plain1_hidden", 1, ("EP",), (0,), 1),
    ("scalar1", 1, ("PS",), (), 1),
    ("quirk_sp_first", 3, ("EP", "PS", "PS"), (0,), 1),
    ("quirk_sp_middle", 3, ("PS", "EP", "PS"), (1,), 1),
    ("identity_plaintext", 6, ("EPZ", "PS", "PS", "PP", "PP", "PS"), (), 1),
]


class Sha512StreamRng(A.ShakeRng):
    """The dumper's StreamRng: block i = SHA-512(seed || le64(i)); every request takes the next bytes of the stream."""

    def __init__(self, seed: bytes):
        super().__init__(seed)
        self.taken = 0

    def fill(self, n: int) -> bytes:
        while len(self.buf) < n:
            self.buf += hashlib.sha512(self.seed + self.ctr.to_bytes(8, "little")).digest()
            self.ctr += 1
        out, self.buf = self.buf[:n], self.buf[n:]
        self.taken += n
        return out


def presentation_stream_words(kinds):
    """Indices of the presentation words that are NOT a function of the stream (challenge + responses of every proof)."""
    kinds = list(kinds)
    n = len(kinds)
    h_s = sum(k == F.KIND_SS for k in kinds)
    free = list(range(0, 4 + h_s))
    pos = 7 + h_s + n + sum(k in (F.KIND_PS, F.KIND_PP) for k in kinds)
    for k in kinds:
        if k == F.KIND_SP:
            free += list(range(pos, pos + 7))
            pos += 14
    return free


def issuance_stream_words(n):
    return list(range(n + 3, 2 * n + 9))


def replay_case(name, n, request, hide, items):
    """-> dict in the dumper's JSON structure (without the corruption lists), produced by the oracle.  The proofs' blindings come
    from a separate SHAKE stream (the reference's come from thread_rng), so challenge / response words differ from any reference
    run by construction; all other words must be identical."""
    seed = SEED_PREFIX + name.encode()
    rng = Sha512StreamRng(seed)
    sp = A.SystemParameters.generate(rng, n)
    issuer = A.Issuer.new(sp, rng)
    blind = A.ShakeRng(b"refvec-blindings/" + name.encode())
    out = {"source": "oracle/pyoracle replay (NOT the reference)", "name": name, "n": n, "seed": seed.hex(), "request": list(request),
           "hide": list(hide), "sysparams": sp.to_bytes().hex(), "issuer_pub": issuer.issuer_parameters.to_bytes().hex(),
           "secret": issuer.amacs_key.to_bytes().hex(), "items": []}
    for item in range(items):
        start = rng.taken
        attrs = []
        for r in request:
            if r == "PS":
                attrs.append(("PS", rng.scalar()))
            elif r == "PP":
                attrs.append(("PP", rng.point()))
            elif r == "EP":
                attrs.append(("EP", A.Plaintext.from_bytes30(rng.fill(30))))
            elif r == "EPZ":
                attrs.append(("EP", A.Plaintext.from_bytes30(bytes(30))))
            else:
                raise ValueError(r)
        # Issuer::issue: Amac::tag draws t then U from the caller's rng (amacs.rs:289-290); the proof's blindings do not
        proof, (amac, _) = issuer.issue(list(attrs), rng, blindings=[blind.scalar() for _ in range(n + 5)])
        iw = F.issuance_to_words(attrs, amac, proof)
        ik = F.request_kinds(attrs)
        iv, _ = F.verify_issuance_flat(sp, issuer.issuer_parameters, ik, iw)
        kp, _ = A.SymmetricKeypair.generate(sp, rng)
        shown = list(attrs)
        for i in hide:
            A.hide_attribute(shown, i)
        z = rng.scalar()                                         # presentation.rs:162
        h_s = sum(1 for k, _ in shown if k == "SS")
        h_p = sum(1 for k, _ in shown if k == "SP")
        pres = A.presentation_prove(sp, issuer.issuer_parameters, amac, shown, kp, z, [blind.scalar() for _ in range(3 + h_s)],
                                    [[blind.scalar() for _ in range(6)] for _ in range(h_p)])
        kinds = F.presentation_kinds(pres)
        words = F.presentation_to_words(pres)
        verdict, _ = F.verify_flat(issuer, kinds, words)
        out["items"].append({"item": item, "stream_start": start, "stream_end": rng.taken, "kinds": list(kinds), "words": [w.hex() for w in words],
                             "verdict": verdict, "corrupted": [], "issuance_kinds": list(ik), "issuance_words": [w.hex() for w in iw],
                             "issuance_verdict": iv, "issuance_corrupted": []})
    return out
