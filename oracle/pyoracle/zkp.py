"""ORACLE (test infrastructure, NOT the product): zkp 0.7 toolbox semantics
(SchnorrCS Prover / Verifier with CompactProof), restated from SURVEY.md Appendix
A.4 (zkp "0.7" is an un-vendored dependency, /root/reference/Cargo.toml:40).

Reference call sites: presentation.rs:187-284,355-435; encryption.rs:81-130,160-209;
issuance.rs:48-128,142-217.

Deviation that is by construction ("Issuance parity"): zkp's prover
draws blindings from a thread_rng()-seeded TranscriptRng, so no implementation can
reproduce a reference *proof* byte-for-byte.  Here the blindings are an explicit
argument; everything downstream of them is deterministic and bit-exact.
"""
from .merlin import Transcript
from .ristretto import (IDENTITY_COMPRESSED, L, Point, decompress, sc_from_wide)


class VerificationFailure(Exception):
    pass


def domain_sep(t: Transcript, label: bytes):
    t.append_message(b"dom-sep", b"schnorrzkp/1.0/ristretto255")
    t.append_message(b"dom-sep", label)


def get_challenge(t: Transcript, label: bytes) -> int:
    return sc_from_wide(t.challenge_bytes(label, 64))


class Verifier:
    def __init__(self, label: bytes, transcript: Transcript):
        self.t = transcript
        domain_sep(self.t, label)
        self.num_scalars = 0
        self.points = []        # compressed encodings
        self.point_labels = []
        self.constraints = []
        self.trace = {"commitments": [], "challenge": None}

    def allocate_scalar(self, label: bytes) -> int:
        self.t.append_message(b"scvar", label)
        self.num_scalars += 1
        return self.num_scalars - 1

    def allocate_point(self, label: bytes, enc: bytes) -> int:
        if enc == IDENTITY_COMPRESSED:
            raise VerificationFailure("identity point variable " + label.decode())
        self.t.append_message(b"ptvar", label)
        self.t.append_message(b"val", enc)
        self.points.append(enc)
        self.point_labels.append(label)
        return len(self.points) - 1

    def constrain(self, lhs: int, lc):
        self.constraints.append((lhs, list(lc)))

    def verify_batchable(self, commitments, responses):
        """zkp 0.7 Verifier::verify_batchable (BatchableProof = commitments + responses; the encoding the reference's commented-out
        BatchVerifier would need, presentation.rs:33-34).  The prover's commitments are validated (identity rejected:
        validate_and_append_blinding_commitment) and fed to the transcript, the challenge is derived from it, and every
        constraint must satisfy  sum resp_k * P_k - c * LHS - R == 0.  zkp checks a random 128-bit linear combination of
        these equations; this restatement checks each one exactly (same verdict up to probability 2^-128)."""
        if len(responses) != self.num_scalars or len(commitments) != len(self.constraints):
            raise VerificationFailure("response / commitment count")
        for (lhs, _lc), enc in zip(self.constraints, commitments):
            if enc == IDENTITY_COMPRESSED:
                raise VerificationFailure("identity blinding commitment")
            self.t.append_message(b"blindcom", self.point_labels[lhs])
            self.t.append_message(b"val", enc)
        c = get_challenge(self.t, b"chal")
        self.trace["challenge"] = c
        pts = [decompress(e) for e in self.points]
        Rs = [decompress(e) for e in commitments]
        if any(p is None for p in pts) or any(r is None for r in Rs):
            raise VerificationFailure("undecodable point")
        minus_c = (-c) % L
        for (lhs, lc), Rw in zip(self.constraints, Rs):
            acc = Point.identity()
            for sv, pv in lc:
                acc = acc + pts[pv] * responses[sv]
            acc = acc + pts[lhs] * minus_c
            self.trace["commitments"].append(acc.compress())
            if not (acc == Rw):
                raise VerificationFailure("constraint does not hold")

    def verify_compact(self, challenge, responses):
        if isinstance(challenge, (list, tuple)):       # a BatchableProof travels through the same call sites as (commitments, responses)
            return self.verify_batchable(list(challenge), responses)
        if len(responses) != self.num_scalars:
            raise VerificationFailure("response count")
        pts = [decompress(e) for e in self.points]
        if any(p is None for p in pts):
            raise VerificationFailure("undecodable point")
        minus_c = (-challenge) % L
        for lhs, lc in self.constraints:
            R = Point.identity()
            for sv, pv in lc:
                R = R + pts[pv] * responses[sv]
            R = R + pts[lhs] * minus_c
            enc = R.compress()
            self.trace["commitments"].append(enc)
            self.t.append_message(b"blindcom", self.point_labels[lhs])
            self.t.append_message(b"val", enc)
        c2 = get_challenge(self.t, b"chal")
        self.trace["challenge"] = c2
        if c2 != challenge:
            raise VerificationFailure("challenge mismatch")


class Prover:
    def __init__(self, label: bytes, transcript: Transcript):
        self.t = transcript
        domain_sep(self.t, label)
        self.scalars = []
        self.points = []
        self.point_labels = []
        self.constraints = []

    def allocate_scalar(self, label: bytes, value: int) -> int:
        self.t.append_message(b"scvar", label)
        self.scalars.append(value % L)
        return len(self.scalars) - 1

    def allocate_point(self, label: bytes, value: Point):
        enc = value.compress()
        self.t.append_message(b"ptvar", label)
        self.t.append_message(b"val", enc)
        self.points.append(value)
        self.point_labels.append(label)
        return len(self.points) - 1, enc

    def constrain(self, lhs: int, lc):
        self.constraints.append((lhs, list(lc)))

    def prove_compact(self, blindings):
        """blindings: one scalar per allocated scalar (supplied; see module docstring)."""
        assert len(blindings) == len(self.scalars)
        commitments = []
        for lhs, lc in self.constraints:
            R = Point.identity()
            for sv, pv in lc:
                R = R + self.points[pv] * blindings[sv]
            enc = R.compress()
            commitments.append(enc)
            self.t.append_message(b"blindcom", self.point_labels[lhs])
            self.t.append_message(b"val", enc)
        c = get_challenge(self.t, b"chal")
        responses = [(s * c + b) % L for s, b in zip(self.scalars, blindings)]
        return c, responses, commitments
