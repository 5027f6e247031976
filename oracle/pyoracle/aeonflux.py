"""ORACLE (test infrastructure, NOT the product): line-by-line restatement of the
aeonflux protocol functions on and around the hot path.  Every function cites the
reference lines it follows (paths relative to /root/reference/src).

Parity status: the reference holds no golden vectors (all its tests draw from
thread_rng) and cannot be compiled here (no Rust toolchain), so this
oracle is pinned by external KATs for the third-party layers (see ristretto.py,
merlin.py) and by reproducing the verdicts of the reference's own tests
(tests/test_oracle_kats.py::test_reference_test_verdicts) -- "parity unpinned by the reference itself" at the byte
level, as DESIGN.md states.
"""
import hashlib

from . import ristretto as R
from .merlin import Transcript
from .ristretto import L, Point
from .zkp import Prover, VerificationFailure, Verifier


class MacCreationError(Exception):
    """CredentialError::MacCreation (errors.rs:141-142) from amacs.rs:285-287."""


# ---- deterministic randomness ----------------------------------

class ShakeRng:
    """Deterministic stand-in for the reference's CryptoRng: SHAKE-256(seed || le32(counter)),
    64-byte blocks."""

    def __init__(self, seed: bytes):
        self.seed = seed
        self.ctr = 0
        self.buf = b""

    def fill(self, n: int) -> bytes:
        while len(self.buf) < n:
            self.buf += hashlib.shake_256(self.seed + self.ctr.to_bytes(4, "little")).digest(64)
            self.ctr += 1
        out, self.buf = self.buf[:n], self.buf[n:]
        return out

    def scalar(self) -> int:
        """Scalar::random: 64 bytes, from_bytes_mod_order_wide."""
        return R.sc_from_wide(self.fill(64))

    def point(self) -> Point:
        """RistrettoPoint::random: 64 bytes, from_uniform_bytes."""
        return R.from_uniform_bytes(self.fill(64))


# ---- parameters.rs ----------------------------------------------------------

class SystemParameters:
    """parameters.rs:62-76."""

    def __init__(self, n, G, G_w, G_w_prime, G_x_0, G_x_1, G_y, G_m, G_V, G_a, G_a0, G_a1):
        self.n = n
        self.G, self.G_w, self.G_w_prime, self.G_x_0, self.G_x_1 = G, G_w, G_w_prime, G_x_0, G_x_1
        self.G_y, self.G_m, self.G_V, self.G_a, self.G_a0, self.G_a1 = G_y, G_m, G_V, G_a, G_a0, G_a1

    @staticmethod
    def generate(rng: ShakeRng, n: int):
        """hash_and_pray, parameters.rs:196-326 (sampling order :215-281)."""
        def sample():
            while True:
                p = R.decompress(rng.fill(32))
                if p is not None:
                    return p
        G_w = sample(); G_w_prime = sample(); G_x_0 = sample(); G_x_1 = sample()
        G_y = [sample() for _ in range(max(n, 3))]
        G_m = [sample() for _ in range(n)]
        G_V = sample(); G_a = sample(); G_a0 = sample(); G_a1 = sample()
        return SystemParameters(n, R.BASEPOINT, G_w, G_w_prime, G_x_0, G_x_1, G_y, G_m, G_V, G_a, G_a0, G_a1)

    def to_bytes(self) -> bytes:
        """parameters.rs:155-184."""
        v = self.n.to_bytes(4, "little")
        for p in [self.G, self.G_w, self.G_w_prime, self.G_x_0, self.G_x_1]:
            v += p.compress()
        for i in range(max(self.n, 3)):
            v += self.G_y[i].compress()
        for i in range(self.n):
            v += self.G_m[i].compress()
        for p in [self.G_V, self.G_a, self.G_a0, self.G_a1]:
            v += p.compress()
        return v

    @staticmethod
    def sizeof(n):
        """parameters.rs:34-40."""
        return 32 * (5 + 3 + n + 4) + 4 if n < 3 else 32 * (5 + 2 * n + 4) + 4

    @staticmethod
    def from_bytes(b: bytes):
        """parameters.rs:92-153."""
        n = int.from_bytes(b[:4], "little")
        if len(b) != SystemParameters.sizeof(n):
            raise ValueError("NoSystemParameters")
        pts = []
        for off in range(4, len(b), 32):
            p = R.decompress(b[off:off + 32])
            if p is None:
                raise ValueError("PointDecompressionError")
            pts.append(p)
        ny = max(n, 3)
        return SystemParameters(n, pts[0], pts[1], pts[2], pts[3], pts[4], pts[5:5 + ny],
                                pts[5 + ny:5 + ny + n], *pts[5 + ny + n:5 + ny + n + 4])


class SecretKey:
    """amacs.rs:52-59."""

    def __init__(self, w, w_prime, x_0, x_1, y, W):
        self.w, self.w_prime, self.x_0, self.x_1, self.y, self.W = w, w_prime, x_0, x_1, y, W

    @staticmethod
    def generate(rng: ShakeRng, sp: SystemParameters):
        """amacs.rs:89-107."""
        w = rng.scalar(); w_prime = rng.scalar(); x_0 = rng.scalar(); x_1 = rng.scalar()
        y = [rng.scalar() for _ in range(sp.n)]
        return SecretKey(w, w_prime, x_0, x_1, y, sp.G_w * w)

    def to_bytes(self) -> bytes:
        """amacs.rs:110-125 (the authoritative layout; from_bytes :128-155 has the y-loop bug)."""
        v = len(self.y).to_bytes(4, "little")
        for s in [self.w, self.w_prime, self.x_0, self.x_1] + list(self.y):
            v += R.sc_to_bytes(s)
        return v + self.W.compress()


class IssuerParameters:
    """parameters.rs:341-362."""

    def __init__(self, C_W, I):
        self.C_W, self.I = C_W, I

    @staticmethod
    def generate(sp: SystemParameters, sk: SecretKey):
        C_W = sp.G_w * sk.w + sp.G_w_prime * sk.w_prime
        I = sp.G_V - sp.G_x_0 * sk.x_0 - sp.G_x_1 * sk.x_1
        for i in range(sp.n):
            I = I - sp.G_y[i] * sk.y[i]
        return IssuerParameters(C_W, I)

    def to_bytes(self) -> bytes:
        """C_W || I, the 64 bytes issuer.rs:155,163 reserves (reference impl is unimplemented!())."""
        return self.C_W.compress() + self.I.compress()


# ---- encoding.rs / symmetric.rs (user side; generator only) -----------------

def encode_to_group(data: bytes):
    """encoding.rs:56-70."""
    assert len(data) <= 30
    b = bytearray(32)
    b[1:1 + len(data)] = data
    for j in range(64):
        b[31] = j
        for i in range(128):
            b[0] = 2 * i
            p = R.decompress(bytes(b))
            if p is not None:
                return p, i + j * 128
    raise RuntimeError("a very unlikely event occurred")


def decode_from_group(p: Point):
    """encoding.rs:75-82."""
    c = p.compress()
    return c[1:31], (c[0] // 2) + c[31] * 128


class Plaintext:
    """symmetric.rs:89-96, From<&[u8;30]> :135-143."""

    def __init__(self, M1, M2, m3):
        self.M1, self.M2, self.m3 = M1, M2, m3

    @staticmethod
    def from_bytes30(src: bytes):
        assert len(src) == 30
        M1, _ = encode_to_group(src)
        h = hashlib.sha512(src).digest()
        return Plaintext(M1, R.from_uniform_bytes(h), R.sc_from_wide(h))

    @staticmethod
    def from_slice(data: bytes):
        """symmetric.rs:118-132."""
        return [Plaintext.from_bytes30(data[i:i + 30].ljust(30, b"\0")) for i in range(0, len(data), 30)]


class SymmetricKeypair:
    """symmetric.rs:74-79."""

    def __init__(self, a, a0, a1, pk):
        self.a, self.a0, self.a1, self.pk = a, a0, a1, pk

    @staticmethod
    def derive(master_secret: bytes, sp: SystemParameters):
        """symmetric.rs:197-215."""
        a = R.sc_from_wide(hashlib.sha512(master_secret).digest())
        a0 = R.sc_from_wide(hashlib.sha512(R.sc_to_bytes(a)).digest())
        a1 = R.sc_from_wide(hashlib.sha512(R.sc_to_bytes(a0)).digest())
        return SymmetricKeypair(a, a0, a1, sp.G_a * a + sp.G_a0 * a0 + sp.G_a1 * a1)

    @staticmethod
    def generate(sp, rng: ShakeRng):
        """symmetric.rs:227-241."""
        ms = rng.fill(64)
        return SymmetricKeypair.derive(ms, sp), ms

    def encrypt(self, pt: Plaintext):
        """symmetric.rs:252-261."""
        E1 = pt.M2 * ((self.a0 + self.a1 * pt.m3) % L)
        E2 = E1 * self.a + pt.M1
        return E1, E2

    def decrypt(self, E1, E2):
        """symmetric.rs:273-289."""
        M1p = E2 - E1 * self.a
        m_prime, _ = decode_from_group(M1p)
        h = hashlib.sha512(m_prime).digest()
        m3p = R.sc_from_wide(h)
        M2p = R.from_uniform_bytes(h)
        if not (E1 == M2p * ((self.a0 + self.a1 * m3p) % L)):
            raise ValueError("UndecryptableAttribute")
        return Plaintext(M1p, M2p, m3p)


# ---- amacs.rs ----------------------------------------------------------------
# Attribute = (kind, value): kinds 'PS','SS' (value int), 'PP' (Point), 'EP','SP' (Plaintext); amacs.rs:168-179

def messages_from_attributes(attrs, sp):
    """Messages::from_attributes, amacs.rs:224-244."""
    out = []
    for i, (k, v) in enumerate(attrs):
        if k in ("PS", "SS"):
            out.append(sp.G_m[i] * v)
        elif k == "PP":
            out.append(v)
        else:
            out.append(v.M1)
    return out


class Amac:
    def __init__(self, t, U, V):
        self.t, self.U, self.V = t, U, V

    @staticmethod
    def compute_V(sp, sk, attrs, t, U):
        """amacs.rs:256-272."""
        M = messages_from_attributes(attrs, sp)
        V = sk.W + U * sk.x_0 + U * ((sk.x_1 * t) % L)
        for y, m in zip(sk.y, M):
            V = V + m * y
        return V

    @staticmethod
    def tag(rng, sp, sk, attrs, t=None, U=None):
        """amacs.rs:276-294.  t/U may be supplied instead of drawn (same order: t then U)."""
        if len(attrs) != sp.n:
            raise MacCreationError()
        if t is None:
            t = rng.scalar()
        if U is None:
            U = rng.point()
        return Amac(t, U, Amac.compute_V(sp, sk, attrs, t, U))


# ---- issuer.rs / nizk/issuance.rs ------------------------------------------

class Issuer:
    """issuer.rs:61-93."""

    def __init__(self, sp, ip, sk):
        self.system_parameters, self.issuer_parameters, self.amacs_key = sp, ip, sk

    @staticmethod
    def new(sp, rng):
        sk = SecretKey.generate(rng, sp)
        return Issuer(sp, IssuerParameters.generate(sp, sk), sk)

    def issue(self, attrs, rng, blindings=None, t=None, U=None):
        """issuer.rs:111-124.  Returns (proof=(c,responses), credential=(amac, attrs))."""
        amac = Amac.tag(rng, self.system_parameters, self.amacs_key, attrs, t, U)
        n = self.system_parameters.n
        if blindings is None:
            blindings = [rng.scalar() for _ in range(n + 5)]
        proof, _ = issuance_prove(self, amac, attrs, blindings)
        return proof, (amac, attrs)

    def verify(self, presentation):
        """issuer.rs:141-147."""
        return presentation_verify(presentation, self)


def issuance_prove(issuer, amac, attrs, blindings):
    """ProofOfIssuance::prove, issuance.rs:40-129."""
    sp, ip, sk = issuer.system_parameters, issuer.issuer_parameters, issuer.amacs_key
    t = Transcript(b"2019/1416 anonymous credential")
    pr = Prover(b"2019/1416 issuance proof", t)
    w = pr.allocate_scalar(b"w", sk.w)
    w_prime = pr.allocate_scalar(b"w'", sk.w_prime)
    x_0 = pr.allocate_scalar(b"x_0", sk.x_0)
    x_1 = pr.allocate_scalar(b"x_1", sk.x_1)
    y = [pr.allocate_scalar(b"y", yi) for yi in sk.y]
    one = pr.allocate_scalar(b"1", 1)
    G_V, _ = pr.allocate_point(b"G_V", sp.G_V)
    G_w, _ = pr.allocate_point(b"G_w", sp.G_w)
    G_w_prime, _ = pr.allocate_point(b"G_w_prime", sp.G_w_prime)
    neg_G_x_0, _ = pr.allocate_point(b"-G_x_0", -sp.G_x_0)
    neg_G_x_1, _ = pr.allocate_point(b"-G_x_1", -sp.G_x_1)
    neg_G_y = [pr.allocate_point(b"-G_y", -g)[0] for g in sp.G_y]
    C_W, _ = pr.allocate_point(b"C_W", ip.C_W)
    I, _ = pr.allocate_point(b"I", ip.I)
    U, _ = pr.allocate_point(b"U", amac.U)
    V, _ = pr.allocate_point(b"V", amac.V)
    tU, _ = pr.allocate_point(b"tU", amac.U * amac.t)
    M = [pr.allocate_point(b"M", m)[0] for m in messages_from_attributes(attrs, sp)]
    pr.constrain(C_W, [(w, G_w), (w_prime, G_w_prime)])
    pr.constrain(I, [(one, G_V), (x_0, neg_G_x_0), (x_1, neg_G_x_1)] + list(zip(y, neg_G_y)))
    pr.constrain(V, [(w, G_w), (x_0, U), (x_1, tU)] + list(zip(y, M)))
    c, responses, commitments = pr.prove_compact(blindings)
    return (c, responses), commitments


def issuance_verify(proof, sp, ip, amac, attrs, trace=None, batchable=False):
    """ProofOfIssuance::verify, issuance.rs:132-218 (via CredentialIssuance::verify, issuer.rs:48-57).
    Raises VerificationFailure.  batchable: proof = (commitments[3], responses), checked with zkp's verify_batchable (the
    BatchVerifier route the reference left commented out, issuance.rs:21-22)."""
    c, responses = proof
    t = Transcript(b"2019/1416 anonymous credential")
    vf = Verifier(b"2019/1416 issuance proof", t)
    w = vf.allocate_scalar(b"w")
    w_prime = vf.allocate_scalar(b"w'")
    x_0 = vf.allocate_scalar(b"x_0")
    x_1 = vf.allocate_scalar(b"x_1")
    y = [vf.allocate_scalar(b"y") for _ in range(sp.n)]
    one = vf.allocate_scalar(b"1")
    if trace is not None:
        trace["verifier"] = vf
    G_V = vf.allocate_point(b"G_V", sp.G_V.compress())
    G_w = vf.allocate_point(b"G_w", sp.G_w.compress())
    G_w_prime = vf.allocate_point(b"G_w_prime", sp.G_w_prime.compress())
    neg_G_x_0 = vf.allocate_point(b"-G_x_0", (-sp.G_x_0).compress())
    neg_G_x_1 = vf.allocate_point(b"-G_x_1", (-sp.G_x_1).compress())
    neg_G_y = [vf.allocate_point(b"-G_y", (-g).compress()) for g in sp.G_y]
    C_W = vf.allocate_point(b"C_W", ip.C_W.compress())
    I = vf.allocate_point(b"I", ip.I.compress())
    U = vf.allocate_point(b"U", amac.U.compress())
    V = vf.allocate_point(b"V", amac.V.compress())
    tU = vf.allocate_point(b"tU", (amac.U * amac.t).compress())
    M = [vf.allocate_point(b"M", m.compress()) for m in messages_from_attributes(attrs, sp)]
    vf.constrain(C_W, [(w, G_w), (w_prime, G_w_prime)])
    vf.constrain(I, [(one, G_V), (x_0, neg_G_x_0), (x_1, neg_G_x_1)] + list(zip(y, neg_G_y)))
    vf.constrain(V, [(w, G_w), (x_0, U), (x_1, tU)] + list(zip(y, M)))
    if batchable:
        vf.verify_batchable(c, responses)
    else:
        vf.verify_compact(c, responses)


# ---- credential.rs -------------------------------------------------------------

def hide_attribute(attrs, index):
    """credential.rs:77-97."""
    k, v = attrs[index]
    if k == "PS":
        attrs[index] = ("SS", v)
    elif k == "EP":
        attrs[index] = ("SP", v)
    elif k == "PP":
        raise ValueError("Public point attributes cannot be converted")


# ---- nizk/encryption.rs ---------------------------------------------------------

class ProofOfEncryption:
    """encryption.rs:32-41."""

    def __init__(self, proof, pk, E1, E2, index, C_y_1, C_y_2, C_y_3, C_y_2_prime):
        self.proof, self.pk, self.E1, self.E2, self.index = proof, pk, E1, E2, index
        self.C_y_1, self.C_y_2, self.C_y_3, self.C_y_2_prime = C_y_1, C_y_2, C_y_3, C_y_2_prime


def encryption_prove(sp, pt: Plaintext, index, kp: SymmetricKeypair, z, blindings):
    """ProofOfEncryption::prove, encryption.rs:58-142."""
    E1, E2 = kp.encrypt(pt)
    C_y_1_ = sp.G_y[0] * z + pt.M1
    C_y_2_ = sp.G_y[1] * z + pt.M2
    C_y_3_ = sp.G_y[2] * z + sp.G_m[index] * pt.m3
    C_y_2_prime_ = C_y_2_ * kp.a1
    z1_ = (-z * (kp.a0 + kp.a1 * pt.m3)) % L
    t = Transcript(b"2019/1416 anonymous credentials")
    pr = Prover(b"2019/1416 proof of encryption", t)
    a = pr.allocate_scalar(b"a", kp.a)
    a0 = pr.allocate_scalar(b"a0", kp.a0)
    a1 = pr.allocate_scalar(b"a1", kp.a1)
    m3 = pr.allocate_scalar(b"m3", pt.m3)
    zv = pr.allocate_scalar(b"z", z)
    z1 = pr.allocate_scalar(b"z1", z1_)
    pk, _ = pr.allocate_point(b"pk", kp.pk)
    G_a, _ = pr.allocate_point(b"G_a", sp.G_a)
    G_a_0, _ = pr.allocate_point(b"G_a_0", sp.G_a0)
    G_a_1, _ = pr.allocate_point(b"G_a_1", sp.G_a1)
    G_y_1, _ = pr.allocate_point(b"G_y_1", sp.G_y[0])
    G_y_2, _ = pr.allocate_point(b"G_y_2", sp.G_y[1])
    G_y_3, _ = pr.allocate_point(b"G_y_3", sp.G_y[2])
    G_m_3, _ = pr.allocate_point(b"G_m_3", sp.G_m[index])
    C_y_2, _ = pr.allocate_point(b"C_y_2", C_y_2_)
    C_y_3, _ = pr.allocate_point(b"C_y_3", C_y_3_)
    C_y_2_prime, _ = pr.allocate_point(b"C_y_2'", C_y_2_prime_)
    C_y_1_minus_E2, _ = pr.allocate_point(b"C_y_1-E2", C_y_1_ - E2)
    E1v, _ = pr.allocate_point(b"E1", E1)
    minus_E1, _ = pr.allocate_point(b"-E1", -E1)
    pr.constrain(pk, [(a, G_a), (a0, G_a_0), (a1, G_a_1)])
    pr.constrain(C_y_1_minus_E2, [(zv, G_y_1), (a, minus_E1)])
    pr.constrain(C_y_2_prime, [(a1, C_y_2)])
    pr.constrain(E1v, [(a0, C_y_2), (m3, C_y_2_prime), (z1, G_y_2)])
    pr.constrain(C_y_3, [(zv, G_y_3), (m3, G_m_3)])
    c, responses, _ = pr.prove_compact(blindings)
    return ProofOfEncryption((c, responses), kp.pk, E1, E2, index, C_y_1_, C_y_2_, C_y_3_, C_y_2_prime_)


def encryption_verify(pe: ProofOfEncryption, sp, trace=None):
    """ProofOfEncryption::verify, encryption.rs:154-210.  Raises VerificationFailure."""
    t = Transcript(b"2019/1416 anonymous credentials")
    vf = Verifier(b"2019/1416 proof of encryption", t)
    a = vf.allocate_scalar(b"a")
    a0 = vf.allocate_scalar(b"a0")
    a1 = vf.allocate_scalar(b"a1")
    m3 = vf.allocate_scalar(b"m3")
    z = vf.allocate_scalar(b"z")
    z1 = vf.allocate_scalar(b"z1")
    if trace is not None:
        trace.setdefault("enc_verifiers", []).append(vf)
    pk = vf.allocate_point(b"pk", pe.pk.compress())
    G_a = vf.allocate_point(b"G_a", sp.G_a.compress())
    G_a_0 = vf.allocate_point(b"G_a_0", sp.G_a0.compress())
    G_a_1 = vf.allocate_point(b"G_a_1", sp.G_a1.compress())
    G_y_1 = vf.allocate_point(b"G_y_1", sp.G_y[0].compress())
    G_y_2 = vf.allocate_point(b"G_y_2", sp.G_y[1].compress())
    G_y_3 = vf.allocate_point(b"G_y_3", sp.G_y[2].compress())
    G_m_3 = vf.allocate_point(b"G_m_3", sp.G_m[pe.index].compress())
    C_y_2 = vf.allocate_point(b"C_y_2", pe.C_y_2.compress())
    C_y_3 = vf.allocate_point(b"C_y_3", pe.C_y_3.compress())
    C_y_2_prime = vf.allocate_point(b"C_y_2'", pe.C_y_2_prime.compress())
    C_y_1_minus_E2 = vf.allocate_point(b"C_y_1-E2", (pe.C_y_1 - pe.E2).compress())
    E1 = vf.allocate_point(b"E1", pe.E1.compress())
    minus_E1 = vf.allocate_point(b"-E1", (-pe.E1).compress())
    vf.constrain(pk, [(a, G_a), (a0, G_a_0), (a1, G_a_1)])
    vf.constrain(C_y_1_minus_E2, [(z, G_y_1), (a, minus_E1)])
    vf.constrain(C_y_2_prime, [(a1, C_y_2)])
    vf.constrain(E1, [(a0, C_y_2), (m3, C_y_2_prime), (z1, G_y_2)])
    vf.constrain(C_y_3, [(z, G_y_3), (m3, G_m_3)])
    vf.verify_compact(*pe.proof)


# ---- nizk/presentation.rs -------------------------------------------------------

class Presentation:
    """ProofOfValidCredential, presentation.rs:118-127.
    encrypted_attributes: list of (kind, value) with kinds 'PS' (int), 'SS' (None), 'PP' (Point), 'SP' (None)."""

    def __init__(self, proof, proofs_of_encryption, encrypted_attributes, hidden_scalar_indices, C_x_0, C_x_1, C_V, C_y):
        self.proof = proof
        self.proofs_of_encryption = proofs_of_encryption
        self.encrypted_attributes = encrypted_attributes
        self.hidden_scalar_indices = hidden_scalar_indices
        self.C_x_0, self.C_x_1, self.C_V, self.C_y = C_x_0, C_x_1, C_V, C_y


LINK_BASE_LABEL, LINK_LABEL = b"G_y-G_y_1", b"C_y-C_y_1"


def presentation_prove(sp, ip, amac, attrs, keypair, z_, blindings, enc_blindings, linked=False):
    """ProofOfValidCredential::prove, presentation.rs:139-321 (via AnonymousCredential::show,
    credential.rs:37-46).  z_, blindings (3+h_s) and enc_blindings (6 per SecretPoint) are
    supplied (see zkp.py docstring).

    linked=True is NOT the reference's protocol: it is the fix its authors left as a TODO (README.md:119-122,
    presentation.rs:292 "don't we also need DLEQ between the plaintext here and that in the commitments above?").  The
    proof of encryption of hidden plaintext i commits to M1 as C_y_1 = z*G_y[0] + M1 while the credential proof commits to
    it as C_y[i] = z*G_y[i] + M1, and nothing ties the two.  The linked statement adds, per hidden plaintext at index i > 0,
    the allocated points G_y[i] - G_y[0] (label "G_y-G_y_1") and C_y[i] - C_y_1 (label "C_y-C_y_1") -- after the G_m
    points, before Z -- and the constraint  C_y[i] - C_y_1 = z * (G_y[i] - G_y[0])  after the C_y constraints: a DLEQ with
    Z = z*I.  At index 0 both commitments use G_y[0], so the verifier requires C_y[0] == C_y_1 outright."""
    if keypair is None and any(k == "SP" for k, _ in attrs):
        raise ValueError("NoSymmetricKey")
    z_0_ = (-amac.t * z_) % L
    C_y_, H_s_ = [], []
    for i, (k, v) in enumerate(attrs):
        if k in ("PP", "EP", "PS"):
            C_y_.append(sp.G_y[i] * z_)
        elif k == "SP":
            C_y_.append(sp.G_y[i] * z_ + v.M1)
        else:  # SS
            C_y_.append(sp.G_y[i] * z_ + sp.G_m[i] * v)
            H_s_.append((i, sp.G_m[i], v))
    C_x_0_ = sp.G_x_0 * z_ + amac.U
    C_x_1_ = sp.G_x_1 * z_ + amac.U * amac.t
    C_V_ = sp.G_V * z_ + amac.V
    Z_ = ip.I * z_
    t = Transcript(b"2019/1416 anonymous credential")
    pr = Prover(b"2019/1416 presentation proof", t)
    z = pr.allocate_scalar(b"z", z_)
    z_0 = pr.allocate_scalar(b"z_0", z_0_)
    tv = pr.allocate_scalar(b"t", amac.t)
    H_s = {}
    hidden_scalar_indices = []
    for i, _bp, m in H_s_:
        H_s[i] = pr.allocate_scalar(b"m", m)
        hidden_scalar_indices.append(i)
    I, _ = pr.allocate_point(b"I", ip.I)
    C_x_1, _ = pr.allocate_point(b"C_x_1", C_x_1_)
    C_x_0, _ = pr.allocate_point(b"C_x_0", C_x_0_)
    G_x_0, _ = pr.allocate_point(b"G_x_0", sp.G_x_0)
    G_x_1, _ = pr.allocate_point(b"G_x_1", sp.G_x_1)
    C_y = []
    for i, cm in enumerate(C_y_):
        if attrs[i][0] == "SP":
            continue
        C_y.append(pr.allocate_point(b"C_y", cm)[0])
    G_y = [pr.allocate_point(b"G_y", g)[0] for g in sp.G_y]
    G_m = {}
    for i, bp, _m in H_s_:
        G_m[i] = pr.allocate_point(b"G_m", bp)[0]
    links = []
    if linked:
        for i, (k, _v) in enumerate(attrs):
            if k == "SP" and i > 0:
                base = sp.G_y[i] - sp.G_y[0]
                Lv = pr.allocate_point(LINK_BASE_LABEL, base)[0]
                Dv = pr.allocate_point(LINK_LABEL, base * z_)[0]       # = C_y[i] - C_y_1
                links.append((Dv, Lv))
    Z, _ = pr.allocate_point(b"Z", Z_)
    pr.constrain(Z, [(z, I)])
    pr.constrain(C_x_1, [(tv, C_x_0), (z_0, G_x_0), (z, G_x_1)])
    # presentation.rs:267-273 -- compacted-index loop: i indexes the compacted
    # C_y list but is used to index attributes / G_y / G_m / H_s.
    for i, C_y_i in enumerate(C_y):
        k = attrs[i][0]
        if k == "SP":
            continue
        elif k == "SS":
            pr.constrain(C_y_i, [(z, G_y[i]), (H_s[i], G_m[i])])
        else:
            pr.constrain(C_y_i, [(z, G_y[i])])
    for Dv, Lv in links:
        pr.constrain(Dv, [(z, Lv)])
    assert len(blindings) == len(pr.scalars)
    c, responses, _ = pr.prove_compact(blindings)
    poes, enc_attrs = [], []
    eb = iter(enc_blindings)
    for i, (k, v) in enumerate(attrs):
        if k == "PS":
            enc_attrs.append(("PS", v))
        elif k == "SS":
            enc_attrs.append(("SS", None))
        elif k == "PP":
            enc_attrs.append(("PP", v))
        elif k == "EP":
            enc_attrs.append(("PP", v.M1))
        else:
            poes.append((i, encryption_prove(sp, v, i, keypair, z_, next(eb))))
            enc_attrs.append(("SP", None))
    return Presentation((c, responses), poes, enc_attrs, hidden_scalar_indices, C_x_0_, C_x_1_, C_V_, C_y_)


def presentation_verify(p: Presentation, issuer: Issuer, trace=None, linked=False):
    """ProofOfValidCredential::verify, presentation.rs:324-443.  Raises VerificationFailure
    (the only CredentialError this path can yield) -- or IndexError/KeyError where the
    Rust panics on structurally malformed input.  linked: see presentation_prove."""
    sp, ip, sk = issuer.system_parameters, issuer.issuer_parameters, issuer.amacs_key
    # :342-352
    Z_ = p.C_V - sk.W - p.C_x_0 * sk.x_0 - p.C_x_1 * sk.x_1
    for i, (k, v) in enumerate(p.encrypted_attributes):
        if k == "PS":
            x = p.C_y[i] + sp.G_m[i] * v
        elif k == "PP":
            x = p.C_y[i] + v
        else:
            x = p.C_y[i]
        Z_ = Z_ - x * sk.y[i]
    if trace is not None:
        trace["Z"] = Z_.compress()
    t = Transcript(b"2019/1416 anonymous credential")
    vf = Verifier(b"2019/1416 presentation proof", t)
    if trace is not None:
        trace["verifier"] = vf
    z = vf.allocate_scalar(b"z")
    z_0 = vf.allocate_scalar(b"z_0")
    tv = vf.allocate_scalar(b"t")
    H_s = {}
    for i in p.hidden_scalar_indices:
        H_s[i] = vf.allocate_scalar(b"m")
    I = vf.allocate_point(b"I", ip.I.compress())
    C_x_1 = vf.allocate_point(b"C_x_1", p.C_x_1.compress())
    C_x_0 = vf.allocate_point(b"C_x_0", p.C_x_0.compress())
    G_x_0 = vf.allocate_point(b"G_x_0", sp.G_x_0.compress())
    G_x_1 = vf.allocate_point(b"G_x_1", sp.G_x_1.compress())
    C_y = []
    for i, cm in enumerate(p.C_y):
        if p.encrypted_attributes[i][0] == "SP":
            continue
        C_y.append(vf.allocate_point(b"C_y", cm.compress()))
    G_y = [vf.allocate_point(b"G_y", g.compress()) for g in sp.G_y]
    G_m = {}
    for i in H_s:
        G_m[i] = vf.allocate_point(b"G_m", sp.G_m[i].compress())
    links = []
    if linked:
        poe_of = dict(p.proofs_of_encryption)
        for i, (k, _v) in enumerate(p.encrypted_attributes):
            if k != "SP":
                continue
            if i == 0:
                if p.C_y[0].compress() != poe_of[0].C_y_1.compress():
                    raise VerificationFailure("C_y[0] is not the proof of encryption's C_y_1")
                continue
            Lv = vf.allocate_point(LINK_BASE_LABEL, (sp.G_y[i] - sp.G_y[0]).compress())
            Dv = vf.allocate_point(LINK_LABEL, (p.C_y[i] - poe_of[i].C_y_1).compress())
            links.append((Dv, Lv))
    Z = vf.allocate_point(b"Z", Z_.compress())
    vf.constrain(Z, [(z, I)])
    vf.constrain(C_x_1, [(tv, C_x_0), (z_0, G_x_0), (z, G_x_1)])
    for i, C_y_i in enumerate(C_y):  # :427-433 compacted-index loop
        k = p.encrypted_attributes[i][0]
        if k == "SP":
            continue
        elif k == "SS":
            vf.constrain(C_y_i, [(z, G_y[i]), (H_s[i], G_m[i])])
        else:
            vf.constrain(C_y_i, [(z, G_y[i])])
    for Dv, Lv in links:
        vf.constrain(Dv, [(z, Lv)])
    vf.verify_compact(*p.proof)
    for _i, poe in p.proofs_of_encryption:  # :438-440
        encryption_verify(poe, sp, trace)
