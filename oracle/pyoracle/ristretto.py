"""ORACLE (test infrastructure, NOT the product): big-int restatement of the
ristretto255 group, GF(2^255-19) and scalars mod l, as used by aeonflux through
curve25519-dalek "2" (/root/reference/Cargo.toml:34, not vendored).

The algorithms are the published ones (RFC 9496 / dalek 2.x); SURVEY.md
Appendix A.1/A.2 is the in-repo spec.  Pinned in tests/test_oracle_kats.py by:
RFC 9496 constants and generator multiples, and libsodium 1.0.20's
crypto_core_ristretto255_* when that library is loadable.

Call sites in the reference that this serves: presentation.rs:342-351,373-412;
encryption.rs:172-185; issuance.rs:74-91,162-189; amacs.rs:234-235,267-270,289-290.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import
this package.
"""

P = 2**255 - 19
L = 2**252 + 27742317777372353535851937790883648493

D = (-121665 * pow(121666, P - 2, P)) % P
SQRT_M1 = pow(2, (P - 1) // 4, P)
ONE_MINUS_D_SQ = (1 - D * D) % P
D_MINUS_ONE_SQ = ((D - 1) * (D - 1)) % P


def _is_neg(x):
    return (x % P) & 1


def _abs(x):
    x %= P
    return P - x if x & 1 else x


def sqrt_ratio_i(u, v):
    """dalek FieldElement::sqrt_ratio_i."""
    u %= P
    v %= P
    v3 = v * v % P * v % P
    v7 = v3 * v3 % P * v % P
    r = (u * v3) % P * pow(u * v7 % P, (P - 5) // 8, P) % P
    check = v * r % P * r % P
    correct = check == u
    flipped = check == (-u) % P
    flipped_i = check == (-u * SQRT_M1) % P
    if flipped or flipped_i:
        r = r * SQRT_M1 % P
    if r & 1:
        r = P - r
    return (correct or flipped), r


def invsqrt(v):
    return sqrt_ratio_i(1, v)


_ok, _r = sqrt_ratio_i(1, (-1 - D) % P)
INVSQRT_A_MINUS_D = _r
# SQRT_AD_MINUS_ONE: dalek's constant is the root of (a*d - 1) = (-d - 1) given in
# RFC 9496 (25063068953384623474111414158702152701244531502492656460079210482610430750235).
SQRT_AD_MINUS_ONE = 25063068953384623474111414158702152701244531502492656460079210482610430750235
assert SQRT_AD_MINUS_ONE * SQRT_AD_MINUS_ONE % P == (-D - 1) % P


def fe_from_bytes(b):
    """dalek FieldElement::from_bytes: little endian, bit 255 ignored."""
    return int.from_bytes(b, "little") & ((1 << 255) - 1)


def fe_to_bytes(x):
    return (x % P).to_bytes(32, "little")


class Point:
    """Extended twisted Edwards point (X:Y:Z:T), a=-1, viewed as a ristretto255 element."""

    __slots__ = ("X", "Y", "Z", "T")

    def __init__(self, X, Y, Z, T):
        self.X, self.Y, self.Z, self.T = X % P, Y % P, Z % P, T % P

    @staticmethod
    def identity():
        return Point(0, 1, 1, 0)

    def __add__(self, o):
        A = (self.Y - self.X) * (o.Y - o.X) % P
        B = (self.Y + self.X) * (o.Y + o.X) % P
        C = 2 * D * self.T % P * o.T % P
        Dd = 2 * self.Z * o.Z % P
        E, F, G, H = B - A, Dd - C, Dd + C, B + A
        return Point(E * F, G * H, F * G, E * H)

    def __neg__(self):
        return Point(-self.X, self.Y, self.Z, -self.T)

    def __sub__(self, o):
        return self + (-o)

    def double(self):
        return self + self

    def __mul__(self, s):
        s = int(s) % L
        acc = Point.identity()
        for i in reversed(range(s.bit_length())):
            acc = acc.double()
            if (s >> i) & 1:
                acc = acc + self
        return acc

    __rmul__ = __mul__

    def __eq__(self, o):
        return (self.X * o.Y - self.Y * o.X) % P == 0 or (self.X * o.X - self.Y * o.Y) % P == 0

    def is_identity(self):
        return self == Point.identity()

    def compress(self):
        """RistrettoPoint::compress."""
        X, Y, Z, T = self.X, self.Y, self.Z, self.T
        u1 = (Z + Y) * (Z - Y) % P
        u2 = X * Y % P
        _, inv = invsqrt(u1 * u2 % P * u2 % P)
        i1 = inv * u1 % P
        i2 = inv * u2 % P
        z_inv = i1 * (i2 * T % P) % P
        den_inv = i2
        if _is_neg(T * z_inv):
            X, Y = Y * SQRT_M1 % P, X * SQRT_M1 % P
            den_inv = i1 * INVSQRT_A_MINUS_D % P
        if _is_neg(X * z_inv):
            Y = (-Y) % P
        s = _abs(den_inv * (Z - Y))
        return fe_to_bytes(s)


def decompress(b):
    """CompressedRistretto::decompress. Returns Point or None."""
    if len(b) != 32:
        return None
    s = fe_from_bytes(b)
    if fe_to_bytes(s) != bytes(b) or (s & 1):
        return None
    ss = s * s % P
    u1 = (1 - ss) % P
    u2 = (1 + ss) % P
    u2s = u2 * u2 % P
    v = (-D * u1 % P * u1 - u2s) % P
    ok, I = invsqrt(v * u2s % P)
    Dx = I * u2 % P
    Dy = I * Dx % P * v % P
    x = _abs(2 * s * Dx)
    y = u1 * Dy % P
    t = x * y % P
    if (not ok) or (t & 1) or y == 0:
        return None
    return Point(x, y, 1, t)


def elligator(r0):
    """RistrettoPoint::elligator_ristretto_flavor (RFC 9496 4.3.4 MAP)."""
    r = SQRT_M1 * r0 % P * r0 % P
    Ns = (r + 1) * ONE_MINUS_D_SQ % P
    c = P - 1
    Dn = (c - D * r) % P * ((r + D) % P) % P
    sq, s = sqrt_ratio_i(Ns, Dn)
    s_prime = (-_abs(s * r0)) % P
    if not sq:
        s = s_prime
        c = r
    Nt = (c * (r - 1) % P * D_MINUS_ONE_SQ - Dn) % P
    W0 = 2 * s * Dn % P
    W1 = Nt * SQRT_AD_MINUS_ONE % P
    W2 = (1 - s * s) % P
    W3 = (1 + s * s) % P
    return Point(W0 * W3, W2 * W1, W1 * W3, W0 * W2)


def from_uniform_bytes(b):
    """RistrettoPoint::from_uniform_bytes: 64 bytes -> two Elligator maps, added."""
    assert len(b) == 64
    return elligator(fe_from_bytes(b[:32])) + elligator(fe_from_bytes(b[32:]))


BASEPOINT_COMPRESSED = bytes.fromhex("e2f2ae0a6abc4e71a884a961c500515f58e30b6aa582dd8db6a65945e08d2d76")
BASEPOINT = decompress(BASEPOINT_COMPRESSED)
IDENTITY_COMPRESSED = bytes(32)


# ---- scalars mod l ---------------------------------------------------------

def sc_from_wide(b):
    """Scalar::from_bytes_mod_order_wide."""
    assert len(b) == 64
    return int.from_bytes(b, "little") % L


def sc_from_canonical(b):
    """Scalar::from_canonical_bytes -> int or None."""
    x = int.from_bytes(b, "little")
    return x if x < L else None


def sc_to_bytes(x):
    return (x % L).to_bytes(32, "little")
