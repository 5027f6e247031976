"""ORACLE (test infrastructure, NOT the product): deterministic synthetic issuers,
credentials and presentations.  All randomness is
SHAKE-256("aeonflux-b200/" || config || le64(item) || le32(counter)).

The C oracle (oracle/c/afx_oracle.c) implements the same draw order so both produce
identical bytes for the same (config, item).

Draw order for one item of a shape (request kinds r[0..n), hidden index set H):
  1. per attribute in order: 'PS' -> scalar (64 B), 'PP' -> point (64 B), 'EP' -> 30 plaintext bytes
  2. t (64 B), U (64 B), issuance blindings: n+5 scalars (64 B each)
  3. symmetric master secret (64 B)            [always drawn]
  4. z (64 B), presentation blindings: 3+h_s scalars, then 6 scalars per hidden plaintext in index order
"""
from . import aeonflux as A
from . import flat as F
from . import ristretto as R

PREFIX = b"aeonflux-b200/"

README4_REQUEST = ("PS", "PS", "PP", "EP")
README4_HIDE = (0, 3)
S16_REQUEST = ("PS", "PS", "PS", "PS", "PS", "PS", "PP", "PP") + ("EP",) * 8
S16_HIDE = (0, 1) + tuple(range(8, 16))
README_MESSAGE = b"This is a tsunami alert test.."


def item_rng(config: bytes, item: int) -> A.ShakeRng:
    return A.ShakeRng(PREFIX + config + item.to_bytes(8, "little"))


def make_issuer(n: int, tag: bytes = b"issuer") -> A.Issuer:
    rng = A.ShakeRng(PREFIX + tag + n.to_bytes(4, "little"))
    sp = A.SystemParameters.generate(rng, n)
    return A.Issuer.new(sp, rng)


def make_item(issuer: A.Issuer, request, hide, config: bytes, item: int, message=None):
    """Full user+issuer flow for one synthetic credential.  Returns a dict with the
    request attributes, issuance (amac, proof), and the presentation."""
    sp, ip = issuer.system_parameters, issuer.issuer_parameters
    rng = item_rng(config, item)
    attrs = []
    for k in request:
        if k == "PS":
            attrs.append(("PS", rng.scalar()))
        elif k == "PP":
            attrs.append(("PP", rng.point()))
        elif k == "EP":
            msg = rng.fill(30)
            if message is not None:
                msg = message
            attrs.append(("EP", A.Plaintext.from_bytes30(msg)))
        else:
            raise ValueError(k)
    n = sp.n
    t = rng.scalar()
    U = rng.point()
    iss_blind = [rng.scalar() for _ in range(n + 5)]
    request_attrs = list(attrs)
    proof, (amac, _) = issuer.issue(list(attrs), None, blindings=iss_blind, t=t, U=U)
    ms = rng.fill(64)
    kp = A.SymmetricKeypair.derive(ms, sp)
    shown = list(attrs)
    for i in hide:
        A.hide_attribute(shown, i)
    z = rng.scalar()
    h_s = sum(1 for k, _ in shown if k == "SS")
    h_p = sum(1 for k, _ in shown if k == "SP")
    blind = [rng.scalar() for _ in range(3 + h_s)]
    enc_blind = [[rng.scalar() for _ in range(6)] for _ in range(h_p)]
    pres = A.presentation_prove(sp, ip, amac, shown, kp if h_p else None, z, blind, enc_blind)
    return {"request_attrs": request_attrs, "amac": amac, "issuance_proof": proof,
            "presentation": pres, "keypair": kp, "z": z}


def presentation_batch(issuer, request, hide, config: bytes, start: int, count: int):
    """-> (kinds bytes, list of per-item word lists)."""
    kinds, items = None, []
    for i in range(start, start + count):
        it = make_item(issuer, request, hide, config, i)
        kinds = F.presentation_kinds(it["presentation"])
        items.append(F.presentation_to_words(it["presentation"]))
    return kinds, items


# ---- corruption classes --------------------------------

CORRUPTIONS = ("response+1", "challenge+1", "C_x_0+B", "C_V+B", "revealed_scalar", "enc_E2+B",
               "enc_response+1", "undecodable_point", "noncanonical_scalar")


def _field_index(kinds):
    """Name -> word index for the flat presentation layout."""
    kinds = list(kinds)
    n = len(kinds)
    h_s = sum(1 for k in kinds if k == F.KIND_SS)
    idx = {"challenge": 0, "responses": 1, "C_x_0": 4 + h_s, "C_x_1": 5 + h_s, "C_V": 6 + h_s, "C_y": 7 + h_s}
    off = 7 + h_s + n
    idx["revealed"] = {}
    for i, k in enumerate(kinds):
        if k in (F.KIND_PS, F.KIND_PP):
            idx["revealed"][i] = off
            off += 1
    idx["enc"] = []
    for i, k in enumerate(kinds):
        if k == F.KIND_SP:
            idx["enc"].append(off)
            off += F.ENC_WORDS
    return idx


def corrupt(kinds, words, cls: str, rng: A.ShakeRng):
    """Return a corrupted copy of a flat presentation, or None if the class does not apply."""
    w = list(words)
    fi = _field_index(kinds)

    def sc_plus1(b):
        return R.sc_to_bytes(int.from_bytes(b, "little") + 1)

    def pt_plusB(b):
        return (R.decompress(b) + R.BASEPOINT).compress()

    if cls == "response+1":
        w[fi["responses"]] = sc_plus1(w[fi["responses"]])
    elif cls == "challenge+1":
        w[0] = sc_plus1(w[0])
    elif cls == "C_x_0+B":
        w[fi["C_x_0"]] = pt_plusB(w[fi["C_x_0"]])
    elif cls == "C_V+B":
        w[fi["C_V"]] = pt_plusB(w[fi["C_V"]])
    elif cls == "revealed_scalar":
        tgt = [i for i, k in enumerate(kinds) if k == F.KIND_PS]
        if not tgt:
            return None
        w[fi["revealed"][tgt[0]]] = R.sc_to_bytes(rng.scalar())
    elif cls == "enc_E2+B":
        if not fi["enc"]:
            return None
        w[fi["enc"][0] + 9] = pt_plusB(w[fi["enc"][0] + 9])
    elif cls == "enc_response+1":
        if not fi["enc"]:
            return None
        w[fi["enc"][0] + 1] = sc_plus1(w[fi["enc"][0] + 1])
    elif cls == "undecodable_point":
        while True:
            b = rng.fill(32)
            if R.decompress(b) is None:
                break
        w[fi["C_x_1"]] = b
    elif cls == "noncanonical_scalar":
        w[fi["responses"] + 1] = (int.from_bytes(w[fi["responses"] + 1], "little") + R.L).to_bytes(32, "little")
    else:
        raise ValueError(cls)
    return w
