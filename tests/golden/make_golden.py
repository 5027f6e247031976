"""Generates tests/golden/*.json from the Python big-int oracle (oracle/pyoracle).

The reference has no golden vectors of its own and cannot be run here, so these
fixtures pin (a) the external KATs the oracle was validated against and (b) the oracle's own
outputs on seeded inputs, so that the C oracle, the CUDA path and future rounds are all held to the
same bytes.  Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.pyoracle import aeonflux as A, flat as F, ristretto as R, synth as S  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def hexl(ws):
    return [w.hex() for w in ws]


def presentation_case(name, n, request, hide, config, items, corrupt=True):
    iss = S.make_issuer(n)
    case = {"name": name, "n": n, "request": list(request), "hide": list(hide), "config": config.decode(),
            "sysparams": iss.system_parameters.to_bytes().hex(), "issuer_pub": iss.issuer_parameters.to_bytes().hex(),
            "secret": iss.amacs_key.to_bytes().hex(), "items": []}
    for i in items:
        it = S.make_item(iss, request, hide, config, i)
        p = it["presentation"]
        kinds = F.presentation_kinds(p)
        words = F.presentation_to_words(p)
        verdict, tr = F.verify_flat(iss, kinds, words)
        iw = F.issuance_to_words(it["request_attrs"], it["amac"], it["issuance_proof"])
        ikinds = F.request_kinds(it["request_attrs"])
        iv, itr = F.verify_issuance_flat(iss.system_parameters, iss.issuer_parameters, ikinds, iw)
        entry = {"item": i, "kinds": list(kinds), "words": hexl(words), "verdict": verdict, "Z": tr["Z"].hex(),
                 "commitments": hexl(tr["commitments"]), "challenges": hexl(tr["challenges"]),
                 "issuance_kinds": list(ikinds), "issuance_words": hexl(iw), "issuance_verdict": iv,
                 "issuance_commitments": hexl(itr["commitments"]), "issuance_challenges": hexl(itr["challenges"]),
                 "corrupted": []}
        if corrupt:
            for cls in S.CORRUPTIONS:
                w2 = S.corrupt(kinds, words, cls, A.ShakeRng(b"corrupt" + cls.encode()))
                if w2 is None:
                    continue
                v2, tr2 = F.verify_flat(iss, kinds, w2)
                entry["corrupted"].append({"class": cls, "words": hexl(w2), "verdict": v2,
                                           "Z": tr2["Z"].hex() if tr2["Z"] else None,
                                           "commitments": hexl(tr2["commitments"]), "challenges": hexl(tr2["challenges"])})
        case["items"].append(entry)
    json.dump(case, open(os.path.join(HERE, name + ".json"), "w"), indent=0)
    print(name, "ok", [e["verdict"] for e in case["items"]])


def main():
    # config 1: the README flow (README.md:44-117), message = the README's string
    iss = S.make_issuer(4)
    it = S.make_item(iss, S.README4_REQUEST, S.README4_HIDE, b"readme", 0, message=S.README_MESSAGE)
    A.issuance_verify(it["issuance_proof"], iss.system_parameters, iss.issuer_parameters, it["amac"], it["request_attrs"])
    iss.verify(it["presentation"])
    presentation_case("readme4", 4, S.README4_REQUEST, S.README4_HIDE, b"readme4", [0, 1, 2])
    presentation_case("s16", 16, S.S16_REQUEST, S.S16_HIDE, b"s16", [0], corrupt=False)
    # shapes from the reference's own tests (presentation.rs:460-638) and the A.6.1 quirk shapes
    presentation_case("revealed10", 10, ("PP", "PP", "PS", "PS", "PP", "PS", "PP", "PS", "PS", "PP"), (), b"revealed10", [0], corrupt=False)
    presentation_case("plain1_hidden", 1, ("EP",), (0,), b"plain1", [0], corrupt=False)
    presentation_case("scalar1", 1, ("PS",), (), b"scalar1", [0])
    presentation_case("quirk_sp_first", 3, ("EP", "PS", "PS"), (0,), b"quirk1", [0], corrupt=False)   # always fails (A.6.1)
    presentation_case("quirk_sp_middle", 3, ("PS", "EP", "PS"), (1,), b"quirk2", [0], corrupt=False)  # passes (A.6.1)
    # primitive vectors
    rng = A.ShakeRng(b"golden-prims")
    prims = {"decompress": [], "from_uniform": [], "scalarmult": [], "wide_reduce": []}
    for _ in range(64):
        b = rng.fill(32)
        p = R.decompress(b)
        prims["decompress"].append({"in": b.hex(), "valid": p is not None})
    for _ in range(16):
        h = rng.fill(64)
        P = R.from_uniform_bytes(h)
        s = rng.scalar()
        prims["from_uniform"].append({"in": h.hex(), "out": P.compress().hex()})
        prims["scalarmult"].append({"s": R.sc_to_bytes(s).hex(), "p": P.compress().hex(), "out": (P * s).compress().hex()})
        prims["wide_reduce"].append({"in": h.hex(), "out": R.sc_to_bytes(R.sc_from_wide(h)).hex()})
    json.dump(prims, open(os.path.join(HERE, "prims.json"), "w"), indent=0)


if __name__ == "__main__":
    main()
