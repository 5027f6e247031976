"""CPU tests of everything around the kernels: the C ABI library loads and exports every symbol the header declares, and
the shape compiler / marshaling / stage logic (compiled for the host by the test-only emulation harness, tests/hostemu)
agree with the oracle on every golden shape.  No CUDA compute happens here."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from tests.common import GOLDEN_SHAPES, REQ, compare_with_oracle_trace, corrupt_batch, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_library_exports_declared_symbols():
    from aeonflux_b200 import build as B
    so = B.build()
    header = open(os.path.join(ROOT, "include", "aeonflux_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = sorted(set(re.findall(r"\b(afx_[a-z_0-9]+)\s*\(", header)))
    assert len(declared) >= 10
    lib = ctypes.CDLL(so)          # loads without a GPU: no CUDA call happens at load time
    for name in declared:
        assert hasattr(lib, name), "missing export " + name
    from aeonflux_b200._binding import Binding
    assert sorted(Binding.SYMBOLS) == declared
    lib.afx_version.restype = ctypes.c_char_p
    assert b"cuda sm_100a" in lib.afx_version()
    out = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_rust_ffi_is_generated_from_the_header():
    """rust/src/ffi.rs (the extern block of the Rust shim, which cannot be compiled here) is exactly what tools/gen_rust_ffi.py makes
    of include/aeonflux_b200.h, and declares every exported function."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_rust_ffi", os.path.join(ROOT, "tools", "gen_rust_ffi.py"))
    g = importlib.util.module_from_spec(spec); spec.loader.exec_module(g)
    text = g.generate()
    assert open(os.path.join(ROOT, "rust", "src", "ffi.rs")).read() == text, "run python tools/gen_rust_ffi.py"
    from aeonflux_b200._binding import Binding
    for name in Binding.SYMBOLS:
        assert ("pub fn %s(" % name) in text, name


def test_product_has_no_cpu_fallback():
    """The package never imports the oracle or the emulation harness, and refuses to run without the CUDA build."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "aeonflux_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".inc")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "hostemu.so" not in src, f
    import aeonflux_b200._lib as lib
    saved, lib._binding, lib.SO_PATH = (lib._binding, lib.SO_PATH), None, "/nonexistent/libaeonflux_b200.so"
    try:
        with pytest.raises(ImportError):
            lib.load()
    finally:
        lib._binding, lib.SO_PATH = saved


def build_hostemu():
    d = os.path.join(ROOT, "tests", "hostemu")
    so, src = os.path.join(d, "libafx_hostemu.so"), os.path.join(d, "hostemu.cpp")
    csrc = os.path.join(ROOT, "aeonflux_b200", "csrc")
    newest = max(os.path.getmtime(os.path.join(csrc, f)) for f in os.listdir(csrc) if not f.endswith(".so"))
    if not os.path.exists(so) or os.path.getmtime(so) < max(newest, os.path.getmtime(src)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-fvisibility=hidden", "-x", "c++", "-o", so, src])
    return so


@pytest.fixture(scope="session")
def emu():
    """The host-emulation build of the same C ABI (tests only)."""
    from aeonflux_b200._binding import Binding
    return Binding(ctypes.CDLL(build_hostemu()))


@pytest.mark.parametrize("name", GOLDEN_SHAPES)
def test_stage_logic_matches_oracle_on_golden_shapes(emu, coracle, name):
    from aeonflux_b200 import Issuer, PresentationBatch
    g = load_golden(name)
    sp, ip, sk = bytes.fromhex(g["sysparams"]), bytes.fromhex(g["issuer_pub"]), bytes.fromhex(g["secret"])
    orc = coracle.Issuer(sp, ip, sk)
    rk = bytes(REQ[k] for k in g["request"])
    kinds, pres, issu = orc.synth(rk, g["hide"], g["config"].encode(), 0, 5)
    e = g["items"][0]
    assert pres[0].tobytes().hex() == "".join(e["words"])
    pres[1, 1, 3] ^= 0x10
    pres[3, 0, 0] ^= 1
    iss = Issuer(sp, ip, sk, max_batch=3, _binding=emu)      # max_batch < count exercises chunking
    v, dbg = iss.verify_batch(PresentationBatch.from_items(kinds, pres), debug=True)
    ov, _, tr = orc.verify_presentations(kinds, pres, trace=True)
    compare_with_oracle_trace(v, dbg, ov, tr)
    assert dbg["Z"][0].tobytes().hex() == e["Z"]
    assert dbg["commitments"][:len(e["commitments"]), 0].tobytes().hex() == "".join(e["commitments"])   # filled as far as the reference gets
    assert dbg["challenges"][:len(e["challenges"]), 0].tobytes().hex() == "".join(e["challenges"])
    ik = bytes(e["issuance_kinds"])
    issu[2, len(ik) + 1, 5] ^= 2
    vi, dbgi = iss.verify_issuance_batch(PresentationBatch.from_items(ik, issu), debug=True)
    ovi, _, tri = orc.verify_issuances(ik, issu, trace=True)
    compare_with_oracle_trace(vi, dbgi, ovi, tri)
    assert dbgi["commitments"][:, 0].tobytes().hex() == "".join(e["issuance_commitments"])
    user = Issuer(sp, ip, None, max_batch=8, _binding=emu)     # user-side context: no secret key
    assert (user.verify_issuance_batch(PresentationBatch.from_items(ik, issu)) == ovi).all()
    # item-major wire blobs (one copy, kernels read the item-major layout): same verdicts; 5 items over max_batch 3 = pipelined chunks
    assert (iss.verify_wire(kinds, pres) == ov).all()
    assert (user.verify_wire(ik, issu, issuance=True) == ovi).all()
    assert len(iss.verify_wire(kinds, pres[:0])) == 0


@pytest.mark.parametrize("name", GOLDEN_SHAPES)
def test_issue_stage_logic_matches_golden_and_oracle(emu, coracle, name):
    """Issuer::issue (issuer.rs:111-124) with supplied rng output: byte-identical to the Python oracle's committed
    fixture and to the C oracle on fresh randomness; malformed requests; the result verifies."""
    from aeonflux_b200 import Issuer, RequestBatch
    from tests.common import golden_issue_request, words
    g = load_golden(name)
    sp, ip, sk = bytes.fromhex(g["sysparams"]), bytes.fromhex(g["issuer_pub"]), bytes.fromhex(g["secret"])
    n = g["n"]
    iss = Issuer(sp, ip, sk, max_batch=3, _binding=emu)
    e = g["items"][0]
    ik = bytes(e["issuance_kinds"])
    attrs, rnd = golden_issue_request(g, e)
    res, st, dbg = iss.issue_batch(RequestBatch.from_request(ik, attrs[None], rnd[None]), debug=True)
    assert st[0] == 0
    assert res.fields[:, 0].tobytes().hex() == "".join(e["issuance_words"])
    assert dbg["commitments"][:, 0].tobytes().hex() == "".join(e["issuance_commitments"])   # prover's blinding commitments == verifier's
    # fresh randomness, 5 items (two chunks), one malformed scalar or point
    orc = coracle.Issuer(sp, ip, sk)
    rng = np.random.default_rng(11)
    A = np.repeat(attrs[None], 5, axis=0)
    R = rng.integers(0, 256, (5, n + 7, 64), dtype=np.uint8)
    A[3, 0] = 0xff
    out, status, _ = orc.issue(ik, A, R)
    res, st = iss.issue_batch(RequestBatch.from_request(ik, A, R))
    assert (st == status).all() and list(status) == [0, 0, 0, 1, 0]
    assert (res.fields.transpose(1, 0, 2)[:, n:] == out).all()
    assert list(iss.verify_issuance_batch(res)) == [0, 0, 0, 1, 0]
    # item-major requests in, item-major issuances out (afx_issue_wire): the same bytes, ready for verify_wire
    reqs = np.ascontiguousarray(RequestBatch.from_request(ik, A, R).fields.transpose(1, 0, 2))
    wout, wst = iss.issue_wire(ik, reqs)
    assert (wst == status).all() and (wout[[0, 1, 2, 4]] == res.fields.transpose(1, 0, 2)[[0, 1, 2, 4]]).all()
    assert not wout[3].any()                                          # the malformed request: all-zero issuance, attributes included
    assert list(iss.verify_wire(ik, wout, issuance=True)) == [0, 0, 0, 1, 0]


@pytest.mark.parametrize("name", GOLDEN_SHAPES)
def test_show_stage_logic_matches_golden_and_oracle(emu, coracle, name):
    """AnonymousCredential::show (credential.rs:37-46) with supplied rng output: byte-identical to the presentations the oracle's
    prover makes from the same credential and rng bytes (item 0 = the committed Python-oracle fixture); the result verifies;
    malformed inputs give status 1 and all-zero words."""
    from aeonflux_b200 import Issuer
    g = load_golden(name)
    sp, ip, sk = bytes.fromhex(g["sysparams"]), bytes.fromhex(g["issuer_pub"]), bytes.fromhex(g["secret"])
    orc = coracle.Issuer(sp, ip, sk)
    rk = bytes(REQ[k] for k in g["request"])
    kinds, pres, _, showin = orc.synth(rk, g["hide"], g["config"].encode(), 0, 5, want_issuances=False, want_show_inputs=True)
    assert pres[0].tobytes().hex() == "".join(g["items"][0]["words"])
    user = Issuer(sp, ip, None, max_batch=3, _binding=emu)          # the user-side context holds no issuer secret
    fields = np.ascontiguousarray(showin.transpose(1, 0, 2))
    fields[2, 3, 0] ^= 1                                            # item 3: V no longer decodes (or is another point)
    res, st, dbg = user.show_batch(kinds, fields, debug=True)
    got = res.fields.transpose(1, 0, 2)
    ok = [0, 1, 2, 4]
    assert (got[ok] == pres[ok]).all()
    assert got[0].tobytes().hex() == "".join(g["items"][0]["words"])
    if st[3]:
        assert not got[3].any()
    wgot, wst = user.show_wire(kinds, np.ascontiguousarray(fields.transpose(1, 0, 2)))      # item-major in and out (afx_show_wire)
    assert (wst == st).all() and (wgot == got).all()
    iss = Issuer(sp, ip, sk, max_batch=8, _binding=emu)
    v = iss.verify_batch(res)
    ov, _ = orc.verify_presentations(kinds, np.ascontiguousarray(got))
    assert (v == ov).all() and list(v[ok]) == [g["items"][0]["verdict"]] * 4
    # the prover's blinding commitments are the ones the verifier recomputes
    _, vdbg = iss.verify_batch(res, debug=True)
    assert (dbg["commitments"][:, ok] == vdbg["commitments"][:, ok]).all() or g["items"][0]["verdict"] == 1


def test_issue_edges(emu, coracle):
    from aeonflux_b200 import Issuer, RequestBatch
    from aeonflux_b200._binding import AfxError
    sp, ip, sk = coracle.make_issuer(4)
    iss = Issuer(sp, ip, sk, max_batch=8, _binding=emu)
    kinds = bytes([0, 0, 2, 2])
    res, st = iss.issue_batch(RequestBatch(kinds, np.zeros((26, 0, 32), np.uint8)))            # empty
    assert res.count == 0 and len(st) == 0
    with pytest.raises(AfxError):                                                              # amacs.rs:285-287: wrong attribute count
        iss.issue_batch(RequestBatch(bytes([0, 0, 2]), np.zeros((23, 1, 32), np.uint8)))
    with pytest.raises(AfxError):                                                              # hidden kinds cannot be issued
        iss.issue_batch(RequestBatch(bytes([0, 1, 2, 3]), np.zeros((26, 1, 32), np.uint8)))
    user = Issuer(sp, ip, None, max_batch=8, _binding=emu)
    with pytest.raises(AfxError):
        user.issue_batch(RequestBatch(kinds, np.zeros((26, 1, 32), np.uint8)))
    # an all-zero 30-byte plaintext encodes to the identity (encoding.rs:55): issue succeeds, issuance verify rejects it
    # (issuance.rs:271-295)
    orc = coracle.Issuer(sp, ip, sk)
    _, _, issu = orc.synth(b"SSPP", [], b"issue-edges", 0, 2)
    A = np.ascontiguousarray(issu[:, :4]); A[1, 3] = 0
    R = np.random.default_rng(3).integers(0, 256, (2, 11, 64), dtype=np.uint8)
    res, st = iss.issue_batch(RequestBatch.from_request(kinds, A, R))
    out, status, _ = orc.issue(kinds, A, R)
    assert (res.fields.transpose(1, 0, 2)[:, 4:] == out).all() and not st.any()
    assert list(iss.verify_issuance_batch(res)) == [0, 1]


def test_corruption_classes_and_edges(emu, coracle):
    from aeonflux_b200 import Issuer, PresentationBatch
    from aeonflux_b200._binding import AfxError
    sp, ip, sk = coracle.make_issuer(4)
    orc = coracle.Issuer(sp, ip, sk)
    kinds, pres, _ = orc.synth(b"SSPE", [0, 3], b"edges", 0, 40, want_issuances=False)
    rng = np.random.default_rng(7)
    pts = pres[:, 5:8].reshape(-1, 32).copy()
    idx = corrupt_batch(pres, kinds, rng, 0.5, pts)
    iss = Issuer(sp, ip, sk, max_batch=64, _binding=emu)
    v, dbg = iss.verify_batch(PresentationBatch.from_items(kinds, pres), debug=True)
    ov, _, tr = orc.verify_presentations(kinds, pres, trace=True)
    compare_with_oracle_trace(v, dbg, ov, tr)
    assert ov[idx].all() and not np.delete(ov, idx).any()
    # empty batch
    assert len(iss.verify_batch(PresentationBatch(kinds, np.zeros((28, 0, 32), np.uint8)))) == 0
    # wrong field count, wrong attribute count, missing secret
    with pytest.raises(AfxError):
        iss.verify_batch(PresentationBatch(kinds, np.zeros((27, 2, 32), np.uint8)))
    with pytest.raises(AfxError):
        iss.verify_batch(PresentationBatch(bytes([0, 0, 0]), np.zeros((12, 1, 32), np.uint8)))
    user = Issuer(sp, ip, None, max_batch=4, _binding=emu)
    with pytest.raises(AfxError):
        user.verify_batch(PresentationBatch.from_items(kinds, pres[:2]))
    # issuer material that does not decode
    bad = bytearray(sp); bad[4 + 32] ^= 1
    with pytest.raises(AfxError):
        Issuer(bytes(bad), ip, sk, _binding=emu)
    with pytest.raises(AfxError):
        Issuer(sp[:-1], ip, sk, _binding=emu)


def test_key_material_layouts_round_trip(emu, coracle):
    """aeonflux_b200.wire: the reference's to_bytes layouts, Issuer::{to,from}_bytes, and the SecretKey::from_bytes y-loop fix."""
    from aeonflux_b200 import Issuer, PresentationBatch, wire
    for name in ("readme4", "scalar1", "s16"):
        g = load_golden(name)
        sp, ip, sk = bytes.fromhex(g["sysparams"]), bytes.fromhex(g["issuer_pub"]), bytes.fromhex(g["secret"])
        n = g["n"]
        assert len(sp) == wire.system_parameters_size(n) and len(sk) == wire.secret_key_size(n)
        P = wire.split_system_parameters(sp)
        assert P["n"] == n and len(P["G_y"]) == max(n, 3) and len(P["G_m"]) == n
        assert P["G"].hex() == "e2f2ae0a6abc4e71a884a961c500515f58e30b6aa582dd8db6a65945e08d2d76"     # the ristretto basepoint
        assert wire.join_system_parameters(P) == sp
        K = wire.split_secret_key(sk)
        assert wire.join_secret_key(K) == sk
        if n > 1:
            assert len(set(K["y"])) == n            # every y_i read from its own bytes (amacs.rs:148-150 would repeat y_0)
        assert wire.issuer_parameters_from_bytes(wire.issuer_parameters_to_bytes(ip[:32], ip[32:])) == (ip[:32], ip[32:])
        blob = wire.issuer_to_bytes(sp, ip, sk)
        assert wire.issuer_from_bytes(blob) == (sp, ip, sk)
    # an Issuer restored from Issuer::to_bytes bytes verifies like the original
    g = load_golden("readme4")
    sp, ip, sk = bytes.fromhex(g["sysparams"]), bytes.fromhex(g["issuer_pub"]), bytes.fromhex(g["secret"])
    iss = Issuer.from_bytes(wire.issuer_to_bytes(sp, ip, sk), max_batch=4, _binding=emu)
    e = g["items"][0]
    from tests.common import words
    assert list(iss.verify_batch(PresentationBatch.from_items(bytes(e["kinds"]), words(e["words"])[None]))) == [0]
    for bad in (sp[:-1], b"\x00\x00\x00\x00" + sp[4:], b""):
        with pytest.raises(ValueError):
            wire.split_system_parameters(bad)
    nc = bytearray(sk); nc[4:36] = (wire.L).to_bytes(32, "little")
    with pytest.raises(ValueError):
        wire.split_secret_key(bytes(nc))
    with pytest.raises(ValueError):
        wire.issuer_from_bytes(blob[:-1])


def check_primitives(iss, coracle):
    """The engine's field / group / scalar primitives against the committed primitive vectors (tests/golden/prims.json, from the
    Python oracle, itself pinned to libsodium and RFC 9496) and, on random inputs, against the C oracle's hooks."""
    import json
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "prims.json")))
    L = coracle.lib()
    hx = lambda rows, k: np.frombuffer(b"".join(bytes.fromhex(r[k]) for r in rows), np.uint8)
    # decompress validity + round trip, incl. the identity, non-canonical and negative encodings
    enc = hx(g["decompress"], "in").reshape(-1, 32)
    extra = np.zeros((4, 32), np.uint8); extra[1, 0] = 1; extra[2] = 0xff; extra[3, :] = 0xed; extra[3, 31] = 0x7f   # 0 (identity), 1 (negative), 2^256-1, p
    enc = np.concatenate([enc, extra])
    out, ok = iss.selftest_primitive("decompress_compress", enc)
    assert list(ok[:64]) == [int(r["valid"]) for r in g["decompress"]] and list(ok[64:]) == [1, 0, 0, 0]
    assert (out[ok == 1] == enc[ok == 1]).all()                      # compress(decompress(x)) == x for every valid encoding
    out, ok = iss.selftest_primitive("from_uniform", hx(g["from_uniform"], "in").reshape(-1, 64))
    assert out.tobytes() == hx(g["from_uniform"], "out").tobytes()
    sm = np.concatenate([hx(g["scalarmult"], "s").reshape(-1, 32), hx(g["scalarmult"], "p").reshape(-1, 32)], axis=1)
    out, ok = iss.selftest_primitive("scalarmult", sm)
    assert ok.all() and out.tobytes() == hx(g["scalarmult"], "out").tobytes()
    out, ok = iss.selftest_primitive("wide_reduce", hx(g["wide_reduce"], "in").reshape(-1, 64))
    assert out.tobytes() == hx(g["wide_reduce"], "out").tobytes()
    # RFC 9496 Appendix A (tests/golden/rfc9496.json): generator multiples, every invalid encoding, hash-to-group
    rfc = json.load(open(os.path.join(ROOT, "tests", "golden", "rfc9496.json")))
    mult = np.frombuffer(bytes.fromhex("".join(rfc["multiples"])), np.uint8).reshape(16, 32)
    ks = np.zeros((16, 32), np.uint8); ks[:, 0] = np.arange(16)
    out, ok = iss.selftest_primitive("scalarmult", np.concatenate([ks, np.repeat(mult[1:2], 16, axis=0)], axis=1))
    assert ok.all() and (out == mult).all()
    out, ok = iss.selftest_primitive("ladder_scalarmult", np.concatenate([ks, np.repeat(mult[1:2], 16, axis=0)], axis=1))
    assert ok.all() and (out == mult).all()
    out, ok = iss.selftest_primitive("decompress_compress", mult)
    assert ok.all() and (out == mult).all()
    inv = np.frombuffer(bytes.fromhex("".join(h for grp in rfc["invalid"].values() for h in grp)), np.uint8).reshape(-1, 32)
    out, ok = iss.selftest_primitive("decompress_compress", inv)
    assert len(inv) == 29 and not ok.any()
    h2g = np.frombuffer(bytes.fromhex("".join(e["input"] for e in rfc["hash_to_group"])), np.uint8).reshape(-1, 64)
    out, ok = iss.selftest_primitive("from_uniform", h2g)
    assert out.tobytes().hex() == "".join(e["output"] for e in rfc["hash_to_group"])
    # random inputs against the C oracle: edge scalars (0, 1, l-1) times a point, wide reductions of extreme values, a*b+c
    rng = np.random.default_rng(17)
    Lm1 = (2**252 + 27742317777372353535851937790883648493 - 1).to_bytes(32, "little")
    pts = out_pts = iss.selftest_primitive("from_uniform", rng.integers(0, 256, (40, 64), dtype=np.uint8))[0]
    scal = rng.integers(0, 256, (40, 32), dtype=np.uint8); scal[:, 31] &= 0x0f
    scal[0] = 0; scal[1] = 0; scal[1, 0] = 1; scal[2] = np.frombuffer(Lm1, np.uint8)
    got, ok = iss.selftest_primitive("scalarmult", np.concatenate([scal, pts], axis=1))
    for i in range(40):
        exp = np.zeros(32, np.uint8)
        assert L.afxo_scalarmult(scal[i].ctypes.data, pts[i].ctypes.data, exp.ctypes.data, 0) == 1
        assert (got[i] == exp).all(), i
    wide = rng.integers(0, 256, (20, 64), dtype=np.uint8); wide[0] = 0xff; wide[1] = 0
    got, _ = iss.selftest_primitive("wide_reduce", wide)
    abc = rng.integers(0, 256, (20, 96), dtype=np.uint8); abc[:, 31] &= 0x0f; abc[:, 63] &= 0x0f; abc[:, 95] &= 0x0f
    abc[0, :32] = np.frombuffer(Lm1, np.uint8); abc[0, 32:64] = np.frombuffer(Lm1, np.uint8); abc[0, 64:] = np.frombuffer(Lm1, np.uint8)
    got2, _ = iss.selftest_primitive("sc_muladd", abc)
    for i in range(20):
        exp = np.zeros(32, np.uint8)
        L.afxo_sc_from_wide(wide[i].ctypes.data, exp.ctypes.data)
        assert (got[i] == exp).all()
        L.afxo_sc_muladd(abc[i, :32].ctypes.data, abc[i, 32:64].ctypes.data, abc[i, 64:].ctypes.data, exp.ctypes.data)
        assert (got2[i] == exp).all()
    # the ladder forms k_ladders / k_msm_ct use (completed coordinates) against the plain extended-coordinates ladder and the oracle
    got3, ok3 = iss.selftest_primitive("ladder_scalarmult", np.concatenate([scal, pts], axis=1))
    got, ok = iss.selftest_primitive("scalarmult", np.concatenate([scal, pts], axis=1))
    assert ok3.all() and (got3 == got).all()
    check_field_edges(iss)
    # biased radix-4096 recoding of the constant-base scalars: the digits recompose to the scalar, for edge patterns of 12-bit fields
    # (0x000 -> digit -2048, 0x800 -> 0, 0xfff -> 2047), the group order's neighbours and random canonical scalars
    ell = 2**252 + 27742317777372353535851937790883648493
    pat = lambda f: sum(f << (12 * j) for j in range(21))
    cases = [0, 1, ell - 1, ell, 2**252, 2**253 - 1, pat(0x800), pat(0x7ff), pat(0xfff), pat(0x000) + (1 << 252), pat(0x801), 2**256 - 2**252 - 1]
    cases += [int.from_bytes(rng.bytes(32), "little") % ell for _ in range(500)]
    cases += [sum(int(rng.choice([0, 0x7ff, 0x800, 0x801, 0xfff, 1])) << (12 * j) for j in range(21)) for _ in range(200)]
    rec_in = np.frombuffer(b"".join(c.to_bytes(32, "little") for c in cases), np.uint8).reshape(-1, 32)
    rec_out, rec_ok = iss.selftest_primitive("recode4096", rec_in)
    assert rec_ok.all() and (rec_out == rec_in).all()


def check_field_edges(iss, n_random=3000):
    """fe_mul / fe_sq / fe_add / fe_sub on raw 256-bit limb vectors -- every value in [0, 2^256) is a legal lazily reduced
    element -- against Python integers: hand-picked extremes (0, 1, p-1, p, p+1, 2p, 2p+1 .. 2^256-1, 2^255, 38-multiples) in all
    pairs, and vectors whose limbs are drawn from edge words (0, 1, 2^32-1, 2^32-38, ...), which is where the carry folds of
    the device code (and nowhere else) take their rare paths."""
    P = 2**255 - 19
    M = 2**256
    special = [0, 1, 2, 18, 19, 37, 38, 39, P - 1, P, P + 1, P + 18, P + 19, 2 * P - 1, 2 * P, 2 * P + 1, M - 39, M - 38, M - 37, M - 2, M - 1,
               2**255, 2**255 - 1, 2**255 + 18, 2**128, 2**128 - 1, M - 2**128, M - 2**32, 2**32 - 1, 2**32, (M - 1) // 3, (M - 1) // 38]
    rng = np.random.default_rng(23)
    words = np.array([0, 1, 2, 37, 38, 0x7fffffff, 0x80000000, 0xfffffffe, 0xffffffff, 0xffffffda, 0xffffffd9, 0xffffffed, 0x0000ffff, 0xffff0000], np.uint64)
    vals = [(a, b) for a in special for b in special]
    for _ in range(n_random):
        pick = lambda: sum(int(words[rng.integers(0, len(words))]) << (32 * i) for i in range(8)) if rng.random() < 0.7 else int.from_bytes(rng.bytes(32), "little")
        vals.append((pick(), pick()))
    inp = np.frombuffer(b"".join(a.to_bytes(32, "little") + b.to_bytes(32, "little") for a, b in vals), np.uint8).reshape(-1, 64)
    for name, f in (("fe_mul", lambda a, b: a * b), ("fe_sq", lambda a, b: a * a), ("fe_add", lambda a, b: a + b), ("fe_sub", lambda a, b: a - b),
                    ("fe_chain", lambda a, b: ((a + b) * (a - b)) ** 2 * (a - b) + a)):
        out, _ = iss.selftest_primitive(name, inp)
        exp = np.frombuffer(b"".join((f(a, b) % P).to_bytes(32, "little") for a, b in vals), np.uint8).reshape(-1, 32)
        bad = np.nonzero((out != exp).any(axis=1))[0]
        assert len(bad) == 0, (name, [hex(v) for v in vals[bad[0]]])


def test_primitives_on_emulation(emu, coracle):
    from aeonflux_b200 import Issuer
    sp, ip, sk = coracle.make_issuer(1)
    check_primitives(Issuer(sp, ip, None, max_batch=4, _binding=emu), coracle)


@pytest.mark.parametrize("name", ["readme4", "s16", "scalar1", "quirk_sp_middle", "revealed10"])
def test_batchable_proofs_exact_and_rlc(emu, coracle, name):
    """BatchableProof presentations (commitments on the wire): the exact per-constraint check and the
    random-linear-combination / Pippenger check agree with the Python restatement of zkp's verify_batchable -- honest items,
    a corrupted commitment, a corrupted response, an identity commitment, an undecodable commitment."""
    from aeonflux_b200 import Issuer, PresentationBatch
    from oracle.pyoracle import aeonflux as A, flat as F, ristretto as R
    from tests.common import to_batchable
    g = load_golden(name)
    sp, ip, sk = bytes.fromhex(g["sysparams"]), bytes.fromhex(g["issuer_pub"]), bytes.fromhex(g["secret"])
    orc = coracle.Issuer(sp, ip, sk)
    rk = bytes(REQ[k] for k in g["request"])
    count = 6
    kinds, pres, _ = orc.synth(rk, g["hide"], g["config"].encode(), 0, count, want_issuances=False)
    ov, _, tr = orc.verify_presentations(kinds, pres, trace=True)
    assert not ov.any()
    bp = to_batchable(kinds, pres, tr["commitments"])
    iss = Issuer(sp, ip, sk, max_batch=4, _binding=emu)
    assert bp.shape[1] == iss._b.L.afx_batchable_num_fields(len(kinds), kinds)
    v, dbg = iss.verify_batchable(PresentationBatch.from_items(kinds, bp), debug=True)
    assert not v.any()
    assert (dbg["challenges"][0] == pres[:, 0]).all()                # the derived challenge is the compact proof's challenge
    vr, fell_back = iss.verify_batchable_rlc(PresentationBatch.from_items(kinds, bp), bytes(range(32)))
    assert not vr.any() and fell_back == 0
    nc = F.presentation_num_constraints(list(kinds))
    bad = bp.copy()
    bad[1, 0, 5] ^= 1                                                 # a commitment (still decodes or not: either way rejected)
    bad[2, nc, 0] ^= 1                                                # a response
    bad[3, 1] = 0                                                     # an identity commitment
    bad[4, 0] = 0xff                                                  # an undecodable commitment
    py = A.Issuer(A.SystemParameters.from_bytes(sp), A.IssuerParameters(R.decompress(ip[:32]), R.decompress(ip[32:])),
                  A.SecretKey(*[int.from_bytes(sk[4 + 32 * i:36 + 32 * i], "little") for i in range(4)],
                              [int.from_bytes(sk[132 + 32 * i:164 + 32 * i], "little") for i in range(g["n"])], R.decompress(sk[-32:])))
    expect = [F.verify_flat(py, kinds, [bad[i, w].tobytes() for w in range(bad.shape[1])], batchable=True)[0] for i in range(count)]
    assert expect == [0, 1, 1, 1, 1, 0]
    v = iss.verify_batchable(PresentationBatch.from_items(kinds, bad))
    assert list(v) == expect
    vr, fell_back = iss.verify_batchable_rlc(PresentationBatch.from_items(kinds, bad), bytes(range(32)))
    assert list(vr) == expect and fell_back == 2                      # both chunks of 4 and 2 items hold a bad item


def test_async_submit_wait(emu, coracle):
    """afx_*_submit / afx_wait: two submissions in flight, waited in and out of order; a third is refused until one is waited."""
    from aeonflux_b200 import Issuer, PresentationBatch
    from aeonflux_b200._binding import AfxError
    sp, ip, sk = coracle.make_issuer(4)
    orc = coracle.Issuer(sp, ip, sk)
    kinds, pres, issu = orc.synth(b"SSPE", [0, 3], b"async", 0, 12)
    pres[2, 1, 0] ^= 1; pres[9, 5, 3] ^= 2
    ov, _ = orc.verify_presentations(kinds, pres)
    iss = Issuer(sp, ip, sk, max_batch=6, _binding=emu)
    a = iss.submit(PresentationBatch.from_items(kinds, pres[:6]))
    b = iss.submit(PresentationBatch.from_items(kinds, pres[6:]))
    with pytest.raises(AfxError):
        iss.submit(PresentationBatch.from_items(kinds, pres[:1]))
    with pytest.raises(AfxError):
        iss.verify_batch(PresentationBatch.from_items(kinds, pres))          # a multi-pass call needs the same buffers
    assert (b.wait() == ov[6:]).all() and (a.wait() == ov[:6]).all()
    with pytest.raises(AfxError):
        a.wait()                                                             # a ticket is good for one wait
    c = iss.submit(PresentationBatch.from_items(bytes([0, 0, 2, 2]), issu[:5]), issuance=True)
    d = iss.submit(PresentationBatch.from_items(kinds, pres[:4]))
    assert not c.wait().any() and (d.wait() == ov[:4]).all()
    assert (iss.verify_batch(PresentationBatch.from_items(kinds, pres)) == ov).all()


def test_host_alloc_buffers(emu, coracle):
    """afx_host_alloc / afx_host_free: a batch built in library-allocated host memory verifies like any other; freeing NULL is a
    no-op and a NULL out pointer is an argument error."""
    import gc
    from aeonflux_b200 import Issuer, PresentationBatch
    sp, ip, sk = coracle.make_issuer(4)
    orc = coracle.Issuer(sp, ip, sk)
    kinds, pres, _ = orc.synth(b"SSPE", [0, 3], b"hostalloc", 0, 5)
    pres[3, 2, 9] ^= 8
    ov, _ = orc.verify_presentations(kinds, pres)
    iss = Issuer(sp, ip, sk, max_batch=4, _binding=emu)
    batch = PresentationBatch.from_items(kinds, pres, host_array=iss.host_array)
    assert batch.fields.shape == (28, 5, 32) and (batch.fields.transpose(1, 0, 2) == pres).all()
    assert (iss.verify_batch(batch) == ov).all() and list(ov) == [0, 0, 0, 1, 0]
    del batch
    gc.collect()                                     # runs afx_host_free through the finalizer
    assert emu.L.afx_host_alloc(None, 16) != 0
    emu.L.afx_host_free(None)
    z = iss.host_array((0, 32))
    assert z.shape == (0, 32)


def test_wide_constant_tables_on_emulation(emu, coracle, monkeypatch):
    """The radix-2^16 constant tables (built on the GPU for issuers with few generators; on the emulation only on request, they
    take seconds on one core): Issuer::verify and CredentialIssuance::verify give the oracle's Z, commitments, challenges and
    verdicts with them, as they do with the radix-4096 tables."""
    from aeonflux_b200 import Issuer, PresentationBatch
    from tests.common import compare_with_oracle_trace
    monkeypatch.setenv("AFX_HOSTEMU_CTAB16", "1")
    sp, ip, sk = coracle.make_issuer(4)
    orc = coracle.Issuer(sp, ip, sk)
    kinds, pres, issu = orc.synth(b"SSPE", [0, 3], b"wide-tables", 0, 9)
    pres[4, 2, 1] ^= 2; pres[7, 20, 30] ^= 0x40
    issu[3, 7, 0] ^= 1
    iss = Issuer(sp, ip, sk, max_batch=5, _binding=emu)
    monkeypatch.delenv("AFX_HOSTEMU_CTAB16")
    assert iss.launch_count == Issuer(sp, ip, sk, max_batch=5, _binding=emu).launch_count + 1      # the extra table-setup pass ran
    v, dbg = iss.verify_batch(PresentationBatch.from_items(kinds, pres), debug=True)
    ov, _, tr = orc.verify_presentations(kinds, pres, trace=True)
    compare_with_oracle_trace(v, dbg, ov, tr)
    assert list(ov) == [0, 0, 0, 0, 1, 0, 0, 1, 0]
    ik = bytes([0, 0, 2, 2])
    oi, _ = orc.verify_issuances(ik, issu)
    assert (iss.verify_issuance_batch(PresentationBatch.from_items(ik, issu)) == oi).all() and oi.sum() == 1


def check_noncanonical_wire_scalars(make_issuer, coracle, count=3):
    """Every scalar word of a presentation / issuance replaced by 256-bit values >= l whose top bits would, unmasked, drive the
    unbiased top digit of the constant-table recodings far out of bounds (response z = 0xffff0000... read 3 MB past the
    radix-2^16 table of I).  The items are rejected like the oracle rejects them, nothing faults, and the context keeps verifying."""
    from aeonflux_b200 import PresentationBatch
    from tests.common import issuance_scalar_words, noncanonical_scalar_batch, presentation_scalar_words
    sp, ip, sk = coracle.make_issuer(4)
    orc = coracle.Issuer(sp, ip, sk)
    kinds, pres, issu = orc.synth(b"SSPE", [0, 3], b"noncanonical", 0, count)
    iss = make_issuer(sp, ip, sk)
    rng = np.random.default_rng(5)
    bad = noncanonical_scalar_batch(pres, presentation_scalar_words(kinds), rng)
    assert len(bad) == 13 * 16
    ov, _ = orc.verify_presentations(kinds, bad)
    assert ov.all()
    assert (iss.verify_batch(PresentationBatch.from_items(kinds, bad)) == 1).all()
    assert (iss.verify_wire(kinds, bad) == 1).all()
    ik = bytes([0, 0, 2, 2])
    ibad = noncanonical_scalar_batch(issu, issuance_scalar_words(ik), rng)
    oi, _ = orc.verify_issuances(ik, ibad)
    assert oi.all()
    assert (iss.verify_issuance_batch(PresentationBatch.from_items(ik, ibad)) == 1).all()
    assert not iss.verify_batch(PresentationBatch.from_items(kinds, pres)).any()
    assert not iss.verify_issuance_batch(PresentationBatch.from_items(ik, issu)).any()


@pytest.mark.parametrize("wide", [False, True])
def test_noncanonical_wire_scalars_never_index_out_of_bounds(emu, coracle, monkeypatch, wide):
    from aeonflux_b200 import Issuer
    if wide:
        monkeypatch.setenv("AFX_HOSTEMU_CTAB16", "1")
    check_noncanonical_wire_scalars(lambda sp, ip, sk: Issuer(sp, ip, sk, max_batch=64, _binding=emu), coracle)


def test_submit_refuses_a_shape_the_inflight_workspace_cannot_hold(emu, coracle):
    """on a fresh context an issuance submit sizes the workspace for the issuance shape only (no aMAC tables,
    no flags, n + 2 ladder tables).  A presentation submit right behind it must be refused while that one is in flight -- not run
    the aMAC through unallocated arrays -- and succeed once the context is idle again."""
    from aeonflux_b200 import Issuer, PresentationBatch
    from aeonflux_b200._binding import AfxError
    sp, ip, sk = coracle.make_issuer(4)
    orc = coracle.Issuer(sp, ip, sk)
    kinds, pres, issu = orc.synth(b"SSPP", [], b"submit-ws", 0, 6)       # all revealed: 2n + 7 fields < the issuance's 2n + 9
    ik = bytes([0, 0, 2, 2])
    iss = Issuer(sp, ip, sk, max_batch=6, _binding=emu)
    a = iss.submit(PresentationBatch.from_items(ik, issu), issuance=True)
    with pytest.raises(AfxError):
        iss.submit(PresentationBatch.from_items(kinds, pres))
    assert not a.wait().any()
    b = iss.submit(PresentationBatch.from_items(kinds, pres))            # idle: the workspace grows
    c = iss.submit(PresentationBatch.from_items(ik, issu), issuance=True)  # fits what the presentation sized
    assert not b.wait().any() and not c.wait().any()


def issuance_to_batchable(issu, commitments):
    """Compact issuances [count][2n+9][32] + the three commitments their proofs recompute [count][3][32] -> the BatchableProof layout
    [count][2n+11][32]: attribute[n], t, U, V, commitments[3], responses[n+5]."""
    n = (issu.shape[1] - 9) // 2
    return np.ascontiguousarray(np.concatenate([issu[:, :n + 3], commitments, issu[:, n + 4:]], axis=1))


def check_issuance_batchable(make_issuer, coracle, count, max_batch):
    """BatchableProof issuances (issuance.rs:21-22's commented-out BatchVerifier): the exact per-constraint check and the random
    linear combination agree with the Python restatement of zkp's verify_batchable -- honest items, corrupted commitment, response,
    attribute, identity commitment, undecodable commitment."""
    from aeonflux_b200 import PresentationBatch
    from oracle.pyoracle import aeonflux as A, flat as F
    sp, ip, sk = coracle.make_issuer(4)
    orc = coracle.Issuer(sp, ip, sk)
    _, _, issu = orc.synth(b"SSPE", [], b"iss-batchable", 0, count)
    ik = bytes([0, 0, 2, 2])
    ov, _, tr = orc.verify_issuances(ik, issu, trace=True)
    assert not ov.any()
    bi = issuance_to_batchable(issu, tr["commitments"])
    iss = make_issuer(sp, ip, sk, max_batch)
    assert bi.shape[1] == iss._b.L.afx_issuance_batchable_num_fields(4) == 19
    v, dbg = iss.verify_issuance_batchable(PresentationBatch.from_items(ik, bi), debug=True)
    assert not v.any()
    assert (dbg["challenges"][0] == issu[:, 7]).all()                 # the derived challenge is the compact proof's
    p0, e0 = iss.rlc_stats()
    vr, fb = iss.verify_batchable_rlc(PresentationBatch.from_items(ik, bi), bytes(32), issuance=True)
    assert not vr.any() and fb == 0
    assert iss.rlc_stats() == (p0 + -(-count // max_batch), e0)      # one pass per chunk, nothing re-verified
    bad = bi.copy()
    bad[1, 7, 5] ^= 1          # commitment C_W
    bad[2, 10, 0] ^= 1         # response w
    bad[3, 0, 0] ^= 1          # a scalar attribute
    bad[4, 8] = 0              # an identity commitment
    bad[5, 9] = 0xff           # an undecodable commitment
    sysp = A.SystemParameters.from_bytes(sp)
    from oracle.pyoracle import ristretto as R
    ipp = A.IssuerParameters(R.decompress(ip[:32]), R.decompress(ip[32:]))
    expect = [F.verify_issuance_flat(sysp, ipp, ik, [bad[i, w].tobytes() for w in range(19)], batchable=True)[0] for i in range(min(count, 8))]
    assert expect[:7] == [0, 1, 1, 1, 1, 1, 0]
    full = np.zeros(count, np.uint8); full[1:6] = 1
    assert (iss.verify_issuance_batchable(PresentationBatch.from_items(ik, bad)) == full).all()
    vr, fb = iss.verify_batchable_rlc(PresentationBatch.from_items(ik, bad), bytes(32), issuance=True)
    assert (vr == full).all() and fb >= 1
    user = make_issuer(sp, ip, None, max_batch)                       # needs no secret key
    assert (user.verify_batchable_rlc(PresentationBatch.from_items(ik, bad), bytes(range(32)), issuance=True)[0] == full).all()


def test_issuance_batchable_exact_and_rlc(emu, coracle):
    from aeonflux_b200 import Issuer
    check_issuance_batchable(lambda sp, ip, sk, mb: Issuer(sp, ip, sk, max_batch=mb, _binding=emu), coracle, count=9, max_batch=7)


def check_rlc_bisection(make_issuer, coracle, count, max_batch, leaf, n_bad, monkeypatch):
    """A chunk whose combination does not vanish is bisected: with a few bad items only the leaves that hold them are re-verified
    exactly, the verdicts are exact, and the number of passes / re-verified items stays far below the chunk."""
    from aeonflux_b200 import PresentationBatch
    from tests.common import to_batchable
    monkeypatch.setenv("AFX_RLC_LEAF", str(leaf))
    sp, ip, sk = coracle.make_issuer(4)
    orc = coracle.Issuer(sp, ip, sk)
    kinds, pres, _ = orc.synth(b"SSPE", [0, 3], b"bisect", 0, min(count, 64), want_issuances=False)
    ov, _, tr = orc.verify_presentations(kinds, pres, trace=True)
    bp = to_batchable(kinds, pres, tr["commitments"])
    rng = np.random.default_rng(8)
    items = bp[rng.integers(0, len(bp), count)].copy()
    bad = np.sort(rng.choice(count, n_bad, replace=False))
    for i in bad:
        items[i, rng.integers(0, items.shape[1]), rng.integers(0, 32)] ^= 1 << rng.integers(0, 8)
    expect = np.zeros(count, np.uint8); expect[bad] = 1
    iss = make_issuer(sp, ip, sk, max_batch)
    p0, e0 = iss.rlc_stats()
    v, fb = iss.verify_batchable_rlc(PresentationBatch.from_items(kinds, items), bytes(32))
    p1, e1 = iss.rlc_stats()
    assert (v == expect).all()
    chunks = -(-count // max_batch)
    assert fb == len({int(i) // max_batch for i in bad})
    assert 0 < e1 - e0 <= n_bad * leaf, (e1 - e0, n_bad, leaf)           # only the suspect leaves were re-verified
    depth = int(np.ceil(np.log2(max_batch / leaf)))
    assert chunks < p1 - p0 <= chunks + n_bad * 2 * (depth + 1), (p1 - p0, depth)
    # many bad items: bisection gives up and the whole chunk is checked exactly -- same verdicts
    for i in range(0, count, 3):
        items[i, 1, 0] ^= 1
        expect[i] = 1
    v, fb = iss.verify_batchable_rlc(PresentationBatch.from_items(kinds, items), bytes(32))
    assert (v == expect).all() and fb == chunks


def test_rlc_bisection_on_emulation(emu, coracle, monkeypatch):
    from aeonflux_b200 import Issuer
    check_rlc_bisection(lambda sp, ip, sk, mb: Issuer(sp, ip, sk, max_batch=mb, _binding=emu), coracle, count=70, max_batch=64, leaf=4, n_bad=2, monkeypatch=monkeypatch)


def check_wide_table_allocation_failure(make_issuer, coracle, monkeypatch):
    """afx_ctx_create when the radix-2^16 constant tables cannot be allocated (the fallback was exercised by no test):
    the context falls back to the radix-4096 tables -- one table-setup launch fewer -- and verifies like any other."""
    from aeonflux_b200 import PresentationBatch
    sp, ip, sk = coracle.make_issuer(4)
    orc = coracle.Issuer(sp, ip, sk)
    kinds, pres, issu = orc.synth(b"SSPE", [0, 3], b"ctab16-fail", 0, 12)
    pres[3, 1, 31] ^= 0x10; issu[5, 9, 0] ^= 1
    wide = make_issuer(sp, ip, sk)
    monkeypatch.setenv("AFX_TEST_FAIL_CTAB16", "1")
    narrow = make_issuer(sp, ip, sk)
    monkeypatch.delenv("AFX_TEST_FAIL_CTAB16")
    assert narrow.launch_count == wide.launch_count - 1
    ov, _ = orc.verify_presentations(kinds, pres)
    oi, _ = orc.verify_issuances(bytes([0, 0, 2, 2]), issu)
    for iss in (wide, narrow):
        assert (iss.verify_batch(PresentationBatch.from_items(kinds, pres)) == ov).all() and ov.sum() == 1
        assert (iss.verify_issuance_batch(PresentationBatch.from_items(bytes([0, 0, 2, 2]), issu)) == oi).all() and oi.sum() == 1


def test_wide_table_allocation_failure_falls_back(emu, coracle, monkeypatch):
    from aeonflux_b200 import Issuer
    monkeypatch.setenv("AFX_HOSTEMU_CTAB16", "1")
    check_wide_table_allocation_failure(lambda sp, ip, sk: Issuer(sp, ip, sk, max_batch=16, _binding=emu), coracle, monkeypatch)


def linked_case(n, request, hide, count, tag, linked=True):
    """Python-oracle material for the linked (DLEQ) presentation tests: issuer bytes, kinds, the struct-of-arrays show input
    [n_show_fields][count][32], the linked presentations the oracle's prover makes from the same rng bytes [count][W][32], and for every
    item a SPLICED copy: the same credential proof with the proof of encryption of another plaintext (same z) attached."""
    from oracle.pyoracle import aeonflux as A, flat as F, ristretto as R, synth as S
    iss = S.make_issuer(n)
    sp, ip = iss.system_parameters, iss.issuer_parameters
    rng = A.ShakeRng(b"linked/" + tag)
    kinds = None
    rows, pres, spliced = [], [], []
    for _ in range(count):
        attrs = []
        for k in request:
            attrs.append(("PS", rng.scalar()) if k == "PS" else ("PP", rng.point()) if k == "PP" else ("EP", A.Plaintext.from_bytes30(rng.fill(30))))
        _, (amac, _) = iss.issue(list(attrs), rng)
        kp, _ = A.SymmetricKeypair.generate(sp, rng)
        shown = list(attrs)
        for i in hide:
            A.hide_attribute(shown, i)
        h_s, h_p = sum(k == "SS" for k, _ in shown), sum(k == "SP" for k, _ in shown)
        draws = [rng.fill(64) for _ in range(1 + 3 + h_s + 6 * h_p)]
        sc = [R.sc_from_wide(d) for d in draws]
        p = A.presentation_prove(sp, ip, amac, shown, kp if h_p else None, sc[0], sc[1:4 + h_s], [sc[4 + h_s + 6 * e:10 + h_s + 6 * e] for e in range(h_p)], linked=linked)
        kinds = F.presentation_kinds(p)
        pres.append(F.presentation_to_words(p))
        w = [R.sc_to_bytes(amac.t), amac.U.compress(), amac.V.compress()]
        for k, v in shown:
            w += [R.sc_to_bytes(v)] if k in ("PS", "SS") else [v.compress()] if k == "PP" else [v.M1.compress()] if k == "EP" else [v.M1.compress(), v.M2.compress(), R.sc_to_bytes(v.m3)]
        if h_p:
            w += [R.sc_to_bytes(kp.a), R.sc_to_bytes(kp.a0), R.sc_to_bytes(kp.a1), kp.pk.compress()]
        for d in draws:
            w += [d[:32], d[32:]]
        rows.append(w)
        if h_p:
            other = A.Plaintext.from_bytes30(rng.fill(30))
            i = [j for j, (k, _) in enumerate(shown) if k == "SP"][0]
            p.proofs_of_encryption[0] = (i, A.encryption_prove(sp, other, i, kp, sc[0], [rng.scalar() for _ in range(6)]))
            spliced.append(F.presentation_to_words(p))
    tow = lambda items: np.frombuffer(b"".join(b"".join(ws) for ws in items), np.uint8).reshape(len(items), -1, 32).copy()
    show_in = np.ascontiguousarray(tow(rows).transpose(1, 0, 2))
    return iss, bytes(kinds), show_in, tow(pres), (tow(spliced) if spliced else None)


def check_linked_presentations(make_issuer, n, request, hide, count, tag):
    """The linked statement (README.md:119-122's TODO, opt-in): afx_show_linked is byte-identical to the oracle's linked prover,
    afx_verify_presentations_linked accepts it with the oracle's Z / commitments / challenges, the reference's verifier and the linked
    one reject each other's proofs (different transcripts) -- and a credential proof with the encryption of ANOTHER plaintext attached,
    which Issuer::verify accepts, is rejected by the linked verifier."""
    from aeonflux_b200 import PresentationBatch
    from oracle.pyoracle import flat as F
    from tests.common import compare_with_oracle_trace
    iss, kinds, show_in, pres, spliced = linked_case(n, request, hide, count, tag)
    sp, ip, sk = iss.system_parameters.to_bytes(), iss.issuer_parameters.to_bytes(), iss.amacs_key.to_bytes()
    eng = make_issuer(sp, ip, sk)
    user = make_issuer(sp, ip, None)
    res, st, sdbg = user.show_batch(kinds, show_in, debug=True, linked=True)
    assert not st.any()
    got = res.fields.transpose(1, 0, 2)
    assert (got == pres).all(), "afx_show_linked differs from the oracle's linked prover"
    wgot, wst = user.show_wire(kinds, np.ascontiguousarray(show_in.transpose(1, 0, 2)), linked=True)
    assert (wgot == pres).all() and not wst.any()
    has_link = any(k == 3 and i > 0 for i, k in enumerate(kinds))
    # linked verifier vs oracle: verdicts, Z, every commitment (incl. the link constraints), challenges
    pres[1 % count, 1, 0] ^= 1
    items = np.concatenate([pres] + ([spliced] if spliced is not None else []))
    overd, otr = zip(*[F.verify_flat(iss, kinds, [items[i, w].tobytes() for w in range(items.shape[1])], linked=True) for i in range(len(items))])
    ncm = eng._b.L.afx_presentation_linked_num_commitments(len(kinds), kinds)
    tr = {"Z": np.zeros((len(items), 32), np.uint8), "commitments": np.zeros((len(items), ncm, 32), np.uint8),
          "challenges": np.zeros((len(items), eng.num_proofs(kinds), 32), np.uint8)}
    for i, t in enumerate(otr):
        if t["Z"]:
            tr["Z"][i] = np.frombuffer(t["Z"], np.uint8)
        for k, cm in enumerate(t["commitments"]):
            tr["commitments"][i, k] = np.frombuffer(cm, np.uint8)
        for k, ch in enumerate(t["challenges"]):
            tr["challenges"][i, k] = np.frombuffer(ch, np.uint8)
    v, dbg = eng.verify_batch(PresentationBatch.from_items(kinds, items), debug=True, linked=True)
    # the flat trace lists the main proof's commitments then each enc proof's; the engine's dump has the same order
    compare_with_oracle_trace(v, dbg, np.array(overd, np.uint8), tr)
    expect = np.zeros(len(items), np.uint8); expect[1 % count] = 1
    if spliced is not None:
        expect[count:] = 1 if has_link or kinds[0] == 3 else 0
    assert (v == expect).all(), (list(v), list(expect))
    assert (eng.verify_wire(kinds, items, linked=True) == expect).all()
    # the reference's verifier: accepts the spliced proofs when they were made by ITS prover -- shown here on the linked prover's
    # output only where the two statements coincide (no link constraint); otherwise the transcripts differ and it rejects
    ref = eng.verify_batch(PresentationBatch.from_items(kinds, pres))
    if has_link:
        assert ref.all()
    else:
        assert list(ref) == list(expect[:count])


@pytest.mark.parametrize("n,request_,hide,tag", [(4, ("PS", "PS", "PP", "EP"), (0, 3), b"readme4"), (1, ("EP",), (0,), b"plain1"),
                                                 (3, ("PS", "EP", "PS"), (1,), b"middle"), (5, ("PS", "PP", "EP", "EP", "EP"), (0, 2, 3, 4), b"three")])
def test_linked_presentations_on_emulation(emu, n, request_, hide, tag):
    from aeonflux_b200 import Issuer
    check_linked_presentations(lambda sp, ip, sk: Issuer(sp, ip, sk, max_batch=3, _binding=emu), n, request_, hide, 4, tag)


def check_reference_accepts_spliced_encryption(make_issuer):
    """Why the link exists: under the REFERENCE's statement a valid credential proof with the proof of encryption of another
    plaintext attached is accepted (presentation.rs:292 TODO) -- by the oracle and, bit for bit, by the engine's Issuer::verify."""
    from aeonflux_b200 import PresentationBatch
    from oracle.pyoracle import flat as F
    iss, kinds, _, pres, spliced = linked_case(4, ("PS", "PS", "PP", "EP"), (0, 3), 3, b"unlinked", linked=False)
    sp, ip, sk = iss.system_parameters.to_bytes(), iss.issuer_parameters.to_bytes(), iss.amacs_key.to_bytes()
    eng = make_issuer(sp, ip, sk)
    assert [F.verify_flat(iss, kinds, [spliced[i, w].tobytes() for w in range(spliced.shape[1])])[0] for i in range(3)] == [0, 0, 0]
    assert not eng.verify_batch(PresentationBatch.from_items(kinds, spliced)).any()
    assert eng.verify_batch(PresentationBatch.from_items(kinds, spliced), linked=True).all()


def test_reference_accepts_spliced_encryption_on_emulation(emu):
    from aeonflux_b200 import Issuer
    check_reference_accepts_spliced_encryption(lambda sp, ip, sk: Issuer(sp, ip, sk, max_batch=3, _binding=emu))


def test_single_pass_wire_call_copies_in_halves(emu, coracle):
    """A one-pass item-major host call copies its batch in two halves (the first half's early stages under the second half's copy):
    same verdicts as the oracle's, odd counts and counts below the split threshold included."""
    from aeonflux_b200 import Issuer
    sp, ip, sk = coracle.make_issuer(4)
    orc = coracle.Issuer(sp, ip, sk)
    kinds, pres, issu = orc.synth(b"SSPE", [0, 3], b"halves", 0, 9)
    pres[0, 1, 0] ^= 1; pres[4, 9, 31] ^= 0x80; pres[8, 20, 5] ^= 2; issu[5, 8, 0] ^= 1
    ov, _ = orc.verify_presentations(kinds, pres)
    oi, _ = orc.verify_issuances(bytes([0, 0, 2, 2]), issu)
    iss = Issuer(sp, ip, sk, max_batch=16, _binding=emu)
    for m in (9, 8, 3, 2, 1):
        assert (iss.verify_wire(kinds, pres[:m]) == ov[:m]).all(), m
    assert (iss.verify_wire(bytes([0, 0, 2, 2]), issu, issuance=True) == oi).all() and list(ov) == [1, 0, 0, 0, 1, 0, 0, 0, 1]


def check_split_amac(make_issuer, coracle, monkeypatch, count=7):
    """Small batches cut the aMAC ladder into parts that run as separate jobs (api_impl.inc:amac_split_terms); whatever the cut -- none,
    one term per part, two, five, the automatic choice -- Z, every commitment, every challenge and every verdict are the oracle's."""
    from aeonflux_b200 import PresentationBatch
    for n, rk, hide in ((4, b"SSPE", [0, 3]), (16, b"SSSSSSPP" + b"E" * 8, [0, 1] + list(range(8, 16))), (1, b"S", [])):
        sp, ip, sk = coracle.make_issuer(n)
        orc = coracle.Issuer(sp, ip, sk)
        kinds, pres, _ = orc.synth(rk, hide, b"split-amac", 0, count, want_issuances=False)
        pres[2, 1, 3] ^= 8
        pres[4, 5 + sum(k == 1 for k in kinds), 0] ^= 2          # C_x_1: a base of the aMAC ladder
        ov, _, tr = orc.verify_presentations(kinds, pres, trace=True)
        assert ov[2] == 1 and ov[4] == 1
        launches = {}
        for mode in ("0", "1", "2", "5", None):
            if mode is None:
                monkeypatch.delenv("AFX_AMAC_SPLIT", raising=False)
            else:
                monkeypatch.setenv("AFX_AMAC_SPLIT", mode)
            iss = make_issuer(sp, ip, sk)
            v, dbg = iss.verify_batch(PresentationBatch.from_items(kinds, pres), debug=True)
            compare_with_oracle_trace(v, dbg, ov, tr)
            assert (iss.verify_wire(kinds, pres) == ov).all()
            launches[mode] = iss.launch_count
            iss.close()
        if n > 1:
            assert launches["1"] > launches["0"] and launches[None] > launches["0"]      # the split path ran (combine + "Z" launches)
    monkeypatch.delenv("AFX_AMAC_SPLIT", raising=False)


def test_split_amac_ladder_on_emulation(emu, coracle, monkeypatch):
    from aeonflux_b200 import Issuer
    check_split_amac(lambda sp, ip, sk: Issuer(sp, ip, sk, max_batch=16, _binding=emu), coracle, monkeypatch)
