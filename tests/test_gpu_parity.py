"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI, against the oracle -- bit-exact compressed
points (Z, every recomputed commitment), challenge scalars and verdicts, including deliberately corrupted proofs."""
import numpy as np
import pytest

from tests.common import GOLDEN_SHAPES, REQ, compare_with_oracle_trace, corrupt_batch, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def readme4(coracle):
    from aeonflux_b200 import Issuer
    sp, ip, sk = coracle.make_issuer(4)
    return coracle.Issuer(sp, ip, sk), Issuer(sp, ip, sk, device=0, max_batch=8192), (sp, ip, sk)


def test_cuda_library_is_the_path_under_test():
    from aeonflux_b200._lib import load
    assert "cuda sm_100a" in load().version()


@pytest.mark.parametrize("name", GOLDEN_SHAPES)
def test_golden_shapes(coracle, name):
    from aeonflux_b200 import Issuer, PresentationBatch
    g = load_golden(name)
    sp, ip, sk = bytes.fromhex(g["sysparams"]), bytes.fromhex(g["issuer_pub"]), bytes.fromhex(g["secret"])
    orc = coracle.Issuer(sp, ip, sk)
    rk = bytes(REQ[k] for k in g["request"])
    count = 70                                                  # ragged: not a multiple of the 128-thread CTA
    kinds, pres, issu = orc.synth(rk, g["hide"], g["config"].encode(), 0, count)
    pres[1, 1, 3] ^= 0x10
    pres[3, 0, 0] ^= 1
    iss = Issuer(sp, ip, sk, device=0, max_batch=64)           # 64 < 70: two chunks
    v, dbg = iss.verify_batch(PresentationBatch.from_items(kinds, pres), debug=True)
    ov, _, tr = orc.verify_presentations(kinds, pres, trace=True)
    compare_with_oracle_trace(v, dbg, ov, tr)
    e = g["items"][0]                                           # committed fixture from the Python big-int oracle
    assert dbg["Z"][0].tobytes().hex() == e["Z"]
    assert dbg["commitments"][:len(e["commitments"]), 0].tobytes().hex() == "".join(e["commitments"])   # filled as far as the reference gets
    assert dbg["challenges"][:len(e["challenges"]), 0].tobytes().hex() == "".join(e["challenges"])
    assert v[0] == e["verdict"]
    for c in e["corrupted"]:
        from tests.common import words
        w = words(c["words"])[None]
        v1, d1 = iss.verify_batch(PresentationBatch.from_items(kinds, w), debug=True)
        assert v1[0] == c["verdict"] == 1, c["class"]
        ncm = len(c["commitments"])
        assert d1["commitments"][:ncm, 0].tobytes().hex() == "".join(c["commitments"]), c["class"]
    ik = bytes(e["issuance_kinds"])
    issu[2, len(ik) + 1, 5] ^= 2
    vi, dbgi = iss.verify_issuance_batch(PresentationBatch.from_items(ik, issu), debug=True)
    ovi, _, tri = orc.verify_issuances(ik, issu, trace=True)
    compare_with_oracle_trace(vi, dbgi, ovi, tri)
    assert dbgi["commitments"][:, 0].tobytes().hex() == "".join(e["issuance_commitments"])


def test_readme4_batch_with_corruptions(readme4):
    from aeonflux_b200 import PresentationBatch
    orc, iss, _ = readme4
    count = 4096 + 37
    kinds, pres, _ = orc.synth(b"SSPE", [0, 3], b"readme4", 1000, count, want_issuances=False)
    rng = np.random.default_rng(1)
    pts = pres[:64, 5:8].reshape(-1, 32).copy()
    idx = corrupt_batch(pres, kinds, rng, 0.05, pts)
    v, dbg = iss.verify_batch(PresentationBatch.from_items(kinds, pres), debug=True)
    ov, _, tr = orc.verify_presentations(kinds, pres, trace=True)
    compare_with_oracle_trace(v, dbg, ov, tr)
    assert ov[idx].all() and int(ov.sum()) == len(idx)


def test_s16_batch_with_corruptions(coracle):
    from aeonflux_b200 import Issuer, PresentationBatch
    sp, ip, sk = coracle.make_issuer(16)
    orc = coracle.Issuer(sp, ip, sk)
    count = 300
    kinds, pres, issu = orc.synth(b"SSSSSSPP" + b"E" * 8, [0, 1] + list(range(8, 16)), b"s16", 0, count)
    rng = np.random.default_rng(2)
    pts = pres[:8, 6:9].reshape(-1, 32).copy()
    idx = corrupt_batch(pres, kinds, rng, 0.1, pts)
    iss = Issuer(sp, ip, sk, device=0, max_batch=512)
    v, dbg = iss.verify_batch(PresentationBatch.from_items(kinds, pres), debug=True)
    ov, _, tr = orc.verify_presentations(kinds, pres, trace=True)
    compare_with_oracle_trace(v, dbg, ov, tr)
    assert int(ov.sum()) == len(idx)
    ik = bytes([0] * 6 + [2] * 10)
    vi, dbgi = iss.verify_issuance_batch(PresentationBatch.from_items(ik, issu), debug=True)
    ovi, _, tri = orc.verify_issuances(ik, issu, trace=True)
    compare_with_oracle_trace(vi, dbgi, ovi, tri)
    assert not ovi.any()


def test_edges(readme4):
    from aeonflux_b200 import Issuer, PresentationBatch
    from aeonflux_b200._binding import AfxError
    orc, iss, (sp, ip, sk) = readme4
    kinds, pres, issu = orc.synth(b"SSPE", [0, 3], b"edges", 0, 3)
    assert len(iss.verify_batch(PresentationBatch(kinds, np.zeros((28, 0, 32), np.uint8)))) == 0
    one = iss.verify_batch(PresentationBatch.from_items(kinds, pres[:1]))
    assert list(one) == [0]
    zeros = iss.verify_batch(PresentationBatch(kinds, np.zeros((28, 5, 32), np.uint8)))     # all-identity / all-zero input
    assert zeros.all()
    ones = iss.verify_batch(PresentationBatch(kinds, np.full((28, 5, 32), 0xff, np.uint8)))  # nothing decodes
    assert ones.all()
    with pytest.raises(AfxError):
        iss.verify_batch(PresentationBatch(kinds, np.zeros((27, 2, 32), np.uint8)))
    user = Issuer(sp, ip, None, device=0, max_batch=16)
    with pytest.raises(AfxError):
        user.verify_batch(PresentationBatch.from_items(kinds, pres))
    assert not user.verify_issuance_batch(PresentationBatch.from_items(bytes([0, 0, 2, 2]), issu)).any()


def test_device_resident_entry_point(readme4):
    """afx_verify_presentations_device: inputs already in HBM, enqueued on the caller's stream (what bench.py times)."""
    import torch
    orc, iss, _ = readme4
    count = 1000
    kinds, pres, _ = orc.synth(b"SSPE", [0, 3], b"devapi", 0, count, want_issuances=False)
    pres[17, 0, 0] ^= 1
    fields = torch.from_numpy(np.ascontiguousarray(pres.transpose(1, 0, 2))).cuda()
    verdicts = torch.full((count,), 7, dtype=torch.uint8, device="cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        iss.verify_batch_device(kinds, count, fields.data_ptr(), verdicts.data_ptr(), s.cuda_stream)
    s.synchronize()
    ov, _ = orc.verify_presentations(kinds, pres)
    assert (verdicts.cpu().numpy() == ov).all() and ov[17] == 1 and ov.sum() == 1


def test_full_size_batch_round_trip_properties(readme4):
    """BASELINE config 2 size through size-independent properties: 65,536 DISTINCT README-4 presentations (issued and shown on
    the device from random attributes, bench.synthesize_on_device) are all accepted, every item with one flipped bit -- any
    word, any byte including byte 31 with bit 255 and the scalars' top bits -- is rejected, and nothing else changes
    (linearity of the corruption set); a sample is verified by the oracle."""
    import torch
    from aeonflux_b200 import Issuer, PresentationBatch
    from bench import KINDS_README4, synthesize_on_device
    orc, _, (sp, ip, sk) = readme4
    kinds = KINDS_README4
    iss = Issuer(sp, ip, sk, device=0, max_batch=65536)
    dev = synthesize_on_device(torch, iss, 65536, 77, torch.cuda.current_stream())      # bench_data's issuer and keypair are make_issuer(4)'s
    big = np.ascontiguousarray(dev.cpu().numpy().transpose(1, 0, 2))
    assert len(np.unique(big[:, 4:7].reshape(65536, -1), axis=0)) == 65536          # distinct commitments per item
    assert not iss.verify_wire(kinds, big).any()
    rng = np.random.default_rng(3)
    bad = rng.choice(65536, 655, replace=False)
    for j, i in enumerate(bad):
        w = rng.integers(0, 28)
        big[i, w, 31 if j % 4 == 0 else rng.integers(0, 32)] ^= 1 << rng.integers(0, 8)
    v = iss.verify_batch(PresentationBatch.from_items(kinds, big))
    expect = np.zeros(65536, np.uint8)
    expect[bad] = 1
    assert (v == expect).all()
    sample = np.concatenate([bad[:96], rng.choice(65536, 64, replace=False)])
    ov, _ = orc.verify_presentations(kinds, np.ascontiguousarray(big[sample]))
    assert (v[sample] == ov).all()


def test_s16_full_size_round_trip_properties(coracle):
    """BASELINE config 4 size (65,536 x S16: 16 attributes, 8 hidden plaintexts, 4,576 bytes per item) through the same properties:
    65,536 distinct presentations issued and shown on the device are all accepted, 500 items with one flipped bit anywhere in their
    143 words are rejected and nothing else is, a sample (honest and corrupted) gets the oracle's verdicts."""
    import torch
    from aeonflux_b200 import Issuer
    from bench import KINDS_S16, load_issuer, synthesize_on_device
    sp, ip, sk = load_issuer("issuer16.bin")
    orc = coracle.Issuer(sp, ip, sk)
    iss = Issuer(sp, ip, sk, device=0, max_batch=65536)
    dev = synthesize_on_device(torch, iss, 65536, 78, torch.cuda.current_stream(), KINDS_S16, "keypair16.bin")
    big = np.ascontiguousarray(dev.cpu().numpy().transpose(1, 0, 2))
    del dev
    assert big.shape == (65536, 143, 32)
    assert not iss.verify_wire(KINDS_S16, big).any()
    rng = np.random.default_rng(4)
    bad = rng.choice(65536, 500, replace=False)
    for j, i in enumerate(bad):
        big[i, rng.integers(0, 143), 31 if j % 4 == 0 else rng.integers(0, 32)] ^= 1 << rng.integers(0, 8)
    v = iss.verify_wire(KINDS_S16, big)
    expect = np.zeros(65536, np.uint8)
    expect[bad] = 1
    assert (v == expect).all()
    sample = np.concatenate([bad[:48], rng.choice(65536, 16, replace=False)])
    ov, _ = orc.verify_presentations(KINDS_S16, np.ascontiguousarray(big[sample]))
    assert (v[sample] == ov).all()
    iss.close()


@pytest.mark.parametrize("name", GOLDEN_SHAPES)
def test_issue_golden_shapes(coracle, name):
    """Issuer::issue (issuer.rs:111-124) on the GPU with supplied rng output: the committed Python-oracle fixture, then a
    ragged batch against the C oracle, including malformed requests; every issued credential verifies."""
    from aeonflux_b200 import Issuer, RequestBatch
    from tests.common import golden_issue_request
    g = load_golden(name)
    sp, ip, sk = bytes.fromhex(g["sysparams"]), bytes.fromhex(g["issuer_pub"]), bytes.fromhex(g["secret"])
    n = g["n"]
    iss = Issuer(sp, ip, sk, device=0, max_batch=64)
    e = g["items"][0]
    ik = bytes(e["issuance_kinds"])
    attrs, rnd = golden_issue_request(g, e)
    res, st, dbg = iss.issue_batch(RequestBatch.from_request(ik, attrs[None], rnd[None]), debug=True)
    assert st[0] == 0
    assert res.fields[:, 0].tobytes().hex() == "".join(e["issuance_words"])
    assert dbg["commitments"][:, 0].tobytes().hex() == "".join(e["issuance_commitments"])
    orc = coracle.Issuer(sp, ip, sk)
    rk = bytes(REQ[k] for k in g["request"])
    count = 70
    _, _, issu = orc.synth(rk, [], g["config"].encode() + b"-issue", 0, count)
    A = np.ascontiguousarray(issu[:, :n])
    R = np.random.default_rng(21).integers(0, 256, (count, n + 7, 64), dtype=np.uint8)
    A[5, 0] = 0xff
    A[33, n - 1, 0] ^= 1
    out, status, _ = orc.issue(ik, A, R)
    res, st = iss.issue_batch(RequestBatch.from_request(ik, A, R))
    assert (st == status).all() and status[5] == 1
    assert (res.fields.transpose(1, 0, 2)[:, n:] == out).all()
    v = iss.verify_issuance_batch(res)
    ov, _ = orc.verify_issuances(ik, np.ascontiguousarray(res.fields.transpose(1, 0, 2)))
    assert (v == ov).all() and (v[status == 1] == 1).all()
    # item-major requests in, item-major issuances out (afx_issue_wire; 70 items over max_batch 64 = two passes)
    wout, wst = iss.issue_wire(ik, np.ascontiguousarray(RequestBatch.from_request(ik, A, R).fields.transpose(1, 0, 2)))
    good = status == 0
    assert (wst == status).all() and (wout[good] == res.fields.transpose(1, 0, 2)[good]).all() and not wout[~good].any()
    assert (iss.verify_wire(ik, wout, issuance=True) == ov).all()


def test_issue_full_size_round_trip(readme4):
    """BASELINE config 3 size: 65,536 revealed 4-attribute requests through issue -> CredentialIssuance::verify on the
    device; every honest issuance verifies, a sample is byte-identical to the oracle, and flipping a bit of an issued
    word makes exactly that item fail."""
    from aeonflux_b200 import Issuer, RequestBatch
    orc, _, (sp, ip, sk) = readme4
    count = 65536
    kinds = bytes([0, 0, 2, 2])
    _, _, issu = orc.synth(b"SSPP", [], b"issue-full", 0, 2048)
    A = np.tile(np.ascontiguousarray(issu[:, :4]), (count // 2048, 1, 1))
    R = np.random.default_rng(22).integers(0, 256, (count, 11, 64), dtype=np.uint8)
    iss = Issuer(sp, ip, sk, device=0, max_batch=count)
    res, st = iss.issue_batch(RequestBatch.from_request(kinds, A, R))
    assert not st.any()
    sample = np.random.default_rng(23).choice(count, 128, replace=False)
    out, status, _ = orc.issue(kinds, np.ascontiguousarray(A[sample]), np.ascontiguousarray(R[sample]))
    assert (res.fields.transpose(1, 0, 2)[sample, 4:] == out).all()
    bad = np.random.default_rng(24).choice(count, 300, replace=False)
    rr = np.random.default_rng(25)
    for i in bad:
        w = 4 + rr.integers(0, 13)
        res.fields[w, i, rr.integers(0, 32)] ^= 1 << rr.integers(0, 8)
    v = iss.verify_issuance_batch(res)
    expect = np.zeros(count, np.uint8)
    expect[bad] = 1
    assert (v == expect).all()


def test_mixed_stream_reduced_config5(coracle):
    """BASELINE configs[4] at reduced size through the sharding layer (world 1): README-4 and S16 presentations interleaved
    item by item, 1 % corrupted; verdicts come back in stream order and match the corrupted set and the oracle."""
    from aeonflux_b200 import Issuer
    from aeonflux_b200.shard import ShardedIssuer
    iss, orc, base = {}, {}, {}
    for n, rk, hide, cfg in ((4, b"SSPE", [0, 3], b"mix4"), (16, b"SSSSSSPP" + b"E" * 8, [0, 1] + list(range(8, 16)), b"mix16")):
        sp, ip, sk = coracle.make_issuer(n)
        orc[n] = coracle.Issuer(sp, ip, sk)
        iss[n] = Issuer(sp, ip, sk, device=0, max_batch=2048)
        kinds, pres, _ = orc[n].synth(rk, hide, cfg, 0, 64, want_issuances=False)
        base[n] = (kinds, pres)
    rng = np.random.default_rng(9)
    total = 6000
    kl, items, expect = [], [], np.zeros(total, np.uint8)
    for i in range(total):
        n = 4 if rng.random() < 0.5 else 16
        kinds, pres = base[n]
        w = pres[rng.integers(0, len(pres))].copy()
        if rng.random() < 0.01:
            w[rng.integers(0, w.shape[0]), rng.integers(0, 32)] ^= 1 << rng.integers(0, 8)
            expect[i] = 1
        kl.append(kinds); items.append(w)
    v = ShardedIssuer(iss[4]).verify_mixed(kl, items, issuers=iss)
    assert (v == expect).all() and expect.sum() > 20
    for n in (4, 16):
        sel = [i for i in range(total) if len(kl[i]) == n][:200]
        ov, _ = orc[n].verify_presentations(base[n][0], np.ascontiguousarray(np.stack([items[i] for i in sel])))
        assert (ov == v[sel]).all()


def test_multi_chunk_host_call_is_pipelined_and_exact(readme4):
    """A host call longer than max_batch takes the double-buffered copy/execute path: same verdicts as the oracle, in order,
    for a ragged number of chunks."""
    from aeonflux_b200 import Issuer, PresentationBatch
    orc, _, (sp, ip, sk) = readme4
    count = 5 * 1024 + 333
    kinds, pres, issu = orc.synth(b"SSPE", [0, 3], b"pipelined", 0, count)
    rng = np.random.default_rng(31)
    bad = rng.choice(count, 97, replace=False)
    for i in bad:
        pres[i, rng.integers(0, 28), rng.integers(0, 32)] ^= 1 << rng.integers(0, 8)
    iss = Issuer(sp, ip, sk, device=0, max_batch=1024)
    v = iss.verify_batch(PresentationBatch.from_items(kinds, pres))
    ov, _ = orc.verify_presentations(kinds, pres)
    assert (v == ov).all() and set(np.where(v)[0]) == set(bad)
    issu[77, 5, 3] ^= 4
    vi = iss.verify_issuance_batch(PresentationBatch.from_items(bytes([0, 0, 2, 2]), issu))
    ovi, _ = orc.verify_issuances(bytes([0, 0, 2, 2]), issu)
    assert (vi == ovi).all() and vi.sum() == 1
    # the same batches as item-major wire blobs (pipelined, 6 chunks) and in one pass
    assert (iss.verify_wire(kinds, pres) == ov).all()
    assert (iss.verify_wire(bytes([0, 0, 2, 2]), issu, issuance=True) == ovi).all()
    big = Issuer(sp, ip, sk, device=0, max_batch=8192)
    assert (big.verify_wire(kinds, pres) == ov).all()


@pytest.mark.parametrize("name", GOLDEN_SHAPES)
def test_show_golden_shapes(coracle, name):
    """AnonymousCredential::show on the GPU with supplied rng output: byte-identical to the oracle prover's presentations (item 0
    is the committed fixture), on a ragged batch over two chunks; the presentations verify like the oracle says."""
    from aeonflux_b200 import Issuer
    g = load_golden(name)
    sp, ip, sk = bytes.fromhex(g["sysparams"]), bytes.fromhex(g["issuer_pub"]), bytes.fromhex(g["secret"])
    orc = coracle.Issuer(sp, ip, sk)
    rk = bytes(REQ[k] for k in g["request"])
    count = 70
    kinds, pres, _, showin = orc.synth(rk, g["hide"], g["config"].encode(), 0, count, want_issuances=False, want_show_inputs=True)
    user = Issuer(sp, ip, None, device=0, max_batch=64)
    fields = np.ascontiguousarray(showin.transpose(1, 0, 2))
    fields[0, 9, 31] = 0xff                                     # item 9: t is not a canonical scalar
    res, st = user.show_batch(kinds, fields)
    got = res.fields.transpose(1, 0, 2)
    ok = np.ones(count, bool); ok[9] = False
    assert st[9] == 1 and not got[9].any() and not st[ok].any()
    assert (got[ok] == pres[ok]).all()
    assert got[0].tobytes().hex() == "".join(g["items"][0]["words"])
    wgot, wst = user.show_wire(kinds, np.ascontiguousarray(fields.transpose(1, 0, 2)))      # afx_show_wire: item-major in and out
    assert (wst == st).all() and (wgot == got).all()
    iss = Issuer(sp, ip, sk, device=0, max_batch=64)
    v = iss.verify_batch(res)
    ov, _ = orc.verify_presentations(kinds, np.ascontiguousarray(got))
    assert (v == ov).all() and (v[ok] == g["items"][0]["verdict"]).all()


def test_show_full_size_round_trip(readme4):
    """65,536 README-4 presentations made on the device from 2,048 credentials x 32 independent rng draws, then verified on the
    device: every one is accepted, presentations of the same credential differ (fresh z), a sample equals the oracle prover."""
    from aeonflux_b200 import Issuer
    orc, _, (sp, ip, sk) = readme4
    base = 2048
    kinds, pres, _, showin = orc.synth(b"SSPE", [0, 3], b"show-full", 0, base, want_issuances=False, want_show_inputs=True)
    count = 65536
    fields = np.tile(np.ascontiguousarray(showin.transpose(1, 0, 2)), (1, count // base, 1))
    nrand = 2 * (1 + 4 + 6)
    rng = np.random.default_rng(41)
    fields[-nrand:, base:] = rng.integers(0, 256, (nrand, count - base, 32), dtype=np.uint8)      # fresh z / blindings beyond the first copy
    user = Issuer(sp, ip, None, device=0, max_batch=count)
    res, st = user.show_batch(kinds, fields)
    assert not st.any()
    got = res.fields.transpose(1, 0, 2)
    assert (got[:base] == pres).all()
    assert (got[base:2 * base, 5] != got[:base, 5]).any(axis=1).all()                            # C_x_0 depends on z
    iss = Issuer(sp, ip, sk, device=0, max_batch=count)
    assert not iss.verify_batch(res).any()


def test_primitives_on_gpu(coracle):
    """Field / group / scalar primitives of the CUDA engine against the committed primitive vectors and the C oracle."""
    from aeonflux_b200 import Issuer
    from tests.test_host_logic import check_field_edges, check_primitives
    sp, ip, sk = coracle.make_issuer(1)
    iss = Issuer(sp, ip, None, device=0, max_batch=4)
    check_primitives(iss, coracle)
    check_field_edges(iss, n_random=100000)            # the PTX carry chains on edge-limb vectors (the emulation has plain C there)
    # 20,000 random 32-byte strings: the decode verdicts equal the oracle's (about 6.8 % decode), valid ones round-trip
    rng = np.random.default_rng(18)
    enc = rng.integers(0, 256, (20000, 32), dtype=np.uint8)
    out, ok = iss.selftest_primitive("decompress_compress", enc)
    L = coracle.lib()
    scratch = np.zeros(32, np.uint8)
    exp_ok = np.array([L.afxo_decompress_compress(enc[i].ctypes.data, scratch.ctypes.data) for i in range(20000)], np.uint8)
    assert (ok == exp_ok).all() and 1000 < ok.sum() < 1800
    assert (out[ok == 1] == enc[ok == 1]).all()


def test_batchable_proofs_exact_and_rlc_on_gpu(readme4):
    """BatchableProof presentations on the GPU: 20,000 README-4 items over three chunks.  All honest -> every chunk's random
    linear combination vanishes (no fallback); a few corrupted items -> exactly those are rejected, by the exact path and by
    the RLC path (which falls back only for the chunks that hold them); different seeds agree."""
    from aeonflux_b200 import Issuer, PresentationBatch
    from tests.common import to_batchable
    orc, _, (sp, ip, sk) = readme4
    base = 1000
    kinds, pres, _ = orc.synth(b"SSPE", [0, 3], b"batchable-gpu", 0, base, want_issuances=False)
    ov, _, tr = orc.verify_presentations(kinds, pres, trace=True)
    assert not ov.any()
    bp = np.tile(to_batchable(kinds, pres, tr["commitments"]), (20, 1, 1))
    count = len(bp)
    iss = Issuer(sp, ip, sk, device=0, max_batch=8192)
    batch = PresentationBatch.from_items(kinds, bp)
    assert not iss.verify_batchable(batch).any()
    v, fell_back = iss.verify_batchable_rlc(batch, bytes(range(32)))
    assert not v.any() and fell_back == 0
    rng = np.random.default_rng(51)
    bad = np.array([5, 8191, 8192, 19999])                      # chunks 0, 0, 1, 2
    for i in bad:
        bp[i, rng.integers(0, bp.shape[1]), rng.integers(0, 32)] ^= 1 << rng.integers(0, 8)
    batch = PresentationBatch.from_items(kinds, bp)
    expect = np.zeros(count, np.uint8); expect[bad] = 1
    assert (iss.verify_batchable(batch) == expect).all()
    for seed in (bytes(range(32)), bytes(32), bytes([7] * 32)):
        v, fell_back = iss.verify_batchable_rlc(batch, seed)
        assert (v == expect).all() and fell_back == 3
    bp[bad[2:]] = np.tile(to_batchable(kinds, pres, tr["commitments"]), (20, 1, 1))[bad[2:]]      # repair chunks 1 and 2
    v, fell_back = iss.verify_batchable_rlc(PresentationBatch.from_items(kinds, bp), bytes(range(32)))
    expect[bad[2:]] = 0
    assert (v == expect).all() and fell_back == 1


def test_async_submit_wait_on_gpu(readme4):
    """afx_verify_presentations_submit / afx_wait on the GPU: a stream of passes with two always in flight (the copy of one under
    the kernels of the other); every verdict vector equals the oracle's."""
    from aeonflux_b200 import Issuer, PresentationBatch
    orc, _, (sp, ip, sk) = readme4
    n_pass, per = 6, 4096
    kinds, pres, _ = orc.synth(b"SSPE", [0, 3], b"async-gpu", 0, 1024, want_issuances=False)
    rng = np.random.default_rng(61)
    items = pres[rng.integers(0, 1024, n_pass * per)].copy()
    bad = rng.choice(n_pass * per, 40, replace=False)
    for i in bad:
        items[i, rng.integers(0, 28), rng.integers(0, 32)] ^= 1 << rng.integers(0, 8)
    expect = np.zeros(n_pass * per, np.uint8); expect[bad] = 1
    iss = Issuer(sp, ip, sk, device=0, max_batch=per)
    # odd passes are staged in page-locked memory from the library's allocator (afx_host_alloc), even ones in numpy memory
    batches = [PresentationBatch.from_items(kinds, items[k * per:(k + 1) * per], host_array=iss.host_array if k % 2 else None) for k in range(n_pass)]
    pending, got = [], []
    for b in batches:
        if len(pending) == 2:
            got.append(pending.pop(0).wait())
        pending.append(iss.submit(b))
    got += [p.wait() for p in pending]
    assert (np.concatenate(got) == expect).all()
    assert (iss.verify_batch(PresentationBatch.from_items(kinds, items)) == expect).all()     # and the synchronous path still works after it


@pytest.mark.parametrize("budget_mb", ["200", "0"])
def test_noncanonical_wire_scalars_never_index_out_of_bounds_on_gpu(coracle, monkeypatch, budget_mb):
    """wire scalars >= l with top-word patterns 0x2001xxxx .. 0xffffffff in every scalar field, under the radix-2^16
    constant tables (default) and the radix-4096 ones (AFX_CTAB16_BUDGET_MB=0): rejected, no illegal address (a fault would poison
    the context and fail every later call), and the same context keeps verifying honest batches."""
    from aeonflux_b200 import Issuer
    from tests.test_host_logic import check_noncanonical_wire_scalars
    monkeypatch.setenv("AFX_CTAB16_BUDGET_MB", budget_mb)
    check_noncanonical_wire_scalars(lambda sp, ip, sk: Issuer(sp, ip, sk, device=0, max_batch=4096), coracle, count=16)


def test_two_contexts_run_the_fused_ladders_concurrently(readme4):
    """k_ladders claims its work by atomic ticket, so the consumer of Z only ever waits for producers that already run -- also when
    a second context's grid competes for the same SMs.  Two contexts on device 0, driven from two threads, several passes each
    with different corruption sets: every verdict vector is exact."""
    from concurrent.futures import ThreadPoolExecutor
    from aeonflux_b200 import Issuer, PresentationBatch
    orc, _, (sp, ip, sk) = readme4
    kinds, pres, _ = orc.synth(b"SSPE", [0, 3], b"two-ctx", 0, 512, want_issuances=False)
    per, passes = 16384, 6

    def worker(seed):
        rng = np.random.default_rng(seed)
        iss = Issuer(sp, ip, sk, device=0, max_batch=per)
        ok = True
        for _ in range(passes):
            items = pres[rng.integers(0, 512, per)].copy()
            bad = rng.choice(per, 50, replace=False)
            for i in bad:
                items[i, rng.integers(0, 28), rng.integers(0, 32)] ^= 1 << rng.integers(0, 8)
            expect = np.zeros(per, np.uint8); expect[bad] = 1
            ok &= bool((iss.verify_wire(kinds, items) == expect).all())
        iss.close()
        return ok

    with ThreadPoolExecutor(max_workers=2) as pool:
        assert all(pool.map(worker, [101, 202]))


def test_issuance_batchable_exact_and_rlc_on_gpu(coracle):
    from aeonflux_b200 import Issuer
    from tests.test_host_logic import check_issuance_batchable
    check_issuance_batchable(lambda sp, ip, sk, mb: Issuer(sp, ip, sk, device=0, max_batch=mb), coracle, count=700, max_batch=512)


def test_rlc_bisection_on_gpu(coracle, monkeypatch):
    """65,536 BatchableProof presentations with three bad items: the library's default 1,024-item leaves -- at most 3 x 1,024 items are
    re-verified exactly and a few dozen small passes run; then 1 % bad items: bisection gives up, whole-chunk exact check."""
    from aeonflux_b200 import Issuer
    from tests.test_host_logic import check_rlc_bisection
    check_rlc_bisection(lambda sp, ip, sk, mb: Issuer(sp, ip, sk, device=0, max_batch=mb), coracle, count=65536, max_batch=65536, leaf=1024, n_bad=3,
                        monkeypatch=monkeypatch)


def test_wide_table_allocation_failure_falls_back_on_gpu(coracle, monkeypatch):
    from aeonflux_b200 import Issuer
    from tests.test_host_logic import check_wide_table_allocation_failure
    check_wide_table_allocation_failure(lambda sp, ip, sk: Issuer(sp, ip, sk, device=0, max_batch=64), coracle, monkeypatch)


@pytest.mark.parametrize("n,request_,hide,tag", [(4, ("PS", "PS", "PP", "EP"), (0, 3), b"readme4"), (1, ("EP",), (0,), b"plain1"),
                                                 (5, ("PS", "PP", "EP", "EP", "EP"), (0, 2, 3, 4), b"three")])
def test_linked_presentations_on_gpu(n, request_, hide, tag):
    from aeonflux_b200 import Issuer
    from tests.test_host_logic import check_linked_presentations
    check_linked_presentations(lambda sp, ip, sk: Issuer(sp, ip, sk, device=0, max_batch=8), n, request_, hide, 12, tag)


def test_reference_accepts_spliced_encryption_on_gpu():
    from aeonflux_b200 import Issuer
    from tests.test_host_logic import check_reference_accepts_spliced_encryption
    check_reference_accepts_spliced_encryption(lambda sp, ip, sk: Issuer(sp, ip, sk, device=0, max_batch=8))


def test_linked_full_size_round_trip(readme4):
    """65,536 linked README-4 presentations made on the device (afx_show_linked) verify under the linked statement, are all rejected by
    the reference's statement (different transcript), and a flipped bit is caught."""
    from aeonflux_b200 import Issuer
    orc, _, (sp, ip, sk) = readme4
    base = 1024
    kinds, _, _, showin = orc.synth(b"SSPE", [0, 3], b"linked-full", 0, base, want_issuances=False, want_show_inputs=True)
    count = 65536
    fields = np.tile(np.ascontiguousarray(showin.transpose(1, 0, 2)), (1, count // base, 1))
    fields[-22:] = np.random.default_rng(5).integers(0, 256, (22, count, 32), dtype=np.uint8)      # fresh z and blindings for every item
    iss = Issuer(sp, ip, sk, device=0, max_batch=count)
    res, st = iss.show_batch(kinds, fields, linked=True)
    assert not st.any()
    wire = np.ascontiguousarray(res.fields.transpose(1, 0, 2))
    assert not iss.verify_wire(kinds, wire, linked=True).any()
    assert iss.verify_wire(kinds, wire[:4096]).all()
    rng = np.random.default_rng(6)
    bad = rng.choice(count, 200, replace=False)
    for i in bad:
        wire[i, rng.integers(0, 28), rng.integers(0, 32)] ^= 1 << rng.integers(0, 8)
    expect = np.zeros(count, np.uint8); expect[bad] = 1
    assert (iss.verify_wire(kinds, wire, linked=True) == expect).all()


def test_split_amac_ladder_on_gpu(coracle, monkeypatch):
    """Small batches run the aMAC ladder in parts (k_ladders_parts, k_amac_combine): same Z, commitments, challenges, verdicts for every
    cut; and a batch above the threshold still takes the single fused ladder."""
    from aeonflux_b200 import Issuer
    from tests.test_host_logic import check_split_amac
    check_split_amac(lambda sp, ip, sk: Issuer(sp, ip, sk, device=0, max_batch=4096), coracle, monkeypatch, count=700)


@pytest.mark.parametrize("count", [9000, 20000, 40000])
def test_medium_batches_place_their_amac_ctas_early(readme4, count):
    """Between the split small-batch path (<= 8,192 items) and the many-wave launch, the fused ladder grid places its aMAC CTAs by how
    long the launch is (afx_b200.cu:be_launch_ladders); the block order changes, the bytes must not: distinct items made on the
    device, flipped bits rejected, nothing else, a sample against the oracle."""
    import torch
    from aeonflux_b200 import Issuer
    from bench import KINDS_README4, synthesize_on_device
    orc, _, (sp, ip, sk) = readme4
    iss = Issuer(sp, ip, sk, device=0, max_batch=count)
    dev = synthesize_on_device(torch, iss, count, 500 + count, torch.cuda.current_stream())
    big = np.ascontiguousarray(dev.cpu().numpy().transpose(1, 0, 2))
    assert not iss.verify_wire(KINDS_README4, big).any()
    rng = np.random.default_rng(count)
    bad = rng.choice(count, 200, replace=False)
    for i in bad:
        big[i, rng.integers(0, 28), rng.integers(0, 32)] ^= 1 << rng.integers(0, 8)
    v = iss.verify_wire(KINDS_README4, big)
    expect = np.zeros(count, np.uint8)
    expect[bad] = 1
    assert (v == expect).all()
    sample = np.concatenate([bad[:40], rng.choice(count, 24, replace=False)])
    ov, _ = orc.verify_presentations(KINDS_README4, np.ascontiguousarray(big[sample]))
    assert (v[sample] == ov).all()
    iss.close()
