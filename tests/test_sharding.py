"""Multi-GPU host logic on CPU: contiguous slicing, verdict bitmaps, shape bucketing, and a world_size-2 gloo run of the
sharded verify path (each rank on the test-only host-emulation engine) checked against the oracle."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slice_bounds_partition_every_batch():
    from aeonflux_b200.shard import slice_bounds
    for total in (0, 1, 7, 45, 65536, 2**22 + 3):
        for world in (1, 2, 4, 8):
            b = [slice_bounds(total, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == total
            assert all(b[r][1] == b[r + 1][0] for r in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        slice_bounds(10, 2, 2)


def test_bitmap_round_trip_and_buckets():
    from aeonflux_b200.shard import bucket_by_shape, pack_bitmap, unpack_bitmap
    rng = np.random.default_rng(0)
    for n in (0, 1, 8, 9, 1000, 65537):
        v = (rng.random(n) < 0.01).astype(np.uint8)
        bits = pack_bitmap(v)
        assert len(bits) == (n + 7) // 8
        assert (unpack_bitmap(bits, n) == v).all()
    b = bucket_by_shape([b"\x01\x00\x02\x03", b"\x00\x00", b"\x01\x00\x02\x03", b"\x00\x00", b"\x02"])
    assert {k: list(v) for k, v in b.items()} == {b"\x01\x00\x02\x03": [0, 2], b"\x00\x00": [1, 3], b"\x02": [4]}


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_verify_gloo(tmp_path, coracle, world):
    from tests.test_host_logic import build_hostemu
    build_hostemu()
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = tmp_path / "result.json"
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_shard_worker.py"), str(out)], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    logs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(logs)
    res = json.load(open(out))
    assert res == {"world": world, "presentations_ok": True, "rejected": [3, 22, 44], "issuances_ok": True, "mixed_ok": True,
                   "wire_ok": True, "stream_ok": True, "mixed_rejected": 4}


def run_multi_gpu_issuer(coracle, binding, devices, count, max_batch):
    from aeonflux_b200 import PresentationBatch
    from aeonflux_b200.shard import MultiGpuIssuer
    sp, ip, sk = coracle.make_issuer(4)
    orc = coracle.Issuer(sp, ip, sk)
    kinds, pres, issu = orc.synth(b"SSPE", [0, 3], b"multi", 0, count)
    for i, w, b in ((1, 1, 0), (count // 2, 9, 5), (count - 1, 20, 31)):
        pres[i, w, b] ^= 1
    issu[count // 3, 6, 2] ^= 4
    ov, _ = orc.verify_presentations(kinds, pres)
    oi, _ = orc.verify_issuances(bytes([0, 0, 2, 2]), issu)
    m = MultiGpuIssuer(sp, ip, sk, devices=devices, max_batch=max_batch, _binding=binding)
    try:
        for _ in range(3):      # repeated: the per-context threads are reused
            assert (m.verify_batch(PresentationBatch.from_items(kinds, pres, host_array=m.host_array)) == ov).all()
        assert (m.verify_issuance_batch(PresentationBatch.from_items(bytes([0, 0, 2, 2]), issu)) == oi).all()
        assert ov.sum() == 3 and oi.sum() == 1
        assert len(m.verify_batch(PresentationBatch.from_items(kinds, pres[:0]))) == 0
        assert (m.verify_batch(PresentationBatch.from_items(kinds, pres[:2])) == ov[:2]).all()      # fewer items than devices
        # item-major wire bytes: each device copies its contiguous byte range
        assert (m.verify_wire(kinds, pres) == ov).all()
        assert (m.verify_wire(bytes([0, 0, 2, 2]), issu, issuance=True) == oi).all()
        # Issuer::issue across the devices: byte-identical to the oracle, and the result verifies everywhere
        from aeonflux_b200 import RequestBatch
        ik = bytes([0, 0, 2, 2])
        attrs = np.ascontiguousarray(issu[:, :4])
        rnd = np.random.default_rng(9).integers(0, 256, (count, 11, 64), dtype=np.uint8)
        issued, status = m.issue_batch(RequestBatch.from_request(ik, attrs, rnd))
        oout, ostatus, _ = orc.issue(ik, attrs, rnd)
        assert (status == ostatus).all() and (issued.fields.transpose(1, 0, 2)[:, 4:] == oout).all()
        assert not m.verify_issuance_batch(issued).any()
        # a structural mistake is an error from the multi call too (wrong field count)
        from aeonflux_b200._binding import AfxError
        with pytest.raises(AfxError):
            m.verify_batch(PresentationBatch(kinds, np.zeros((27, 4, 32), np.uint8)))
    finally:
        m.close()


def test_multi_gpu_issuer_threads_on_emulation(coracle):
    """One process, three contexts, three host threads (the single-process form of the sharding) on the host emulation."""
    import ctypes
    from aeonflux_b200._binding import Binding
    from tests.test_host_logic import build_hostemu
    run_multi_gpu_issuer(coracle, Binding(ctypes.CDLL(build_hostemu())), devices=[0, 0, 0], count=23, max_batch=4)


@pytest.mark.gpu
def test_multi_gpu_issuer_threads_on_gpu(coracle):
    """The same on the CUDA library: one context per visible GPU plus a second context on GPU 0, each driven by its own thread."""
    import torch
    devices = list(range(torch.cuda.device_count())) + [0]
    run_multi_gpu_issuer(coracle, None, devices=devices, count=5000, max_batch=1024)


def run_mixed_stream(coracle, binding, device, n4, n16, max4, max16, share_context=False):
    """BASELINE configs[4] in small: an item-mixed stream of README-4 and S16 presentations (two issuers) plus README-4-issuer
    issuances, ~3 % corrupted, pushed in several pieces through the library's stream object; verdicts equal the oracle's at
    every stream position."""
    from aeonflux_b200 import Issuer
    from aeonflux_b200.shard import MixedStream, interleave_records
    rng = np.random.default_rng(31)
    sp4, ip4, sk4 = coracle.make_issuer(4)
    sp16, ip16, sk16 = coracle.make_issuer(16)
    o4, o16 = coracle.Issuer(sp4, ip4, sk4), coracle.Issuer(sp16, ip16, sk16)
    k4, p4, i4 = o4.synth(b"SSPE", [0, 3], b"stream4", 0, n4)
    k16, p16, _ = o16.synth(b"SSSSSSPP" + b"E" * 8, [0, 1] + list(range(8, 16)), b"stream16", 0, n16, want_issuances=False)
    ik = bytes([0, 0, 2, 2])
    for pool in (p4, p16, i4):
        for i in rng.choice(len(pool), max(1, len(pool) // 30), replace=False):
            pool[i, rng.integers(0, pool.shape[1]), rng.integers(0, 32)] ^= 1 << rng.integers(0, 8)
    e4, _ = o4.verify_presentations(k4, p4)
    e16, _ = o16.verify_presentations(k16, p16)
    ei, _ = o4.verify_issuances(ik, i4)
    order = rng.permutation(np.concatenate([np.zeros(n4, np.uint8), np.ones(n16, np.uint8), np.full(n4, 2, np.uint8)]))
    blob, offsets = interleave_records([p4, p16, i4], order)
    expect = np.empty(len(order), np.uint8)
    expect[order == 0], expect[order == 1], expect[order == 2] = e4, e16, ei
    kw = {"_binding": binding} if binding is not None else {"device": device}
    iss4 = Issuer(sp4, ip4, sk4, max_batch=max4, **kw)
    iss16 = Issuer(sp16, ip16, sk16, max_batch=max16, **kw)
    iss4b = iss4 if share_context else Issuer(sp4, ip4, None, max_batch=max4, **kw)       # issuances need no secret key
    st = MixedStream(_binding=binding)
    assert [st.add_shape(iss4, k4), st.add_shape(iss16, k16), st.add_shape(iss4b, ik, issuance=True)] == [0, 1, 2]
    assert st.record_bytes == [28 * 32, 143 * 32, 17 * 32]
    verdicts = np.full(len(order), 9, np.uint8)
    cuts = [0, len(order) // 3, len(order) // 3 + 1, len(order)]
    for a, b in zip(cuts, cuts[1:]):
        st.push(blob, offsets[a:b], order[a:b], verdicts[a:b])
    st.flush()
    assert (verdicts == expect).all() and expect.any() and not expect.all()
    assert st.buckets_submitted >= -(-n4 // max4) + -(-n16 // max16) + -(-n4 // max4)
    # the stream object is reusable after a flush, and an unknown shape id is an argument error
    v2 = np.full(5, 9, np.uint8)
    st.push(blob, offsets[:5], order[:5], v2)
    st.flush()
    assert (v2 == expect[:5]).all()
    from aeonflux_b200._binding import AfxError
    with pytest.raises(AfxError):
        st.push(blob, offsets[:1], np.array([7], np.uint8), v2[:1])
    st.close()


@pytest.mark.parametrize("share", [False, True])
def test_mixed_stream_on_emulation(coracle, share):
    import ctypes
    from aeonflux_b200._binding import Binding
    from tests.test_host_logic import build_hostemu
    run_mixed_stream(coracle, Binding(ctypes.CDLL(build_hostemu())), 0, n4=23, n16=7, max4=5, max16=3, share_context=share)


@pytest.mark.gpu
@pytest.mark.parametrize("share", [False, True])
def test_mixed_stream_on_gpu(coracle, share):
    run_mixed_stream(coracle, None, 0, n4=3000, n16=700, max4=512, max16=256, share_context=share)
