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
                   "mixed_rejected": 4}


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
