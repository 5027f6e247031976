"""Worker for tests/test_sharding.py: one process of a world_size-N gloo group.  Each rank holds a replicated issuer
context on the TEST-ONLY host-emulation build of the engine (there is no GPU here), verifies its contiguous slice of a
global batch and takes part in the verdict-bitmap gather; rank 0 checks the gathered verdicts against the oracle."""
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    out_path = sys.argv[1]
    import torch.distributed as dist
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    from aeonflux_b200 import Issuer, PresentationBatch
    from aeonflux_b200._binding import Binding
    from aeonflux_b200.shard import ShardedIssuer, slice_bounds
    from oracle import coracle as C
    emu = Binding(ctypes.CDLL(os.path.join(ROOT, "tests", "hostemu", "libafx_hostemu.so")))
    sp, ip, sk = C.make_issuer(4)
    orc = C.Issuer(sp, ip, sk)
    count = 45                                                   # not a multiple of the world size
    kinds, pres, issu = orc.synth(b"SSPE", [0, 3], b"shard", 0, count, threads=2)
    pres[3, 1, 0] ^= 1; pres[22, 9, 4] ^= 8; pres[44, 0, 0] ^= 1    # one bad item in (almost) every slice
    sh = ShardedIssuer(Issuer(sp, ip, sk, max_batch=8, _binding=emu))
    assert (sh.rank, sh.world) == (rank, world)
    before = sh.issuer.launch_count
    sh.issuer.verify_batch(PresentationBatch.from_items(kinds, pres[:1]))
    per_chunk = sh.issuer.launch_count - before
    before = sh.issuer.launch_count
    v = sh.verify_batch(PresentationBatch.from_items(kinds, pres))
    lo, hi = slice_bounds(count, rank, world)
    chunks = -(-(hi - lo) // 8)
    assert sh.issuer.launch_count - before == per_chunk * chunks, "a rank must only run its own slice"
    vi = sh.verify_issuance_batch(PresentationBatch.from_items(bytes([0, 0, 2, 2]), issu))
    # mixed stream: README-4 and a 4-attribute all-revealed shape interleaved
    k2, p2, _ = orc.synth(b"SSPP", [], b"shard2", 0, 20, threads=2)
    p2[7, 2, 1] ^= 2
    kl, items = [], []
    for i in range(count + 20):
        if i % 3 == 2 and i // 3 < 20:
            kl.append(k2); items.append(p2[i // 3])
        else:
            j = i - min(20, (i + 1) // 3)
            kl.append(kinds); items.append(pres[j])
    vm = sh.verify_mixed(kl, items)
    # item-major wire bytes, and the same mixed stream as records through the library's stream object (afx_stream_*)
    vw = sh.verify_wire(kinds, pres)
    from aeonflux_b200.shard import MixedStream, interleave_records
    order = np.array([1 if (i % 3 == 2 and i // 3 < 20) else 0 for i in range(count + 20)], np.uint8)
    blob, offsets = interleave_records([pres, p2], order)
    ms = MixedStream(_binding=emu)
    iss2 = Issuer(sp, ip, sk, max_batch=4, _binding=emu)
    assert [ms.add_shape(sh.issuer, kinds), ms.add_shape(iss2, k2)] == [0, 1]
    vs = sh.verify_stream(ms, blob, offsets, order)
    if rank == 0:
        ov, _ = orc.verify_presentations(kinds, pres, threads=2)
        ovi, _ = orc.verify_issuances(bytes([0, 0, 2, 2]), issu, threads=2)
        ov2, _ = orc.verify_presentations(k2, p2, threads=2)
        exp_m = []
        for i in range(count + 20):
            if i % 3 == 2 and i // 3 < 20:
                exp_m.append(int(ov2[i // 3]))
            else:
                exp_m.append(int(ov[i - min(20, (i + 1) // 3)]))
        json.dump({"world": world, "presentations_ok": bool((v == ov).all()), "rejected": [int(i) for i in np.where(v)[0]],
                   "issuances_ok": bool((vi == ovi).all()), "mixed_ok": bool((vm == np.asarray(exp_m, np.uint8)).all()),
                   "wire_ok": bool((vw == ov).all()), "stream_ok": bool((vs == np.asarray(exp_m, np.uint8)).all()),
                   "mixed_rejected": int(vm.sum())}, open(out_path, "w"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
