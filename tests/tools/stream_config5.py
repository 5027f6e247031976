#!/usr/bin/env python
"""BASELINE configs[4]: streamed verification of N mixed presentations (50 % README-4, 50 % S16, 1 % corrupted), bucketed by
shape into chunks of 65,536 and sharded over the ranks of the process group (contiguous slices, verdict bitmap gathered).

    python tests/tools/stream_config5.py [--items 4194304] [--chunk 65536]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P tests/tools/stream_config5.py ...

Inputs are tiled from the committed bench fixtures (every item is independent, so tiling changes neither work nor traffic);
corruption = one flipped bit in a random word, which every verdict must catch (every word of a presentation is bound by a
transcript).  A random sample of each chunk is cross-checked against the CPU oracle.  Prints one JSON line on rank 0."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

K4 = bytes([1, 0, 2, 3])
K16 = bytes([1, 1, 0, 0, 0, 0, 2, 2] + [3] * 8)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--items", type=int, default=1 << 22)
    ap.add_argument("--chunk", type=int, default=65536, help="items per host call (one shape bucket)")
    ap.add_argument("--max-batch", type=int, default=0, help="items per device pass and per GPU (default: chunk / world); a host call "
                    "longer than this is pipelined inside the library (H2D of pass i+1 under the kernels of pass i)")
    ap.add_argument("--oracle-sample", type=int, default=32)
    ap.add_argument("--pageable", action="store_true", help="hand the library ordinary (pageable) numpy memory instead of afx_host_alloc buffers")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from aeonflux_b200 import Issuer, PresentationBatch
    from aeonflux_b200.shard import ShardedIssuer, slice_bounds
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    b4 = open(os.path.join(ROOT, "bench_data", "issuer4.bin"), "rb").read()
    b16 = open(os.path.join(ROOT, "bench_data", "issuer16.bin"), "rb").read()
    p4 = np.fromfile(os.path.join(ROOT, "bench_data", "readme4_1024.bin"), np.uint8).reshape(-1, 28, 32)
    p16 = np.fromfile(os.path.join(ROOT, "bench_data", "s16_256.bin"), np.uint8).reshape(-1, 143, 32)
    per_rank = args.max_batch or -(-args.chunk // world)
    iss = {4: Issuer(b4[:548], b4[548:612], b4[612:], device=local, max_batch=per_rank),
           16: Issuer(b16[:1316], b16[1316:1380], b16[1380:], device=local, max_batch=per_rank)}
    sh = {n: ShardedIssuer(i) for n, i in iss.items()}
    # struct-of-arrays staging of one chunk per shape, in page-locked memory from the library's allocator (afx_host_alloc)
    stage = {} if args.pageable else {4: iss[4].host_array((28, args.chunk, 32)), 16: iss[16].host_array((143, args.chunk, 32))}

    def make_batch(n, kinds, items):
        if args.pageable or len(items) != args.chunk:
            return PresentationBatch.from_items(kinds, items)
        stage[n][:] = items.transpose(1, 0, 2)
        return PresentationBatch(kinds, stage[n])
    orc = None
    if rank == 0 and args.oracle_sample:
        from oracle import coracle as C
        orc = {4: C.Issuer(b4[:548], b4[548:612], b4[612:]), 16: C.Issuer(b16[:1316], b16[1316:1380], b16[1380:])}
    rng = np.random.default_rng(5)          # same stream on every rank: each rank builds the chunk and takes its slice
    srng = np.random.default_rng(6)         # rank 0's oracle sampling draws must not disturb the shared stream
    n_chunks = args.items // args.chunk
    # one-time costs (workspace allocation, shape compilation, pinned staging) are paid by a warm-up call per shape and
    # reported separately: the stream figure is the steady state
    t_setup = time.perf_counter()
    for n, kinds, base in ((4, K4, p4), (16, K16, p16)):
        w = base[np.arange(min(args.chunk, 2 * per_rank * world)) % len(base)]
        assert not sh[n].verify_batch(make_batch(n, kinds, w)).any()
    setup = time.perf_counter() - t_setup
    done, rejected, expected_rejected, mism, oracle_checked, busy = 0, 0, 0, 0, 0, 0.0
    by_shape = {4: 0.0, 16: 0.0}
    t_start = time.perf_counter()
    for c in range(n_chunks):
        n, kinds, base = (4, K4, p4) if c % 2 == 0 else (16, K16, p16)     # chunks alternate between the two shape buckets
        start = int(rng.integers(0, len(base)))
        idx = (start + np.arange(args.chunk)) % len(base)
        items = base[idx]                                                    # [chunk][W][32]
        bad = rng.choice(args.chunk, args.chunk // 100, replace=False)
        items[bad, rng.integers(0, items.shape[1], len(bad)), rng.integers(0, 31, len(bad))] ^= (1 << rng.integers(0, 8, len(bad))).astype(np.uint8)
        expect = np.zeros(args.chunk, np.uint8); expect[bad] = 1
        batch = make_batch(n, kinds, items)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        v = sh[n].verify_batch(batch)
        dt = time.perf_counter() - t0
        busy += dt; by_shape[n] += dt
        mism += int((v != expect).sum())
        rejected += int(v.sum()); expected_rejected += len(bad); done += args.chunk
        if orc is not None:
            sample = np.concatenate([bad[:args.oracle_sample // 2], srng.choice(args.chunk, args.oracle_sample // 2, replace=False)])
            ov, _ = orc[n].verify_presentations(kinds, np.ascontiguousarray(items[sample]))
            mism += int((ov != v[sample]).sum()); oracle_checked += len(sample)
    wall = time.perf_counter() - t_start
    if rank == 0:
        print(json.dumps({"workload": "streamed verification of %d mixed presentations (README-4 / S16 chunks of %d alternating, 1%% corrupted)" % (done, args.chunk),
                          "n_gpus": world, "items": done, "host_memory": "pageable" if args.pageable else "page-locked (afx_host_alloc)", "verify_seconds": busy, "setup_seconds_warmup_call_per_shape": setup, "wall_seconds_incl_input_synthesis": wall,
                          "presentations_per_s": done / busy, "seconds_by_shape": {"readme4": by_shape[4], "s16": by_shape[16]}, "rejected": rejected, "expected_rejected": expected_rejected,
                          "mismatches": mism, "oracle_cross_checked": oracle_checked}))
    if world > 1:
        dist.destroy_process_group()
    if mism:
        raise SystemExit("verdict mismatches: %d" % mism)


if __name__ == "__main__":
    main()
