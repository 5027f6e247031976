#!/usr/bin/env python
"""Host-side memory-safety fuzz (no GPU): the test-only host emulation of the engine (tests/hostemu: the same C-ABI implementation,
shape compiler, marshaling and stage bodies as the CUDA library, plain loops for kernels) is compiled with AddressSanitizer and
UndefinedBehaviorSanitizer and driven with adversarial wire input through every batch entry point:

  * differential runs against the C oracle on mutated presentations / issuances (tests/common.py:fuzz_mutations: edge scalars and
    encodings, random bytes, valid words in the wrong place, bit flips incl. the top byte), both constant-table radices;
  * random attribute shapes through verify / show / issue / BatchableProof exact + RLC (tests/test_random_shapes.py);
  * garbage into the prover calls (afx_issue*, afx_show*: undecodable attribute points, scalars >= l, random rng bytes) and wholly
    random items into every verifier call, SoA and item-major, chunked and one-pass;
  * every third round the library-owned host layers: afx_stream_* (mixed shapes, corrupted records, several pushes, reuse after a
    flush) and afx_multi_* (three contexts, three library threads).

Any out-of-bounds access, use after free, misaligned or overflowing arithmetic aborts the run (the GPU-side counterpart is
compute-sanitizer over tests/tools/sanitize_small.py, profiles/r02_compute_sanitizer.txt).

    python tests/tools/fuzz_asan.py [seconds] [light]  # default 60; `light` = one short round (the pytest entry, tests/test_fuzz.py)
"""
import ctypes
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
EMU_DIR = os.path.join(ROOT, "tests", "hostemu")
SO = os.path.join(EMU_DIR, "libafx_hostemu_asan.so")


def build():
    src = os.path.join(EMU_DIR, "hostemu.cpp")
    csrc = os.path.join(ROOT, "aeonflux_b200", "csrc")
    newest = max(os.path.getmtime(os.path.join(csrc, f)) for f in os.listdir(csrc) if not f.endswith(".so"))
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(newest, os.path.getmtime(src)):
        subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-pthread", "-fvisibility=hidden",
                               "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer",
                               "-x", "c++", "-o", SO, src])
    return SO


def asan_runtime():
    out = subprocess.check_output(["gcc", "-print-file-name=libasan.so"], text=True).strip()
    return out if os.path.isabs(out) and os.path.exists(out) else None


def main(seconds, light=False):
    sys.path.insert(0, ROOT)
    import numpy as np
    from aeonflux_b200 import Issuer, PresentationBatch, RequestBatch
    from aeonflux_b200._binding import AfxError, Binding
    from oracle import coracle
    from tests.common import differential_fuzz, fuzz_mutations
    from tests.test_random_shapes import run_shapes
    coracle.build()
    emu = Binding(ctypes.CDLL(SO))
    t0, rounds, tot_acc, tot_rej = time.time(), 0, 0, 0
    while True:
        seed = 1000 + rounds
        rng = np.random.default_rng(seed)
        # 1. differential runs, both table radices, several chunkings (max_batch below / above the count)
        for wide in ((False,) if light else (False, True)):
            if wide:
                os.environ["AFX_HOSTEMU_CTAB16"] = "1"
            mb = int(rng.choice([3, 17, 64]))
            a, r = differential_fuzz(lambda sp, ip, sk: Issuer(sp, ip, sk, max_batch=mb, _binding=emu), coracle, seed, 24 if wide else (12 if light else 48),
                                     with_trace=not wide)
            os.environ.pop("AFX_HOSTEMU_CTAB16", None)
            tot_acc += a; tot_rej += r
        # 2. random shapes through verify / show / issue / BatchableProof
        if not light:
            run_shapes(coracle, emu, seed, 2, 4, {})
        # 3. garbage into the prover calls and wholly random items into the verifier calls
        n = int(rng.integers(1, 7))
        rk = bytes(rng.choice([ord("S"), ord("P"), ord("E")], n).tolist())
        hide = [i for i in range(n) if rk[i] != ord("P") and rng.random() < 0.5]
        sp, ip, sk = coracle.make_issuer(n)
        orc = coracle.Issuer(sp, ip, sk)
        kinds, pres, issu, showin = orc.synth(rk, hide, b"asan-%d" % seed, 0, 6, want_show_inputs=True)
        iss = Issuer(sp, ip, sk, max_batch=4, _binding=emu)
        ik = bytes(0 if c == ord("S") else 2 for c in rk)
        junk = rng.integers(0, 256, pres.shape, dtype=np.uint8)
        assert iss.verify_wire(kinds, junk).all() and iss.verify_batch(PresentationBatch.from_items(kinds, junk)).all()
        ijunk = rng.integers(0, 256, issu.shape, dtype=np.uint8)
        assert iss.verify_wire(ik, ijunk, issuance=True).all()
        A = fuzz_mutations(np.ascontiguousarray(issu[:, :n]), rng, 12)
        R = rng.integers(0, 256, (12, n + 7, 64), dtype=np.uint8)
        out, status, _ = orc.issue(ik, A, R)
        res, st = iss.issue_batch(RequestBatch.from_request(ik, A, R))
        assert (st == status).all() and (res.fields.transpose(1, 0, 2)[:, n:] == out).all()
        reqs = np.ascontiguousarray(RequestBatch.from_request(ik, A, R).fields.transpose(1, 0, 2))
        wout, wst = iss.issue_wire(ik, reqs)
        assert (wst == status).all() and (wout[:, n:] == out).all()
        sjunk = fuzz_mutations(showin, rng, 12)
        try:
            sres, sst = iss.show_batch(kinds, np.ascontiguousarray(sjunk.transpose(1, 0, 2)))
            wres, wsst = iss.show_wire(kinds, sjunk)
            assert (sst == wsst).all() and (sres.fields.transpose(1, 0, 2) == wres).all()
            good = sst == 0                  # whatever show accepted as input must come out in a form the verifier can parse
            if good.any():
                iss.verify_wire(kinds, np.ascontiguousarray(wres[good]))
        except AfxError:
            pass
        iss.close()
        # 4. the library-owned host layers: the mixed-shape stream object and the multi-device handle (their threads, buckets, slices)
        if not light and rounds % 3 == 0:
            from tests.test_sharding import run_mixed_stream, run_multi_gpu_issuer
            run_mixed_stream(coracle, emu, 0, n4=11, n16=3, max4=int(rng.integers(2, 6)), max16=2, share_context=bool(rounds % 2))
            run_multi_gpu_issuer(coracle, emu, devices=[0, 0, 0], count=int(rng.integers(7, 20)), max_batch=4)
        rounds += 1
        if time.time() - t0 > seconds:
            break
    print("fuzz_asan ok: %d rounds in %.0f s, %d mutated items accepted / %d rejected by both the engine and the oracle, no sanitizer report"
          % (rounds, time.time() - t0, tot_acc, tot_rej))


if __name__ == "__main__":
    secs = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
    light = len(sys.argv) > 2 and sys.argv[2] == "light"
    if os.environ.get("AFX_FUZZ_CHILD") != "1":
        build()
        rt = asan_runtime()
        if rt is None:
            print("fuzz_asan skipped: no libasan runtime next to gcc"); sys.exit(77)
        env = dict(os.environ, AFX_FUZZ_CHILD="1", LD_PRELOAD=rt, ASAN_OPTIONS="detect_leaks=0:abort_on_error=1:handle_segv=1",
                   UBSAN_OPTIONS="print_stacktrace=1:halt_on_error=1")
        sys.exit(subprocess.call([sys.executable, os.path.abspath(__file__)] + sys.argv[1:], env=env))
    main(secs, light)
