#!/usr/bin/env python
"""bench.py -- presentations verified/sec (Issuer::verify, 4 attributes) on N B200s, with the integer-pipe roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl native|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A step = one pass of the hot path (afx_verify_presentations) over one batch of B = 65,536 README-4 presentations per GPU
(BASELINE.json configs[1]).  `value` is timed with the batch resident in HBM (CUDA events on the launching stream, L2
flushed between steps); `e2e` is the same call through the host-buffer C ABI with the H2D copy of the 896-byte items and
the D2H copy of the verdicts inside the timed region.  Ranks are independent (weak scaling: every rank verifies its own
65,536-item slice with a replicated issuer context; only the accept/reject counts are gathered).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

KINDS_README4 = bytes([1, 0, 2, 3])   # [SecretScalar, PublicScalar, PublicPoint, SecretPoint]  (README.md:44-117, attrs 0 and 3 hidden)
WORDS = 28

# ---- algorithmic work model: limb-products per field op, 1 limb-product = 2 IMAD issue slots (IMAD.WIDE is half rate,
# measured: profiles/r01_microbench_imad.json) ----------------------------------------------------------------------------------
S_LP, M_LP = 44, 72
DBL, ADD, MADD = 4 * S_LP + 4 * M_LP, 8 * M_LP, 7 * M_LP
DECOMPRESS, COMPRESS = 258 * S_LP + 24 * M_LP, 256 * S_LP + 30 * M_LP
TABLE = DBL + 7 * ADD


def work_model(kinds):
    """Algorithmic limb-products per item and per stage for a presentation shape (minimal schedule of SURVEY 8d)."""
    n = len(kinds)
    r_s, h_s = sum(k == 0 for k in kinds), sum(k == 1 for k in kinds)
    r_p, h_p = sum(k == 2 for k in kinds), sum(k == 3 for k in kinds)
    n_dec = 3 + n + r_p + 7 * h_p
    n_msm = 2 + h_s + r_s + r_p + 5 * h_p
    var_terms = 1 + 2 + h_s + r_s + r_p + h_p * (1 + 2 + 2 + 3 + 1)
    con_terms = 1 + 2 + 2 * h_s + r_s + r_p + h_p * (3 + 1 + 0 + 1 + 2)
    points = n_dec * DECOMPRESS + (var_terms - 1) * TABLE + 2 * h_p * COMPRESS
    amac = 252 * DBL + (2 + n) * (7 + 64) * ADD + r_s * 64 * MADD + (r_s + r_p + 2) * ADD + COMPRESS + TABLE
    msm = n_msm * 253 * DBL + var_terms * (253 / 6) * ADD + con_terms * (253 / 9) * MADD + n_msm * COMPRESS
    total = points + amac + msm
    return {"points": points, "amac": amac, "msm": msm, "total": total}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons (B200_PROFILING.md).  Started before the warm-up (nvidia-smi needs a moment to
    come up, longer on an 8-GPU box) and polled every 25 ms; rows are time-stamped on arrival and summary() only uses those
    that arrived inside the marked timed windows."""

    def __init__(self, index):
        self.index, self.rows, self.proc, self.windows = index, [], None, []

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: [self.rows.append((time.time(), line)) for line in self.proc.stdout], daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def window(self):
        sampler = self

        class _W:
            def __enter__(self):
                self.t0 = time.time()

            def __exit__(self, *a):
                sampler.windows.append((self.t0, time.time()))
        return _W()

    def stop(self):
        if self.proc:
            time.sleep(0.05)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for ts, line in self.rows:
            if not any(a <= ts <= b + 0.03 for a, b in self.windows):
                continue
            p = [x.strip() for x in line.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"], "samples": 0, "rows_total": len(self.rows)}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm),
                "windows": "device-resident and end-to-end timed regions"}


def load_fixture(batch):
    pres = np.fromfile(os.path.join(ROOT, "bench_data", "readme4_1024.bin"), np.uint8).reshape(-1, WORDS, 32)
    blob = open(os.path.join(ROOT, "bench_data", "issuer4.bin"), "rb").read()
    sp, ip, sk = blob[:548], blob[548:612], blob[612:]
    reps = (batch + len(pres) - 1) // len(pres)
    items = np.tile(pres, (reps, 1, 1))[:batch]
    return sp, ip, sk, items


def synthesize_on_device(torch, issuer, B, seed, stream):
    """65,536 DISTINCT honest README-4 presentations made on the GPU itself: random attributes -> Issuer::issue (credentials) ->
    AnonymousCredential::show with attributes 0 and 3 hidden (fresh z and blindings per item, one symmetric keypair).  Returns the
    device-resident struct-of-arrays batch [28][B][32].  (Needs no CPU oracle; a sample is cross-checked against it afterwards.)"""
    rng = np.random.default_rng(seed)
    s = stream.cuda_stream

    def rand_words(k):
        return torch.from_numpy(rng.integers(0, 256, (k, B, 32), dtype=np.uint8)).cuda()

    def rand_scalars(k):
        w = rng.integers(0, 256, (k, B, 32), dtype=np.uint8); w[:, :, 31] &= 0x0f       # < 2^252 < l
        return torch.from_numpy(w).cuda()

    status = torch.empty(B, dtype=torch.uint8, device="cuda")
    # valid, distinct points: the U = RistrettoPoint::from_uniform_bytes(rng) outputs of three throw-away issuances
    pts = []
    for _ in range(3):
        req = torch.cat([rand_scalars(4), rand_words(22)])
        out = torch.empty((13, B, 32), dtype=torch.uint8, device="cuda")
        issuer.issue_batch_device(bytes([0, 0, 0, 0]), B, req.data_ptr(), out.data_ptr(), status.data_ptr(), s)
        torch.cuda.synchronize()
        pts.append(out[1].clone())
    P2, M1, M2 = pts
    m = rand_scalars(3)                                                                  # m0, m1, m3
    # credentials over attributes [m0, m1, P2, plaintext(M1, M2, m3)] (the plaintext enters the aMAC through M1, amacs.rs:241)
    req = torch.cat([m[0:2], P2[None], M1[None], rand_words(22)])
    cred = torch.empty((13, B, 32), dtype=torch.uint8, device="cuda")
    issuer.issue_batch_device(bytes([0, 0, 2, 2]), B, req.data_ptr(), cred.data_ptr(), status.data_ptr(), s)
    torch.cuda.synchronize()
    assert int(status.sum().item()) == 0
    kp = np.frombuffer(open(os.path.join(ROOT, "bench_data", "keypair4.bin"), "rb").read(), np.uint8).reshape(4, 1, 32)
    kp_dev = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(kp, (4, B, 32)))).cuda()
    show_in = torch.cat([cred[0:3], m[0:2], P2[None], M1[None], M2[None], m[2:3], kp_dev, rand_words(22)])   # 35 fields
    pres = torch.empty((WORDS, B, 32), dtype=torch.uint8, device="cuda")
    issuer.show_batch_device(KINDS_README4, B, show_in.data_ptr(), pres.data_ptr(), status.data_ptr(), s)
    torch.cuda.synchronize()
    assert int(status.sum().item()) == 0
    return pres


def cpu_leg(sp, ip, sk, items, sample, threads):
    """The restated reference CPU path (oracle/c, reference schedule) on `sample` items with `threads` host threads."""
    from oracle import coracle as C
    C.build()
    orc = C.Issuer(sp, ip, sk)
    sub = np.ascontiguousarray(items[:sample])
    t0 = time.perf_counter()
    verdicts, _ = orc.verify_presentations(KINDS_README4, sub, threads=threads)
    wall = time.perf_counter() - t0
    return verdicts, sample / wall, wall



def time_device(torch, stream, flush, fn, steps, warmup=2):
    """Device ms per call of fn (enqueues on `stream`), L2 flushed before each timed call."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    total = 0.0
    for _ in range(steps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); fn(); e1.record(stream)
        torch.cuda.synchronize()
        total += e0.elapsed_time(e1)
    return total / steps


def secondary_measurements(torch, issuer4, items4, local, stream, flush, B, steps):
    """BASELINE configs[2] and configs[3] on the same GPU (device-resident inputs, CUDA events): batch Issuer::issue and
    CredentialIssuance::verify of B revealed 4-attribute requests, and Issuer::verify of S16 presentations.  Reported under
    `secondary`; the headline stays configs[1]."""
    from aeonflux_b200 import Issuer
    out = {}
    # ---- issuance: requests [scalar, scalar, point, point]; the point attributes are valid encodings taken from the fixture
    kinds = bytes([0, 0, 2, 2])
    rng = np.random.default_rng(1234)
    n, nf = 4, 3 * 4 + 14
    req = np.empty((nf, B, 32), np.uint8)
    sc = rng.integers(0, 256, (2, B, 32), dtype=np.uint8); sc[:, :, 31] &= 0x0f       # < 2^252 < l: canonical scalars
    req[0:2] = sc
    req[2] = items4[:B, 5]; req[3] = items4[:B, 6]                                     # C_x_0, C_x_1 of the fixture: valid points
    req[4:] = rng.integers(0, 256, (nf - 4, B, 32), dtype=np.uint8)                    # rng output for t, U, blindings
    req_dev = torch.from_numpy(req).cuda()
    iss_dev = torch.empty((2 * n + 9, B, 32), dtype=torch.uint8, device="cuda")
    iss_dev[:n] = req_dev[:n]
    status_dev = torch.empty(B, dtype=torch.uint8, device="cuda")
    ms = time_device(torch, stream, flush, lambda: issuer4.issue_batch_device(kinds, B, req_dev.data_ptr(), iss_dev[n:].data_ptr(), status_dev.data_ptr(), stream.cuda_stream), steps)
    assert int(status_dev.sum().item()) == 0
    out["issue_4attr"] = {"workload": "batch Issuer::issue (Amac::tag + ProofOfIssuance::prove) of %d revealed 4-attribute requests (BASELINE configs[2])" % B,
                          "value": B / (ms * 1e-3), "unit": "issuances/s", "ms_per_step": ms}
    verdicts_dev = torch.empty(B, dtype=torch.uint8, device="cuda")
    ms = time_device(torch, stream, flush, lambda: issuer4.verify_issuance_batch_device(kinds, B, iss_dev.data_ptr(), verdicts_dev.data_ptr(), stream.cuda_stream), steps)
    assert int(verdicts_dev.sum().item()) == 0, "issued credentials failed CredentialIssuance::verify"
    out["verify_issuance_4attr"] = {"workload": "batch CredentialIssuance::verify of the %d issuances above" % B,
                                    "value": B / (ms * 1e-3), "unit": "issuances/s", "ms_per_step": ms}
    # ---- AnonymousCredential::show (user-side prover): inputs = the issuances above as credentials of shape
    # [PS, PS, PP, PP] (nothing hidden), fresh rng bytes; the presentations are then verified
    k_show = bytes([0, 0, 2, 2])
    nf_in, nf_out = 3 + 4 + 2 * (1 + 3), 1 + 3 + 3 + 4 + 4
    sh = torch.empty((nf_in, B, 32), dtype=torch.uint8, device="cuda")
    sh[0:3] = iss_dev[n:n + 3]; sh[3:7] = iss_dev[0:4]
    sh[7:] = torch.from_numpy(rng.integers(0, 256, (nf_in - 7, B, 32), dtype=np.uint8)).cuda()
    pres_dev = torch.empty((nf_out, B, 32), dtype=torch.uint8, device="cuda")
    ms = time_device(torch, stream, flush, lambda: issuer4.show_batch_device(k_show, B, sh.data_ptr(), pres_dev.data_ptr(), status_dev.data_ptr(), stream.cuda_stream), steps)
    assert int(status_dev.sum().item()) == 0
    issuer4.verify_batch_device(k_show, B, pres_dev.data_ptr(), verdicts_dev.data_ptr(), stream.cuda_stream)
    torch.cuda.synchronize()
    assert int(verdicts_dev.sum().item()) == 0, "presentations made on the device failed Issuer::verify"
    out["show_4attr_revealed"] = {"workload": "batch AnonymousCredential::show of %d all-revealed 4-attribute credentials (user-side prover), verified afterwards" % B,
                                  "value": B / (ms * 1e-3), "unit": "presentations/s", "ms_per_step": ms}
    # ---- BatchableProof mode (opt-in, not the reference's encoding): the same presentations re-encoded with their commitments,
    # verified exactly and by one random linear combination per batch (Pippenger); host-buffer calls, wall clock
    from aeonflux_b200 import PresentationBatch, compact_to_batchable
    comp = PresentationBatch.from_items(KINDS_README4, items4[:B])
    _, dbg = issuer4.verify_batch(comp, debug=True)
    bb_host = torch.from_numpy(compact_to_batchable(KINDS_README4, comp.fields, dbg["commitments"])).pin_memory()
    bb = PresentationBatch(KINDS_README4, bb_host.numpy())
    for name, fn in (("verify_batchable_exact", lambda: issuer4.verify_batchable(bb)), ("verify_batchable_rlc", lambda: issuer4.verify_batchable_rlc(bb, bytes(range(32)))[0])):
        fn()
        t0 = time.perf_counter()
        for _ in range(steps):
            v = fn()
        dt = (time.perf_counter() - t0) / steps
        assert not v.any()
        out[name] = {"workload": "%d README-4 presentations in BatchableProof form through the host-buffer call (%s), all valid" % (B, "one MSM per constraint" if "exact" in name else "one random linear combination per batch, Pippenger"),
                     "value": B / dt, "unit": "presentations/s", "ms_per_step": dt * 1e3, "timing": "end to end (H2D + kernels + D2H), wall clock"}
    # ---- S16 presentations (configs[3]) at the config's full batch: 65,536 x 4,576 B in, 4.6 GB of ladder tables
    try:
        blob = open(os.path.join(ROOT, "bench_data", "issuer16.bin"), "rb").read()
        pres = np.fromfile(os.path.join(ROOT, "bench_data", "s16_256.bin"), np.uint8).reshape(-1, 143, 32)
    except OSError:
        return out
    sp, ip, sk = blob[:1316], blob[1316:1380], blob[1380:]
    B16 = B
    k16 = bytes([1, 1, 0, 0, 0, 0, 2, 2] + [3] * 8)
    issuer16 = Issuer(sp, ip, sk, device=local, max_batch=B16)
    f16 = torch.from_numpy(np.ascontiguousarray(np.tile(pres, ((B16 + 255) // 256, 1, 1))[:B16].transpose(1, 0, 2))).cuda()
    v16 = torch.empty(B16, dtype=torch.uint8, device="cuda")
    ms = time_device(torch, stream, flush, lambda: issuer16.verify_batch_device(k16, B16, f16.data_ptr(), v16.data_ptr(), stream.cuda_stream), steps)
    assert int(v16.sum().item()) == 0
    wm = work_model(k16)
    out["verify_s16"] = {"workload": "batch Issuer::verify of %d 16-attribute presentations, 8 hidden plaintext attributes (BASELINE configs[3])" % B16,
                         "value": B16 / (ms * 1e-3), "unit": "presentations/s", "ms_per_step": ms,
                         "pipeline_frac_of_imad_peak": (B16 * 2 * wm["total"] / (ms * 1e-3)) / (148 * 64 * 1965e6)}
    issuer16.close()
    return out

def run_reference(args, rank):
    """--impl reference: the reference's own CPU schedule of Issuer::verify on the host cores.  The Rust crate cannot be
    built in this image (no rustc/cargo, un-vendored deps), so this is the C restatement of its schedule (oracle/c)."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sp, ip, sk, items = load_fixture(args.batch)
    sample = min(args.batch, cores * 2048)          # ~2 s of work per step on every host core
    for _ in range(args.warmup):
        cpu_leg(sp, ip, sk, items, min(sample, cores * 16), cores)
    t_total, n_total = 0.0, 0
    for _ in range(args.steps):
        v, rate, wall = cpu_leg(sp, ip, sk, items, sample, cores)
        assert not v.any()
        t_total += wall; n_total += sample
    value = n_total / t_total
    line = {"impl": "reference", "metric": "presentations_verified_per_sec", "value": value, "unit": "presentations/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64 (5x51-bit limbs, like dalek u64_backend)", "data": "synthetic",
            "config": {"workload": "Issuer::verify, README-4 presentations [SS,PS,PP,SP], %d-item sample per step of the 65,536-item batch" % sample},
            "cpu_baseline": {"value": value, "unit": "presentations/s", "cores": cores, "kind": "port",
                             "sample": "%d items/step x %d steps, C restatement of the reference schedule (not the Rust binary)" % (sample, args.steps)},
            "e2e": {"value": value, "unit": "presentations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the issuance / S16 measurements (configs[2], configs[3])")
    ap.add_argument("--tiled-input", action="store_true", help="tile the 1,024-item CPU-made fixture instead of synthesizing the batch on the device")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank)

    import torch
    import torch.distributed as dist
    from aeonflux_b200 import Issuer, PresentationBatch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = args.batch
    sp, ip, sk, items = load_fixture(B)
    if world > 1:   # each rank takes its own slice of the global stream: rotate so the ranks do not hold identical bytes
        items = np.roll(items, rank * 131, axis=0)
    issuer = Issuer(sp, ip, sk, device=local, max_batch=B)
    # host staging in pinned memory (the e2e leg copies from here); SoA [field][item][32]
    host = torch.empty((WORDS, B, 32), dtype=torch.uint8).pin_memory()
    if args.tiled_input:
        host.numpy()[:] = items.transpose(1, 0, 2)
        fields_dev = host.cuda(non_blocking=False)
        input_note = "1,024 distinct presentations tiled to the batch (tests/golden/make_bench_fixture.py)"
    else:
        fields_dev = synthesize_on_device(torch, issuer, B, 1000 + rank, torch.cuda.current_stream())
        host.copy_(fields_dev)
        items = np.ascontiguousarray(host.numpy().transpose(1, 0, 2))
        input_note = "%d distinct presentations per GPU, issued and shown on the device from random attributes (afx_issue -> afx_show); " \
                     "a sample is cross-checked against the CPU oracle" % B
    verdicts_dev = torch.empty(B, dtype=torch.uint8, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2
    stream = torch.cuda.current_stream()
    kinds = KINDS_README4

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        issuer.verify_batch_device(kinds, B, fields_dev.data_ptr(), verdicts_dev.data_ptr(), stream.cuda_stream)
        e1.record(stream)
        return e0, e1

    # ---- kernel-only: inputs resident in HBM --------------------------------------------------------------------
    clocks = ClockSampler(local).start()
    issuer.set_stage_timing(True)
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    launches0 = issuer.launch_count
    stage_sum = {k: 0.0 for k in Issuer.STAGES}
    with clocks.window():
        evs = []
        for _ in range(args.steps):
            evs.append(step_device())
            torch.cuda.synchronize()
            for k, v in issuer.stage_times_ms().items():
                stage_sum[k] += v
        barrier()
    launches = issuer.launch_count - launches0
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    assert int(verdicts_dev.sum().item()) == 0, "honest presentations were rejected"
    issuer.set_stage_timing(False)

    # ---- end to end: host buffers through the C ABI, H2D + kernels + D2H per step --------------------------------------
    # headline e2e: afx_verify_presentations_wire on item-major bytes in pinned host memory (one H2D copy of the batch);
    # also timed: afx_verify_presentations on the struct-of-arrays form (one copy per field).
    wire_host = torch.empty((B, WORDS, 32), dtype=torch.uint8).pin_memory()
    wire_host.numpy()[:] = items
    batch = PresentationBatch(kinds, host.numpy())
    for _ in range(2):
        issuer.verify_wire(kinds, wire_host.numpy())
        issuer.verify_batch(batch)
    barrier()
    e2e_s, e2e_soa_s = 0.0, 0.0
    with clocks.window():
        for _ in range(args.steps):
            t0 = time.perf_counter()
            v = issuer.verify_wire(kinds, wire_host.numpy())
            e2e_s += time.perf_counter() - t0
            assert not v.any()
    barrier()
    for _ in range(args.steps):
        t0 = time.perf_counter()
        v = issuer.verify_batch(batch)
        e2e_soa_s += time.perf_counter() - t0
        assert not v.any()
    barrier()
    # streamed: afx_verify_presentations_submit / afx_wait with two submissions in flight -- every step still copies its own
    # inputs H2D and its verdicts D2H inside the timed region, but the copy of step k+1 runs under the kernels of step k
    for p in [issuer.submit(batch), issuer.submit(batch)]:
        p.wait()
    barrier()
    with clocks.window():
        t0 = time.perf_counter()
        pend = []
        for _ in range(args.steps):
            pend.append(issuer.submit(batch))
            if len(pend) == 2:
                assert not pend.pop(0).wait().any()
        while pend:
            assert not pend.pop(0).wait().any()
        e2e_stream_s = time.perf_counter() - t0
    barrier()
    clocks.stop()

    secondary = None
    if world == 1 and not args.no_secondary:
        secondary = secondary_measurements(torch, issuer, items, local, stream, flush, B, min(args.steps, 3))

    t = torch.tensor([dev_ms, e2e_s * 1e3, e2e_soa_s * 1e3, e2e_stream_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_ms_max, e2e_soa_ms_max, e2e_stream_ms_max = float(t[0]), float(t[1]), float(t[2]), float(t[3])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    total_items = world * B * args.steps
    value = total_items / (dev_ms_max * 1e-3)
    e2e_value = total_items / (e2e_ms_max * 1e-3)
    wm = work_model(kinds)
    clk = clocks.summary()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    sm_max = clk.get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0
    imad_peak = 148 * 64 * sm_max * 1e6                       # IMAD issue slots/s at max clock (measured 18.52e12 by tools/microbench)
    # dominant kernel: k_ladders = the aMAC ladder + every constraint MSM of the batch in one launch (stages "amac" + "msm")
    msm_ms = (stage_sum["msm"] + stage_sum["amac"]) / args.steps
    msm_imad = 2 * (wm["msm"] + wm["amac"]) * B               # algorithmic IMAD slots per launch of k_ladders
    achieved = msm_imad / (msm_ms * 1e-3)
    traffic = None
    try:   # DRAM bytes per launch of the same kernel from the committed `ncu --set full` capture (profiles/), not a live number
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get("k_ladders")
    except (OSError, ValueError):
        pass
    hbm_bytes = B * (WORDS * 32 + 1)
    roofline = {"bound": "imad", "kernel": "k_ladders", "achieved": achieved / 1e12, "peak": imad_peak / 1e12, "unit": "TIMAD/s", "frac": achieved / imad_peak,
                "traffic": traffic,
                "traffic_note": "DRAM bytes of one k_ladders launch (ncu): ~25 GB are the aMAC ladder's constant-address table scans (every entry of "
                                "every per-item table is read at every step so that no address depends on an issuer secret), the rest per-item "
                                "ladder tables; 15 % of DRAM throughput, the kernel is bound by the IMAD pipe (fma-heavy 81 % busy)", "peak_source": "148 SMs x 64 IMAD/clk x sm_max_mhz; tools/microbench measured 18.52 T IMAD/s and 9.12 T IMAD.WIDE/s (profiles/r01_microbench_imad.json)",
                "algorithmic_imad_per_item": {k: 2 * v for k, v in wm.items()},
                "stage_ms_per_step": {k: v / args.steps for k, v in stage_sum.items()},
                "stage_frac_of_imad_peak": {"k_points": (2 * wm["points"] * B / (stage_sum["points"] / args.steps * 1e-3)) / imad_peak,
                                            "k_ladders": achieved / imad_peak},
                "pipeline_frac_of_imad_peak": (B * args.steps * 2 * wm["total"] / (dev_ms * 1e-3)) / imad_peak,
                "pipeline_frac_at_observed_clock": ((B * args.steps * 2 * wm["total"] / (dev_ms * 1e-3)) / (148 * 64 * clk["sm_mhz"] * 1e6)) if clk.get("sm_mhz") else None,
                "hbm": {"algorithmic_GBps": hbm_bytes * args.steps / (dev_ms * 1e-3) / 1e9, "peak_GBps": peaks.get("hbm_gbs", 6650.0),
                        "note": "non-binding: 897 B of input/output per presentation"}}
    line = {"metric": "presentations_verified_per_sec", "value": value, "unit": "presentations/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 (8x32-bit limbs, IMAD.WIDE carry chains)",
            "data": "synthetic", "config": {"workload": "batch Issuer::verify of 65,536 README-4 presentations [SS,PS,PP,SP] per GPU (BASELINE configs[1])",
                                            "batch_per_gpu": B, "kinds": list(kinds), "bytes_per_item": WORDS * 32,
                                            "l2": "256 MiB flush write between timed steps; 1.2 GB workspace per step exceeds L2",
                                            "input": input_note},
            "clocks": clk, "gpu_launches": launches,
            "e2e": {"value": e2e_value, "unit": "presentations/s", "h2d_bytes_per_step": B * WORDS * 32, "d2h_bytes_per_step": B, "ms_per_step": e2e_ms_max / args.steps,
                    "api": "afx_verify_presentations_wire: item-major bytes in pinned host memory -> verdict bytes in host memory",
                    "soa_api_value": total_items / (e2e_soa_ms_max * 1e-3),
                    "streamed_value": total_items / (e2e_stream_ms_max * 1e-3),
                    "streamed_api": "afx_verify_presentations_submit / afx_wait, two submissions in flight (same per-step H2D and D2H bytes; "
                                    "the copy of step k+1 overlaps the kernels of step k)"},
            "roofline": roofline}
    if secondary:
        line["secondary"] = secondary
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sample = min(B, cores * 8192)                  # ~10 s of CPU work at ~1 k presentations/s/core
        cv, rate, wall = cpu_leg(sp, ip, sk, items, sample, cores)
        gv = issuer.verify_batch(PresentationBatch.from_items(kinds, items[:sample]))
        assert (cv == gv).all(), "GPU verdicts differ from the CPU restatement"
        _, rate1, _ = cpu_leg(sp, ip, sk, items, max(sample // cores, 256), 1)
        line["cpu_baseline"] = {"value": rate, "unit": "presentations/s", "cores": cores, "kind": "port", "single_core_value": rate1,
                                "sample": "%d items of the same batch, %.1f s, C restatement of the reference CPU schedule (oracle/c), verdicts cross-checked with the GPU" % (sample, wall)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
