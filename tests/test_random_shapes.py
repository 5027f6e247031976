"""Randomly drawn attribute shapes (1..8 attributes plus the maximum, 32; scalars / points / plaintexts, random hidden sets --
including the shapes that the reference's compacted-index quirk makes unverifiable) through all four batch
operations against the oracle: Issuer::verify (with Z / commitment / challenge traces), AnonymousCredential::show,
CredentialIssuance::verify and Issuer::issue.  CPU: the test-only host emulation of the engine; -m gpu: the CUDA library."""
import numpy as np
import pytest


def run_shapes(coracle, binding, seed, trials, count, device_kw):
    from aeonflux_b200 import Issuer, PresentationBatch, RequestBatch
    from tests.common import compare_with_oracle_trace
    rng = np.random.default_rng(seed)
    seen_reject = seen_accept = 0
    for trial in range(trials):
        n = 32 if trial == 0 else int(rng.integers(1, 9))
        rk = bytes(rng.choice([ord("S"), ord("P"), ord("E")], n).tolist())
        hide = [i for i in range(n) if rk[i] != ord("P") and rng.random() < 0.5]
        sp, ip, sk = coracle.make_issuer(n)
        orc = coracle.Issuer(sp, ip, sk)
        kinds, pres, issu, showin = orc.synth(rk, hide, b"rand-%d-%d" % (seed, trial), 0, count, want_show_inputs=True)
        pres[count // 2, 1, 7] ^= 4                                              # one corrupted response
        iss = Issuer(sp, ip, sk, max_batch=max(2, count // 2 + 1), _binding=binding, **device_kw)
        v, dbg = iss.verify_batch(PresentationBatch.from_items(kinds, pres), debug=True)
        ov, _, tr = orc.verify_presentations(kinds, pres, trace=True)
        compare_with_oracle_trace(v, dbg, ov, tr)
        assert ov[count // 2] == 1
        seen_reject += int(ov[0] == 1); seen_accept += int(ov[0] == 0)
        pres[count // 2, 1, 7] ^= 4
        if not ov[0]:            # BatchableProof form of the same (honest) presentations: exact and random-linear-combination checks
            from aeonflux_b200 import compact_to_batchable
            comp = PresentationBatch.from_items(kinds, pres)
            _, cdbg = iss.verify_batch(comp, debug=True)
            bf = compact_to_batchable(kinds, comp.fields, cdbg["commitments"])
            assert not iss.verify_batchable(PresentationBatch(kinds, bf)).any()
            vr, fb = iss.verify_batchable_rlc(PresentationBatch(kinds, bf), bytes([trial] * 32))
            assert not vr.any() and fb == 0, (n, rk, hide)
            bf[int(rng.integers(0, bf.shape[0])), count - 1, int(rng.integers(0, 32))] ^= 1 << int(rng.integers(0, 8))
            expect = np.zeros(count, np.uint8); expect[count - 1] = 1
            assert (iss.verify_batchable(PresentationBatch(kinds, bf)) == expect).all()
            vr, fb = iss.verify_batchable_rlc(PresentationBatch(kinds, bf), bytes([trial] * 32))
            assert (vr == expect).all() and fb >= 1, (n, rk, hide)
        res, st = iss.show_batch(kinds, np.ascontiguousarray(showin.transpose(1, 0, 2)))
        assert not st.any() and (res.fields.transpose(1, 0, 2) == pres).all(), (n, rk, hide)
        ik = bytes(0 if c == ord("S") else 2 for c in rk)
        vi = iss.verify_issuance_batch(PresentationBatch.from_items(ik, issu))
        ovi, _ = orc.verify_issuances(ik, issu)
        assert (vi == ovi).all()
        R = rng.integers(0, 256, (count, n + 7, 64), dtype=np.uint8)
        A = np.ascontiguousarray(issu[:, :n])
        out, status, _ = orc.issue(ik, A, R)
        ires, ist = iss.issue_batch(RequestBatch.from_request(ik, A, R))
        assert (ist == status).all() and (ires.fields.transpose(1, 0, 2)[:, n:] == out).all(), (n, rk)
        iss.close()
    return seen_accept, seen_reject


def test_random_shapes_on_emulation(coracle):
    import ctypes
    from aeonflux_b200._binding import Binding
    from tests.test_host_logic import build_hostemu
    acc, rej = run_shapes(coracle, Binding(ctypes.CDLL(build_hostemu())), seed=1, trials=6, count=3, device_kw={})
    assert acc >= 2                                    # and at least the honest shapes verify


@pytest.mark.gpu
def test_random_shapes_on_gpu(coracle):
    acc, rej = run_shapes(coracle, None, seed=2, trials=14, count=150, device_kw={"device": 0})
    assert acc >= 4 and rej >= 1
