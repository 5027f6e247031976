"""Writes bench_data/readme4_1024.bin: 1,024 distinct honest README-4 presentations (kinds [SS,PS,PP,SP], 28 words of 32 B
each, item-major) plus the issuer they verify under (bench_data/issuer4.bin = sysparams || C_W||I || secret key), generated
with the oracle's reference-schedule prover from the fixed SHAKE-256 seeds of SURVEY 8d (config "bench-readme4").
bench.py tiles these to the 65,536-item batch of BASELINE config 2: every item is independent and every stage's control flow
is data-independent, so tiling changes neither the work nor the memory traffic (each item still owns its own workspace rows).
Also writes bench_data/s16_256.bin / issuer16.bin: 256 honest S16 presentations (kinds [SS,SS,PS,PS,PS,PS,PP,PP,SPx8], 143 words,
BASELINE config 4) for bench.py's secondary measurements.
Run from the repo root:  python tests/golden/make_bench_fixture.py   (test infrastructure: it uses the oracle; bench.py only reads the files)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import coracle as C  # noqa: E402

HERE = os.path.join(ROOT, "bench_data")
sp, ip, sk = C.make_issuer(4)
iss = C.Issuer(sp, ip, sk)
kinds, pres, _ = iss.synth(b"SSPE", [0, 3], b"bench-readme4", 0, 1024, want_issuances=False)
assert list(kinds) == [1, 0, 2, 3]
v, _ = iss.verify_presentations(kinds, pres)
assert not v.any()
pres.tofile(os.path.join(HERE, "readme4_1024.bin"))
open(os.path.join(HERE, "issuer4.bin"), "wb").write(sp + ip + sk)
print("wrote", pres.shape, len(sp), len(ip), len(sk))

# one symmetric keypair for the README-4 issuer (a, a0, a1, pk), used by bench.py when it synthesizes presentations on the device
from oracle.pyoracle import aeonflux as A, ristretto as R  # noqa: E402
kp = A.SymmetricKeypair.derive(b"aeonflux-b200 bench keypair".ljust(64, b"\0"), A.SystemParameters.from_bytes(sp))
open(os.path.join(HERE, "keypair4.bin"), "wb").write(R.sc_to_bytes(kp.a) + R.sc_to_bytes(kp.a0) + R.sc_to_bytes(kp.a1) + kp.pk.compress())
sp, ip, sk = C.make_issuer(16)
iss = C.Issuer(sp, ip, sk)
kinds, pres, _ = iss.synth(b"SSSSSSPP" + b"E" * 8, [0, 1] + list(range(8, 16)), b"bench-s16", 0, 256, want_issuances=False)
assert list(kinds) == [1, 1, 0, 0, 0, 0, 2, 2] + [3] * 8
v, _ = iss.verify_presentations(kinds, pres)
assert not v.any()
pres.tofile(os.path.join(HERE, "s16_256.bin"))
open(os.path.join(HERE, "issuer16.bin"), "wb").write(sp + ip + sk)
kp = A.SymmetricKeypair.derive(b"aeonflux-b200 bench keypair".ljust(64, b"\0"), A.SystemParameters.from_bytes(sp))
open(os.path.join(HERE, "keypair16.bin"), "wb").write(R.sc_to_bytes(kp.a) + R.sc_to_bytes(kp.a0) + R.sc_to_bytes(kp.a1) + kp.pk.compress())
print("wrote", pres.shape, len(sp), len(ip), len(sk))
