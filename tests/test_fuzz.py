"""Differential fuzz: mutated wire input (tests/common.py:fuzz_mutations -- edge scalars and point encodings, random bytes, valid
words in the wrong place, bit flips incl. the top byte) through the engine and the C oracle; verdicts and, as far as the reference's
early return gets, Z / commitments / challenges must be identical.  CPU: the test-only host emulation, plain and under
AddressSanitizer + UndefinedBehaviorSanitizer (tests/tools/fuzz_asan.py); -m gpu: the CUDA library through the C ABI."""
import ctypes
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def coracle():
    from oracle import coracle as C
    C.build()
    return C


def test_differential_fuzz_on_emulation(coracle):
    from aeonflux_b200 import Issuer
    from aeonflux_b200._binding import Binding
    from tests.common import differential_fuzz
    from tests.test_host_logic import build_hostemu
    emu = Binding(ctypes.CDLL(build_hostemu()))
    acc, rej = differential_fuzz(lambda sp, ip, sk: Issuer(sp, ip, sk, max_batch=40, _binding=emu), coracle, 21, 100)
    assert rej > 300 and acc + rej == 400


def test_emulation_under_address_and_ub_sanitizers():
    """One short round of tests/tools/fuzz_asan.py (the long form is run by hand: `python tests/tools/fuzz_asan.py 600`)."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tools", "fuzz_asan.py"), "0", "light"], capture_output=True, text=True, timeout=900)
    if r.returncode == 77:
        pytest.skip("no AddressSanitizer runtime in this image")
    assert r.returncode == 0 and "fuzz_asan ok" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])


@pytest.mark.gpu
def test_differential_fuzz_on_gpu(coracle):
    """4,096 mutated presentations and issuances per shape (README-4, a 3-attribute shape, S16) on the GPU vs the C oracle."""
    from aeonflux_b200 import Issuer
    from tests.common import differential_fuzz
    shapes = ((4, b"SSPE", [0, 3]), (3, b"SPE", [2]), (16, b"SSSSSSPP" + b"E" * 8, [0, 1] + list(range(8, 16))))
    acc, rej = differential_fuzz(lambda sp, ip, sk: Issuer(sp, ip, sk, device=0, max_batch=4096), coracle, 31, 4096, shapes=shapes)
    assert rej > 20000 and acc + rej == 6 * 4096
