"""The C++ host-side mirror of the reference's interface (include/aeonflux_b200.hpp) driven through the README flow in batch
form by tests/cpp/host_parity.cpp, against oracle expectations: on the test-only host emulation of the C ABI here (no GPU), and
on the CUDA library under -m gpu."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "host_parity.cpp")


def write_fixture(path, coracle, count):
    sp, ip, sk = coracle.make_issuer(4)
    orc = coracle.Issuer(sp, ip, sk)
    kinds, pres, issu, showin = orc.synth(b"SSPE", [0, 3], b"cpp-host", 0, count, want_show_inputs=True)
    n = 4
    rk = bytes([0, 0, 2, 2])
    attrs = np.ascontiguousarray(issu[:, :n])
    rnd = np.random.default_rng(77).integers(0, 256, (count, n + 7, 64), dtype=np.uint8)
    issued, status, _ = orc.issue(rk, attrs, rnd)
    assert not status.any()
    bad = pres.copy()
    for i in range(0, count, 7):
        bad[i, (3 * i) % 28, 5] ^= 0x20
    ov, _ = orc.verify_presentations(kinds, bad)
    assert ov.sum() == len(range(0, count, 7))
    blob = sp + ip + sk
    with open(path, "wb") as f:
        f.write(struct.pack("<II", n, count))
        f.write(struct.pack("<I", len(blob))); f.write(blob)
        f.write(rk); f.write(attrs.tobytes()); f.write(rnd.tobytes()); f.write(issued.tobytes())
        f.write(kinds); f.write(struct.pack("<II", showin.shape[1], pres.shape[1]))
        f.write(showin.tobytes()); f.write(pres.tobytes()); f.write(bad.tobytes()); f.write(ov.tobytes())


def build_driver(out, lib_dir, lib_name):
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-pthread", "-o", out, SRC, "-L" + lib_dir, "-l" + lib_name, "-Wl,-rpath," + lib_dir])


def test_cpp_host_mirror_on_emulation(tmp_path, coracle):
    from tests.test_host_logic import build_hostemu
    so = build_hostemu()
    fx = tmp_path / "fixture.bin"
    write_fixture(fx, coracle, 20)
    exe = str(tmp_path / "host_parity_emu")
    build_driver(exe, os.path.dirname(so), "afx_hostemu")
    out = subprocess.run([exe, str(fx), "8"], capture_output=True, text=True, timeout=600)     # max_batch 8: three chunks
    assert out.returncode == 0, out.stdout + out.stderr
    assert "host_parity ok: 20 items, 3 rejected" in out.stdout and "host-emulation" in out.stdout


@pytest.mark.gpu
def test_cpp_host_mirror_on_gpu(tmp_path, coracle):
    from aeonflux_b200.build import CSRC
    fx = tmp_path / "fixture.bin"
    write_fixture(fx, coracle, 1000)
    exe = str(tmp_path / "host_parity_cuda")
    build_driver(exe, CSRC, "aeonflux_b200")
    out = subprocess.run([exe, str(fx), "512"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "host_parity ok: 1000 items, 143 rejected" in out.stdout and "cuda sm_100a" in out.stdout


STRESS = os.path.join(ROOT, "tests", "cpp", "thread_stress.cpp")


def test_threads_on_one_context_under_thread_sanitizer(tmp_path, coracle):
    """tests/cpp/thread_stress.cpp: four threads making synchronous calls on one context, then a submit/wait streamer beside two
    synchronous callers -- right verdicts or AFX_ERR_ARG, never a wrong verdict -- and the C++ host mirror incl. the library's
    afx_multi_* device threads (host_parity.cpp), all against an emulation build compiled with ThreadSanitizer."""
    rt = subprocess.run(["gcc", "-print-file-name=libtsan.so"], capture_output=True, text=True).stdout.strip()
    tsan = ["-fsanitize=thread", "-g"] if os.path.isabs(rt) and os.path.exists(rt) else []
    emu_src = os.path.join(ROOT, "tests", "hostemu", "hostemu.cpp")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-pthread", "-fvisibility=hidden"] + tsan + ["-x", "c++", "-o", str(tmp_path / "libafx_hostemu.so"), emu_src])
    fx = tmp_path / "fixture.bin"
    write_fixture(fx, coracle, 20)
    for src, args, expect in ((STRESS, ["8", "3"], "thread_stress ok"), (SRC, ["8"], "host_parity ok: 20 items, 3 rejected")):
        exe = str(tmp_path / os.path.basename(src).replace(".cpp", ""))
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-pthread"] + tsan + ["-o", exe, src, "-L" + str(tmp_path), "-lafx_hostemu", "-Wl,-rpath," + str(tmp_path)])
        out = subprocess.run([exe, str(fx)] + args, capture_output=True, text=True, timeout=900)
        assert out.returncode == 0 and expect in out.stdout and "ThreadSanitizer" not in out.stderr, out.stdout + out.stderr


@pytest.mark.gpu
def test_threads_on_one_context_on_gpu(tmp_path, coracle):
    from aeonflux_b200.build import CSRC
    fx = tmp_path / "fixture.bin"
    write_fixture(fx, coracle, 2000)
    exe = str(tmp_path / "thread_stress_cuda")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-pthread", "-o", exe, STRESS, "-L" + CSRC, "-laeonflux_b200", "-Wl,-rpath," + CSRC])
    out = subprocess.run([exe, str(fx), "600", "20"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "thread_stress ok" in out.stdout and "cuda sm_100a" in out.stdout, out.stdout + out.stderr
