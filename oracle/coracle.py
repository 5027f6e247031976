"""ORACLE (test infrastructure, NOT the product): ctypes loader for oracle/_build/libafx_oracle.so
(the C restatement in the reference's CPU schedule, oracle/c/afx_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libafx_oracle.so")
_lib = None

KIND_PS, KIND_SS, KIND_PP, KIND_SP = 0, 1, 2, 3


def build(force=False):
    """Compile the C oracle (gcc) if missing or stale."""
    srcs = [os.path.join(_HERE, "c", f) for f in ("afx_oracle.c", "prim.h", "consts.h")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE, "_build/libafx_oracle.so"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        u8p, c_int, c_u64, c_dbl = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64, ctypes.c_double
        L.afxo_issuer_new.restype = ctypes.c_void_p
        L.afxo_issuer_new.argtypes = [u8p, ctypes.c_size_t, u8p, u8p, ctypes.c_size_t]
        L.afxo_issuer_free.argtypes = [ctypes.c_void_p]
        L.afxo_make_issuer.argtypes = [ctypes.c_uint32, ctypes.c_char_p, u8p, u8p, u8p]
        L.afxo_verify_presentations.restype = c_dbl
        L.afxo_verify_presentations.argtypes = [ctypes.c_void_p, u8p, c_int, u8p, c_u64, c_int, u8p, u8p, u8p, c_int, u8p, c_int]
        L.afxo_synth.restype = c_dbl
        L.afxo_synth.argtypes = [ctypes.c_void_p, u8p, u8p, c_int, ctypes.c_char_p, c_u64, c_u64, c_int, u8p, u8p, u8p]
        L.afxo_verify_issuances.restype = c_dbl
        L.afxo_verify_issuances.argtypes = [ctypes.c_void_p, u8p, c_int, u8p, c_u64, c_int, u8p, u8p, u8p]
        L.afxo_issue.restype = c_dbl
        L.afxo_issue.argtypes = [ctypes.c_void_p, u8p, c_int, u8p, u8p, c_u64, c_int, u8p, u8p]
        L.afxo_decompress_compress.argtypes = [u8p, u8p]
        L.afxo_from_uniform.argtypes = [u8p, u8p]
        L.afxo_scalarmult.argtypes = [u8p, u8p, u8p, c_int]
        L.afxo_sc_from_wide.argtypes = [u8p, u8p]
        L.afxo_sc_muladd.argtypes = [u8p, u8p, u8p, u8p]
        L.afxo_sysparams_size.argtypes = [ctypes.c_uint32]
        L.afxo_secret_size.argtypes = [ctypes.c_uint32]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _buf(b):
    return np.frombuffer(bytes(b), dtype=np.uint8).copy()


def presentation_words(kinds) -> int:
    kinds = list(kinds)
    n = len(kinds)
    return (1 + 3 + sum(k == KIND_SS for k in kinds) + 3 + n + sum(k in (KIND_PS, KIND_PP) for k in kinds)
            + 14 * sum(k == KIND_SP for k in kinds))


def presentation_counts(kinds):
    """-> (#commitments, #proofs) a full verification of this shape recomputes."""
    kinds = list(kinds)
    h_p = sum(k == KIND_SP for k in kinds)
    n_nsp = len(kinds) - h_p
    main = 2 + sum(1 for i in range(n_nsp) if kinds[i] != KIND_SP)  # compacted-index loop
    return main + 5 * h_p, 1 + h_p


def make_issuer(n: int, tag: bytes = b"issuer"):
    """-> (sysparams bytes, issuer_pub 64 B, secret bytes); same bytes as pyoracle.synth.make_issuer."""
    L = lib()
    sp = np.zeros(L.afxo_sysparams_size(n), np.uint8)
    ip = np.zeros(64, np.uint8)
    sk = np.zeros(L.afxo_secret_size(n), np.uint8)
    assert L.afxo_make_issuer(n, tag, _p(sp), _p(ip), _p(sk)) == 0
    return sp.tobytes(), ip.tobytes(), sk.tobytes()


class Issuer:
    def __init__(self, sysparams: bytes, issuer_pub: bytes, secret: bytes = None):
        L = lib()
        self.n = int.from_bytes(sysparams[:4], "little")
        sp, ip = _buf(sysparams), _buf(issuer_pub)
        sk = _buf(secret) if secret is not None else None
        self._h = L.afxo_issuer_new(_p(sp), len(sp), _p(ip), _p(sk), 0 if sk is None else len(sk))
        if not self._h:
            raise ValueError("bad issuer encoding")

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:      # at interpreter shutdown the module globals may already be gone
            _lib.afxo_issuer_free(self._h)
            self._h = None

    def synth(self, request_kinds: bytes, hide, config: bytes, start: int, count: int, threads: int = 0, want_issuances=True,
              want_show_inputs=False):
        """request_kinds: bytes of b'S' / b'P' / b'E'.  -> (presentation kinds, presentations [count][W][32] u8,
        issuances [count][Wi][32] u8); with want_show_inputs also the flat input of the batch AnonymousCredential::show that
        produces exactly those presentations ([count][Ws][32]: credential, attributes, keypair, rng bytes)."""
        n = self.n
        assert len(request_kinds) == n
        hide_flags = np.zeros(n, np.uint8)
        for i in hide:
            hide_flags[i] = 1
        kinds = []
        for i, k in enumerate(request_kinds):
            if k == ord("S"):
                kinds.append(KIND_SS if hide_flags[i] else KIND_PS)
            elif k == ord("P"):
                kinds.append(KIND_PP)
            else:
                kinds.append(KIND_SP if hide_flags[i] else KIND_PP)
        W = presentation_words(kinds)
        Wi = n + 3 + 1 + n + 5
        pres = np.zeros((count, W, 32), np.uint8)
        iss = np.zeros((count, Wi, 32), np.uint8) if want_issuances else None
        rk = _buf(request_kinds)
        show = None
        if want_show_inputs:
            hs, hp = sum(k == KIND_SS for k in kinds), sum(k == KIND_SP for k in kinds)
            Ws = 3 + sum(3 if k == KIND_SP else 1 for k in kinds) + (4 if hp else 0) + 2 * (1 + 3 + hs + 6 * hp)
            show = np.zeros((count, Ws, 32), np.uint8)
        t = lib().afxo_synth(self._h, _p(rk), _p(hide_flags), n, config, start, count, threads or os.cpu_count(), _p(pres), _p(iss), _p(show))
        assert t >= 0
        if want_show_inputs:
            return bytes(kinds), pres, iss, show
        return bytes(kinds), pres, iss

    def verify_presentations(self, kinds: bytes, items: np.ndarray, threads: int = 0, trace=False):
        """items [count][W][32] u8 -> verdicts u8[count] (+ trace dict) ; also returns seconds inside verify."""
        count = items.shape[0]
        assert items.dtype == np.uint8 and items.shape[1:] == (presentation_words(kinds), 32) and items.flags.c_contiguous
        verdicts = np.zeros(count, np.uint8)
        kb = _buf(kinds)
        ncm, nch = presentation_counts(kinds)
        tz = np.zeros((count, 32), np.uint8) if trace else None
        tc = np.zeros((count, ncm, 32), np.uint8) if trace else None
        th = np.zeros((count, nch, 32), np.uint8) if trace else None
        secs = lib().afxo_verify_presentations(self._h, _p(kb), len(kinds), _p(items), count, threads or os.cpu_count(),
                                               _p(verdicts), _p(tz), _p(tc), ncm, _p(th), nch)
        assert secs >= 0
        if trace:
            return verdicts, secs, {"Z": tz, "commitments": tc, "challenges": th}
        return verdicts, secs

    def verify_issuances(self, kinds: bytes, items: np.ndarray, threads: int = 0, trace=False):
        count = items.shape[0]
        n = len(kinds)
        assert items.dtype == np.uint8 and items.shape[1:] == (2 * n + 9, 32) and items.flags.c_contiguous
        verdicts = np.zeros(count, np.uint8)
        kb = _buf(kinds)
        tc = np.zeros((count, 3, 32), np.uint8) if trace else None
        th = np.zeros((count, 1, 32), np.uint8) if trace else None
        secs = lib().afxo_verify_issuances(self._h, _p(kb), n, _p(items), count, threads or os.cpu_count(), _p(verdicts), _p(tc), _p(th))
        assert secs >= 0
        if trace:
            return verdicts, secs, {"commitments": tc, "challenges": th}
        return verdicts, secs

    def issue(self, kinds: bytes, attrs: np.ndarray, randomness: np.ndarray, threads: int = 0):
        """attrs [count][n][32]; randomness [count][2+n+5][64] -> (out [count][3+1+n+5][32] = t,U,V,c,responses; status)"""
        count, n = attrs.shape[0], len(kinds)
        assert attrs.shape == (count, n, 32) and randomness.shape == (count, n + 7, 64)
        out = np.zeros((count, n + 9, 32), np.uint8)
        status = np.zeros(count, np.uint8)
        kb = _buf(kinds)
        secs = lib().afxo_issue(self._h, _p(kb), n, _p(np.ascontiguousarray(attrs)), _p(np.ascontiguousarray(randomness)), count,
                                threads or os.cpu_count(), _p(out), _p(status))
        assert secs >= 0
        return out, status, secs
