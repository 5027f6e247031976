"""Byte layouts of the issuer-side key material, parsed and emitted exactly as the reference's to_bytes() methods lay them
out, plus the pieces the reference leaves unimplemented or broken:

  SystemParameters::to_bytes   /root/reference/src/parameters.rs:155-184   n:u32le || G || G_w || G_w' || G_x_0 || G_x_1 ||
                               (size :34-40)                                G_y x max(n,3) || G_m x n || G_V || G_a || G_a0 || G_a1
  SecretKey::to_bytes          src/amacs.rs:110-125 (size :44-46)          n:u32le || w || w' || x_0 || x_1 || y x n || W
  SecretKey::from_bytes        src/amacs.rs:128-155 never advances `chunk` inside the y loop (:148-150), so every y_i
                               deserializes as y_0; `split_secret_key` reads each y_i from its own 32 bytes.
  IssuerParameters::{to,from}_bytes  src/parameters.rs:365-372 are unimplemented!(); the 64 bytes Issuer::to_bytes reserves
                               (src/issuer.rs:155,163) are C_W || I.
  Issuer::{to,from}_bytes      src/issuer.rs:152-174 = sysparams || issuer params || secret key (panics in the reference
                               because of the unimplemented!() above).

Pure byte bookkeeping: no curve arithmetic happens here; encodings are validated on the device by afx_ctx_create.
"""
import struct

L = 2**252 + 27742317777372353535851937790883648493


def system_parameters_size(n: int) -> int:
    """parameters.rs:34-40."""
    return 32 * (5 + 3 + n + 4) + 4 if n < 3 else 32 * (5 + 2 * n + 4) + 4


def secret_key_size(n: int) -> int:
    """amacs.rs:44-46."""
    return 32 * (5 + n) + 4


def split_system_parameters(b: bytes) -> dict:
    """SystemParameters::from_bytes (parameters.rs:92-153) without decompression: named 32-byte encodings."""
    if len(b) < 4:
        raise ValueError("NoSystemParameters")
    n = struct.unpack_from("<I", b)[0]
    if n == 0 or len(b) != system_parameters_size(n):
        raise ValueError("NoSystemParameters")
    w = [b[4 + 32 * i: 36 + 32 * i] for i in range((len(b) - 4) // 32)]
    ny = max(n, 3)
    return {"n": n, "G": w[0], "G_w": w[1], "G_w_prime": w[2], "G_x_0": w[3], "G_x_1": w[4], "G_y": w[5:5 + ny], "G_m": w[5 + ny:5 + ny + n],
            "G_V": w[5 + ny + n], "G_a": w[6 + ny + n], "G_a0": w[7 + ny + n], "G_a1": w[8 + ny + n]}


def join_system_parameters(p: dict) -> bytes:
    n = p["n"]
    if len(p["G_y"]) != max(n, 3) or len(p["G_m"]) != n:
        raise ValueError("G_y must have max(n,3) entries and G_m n entries (parameters.rs:235-251)")
    words = [p["G"], p["G_w"], p["G_w_prime"], p["G_x_0"], p["G_x_1"]] + list(p["G_y"]) + list(p["G_m"]) + [p["G_V"], p["G_a"], p["G_a0"], p["G_a1"]]
    if any(len(x) != 32 for x in words):
        raise ValueError("every generator is a 32-byte CompressedRistretto")
    return struct.pack("<I", n) + b"".join(words)


def split_secret_key(b: bytes) -> dict:
    """SecretKey::from_bytes with the y loop fixed: y_i is read from bytes [132 + 32 i, 164 + 32 i).  Scalars must be canonical
    (Scalar::from_canonical_bytes, amacs.rs:141)."""
    if len(b) < 4:
        raise ValueError("MacError::KeypairDeserialisation")
    n = struct.unpack_from("<I", b)[0]
    if n == 0 or len(b) != secret_key_size(n):
        raise ValueError("MacError::KeypairDeserialisation")
    w = [b[4 + 32 * i: 36 + 32 * i] for i in range(5 + n)]
    for s in w[:4 + n]:
        if int.from_bytes(s, "little") >= L:
            raise ValueError("MacError::KeypairDeserialisation")
    return {"n": n, "w": w[0], "w_prime": w[1], "x_0": w[2], "x_1": w[3], "y": w[4:4 + n], "W": w[4 + n]}


def join_secret_key(k: dict) -> bytes:
    words = [k["w"], k["w_prime"], k["x_0"], k["x_1"]] + list(k["y"]) + [k["W"]]
    if len(k["y"]) != k["n"] or any(len(x) != 32 for x in words):
        raise ValueError("malformed secret key")
    return struct.pack("<I", k["n"]) + b"".join(words)


def issuer_parameters_to_bytes(C_W: bytes, I: bytes) -> bytes:
    """IssuerParameters::to_bytes (unimplemented!() in the reference): C_W || I."""
    if len(C_W) != 32 or len(I) != 32:
        raise ValueError("C_W and I are 32-byte CompressedRistretto encodings")
    return bytes(C_W) + bytes(I)


def issuer_parameters_from_bytes(b: bytes):
    if len(b) != 64:
        raise ValueError("IssuerParameters is C_W || I (64 bytes)")
    return bytes(b[:32]), bytes(b[32:])


def issuer_to_bytes(system_parameters: bytes, issuer_parameters: bytes, amacs_key: bytes) -> bytes:
    """Issuer::to_bytes, issuer.rs:162-174."""
    n = split_system_parameters(system_parameters)["n"]
    issuer_parameters_from_bytes(issuer_parameters)
    if split_secret_key(amacs_key)["n"] != n:
        raise ValueError("secret key and system parameters disagree on the number of attributes")
    return bytes(system_parameters) + bytes(issuer_parameters) + bytes(amacs_key)


def issuer_from_bytes(b: bytes):
    """Issuer::from_bytes, issuer.rs:152-159 -> (system_parameters, issuer_parameters, amacs_key) byte strings."""
    if len(b) < 4:
        raise ValueError("NoIssuerParameters")
    n = struct.unpack_from("<I", b)[0]
    a = system_parameters_size(n)
    if n == 0 or len(b) != a + 64 + secret_key_size(n):
        raise ValueError("NoIssuerParameters")
    sp, ip, sk = bytes(b[:a]), bytes(b[a:a + 64]), bytes(b[a + 64:])
    split_system_parameters(sp); split_secret_key(sk)
    return sp, ip, sk
