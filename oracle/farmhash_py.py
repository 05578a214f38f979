"""oracle/farmhash_py.py -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).

Pure-Python restatement of FarmHash ``Fingerprint64`` (== farmhashna ``Hash64``; google/farmhash, MIT licence),
the hash behind ``tf.feature_column.categorical_column_with_hash_bucket`` (tf.strings.to_hash_bucket_fast),
which the reference's examples use for the user / movie id slots
(examples/train_deepfm_on_movielens_keras.py:12-13,20-21).  FarmHash is a third-party dependency of TensorFlow,
not vendored in /root/reference and not installed here, so the published algorithm is restated.

Pinned by (tests/test_oracle.py, tests/test_cpu_boundary.py):
  * Fingerprint64("") = k2 = 0x9ae16a3b2f90404f (by construction of HashLen0to16),
  * Fingerprint64("abc") = 2640714258260161385, Fingerprint64("hello") = 13009744463427800296 (pyfarmhash docs),
  * tf.strings.to_hash_bucket_fast(["Hello", "TensorFlow", "2.x"], 3) = [0, 2, 2] (TensorFlow API docs).
No offline known answer exists for inputs longer than 64 bytes (the streaming branch); there the check is
agreement between this file and the separately written C twin (deep_recommenders_b200/csrc/farmhash.cuh).
"""
import struct

_M = (1 << 64) - 1
_K0 = 0xC3A5C85C97CB3127
_K1 = 0xB492B66FBE98F273
_K2 = 0x9AE16A3B2F90404F


def _rot(v: int, s: int) -> int:
    return v if s == 0 else ((v >> s) | (v << (64 - s))) & _M


def _smix(v: int) -> int:
    return v ^ (v >> 47)


def _f64(b: bytes, i: int) -> int:
    return struct.unpack_from("<Q", b, i)[0]


def _f32(b: bytes, i: int) -> int:
    return struct.unpack_from("<I", b, i)[0]


def _hl16(u: int, v: int, mul: int) -> int:
    a = ((u ^ v) * mul) & _M
    a ^= a >> 47
    b = ((v ^ a) * mul) & _M
    b ^= b >> 47
    return (b * mul) & _M


def _h0to16(s: bytes) -> int:
    n = len(s)
    if n >= 8:
        mul = (_K2 + n * 2) & _M
        a = (_f64(s, 0) + _K2) & _M
        b = _f64(s, n - 8)
        c = (_rot(b, 37) * mul + a) & _M
        d = ((_rot(a, 25) + b) * mul) & _M
        return _hl16(c, d, mul)
    if n >= 4:
        mul = (_K2 + n * 2) & _M
        a = _f32(s, 0)
        return _hl16((n + (a << 3)) & _M, _f32(s, n - 4), mul)
    if n > 0:
        a, b, c = s[0], s[n >> 1], s[n - 1]
        y = (a + (b << 8)) & 0xFFFFFFFF
        z = (n + (c << 2)) & 0xFFFFFFFF
        return (_smix(((y * _K2) & _M) ^ ((z * _K0) & _M)) * _K2) & _M
    return _K2


def _h17to32(s: bytes) -> int:
    n = len(s)
    mul = (_K2 + n * 2) & _M
    a = (_f64(s, 0) * _K1) & _M
    b = _f64(s, 8)
    c = (_f64(s, n - 8) * mul) & _M
    d = (_f64(s, n - 16) * _K2) & _M
    return _hl16((_rot((a + b) & _M, 43) + _rot(c, 30) + d) & _M,
                 (a + _rot((b + _K2) & _M, 18) + c) & _M, mul)


def _h33to64(s: bytes) -> int:
    n = len(s)
    mul = (_K2 + n * 2) & _M
    a = (_f64(s, 0) * _K2) & _M
    b = _f64(s, 8)
    c = (_f64(s, n - 8) * mul) & _M
    d = (_f64(s, n - 16) * _K2) & _M
    y = (_rot((a + b) & _M, 43) + _rot(c, 30) + d) & _M
    z = _hl16(y, (a + _rot((b + _K2) & _M, 18) + c) & _M, mul)
    e = (_f64(s, 16) * mul) & _M
    f = _f64(s, 24)
    g = ((y + _f64(s, n - 32)) * mul) & _M
    h = ((z + _f64(s, n - 24)) * mul) & _M
    return _hl16((_rot((e + f) & _M, 43) + _rot(g, 30) + h) & _M,
                 (e + _rot((f + a) & _M, 18) + g) & _M, mul)


def _weak32(s: bytes, i: int, a: int, b: int):
    w, x, y, z = _f64(s, i), _f64(s, i + 8), _f64(s, i + 16), _f64(s, i + 24)
    a = (a + w) & _M
    b = _rot((b + a + z) & _M, 21)
    c = a
    a = (a + x) & _M
    a = (a + y) & _M
    b = (b + _rot(a, 44)) & _M
    return (a + z) & _M, (b + c) & _M


def fingerprint64(s: bytes) -> int:
    """farmhash::Fingerprint64 (farmhashna::Hash64)."""
    n = len(s)
    if n <= 16:
        return _h0to16(s)
    if n <= 32:
        return _h17to32(s)
    if n <= 64:
        return _h33to64(s)
    seed = 81
    x = seed
    y = (seed * _K1 + 113) & _M
    z = (_smix((y * _K2 + 113) & _M) * _K2) & _M
    v = (0, 0)
    w = (0, 0)
    x = (x * _K2 + _f64(s, 0)) & _M
    end = ((n - 1) // 64) * 64
    last64 = end + ((n - 1) & 63) - 63
    i = 0
    while True:
        x = (_rot((x + y + v[0] + _f64(s, i + 8)) & _M, 37) * _K1) & _M
        y = (_rot((y + v[1] + _f64(s, i + 48)) & _M, 42) * _K1) & _M
        x ^= w[1]
        y = (y + v[0] + _f64(s, i + 40)) & _M
        z = (_rot((z + w[0]) & _M, 33) * _K1) & _M
        v = _weak32(s, i, (v[1] * _K1) & _M, (x + w[0]) & _M)
        w = _weak32(s, i + 32, (z + w[1]) & _M, (y + _f64(s, i + 16)) & _M)
        z, x = x, z
        i += 64
        if i == end:
            break
    mul = (_K1 + ((z & 0xFF) << 1)) & _M
    i = last64
    w = ((w[0] + ((n - 1) & 63)) & _M, w[1])
    v = ((v[0] + w[0]) & _M, v[1])
    w = ((w[0] + v[0]) & _M, w[1])
    x = (_rot((x + y + v[0] + _f64(s, i + 8)) & _M, 37) * mul) & _M
    y = (_rot((y + v[1] + _f64(s, i + 48)) & _M, 42) * mul) & _M
    x ^= (w[1] * 9) & _M
    y = (y + v[0] * 9 + _f64(s, i + 40)) & _M
    z = (_rot((z + w[0]) & _M, 33) * mul) & _M
    v = _weak32(s, i, (v[1] * mul) & _M, (x + w[0]) & _M)
    w = _weak32(s, i + 32, (z + w[1]) & _M, (y + _f64(s, i + 16)) & _M)
    z, x = x, z
    return _hl16((_hl16(v[0], w[0], mul) + (_smix(y) * _K0) + z) & _M,
                 (_hl16(v[1], w[1], mul) + x) & _M, mul)


def hash_bucket_py(values, num_buckets: int):
    """to_hash_bucket_fast over python values (ints are hashed through their decimal string, as TF does)."""
    out = []
    for v in values:
        if isinstance(v, bytes):
            b = v
        elif isinstance(v, str):
            b = v.encode("utf-8")
        else:
            b = str(int(v)).encode()
        out.append(fingerprint64(b) % num_buckets)
    return out
