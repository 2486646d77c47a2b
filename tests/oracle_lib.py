"""ctypes binding of the CPU parity oracle (oracle/_build/libszl_oracle.so).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs import this.  The product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "_build", "libszl_oracle.so")


def build(force=False):
    srcs = [os.path.join(_ROOT, "oracle", f) for f in
            ("szl_deflate.cpp", "szl_inflate.cpp", "szl_capi.cpp", "szl_crypto.cpp", "szl_oracle.hpp")]
    if (not force and os.path.exists(_SO)
            and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in srcs if os.path.exists(s))):
        return _SO
    subprocess.check_call(["make", "-s", "-C", os.path.join(_ROOT, "oracle")])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_SO)
    u8p = C.c_void_p
    L.szl_last_error.restype = C.c_char_p
    for name in ("szl_crc32", "szl_adler32", "szl_crc32_bytewise", "szl_adler32_bytewise"):
        f = getattr(L, name)
        f.restype = C.c_uint32
        f.argtypes = [u8p, C.c_uint64]
    for name in ("szl_crc32_update", "szl_adler32_update"):
        f = getattr(L, name)
        f.restype = C.c_uint32
        f.argtypes = [C.c_uint32, u8p, C.c_uint64]
    L.szl_deflate_oneshot.restype = C.c_int
    L.szl_deflate_oneshot.argtypes = [u8p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int32, C.c_int32,
                                      u8p, C.c_int64, C.POINTER(C.c_int64)]
    L.szl_deflate_trace.restype = C.c_int
    L.szl_deflate_trace.argtypes = [u8p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int32,
                                    u8p, C.c_int64, C.POINTER(C.c_int64), u8p, C.c_int32, C.POINTER(C.c_int32)]
    L.szl_inflate_oneshot.restype = C.c_int
    L.szl_inflate_oneshot.argtypes = [u8p, C.c_int64, C.c_int, C.c_int32, u8p, C.c_int64, C.POINTER(C.c_int64),
                                      C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.szl_batch.restype = C.c_int
    L.szl_batch.argtypes = [C.c_int, u8p, u8p, u8p, C.c_int32, C.c_int, C.c_int, u8p, u8p, u8p, u8p, C.c_int32]
    # entry ciphers (oracle/szl_crypto.cpp)
    L.szl_aes_encrypt_block.argtypes = [u8p, C.c_int, u8p, u8p]
    L.szl_sha1.argtypes = [u8p, C.c_uint64, u8p]
    L.szl_hmac_sha1.argtypes = [u8p, C.c_uint64, u8p, C.c_uint64, u8p]
    L.szl_pbkdf2_sha1.argtypes = [u8p, C.c_uint64, u8p, C.c_uint64, C.c_int, u8p, C.c_uint64]
    L.szl_zip_aes.argtypes = [u8p, C.c_uint64, u8p, C.c_int, C.c_int, u8p, C.c_uint64, C.c_uint64, u8p, u8p, u8p]
    L.szl_pkzip_generate_keys.argtypes = [u8p, C.c_uint64, u8p]
    L.szl_pkzip_transform.argtypes = [u8p, C.c_int, u8p, C.c_uint64, u8p]
    for name in ("szl_aes_encrypt_block", "szl_sha1", "szl_hmac_sha1", "szl_pbkdf2_sha1", "szl_zip_aes", "szl_pkzip_generate_keys",
                 "szl_pkzip_transform"):
        getattr(L, name).restype = None
    # handles
    L.szl_deflater_new.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    L.szl_deflater_free.argtypes = [C.c_void_p]
    L.szl_deflater_reset.argtypes = [C.c_void_p]
    L.szl_deflater_set_input.argtypes = [C.c_void_p, u8p, C.c_int32, C.c_int32]
    L.szl_deflater_set_level.argtypes = [C.c_void_p, C.c_int]
    L.szl_deflater_get_level.argtypes = [C.c_void_p]
    L.szl_deflater_set_strategy.argtypes = [C.c_void_p, C.c_int]
    L.szl_deflater_set_dictionary.argtypes = [C.c_void_p, u8p, C.c_int32, C.c_int32]
    L.szl_deflater_flush.argtypes = [C.c_void_p]
    L.szl_deflater_flush.restype = None
    L.szl_deflater_finish.argtypes = [C.c_void_p]
    L.szl_deflater_finish.restype = None
    L.szl_deflater_deflate.argtypes = [C.c_void_p, u8p, C.c_int32, C.c_int32, C.POINTER(C.c_int32)]
    L.szl_deflater_needs_input.argtypes = [C.c_void_p]
    L.szl_deflater_is_finished.argtypes = [C.c_void_p]
    L.szl_deflater_total_in.argtypes = [C.c_void_p]
    L.szl_deflater_total_in.restype = C.c_int64
    L.szl_deflater_total_out.argtypes = [C.c_void_p]
    L.szl_deflater_total_out.restype = C.c_int64
    L.szl_deflater_adler.argtypes = [C.c_void_p]
    L.szl_deflater_adler.restype = C.c_uint32
    L.szl_inflater_new.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    L.szl_inflater_free.argtypes = [C.c_void_p]
    L.szl_inflater_reset.argtypes = [C.c_void_p]
    L.szl_inflater_set_input.argtypes = [C.c_void_p, u8p, C.c_int32, C.c_int32]
    L.szl_inflater_set_dictionary.argtypes = [C.c_void_p, u8p, C.c_int32, C.c_int32]
    L.szl_inflater_inflate.argtypes = [C.c_void_p, u8p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32)]
    for n in ("needs_input", "needs_dictionary", "is_finished", "remaining_input"):
        getattr(L, "szl_inflater_" + n).argtypes = [C.c_void_p]
    L.szl_inflater_total_in.argtypes = [C.c_void_p]
    L.szl_inflater_total_in.restype = C.c_int64
    L.szl_inflater_total_out.argtypes = [C.c_void_p]
    L.szl_inflater_total_out.restype = C.c_int64
    L.szl_inflater_adler.argtypes = [C.c_void_p]
    L.szl_inflater_adler.restype = C.c_uint32
    _lib = L
    return L


class OracleError(Exception):
    def __init__(self, kind, msg):
        super().__init__("oracle error kind=%d: %s" % (kind, msg))
        self.kind = kind
        self.msg = msg


def _check(rc):
    if rc != 0:
        raise OracleError(rc, lib().szl_last_error().decode("latin1"))


def _arr(data):
    a = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    return np.ascontiguousarray(a, dtype=np.uint8)


def _ptr(a):
    return a.ctypes.data if a.size else None


def crc32(data):
    a = _arr(data)
    return lib().szl_crc32(a.ctypes.data, a.size)


def adler32(data):
    a = _arr(data)
    return lib().szl_adler32(a.ctypes.data, a.size)


def crc32_update(value, data):
    a = _arr(data)
    return lib().szl_crc32_update(value, a.ctypes.data, a.size)


def adler32_update(value, data):
    a = _arr(data)
    return lib().szl_adler32_update(value, a.ctypes.data, a.size)


def aes_encrypt_block(key, block16):
    k, b, o = _arr(key), _arr(block16), np.zeros(16, np.uint8)
    lib().szl_aes_encrypt_block(k.ctypes.data, k.size, b.ctypes.data, o.ctypes.data)
    return o.tobytes()


def sha1(data):
    a, o = _arr(data), np.zeros(20, np.uint8)
    lib().szl_sha1(_ptr(a), a.size, o.ctypes.data)
    return o.tobytes()


def hmac_sha1(key, data):
    k, a, o = _arr(key), _arr(data), np.zeros(20, np.uint8)
    lib().szl_hmac_sha1(_ptr(k), k.size, _ptr(a), a.size, o.ctypes.data)
    return o.tobytes()


def pbkdf2_sha1(password, salt, rounds, n):
    p, s, o = _arr(password), _arr(salt), np.zeros(n, np.uint8)
    lib().szl_pbkdf2_sha1(_ptr(p), p.size, _ptr(s), s.size, rounds, o.ctypes.data, n)
    return o.tobytes()


def zip_aes(password, salt, block_size, write_mode, data, piece=0):
    """new ZipAESTransform(password, salt, blockSize, writeMode).TransformBlock(data) -> (bytes, PwdVerifier, GetAuthCode())"""
    p, s, a = _arr(password), _arr(salt), _arr(data)
    o, v, h = np.zeros(a.size, np.uint8), np.zeros(2, np.uint8), np.zeros(20, np.uint8)
    lib().szl_zip_aes(_ptr(p), p.size, s.ctypes.data, block_size, 1 if write_mode else 0, _ptr(a), a.size, piece, _ptr(o),
                      v.ctypes.data, h.ctypes.data)
    return o.tobytes(), v.tobytes(), h.tobytes()


def pkzip_generate_keys(seed):
    s, o = _arr(seed), np.zeros(12, np.uint8)
    lib().szl_pkzip_generate_keys(_ptr(s), s.size, o.ctypes.data)
    return o.tobytes()


def pkzip_transform(keys12, encrypt, data):
    """PkzipClassic{Encrypt,Decrypt}CryptoTransform.TransformBlock -> (bytes, keys after)"""
    k, a = np.array(np.frombuffer(keys12, dtype=np.uint8)), _arr(data)
    o = np.zeros(a.size, np.uint8)
    lib().szl_pkzip_transform(k.ctypes.data, 1 if encrypt else 0, _ptr(a), a.size, _ptr(o))
    return o.tobytes(), k.tobytes()


def dotnet_random_bytes(seed, n):
    """`new System.Random(seed).NextBytes(new byte[n])` -- the data source of the reference's own tests."""
    out = np.empty(n, np.uint8)
    f = lib().szl_dotnet_random_bytes
    f.restype = None
    f.argtypes = [C.c_int32, C.c_void_p, C.c_uint64]
    f(seed, out.ctypes.data, n)
    return out


def deflate_bound(n):
    return n + (n >> 3) + 1024


def deflate(data, level=6, nowrap=True, strategy=0, pattern=0, chunk=0, obuf=512):
    """One-shot reference-pattern deflate.  pattern 0: SetInput*/Finish/drain; 1: Write*/Flush/Finish;
    2: Write*/Finish (DeflaterOutputStream)."""
    a = _arr(data)
    cap = deflate_bound(a.size) + 64
    out = np.empty(cap, dtype=np.uint8)
    olen = C.c_int64(0)
    _check(lib().szl_deflate_oneshot(a.ctypes.data, a.size, level, 1 if nowrap else 0, strategy, pattern, chunk, obuf,
                                     out.ctypes.data, cap, C.byref(olen)))
    return out[:olen.value].tobytes()


def deflate_trace(data, level=6, nowrap=True, strategy=0, pattern=0, chunk=0):
    a = _arr(data)
    cap = deflate_bound(a.size) + 64
    out = np.empty(cap, dtype=np.uint8)
    olen = C.c_int64(0)
    maxrows = a.size // 4 + 64
    rows = np.zeros((maxrows, 5), dtype=np.int32)
    nrows = C.c_int32(0)
    _check(lib().szl_deflate_trace(a.ctypes.data, a.size, level, 1 if nowrap else 0, strategy, pattern, chunk,
                                   out.ctypes.data, cap, C.byref(olen), rows.ctypes.data, maxrows, C.byref(nrows)))
    return out[:olen.value].tobytes(), rows[:nrows.value].copy()


def inflate(data, nowrap=True, max_out=None, ibuf=4096):
    """Returns (bytes, remaining_input, finished)."""
    a = _arr(data)
    cap = max_out if max_out is not None else max(1024, a.size * 1040 + 4096)
    out = np.empty(cap, dtype=np.uint8)
    olen = C.c_int64(0)
    rem = C.c_int32(0)
    fin = C.c_int32(0)
    _check(lib().szl_inflate_oneshot(a.ctypes.data, a.size, 1 if nowrap else 0, ibuf, out.ctypes.data, cap,
                                     C.byref(olen), C.byref(rem), C.byref(fin)))
    return out[:olen.value].tobytes(), rem.value, bool(fin.value)


def batch(direction, buffers, level=6, nowrap=True, threads=1, out_caps=None):
    """Deflate (0) or inflate (1) independent buffers on `threads` host threads.  Returns list of bytes."""
    n = len(buffers)
    lens = np.array([len(b) for b in buffers], dtype=np.int64)
    offs = np.zeros(n, dtype=np.int64)
    if n:
        offs[1:] = np.cumsum(lens)[:-1]
    blob = np.frombuffer(b"".join(bytes(b) for b in buffers), dtype=np.uint8) if n else np.zeros(0, np.uint8)
    if out_caps is None:
        out_caps = [deflate_bound(int(l)) + 64 for l in lens]
    caps = np.array(out_caps, dtype=np.int64)
    ooffs = np.zeros(n, dtype=np.int64)
    if n:
        ooffs[1:] = np.cumsum(caps)[:-1]
    out = np.empty(int(caps.sum()), dtype=np.uint8)
    olens = np.zeros(n, dtype=np.int64)
    _check(lib().szl_batch(direction, _ptr(blob), offs.ctypes.data, lens.ctypes.data, n, level, 1 if nowrap else 0,
                           out.ctypes.data, ooffs.ctypes.data, caps.ctypes.data, olens.ctypes.data, threads))
    return [out[ooffs[i]:ooffs[i] + olens[i]].tobytes() for i in range(n)]


class BatchJob:
    """Pre-packed batch for timing: run() is only the C call (szl_batch), so Python overhead stays out of the CPU baseline."""

    def __init__(self, direction, buffers, level=6, nowrap=True, out_caps=None):
        self.direction, self.level, self.nowrap = direction, level, nowrap
        n = self.n = len(buffers)
        self.lens = np.array([len(b) for b in buffers], dtype=np.int64)
        self.offs = np.zeros(n, dtype=np.int64)
        if n:
            self.offs[1:] = np.cumsum(self.lens)[:-1]
        self.blob = np.frombuffer(b"".join(bytes(b) for b in buffers), dtype=np.uint8).copy() if n else np.zeros(1, np.uint8)
        if out_caps is None:
            out_caps = [deflate_bound(int(l)) + 64 for l in self.lens]
        self.caps = np.array(out_caps, dtype=np.int64)
        self.ooffs = np.zeros(n, dtype=np.int64)
        if n:
            self.ooffs[1:] = np.cumsum(self.caps)[:-1]
        self.out = np.empty(int(self.caps.sum()) + 1, dtype=np.uint8)
        self.olens = np.zeros(n, dtype=np.int64)

    def run(self, threads=1):
        _check(lib().szl_batch(self.direction, self.blob.ctypes.data, self.offs.ctypes.data, self.lens.ctypes.data, self.n,
                               self.level, 1 if self.nowrap else 0, self.out.ctypes.data, self.ooffs.ctypes.data,
                               self.caps.ctypes.data, self.olens.ctypes.data, threads))

    def results(self):
        return [self.out[self.ooffs[i]:self.ooffs[i] + self.olens[i]].tobytes() for i in range(self.n)]


class Deflater:
    """Thin handle over the oracle's restated Deflater (Zip/Compression/Deflater.cs)."""

    def __init__(self, level=-1, nowrap=False):
        h = C.c_void_p()
        _check(lib().szl_deflater_new(level, 1 if nowrap else 0, C.byref(h)))
        self._h = h

    def __del__(self):
        if getattr(self, "_h", None):
            lib().szl_deflater_free(self._h)
            self._h = None

    def reset(self):
        _check(lib().szl_deflater_reset(self._h))

    def set_input(self, data):
        a = _arr(data)
        buf = a if a.size else np.zeros(1, np.uint8)
        _check(lib().szl_deflater_set_input(self._h, buf.ctypes.data, 0, a.size))

    def set_level(self, level):
        _check(lib().szl_deflater_set_level(self._h, level))

    def set_strategy(self, s):
        _check(lib().szl_deflater_set_strategy(self._h, s))

    def set_dictionary(self, data):
        a = _arr(data)
        _check(lib().szl_deflater_set_dictionary(self._h, a.ctypes.data, 0, a.size))

    def flush(self):
        lib().szl_deflater_flush(self._h)

    def finish(self):
        lib().szl_deflater_finish(self._h)

    def deflate(self, n):
        out = np.empty(max(n, 1), dtype=np.uint8)
        p = C.c_int32(0)
        _check(lib().szl_deflater_deflate(self._h, out.ctypes.data, 0, n, C.byref(p)))
        return out[:p.value].tobytes()

    @property
    def needs_input(self):
        return bool(lib().szl_deflater_needs_input(self._h))

    @property
    def finished(self):
        return bool(lib().szl_deflater_is_finished(self._h))

    @property
    def total_in(self):
        return lib().szl_deflater_total_in(self._h)

    @property
    def total_out(self):
        return lib().szl_deflater_total_out(self._h)

    @property
    def adler(self):
        return lib().szl_deflater_adler(self._h)


class Inflater:
    """Thin handle over the oracle's restated Inflater (Zip/Compression/Inflater.cs)."""

    def __init__(self, nowrap=False):
        h = C.c_void_p()
        _check(lib().szl_inflater_new(1 if nowrap else 0, C.byref(h)))
        self._h = h

    def __del__(self):
        if getattr(self, "_h", None):
            lib().szl_inflater_free(self._h)
            self._h = None

    def reset(self):
        _check(lib().szl_inflater_reset(self._h))

    def set_input(self, data):
        a = _arr(data)
        buf = a if a.size else np.zeros(2, np.uint8)
        _check(lib().szl_inflater_set_input(self._h, buf.ctypes.data, 0, a.size))

    def set_dictionary(self, data):
        a = _arr(data)
        _check(lib().szl_inflater_set_dictionary(self._h, a.ctypes.data, 0, a.size))

    def inflate(self, n):
        out = np.empty(max(n, 1), dtype=np.uint8)
        p = C.c_int32(0)
        _check(lib().szl_inflater_inflate(self._h, out.ctypes.data, max(n, 1), 0, n, C.byref(p)))
        return out[:p.value].tobytes()

    @property
    def needs_input(self):
        return bool(lib().szl_inflater_needs_input(self._h))

    @property
    def needs_dictionary(self):
        return bool(lib().szl_inflater_needs_dictionary(self._h))

    @property
    def finished(self):
        return bool(lib().szl_inflater_is_finished(self._h))

    @property
    def remaining_input(self):
        return lib().szl_inflater_remaining_input(self._h)

    @property
    def total_in(self):
        return lib().szl_inflater_total_in(self._h)

    @property
    def total_out(self):
        return lib().szl_inflater_total_out(self._h)

    @property
    def adler(self):
        return lib().szl_inflater_adler(self._h)
