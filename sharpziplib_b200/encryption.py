"""Mirror of the reference's entry ciphers over libb200z.so: same class and member names, argument
meaning and errors as Encryption/ZipAESTransform.cs and Encryption/PkzipClassic.cs, so that tests read like the reference's.
The arithmetic runs on the GPU (csrc/b200z_crypto.cu); there is no CPU implementation here."""
import ctypes as C

import numpy as np

from . import _lib

AUTH_CODE_LENGTH = 10  # Encryption/ZipAESStream.cs: the archive keeps the first 10 bytes of the HMAC


def _u8(data):
    return np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8))


class ZipAESTransform:
    """ICSharpCode.SharpZipLib.Encryption.ZipAESTransform (ZipAESTransform.cs:41): AES in CTR mode keyed by PBKDF2 from the
    password, HMAC-SHA1 over the ciphertext."""

    def __init__(self, key, saltBytes, blockSize, writeMode):
        self._h = C.c_void_p()
        if blockSize not in (16, 32):
            raise ValueError("Invalid blocksize %d. Must be 16 or 32." % blockSize)
        if len(saltBytes) != blockSize // 2:
            raise ValueError("Invalid salt len. Must be %d for blocksize %d" % (blockSize // 2, blockSize))
        pw = key.encode("utf-8") if isinstance(key, str) else bytes(key)
        p, s = _u8(pw), _u8(saltBytes)
        _lib.raise_for(_lib.lib().b200z_aes_transform_create(p.ctypes.data if p.size else None, p.size, s.ctypes.data, blockSize,
                                                             1 if writeMode else 0, C.byref(self._h)))
        self._blockSize = blockSize

    def TransformBlock(self, inputBuffer, inputOffset, inputCount, outputBuffer, outputOffset):
        src = _u8(inputBuffer)[inputOffset:inputOffset + inputCount]
        if src.size != inputCount:
            raise ValueError("inputCount")
        out = np.empty(inputCount, dtype=np.uint8)
        _lib.raise_for(_lib.lib().b200z_aes_transform_block(self._h, src.ctypes.data if inputCount else None, inputCount,
                                                            out.ctypes.data if inputCount else None))
        outputBuffer[outputOffset:outputOffset + inputCount] = out.tobytes()
        return inputCount

    @property
    def PwdVerifier(self):
        v = np.zeros(2, np.uint8)
        _lib.raise_for(_lib.lib().b200z_aes_transform_pwd_verifier(self._h, v.ctypes.data))
        return v.tobytes()

    def GetAuthCode(self):
        v = np.zeros(20, np.uint8)
        _lib.raise_for(_lib.lib().b200z_aes_transform_auth_code(self._h, v.ctypes.data))
        return v.tobytes()

    InputBlockSize = property(lambda self: self._blockSize)
    OutputBlockSize = property(lambda self: self._blockSize)
    CanTransformMultipleBlocks = True
    CanReuseTransform = True

    def Dispose(self):
        if self._h:
            _lib.lib().b200z_aes_transform_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = Dispose


def aes_derive_keys(passwords, salts, key_bytes):
    """ZipAESTransform's constructor for a batch of entries -> n x (key1 | key2 | verifier) as one uint8 array"""
    n = len(passwords)
    blob = b"".join(passwords)
    off = np.zeros(n + 1, np.int64)
    off[1:] = np.cumsum([len(p) for p in passwords])
    b = _u8(blob + b"\0")
    s = _u8(b"".join(salts))
    out = np.zeros(n * (2 * key_bytes + 2), np.uint8)
    _lib.raise_for(_lib.lib().b200z_aes_derive_keys(b.ctypes.data, off.ctypes.data, s.ctypes.data, key_bytes, n, out.ctypes.data))
    return out.reshape(n, 2 * key_bytes + 2)


def aes_batch(buffers, keys, key_bytes, write_mode):
    """TransformBlock over whole entries + GetAuthCode(): (outputs, n x 20 auth codes)"""
    n = len(buffers)
    ins = [_u8(b) for b in buffers]
    lens = np.array([a.size for a in ins], np.int64)
    outs = [np.empty(max(a.size, 1), np.uint8) for a in ins]
    auth = np.zeros((n, 20), np.uint8)
    k = np.ascontiguousarray(keys, dtype=np.uint8)
    pin = (C.c_void_p * n)(*[a.ctypes.data for a in ins])
    pout = (C.c_void_p * n)(*[a.ctypes.data for a in outs])
    _lib.raise_for(_lib.lib().b200z_aes_batch(pin, lens.ctypes.data, n, key_bytes, k.ctypes.data, 1 if write_mode else 0, pout,
                                              auth.ctypes.data))
    return [outs[i][:lens[i]].tobytes() for i in range(n)], auth


class PkzipClassic:
    """ICSharpCode.SharpZipLib.Encryption.PkzipClassic (PkzipClassic.cs:13)"""

    @staticmethod
    def GenerateKeys(seed):
        if seed is None:
            raise ValueError("seed")
        s = _u8(seed)
        if s.size == 0:
            raise ValueError("Length is zero")
        out = np.zeros(12, np.uint8)
        _lib.raise_for(_lib.lib().b200z_pkzip_generate_keys(s.ctypes.data, s.size, out.ctypes.data))
        return out.tobytes()


class _PkzipTransform:
    _encrypt = 0

    def __init__(self, keyBlock):
        if keyBlock is None:
            raise ValueError("keyData")
        if len(keyBlock) != 12:
            raise _lib.InvalidOperationException("Key length is not valid")
        self._keys = np.array(np.frombuffer(bytes(keyBlock), dtype=np.uint8))

    def TransformBlock(self, inputBuffer, inputOffset, inputCount, outputBuffer, outputOffset):
        src = _u8(inputBuffer)[inputOffset:inputOffset + inputCount]
        out = np.empty(max(inputCount, 1), dtype=np.uint8)
        lens = np.array([inputCount], np.int64)
        pin, pout = (C.c_void_p * 1)(src.ctypes.data if inputCount else None), (C.c_void_p * 1)(out.ctypes.data)
        _lib.raise_for(_lib.lib().b200z_pkzip_batch(pin, lens.ctypes.data, 1, self._keys.ctypes.data, self._encrypt, pout))
        outputBuffer[outputOffset:outputOffset + inputCount] = out[:inputCount].tobytes()
        return inputCount

    def TransformFinalBlock(self, inputBuffer, inputOffset, inputCount):
        out = bytearray(inputCount)
        self.TransformBlock(inputBuffer, inputOffset, inputCount, out, 0)
        return bytes(out)

    InputBlockSize = OutputBlockSize = 1
    CanTransformMultipleBlocks = True
    CanReuseTransform = True


class PkzipClassicEncryptCryptoTransform(_PkzipTransform):
    """PkzipClassic.cs:129"""
    _encrypt = 1


class PkzipClassicDecryptCryptoTransform(_PkzipTransform):
    """PkzipClassic.cs:239"""
    _encrypt = 0


def pkzip_batch(buffers, keys12, encrypt):
    """TransformBlock over n whole streams: (outputs, keys after); keys12: n x 12 bytes"""
    n = len(buffers)
    ins = [_u8(b) for b in buffers]
    lens = np.array([a.size for a in ins], np.int64)
    outs = [np.empty(max(a.size, 1), np.uint8) for a in ins]
    k = np.array(np.ascontiguousarray(keys12, dtype=np.uint8)).reshape(n, 12)
    pin = (C.c_void_p * n)(*[a.ctypes.data for a in ins])
    pout = (C.c_void_p * n)(*[a.ctypes.data for a in outs])
    _lib.raise_for(_lib.lib().b200z_pkzip_batch(pin, lens.ctypes.data, n, k.ctypes.data, 1 if encrypt else 0, pout))
    return [outs[i][:lens[i]].tobytes() for i in range(n)], k
