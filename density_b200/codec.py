"""Host-side mirror of the reference's codec interface for the accelerated path.

Mirrors `trait Codec` (/root/reference/src/codec/codec.rs:12-127) and the inherent associated functions of
Chameleon / Cheetah / Lion (/root/reference/src/algorithms/chameleon/chameleon.rs:39-53, cheetah.rs:47-65,
lion.rs:64-82): same names, same argument meaning (`encode(input, output) -> bytes written`), same sizing
contract (`safe_encode_buffer_size`). All work happens in libdensity_b200.so (CUDA, sm_100a).

Buffers may be: bytes / bytearray / memoryview / numpy uint8 arrays (host) or torch CUDA uint8 tensors
(device-resident, no staging copies).
"""
import ctypes

import numpy as np

from . import _lib

ALG_IDS = {"chameleon": 0, "cheetah": 1, "lion": 2}


class EncodeError(Exception):
    """/root/reference/src/errors/encode_error.rs:4-13"""


class DecodeError(Exception):
    """/root/reference/src/errors/decode_error.rs:4-13"""


def _ptr_len(buf, writable=False):
    """-> (address, nbytes, keepalive)"""
    try:
        import torch
        if isinstance(buf, torch.Tensor):
            if buf.dtype != torch.uint8 or not buf.is_contiguous():
                raise TypeError("torch buffers must be contiguous uint8")
            return buf.data_ptr(), buf.numel(), buf
    except ImportError:  # pragma: no cover
        pass
    if isinstance(buf, np.ndarray):
        if buf.dtype != np.uint8 or not buf.flags.c_contiguous:
            raise TypeError("numpy buffers must be contiguous uint8")
        if writable and not buf.flags.writeable:
            raise TypeError("output buffer is read-only")
        return buf.ctypes.data, buf.size, buf
    if isinstance(buf, (bytes, bytearray, memoryview)):
        if writable:
            if isinstance(buf, bytes):
                raise TypeError("output buffer must be writable (bytearray / numpy / torch)")
            a = np.frombuffer(buf, dtype=np.uint8)
        else:
            a = np.frombuffer(buf, dtype=np.uint8)
        return a.ctypes.data, a.size, a
    raise TypeError(f"unsupported buffer type {type(buf)!r}")


class _Codec:
    """One algorithm. The reference's instances own a dictionary (`state`); every public entry point used by its
    benches and FFI builds a fresh one per call (chameleon.rs:45-53), which is what this path accelerates."""
    NAME = None
    _BLOCK = None
    _UNIT = None
    _SIG = None

    @classmethod
    def block_size(cls):
        return cls._BLOCK

    @classmethod
    def decode_unit_size(cls):
        return cls._UNIT

    @classmethod
    def signature_significant_bytes(cls):
        return cls._SIG

    @classmethod
    def safe_encode_buffer_size(cls, size):
        """codec.rs:18-21 (computed by the library)."""
        return getattr(_lib.load(), f"{cls.NAME}_safe_encode_buffer_size")(size)

    @classmethod
    def encode(cls, input, output):
        """Encode `input` into `output`; returns the number of bytes written (codec.rs:72-80)."""
        ip, n, k1 = _ptr_len(input)
        op, cap, k2 = _ptr_len(output, writable=True)
        r = getattr(_lib.load(), f"{cls.NAME}_encode")(ip, n, op, cap)
        if r == 0 and n != 0:
            raise EncodeError(_lib.last_error() or "encode failed")
        return r

    @classmethod
    def decode(cls, input, output):
        """Decode `input` into `output`; returns the number of bytes written (codec.rs:82-126)."""
        ip, n, k1 = _ptr_len(input)
        op, cap, k2 = _ptr_len(output, writable=True)
        r = getattr(_lib.load(), f"{cls.NAME}_decode")(ip, n, op, cap)
        if r == 0 and n != 0:
            raise DecodeError(_lib.last_error() or "decode failed")
        return r

    # convenience: bytes in, bytes out
    @classmethod
    def encode_bytes(cls, data):
        a = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else data
        out = np.empty(max(1, cls.safe_encode_buffer_size(a.size)), dtype=np.uint8)
        n = cls.encode(a, out)
        return out[:n].tobytes()

    @classmethod
    def decode_bytes(cls, data, original_size):
        a = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else data
        out = np.empty(max(1, original_size), dtype=np.uint8)
        n = cls.decode(a, out)
        return out[:n].tobytes()


class Chameleon(_Codec):
    """chameleon.rs:138-147"""
    NAME, _BLOCK, _UNIT, _SIG = "chameleon", 256, 8, 8


class Cheetah(_Codec):
    """cheetah.rs:188-197"""
    NAME, _BLOCK, _UNIT, _SIG = "cheetah", 128, 4, 8


class Lion(_Codec):
    """lion.rs:317-326"""
    NAME, _BLOCK, _UNIT, _SIG = "lion", 64, 4, 6


CODECS = {"chameleon": Chameleon, "cheetah": Cheetah, "lion": Lion}


class CodecInstance:
    """A Codec instance that is reused across calls (`let mut c = Chameleon::new(); c.encode(a, ..); c.encode(b, ..)`,
    codec.rs:16,72,82): the dictionary survives until clear_state(); the protection state is fresh in every call."""

    def __init__(self, alg):
        self.alg = alg
        self._lib = _lib.load()
        self._h = self._lib.density_b200_codec_create(ALG_IDS[alg])
        if not self._h:
            raise _lib.DensityB200Error(_lib.last_error())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.density_b200_codec_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def clear_state(self):
        if self._lib.density_b200_codec_clear_state(self._h):
            raise _lib.DensityB200Error(_lib.last_error())

    def encode(self, input, output):
        ip, n, k1 = _ptr_len(input)
        op, cap, k2 = _ptr_len(output, writable=True)
        r = self._lib.density_b200_codec_encode(self._h, ip, n, op, cap)
        if r == 0 and n != 0:
            raise EncodeError(_lib.last_error() or "encode failed")
        return r

    def decode(self, input, output):
        ip, n, k1 = _ptr_len(input)
        op, cap, k2 = _ptr_len(output, writable=True)
        r = self._lib.density_b200_codec_decode(self._h, ip, n, op, cap)
        if r == 0 and n != 0:
            raise DecodeError(_lib.last_error() or "decode failed")
        return r


# ---- stream-ordered device API (torch tensors) ------------------------------------------------------------------
def _stream_handle(stream):
    import torch
    s = torch.cuda.current_stream() if stream is None else stream
    return ctypes.c_void_p(s.cuda_stream)


def encode_device(alg, d_in, d_out, d_out_size, stream=None, path=0):
    """Enqueue an encode of CUDA uint8 tensor `d_in` into `d_out` on `stream` (default: torch's current stream).
    `d_out_size` is a CUDA int64/uint64 tensor with one element that receives the encoded size. No synchronisation."""
    L = _lib.load()
    rc = L.density_b200_encode_device_path(ALG_IDS[alg], d_in.data_ptr(), d_in.numel(), d_out.data_ptr(), d_out.numel(),
                                           d_out_size.data_ptr(), _stream_handle(stream), path)
    if rc != 0:
        raise EncodeError(f"density_b200_encode_device rc={rc}: {_lib.last_error()}")


def decode_device(alg, d_in, n_in, d_out, d_out_size, stream=None, path=0):
    L = _lib.load()
    rc = L.density_b200_decode_device_path(ALG_IDS[alg], d_in.data_ptr(), n_in, d_out.data_ptr(), d_out.numel(),
                                           d_out_size.data_ptr(), _stream_handle(stream), path)
    if rc != 0:
        raise DecodeError(f"density_b200_decode_device rc={rc}: {_lib.last_error()}")
