"""CPU oracle for the density hot path — TEST INFRASTRUCTURE ONLY.

ctypes wrapper over oracle/libdensity_oracle.so (built from density_oracle.c, a C
restatement of the reference: /root/reference/src/codec/codec.rs, protection_state.rs,
algorithms/{chameleon,cheetah,lion}/*.rs). Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference leg may import this module; density_b200/
never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libdensity_oracle.so")
ALGS = {"chameleon": 0, "cheetah": 1, "lion": 2}
BLOCK = {"chameleon": 256, "cheetah": 128, "lion": 64}
SIG = {"chameleon": 8, "cheetah": 8, "lion": 6}


def build(force=False):
    src = os.path.join(_HERE, "density_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "clean", "all"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        for name in ("oracle_encode", "oracle_decode"):
            f = getattr(L, name)
            f.restype = ctypes.c_size_t
            f.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
        L.oracle_encode_stats.restype = ctypes.c_size_t
        L.oracle_encode_stats.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                                          ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint64)]
        L.oracle_codec_new.restype = ctypes.c_void_p
        L.oracle_codec_new.argtypes = [ctypes.c_int]
        L.oracle_codec_free.restype = None
        L.oracle_codec_free.argtypes = [ctypes.c_void_p]
        L.oracle_codec_clear_state.restype = None
        L.oracle_codec_clear_state.argtypes = [ctypes.c_void_p]
        for name in ("oracle_codec_encode", "oracle_codec_decode"):
            f = getattr(L, name)
            f.restype = ctypes.c_size_t
            f.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
        _lib = L
    return _lib


class Codec:
    """A reference Codec INSTANCE that is reused across calls (codec.rs:16,72,82): the dictionary survives until clear_state()."""

    def __init__(self, alg):
        self.alg = alg
        self._h = lib().oracle_codec_new(ALGS[alg])

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_codec_free(self._h)
            self._h = None

    def clear_state(self):
        lib().oracle_codec_clear_state(self._h)

    def encode(self, data):
        a = _as_u8(data)
        cap = safe_encode_buffer_size(self.alg, a.size)
        out = np.empty(max(cap, 1), dtype=np.uint8)
        n = lib().oracle_codec_encode(self._h, a.ctypes.data, a.size, out.ctypes.data, cap)
        return out[:n].copy()

    def decode(self, data, out_size):
        a = _as_u8(data)
        out = np.empty(max(out_size, 1), dtype=np.uint8)
        n = lib().oracle_codec_decode(self._h, a.ctypes.data, a.size, out.ctypes.data, out_size)
        return out[:n].copy()


def safe_encode_buffer_size(alg, size):
    """codec.rs:18-21"""
    b, s = BLOCK[alg], SIG[alg]
    return size + (size // b) * s + (s if size % b else 0)


def _as_u8(data):
    if isinstance(data, np.ndarray):
        a = np.ascontiguousarray(data).view(np.uint8).reshape(-1)
    else:
        a = np.frombuffer(bytes(data), dtype=np.uint8)
    return a


def encode(alg, data, cap=None, return_copied=False):
    a = _as_u8(data)
    cap = safe_encode_buffer_size(alg, a.size) if cap is None else cap
    out = np.empty(max(cap, 1), dtype=np.uint8)
    copied = ctypes.c_uint64(0)
    n = lib().oracle_encode_stats(ALGS[alg], a.ctypes.data, a.size, out.ctypes.data, cap, ctypes.byref(copied))
    res = out[:n].copy()
    return (res, copied.value) if return_copied else res


def decode(alg, data, out_size):
    a = _as_u8(data)
    out = np.empty(max(out_size, 1), dtype=np.uint8)
    n = lib().oracle_decode(ALGS[alg], a.ctypes.data, a.size, out.ctypes.data, out_size)
    return out[:n].copy()
