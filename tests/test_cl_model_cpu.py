"""CPU check of the run-parallel Cheetah / Lion decoder SCHEME (symbolic chunk-map pass + fold, iterated prediction rounds with
snapshots / slot tags / unknown propagation, context sweep, in-order tail): tests/cl_model.cpp executes it with the table logic the
CUDA kernels share (density_b200/csrc/cl_core.cuh), one run after the other, and must reproduce the input of every oracle-encoded
stream for any run count. The kernels themselves are checked on the GPU (tests/test_gpu_parity.py)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import oracle
from conftest import payload

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_cl_model.so")


@pytest.fixture(scope="module")
def model():
    src = os.path.join(HERE, "cl_model.cpp")
    core = os.path.join(HERE, "..", "density_b200", "csrc", "cl_core.cuh")
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(src), os.path.getmtime(core)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", src, "-o", SO])
    L = ctypes.CDLL(SO)
    L.cl_model_decode.restype = ctypes.c_size_t
    L.cl_model_decode.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint32,
                                  ctypes.c_uint32, ctypes.POINTER(ctypes.c_uint32)]
    return L


def run_model(L, alg, data, nruns, max_rounds=64):
    enc = oracle.encode(alg, data)
    out = np.zeros(data.size + 64, np.uint8)
    stats = (ctypes.c_uint32 * 4)()
    n = L.cl_model_decode(oracle.ALGS[alg], enc.ctypes.data, enc.size, out.ctypes.data, data.size, nruns, max_rounds, stats)
    return n, out[:n], stats[0], stats[1]


@pytest.mark.parametrize("alg", ["cheetah", "lion"])
@pytest.mark.parametrize("kind,nbytes", [("text", 200000), ("mixed", 150001), ("random", 40003), ("zeros", 60000), ("low", 70002),
                                         ("text", 127), ("text", 129), ("text", 4096 + 3), ("zeros", 5)])
@pytest.mark.parametrize("nruns", [1, 7, 64])
def test_scheme_reproduces_the_input(model, alg, kind, nbytes, nruns):
    data = payload(kind, nbytes, seed=11)
    n, got, rounds, conv = run_model(model, alg, data, nruns)
    assert conv == 1, (alg, kind, nruns, rounds)
    assert n == data.size and (got == data).all(), (alg, kind, nruns)


def test_cheetah_rounds_stay_small_on_text(model, dickens200k):
    """Cheetah: the prediction table is overwritten by 92 % of the quads (their hashes are in the stream), so a wrong context dies out
    within a few quads and the rounds do not grow with the run count."""
    for nruns in (4, 16, 64, 256):
        n, got, rounds, conv = run_model(model, "cheetah", dickens200k, nruns)
        assert conv == 1 and n == dickens200k.size and (got == dickens200k).all()
        assert rounds <= 12, (nruns, rounds)


def test_lion_rounds_grow_with_the_run_count(model):
    """Lion (documented negative result): a misplaced operation desynchronises a whole 5-deep move-to-front list, wrong reads
    keep producing wrong contexts, and exactness advances one run per round — still exact, but no better than in order. This is why
    lion_decode stays on the in-order kernel."""
    from density_b200 import synth
    data = synth.synth_text(2 << 20).numpy()
    n, got, rounds, conv = run_model(model, "lion", data, 16, max_rounds=64)
    assert conv == 1 and n == data.size and (got == data).all()
    assert rounds >= 8, rounds
