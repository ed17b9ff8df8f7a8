"""CPU checks of the parallel FORMULATIONS the CUDA kernels implement (no GPU): the Python models under tools/ reproduce the oracle's
stream byte for byte, and the closed-form automaton jump of the decode boundary walk equals the step-by-step automaton.

These are the arguments DESIGN.md §3/§4/§4b rest on; the kernels themselves are checked on the GPU (tests/test_gpu_parity.py)."""
import os
import sys

import numpy as np
import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools import proto_cheetah_runs, proto_lion_runs  # noqa: E402


def _cases():
    d = np.fromfile(os.path.join(ROOT, "tests", "golden", "dickens_200k.bin"), np.uint8)
    rng = np.random.default_rng(3)
    return {
        "kat": np.frombuffer(b"test" * 31 + b"t", np.uint8),
        "dickens": d[:12003],
        "mixed": np.concatenate([d[:6000], rng.integers(0, 256, 3000, dtype=np.uint8), np.zeros(2000, np.uint8), d[30000:36002]]),
        "low": rng.integers(0, 3, 8000, dtype=np.uint8),
    }


@pytest.mark.parametrize("name", ["kat", "dickens", "mixed", "low"])
@pytest.mark.parametrize("nruns", [1, 5])
def test_cheetah_run_decomposition_reproduces_the_oracle(name, nruns):
    """cheetah.rs:121-150 = previous-in-context equality + MRU-2 chunk map on the non-predicted quads + fixed-point copy map,
    with at most 1 + 2 undecided accesses per run and key (cheetah_encode.cu)."""
    data = _cases()[name]
    got, _, _ = proto_cheetah_runs.encode(data, nruns)
    want = oracle.encode("cheetah", data)
    assert got.size == want.size and (got == want).all()


@pytest.mark.parametrize("name", ["kat", "dickens", "mixed", "low"])
@pytest.mark.parametrize("nruns", [1, 5])
def test_lion_run_decomposition_reproduces_the_oracle(name, nruns):
    """lion.rs:209-271: 5-deep move-to-front lists with at most 5 undecided accesses per run and context (cheetah_encode.cu)."""
    data = _cases()[name]
    got, _, _ = proto_lion_runs.encode(data, nruns)
    want = oracle.encode("lion", data)
    assert got.size == want.size and (got == want).all()


class _Protection:
    """protection_state.rs:9-47"""
    def __init__(self, penalty=0, start=1, prev=False, counter=0):
        self.penalty, self.start, self.prev, self.counter = penalty, start, prev, counter

    def revert_to_copy(self):
        if (self.counter & 0xF) == 0 and self.start > 1:
            self.start >>= 1
        self.counter += 1
        return self.penalty > 0

    def update(self, inc):
        if inc:
            if self.prev:
                self.penalty = self.start
            self.prev = True
        else:
            self.prev = False


def _sw_jump(ps, nb, last_inc):
    """model of sw_jump() in density_b200/csrc/chameleon_decode.cu"""
    k = (ps.counter + nb + 15) // 16 - (ps.counter + 15) // 16
    if ps.start > 1:
        ps.start = max(1, ps.start >> min(k, 8))
    ps.counter += nb
    ps.prev = last_inc


def test_decode_walk_jump_equals_the_stepwise_automaton():
    """A chunk/group may be jumped when penalty == 0 on entry and no two consecutive blocks (including the seam) are incompressible:
    then the automaton state after nb blocks is (penalty 0, start halved once per 16th block, prev = last block's bit)."""
    rng = np.random.default_rng(11)
    for _ in range(3000):
        start = int(rng.integers(1, 40)); counter = int(rng.integers(0, 1 << 20)); prev = bool(rng.integers(0, 2))
        nb = int(rng.integers(1, 9000))
        inc = rng.random(nb) < 0.3
        inc[1:] &= ~inc[:-1]                       # no two consecutive incompressible blocks inside
        if prev:
            inc[0] = False                         # nor across the entry seam
        a = _Protection(0, start, prev, counter)
        for b in range(nb):
            assert not a.revert_to_copy()
            a.update(bool(inc[b]))
        j = _Protection(0, start, prev, counter)
        _sw_jump(j, nb, bool(inc[-1]))
        assert (a.penalty, a.start, a.prev, a.counter) == (0, j.start, j.prev, j.counter)


@pytest.mark.parametrize("seed", range(6))
def test_decode_boundary_walk_model_equals_in_order_parse(seed):
    """dec_chunk_walk / dec_group_compose / dec_seq_walk (chunk and group jumps) / dec_chunk_entries / dec_block_offsets, modelled in
    tools/proto_decode_walk.py with small chunks, against the plain in-order parse of oracle streams with copy-mode episodes."""
    from tools import proto_decode_walk as W
    rng = np.random.default_rng(100 + seed)
    d = np.fromfile(os.path.join(ROOT, "tests", "golden", "dickens_200k.bin"), np.uint8)
    parts = []
    pos = int(rng.integers(0, 50000))
    for _ in range(int(rng.integers(2, 7))):
        ln = int(rng.integers(2000, 40000)); parts.append(d[pos:pos + ln]); pos = (pos + ln) % 150000
        kind = int(rng.integers(0, 4))
        if kind == 0:
            parts.append(rng.integers(0, 256, int(rng.integers(300, 20000)), dtype=np.uint8))       # incompressible burst
        elif kind == 1:
            parts.append(np.arange(int(rng.integers(100, 3000)), dtype=np.uint32).view(np.uint8))    # all-plain counters
        elif kind == 2:
            parts.append(np.zeros(int(rng.integers(100, 5000)), np.uint8))
    data = np.concatenate(parts)
    enc, copied = oracle.encode("chameleon", data, return_copied=True)
    want = W.reference_parse(enc)
    assert sum(1 for o in want[0] if o & W.COPY) > 0 or copied == 0
    got = W.model_parse(enc, CH=2048, GROUP=4)
    assert got == want


@pytest.mark.parametrize("name", ["kat", "dickens", "low"])
def test_planned_cheetah_decoder_model_round_trips(name):
    """tools/proto_cheetah_decode_full.py (the decomposition the next round's Cheetah decoder is built on: chunk-map values run-parallel,
    predicted values by iteration on the contexts, copy-mode blocks and tail in order) reproduces the input from the oracle's stream."""
    from tools import proto_cheetah_decode_full as D
    data = _cases()[name]
    enc = oracle.encode("cheetah", data)
    got, rounds, ncopy, produced = D.decode(enc, data.size, nruns=3)
    assert produced == data.size and (got == data).all()


@pytest.mark.parametrize("name", ["kat", "dickens", "mixed"])
def test_planned_lion_decoder_model_round_trips(name):
    """tools/proto_lion_decode_full.py: 6-byte signatures, 3-bit flags, 5 hashes per context in the serial hash chain, values per context."""
    from tools import proto_lion_decode_full as D
    data = _cases()[name]
    enc = oracle.encode("lion", data)
    got, ncopy, produced = D.decode(enc, data.size, nruns=3)
    assert produced == data.size and (got == data).all()
