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


def _flag_cases():
    from tools import proto_tile_protocol_v6 as m6
    d = np.fromfile(os.path.join(ROOT, "tests", "golden", "dickens_200k.bin"), np.uint8)
    rng = np.random.default_rng(9)
    text = d[:160000].view(np.uint32).copy()
    # two quads of one bucket
    seen, q = {}, 0x12345678
    while True:
        q = (q * 1103515245 + 12345) & 0xFFFFFFFF
        hh = ((q * m6.M) & 0xFFFFFFFF) >> 16
        if hh in seen and seen[hh] != q:
            q1, q2 = seen[hh], q
            break
        seen[hh] = q
    runs = []
    pos = 0
    for k, L in enumerate([1, 2, 3, 31, 32, 33, 255, 4096, 4097, 700, 5, 64]):
        runs.append(text[pos:pos + 1500]); pos += 1500
        v = [0, 0xFFFFFFFF, 0x20202020, q1, q2][k % 5]
        r = np.full(L, v, np.uint32)
        if k % 2 and L > 2:
            r[L // 2] = q2 if v != q2 else q1           # a run cut by another quad (of the same bucket for q1 / q2)
        runs.append(r)
    burst = np.empty(80, np.uint32); burst[0::2] = q1; burst[1::2] = text[:40]; burst[0::4] = q2      # 40 members of one bucket, not a run
    return {
        "dickens": text,
        "noise": rng.integers(0, 2 ** 32, 20000, dtype=np.uint32),
        "low": rng.integers(0, 4, 30000, dtype=np.uint32) * 0x01010101,
        "runs": np.concatenate(runs + [text[pos:pos + 3000]]),
        "pileup": np.concatenate([text[:5000], burst, text[5000:9000], np.tile(np.array([q1, q2], np.uint32), 300), text[9000:12000]]),
        "zeros": np.zeros(9000, np.uint32),
    }


@pytest.mark.parametrize("name", ["dickens", "noise", "low", "runs", "pileup", "zeros"])
def test_write_verify_mailbox_tile_protocol_equals_the_in_order_walk(name):
    """cham_flag_pass6: racy publish, verify by re-reading, dirty members resolved through per-bucket mailboxes by
    record index, runs of equal quads dropped, overflow mailboxes, in-order replay of a tile whose mailboxes overflow — for ANY winner
    of the racy publishes the flags, the first touches and the final dictionary are those of chameleon.rs:86-101."""
    from tools import proto_tile_protocol_v6 as m6
    q = _flag_cases()[name]
    want, want_tab = m6.reference_flags(q)
    for seed in (1, 2, 3):
        stats = {}
        got, tab, touched = m6.flag_pass(q, seed=seed, stats=stats)
        assert (got == want).all(), (name, seed, int((got != want).sum()))
        assert {int(b): int(tab[b]) for b in np.flatnonzero(touched)} == want_tab
        if name == "pileup":
            assert stats["overflow"] >= 1          # the replay path is part of what is checked


def _candidate_copy_map(inc, PSEG=256, GROUP=8, NS=10, NP=10):
    """Model of prot_iterate's candidate-state evaluation (chameleon_encode.cu, PC_NC): transfer table per segment over the candidate
    incoming states, composed per group and in order at the top, true states handed back down, final walk. Returns the copy map, or
    None when the true path leaves the candidate set (the kernel then falls back to the relaxation)."""
    nb = inc.size
    nseg = (nb + PSEG - 1) // PSEG
    NC, ESC = 2 * NS * NP, -1

    def dec(c, counter):
        return _Protection(c % NP, (c // NP) % NS + 1, bool(c // (NP * NS)), counter)

    def enc(ps):
        if ps.penalty >= NP or ps.start < 1 or ps.start > NS:
            return ESC
        return (int(ps.prev) * NS + (ps.start - 1)) * NP + ps.penalty

    def walk(ps, s, cm=None):
        for b in range(s * PSEG, min((s + 1) * PSEG, nb)):
            if ps.revert_to_copy():
                if cm is not None:
                    cm[b] = 1
                ps.penalty = (ps.penalty - 1) & 0xFF          # decay(), protection_state.rs:29-35
                if ps.penalty == 0:
                    ps.start = (ps.start + 1) & 0xFF
            else:
                if cm is not None:
                    cm[b] = 0
                ps.update(bool(inc[b]))
        return ps

    T = [[enc(walk(dec(c, s * PSEG), s)) for c in range(NC)] for s in range(nseg)]
    ngrp = (nseg + GROUP - 1) // GROUP
    GT = []
    for g in range(ngrp):
        row = []
        for c in range(NC):
            x = c
            for s in range(g * GROUP, min((g + 1) * GROUP, nseg)):
                if x != ESC:
                    x = T[s][x]
            row.append(x)
        GT.append(row)
    gin, x = [], 0
    for g in range(ngrp):
        gin.append(x)
        if x != ESC:
            x = GT[g][x]
    if x == ESC:
        return None
    cm = np.zeros(nb, np.uint8)
    for g in range(ngrp):
        x = gin[g]
        for s in range(g * GROUP, min((g + 1) * GROUP, nseg)):
            walk(dec(x, s * PSEG), s, cm)
            x = T[s][x]
    return cm


@pytest.mark.parametrize("kind", ["noise", "bursts", "alternating", "quiet"])
def test_candidate_state_evaluation_of_the_protection_automaton(kind):
    """prot_iterate (Protection): walking every segment from every candidate incoming state and composing the transfer
    tables gives the copy map of the in-order automaton (codec.rs:35-37,68; protection_state.rs:18-47) on any incompressible-bit
    sequence whose seam states stay inside the candidate set — noise (every block incompressible) included."""
    rng = np.random.default_rng(21)
    nb = 20000
    if kind == "noise":
        inc = np.ones(nb, bool)
    elif kind == "bursts":
        inc = np.zeros(nb, bool)
        for _ in range(40):
            a = int(rng.integers(0, nb - 600)); inc[a:a + int(rng.integers(2, 600))] = True
    elif kind == "alternating":
        inc = rng.random(nb) < 0.6
    else:
        inc = rng.random(nb) < 0.02
        inc[1:] &= ~inc[:-1]
    want = np.zeros(nb, np.uint8)
    ps = _Protection()
    for b in range(nb):
        if ps.revert_to_copy():
            want[b] = 1
            ps.penalty = (ps.penalty - 1) & 0xFF
            if ps.penalty == 0:
                ps.start = (ps.start + 1) & 0xFF
        else:
            ps.update(bool(inc[b]))
    got = _candidate_copy_map(inc, PSEG=64, GROUP=8)
    assert got is not None, "the true path left the candidate set"
    assert (got == want).all()
    if kind == "quiet":
        assert want.sum() == 0


@pytest.mark.parametrize("name", ["dickens", "noise", "low", "runs", "pileup", "zeros"])
def test_decode_mark_map_mailbox_protocol_equals_the_in_order_decoder(name):
    """cham_decode_pass7: hashed mark map (false positives only make suspects), writers-only mailboxes, the writer
    without a successor leaves the bucket's value, in-order replay of an overflowing tile: the decoded quads and the final dictionary
    are those of chameleon.rs:55-68 on the flag / payload sequence the in-order ENCODER produces for the same quads."""
    from tools import proto_tile_protocol_v6 as m6
    q = _flag_cases()[name]
    flags, _ = m6.reference_flags(q)
    # the encoder's view with a zero-initialised dictionary: first touches are hits only for quad 0 in bucket 0 (chameleon.rs:41,89-91)
    h, _f = m6.hf(q)
    hit = (flags == 1) | ((flags == 2) & (q == 0))
    is_plain = ~hit
    payload = np.where(is_plain, q.astype(np.uint64), h.astype(np.uint64))
    want, want_dic = m6.decode_reference(is_plain, payload)
    assert (want == q).all()                                        # the in-order decoder inverts the in-order encoder
    stats = {}
    got, dic = m6.decode_pass(is_plain, payload, stats=stats)
    assert (got == want).all(), (name, int((got != want).sum()))
    assert dic == want_dic
    if name == "pileup":
        assert stats["overflow"] >= 1
