"""The oracle against the reference's own golden facts (CPU only).

Pins: known-answer vectors /root/reference/src/lib.rs:19,28,50,72; round trip benches/density.rs:42-45;
dickens ratios benchmark.log:17,22,27; digests SURVEY.md §8c (independent restatement)."""
import numpy as np
import pytest

import oracle
from conftest import ALGS, payload, sha256

TEST_DATA = b"test" * 31 + b"t"  # lib.rs:19
KAT = {
    # lib.rs:28
    "chameleon": [0xfe, 0xff, 0xff, 0x7f, 0, 0, 0, 0, 116, 101, 115, 116] + [112, 251] * 30 + [116],
    # lib.rs:50
    "cheetah": [244, 255, 255, 255, 255, 255, 255, 63, 116, 101, 115, 116, 112, 251, 116],
    # lib.rs:72
    "lion": [112, 146, 36, 73, 146, 36, 116, 101, 115, 116, 112, 251, 73, 146, 36, 73, 146, 4, 116],
}


@pytest.mark.parametrize("alg", ALGS)
def test_reference_known_answer_vectors(alg):
    # the reference encodes into a buffer of TEST_DATA.len() bytes (lib.rs:24,46,68), smaller than the safe size
    enc = oracle.encode(alg, TEST_DATA, cap=len(TEST_DATA))
    assert enc.tolist() == KAT[alg]
    assert bytes(oracle.decode(alg, enc, len(TEST_DATA))) == TEST_DATA


@pytest.mark.parametrize("alg", ALGS)
def test_golden_fixtures(alg, golden, golden_inputs):
    for name, data in golden_inputs.items():
        g = golden[name]
        assert g["input_len"] == data.size and g["input_sha256"] == sha256(data), name
        enc, copied = oracle.encode(alg, data, return_copied=True)
        e = g["alg"][alg]
        assert (enc.size, sha256(enc), copied) == (e["size"], e["sha256"], e["copied_blocks"]), name
        dec = oracle.decode(alg, enc, data.size)
        assert dec.size == data.size and (dec == data).all(), name


def test_published_dickens_ratios(golden):
    # benchmark.log:13,17,22,27 — 10,192,446 bytes; (1.749x) (1.860x) (1.966x)
    n = golden["dickens_full"]["input_len"]
    assert n == 10192446
    for alg, ratio in (("chameleon", "1.749"), ("cheetah", "1.860"), ("lion", "1.966")):
        assert f"{n / golden['dickens_full']['alg'][alg]['size']:.3f}" == ratio


@pytest.mark.parametrize("alg", ALGS)
def test_safe_encode_buffer_size(alg):
    # codec.rs:18-21; SURVEY §8(a15): 1 GiB -> +32 / +64 / +96 MiB
    add = {"chameleon": 32, "cheetah": 64, "lion": 96}[alg] << 20
    assert oracle.safe_encode_buffer_size(alg, 1 << 30) == (1 << 30) + add
    assert oracle.safe_encode_buffer_size(alg, 0) == 0
    assert oracle.safe_encode_buffer_size(alg, 1) == 1 + oracle.SIG[alg]


@pytest.mark.parametrize("alg", ALGS)
@pytest.mark.parametrize("kind", ["text", "random", "zeros", "low", "mixed"])
def test_round_trip_tail_sweep(alg, kind):
    lengths = [0, 1, 2, 3, 4, 5, 7, 8, 15, 16, 17, 63, 64, 65, 127, 128, 129, 255, 256, 257, 260, 511, 512, 513, 1000, 2999, 70001]
    for n in lengths:
        data = payload(kind, n, seed=n)
        enc = oracle.encode(alg, data)
        assert enc.size <= oracle.safe_encode_buffer_size(alg, n)
        dec = oracle.decode(alg, enc, n)
        assert dec.size == n and (dec == data).all(), (alg, kind, n)


def test_all_zero_is_a_hit_on_first_sight():
    # hash(0)=0 and tables start at 0 => quad 0 maps immediately (chameleon.rs:41,89-91)
    enc = oracle.encode("chameleon", np.zeros(256, np.uint8))
    assert enc.size == 8 + 2 * 64 and enc[:8].tolist() == [0xff] * 8


def test_copy_mode_engages_on_random():
    from conftest import splitmix_bytes
    data = splitmix_bytes(1 << 16, 3)
    for alg in ALGS:
        enc, copied = oracle.encode(alg, data, return_copied=True)
        nblocks = data.size // oracle.BLOCK[alg]
        assert copied > nblocks // 2
        assert (oracle.decode(alg, enc, data.size) == data).all()


def test_undersized_output_reports_zero():
    data = payload("random", 4096, 1)
    assert oracle.encode("chameleon", data, cap=100).size == 0


@pytest.mark.parametrize("alg", ["chameleon", "cheetah", "lion"])
def test_codec_instance_keeps_its_dictionary(alg, dickens200k):
    """codec.rs:16,72,82: `encode` / `decode` are methods of an instance; the dictionary survives from call to call until clear_state()
    (chameleon.rs:148-150), the protection state is fresh in every call (codec.rs:75,85)."""
    a, b = dickens200k[:70000], dickens200k[70000:150001]
    inst = oracle.Codec(alg)
    ea, eb = inst.encode(a), inst.encode(b)
    assert (ea == oracle.encode(alg, a)).all()                       # a fresh instance behaves like the inherent X::encode
    fresh_b = oracle.encode(alg, b)
    assert eb.size != fresh_b.size or (eb != fresh_b).any()          # the second call saw the first call's dictionary
    dec = oracle.Codec(alg)
    assert (dec.decode(ea, a.size) == a).all() and (dec.decode(eb, b.size) == b).all()
    inst.clear_state()
    assert (inst.encode(b) == fresh_b).all()
