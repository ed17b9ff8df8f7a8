"""Python model of the run-parallel Cheetah encode decomposition planned for round 2.

Validates, against the oracle's byte stream, that Cheetah's flags (cheetah.rs:121-150) decompose into

  1. PREDICTED_i  <=>  quad_i == (the quad that followed the previous occurrence of the same CONTEXT),
     context_i = hash of the previous *encoded* quad (0 at the stream start); pred table starts as "0 everywhere".
     -> a previous-in-bucket equality problem on (context, quad): every run can be processed independently with an
        "unresolved" first touch per context that is patched from a per-run carry-in table (left fold of the runs' last-writer tables).
  2. On the subsequence of non-predicted quads, per hash bucket an MRU-2 of fingerprints (a, b):
        v == a -> MAP_A (no change); v == b -> MAP_B, (a, b) <- (v, a); else PLAIN, (a, b) <- (v, a).
     Per run and bucket at most TWO accesses cannot be decided without the carried-in state (first touch; first access that differs
     from it). Runs export (T0 untouched | T1 a known | T2 a,b known); the fold across runs and the late resolution are below.
  3. Copy-mode blocks (codec/protection_state.rs) are hidden from both tables and from the context chain; the copy map is the fixed
     point of  M -> automaton(incompressible bits computed under M)  (same iteration as the Chameleon encoder).

The model reproduces the oracle's stream byte for byte (flags -> signatures + payload) on text, random, zeros and mixed inputs.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402

M32 = np.uint64(0x9D6EF916)
BLOCK_Q = 32


def hashes(q):
    return ((q.astype(np.uint64) * M32) & np.uint64(0xFFFFFFFF)).astype(np.uint32) >> np.uint32(16)


def flags_given_copymap(q, copied, nruns):
    """q: uint32 quads (whole quads only); copied: bool per block. Returns flags (0 plain,1 A,2 B,3 P) for quads of encoded blocks."""
    n = q.size
    h = hashes(q).astype(np.int64)
    nblk = (n + BLOCK_Q - 1) // BLOCK_Q
    enc = ~np.repeat(copied[:nblk], BLOCK_Q)[:n]
    idx = np.nonzero(enc)[0]                       # encoded quads in stream order
    ctx = np.zeros(n, np.int64)
    ctx[idx[1:]] = h[idx[:-1]]                     # context = hash of the previous encoded quad; first one: 0
    flags = np.zeros(n, np.int8)
    bounds = [(nblk * r // nruns) * BLOCK_Q for r in range(nruns)] + [n]

    # ---- pass P per run, then fold + resolve ------------------------------------------------------------------------
    P = np.zeros(n, bool)
    unres_p = []                                   # (run, i)
    finals_p = []
    for r in range(nruns):
        tab = {}                                   # ctx -> quad
        for i in range(bounds[r], bounds[r + 1]):
            if not enc[i]:
                continue
            c = ctx[i]
            if c in tab:
                P[i] = tab[c] == q[i]
            else:
                unres_p.append((r, i))
            tab[c] = q[i]
        finals_p.append(tab)
    carry = {}
    carries = []
    for r in range(nruns):
        carries.append(dict(carry))
        carry.update(finals_p[r])
    for r, i in unres_p:
        P[i] = carries[r].get(ctx[i], np.uint32(0)) == q[i]     # pred table starts as 0 everywhere (cheetah.rs:53)

    # ---- pass C per run on the non-predicted quads ----------------------------------------------------------------
    code = np.zeros(n, np.int8)                    # 0 plain 1 A 2 B
    finals_c = []
    unres_c = []                                   # (run, bucket, i1, i2)
    for r in range(nruns):
        st = {}                                    # bucket -> [T, a, b, i1, i2]
        for i in range(bounds[r], bounds[r + 1]):
            if not enc[i] or P[i]:
                continue
            b = h[i]; v = q[i]
            s = st.get(b)
            if s is None:
                st[b] = [1, v, None, i, None]      # T1: a known; flag of i unresolved (#1)
            elif s[0] == 1:
                if v == s[1]:
                    code[i] = 1
                else:
                    s[4] = i                       # unresolved #2: MAP_B iff v == (unknown) b
                    s[0], s[2], s[1] = 2, s[1], v
            else:
                if v == s[1]:
                    code[i] = 1
                else:
                    code[i] = 2 if v == s[2] else 0
                    s[2], s[1] = s[1], v
        finals_c.append(st)
    carry = {}                                     # bucket -> (a, b); missing = (0, 0) (cheetah.rs:52)
    for r in range(nruns):
        for b, s in finals_c[r].items():
            a0, b0 = carry.get(b, (np.uint32(0), np.uint32(0)))
            i1, i2 = s[3], s[4]
            v1 = q[i1]
            if v1 == a0:
                code[i1] = 1; b1 = b0
            else:
                code[i1] = 2 if v1 == b0 else 0; b1 = a0
            if i2 is not None:
                code[i2] = 2 if q[i2] == b1 else 0
            # state after the run
            if s[0] == 2:
                carry[b] = (s[1], s[2])
            else:                                   # only one value accessed in this run
                carry[b] = (a0, b0) if v1 == a0 else (v1, a0)
    flags[:] = np.where(P, 3, code)
    flags[~enc] = -1
    return flags


def automaton(inc, nblk):
    """codec/protection_state.rs over blocks; inc[b] consumed only for encoded blocks. Returns copied[]."""
    copied = np.zeros(nblk, bool)
    pen, start, prev = 0, 1, False
    for b in range(nblk):
        if (b & 15) == 0 and start > 1:
            start >>= 1
        if pen > 0:
            copied[b] = True
            pen -= 1
            if pen == 0:
                start += 1
        else:
            if inc[b]:
                if prev:
                    pen = start
                prev = True
            else:
                prev = False
    return copied


def encode(data, nruns):
    n = data.size
    nq = n // 4
    q = data[:nq * 4].view(np.uint32)
    nblk = (n + 127) // 128
    copied = np.zeros(nblk, bool)
    inc = np.zeros(nblk, bool)
    rounds = 0
    while True:
        flags = flags_given_copymap(q, copied, nruns)
        # incompressible bits of the encoded blocks (codec.rs:68): 8 + 4*plain + 2*map (+ tail) >= 128
        for b in range(nblk):
            if copied[b]:
                continue
            fb = flags[b * 32:(b + 1) * 32]
            blen = min(128, n - b * 128)
            size = 8 + 4 * int((fb == 0).sum()) + 2 * int(((fb == 1) | (fb == 2)).sum()) + (blen & 3)
            inc[b] = size >= 128
        new = automaton(inc, nblk)
        rounds += 1
        if (new == copied).all():
            break
        copied = new
    # ---- emit --------------------------------------------------------------------------------------------------------
    h = hashes(q)
    out = bytearray()
    for b in range(nblk):
        blk = data[b * 128:(b + 1) * 128]
        if copied[b]:
            out += blk.tobytes()
            continue
        sig = 0; payload = bytearray()
        for k in range(len(blk) // 4):
            i = b * 32 + k
            fl = int(flags[i])
            sig |= fl << (2 * k)
            if fl == 0:
                payload += int(q[i]).to_bytes(4, "little")
            elif fl in (1, 2):
                payload += int(h[i]).to_bytes(2, "little")
        out += sig.to_bytes(8, "little") + payload + blk[(len(blk) // 4) * 4:].tobytes()
    return np.frombuffer(bytes(out), np.uint8), rounds, int(copied.sum())


def main():
    d = np.fromfile(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "dickens_200k.bin"), np.uint8)
    rng = np.random.default_rng(1)
    cases = {
        "kat": np.frombuffer(b"test" * 31 + b"t", np.uint8),
        "dickens60k": d[:60003],
        "zeros": np.zeros(20000, np.uint8),
        "random": rng.integers(0, 256, 30001, dtype=np.uint8),
        "mixed": np.concatenate([d[:20000], rng.integers(0, 256, 9000, dtype=np.uint8), np.zeros(5000, np.uint8), d[50000:70002]]),
        "low": rng.integers(0, 3, 30000, dtype=np.uint8),
    }
    ok = True
    for name, data in cases.items():
        want = oracle.encode("cheetah", data)
        for nruns in (1, 3, 7):
            got, rounds, ncopied = encode(data, nruns)
            good = got.size == want.size and bool((got == want).all())
            ok &= good
            print(f"{name:10s} runs={nruns} bytes={data.size:6d} out={got.size:6d} copied_blocks={ncopied:4d} fixed-point rounds={rounds} {'OK' if good else 'MISMATCH'}")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
