"""Python model of the run-parallel Lion encode decomposition.

Validates against the oracle's byte stream that Lion's flags (lion.rs:209-271) decompose into

  1. per CONTEXT (hash of the previous *encoded* quad, 0 at the stream start) a 5-deep move-to-front list (lion.rs:43-57,214-262):
     found at depth k -> PREDICTED_{A..E}[k], entries [0..k] rotate; not found -> the quad goes through the chunk map and is shifted in.
     Inside a run the values it has already put into the list sit at the FRONT in recency order, ahead of what is left of the
     carried-in list; so an access is decidable locally unless the quad is not in the run-local list while that list has m < 5
     entries -> at most 5 undecided accesses per run and context. The fold walks the runs in order per context, replays those
     accesses against the carried-in list (hit at position j of the remainder -> depth m + j, that entry leaves the remainder; miss ->
     the remainder's visible part shrinks by one) and carries  local list + remainder, cut to 5  on.
  2. the chunk map on the NOT-predicted quads is Cheetah's MRU-2 (tools/proto_cheetah_runs.py).
  3. copy-mode blocks (64 B) hidden as in Cheetah; copy map by fixed-point iteration.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
from tools.proto_cheetah_runs import hashes, automaton  # noqa: E402

BLOCK_Q = 16


def flags_given_copymap(q, copied, nruns):
    n = q.size
    h = hashes(q).astype(np.int64)
    nblk = (n + BLOCK_Q - 1) // BLOCK_Q
    enc = ~np.repeat(copied[:nblk], BLOCK_Q)[:n]
    idx = np.nonzero(enc)[0]
    ctx = np.zeros(n, np.int64)
    ctx[idx[1:]] = h[idx[:-1]]
    bounds = [(nblk * r // nruns) * BLOCK_Q for r in range(nruns)] + [n]

    pcode = np.zeros(n, np.int8)                   # 0 not predicted, 1..5 depth + 1
    finals = []                                    # per run: ctx -> (local list, [undecided quad indices])
    for r in range(nruns):
        tab = {}
        for i in range(bounds[r], bounds[r + 1]):
            if not enc[i]:
                continue
            loc, und = tab.setdefault(ctx[i], ([], []))
            v = int(q[i])
            if v in loc:
                k = loc.index(v)
                pcode[i] = k + 1
                loc.insert(0, loc.pop(k))
            else:
                if len(loc) < 5:
                    und.append(i)                  # depends on the carried-in list
                else:
                    loc.pop()
                loc.insert(0, v)
        finals.append(tab)
    carry = {}                                     # ctx -> list of 5 (missing = five zeros, lion.rs:64-72)
    for r in range(nruns):
        for c, (loc, und) in finals[r].items():
            rem = list(carry.get(c, [0, 0, 0, 0, 0]))
            for m, i in enumerate(und):            # the m-th undecided access happened with m local entries in front
                v = int(q[i])
                vis = rem[:5 - m]
                if v in vis:
                    j = vis.index(v)
                    pcode[i] = m + j + 1
                    del rem[j]
                else:
                    pcode[i] = 0
                    rem = rem[:5 - m - 1]
            carry[c] = (loc + rem)[:5]

    # chunk map on the non-predicted quads: MRU-2 per bucket, as in Cheetah (sequential here; the run decomposition of this part is
    # validated by proto_cheetah_runs.py)
    code = np.zeros(n, np.int8)
    cm = {}
    for i in idx:
        if pcode[i]:
            continue
        a, b = cm.get(h[i], (0, 0))
        v = int(q[i])
        if v == a:
            code[i] = 6
        else:
            code[i] = 7 if v == b else 0
            cm[h[i]] = (v, a)
    flags = np.where(pcode > 0, pcode, code)
    flags[~enc] = -1
    return flags


def encode(data, nruns):
    n = data.size
    nq = n // 4
    q = data[:nq * 4].view(np.uint32)
    nblk = (n + 63) // 64
    copied = np.zeros(nblk, bool)
    inc = np.zeros(nblk, bool)
    rounds = 0
    while True:
        flags = flags_given_copymap(q, copied, nruns)
        for b in range(nblk):
            if copied[b]:
                continue
            fb = flags[b * 16:(b + 1) * 16]
            blen = min(64, n - b * 64)
            size = 6 + 4 * int((fb == 0).sum()) + 2 * int((fb >= 6).sum()) + (blen & 3)
            inc[b] = size >= 64
        new = automaton(inc, nblk)
        rounds += 1
        if (new == copied).all():
            break
        copied = new
    h = hashes(q)
    out = bytearray()
    for b in range(nblk):
        blk = data[b * 64:(b + 1) * 64]
        if copied[b]:
            out += blk.tobytes()
            continue
        sig = 0; payload = bytearray()
        for k in range(len(blk) // 4):
            i = b * 16 + k
            fl = int(flags[i])
            sig |= fl << (3 * k)
            if fl == 0:
                payload += int(q[i]).to_bytes(4, "little")
            elif fl >= 6:
                payload += int(h[i]).to_bytes(2, "little")
        out += sig.to_bytes(6, "little") + payload + blk[(len(blk) // 4) * 4:].tobytes()
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
        "low2": np.repeat(rng.integers(0, 2, 12000, dtype=np.uint8), 2)[:20001],
    }
    ok = True
    for name, data in cases.items():
        want = oracle.encode("lion", data)
        for nruns in (1, 3, 7, 40):
            got, rounds, ncopied = encode(data, nruns)
            good = got.size == want.size and bool((got == want).all())
            ok &= good
            print(f"{name:10s} runs={nruns:2d} bytes={data.size:6d} out={got.size:6d} copied_blocks={ncopied:4d} fixed-point rounds={rounds} {'OK' if good else 'MISMATCH'}")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
