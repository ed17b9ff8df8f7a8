"""Research model for the next round: Cheetah DECODE split into two parallel passes around one small serial one.

Observation: a PREDICTED quad never changes the prediction table's VALUES (cheetah.rs:125,144,148 — it would store back what it just
read), so   value(P quad i) = value of the latest NON-predicted quad j < i with context_j == context_i   (0 if none),
where context_k = hash of quad k-1. Values of non-predicted quads do not depend on predicted quads at all (PLAIN: literal; MAP_A/B: the
chunk map, which only non-predicted quads touch). The only thing that is serial is the CONTEXT of a quad that follows a predicted quad:
hash(value of that predicted quad) = hash of its source quad. In hash space that is a tiny automaton over a table T: context -> 16-bit
hash (128 KiB: fits one SM's shared memory):

    for i in order:  c = H[i-1];  if predicted(i): H[i] = T[c]   else: H[i] = (from the stream / literal);  T[c] = H[i]

Pass 1 (parallel, encoder-like run/fold): chunk-map values of the MAP quads.          Pass 2 (serial state, one SM): H[] as above.
Pass 3 (parallel): with every context known, a predicted quad's value is a previous-occurrence lookup keyed by context.

This script checks the decomposition on dickens (copy mode ignored: boundaries / copy flags are the boundary walk's job).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.proto_cheetah_decode_jacobi import true_flags  # noqa: E402

M = 0x9D6EF916


def main():
    d = np.fromfile(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "dickens_200k.bin"), np.uint8)
    q = d[:200000].view(np.uint32)
    flags, h = true_flags(q)
    ql = q.tolist()
    n = len(ql)
    # pass 1: values of the non-predicted quads (sequential here; per bucket MRU-2, independent of predicted quads)
    val = [None] * n
    cm = {}
    for i in range(n):
        f = flags[i]
        if f == 3:
            continue
        hh = h[i]
        a, b = cm.get(hh, (0, 0))
        if f == 0:
            v = ql[i]; cm[hh] = (v, a)
        elif f == 1:
            v = a
        else:
            v = b; cm[hh] = (b, a)
        val[i] = v
    assert all(val[i] == ql[i] for i in range(n) if flags[i] != 3)
    # pass 2: hash chain (the only serial state: 65536 x 16 bit)
    T = [0] * 65536            # hash of 0 is 0: the table starts as "value 0 everywhere" (cheetah.rs:53)
    H = [0] * n
    c = 0
    for i in range(n):
        if flags[i] == 3:
            H[i] = T[c]
        else:
            H[i] = h[i]
            T[c] = H[i]
        c = H[i]
    assert H == h, "hash chain differs"
    # pass 3: predicted values = latest earlier non-predicted quad with the same context (contexts now all known)
    last = {}
    ctx = 0
    wrong = 0
    for i in range(n):
        if flags[i] == 3:
            v = last.get(ctx, 0)
            wrong += v != ql[i]
        else:
            last[ctx] = val[i]
        ctx = H[i]
    npred = sum(1 for f in flags if f == 3)
    runs = []
    k = 0
    for f in flags:
        if f == 3:
            k += 1
        else:
            if k: runs.append(k)
            k = 0
    print(f"{n} quads, {npred} predicted ({100 * npred / n:.1f} %), wrong after the three passes: {wrong}")
    print(f"consecutive-predicted chains: {len(runs)}, mean length {np.mean(runs):.2f}, max {max(runs)} "
          f"(the serial pass is latency-bound only along these)")
    sys.exit(1 if wrong else 0)


if __name__ == "__main__":
    main()
