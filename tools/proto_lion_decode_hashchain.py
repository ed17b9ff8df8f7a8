"""Research model for the next round, Lion flavour of tools/proto_cheetah_decode_hashchain.py.

Lion's predicted quads DO reorder the context's 5-deep list (lion.rs:214-262), but only as a function of the flag (the depth k is in the
stream): in HASH space the serial state is a table context -> 5 x 16-bit hashes (640 KiB: a 4-CTA cluster's distributed shared memory):

    c = H[i-1];  flag 1..5 (depth k = flag-1): H[i] = T[c][k], rotate T[c][0..k];   else: H[i] = hash from the stream, shift it into T[c]

Values then follow per context: the same list operations on 32-bit values, replayed per context (contexts are independent once known),
and the chunk-map values of the MAP quads never depend on predicted quads. Checked here on dickens (copy mode ignored)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.proto_cheetah_runs import hashes  # noqa: E402


def lion_flags(q):
    n = q.size
    h = hashes(q).tolist(); ql = q.tolist()
    pred = {}; cm = {}; flags = [0] * n; last = 0
    for i in range(n):
        v = ql[i]; hh = h[i]
        p = pred.setdefault(last, [0, 0, 0, 0, 0])
        if v in p:
            k = p.index(v); flags[i] = k + 1
            p.insert(0, p.pop(k))
        else:
            a, b = cm.get(hh, (0, 0))
            if a == v:
                flags[i] = 6
            else:
                flags[i] = 7 if b == v else 0
                cm[hh] = (v, a)
            p.pop(); p.insert(0, v)
        last = hh
    return flags, h


def main():
    d = np.fromfile(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "dickens_200k.bin"), np.uint8)
    q = d[:200000].view(np.uint32)
    flags, h = lion_flags(q)
    ql = q.tolist(); n = len(ql)
    # pass 1: chunk-map values (non-predicted quads only)
    val = [None] * n; cm = {}
    for i in range(n):
        f = flags[i]
        if 1 <= f <= 5:
            continue
        a, b = cm.get(h[i], (0, 0))
        if f == 0:
            v = ql[i]; cm[h[i]] = (v, a)
        elif f == 6:
            v = a
        else:
            v = b; cm[h[i]] = (b, a)
        val[i] = v
    # pass 2: hash chain over a table of 5 hashes per context
    T = {}; H = [0] * n; c = 0
    for i in range(n):
        p = T.setdefault(c, [0, 0, 0, 0, 0])
        f = flags[i]
        if 1 <= f <= 5:
            H[i] = p[f - 1]; p.insert(0, p.pop(f - 1))
        else:
            H[i] = h[i]; p.pop(); p.insert(0, H[i])
        c = H[i]
    assert H == h, "hash chain differs"
    # pass 3: values, replayed per context (contexts independent)
    by_ctx = {}
    ctx = 0
    for i in range(n):
        by_ctx.setdefault(ctx, []).append(i); ctx = H[i]
    wrong = 0
    for cx, idxs in by_ctx.items():
        p = [0, 0, 0, 0, 0]
        for i in idxs:
            f = flags[i]
            if 1 <= f <= 5:
                v = p[f - 1]; p.insert(0, p.pop(f - 1)); wrong += v != ql[i]
            else:
                p.pop(); p.insert(0, val[i])
    npred = sum(1 for f in flags if 1 <= f <= 5)
    print(f"{n} quads, {npred} predicted ({100 * npred / n:.1f} %), wrong after the three passes: {wrong}")
    sys.exit(1 if wrong else 0)


if __name__ == "__main__":
    main()
