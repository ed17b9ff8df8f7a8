"""Research model for the next round (DESIGN.md §9 item 2, step ii): the chunk-map VALUES of a Cheetah stream, run-parallel.

In decode the flags are given and the values are not: per bucket  PLAIN(v): (a,b) <- (v,a), value v;  MAP_A: value a;
MAP_B: value b, (a,b) <- (b,a)   (cheetah.rs:76-96). Predicted quads never touch the chunk map. A run that does not know the carried-in
(a0,b0) of a bucket cannot name the values of its MAP quads, and there can be any number of those — so, unlike the encoder, a
pre-pass computes only each run's TRANSFER FUNCTION per bucket (final a and b, each either a literal or the symbol A0 / B0), the fold
substitutes run after run, and the real pass starts every run from its concrete carry-in (the same two-pass shape as the Chameleon
decoder's writer pre-pass). Checked against the sequential definition on dickens with several run counts."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.proto_cheetah_decode_jacobi import true_flags  # noqa: E402

A0, B0 = ("sym", "a0"), ("sym", "b0")


def run_parallel_values(flags, h, ql, nruns):
    n = len(ql)
    bounds = [n * r // nruns for r in range(nruns)] + [n]
    # pre-pass: transfer function per run and bucket
    tf = []
    for r in range(nruns):
        st = {}
        for i in range(bounds[r], bounds[r + 1]):
            f = flags[i]
            if f == 3 or f == 1:
                continue
            a, b = st.get(h[i], (A0, B0))
            st[h[i]] = (ql[i], a) if f == 0 else (b, a)
        tf.append(st)
    # fold
    carry = {}
    carries = []
    for r in range(nruns):
        carries.append(dict(carry))
        for k, (a, b) in tf[r].items():
            a0, b0 = carry.get(k, (0, 0))
            sub = lambda x: a0 if x == A0 else b0 if x == B0 else x
            carry[k] = (sub(a), sub(b))
    # real pass
    val = [None] * n
    for r in range(nruns):
        cm = dict(carries[r])
        for i in range(bounds[r], bounds[r + 1]):
            f = flags[i]
            if f == 3:
                continue
            a, b = cm.get(h[i], (0, 0))
            if f == 0:
                v = ql[i]; cm[h[i]] = (v, a)
            elif f == 1:
                v = a
            else:
                v = b; cm[h[i]] = (b, a)
            val[i] = v
    return val


def main():
    d = np.fromfile(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "dickens_200k.bin"), np.uint8)
    q = d[:200000].view(np.uint32)
    flags, h = true_flags(q)
    ql = q.tolist()
    ok = True
    for nruns in (1, 3, 17, 64):
        val = run_parallel_values(flags, h, ql, nruns)
        wrong = sum(1 for i in range(len(ql)) if flags[i] != 3 and val[i] != ql[i])
        ok &= wrong == 0
        print(f"runs={nruns:3d}: wrong chunk-map values {wrong}")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
