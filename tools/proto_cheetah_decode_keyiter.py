"""Research model: Cheetah decode by GLOBAL fixed-point iteration on the contexts.

Facts used (tools/proto_cheetah_decode_hashchain.py): value(predicted quad i) = value of the latest non-predicted quad j < i with
context_j == context_i; only the contexts of quads that FOLLOW a predicted quad are not in the stream. Iteration: keep an estimate of
every context; one round = a fully parallel previous-occurrence pass ("latest writer with my key", the encoder's pass P + fold) under
the current estimates, which yields new hashes for the predicted quads and hence new estimates for their successors. Round 1 leaves
the writers with unknown context out. Exact at the fixed point (induction over the position). How many rounds?
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.proto_cheetah_decode_jacobi import true_flags  # noqa: E402

M = 0x9D6EF916


def iterate(q, max_rounds=60):
    flags, h = true_flags(q)
    ql = q.tolist()
    n = len(ql)
    pred = [f == 3 for f in flags]
    # contexts: exact where the previous quad is non-predicted (its hash is in the stream / computable), else estimated
    ctx = [None] * n
    ctx[0] = 0
    for i in range(1, n):
        if not pred[i - 1]:
            ctx[i] = h[i - 1]
    true_ctx = [0] + h[:-1]
    hist = []
    for rnd in range(1, max_rounds + 1):
        last = {}                      # key -> value of the latest non-predicted quad with that (estimated) context
        newH = [None] * n
        for i in range(n):             # sequential scan = what the parallel previous-occurrence pass computes for fixed keys
            c = ctx[i]
            if pred[i]:
                if c is not None:
                    v = last.get(c, 0)
                    newH[i] = ((v * M) & 0xFFFFFFFF) >> 16
            elif c is not None:
                last[c] = ql[i]
        changed = 0
        for i in range(1, n):
            if pred[i - 1]:
                if ctx[i] != newH[i - 1]:
                    changed += 1
                ctx[i] = newH[i - 1]
        wrong = sum(1 for i in range(n) if ctx[i] != true_ctx[i])
        hist.append((changed, wrong))
        if changed == 0:
            break
    return hist, sum(pred), n


def main():
    path = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "dickens_200k.bin")
    nbytes = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
    d = np.fromfile(path, np.uint8)[:nbytes]
    q = d[:d.size // 4 * 4].view(np.uint32)
    hist, npred, n = iterate(q)
    print(f"{n} quads, {npred} predicted: rounds {len(hist)}; (contexts changed, contexts wrong) per round: {hist}")


if __name__ == "__main__":
    main()
