"""Research model for the next round: run-parallel Cheetah DECODE by Jacobi iteration over runs.

Every run is decoded in order (exact reference semantics, cheetah.rs:67-103) from a *carried-in snapshot* of the two tables; the
snapshots of round k+1 are folded from the runs' final tables of round k (touched entries override, untouched inherit). Run 0's
snapshot is exact from the start, so by induction the fixed point is the true decode. Question answered here: how many rounds until
every run's output is exact, as a function of the run count? (Copy-mode blocks and block boundaries are taken as known.)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.proto_cheetah_runs import hashes  # noqa: E402


def true_flags(q):
    """sequential Cheetah encoder (no copy mode): flags + what the decoder reads from the stream"""
    n = q.size
    h = hashes(q).tolist()
    ql = q.tolist()
    pred = {}
    cm = {}
    flags = [0] * n
    last = 0
    for i in range(n):
        v = ql[i]; hh = h[i]
        if pred.get(last, 0) == v:
            flags[i] = 3
        else:
            a, b = cm.get(hh, (0, 0))
            if a == v:
                flags[i] = 1
            else:
                flags[i] = 2 if b == v else 0
                cm[hh] = (v, a)
            pred[last] = v
        last = hh
    return flags, h


def decode_run(lo, hi, flags, h, ql, pred, cm, last):
    """decode quads [lo, hi) from the given table state (dicts are modified); returns decoded values, number wrong, touched keys"""
    out = [0] * (hi - lo)
    tp, tc = set(), set()
    M = 0x9D6EF916
    for i in range(lo, hi):
        f = flags[i]
        if f == 3:
            v = pred.get(last, 0)
            hh = ((v * M) & 0xFFFFFFFF) >> 16                    # cheetah.rs:99-102: a predicted quad re-hashes itself
        else:
            hh = h[i]                                            # in the stream for MAP, computable for PLAIN
            if f == 0:
                v = ql[i]
                a, b = cm.get(hh, (0, 0)); cm[hh] = (v, a); tc.add(hh)
            elif f == 1:
                v = cm.get(hh, (0, 0))[0]
            else:
                a, b = cm.get(hh, (0, 0)); v = b; cm[hh] = (b, a); tc.add(hh)
            pred[last] = v; tp.add(last)
        out[i - lo] = v
        last = hh
    return out, tp, tc, last


def jacobi(q, nruns, max_rounds=40):
    n = q.size
    flags, h = true_flags(q)
    ql = q.tolist()
    bounds = [n * r // nruns for r in range(nruns)] + [n]
    snap_p = [dict() for _ in range(nruns)]                      # carried-in snapshots (round 0: empty = all zero, wrong except run 0)
    snap_c = [dict() for _ in range(nruns)]
    snap_last = [0] * nruns
    hist = []
    for rnd in range(1, max_rounds + 1):
        finals = []
        wrong_runs = 0; wrong_quads = 0
        for r in range(nruns):
            pred = dict(snap_p[r]); cm = dict(snap_c[r])
            out, tp, tc, last = decode_run(bounds[r], bounds[r + 1], flags, h, ql, pred, cm, snap_last[r])
            w = sum(1 for k, v in enumerate(out) if v != ql[bounds[r] + k])
            wrong_quads += w; wrong_runs += w > 0
            finals.append((pred, cm, tp, tc, last))
        hist.append((wrong_runs, wrong_quads))
        # fold: snapshot of run r+1 = snapshot of run r overridden by what run r touched
        new_p, new_c, new_last = [dict()], [dict()], [0]
        for r in range(nruns - 1):
            pred, cm, tp, tc, last = finals[r]
            p = dict(new_p[r]); c = dict(new_c[r])
            for k in tp: p[k] = pred[k]
            for k in tc: c[k] = cm[k]
            new_p.append(p); new_c.append(c); new_last.append(last)
        same = new_p == snap_p and new_c == snap_c and new_last == snap_last
        snap_p, snap_c, snap_last = new_p, new_c, new_last
        if same:
            break
    return hist


def main():
    d = np.fromfile(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "dickens_200k.bin"), np.uint8)
    nbytes = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
    if len(sys.argv) > 2:
        d = np.fromfile(sys.argv[2], np.uint8)
    q = d[:nbytes // 4 * 4].view(np.uint32)
    for nruns in (4, 16, 64):
        hist = jacobi(q, nruns)
        print(f"{q.size} quads, {nruns:3d} runs: rounds to fixed point {len(hist)}; (wrong runs, wrong quads) per round: {hist}")


if __name__ == "__main__":
    main()
