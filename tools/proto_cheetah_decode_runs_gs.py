"""Research model: prediction pass with IN-ORDER evaluation inside a run (so chains of consecutive predicted quads
advance a whole run per round) + carried-in snapshots from the previous round + "unknown" instead of garbage:

  * a read of a context the run has not written yet uses the snapshot folded in the PREVIOUS round; in round 1 only run 0 has one
    (the zero table), elsewhere the read is UNKNOWN: its hash is unknown, so the next quad's context is unknown, so a non-predicted
    quad's write is SKIPPED this round (nothing wrong is ever written) and a predicted quad stays unknown;
  * after the round the snapshots are refolded from the runs' (known) writes on top of the zero table.

Compare tools/proto_cheetah_decode_jacobi.py (same structure, but unknowns were decoded as garbage: one run per round)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.proto_cheetah_decode_jacobi import true_flags  # noqa: E402

M = 0x9D6EF916
UNK = None


def chase(start, steps, table):
    """x -> hash(table[x]) for `steps` links on a STATIC table (nothing writes inside a stretch of predicted quads): stop early at a
    fixed point or cycle and jump the rest."""
    x = start; seen = {}; k = 0
    while k < steps:
        if x in seen:                                   # cycle of length p: skip whole periods
            p = k - seen[x]
            k += ((steps - k) // p) * p
            seen = {}
            if k >= steps:
                break
        seen[x] = k
        x = ((table.get(x, 0) * M) & 0xFFFFFFFF) >> 16
        k += 1
    return x


def iterate(q, nruns, max_rounds=200, sweep=False):
    flags, h = true_flags(q)
    ql = q.tolist(); n = len(ql)
    bounds = [n * r // nruns for r in range(nruns)] + [n]
    snap = [dict() if r == 0 else None for r in range(nruns)]       # None: no snapshot yet (every untouched read is unknown)
    carry_ctx = [0] + [UNK] * (nruns - 1)                            # context of the first quad of each run
    Hprev = None
    hist = []
    for rnd in range(1, max_rounds + 1):
        finals = []; H = [UNK] * n; val = [UNK] * n
        for r in range(nruns):
            tab = {}                                                 # this run's writes (known contexts only)
            ctx = carry_ctx[r]
            for i in range(bounds[r], bounds[r + 1]):
                if flags[i] == 3:
                    if ctx is UNK:
                        hh = UNK
                    elif ctx in tab:
                        val[i] = tab[ctx]; hh = ((val[i] * M) & 0xFFFFFFFF) >> 16
                    elif snap[r] is not None:
                        val[i] = snap[r].get(ctx, 0); hh = ((val[i] * M) & 0xFFFFFFFF) >> 16
                    else:
                        hh = UNK
                else:
                    val[i] = ql[i]; hh = h[i]                        # chunk-map values are known (pass 1)
                    if ctx is not UNK:
                        tab[ctx] = ql[i]
                H[i] = hh
                ctx = hh
            finals.append((tab, ctx))
        # refold
        acc = {}
        new_snap = []; new_carry = []
        c = 0
        for r in range(nruns):
            new_snap.append(dict(acc)); new_carry.append(c)
            acc.update(finals[r][0]); c = finals[r][1]
        if sweep:
            # carry the context through runs that consist of predicted quads only (flags are known): their tables are static, so the last
            # hash of run r is a chase of (run length) links from its first context on the snapshot run r starts from
            for r in range(1, nruns):
                if new_carry[r] is UNK and new_carry[r - 1] is not UNK and all(f == 3 for f in flags[bounds[r - 1]:bounds[r]]):
                    new_carry[r] = chase(new_carry[r - 1], bounds[r] - bounds[r - 1], new_snap[r - 1])
        wrong = sum(1 for i in range(n) if flags[i] == 3 and val[i] != ql[i])
        unknown = sum(1 for i in range(n) if H[i] is UNK)
        same = H == Hprev and new_snap == snap and new_carry == carry_ctx
        hist.append((unknown, wrong))
        snap, carry_ctx, Hprev = new_snap, new_carry, H
        if same:
            break
    return hist


def main():
    path = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "dickens_200k.bin")
    nbytes = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
    d = np.fromfile(path, np.uint8)[:nbytes]
    cases = {"text": d, "zeros": np.zeros(40000, np.uint8),
             "mixed": np.concatenate([d[:20000], np.zeros(8000, np.uint8), d[30000:50000]])}
    for name, data in cases.items():
        q = data[:data.size // 4 * 4].view(np.uint32)
        for nruns in (4, 16, 64):
            for sweep in (False, True):
                hist = iterate(q, nruns, sweep=sweep)
                print(f"{name:6s} {q.size:7d} quads {nruns:3d} runs, carry sweep {'on ' if sweep else 'off'}: rounds {len(hist):3d}; (unknown hashes, wrong predicted values) per round: {hist[:8]}{' ...' if len(hist) > 8 else ''}")


if __name__ == "__main__":
    main()
