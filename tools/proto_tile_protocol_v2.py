"""Python model of the round-2 Chameleon tile protocol of `cham_flag_pass` (phases A / B / C / F), checked against the in-order
dictionary walk (chameleon.rs:86-101). One "run" with a zero-initialised table; copy mode is not modelled here (the generic tile
variant only masks quads out).

  A  every quad reads old = tab[h]; misser <=> old != f or f == 0 (fingerprint 0 is ambiguous with "empty": always the exact path)
  B  missers publish f (racy: any one of them wins), atomicMin(side[h & (SIDE_N-1)], pos << 16 | h), join their hash-class list
  C  hit candidates re-read tab[h]: unchanged -> flag 1; changed -> flag 1 if the side slot is my bucket's and every misser comes
     later, else the quad joins the class list too ("affected")
  F  per class: exact resolution among the members — predecessor = same-bucket member with the largest smaller position, else the
     pre-tile value; the last member of a bucket leaves its value in the table
"""
import os
import sys

import numpy as np

M = 0x9D6EF916
TILE = 4096
SIDE_N = 4096


def hf(q):
    p = (q.astype(np.uint64) * M) & 0xFFFFFFFF
    return (p >> 16).astype(np.int64), ((p & 0xFFFE) | (q.astype(np.uint64) >> 31)).astype(np.int64)


def reference_flags(q):
    h, f = hf(q)
    tab = {}
    out = np.zeros(q.size, np.uint8)
    for i in range(q.size):
        k = int(h[i])
        cur = tab.get(k)
        if cur is None:
            hit = False      # untouched bucket holds quad 0: only quad 0 (hash 0, fp 0) could match; unresolved first touches are
            out[i] = 2       # reported separately (2 = "first touch of the run": decided by the carry-in)
            tab[k] = int(f[i])
            continue
        hit = cur == int(f[i])
        out[i] = 1 if hit else 0
        if not hit:
            tab[k] = int(f[i])
    return out


def protocol_flags(q, rng):
    h, f = hf(q)
    tab = np.zeros(65536, np.int64)
    vbit = np.zeros(65536, bool)
    out = np.zeros(q.size, np.uint8)
    stats = dict(miss=0, affected=0, maxcls=0)
    for t0 in range(0, q.size, TILE):
        sl = slice(t0, min(q.size, t0 + TILE))
        hh, ff = h[sl], f[sl]
        n = hh.size
        old = tab[hh].copy()
        touched = (old != 0) | vbit[hh]
        miss = (old != ff) | (ff == 0)
        side = np.full(SIDE_N, 0xFFFFFFFF, np.int64)
        members = []                    # (pos, h, f, touched, old)
        midx = np.nonzero(miss)[0]
        order = rng.permutation(midx)   # racy publish order
        for p in order:
            tab[hh[p]] = ff[p]
            s = hh[p] & (SIDE_N - 1)
            side[s] = min(side[s], (p << 16) | hh[p])
            members.append((int(p), int(hh[p]), int(ff[p]), bool(touched[p]), int(old[p]), True))
        ok = np.zeros(n, bool)
        for p in np.nonzero(~miss)[0]:
            if tab[hh[p]] == ff[p]:
                ok[p] = True
                continue
            slot = side[hh[p] & (SIDE_N - 1)]
            if (slot & 0xFFFF) == hh[p] and p < (slot >> 16):
                ok[p] = True
            else:
                members.append((int(p), int(hh[p]), int(ff[p]), True, int(ff[p]), False))
                stats["affected"] += 1
        stats["miss"] += midx.size
        # F: exact resolution per bucket (class lists only partition the work)
        bybucket = {}
        for m in members:
            bybucket.setdefault(m[1], []).append(m)
        cls = {}
        for m in members:
            cls[m[1] >> 11] = cls.get(m[1] >> 11, 0) + 1
        if cls:
            stats["maxcls"] = max(stats["maxcls"], max(cls.values()))
        flags = ok.astype(np.uint8)
        for b, lst in bybucket.items():
            for (p, _, fv, tch, oldv, _) in lst:
                pred = [m for m in lst if m[0] < p]
                if pred:
                    best = max(pred, key=lambda m: m[0])
                    flags[p] = 1 if best[2] == fv else 0
                elif tch:
                    flags[p] = 1 if oldv == fv else 0
                else:
                    flags[p] = 2
            last = max(lst, key=lambda m: m[0])
            tab[b] = last[2]
            if last[2] == 0:
                vbit[b] = True
        out[sl] = flags
    return out, stats


def main():
    path = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "dickens_200k.bin")
    d = np.fromfile(path, np.uint8)
    rng = np.random.default_rng(3)
    cases = {
        "dickens": d[:160000],
        "zeros": np.zeros(40000, np.uint8),
        "low": rng.integers(0, 3, 60000, dtype=np.uint8),
        "random": rng.integers(0, 256, 60000, dtype=np.uint8),
        "fp0": np.tile(np.array([0, 0, 0x80000000, 0, 7, 0x80000000], np.uint32).view(np.uint8), 3000),
    }
    ok = True
    for name, data in cases.items():
        q = data[: data.size // 4 * 4].view(np.uint32)
        want = reference_flags(q)
        got, st = protocol_flags(q, rng)
        good = bool((want == got).all())
        ok &= good
        print(f"{name:8s} quads {q.size:6d} missers {st['miss']:6d} affected {st['affected']:5d} max class {st['maxcls']:4d} {'OK' if good else 'MISMATCH'}")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
