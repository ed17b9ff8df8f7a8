"""Numpy model of the Chameleon flag-pass tile protocol (density_b200/csrc/chameleon_encode.cu).

Design-validation tool, not product code: simulates the racy shared-memory rounds with a
random winner per address and checks the resulting flags against a sequential evaluation of
chameleon.rs:86-101 (protection ignored: the fast path is only accepted when it is quiet).
"""
import sys
import numpy as np

M = np.uint32(0x9D6EF916)
MARK = 0xFFFF


def hf(q):
    p = (q.astype(np.uint64) * np.uint64(M)).astype(np.uint32)
    h = (p >> np.uint32(16)).astype(np.int64)
    f = ((p & np.uint32(0xFFFE)) | (q >> np.uint32(31))).astype(np.int64)
    return h, f


def seq_flags(q):
    h, f = hf(q)
    d = np.zeros(65536, np.uint32)
    out = np.zeros(q.size, np.uint8)
    qq = q.tolist(); hh = h.tolist()
    dl = d.tolist()
    for i in range(q.size):
        if dl[hh[i]] == qq[i]:
            out[i] = 1
        else:
            dl[hh[i]] = qq[i]
    return out


def racy_store(arr, idx, val, rng):
    """all stores land; a random one wins per address"""
    if idx.size == 0:
        return
    perm = rng.permutation(idx.size)
    arr[idx[perm]] = val[perm]


def run_pass(q, T, S, rng):
    """one run (protocol v2, as implemented in chameleon_encode.cu). Phases separated by CTA barriers:
       A read old | B missers: racy store f, atomicMin side | C readback + conflict bits | D classify, publish, restore | F slow path"""
    h, f = hf(q)
    n = q.size
    tab = np.zeros(65536, np.int64)
    vbit = np.zeros(65536, bool)
    side = np.full(S, 0xFFFFFFFF, np.int64)
    conf = np.zeros(65536, bool)
    flags = np.zeros(n, np.uint8)
    unres = []
    nslow = 0
    for t0 in range(0, n, T):
        sl = slice(t0, min(n, t0 + T))
        th, tf = h[sl], f[sl]
        m = th.size
        pos = np.arange(m)
        ss = th & (S - 1)
        key = (pos << 16) | th
        # A
        old = tab[th].copy()
        touched = (old != 0) | vbit[th]
        hit = touched & (old == tf)
        miss = ~hit
        # B
        mi = np.nonzero(miss)[0]
        racy_store(tab, th[mi], tf[mi], rng)
        np.minimum.at(side, ss[mi], key[mi])
        # C
        w = tab[th].copy()
        slot = side[ss].copy()
        slot_h = slot & 0xFFFF
        slot_pos = slot >> 16
        same = slot_h == th
        slow = np.zeros(m, bool)
        done1 = np.zeros(m, bool)
        # hit members
        hm = hit & (w != tf)                    # a misser exists in my bucket
        before = hm & same & (pos < slot_pos)   # all missers of my bucket come after me
        done1 |= hit & ~hm
        done1 |= before
        hslow = hm & ~before
        # missers
        mconf = miss & (~same | (w != tf))
        setters = hslow | mconf
        conf[th[setters]] = True
        slow |= setters
        # D
        tent = miss & ~mconf
        dconf = tent & conf[th]
        slow |= dconf
        uni = tent & ~dconf
        first = uni & (slot == key)
        flags[t0 + np.nonzero(done1)[0]] = 1
        flags[t0 + np.nonzero(uni & ~first)[0]] = 1
        for i in np.nonzero(first & ~touched)[0]:
            unres.append((t0 + i, th[i], tf[i]))
        z = np.nonzero(first & (tf == 0))[0]
        vbit[th[z]] = True
        # restore + resets
        si = np.nonzero(slow)[0]
        tab[th[si]] = old[si]
        conf[th[setters]] = False
        side[ss[mi]] = 0xFFFFFFFF
        assert not conf.any()
        # F: slow path in position order
        nslow += si.size
        for i in si:
            hh, ff = th[i], tf[i]
            tch = tab[hh] != 0 or vbit[hh]
            if tch and tab[hh] == ff:
                flags[t0 + i] = 1
            else:
                if not tch:
                    unres.append((t0 + i, hh, ff))
                tab[hh] = ff
                if ff == 0:
                    vbit[hh] = True
    touched = (tab != 0) | vbit
    return flags, unres, touched, tab, nslow


def encode_flags(q, nruns, T, S, seed):
    rng = np.random.default_rng(seed)
    n = q.size
    bounds = [(n * r) // nruns for r in range(nruns + 1)]
    # align run starts to blocks of 64 quads
    bounds = [min(n, (b // 64) * 64) for b in bounds[:-1]] + [n]
    flags = np.zeros(n, np.uint8)
    carry_valid = np.zeros(65536, bool); carry_valid[0] = True   # initial dictionary: bucket 0 holds quad 0
    carry_f = np.zeros(65536, np.int64)
    tot_unres = 0; tot_slow = 0
    finals = []
    for r in range(nruns):
        a, b = bounds[r], bounds[r + 1]
        fl, unres, touched, tab, nslow = run_pass(q[a:b], T, S, rng)
        flags[a:b] = fl
        tot_slow += nslow
        # resolve against carry-in (in the kernel this happens after all runs finished)
        for (pos, hh, ff) in unres:
            if carry_valid[hh] and carry_f[hh] == ff:
                flags[a + pos] = 1
        tot_unres += len(unres)
        carry_f = np.where(touched, tab, carry_f)
        carry_valid = carry_valid | touched
    return flags, tot_unres, tot_slow


def main():
    d = np.frombuffer(open('/root/reference/benches/data/dickens.txt', 'rb').read(), np.uint8)
    rng = np.random.default_rng(1)
    cases = {
        'dickens256k': d[:262144],
        'zeros': np.zeros(65536, np.uint8),
        'random': rng.integers(0, 256, 131072, dtype=np.uint8),
        'lowentropy': rng.integers(0, 3, 131072, dtype=np.uint8),
        'mixed': np.concatenate([d[:50000 - 50000 % 4], rng.integers(0, 256, 20000, dtype=np.uint8), np.zeros(8000, np.uint8), d[70000:130000]]),
        'period': np.tile(np.arange(0, 4096, dtype=np.uint32).view(np.uint8), 4),
    }
    ok = True
    for name, data in cases.items():
        q = data[:data.size - data.size % 4].view(np.uint32)
        ref = seq_flags(q)
        for (nruns, T, S) in [(1, 4096, 8192), (3, 4096, 8192), (5, 256, 64), (2, 1024, 1024)]:
            fl, nu, ns = encode_flags(q, nruns, T, S, seed=nruns * 7 + T)
            good = bool((fl == ref).all())
            ok &= good
            print(f"{name:12s} runs={nruns} T={T:5d} S={S:5d} quads={q.size:7d} hits={int(ref.sum()):7d} unres={nu:6d} slow={ns:6d} ({100.0*ns/q.size:5.1f}%) {'OK' if good else 'MISMATCH'}")
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
