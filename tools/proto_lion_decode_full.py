"""End-to-end Python model of the planned parallel Lion DECODER (see tools/proto_cheetah_decode_full.py): stream bytes -> original bytes with
copy-mode blocks and the tail, checked against oracle-encoded inputs. Lion specifics: 64-byte blocks of 16 quads, 3-bit flags in a
6-byte signature (lion.rs:317-351), flags 1..5 = predicted at depth k (value = the context's list entry k, entries [0..k] rotate),
6 / 7 = MAP_A / MAP_B, 0 = PLAIN; non-predicted quads are shifted into the context's list (lion.rs:84-186).

Stages: boundaries in order (block = 6 + 4*plain + 2*map bytes; main loop while remaining >= 6 + 64) -> unpack -> chunk-map values
run-parallel (as Cheetah) -> the hash chain over 5 hashes per context (serial state; here in order) -> values replayed per context
(contexts independent) -> tail in order."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
from tools.proto_decode_walk import Prot  # noqa: E402
from tools.proto_cheetah_decode_full import hash16, A0, B0  # noqa: E402


def decode(s, out_size, nruns=5):
    n = s.size
    sb = s.tobytes()
    u16 = lambda o: sb[o] | (sb[o + 1] << 8)
    u32 = lambda o: u16(o) | (u16(o + 2) << 16)
    sig48 = lambda o: int.from_bytes(sb[o:o + 6], "little")
    ps = Prot(); idx = 0; blocks = []
    while n - idx >= 6 + 64:
        if ps.revert_to_copy():
            blocks.append((idx, True)); idx += 64; ps.decay()
        else:
            sig = sig48(idx); size = 6
            for k in range(16):
                f = (sig >> (3 * k)) & 7
                size += 4 if f == 0 else 2 if f >= 6 else 0
            blocks.append((idx, False)); idx += size; ps.update(size >= 64)
    tail_off = idx
    nb = len(blocks); nq = nb * 16
    flag = [0] * nq; lit = [0] * nq; copied = [False] * nq; out = [0] * nq
    for b, (o, cp) in enumerate(blocks):
        if cp:
            for k in range(16):
                copied[b * 16 + k] = True; out[b * 16 + k] = u32(o + 4 * k)
            continue
        sig = sig48(o); p = o + 6
        for k in range(16):
            f = (sig >> (3 * k)) & 7
            flag[b * 16 + k] = f
            if f == 0:
                lit[b * 16 + k] = u32(p); p += 4
            elif f >= 6:
                lit[b * 16 + k] = u16(p); p += 2
    enc = [i for i in range(nq) if not copied[i]]
    isP = lambda f: 1 <= f <= 5
    hN = {i: (hash16(lit[i]) if flag[i] == 0 else lit[i]) for i in enc if not isP(flag[i])}
    # chunk-map values, run-parallel (transfer functions, fold, concrete pass)
    bounds = [(nb * r // nruns) * 16 for r in range(nruns)] + [nq]
    tf = []
    for r in range(nruns):
        st = {}
        for i in range(bounds[r], bounds[r + 1]):
            if copied[i] or isP(flag[i]) or flag[i] == 6:
                continue
            a, b = st.get(hN[i], (A0, B0))
            st[hN[i]] = (lit[i], a) if flag[i] == 0 else (b, a)
        tf.append(st)
    carry = {}; cin = []
    for r in range(nruns):
        cin.append(dict(carry))
        for k, (a, b) in tf[r].items():
            a0, b0 = carry.get(k, (0, 0))
            sub = lambda x: a0 if x == A0 else b0 if x == B0 else x
            carry[k] = (sub(a), sub(b))
    cm_final = carry
    for r in range(nruns):
        cm = dict(cin[r])
        for i in range(bounds[r], bounds[r + 1]):
            if copied[i] or isP(flag[i]):
                continue
            a, b = cm.get(hN[i], (0, 0))
            if flag[i] == 0:
                v = lit[i]; cm[hN[i]] = (v, a)
            elif flag[i] == 6:
                v = a
            else:
                v = b; cm[hN[i]] = (b, a)
            out[i] = v
    # hash chain (serial state: 5 hashes per context)
    T = {}; H = {}; c = 0
    for i in enc:
        p = T.setdefault(c, [0, 0, 0, 0, 0])
        f = flag[i]
        if isP(f):
            H[i] = p[f - 1]; p.insert(0, p.pop(f - 1))
        else:
            H[i] = hN[i]; p.pop(); p.insert(0, H[i])
        c = H[i]
    last_hash = c
    # values per context
    by_ctx = {}; ctx = 0
    for i in enc:
        by_ctx.setdefault(ctx, []).append(i); ctx = H[i]
    lists = {}
    for cx, idxs in by_ctx.items():
        p = [0, 0, 0, 0, 0]
        for i in idxs:
            f = flag[i]
            if isP(f):
                out[i] = p[f - 1]; p.insert(0, p.pop(f - 1))
            else:
                p.pop(); p.insert(0, out[i])
        lists[cx] = p
    # tail
    res = bytearray()
    for v in out:
        res += int(v).to_bytes(4, "little")
    cm = dict(cm_final)
    idx = tail_off
    while n - idx > 0:
        if ps.revert_to_copy():
            rem = n - idx
            if rem > 64:
                res += sb[idx:idx + 64]; idx += 64
            else:
                res += sb[idx:]; idx = n; break
            ps.decay()
        else:
            mark = idx
            if n - idx < 6:
                break
            sig = sig48(idx); idx += 6
            end = False
            for k in range(16):
                f = (sig >> (3 * k)) & 7
                if (n - idx) < 4 and f == 0:                 # decode_partial_unit (lion.rs:291-314)
                    res += sb[idx:]; idx = n; end = True; break
                p = lists.setdefault(last_hash, [0, 0, 0, 0, 0])
                if isP(f):
                    v = p[f - 1]; p.insert(0, p.pop(f - 1)); hh = hash16(v)
                else:
                    if f == 0:
                        v = u32(idx); idx += 4; hh = hash16(v)
                        a, b = cm.get(hh, (0, 0)); cm[hh] = (v, a)
                    else:
                        hh = u16(idx); idx += 2
                        a, b = cm.get(hh, (0, 0))
                        if f == 6:
                            v = a
                        else:
                            v = b; cm[hh] = (b, a)
                    p.pop(); p.insert(0, v)
                res += int(v).to_bytes(4, "little")
                last_hash = hh
            if end:
                break
            ps.update(idx - mark >= 64)
    return np.frombuffer(bytes(res[:out_size]), np.uint8), sum(1 for _, c in blocks if c), len(res)


def main():
    d = np.fromfile(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "dickens_200k.bin"), np.uint8)
    rng = np.random.default_rng(5)
    cases = {
        "kat": np.frombuffer(b"test" * 31 + b"t", np.uint8),
        "dickens": d[:50003],
        "mixed": np.concatenate([d[:20000], rng.integers(0, 256, 9000, dtype=np.uint8), np.zeros(5000, np.uint8), d[50000:70002]]),
        "random": rng.integers(0, 256, 20001, dtype=np.uint8),
        "zeros": np.zeros(30000, np.uint8),
        "low": rng.integers(0, 3, 30000, dtype=np.uint8),
    }
    ok = True
    for name, data in cases.items():
        enc = oracle.encode("lion", data)
        for nruns in (1, 5):
            got, ncopy, produced = decode(enc, data.size, nruns)
            good = produced == data.size and got.size == data.size and bool((got == data).all())
            ok &= good
            print(f"{name:8s} runs={nruns} in={data.size:6d} stream={enc.size:6d} copy-mode blocks={ncopy:4d} {'OK' if good else 'MISMATCH'}")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
