"""End-to-end Python model of the planned parallel Cheetah DECODER, from the stream bytes to the original bytes,
copy-mode blocks and the tail loop included, checked against the inputs the oracle encoded. Every stage is written the way the kernels
are meant to work (what is "parallel" is computed from whole arrays / per-run loops without looking at earlier results of the same stage):

  0. boundaries: in-order walk of the signatures with the protection automaton (the Chameleon machinery with 2-bit flags;
     block = 8 + 4*plain + 2*map bytes, copy-mode block = 128 raw bytes; main loop while remaining >= 8 + 128, codec.rs:88-100)
  1. unpack: flags + literals per quad (cd_unpack)
  2. chunk-map values, run-parallel: transfer functions -> fold -> concrete pass (cd_cmap_*)
  3. predicted values: global iteration on the contexts (cd_pred_*), contexts skip copy-mode blocks
  4. tail (codec.rs:102-123): in order from the folded tables, literal control flow
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
from tools.proto_decode_walk import Prot  # noqa: E402

M = 0x9D6EF916
A0, B0 = ("sym", "a0"), ("sym", "b0")


def hash16(v):
    return ((v * M) & 0xFFFFFFFF) >> 16


def decode(s, out_size, nruns=5):
    n = s.size
    sb = s.tobytes()
    u16 = lambda o: sb[o] | (sb[o + 1] << 8)
    u32 = lambda o: u16(o) | (u16(o + 2) << 16)
    # ---- 0. boundaries --------------------------------------------------------------------------------------------------
    ps = Prot(); idx = 0; blocks = []
    while n - idx >= 8 + 128:
        if ps.revert_to_copy():
            blocks.append((idx, True)); idx += 128; ps.decay()
        else:
            sig = int.from_bytes(sb[idx:idx + 8], "little")
            size = 8
            for k in range(32):
                f = (sig >> (2 * k)) & 3
                size += 4 if f == 0 else 2 if f != 3 else 0
            blocks.append((idx, False)); idx += size; ps.update(size - 0 >= 128 + 0 if False else (size >= 128))
    tail_off = idx
    nb = len(blocks)
    # ---- 1. unpack -----------------------------------------------------------------------------------------------------------
    nq = nb * 32
    flag = [0] * nq; lit = [0] * nq; copied = [False] * nq; out = [0] * nq
    for b, (o, cp) in enumerate(blocks):
        if cp:
            for k in range(32):
                copied[b * 32 + k] = True; out[b * 32 + k] = u32(o + 4 * k)
            continue
        sig = int.from_bytes(sb[o:o + 8], "little"); p = o + 8
        for k in range(32):
            f = (sig >> (2 * k)) & 3
            flag[b * 32 + k] = f
            if f == 0:
                lit[b * 32 + k] = u32(p); p += 4
            elif f != 3:
                lit[b * 32 + k] = u16(p); p += 2
    enc = [i for i in range(nq) if not copied[i]]          # the encoded quads in stream order (copy-mode blocks touch nothing)
    hN = {i: (hash16(lit[i]) if flag[i] == 0 else lit[i]) for i in enc if flag[i] != 3}
    # ---- 2. chunk-map values (runs = contiguous slices of the block list) ------------------------------------------------------
    bounds = [(nb * r // nruns) * 32 for r in range(nruns)] + [nq]
    tf = []
    for r in range(nruns):
        st = {}
        for i in range(bounds[r], bounds[r + 1]):
            if copied[i] or flag[i] in (1, 3):
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
            if copied[i] or flag[i] == 3:
                continue
            a, b = cm.get(hN[i], (0, 0))
            if flag[i] == 0:
                v = lit[i]; cm[hN[i]] = (v, a)
            elif flag[i] == 1:
                v = a
            else:
                v = b; cm[hN[i]] = (b, a)
            out[i] = v
    # ---- 3. predicted values: global iteration on the contexts ---------------------------------------------------------------------
    H = {i: hN.get(i) for i in enc}                        # None = unknown (predicted)
    rounds = 0
    while True:
        rounds += 1
        last = {}; newH = {}; pv = {}
        ctx = 0                                            # last_hash starts as 0 (cheetah.rs:54)
        for i in enc:                                      # (a previous-occurrence pass under the current estimates)
            if flag[i] == 3:
                if ctx is not None:
                    v = last.get(ctx, 0); pv[i] = v; newH[i] = hash16(v)
                else:
                    newH[i] = None
            else:
                if ctx is not None:
                    last[ctx] = out[i]
            ctx = H[i]
        changed = sum(1 for i in newH if newH[i] != H[i])
        for i in newH:
            H[i] = newH[i]
        if changed == 0:
            break
    for i, v in pv.items():
        out[i] = v
    # tables after the main loop, for the tail: prediction table = latest writer per context under the final contexts
    pred = {}; ctx = 0
    for i in enc:
        if flag[i] != 3:
            pred[ctx] = out[i]
        ctx = H[i]
    last_hash = ctx
    # ---- 4. tail loop, in order (codec.rs:102-123; cheetah.rs:152-185) ------------------------------------------------------------------
    res = bytearray()
    for v in out:
        res += int(v).to_bytes(4, "little")
    cm = dict(cm_final)
    idx = tail_off
    while n - idx > 0:
        if ps.revert_to_copy():
            rem = n - idx
            if rem > 128:
                res += sb[idx:idx + 128]; idx += 128
            else:
                res += sb[idx:]; idx = n; break
            ps.decay()
        else:
            mark = idx
            sig = int.from_bytes(sb[idx:idx + 8], "little"); idx += 8
            end = False
            for k in range(32):
                f = (sig >> (2 * k)) & 3
                if (n - idx) < 4:                          # decode_partial_unit (cheetah.rs:165-185): a PLAIN flag with < 4 bytes left ends the stream
                    if f == 0:
                        res += sb[idx:]; idx = n; end = True; break
                if f == 3:
                    v = pred.get(last_hash, 0); hh = hash16(v)
                else:
                    if f == 0:
                        v = u32(idx); idx += 4; hh = hash16(v)
                        a, b = cm.get(hh, (0, 0)); cm[hh] = (v, a)
                    else:
                        hh = u16(idx); idx += 2
                        a, b = cm.get(hh, (0, 0))
                        if f == 1:
                            v = a
                        else:
                            v = b; cm[hh] = (b, a)
                    pred[last_hash] = v
                res += int(v).to_bytes(4, "little")
                last_hash = hh
            if end:
                break
            ps.update(idx - mark >= 128)
    return np.frombuffer(bytes(res[:out_size]), np.uint8), rounds, sum(1 for _, c in blocks if c), len(res)


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
        enc = oracle.encode("cheetah", data)
        for nruns in (1, 5):
            got, rounds, ncopy, produced = decode(enc, data.size, nruns)
            good = produced == data.size and got.size == data.size and bool((got == data).all())
            ok &= good
            print(f"{name:8s} runs={nruns} in={data.size:6d} stream={enc.size:6d} copy-mode blocks={ncopy:4d} context rounds={rounds:2d} {'OK' if good else 'MISMATCH'}")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
