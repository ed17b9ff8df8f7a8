"""tools/proto_cheetah_decode_keyiter.py at scale (numpy): how many global context-iteration rounds does a Cheetah decode need, and how
sparse are the later rounds? Usage: python tools/proto_cheetah_decode_keyiter_np.py <MiB of synthetic text | path> [MiB]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.proto_cheetah_decode_jacobi import true_flags  # noqa: E402

M = np.uint64(0x9D6EF916)


def hash32(v):
    return ((v.astype(np.uint64) * M) & np.uint64(0xFFFFFFFF)).astype(np.int64) >> 16


def main():
    arg = sys.argv[1] if len(sys.argv) > 1 else "8"
    if arg.isdigit():
        import torch  # noqa: F401
        from density_b200 import synth
        data = synth.synth_text(int(arg) << 20).numpy()
    else:
        data = np.fromfile(arg, np.uint8)
        if len(sys.argv) > 2:
            data = data[:int(sys.argv[2]) << 20]
    q = data[:data.size // 4 * 4].view(np.uint32)
    t0 = time.time()
    flags, h = true_flags(q)
    flags = np.array(flags, np.int8); h = np.array(h, np.int64)
    n = q.size
    print(f"{n} quads, flags in {time.time() - t0:.0f} s; predicted {100 * (flags == 3).mean():.1f} %", flush=True)
    pred = flags == 3
    true_ctx = np.concatenate([[0], h[:-1]])
    ctx = np.where(np.concatenate([[False], pred[:-1]]), -1, true_ctx)       # unknown where the previous quad is predicted
    val = q.astype(np.int64)
    pos = np.arange(n)
    for rnd in range(1, 80):
        act = ctx >= 0
        idx = pos[act]
        order = np.argsort(ctx[idx], kind="stable")                           # by key, positions ascending inside a key
        si = idx[order]; sk = ctx[si]
        gstart_flag = np.concatenate([[True], sk[1:] != sk[:-1]])
        gstart = np.maximum.accumulate(np.where(gstart_flag, np.arange(si.size), 0))
        is_w = ~pred[si]
        lastw = np.maximum.accumulate(np.where(is_w, np.arange(si.size), -1))
        has = lastw >= gstart
        src_val = np.where(has, val[si[np.maximum(lastw, 0)]], 0)
        newH = np.full(n, -1, np.int64)
        rd = pred[si]
        newH[si[rd]] = hash32(src_val[rd].astype(np.uint32))
        nxt = np.nonzero(pred[:-1])[0] + 1                                     # quads whose context is an estimate
        new_ctx = ctx.copy()
        new_ctx[nxt] = newH[nxt - 1]
        changed = int((new_ctx != ctx).sum())
        ctx = new_ctx
        wrong = int((ctx != true_ctx).sum())
        print(f"round {rnd:2d}: contexts changed {changed:9d}  wrong {wrong:9d}", flush=True)
        if changed == 0:
            break


if __name__ == "__main__":
    main()
