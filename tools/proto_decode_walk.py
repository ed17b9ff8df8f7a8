"""Python model of the Chameleon decode boundary machinery for streams WITH copy-mode blocks (chameleon_decode.cu: dec_chunk_walk,
dec_group_compose, dec_seq_walk with chunk / group jumps, dec_chunk_entries, dec_block_offsets), checked against a plain in-order
parse (codec.rs:88-100 + protection_state.rs). Mirrors the kernels' conditions one to one; small CH / GROUP so that a few hundred KiB of
stream exercise every branch."""
import numpy as np

TERM = 0xFF
G_SKIP = 0xFE
COPY = 1 << 63


def popc8(s, off):
    return int(np.unpackbits(s[off:off + 8]).sum())


class Prot:
    def __init__(self):
        self.penalty, self.start, self.prev, self.counter = 0, 1, 0, 0

    def revert_to_copy(self):
        if (self.counter & 0xF) == 0 and self.start > 1:
            self.start >>= 1
        self.counter += 1
        return self.penalty > 0

    def decay(self):
        self.penalty = (self.penalty - 1) & 0xFF
        if self.penalty == 0:
            self.start = (self.start + 1) & 0xFF

    def update(self, inc):
        if inc:
            if self.prev:
                self.penalty = self.start
            self.prev = 1
        else:
            self.prev = 0

    def jump(self, nb, last_inc):                      # sw_jump()
        k = (self.counter + nb + 15) // 16 - (self.counter + 15) // 16
        if self.start > 1:
            self.start = max(1, self.start >> min(k, 8))
        self.counter += nb
        self.prev = last_inc


def reference_parse(s):
    """in-order main loop: list of (offset | COPY) per block, tail offset"""
    n = s.size
    ps = Prot(); idx = 0; out = []
    while n - idx >= 264:
        if ps.revert_to_copy():
            out.append(idx | COPY); idx += 256; ps.decay()
        else:
            consumed = 264 - 2 * popc8(s, idx)
            out.append(idx); idx += consumed; ps.update(consumed >= 256)
    return out, idx, (ps.penalty, ps.start, ps.prev)


def model_parse(s, CH=2048, GROUP=4):
    n = s.size
    NC = 132
    nchunks = (n + CH - 1) // CH
    ngroups = (nchunks + GROUP - 1) // GROUP
    # dec_chunk_walk
    res = np.zeros((nchunks, NC), np.uint32)
    for c in range(nchunks):
        base = c * CH
        for cand in range(NC):
            off = cand * 2; nb = 0; ex = TERM; term = 0; pair = first = prev = 0
            while True:
                if off >= CH:
                    ex = (off - CH) >> 1; break
                if base + off + 264 > n:
                    term = off; break
                hits = popc8(s, base + off)
                inc = 1 if hits <= 4 else 0
                if nb == 0:
                    first = inc
                pair |= inc & prev
                prev = inc
                off += 264 - 2 * hits
                nb += 1
            res[c, cand] = ex | (nb << 8) | ((term if ex == TERM else (pair | (first << 1) | (prev << 2))) << 16)
    # dec_group_compose
    gres = np.zeros((ngroups, NC, 4), np.int64)
    for g in range(ngroups):
        c0, c1 = g * GROUP, min(nchunks, (g + 1) * GROUP)
        for cand in range(NC):
            idx = cand; blocks = 0; z = w = 0; pair = first = last = have = 0
            for c in range(c0, c1):
                r = int(res[c, idx]); nb = (r >> 8) & 0xFF
                blocks += nb; idx = r & 0xFF
                if idx == TERM:
                    z, w = c, r >> 16; break
                if nb:
                    fl = r >> 16
                    if not have:
                        first = (fl >> 1) & 1; have = 1
                    else:
                        pair |= last & (fl >> 1) & 1
                    pair |= fl & 1
                    last = (fl >> 2) & 1
            if idx != TERM:
                z = pair | (first << 1) | (last << 2) | (8 if c1 - c0 < GROUP else 0)
            gres[g, cand] = (idx, blocks, z, w)
    # dec_seq_walk
    ps = Prot(); idx = 0; b = 0; g_next = 0
    g_entry = [None] * ngroups; g_bb = [0] * ngroups
    c_entry = [None] * nchunks; c_bb = [0] * nchunks
    blk = {}
    while n - idx >= 264:
        c = idx // CH
        e = (idx - c * CH) >> 1
        assert e < NC
        if c // GROUP == g_next:
            g = g_next; g_next += 1
            gx, gy, gz, _ = (int(v) for v in gres[g, e])
            if ps.penalty == 0 and gx != TERM and not (gz & 9) and not (ps.prev and (gz & 2)):
                g_entry[g] = e; g_bb[g] = b
                ps.jump(gy, (gz >> 2) & 1); b += gy
                idx = (g + 1) * GROUP * CH + 2 * gx
                continue
            g_entry[g] = G_SKIP
        r = int(res[c, e]); ex = r & 0xFF; fl = r >> 16
        if ps.penalty == 0 and ex != TERM and not (fl & 1) and not (ps.prev and (fl & 2)):
            nb = (r >> 8) & 0xFF
            c_entry[c] = e; c_bb[c] = b
            ps.jump(nb, (fl >> 2) & 1); b += nb
            idx = (c + 1) * CH + 2 * ex
        else:
            c_entry[c] = TERM
            wend = (c + 1) * CH
            while idx < wend and n - idx >= 264:
                if ps.revert_to_copy():
                    blk[b] = idx | COPY; b += 1; idx += 256; ps.decay()
                else:
                    consumed = 264 - 2 * popc8(s, idx)
                    blk[b] = idx; b += 1; idx += consumed; ps.update(consumed >= 256)
    for c in range(idx // CH, nchunks):
        c_entry[c] = TERM
    for g in range(g_next, ngroups):
        g_entry[g] = G_SKIP
    # dec_chunk_entries (jumped groups)
    for g in range(ngroups):
        i = g_entry[g]
        if i == G_SKIP:
            continue
        blocks = g_bb[g]
        for c in range(g * GROUP, min(nchunks, (g + 1) * GROUP)):
            c_entry[c] = i; c_bb[c] = blocks
            if i == TERM:
                continue
            r = int(res[c, i]); blocks += (r >> 8) & 0xFF; i = r & 0xFF
    # dec_block_offsets (jumped chunks)
    for c in range(nchunks):
        e = c_entry[c]
        if e == TERM or e is None:
            continue
        base = c * CH; bb = c_bb[c]; off = e * 2
        while off < CH and base + off + 264 <= n:
            blk[bb] = base + off; bb += 1
            off += 264 - 2 * popc8(s, base + off)
    return [blk[k] for k in range(b)], idx, (ps.penalty, ps.start, ps.prev)
