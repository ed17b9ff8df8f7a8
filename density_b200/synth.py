"""Deterministic synthetic corpora for the BASELINE.json configurations.

Counter-based (splitmix64 of the byte/slot index), integer-only on the data path, written with torch ops so that the
same function produces bit-identical bytes on CPU (tests, CPU baseline sample) and on CUDA (bench at 1-64 GiB without
crossing PCIe). No dataset or RNG state is involved.

  text   64 KiB pages of pseudo-English: Zipf-distributed words from a 32,768-word synthetic vocabulary, separated by
         ' ' / ', ' / '. ' / '\\n'. Tuned so that Chameleon compresses it ~1.7x like Silesia/dickens (1.749x,
         /root/reference/benchmark.log:17).
  mixed  256 KiB regions: 4/8 text, 1/8 uniform random bytes, 1/8 zeros, 1/8 little-endian u32 counter table,
         1/8 repeating 24-byte records with two random fields.
"""
import math

import numpy as np
import torch

PAGE = 65536
SLOTS = 16384          # token slots per page (more than enough to fill 64 KiB)
VOCAB = 32768
MAXTOK = 16
TEXT_SEED = 0xD3A517E5
MIXED_SEED = 0xB10CB200
REGION = 262144

_MASK64 = (1 << 64) - 1


def _sm64_py(x):
    x = (x + 0x9E3779B97F4A7C15) & _MASK64
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _MASK64
    return z ^ (z >> 31)


def _i64(v):
    v &= _MASK64
    return v - (1 << 64) if v >= (1 << 63) else v


def _lsr(x, k):
    """logical shift right on int64 tensors"""
    return (x >> k) & ((1 << (64 - k)) - 1)


def splitmix64(x):
    """splitmix64 finaliser of (x + golden gamma), elementwise on an int64 tensor (two's complement wrap-around)."""
    x = x + _i64(0x9E3779B97F4A7C15)
    z = x
    z = (z ^ _lsr(z, 30)) * _i64(0xBF58476D1CE4E5B9)
    z = (z ^ _lsr(z, 27)) * _i64(0x94D049BB133111EB)
    return z ^ _lsr(z, 31)


_VOCAB_CACHE = {}


def _vocab_tables(seed):
    """token table: VOCAB words x 4 separator kinds -> (bytes[VOCAB*4, MAXTOK] uint8, len[VOCAB*4] int64) and the
    65536-entry Zipf lookup (uniform u16 -> word rank). Built on the host once (tiny), integer hashing only."""
    if seed in _VOCAB_CACHE:
        return _VOCAB_CACHE[seed]
    onsets = ["", "b", "c", "d", "f", "g", "h", "l", "m", "n", "p", "r", "s", "t", "w", "th", "st", "ch", "sh", "pr", "tr", "wh"]
    nuclei = ["a", "e", "i", "o", "u", "ea", "ou", "ee", "ai", "oo"]
    codas = ["", "", "n", "r", "s", "t", "d", "l", "m", "ng", "nd", "st", "rt", "ll", "ck"]
    words, seen = [], set()
    common = ["the", "of", "and", "to", "a", "in", "that", "he", "was", "it", "his", "i", "with", "as", "had", "for", "at",
              "her", "you", "not", "be", "on", "is", "said", "she", "but", "have", "him", "by", "my", "which", "all", "so",
              "this", "from", "mr", "they", "were", "me", "no", "one", "there", "would", "what", "when", "been", "if", "we"]
    for w in common:
        words.append(w); seen.add(w)
    ctr = 0
    while len(words) < VOCAB:
        r = _sm64_py(seed * 1315423911 + ctr); ctr += 1
        nsyl = 1 + (r & 3) % 3 + (1 if (r >> 2) & 7 == 0 else 0)
        w = ""
        rr = r >> 8
        for _ in range(nsyl):
            w += onsets[rr % len(onsets)]; rr //= len(onsets)
            w += nuclei[rr % len(nuclei)]; rr //= len(nuclei)
            w += codas[rr % len(codas)]; rr //= len(codas)
        w = w[:12]
        if (r >> 60) == 0:  # ~6% capitalised
            w = w.capitalize()
        if w and w not in seen:
            seen.add(w); words.append(w)
    seps = [" ", ", ", ". ", "\n"]
    tb = np.full((VOCAB * 4, MAXTOK), ord(" "), dtype=np.uint8)
    tl = np.zeros(VOCAB * 4, dtype=np.int64)
    for i, w in enumerate(words):
        for k, s in enumerate(seps):
            t = (w + s).encode()
            tb[i * 4 + k, :len(t)] = np.frombuffer(t, dtype=np.uint8)
            tl[i * 4 + k] = len(t)
    # Zipf(s~1) by inverse transform on a 16-bit uniform: rank = floor(V^u) - 1
    u = (np.arange(65536, dtype=np.float64) + 0.5) / 65536.0
    zipf = np.minimum(VOCAB - 1, np.floor(np.exp(u * math.log(VOCAB))).astype(np.int64) - 1)
    zipf = np.maximum(zipf, 0)
    out = (torch.from_numpy(tb), torch.from_numpy(tl), torch.from_numpy(zipf))
    _VOCAB_CACHE[seed] = out
    return out


def _text_pages(page_ids, seed, device):
    """page_ids: int64 tensor [P] -> uint8 tensor [P, PAGE]"""
    tb, tl, zipf = (t.to(device) for t in _vocab_tables(seed))
    P = page_ids.numel()
    slot = torch.arange(SLOTS, dtype=torch.int64, device=device)
    key = splitmix64((page_ids.to(device)[:, None] * SLOTS + slot[None, :]) ^ _i64(seed * 0x632BE59BD9B4E019))
    word = zipf[key & 0xFFFF]
    sb = _lsr(key, 16) & 0xFF
    sep = torch.zeros_like(sb)
    sep = torch.where(sb >= 228, torch.ones_like(sb), sep)         # ', '  ~6%
    sep = torch.where(sb >= 243, torch.full_like(sb, 2), sep)      # '. '  ~4%
    sep = torch.where(sb >= 253, torch.full_like(sb, 3), sep)      # '\n'  ~1%
    tok = word * 4 + sep
    ln = tl[tok]
    end = torch.cumsum(ln, dim=1)
    start = end - ln
    out = torch.full((P, PAGE), ord(" "), dtype=torch.uint8, device=device)
    flat = out.view(-1)
    rowbase = (torch.arange(P, dtype=torch.int64, device=device) * PAGE)[:, None]
    for c in range(MAXTOK):
        pos = start + c
        m = (ln > c) & (pos < PAGE)
        idx = (rowbase + pos)[m]
        flat[idx] = tb[tok[m], c]
    return out


def synth_text(nbytes, seed=TEXT_SEED, device="cpu", first_page=0, pages_per_chunk=256):
    """`nbytes` of synthetic English-like text starting at page `first_page` of the infinite corpus."""
    device = torch.device(device)
    npages = (nbytes + PAGE - 1) // PAGE
    out = torch.empty(npages * PAGE, dtype=torch.uint8, device=device)
    for p0 in range(0, npages, pages_per_chunk):
        p1 = min(npages, p0 + pages_per_chunk)
        ids = torch.arange(first_page + p0, first_page + p1, dtype=torch.int64, device=device)
        out[p0 * PAGE:p1 * PAGE] = _text_pages(ids, seed, device).view(-1)
    return out[:nbytes]


def random_bytes(nbytes, seed, device="cpu", offset_words=0):
    """uniform bytes: splitmix64 of the 8-byte word index, little-endian"""
    device = torch.device(device)
    nw = (nbytes + 7) // 8
    w = splitmix64(torch.arange(offset_words, offset_words + nw, dtype=torch.int64, device=device) ^ _i64(seed * 0x9E3779B97F4A7C15))
    return w.view(torch.uint8)[:nbytes] if w.is_contiguous() else w.contiguous().view(torch.uint8)[:nbytes]


def synth_mixed(nbytes, seed=MIXED_SEED, device="cpu", first_region=0):
    """mixed text/binary buffer of 256 KiB regions."""
    device = torch.device(device)
    nreg = (nbytes + REGION - 1) // REGION
    out = torch.empty(nreg * REGION, dtype=torch.uint8, device=device)
    for r in range(nreg):
        rid = first_region + r
        kind = _sm64_py(seed ^ (rid * 0x2545F4914F6CDD1D)) % 8
        dst = out[r * REGION:(r + 1) * REGION]
        if kind < 4:
            dst.copy_(synth_text(REGION, seed=TEXT_SEED, device=device, first_page=rid * (REGION // PAGE)))
        elif kind == 4:
            dst.copy_(random_bytes(REGION, seed + 1, device=device, offset_words=rid * (REGION // 8)))
        elif kind == 5:
            dst.zero_()
        elif kind == 6:
            ctr = torch.arange(rid * (REGION // 4), (rid + 1) * (REGION // 4), dtype=torch.int64, device=device).to(torch.int32)
            dst.copy_(ctr.view(torch.uint8))
        else:
            # repeating 24-byte records: 16 constant bytes + two random u32 fields
            nrec = REGION // 24 + 1
            rec = torch.zeros((nrec, 24), dtype=torch.uint8, device=device)
            rec[:, :16] = torch.tensor(list(b"REC0\x01\x00\x00\x00density_"), dtype=torch.uint8, device=device)
            rnd = random_bytes(nrec * 8, seed + 2, device=device, offset_words=rid * nrec).view(nrec, 8)
            rec[:, 16:24] = rnd
            dst.copy_(rec.view(-1)[:REGION])
    return out[:nbytes]
