/*
 * density_oracle.c — CPU restatement of g1mv/density's Chameleon / Cheetah / Lion
 * encode/decode path (density-rs 0.16.6).
 *
 * THIS FILE IS TEST INFRASTRUCTURE. It is the parity checker for the CUDA path
 * (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference
 * leg). Nothing in density_b200/ may import, link or call it: the product path
 * fails loudly when the CUDA library is missing.
 *
 * Parity pinning: the reference is Rust and no Rust toolchain exists in the build
 * image, so oracle/_ref cannot be built ("reference unbuildable here"). The oracle
 * is pinned on the reference's own known-answer vectors (src/lib.rs:19,28,50,72),
 * on its round-trip property (benches/density.rs:42-45) and on its published
 * dickens ratios 1.749x/1.860x/1.966x (benchmark.log:17,22,27) — see
 * tests/test_oracle_golden.py.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/src). The code is a restatement, not a transcription: the
 * reference's ReadBuffer/WriteBuffer/Signature objects are folded into plain
 * cursors and the three algorithms share one block driver parameterised by a
 * small vtable.
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define HASH_MULT 0x9D6EF916u /* algorithms/chameleon/chameleon.rs:15, cheetah.rs:15, lion.rs:15 */

static inline uint32_t hash16(uint32_t quad) {
    /* chameleon.rs:89 = cheetah.rs:124 = lion.rs:212: (quad * M) >> (32 - 16) */
    return (uint32_t)(quad * HASH_MULT) >> 16;
}
static inline uint32_t ld32(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16 | (uint32_t)p[3] << 24; }
static inline uint32_t ld16(const uint8_t *p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8; }
static inline void st32(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }
static inline void st16(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }

/* ---- codec/protection_state.rs:1-47 -------------------------------------------------- */
typedef struct {
    uint8_t copy_penalty;        /* :2 */
    uint8_t copy_penalty_start;  /* :3 */
    int previous_incompressible; /* :4 */
    uint64_t counter;            /* :5 */
} protection_t;

static void prot_init(protection_t *p) { /* :9-16 */
    p->copy_penalty = 0;
    p->copy_penalty_start = 1;
    p->previous_incompressible = 0;
    p->counter = 0;
}
static int prot_revert_to_copy(protection_t *p) { /* :18-27 */
    if ((p->counter & 0xf) == 0 && p->copy_penalty_start > 1) p->copy_penalty_start >>= 1;
    p->counter++;
    return p->copy_penalty > 0;
}
static void prot_decay(protection_t *p) { /* :29-35 */
    p->copy_penalty--;
    if (p->copy_penalty == 0) p->copy_penalty_start++; /* u8; reference would panic on 255+1 in debug, wraps in release */
}
static void prot_update(protection_t *p, int incompressible) { /* :37-47 */
    if (incompressible) {
        if (p->previous_incompressible) p->copy_penalty = p->copy_penalty_start;
        p->previous_incompressible = 1;
    } else {
        p->previous_incompressible = 0;
    }
}

/* ---- per-algorithm state -------------------------------------------------------------- */
enum { ALG_CHAMELEON = 0, ALG_CHEETAH = 1, ALG_LION = 2 };

typedef struct {
    int alg;
    unsigned flag_bits;   /* chameleon.rs:17, cheetah.rs:18, lion.rs:18 */
    size_t block_size;    /* chameleon.rs:140 (256), cheetah.rs:190 (128), lion.rs:319 (64) */
    size_t unit_size;     /* chameleon.rs:143 (8), cheetah.rs:193 (4), lion.rs:322 (4) */
    size_t sig_bytes;     /* chameleon.rs:146 (8), cheetah.rs:196 (8), lion.rs:325 (6) */
    uint32_t last_hash;   /* cheetah.rs:26, lion.rs:30 */
    uint32_t *chunk_a;    /* chameleon: chunk_map (chameleon.rs:31); cheetah/lion: chunk_a (cheetah.rs:33, lion.rs:37) */
    uint32_t *chunk_b;    /* cheetah.rs:34, lion.rs:38 */
    uint32_t *pred;       /* cheetah: next (cheetah.rs:39) [1 per hash]; lion: next_a..e (lion.rs:43-47) [5 per hash] */
} codec_t;

static int codec_init(codec_t *c, int alg) {
    memset(c, 0, sizeof *c);
    c->alg = alg;
    switch (alg) {
    case ALG_CHAMELEON: c->flag_bits = 1; c->block_size = 256; c->unit_size = 8; c->sig_bytes = 8; break;
    case ALG_CHEETAH:   c->flag_bits = 2; c->block_size = 128; c->unit_size = 4; c->sig_bytes = 8; break;
    case ALG_LION:      c->flag_bits = 3; c->block_size = 64;  c->unit_size = 4; c->sig_bytes = 6; break;
    default: return -1;
    }
    /* X::new(): zero-initialised tables (chameleon.rs:39-43, cheetah.rs:47-55, lion.rs:64-72) */
    c->chunk_a = (uint32_t *)calloc(1u << 16, sizeof(uint32_t));
    if (!c->chunk_a) return -1;
    if (alg != ALG_CHAMELEON) {
        c->chunk_b = (uint32_t *)calloc(1u << 16, sizeof(uint32_t));
        c->pred = (uint32_t *)calloc((size_t)(alg == ALG_LION ? 5 : 1) << 16, sizeof(uint32_t));
        if (!c->chunk_b || !c->pred) return -1;
    }
    return 0;
}
static void codec_free(codec_t *c) { free(c->chunk_a); free(c->chunk_b); free(c->pred); }

/* ---- output / signature cursors (io/write_buffer.rs, io/write_signature.rs) ----------- */
typedef struct {
    uint8_t *buf; size_t cap; size_t idx; int overflow; int unchecked;
    size_t sig_pos; uint64_t sig_val; unsigned sig_shift; /* write_signature.rs:2-6 */
} wr_t;

static inline int wr_room(wr_t *w, size_t n) {
    /* The reference panics on an undersized slice (write_buffer.rs:19); the oracle reports it.
       When the caller's buffer is >= safe_encode_buffer_size the checks are skipped (same work
       per quad as the reference's release build, so the CPU baseline timing is fair). */
    if (w->unchecked) return 1;
    if (w->idx + n > w->cap) { w->overflow = 1; return 0; }
    return 1;
}
static inline void wr_push(wr_t *w, const uint8_t *p, size_t n) { /* write_buffer.rs:29-31 */
    if (!wr_room(w, n)) return;
    memcpy(w->buf + w->idx, p, n);
    w->idx += n;
}
static inline void wr_push32(wr_t *w, uint32_t v) { if (wr_room(w, 4)) { st32(w->buf + w->idx, v); w->idx += 4; } }
static inline void wr_push16(wr_t *w, uint32_t v) { if (wr_room(w, 2)) { st16(w->buf + w->idx, v); w->idx += 2; } }
static inline void sig_push(wr_t *w, uint64_t flag, unsigned bits) { /* write_signature.rs:13-16 */
    w->sig_val |= flag << w->sig_shift;
    w->sig_shift += bits;
}

/* ---- encode_quad ---------------------------------------------------------------------- */
static inline __attribute__((always_inline)) void chameleon_encode_quad(codec_t *c, uint32_t quad, wr_t *w) { /* chameleon.rs:86-101 */
    uint32_t h = hash16(quad);
    if (c->chunk_a[h] != quad) {
        sig_push(w, 0, 1);   /* PLAIN_FLAG, algorithms.rs:5 */
        wr_push32(w, quad);
        c->chunk_a[h] = quad;
    } else {
        sig_push(w, 1, 1);   /* MAP_FLAG, chameleon.rs:18 */
        wr_push16(w, h);
    }
}

static inline __attribute__((always_inline)) void cheetah_encode_quad(codec_t *c, uint32_t quad, wr_t *w) { /* cheetah.rs:121-150 */
    uint32_t h = hash16(quad);
    uint32_t *predicted = &c->pred[c->last_hash];
    if (*predicted != quad) {
        uint32_t a = c->chunk_a[h];
        if (a != quad) {
            if (c->chunk_b[h] != quad) { sig_push(w, 0, 2); wr_push32(w, quad); } /* plain */
            else                       { sig_push(w, 2, 2); wr_push16(w, h); }    /* MAP_B, cheetah.rs:20 */
            c->chunk_b[h] = a;
            c->chunk_a[h] = quad;
        } else {
            sig_push(w, 1, 2); wr_push16(w, h);                                   /* MAP_A, cheetah.rs:19 */
        }
        *predicted = quad;
    } else {
        sig_push(w, 3, 2);                                                        /* PREDICTED, cheetah.rs:21 */
    }
    c->last_hash = h;
}

static inline void lion_shift(uint32_t *p, uint32_t quad) { /* lion.rs:50-57 */
    p[4] = p[3]; p[3] = p[2]; p[2] = p[1]; p[1] = p[0]; p[0] = quad;
}
static inline __attribute__((always_inline)) void lion_encode_quad(codec_t *c, uint32_t quad, wr_t *w) { /* lion.rs:209-271 */
    uint32_t h = hash16(quad);
    uint32_t *p = &c->pred[(size_t)c->last_hash * 5];
    int k;
    for (k = 0; k < 5; k++) if (p[k] == quad) break;
    if (k < 5) {
        /* PREDICTED_A..E = 1..5 (lion.rs:19-23); rotate entries [0..k] so quad is first.
           k==0: nothing; k==4 is a full shift (lion.rs:240-243). */
        sig_push(w, (uint64_t)(k + 1), 3);
        for (int j = k; j > 0; j--) p[j] = p[j - 1];
        p[0] = quad;
    } else {
        uint32_t a = c->chunk_a[h];
        if (a != quad) {
            if (c->chunk_b[h] != quad) { sig_push(w, 0, 3); wr_push32(w, quad); } /* plain */
            else                       { sig_push(w, 7, 3); wr_push16(w, h); }    /* MAP_B, lion.rs:25 */
            c->chunk_b[h] = a;
            c->chunk_a[h] = quad;
        } else {
            sig_push(w, 6, 3); wr_push16(w, h);                                   /* MAP_A, lion.rs:24 */
        }
        lion_shift(p, quad);
    }
    c->last_hash = h; /* lion.rs:270 */
}

static inline __attribute__((always_inline)) void encode_quad(codec_t *c, const int alg, uint32_t quad, wr_t *w) {
    switch (alg) {
    case ALG_CHAMELEON: chameleon_encode_quad(c, quad, w); break;
    case ALG_CHEETAH:   cheetah_encode_quad(c, quad, w); break;
    default:            lion_encode_quad(c, quad, w); break;
    }
}

/* ---- codec/codec.rs:34-70 encode_block, :72-80 encode ---------------------------------- */
static size_t safe_size(int alg, size_t size);
static inline __attribute__((always_inline)) size_t codec_encode_impl(codec_t *c, const int alg, const uint8_t *in, size_t n, uint8_t *out, size_t cap,
                           uint64_t *copied_blocks) {
    wr_t w; memset(&w, 0, sizeof w); w.buf = out; w.cap = cap; w.unchecked = cap >= safe_size(alg, n);
    protection_t ps; prot_init(&ps);
    uint64_t copied = 0;
    for (size_t off = 0; off < n; off += c->block_size) {          /* input.chunks(block_size), :76 */
        size_t blen = n - off < c->block_size ? n - off : c->block_size;
        const uint8_t *blk = in + off;
        if (prot_revert_to_copy(&ps)) {                               /* :35-37 */
            wr_push(&w, blk, blen);
            prot_decay(&ps);
            copied++;
        } else {
            size_t mark = w.idx;                                      /* :39 */
            w.sig_pos = w.idx; w.sig_val = 0; w.sig_shift = 0;        /* signature.init, :40 */
            if (!wr_room(&w, c->sig_bytes)) break;
            w.idx += c->sig_bytes;                                    /* skip, :41 */
            size_t q = 0;
            for (; q + 4 <= blen; q += 4) encode_quad(c, alg, ld32(blk + q), &w); /* :42-49 and :53-57 collapse to this */
            if (q < blen) wr_push(&w, blk + q, blen - q);             /* 1..3 raw tail bytes, implicit plain flag, :58-61 */
            if (w.overflow) break;
            /* write_signature: ink 8 LE bytes at sig_pos (codec.rs:24-26, write_buffer.rs:24-26);
               Lion overrides to 6 bytes (lion.rs:333-336). With ink the 8-byte store may run past
               the 8 reserved bytes only never: sig_bytes==8 there. */
            for (size_t i = 0; i < c->sig_bytes; i++) w.buf[w.sig_pos + i] = (uint8_t)(w.sig_val >> (8 * i));
            prot_update(&ps, w.idx - mark >= c->block_size);         /* :68 */
        }
        if (w.overflow) break;
    }
    if (copied_blocks) *copied_blocks = copied;
    return w.overflow ? 0 : w.idx;
}

static size_t codec_encode(codec_t *c, const uint8_t *in, size_t n, uint8_t *out, size_t cap, uint64_t *copied_blocks) {
    switch (c->alg) {
    case ALG_CHAMELEON: return codec_encode_impl(c, ALG_CHAMELEON, in, n, out, cap, copied_blocks);
    case ALG_CHEETAH:   return codec_encode_impl(c, ALG_CHEETAH, in, n, out, cap, copied_blocks);
    default:            return codec_encode_impl(c, ALG_LION, in, n, out, cap, copied_blocks);
    }
}

/* ---- decode --------------------------------------------------------------------------- */
typedef struct { const uint8_t *buf; size_t len; size_t idx; int bad; } rd_t;
static inline size_t rd_remaining(const rd_t *r) { return r->len - r->idx; }

static inline uint32_t rd32(rd_t *r) { if (rd_remaining(r) < 4) { r->bad = 1; return 0; } uint32_t v = ld32(r->buf + r->idx); r->idx += 4; return v; }
static inline uint32_t rd16(rd_t *r) { if (rd_remaining(r) < 2) { r->bad = 1; return 0; } uint32_t v = ld16(r->buf + r->idx); r->idx += 2; return v; }

/* codec.rs:29-31 (u64 LE) and the Lion override lion.rs:338-351 (6 significant bytes) */
static uint64_t read_signature(codec_t *c, rd_t *r) {
    uint64_t v = 0;
    size_t n = c->sig_bytes;
    if (rd_remaining(r) < n) { r->bad = 1; return 0; }
    for (size_t i = 0; i < n; i++) v |= (uint64_t)r->buf[r->idx + i] << (8 * i);
    r->idx += n;
    return v;
}

/* One quad, flag already extracted. Returns 0 normally, 1 when the stream ended (partial unit). */
static inline uint32_t chameleon_dec_plain(codec_t *c, rd_t *r) { /* chameleon.rs:55-61 */
    uint32_t q = rd32(r); c->chunk_a[hash16(q)] = q; return q;
}
static inline uint32_t chameleon_dec_map(codec_t *c, rd_t *r) { /* chameleon.rs:63-68 */
    return c->chunk_a[rd16(r) & 0xffff];
}

/* cheetah.rs:67-103 / lion.rs:84-186: returns quad, sets *hash */
static inline __attribute__((always_inline)) uint32_t cl_dec_plain(codec_t *c, const int alg, rd_t *r, uint32_t *hash) {
    uint32_t q = rd32(r); uint32_t h = hash16(q);
    c->chunk_b[h] = c->chunk_a[h]; c->chunk_a[h] = q;
    if (alg == ALG_CHEETAH) c->pred[c->last_hash] = q; else lion_shift(&c->pred[(size_t)c->last_hash * 5], q);
    *hash = h; return q;
}
static inline __attribute__((always_inline)) uint32_t cl_dec_map_a(codec_t *c, const int alg, rd_t *r, uint32_t *hash) {
    uint32_t h = rd16(r) & 0xffff; uint32_t q = c->chunk_a[h];
    if (alg == ALG_CHEETAH) c->pred[c->last_hash] = q; else lion_shift(&c->pred[(size_t)c->last_hash * 5], q);
    *hash = h; return q;
}
static inline __attribute__((always_inline)) uint32_t cl_dec_map_b(codec_t *c, const int alg, rd_t *r, uint32_t *hash) {
    uint32_t h = rd16(r) & 0xffff; uint32_t q = c->chunk_b[h];
    c->chunk_b[h] = c->chunk_a[h]; c->chunk_a[h] = q;
    if (alg == ALG_CHEETAH) c->pred[c->last_hash] = q; else lion_shift(&c->pred[(size_t)c->last_hash * 5], q);
    *hash = h; return q;
}
static inline uint32_t cheetah_dec_predicted(codec_t *c, uint32_t *hash) { /* cheetah.rs:98-103 */
    uint32_t q = c->pred[c->last_hash]; *hash = hash16(q); return q;
}
static inline uint32_t lion_dec_predicted(codec_t *c, int k, uint32_t *hash) { /* lion.rs:119-186 */
    uint32_t *p = &c->pred[(size_t)c->last_hash * 5];
    uint32_t q = p[k];
    for (int j = k; j > 0; j--) p[j] = p[j - 1];
    p[0] = q;
    *hash = hash16(q); return q;
}

typedef struct { uint8_t *buf; size_t cap; size_t idx; int overflow; } ow_t;
static inline void ow_push32(ow_t *o, uint32_t v) { if (o->idx + 4 > o->cap) { o->overflow = 1; return; } st32(o->buf + o->idx, v); o->idx += 4; }
static inline void ow_push(ow_t *o, const uint8_t *p, size_t n) { if (o->idx + n > o->cap) { o->overflow = 1; return; } memcpy(o->buf + o->idx, p, n); o->idx += n; }

/* Decode ONE quad (flag_bits wide flag from *sig). `checked` selects decode_partial_unit semantics
   (chameleon.rs:116-135, cheetah.rs:165-185, lion.rs:291-314): a PLAIN flag with <4 bytes left ends the stream. */
static inline __attribute__((always_inline)) int decode_one(codec_t *c, const int alg, rd_t *r, uint64_t *sig, ow_t *o, int checked) {
    const unsigned fb = alg == ALG_CHAMELEON ? 1 : alg == ALG_CHEETAH ? 2 : 3;
    uint64_t flag = *sig & ((1u << fb) - 1);
    *sig >>= fb;
    uint32_t q, h = 0;
    if (checked && flag == 0) {
        size_t rem = rd_remaining(r);
        if (rem == 0) return 1;
        if (rem < 4) { ow_push(o, r->buf + r->idx, rem); r->idx += rem; return 1; }
    }
    if (alg == ALG_CHAMELEON) {
        q = flag ? chameleon_dec_map(c, r) : chameleon_dec_plain(c, r);
    } else if (alg == ALG_CHEETAH) {
        switch (flag) { /* cheetah.rs:154-159 */
        case 0: q = cl_dec_plain(c, alg, r, &h); break;
        case 1: q = cl_dec_map_a(c, alg, r, &h); break;
        case 2: q = cl_dec_map_b(c, alg, r, &h); break;
        default: q = cheetah_dec_predicted(c, &h); break;
        }
        c->last_hash = h;
    } else {
        switch (flag) { /* lion.rs:275-284 */
        case 0: q = cl_dec_plain(c, alg, r, &h); break;
        case 6: q = cl_dec_map_a(c, alg, r, &h); break;
        case 7: q = cl_dec_map_b(c, alg, r, &h); break;
        default: q = lion_dec_predicted(c, (int)flag - 1, &h); break;
        }
        c->last_hash = h;
    }
    ow_push32(o, q);
    return 0;
}

/* codec.rs:82-126 */
static inline __attribute__((always_inline)) size_t codec_decode_impl(codec_t *c, const int alg, const uint8_t *in, size_t n, uint8_t *out, size_t cap) {
    rd_t r = { in, n, 0, 0 };
    ow_t o = { out, cap, 0, 0 };
    protection_t ps; prot_init(&ps);
    size_t quads_per_block = c->block_size / 4;

    /* main loop, no per-unit bounds checks (:88-100) */
    while (rd_remaining(&r) >= c->sig_bytes + c->block_size) {
        if (prot_revert_to_copy(&ps)) {
            ow_push(&o, r.buf + r.idx, c->block_size); r.idx += c->block_size;
            prot_decay(&ps);
        } else {
            size_t mark = r.idx;
            uint64_t sig = read_signature(c, &r);
            /* iterations * (unit_size/4) quads. Chameleon's decode_unit consumes 2 flags per unit
               (chameleon.rs:103-114) — identical to two single-quad steps. */
            for (size_t i = 0; i < quads_per_block; i++) decode_one(c, alg, &r, &sig, &o, 0);
            prot_update(&ps, r.idx - mark >= c->block_size);
        }
        if (r.bad || o.overflow) return 0;
    }
    /* tail loop (:102-123) */
    while (rd_remaining(&r) > 0) {
        if (prot_revert_to_copy(&ps)) {
            if (rd_remaining(&r) > c->block_size) {
                ow_push(&o, r.buf + r.idx, c->block_size); r.idx += c->block_size;
            } else {
                size_t rem = rd_remaining(&r);
                ow_push(&o, r.buf + r.idx, rem); r.idx += rem;
                break;
            }
            prot_decay(&ps);
        } else {
            size_t mark = r.idx;
            uint64_t sig = read_signature(c, &r);
            int end = 0;
            size_t units = c->block_size / c->unit_size;
            size_t qpu = c->unit_size / 4;
            for (size_t u = 0; u < units && !end; u++) {
                if (rd_remaining(&r) >= c->unit_size) {
                    for (size_t k = 0; k < qpu; k++) decode_one(c, alg, &r, &sig, &o, 0);   /* decode_unit */
                } else {
                    for (size_t k = 0; k < qpu && !end; k++) end = decode_one(c, alg, &r, &sig, &o, 1); /* decode_partial_unit */
                }
                if (r.bad || o.overflow) return 0;
            }
            if (end) break;
            prot_update(&ps, r.idx - mark >= c->block_size);
        }
        if (r.bad || o.overflow) return 0;
    }
    return o.overflow ? 0 : o.idx;
}

static size_t codec_decode(codec_t *c, const uint8_t *in, size_t n, uint8_t *out, size_t cap) {
    switch (c->alg) {
    case ALG_CHAMELEON: return codec_decode_impl(c, ALG_CHAMELEON, in, n, out, cap);
    case ALG_CHEETAH:   return codec_decode_impl(c, ALG_CHEETAH, in, n, out, cap);
    default:            return codec_decode_impl(c, ALG_LION, in, n, out, cap);
    }
}

/* ---- exported C ABI, same shapes as chameleon.rs:70-83, cheetah.rs:105-118, lion.rs:193-206 ---- */
static size_t safe_size(int alg, size_t size) { /* codec.rs:18-21 */
    size_t B = alg == ALG_CHAMELEON ? 256 : alg == ALG_CHEETAH ? 128 : 64;
    size_t S = alg == ALG_LION ? 6 : 8;
    return size + (size / B) * S + (size % B ? S : 0);
}

size_t oracle_encode(int alg, const uint8_t *in, size_t n, uint8_t *out, size_t cap) {
    codec_t c; if (codec_init(&c, alg)) { codec_free(&c); return 0; }
    size_t r = codec_encode(&c, in, n, out, cap, NULL);
    codec_free(&c); return r;
}
size_t oracle_decode(int alg, const uint8_t *in, size_t n, uint8_t *out, size_t cap) {
    codec_t c; if (codec_init(&c, alg)) { codec_free(&c); return 0; }
    size_t r = codec_decode(&c, in, n, out, cap);
    codec_free(&c); return r;
}
/* ---- Codec INSTANCE reuse (codec.rs:16,72,82: `encode` / `decode` are &mut self methods; the inherent X::encode builds a fresh
   instance per call, a caller that keeps one gets a dictionary that survives from call to call until clear_state();
   ProtectionState is created inside every call, codec.rs:75,85) ------------------------------------------------------------------ */
void *oracle_codec_new(int alg) {
    codec_t *c = (codec_t *)malloc(sizeof(codec_t));
    if (!c) return NULL;
    if (codec_init(c, alg)) { codec_free(c); free(c); return NULL; }
    return c;
}
void oracle_codec_free(void *h) { if (h) { codec_free((codec_t *)h); free(h); } }
void oracle_codec_clear_state(void *h) { /* chameleon.rs:148-150, cheetah.rs:198-202, lion.rs:327-331 */
    codec_t *c = (codec_t *)h;
    c->last_hash = 0;
    memset(c->chunk_a, 0, sizeof(uint32_t) << 16);
    if (c->chunk_b) memset(c->chunk_b, 0, sizeof(uint32_t) << 16);
    if (c->pred) memset(c->pred, 0, (sizeof(uint32_t) * (c->alg == ALG_LION ? 5 : 1)) << 16);
}
size_t oracle_codec_encode(void *h, const uint8_t *in, size_t n, uint8_t *out, size_t cap) { return codec_encode((codec_t *)h, in, n, out, cap, NULL); }
size_t oracle_codec_decode(void *h, const uint8_t *in, size_t n, uint8_t *out, size_t cap) { return codec_decode((codec_t *)h, in, n, out, cap); }

/* encode + number of copy-mode blocks (diagnostic for the tests; not in the reference ABI) */
size_t oracle_encode_stats(int alg, const uint8_t *in, size_t n, uint8_t *out, size_t cap, uint64_t *copied_blocks) {
    codec_t c; if (codec_init(&c, alg)) { codec_free(&c); return 0; }
    size_t r = codec_encode(&c, in, n, out, cap, copied_blocks);
    codec_free(&c); return r;
}

size_t oracle_chameleon_encode(const uint8_t *i, size_t n, uint8_t *o, size_t c) { return oracle_encode(ALG_CHAMELEON, i, n, o, c); }
size_t oracle_chameleon_decode(const uint8_t *i, size_t n, uint8_t *o, size_t c) { return oracle_decode(ALG_CHAMELEON, i, n, o, c); }
size_t oracle_chameleon_safe_encode_buffer_size(size_t s) { return safe_size(ALG_CHAMELEON, s); }
size_t oracle_cheetah_encode(const uint8_t *i, size_t n, uint8_t *o, size_t c) { return oracle_encode(ALG_CHEETAH, i, n, o, c); }
size_t oracle_cheetah_decode(const uint8_t *i, size_t n, uint8_t *o, size_t c) { return oracle_decode(ALG_CHEETAH, i, n, o, c); }
size_t oracle_cheetah_safe_encode_buffer_size(size_t s) { return safe_size(ALG_CHEETAH, s); }
size_t oracle_lion_encode(const uint8_t *i, size_t n, uint8_t *o, size_t c) { return oracle_encode(ALG_LION, i, n, o, c); }
size_t oracle_lion_decode(const uint8_t *i, size_t n, uint8_t *o, size_t c) { return oracle_decode(ALG_LION, i, n, o, c); }
size_t oracle_lion_safe_encode_buffer_size(size_t s) { return safe_size(ALG_LION, s); }
