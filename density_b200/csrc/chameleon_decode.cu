// chameleon_decode.cu — parallel Chameleon decode for sm_100a.
//
// Replaces /root/reference/src/algorithms/chameleon/chameleon.rs:55-68,103-135 (decode_plain / decode_map / decode_unit /
// decode_partial_unit) driven by /root/reference/src/codec/codec.rs:82-126 (Codec::decode), bit-exactly.
//
// The reference decoder has three serial chains: block boundaries (block b+1 starts where block b's
// payload ends; nothing in the stream says where), the protection automaton, and the dictionary (a MAP quad reads the most
// recent PLAIN quad with that hash). This file breaks them as follows.
//
//  1. Boundaries.  A full encoded block is 8 + 256 - 2*popc(sig) bytes. The stream is cut into 16 KiB chunks; for each
//     chunk and each of the 132 possible (even) entry offsets in its first 264 bytes, `dec_chunk_walk` walks the chunk in
//     shared memory and records where that walk leaves the chunk and how many blocks it saw. Composing these 132-entry maps
//     (per group of 64 chunks, then over the groups, then back down) yields every chunk's true entry point and block
//     index; `dec_block_offsets` re-walks each chunk from its true entry and writes one offset per block.
//  2. Protection.  `dec_quiet_check`: if no two consecutive blocks are incompressible (>= 256 bytes consumed, codec.rs:98)
//     the automaton never leaves its initial state and no block is in copy mode (same argument as the encoder). Otherwise
//     the candidate walks are void (a copy-mode block has no signature) and `dec_seq_walk` redoes the boundaries in order
//     with the exact automaton: chunks in which the automaton provably stays in encoded mode are jumped in O(1) from the
//     candidate table, the others are walked block by block from shared memory and their copy-mode blocks marked; the
//     dictionary passes below then run unchanged (a copy-mode block is 64 raw quads that neither read nor write the
//     dictionary, codec.rs:89-92).
//  3. Dictionary.  `cham_decode_pass`: one persistent CTA per contiguous run of blocks, the run's dictionary in shared
//     memory as 16-bit fingerprints (common.cuh). Per tile of 4096 quads: PLAIN quads (~8 %) are the writers, MAP quads the
//     readers; barrier-phased optimistic protocol — A readers read / B writers publish / C readers re-read; unchanged means
//     no writer touched the bucket in this tile / D the few readers whose bucket was written resolve exactly (all writers
//     later than me -> pre-tile value; writers agree and one precedes me -> that value; else search the tile's writer list).
//     A reader whose bucket has not been written in this run yet cannot know the carried-in dictionary: it is recorded
//     and patched afterwards from the per-run carry-in tables (same machinery as the encoder).
//  4. Tail.  The last < 264 bytes of the stream (codec.rs:102-123: per-unit bounds checks, partial units, 1-3 raw bytes) are
//     decoded by one thread with the reference's literal control flow, starting from the folded dictionary.
#include "common.cuh"
#include "encode_internal.cuh"
#include "decode_bounds.cuh"

namespace dns {
namespace chamdec {

using bounds::DecStatus;
using bounds::BLK_COPY;
using bounds::ldu16;
using T = bounds::ChamT;     // boundaries: decode_bounds.cuh (shared with the Cheetah decoder)

// ---- 3. decode pass ---------------------------------------------------------------------------------------------------------
constexpr int DP_THREADS = 1024;
constexpr int DP_QPT = 4;
constexpr int TILE_Q = DP_THREADS * DP_QPT;   // 4096 quads = 64 blocks
constexpr int SIDE_N = 4096;
constexpr uint32_t SIDE_EMPTY = 0xFFFFFFFFu;

// writer record: x = hash | fp << 16, y = pos(12) | agree-flag etc.   suspect reader record: x = hash | w << 16, y = pos | touched << 12 | fa << 16
constexpr uint32_t W_CONF = 1u << 12;
constexpr uint32_t S_TOUCHED = 1u << 12;

struct DecSmem {
    uint16_t tab[65536];
    uint32_t vbit[2048];
    uint32_t conf[2048];          // per-tile: writers of the bucket disagree
    uint32_t wbit[2048];          // per-tile: bucket has a writer in this tile
    uint32_t side[SIDE_N];        // per-tile min over writers of (pos << 16 | hash)
    uint2 wrec[TILE_Q];           // writers (plain quads) of the tile from the front; suspect readers (readers whose bucket is written in
                                  // this tile) from the back: a quad is one or the other, so the two lists never meet
    unsigned long long boff[2][64];  // stream offset of each block of the tile (double buffered: next tile staged early)
    uint32_t bsig[2][128];           // signature halves
    uint32_t bcopy[2][2];            // copy-mode blocks of the tile (bit per block)
    uint32_t nw, ns;
};
static_assert(sizeof(DecSmem) <= 227 * 1024, "decode pass shared memory");

__device__ __forceinline__ bool bit_test(const uint32_t* bm, uint32_t i) { return (bm[i >> 5] >> (i & 31)) & 1u; }

// WONLY = true: "writer pass" — only the PLAIN quads are looked at; produces each run's last-writer table so that the
// carry-in dictionary of every run is known before the real decode pass starts (a MAP quad cannot tell what its bucket
// held at the start of the run). WONLY = false: the decode pass proper, dictionary preloaded from `carry`.
// Stage block offsets + signatures of the tile starting at block b0 into buffer `buf` (first 64 threads).
__device__ __forceinline__ void stage_tile(DecSmem& S, int buf, const uint8_t* __restrict__ in, const uint64_t* __restrict__ blk_off,
                                           uint64_t b0, uint64_t nblocks) {
    const uint32_t tid = threadIdx.x;
    if (tid < 64) {     // warps 0 and 1, whole warps
        unsigned long long o = 0; uint32_t lo = 0, hi = 0; bool copied = false;
        if (b0 + tid < nblocks) {
            o = blk_off[b0 + tid];
            copied = (o & BLK_COPY) != 0;
            o &= ~BLK_COPY;
            if (!copied) {
                const uint8_t* p = in + o;
                lo = ldu16(p) | (ldu16(p + 2) << 16); hi = ldu16(p + 4) | (ldu16(p + 6) << 16);
                o += 8;                       // payload starts behind the signature
            }                                 // copy-mode block: 64 raw quads = "all PLAIN" payload right at the block start
        }
        S.boff[buf][tid] = o; S.bsig[buf][2 * tid] = lo; S.bsig[buf][2 * tid + 1] = hi;
        const uint32_t cmask = __ballot_sync(0xFFFFFFFFu, copied);
        if ((tid & 31) == 0) S.bcopy[buf][tid >> 5] = cmask;
    }
}
// Issue the payload loads of my DP_QPT quads of the tile staged in `buf`: MAP -> 16-bit hash, PLAIN -> the quad.
template <bool WONLY>
__device__ __forceinline__ void fetch_payload(const DecSmem& S, int buf, const uint8_t* __restrict__ in, uint32_t nb_tile, uint32_t (&v)[DP_QPT]) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < DP_QPT; ++j) {
        const uint32_t bl = warp * 2 + (j >> 1);
        const uint32_t k = (j & 1) * 32 + lane;
        v[j] = 0;
        if (bl < nb_tile) {
            const uint32_t lo = S.bsig[buf][2 * bl], hi = S.bsig[buf][2 * bl + 1];
            const uint32_t flag = (((j & 1) ? hi : lo) >> lane) & 1u;
            const uint32_t before = (j & 1) ? (__popc(lo) + __popc(hi & lanemask_lt())) : __popc(lo & lanemask_lt());
            const uint8_t* p = in + S.boff[buf][bl] + 4 * k - 2 * before;
            if (flag) { if (!WONLY) v[j] = ldu16(p); }                       // decode_map reads the 16-bit hash (chameleon.rs:64)
            else v[j] = ldu16(p) | (ldu16(p + 2) << 16);                     // decode_plain reads the quad (chameleon.rs:56)
        }
    }
}

template <bool WONLY>
__global__ void __launch_bounds__(DP_THREADS, 1)
cham_decode_pass(const uint8_t* __restrict__ in, const uint64_t* __restrict__ blk_off, DecStatus* st,
                 uint32_t nruns, uint32_t* __restrict__ out /* quads */, const uint32_t* __restrict__ carry,
                 uint32_t* __restrict__ final_tab) {
    if (st->nonquiet || st->error) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DecSmem& S = *reinterpret_cast<DecSmem*>(smem_raw);
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t run = blockIdx.x;
    const uint64_t nblocks = st->main_blocks;
    const uint64_t ntiles = (nblocks + 63) / 64;
    const uint64_t t_begin = (uint64_t)run * ntiles / nruns, t_end = (uint64_t)(run + 1) * ntiles / nruns;

    {
        uint4 z = make_uint4(0, 0, 0, 0);
        uint4* t4 = reinterpret_cast<uint4*>(S.tab);
        for (uint32_t i = tid; i < 65536 * 2 / 16; i += DP_THREADS) t4[i] = z;
        for (uint32_t i = tid; i < 2048; i += DP_THREADS) { S.vbit[i] = 0; S.conf[i] = 0; S.wbit[i] = 0; }
        if (!WONLY) {
            __syncthreads();
            const uint32_t* __restrict__ cr = carry + (size_t)run * 65536;   // dictionary before this run
            for (uint32_t i = tid; i < 65536; i += DP_THREADS) {
                const uint32_t c = cr[i];
                if (c & 0x10000u) {
                    S.tab[i] = (uint16_t)c;
                    if ((c & 0xFFFFu) == 0) atomicOr(&S.vbit[i >> 5], 1u << (i & 31));
                }
            }
        }
        for (uint32_t i = tid; i < SIDE_N; i += DP_THREADS) S.side[i] = SIDE_EMPTY;
        if (tid == 0) { S.nw = 0; S.ns = 0; }
    }
    __syncthreads();

    // prologue: stage the first tile and issue its payload loads
    uint32_t nval[DP_QPT];
    if (t_begin < t_end) stage_tile(S, 0, in, blk_off, t_begin * 64, nblocks);
    __syncthreads();
    if (t_begin < t_end) fetch_payload<WONLY>(S, 0, in, (uint32_t)((nblocks - t_begin * 64 < 64) ? (nblocks - t_begin * 64) : 64), nval);

    for (uint64_t t = t_begin; t < t_end; ++t) {
        const uint64_t b0 = t * 64;
        const uint32_t nb_tile = (uint32_t)((nblocks - b0 < 64) ? (nblocks - b0) : 64);
        const int cur = (int)((t - t_begin) & 1);
        // stage the NEXT tile's offsets + signatures now; they become visible at S1 and feed the payload prefetch
        if (t + 1 < t_end) stage_tile(S, cur ^ 1, in, blk_off, b0 + 64, nblocks);

        // ---- phase A: my quads (prefetched); writers compact themselves; readers read the pre-tile dictionary ----------
        uint32_t val[DP_QPT];       // PLAIN: the quad; MAP: hash from the stream
        uint32_t fa[DP_QPT];        // readers: pre-tile fingerprint
        uint32_t kind = 0;          // per sub-row: bit j = active, bit 4+j = writer, bit 8+j = reader bucket touched pre-tile, bit 12+j = raw (copy mode)
        uint32_t wb[DP_QPT], wtot = 0;
#pragma unroll
        for (int j = 0; j < DP_QPT; ++j) {
            const uint32_t bl = warp * 2 + (j >> 1);
            bool active = bl < nb_tile, writer = false;
            val[j] = nval[j]; fa[j] = 0;
            if (active) {
                const uint32_t flag = (S.bsig[cur][2 * bl + (j & 1)] >> lane) & 1u;
                if ((S.bcopy[cur][bl >> 5] >> (bl & 31)) & 1u) {
                    kind |= 1u << (12 + j);             // raw quad of a copy-mode block: no dictionary access at all
                } else if (flag) {
                    if (!WONLY) {
                        fa[j] = S.tab[val[j]];
                        if (fa[j] != 0 || bit_test(S.vbit, val[j])) kind |= 1u << (8 + j);
                    }
                } else {
                    writer = true;
                }
                kind |= 1u << j;
            }
            if (writer) kind |= 1u << (4 + j);
            wb[j] = __ballot_sync(0xFFFFFFFFu, writer);
            wtot += __popc(wb[j]);
        }
        if (wtot) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(&S.nw, wtot);
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
#pragma unroll
            for (int j = 0; j < DP_QPT; ++j) {
                if (kind & (1u << (4 + j))) {
                    const uint32_t p = hash_prod(val[j]);
                    S.wrec[base + __popc(wb[j] & lanemask_lt())] = make_uint2(prod_hash(p) | (prod_fp(p, val[j]) << 16), warp * 128 + j * 32 + lane);
                }
                base += __popc(wb[j]);
            }
        }
        __syncthreads();  // S1: readers have read tab; writer list complete; next tile's signatures staged
        const uint32_t nw = S.nw;
        if (t + 1 < t_end)   // payload of the next tile: in flight during phases B-D
            fetch_payload<WONLY>(S, cur ^ 1, in, (uint32_t)((nblocks - (b0 + 64) < 64) ? (nblocks - (b0 + 64)) : 64), nval);

        // ---- phase B: writers publish ----------------------------------------------------------------------------------
#pragma unroll 1
        for (uint32_t i = tid; i < nw; i += DP_THREADS) {
            const uint2 r = S.wrec[i];
            const uint32_t hh = r.x & 0xFFFFu;
            S.tab[hh] = (uint16_t)(r.x >> 16);
            atomicMin(&S.side[hh & (SIDE_N - 1)], ((r.y & 0xFFFu) << 16) | hh);
            atomicOr(&S.wbit[hh >> 5], 1u << (hh & 31));
        }
        __syncthreads();  // S2

        // ---- phase C: readers re-read; writers check agreement ---------------------------------------------------------
        const uint64_t q0 = b0 * 64;   // first output quad of the tile
        if (!WONLY) {
#pragma unroll
            for (int j = 0; j < DP_QPT; ++j) {
                const uint32_t pos = warp * 128 + j * 32 + lane;
                const bool active = kind & (1u << j), writer = kind & (1u << (4 + j));
                const bool touched = kind & (1u << (8 + j));
                bool suspect = false;
                if (active) {
                    if (writer || (kind & (1u << (12 + j)))) {
                        out[q0 + pos] = val[j];
                    } else if (!bit_test(S.wbit, val[j])) {
                        // no PLAIN quad of this tile falls into my bucket: the pre-tile dictionary decides (empty -> 0, chameleon.rs:41)
                        out[q0 + pos] = touched ? quad_from_hf(val[j], fa[j]) : 0u;
                    } else {
                        suspect = true;
                    }
                }
                const uint32_t sm = __ballot_sync(0xFFFFFFFFu, suspect);
                if (sm) {
                    uint32_t base = 0;
                    if (lane == 0) base = atomicAdd(&S.ns, (uint32_t)__popc(sm));
                    base = __shfl_sync(0xFFFFFFFFu, base, 0);
                    if (suspect) {
                        const uint32_t e = base + __popc(sm & lanemask_lt());
                        S.wrec[TILE_Q - 1 - e] = make_uint2(val[j], pos | (touched ? S_TOUCHED : 0u) | (fa[j] << 16));
                    }
                }
            }
        }
#pragma unroll 1
        for (uint32_t i = tid; i < nw; i += DP_THREADS) {
            const uint2 r = S.wrec[i];
            const uint32_t hh = r.x & 0xFFFFu;
            const uint32_t slot = S.side[hh & (SIDE_N - 1)];
            if ((slot & 0xFFFFu) != hh || S.tab[hh] != (r.x >> 16)) {   // foreign slot owner, or the writers of my bucket disagree
                S.wrec[i].y = r.y | W_CONF;
                atomicOr(&S.conf[hh >> 5], 1u << (hh & 31));
            }
        }
        __syncthreads();  // S3
        const uint32_t ns = S.ns;

        // ---- phase D: suspect readers resolve; writers of disagreeing buckets leave the last value ----------------------
        if (!WONLY) {
#pragma unroll 1
            for (uint32_t base = warp * 32; base < ns; base += DP_THREADS) {
                const uint32_t i = base + lane;
                uint32_t hs = 0, pos = 0, fval = 0; bool have = false, search = false;
                if (i < ns) {
                    const uint2 r = S.wrec[TILE_Q - 1 - i];
                    hs = r.x & 0xFFFFu; pos = r.y & 0xFFFu;
                    fval = r.y >> 16; have = (r.y & S_TOUCHED) != 0;             // pre-tile value
                    const uint32_t slot = S.side[hs & (SIDE_N - 1)];
                    if ((slot & 0xFFFFu) == hs && pos < (slot >> 16)) {
                        // every writer of my bucket comes after me: pre-tile value
                    } else if ((slot & 0xFFFFu) == hs && !bit_test(S.conf, hs)) {
                        fval = S.tab[hs]; have = true;                           // the writers agree and the first one precedes me
                    } else {
                        search = true;                                           // disagreeing writers, or the side slot belongs to another bucket
                    }
                }
                // warp-cooperative search of the tile's writer list for the predecessor of each lane that needs it
                uint32_t todo = __ballot_sync(0xFFFFFFFFu, search);
                while (todo) {
                    const int src = __ffs(todo) - 1; todo &= todo - 1;
                    const uint32_t shs = __shfl_sync(0xFFFFFFFFu, hs, src), spos = __shfl_sync(0xFFFFFFFFu, pos, src);
                    uint32_t key = 0;
                    for (uint32_t k = lane; k < nw; k += 32) {
                        const uint2 d = S.wrec[k];
                        const uint32_t pk = d.y & 0xFFFu;
                        if ((d.x & 0xFFFFu) == shs && pk < spos) key = max(key, ((pk + 1) << 16) | (d.x >> 16));
                    }
                    key = __reduce_max_sync(0xFFFFFFFFu, key);
                    if ((int)lane == src && key) { fval = key & 0xFFFFu; have = true; }
                }
                if (i < ns) out[q0 + pos] = have ? quad_from_hf(hs, fval) : 0u;
            }
        }
#pragma unroll 1
        for (uint32_t base = warp * 32; base < nw; base += DP_THREADS) {
            const uint32_t i = base + lane;
            uint32_t hh = 0x10000u, ff = 0, pos = 0; bool confw = false;
            if (i < nw) {
                const uint2 r = S.wrec[i];
                hh = r.x & 0xFFFFu; ff = r.x >> 16; pos = r.y & 0xFFFu; confw = (r.y & W_CONF) != 0;
                if (ff == 0) atomicOr(&S.vbit[hh >> 5], 1u << (hh & 31));
            }
            // writers of disagreeing buckets: the last one (largest position) leaves its value (chameleon.rs:59)
            uint32_t todo = __ballot_sync(0xFFFFFFFFu, confw);
            while (todo) {
                const int src = __ffs(todo) - 1; todo &= todo - 1;
                const uint32_t shh = __shfl_sync(0xFFFFFFFFu, hh, src), spos = __shfl_sync(0xFFFFFFFFu, pos, src);
                bool later = false;
                for (uint32_t k = lane; k < nw; k += 32) {
                    const uint2 d = S.wrec[k];
                    later |= (d.x & 0xFFFFu) == shh && (d.y & 0xFFFu) > spos;
                }
                later = __any_sync(0xFFFFFFFFu, later);
                if ((int)lane == src && !later) S.tab[hh] = (uint16_t)ff;
            }
        }
        __syncthreads();  // S4: all readers of conf/side/lists done
#pragma unroll 1
        for (uint32_t i = tid; i < nw; i += DP_THREADS) {
            const uint2 r = S.wrec[i];
            if (r.y & W_CONF) atomicAnd(&S.conf[(r.x & 0xFFFFu) >> 5], ~(1u << (r.x & 31)));
            atomicAnd(&S.wbit[(r.x & 0xFFFFu) >> 5], ~(1u << (r.x & 31)));
            S.side[(r.x & 0xFFFFu) & (SIDE_N - 1)] = SIDE_EMPTY;
        }
        if (tid == 0) { S.nw = 0; S.ns = 0; }
        __syncthreads();  // S5: counters reset before the next tile's phase A
    }

    for (uint32_t i = tid; i < 65536; i += DP_THREADS) {
        uint32_t v = S.tab[i];
        uint32_t tch = (v != 0 || bit_test(S.vbit, i)) ? 0x10000u : 0u;
        final_tab[(size_t)run * 65536 + i] = v | tch;
    }
}

// ------------------------------------------------------------------------------------------------------------------------------
// Decode pass, second formulation (same contract as cham_decode_pass; the structure of the round-2 encoder flag pass,
// chameleon_encode.cu cham_flag_pass6: nothing but loads, compares and ballots on the per-quad path; atomics and searches only on
// the few "dirty" quads, one record per lane).  512 threads, 8 quads per thread, tile = 4096 quads = 64 blocks:
//   A  readers (MAP quads) read the pre-tile dictionary; writers (PLAIN quads, ~7 %) compute hash and fingerprint.        (barrier)
//   B  writers store their fingerprint (racy on purpose) and raise the bucket's byte in a hashed per-tile mark map.        (barrier)
//   C  raw and PLAIN quads go out as they are; a reader whose mark byte is clear saw no writer of its bucket in this tile: the
//      pre-tile value is its value (coalesced store). Writers and the remaining readers (suspects) are compacted in stream order into
//      the warp's record region; writers drop their record index into the mailbox of their bucket (4096 slots x 4 + overflow).  (barrier)
//   D  one record per lane: a suspect takes the fingerprint of the writer with the largest smaller index in its bucket, or its own
//      pre-tile value; the writer without a successor leaves the bucket's final fingerprint and clears the mark.            (barrier)
// WONLY (writer pass: the run's last-writer table only) needs A, the deposit and D.
// A mailbox overflow (more than ~20 PLAIN quads of one bucket in one tile) sends the tile to d7_replay (one warp, in order).
// ------------------------------------------------------------------------------------------------------------------------------
constexpr int D7_THREADS = 512, D7_QPT = 8, D7_NW = D7_THREADS / 32, D7_WQ = 32 * D7_QPT;
static_assert(D7_THREADS * D7_QPT == TILE_Q, "tile geometry");
constexpr int D7_MB_SLOTS = 4096, D7_MB_CAP = 4, D7_SEC_SLOTS = 64, D7_SEC_CAP = 16, D7_MARK_N = 8192;
constexpr uint32_t D7_TOUCHED = 1u << 12, D7_WRITER = 1u << 13;   // record.y: pos (12) | touched | writer | pre-tile fingerprint << 16

struct Dec7Smem {
    uint16_t tab[65536];
    uint32_t vbit[2048];
    uint2 rec[TILE_Q];            // warp w: records [256 w, 256 w + cnt[w]) in stream order. x = hash | fp << 16 (fp: writers), y see above
    union {
        uint16_t mb[D7_MB_SLOTS][D7_MB_CAP];     // writers only: (hash >> 12) << 12 | record index
        uint32_t wseen[2048];                    // fallback only: bucket written so far in this tile
    };
    uint32_t mbcnt[2][D7_MB_SLOTS / 4];
    __align__(16) uint32_t sec[D7_SEC_SLOTS][D7_SEC_CAP];
    uint32_t seccnt[2][D7_SEC_SLOTS];
    uint8_t wmark[D7_MARK_N];     // per tile: some writer's bucket hashes here
    unsigned long long boff[2][64];
    uint32_t bsig[2][128];
    uint32_t bcopy[2][2];
    uint32_t cnt[32];
    uint32_t overflow;
};
static_assert(sizeof(Dec7Smem) <= 227 * 1024, "decode pass shared memory");

// Staging of a tile's block offsets and signatures, three tiles deep in registers of threads 0..63 so that neither of the two
// dependent global loads (offset -> signature) is waited for: at the top of tile t the values of tile t+1 (loads issued during tile
// t-1) go to shared memory, the signature loads of tile t+2 are issued from its offsets (loaded during tile t-1) and the offset loads
// of tile t+3 are issued.
struct Stage7 {
    unsigned long long o_sig;    // tile t+1 (then t+2): payload offset / copy flag as loaded
    uint32_t lo, hi;             // its signature halves (in flight)
    unsigned long long o_next;   // tile t+2 (then t+3): raw blk_off entry (in flight)
};
__device__ __forceinline__ unsigned long long stage7_load_off(const uint64_t* __restrict__ blk_off, uint64_t b, uint64_t nblocks) {
    return b < nblocks ? __ldg(reinterpret_cast<const unsigned long long*>(blk_off) + b) : ~0ull;     // ~0: no such block
}
__device__ __forceinline__ void stage7_load_sig(const uint8_t* __restrict__ in, unsigned long long o, uint32_t& lo, uint32_t& hi) {
    lo = 0; hi = 0;
    if (o != ~0ull && !(o & BLK_COPY)) {
        const uint8_t* p = in + o;
        lo = ldu16(p) | (ldu16(p + 2) << 16); hi = ldu16(p + 4) | (ldu16(p + 6) << 16);
    }
}
__device__ __forceinline__ void stage7_commit(Dec7Smem& S, int buf, unsigned long long o, uint32_t lo, uint32_t hi) {
    const uint32_t tid = threadIdx.x;     // < 64: warps 0 and 1, whole warps
    const bool none = o == ~0ull;
    const bool copied = !none && (o & BLK_COPY) != 0;
    unsigned long long off = none ? 0ull : (o & ~BLK_COPY);
    if (!none && !copied) off += 8;       // payload starts behind the signature; a copy-mode block is 64 raw quads at the block start
    S.boff[buf][tid] = off; S.bsig[buf][2 * tid] = lo; S.bsig[buf][2 * tid + 1] = hi;
    const uint32_t cmask = __ballot_sync(0xFFFFFFFFu, copied);
    if ((tid & 31) == 0) S.bcopy[buf][tid >> 5] = cmask;
}
// predicated 16-bit load without a branch: the loads of all sub-rows leave back to back (a branch per sub-row would make every
// sub-row wait for its own load: 8 exposed HBM latencies per tile)
__device__ __forceinline__ uint32_t ldu16_if(const uint8_t* p, bool pred) {
    uint32_t v;
    asm volatile("{ .reg .pred q; setp.ne.u32 q, %2, 0; mov.u32 %0, 0; @q ld.global.nc.u16 %0, [%1]; }" : "=r"(v) : "l"(p), "r"((uint32_t)pred));
    return v;
}
template <bool WONLY>
__device__ __forceinline__ void fetch_payload7(const Dec7Smem& S, int buf, const uint8_t* __restrict__ in, uint32_t nb_tile, uint32_t (&v)[D7_QPT]) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t a[D7_QPT], b[D7_QPT];
#pragma unroll
    for (int j = 0; j < D7_QPT; ++j) {
        const uint32_t bl = warp * (D7_WQ / 64) + (j >> 1);
        const uint32_t k = (j & 1) * 32 + lane;
        const bool in_tile = bl < nb_tile;
        const uint32_t blc = in_tile ? bl : 0u;
        const uint32_t lo = S.bsig[buf][2 * blc], hi = S.bsig[buf][2 * blc + 1];
        const uint32_t flag = (((j & 1) ? hi : lo) >> lane) & 1u;
        const uint32_t before = (j & 1) ? (__popc(lo) + __popc(hi & lanemask_lt())) : __popc(lo & lanemask_lt());
        const uint8_t* p = in + S.boff[buf][blc] + 4 * k - 2 * before;
        a[j] = ldu16_if(p, in_tile && !(WONLY && flag));        // MAP: the 16-bit hash (chameleon.rs:64); PLAIN: low half of the quad (:56)
        b[j] = ldu16_if(p + 2, in_tile && !flag);               // PLAIN: high half
    }
#pragma unroll
    for (int j = 0; j < D7_QPT; ++j) v[j] = a[j] | (b[j] << 16);
}

// Fallback: the tile's writers and suspects in stream order by one warp.
template <bool WONLY>
__device__ __noinline__ void d7_replay(Dec7Smem& S, uint32_t* __restrict__ out, uint64_t q0) {
    const uint32_t lane = threadIdx.x & 31;
    for (uint32_t i = lane; i < 2048; i += 32) S.wseen[i] = 0;
    const uint32_t c = lane < (uint32_t)D7_NW ? S.cnt[lane] : 0u;
    uint32_t incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if ((int)lane >= d) incl += t; }
    const uint32_t excl = incl - c;
    const uint32_t n = __shfl_sync(0xFFFFFFFFu, incl, 31);
    __syncwarp();
    #pragma unroll 1
    for (uint32_t i0 = 0; i0 < n; i0 += 32) {
        const uint32_t i = i0 + lane;
        const bool valid = i < n;
        uint32_t w = 0;
#pragma unroll
        for (int b = 16; b >= 1; b >>= 1) { const uint32_t t = __shfl_sync(0xFFFFFFFFu, incl, (w + b - 1) & 31); if (t <= i) w += b; }
        const uint32_t e = __shfl_sync(0xFFFFFFFFu, excl, w & 31);
        uint2 r = make_uint2(0, 0);
        if (valid) r = S.rec[w * D7_WQ + (i - e)];
        const uint32_t hh = r.x & 0xFFFFu, ff = r.x >> 16, pos = r.y & 0xFFFu;
        const bool writer = valid && (r.y & D7_WRITER);
        const uint32_t grp = __match_any_sync(0xFFFFFFFFu, valid ? hh : 0x10000u + lane);
        const uint32_t wm = __ballot_sync(0xFFFFFFFFu, writer);
        const uint32_t lw = grp & wm & lanemask_lt();              // earlier writers of my bucket inside the step
        const uint32_t fprev = __shfl_sync(0xFFFFFFFFu, ff, lw ? 31 - __clz(lw) : 0);
        if (valid && !writer && !WONLY) {
            uint32_t fv = r.y >> 16; bool have = (r.y & D7_TOUCHED) != 0;
            if (lw) { fv = fprev; have = true; }
            else if ((S.wseen[hh >> 5] >> (hh & 31)) & 1u) { fv = S.tab[hh]; have = true; }
            out[q0 + pos] = have ? quad_from_hf(hh, fv) : 0u;
        }
        __syncwarp();
        if (writer && (grp & wm & lanemask_gt()) == 0) {           // last writer of the bucket inside the step
            S.tab[hh] = (uint16_t)ff;
            if (ff == 0) atomicOr(&S.vbit[hh >> 5], 1u << (hh & 31));
            atomicOr(&S.wseen[hh >> 5], 1u << (hh & 31));
            S.wmark[hh & (D7_MARK_N - 1)] = 0;
        }
        __syncwarp();
    }
}

template <bool WONLY>
__global__ void __launch_bounds__(D7_THREADS, 1)
cham_decode_pass7(const uint8_t* __restrict__ in, const uint64_t* __restrict__ blk_off, DecStatus* st,
                  uint32_t nruns, uint32_t* __restrict__ out /* quads */, const uint32_t* __restrict__ carry,
                  uint32_t* __restrict__ final_tab) {
    if (st->nonquiet || st->error) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Dec7Smem& S = *reinterpret_cast<Dec7Smem*>(smem_raw);
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t run = blockIdx.x;
    const uint64_t nblocks = st->main_blocks;
    const uint64_t ntiles = (nblocks + 63) / 64;
    const uint64_t t_begin = (uint64_t)run * ntiles / nruns, t_end = (uint64_t)(run + 1) * ntiles / nruns;
    {
        const uint4 z = make_uint4(0, 0, 0, 0);
        uint4* t4 = reinterpret_cast<uint4*>(S.tab);
        for (uint32_t i = tid; i < 65536 * 2 / 16; i += D7_THREADS) t4[i] = z;
        for (uint32_t i = tid; i < 2048; i += D7_THREADS) S.vbit[i] = 0;
        for (uint32_t i = tid; i < D7_MB_SLOTS / 4; i += D7_THREADS) { S.mbcnt[0][i] = 0; S.mbcnt[1][i] = 0; }
        for (uint32_t i = tid; i < D7_MARK_N / 4; i += D7_THREADS) reinterpret_cast<uint32_t*>(S.wmark)[i] = 0;
        if (tid < D7_SEC_SLOTS) { S.seccnt[0][tid] = 0; S.seccnt[1][tid] = 0; }
        if (tid == 0) S.overflow = 0;
        if (!WONLY) {
            __syncthreads();
            const uint32_t* __restrict__ cr = carry + (size_t)run * 65536;   // dictionary before this run
            for (uint32_t i = tid; i < 65536; i += D7_THREADS) {
                const uint32_t c = cr[i];
                if (c & 0x10000u) {
                    S.tab[i] = (uint16_t)c;
                    if ((c & 0xFFFFu) == 0) atomicOr(&S.vbit[i >> 5], 1u << (i & 31));
                }
            }
        }
    }
    __syncthreads();
    uint32_t nval[D7_QPT];
    Stage7 sg; sg.o_sig = ~0ull; sg.lo = 0; sg.hi = 0; sg.o_next = ~0ull;
    if (t_begin < t_end && tid < 64) {
        const unsigned long long o0 = stage7_load_off(blk_off, t_begin * 64 + tid, nblocks);
        uint32_t lo0, hi0;
        stage7_load_sig(in, o0, lo0, hi0);
        stage7_commit(S, 0, o0, lo0, hi0);
        if (t_begin + 1 < t_end) { sg.o_sig = stage7_load_off(blk_off, (t_begin + 1) * 64 + tid, nblocks); stage7_load_sig(in, sg.o_sig, sg.lo, sg.hi); }
        if (t_begin + 2 < t_end) sg.o_next = stage7_load_off(blk_off, (t_begin + 2) * 64 + tid, nblocks);
    }
    __syncthreads();
    if (t_begin < t_end) fetch_payload7<WONLY>(S, 0, in, (uint32_t)((nblocks - t_begin * 64 < 64) ? (nblocks - t_begin * 64) : 64), nval);

    for (uint64_t t = t_begin; t < t_end; ++t) {
        const uint64_t b0 = t * 64;
        const uint32_t nb_tile = (uint32_t)((nblocks - b0 < 64) ? (nblocks - b0) : 64);
        const int cur = (int)((t - t_begin) & 1);
        const uint32_t buf = (uint32_t)cur;
        if (tid < 64) {
            if (t + 1 < t_end) stage7_commit(S, cur ^ 1, sg.o_sig, sg.lo, sg.hi);          // tile t+1: loaded during tile t-1
            sg.o_sig = sg.o_next;                                                           // tile t+2: its signatures leave now
            sg.lo = 0; sg.hi = 0;
            if (t + 2 < t_end) stage7_load_sig(in, sg.o_sig, sg.lo, sg.hi);
            sg.o_next = (t + 3 < t_end) ? stage7_load_off(blk_off, (t + 3) * 64 + tid, nblocks) : ~0ull;   // tile t+3: its offsets leave now
        }
#pragma unroll
        for (int k = 0; k < D7_MB_SLOTS / 4 / D7_THREADS; ++k) S.mbcnt[buf ^ 1u][tid + k * D7_THREADS] = 0;
        if (tid < D7_SEC_SLOTS) S.seccnt[buf ^ 1u][tid] = 0;

        // ---- A
        uint32_t val[D7_QPT], hk[D7_QPT], fa[D7_QPT];   // val: the quad (PLAIN / raw) or the hash (MAP); hk: hash; fa: readers pre-tile fp, writers own fp
        uint32_t act = 0, wr = 0, raw = 0, tch = 0;     // bit j
#pragma unroll
        for (int j = 0; j < D7_QPT; ++j) {
            const uint32_t bl = warp * (D7_WQ / 64) + (j >> 1);
            val[j] = nval[j]; hk[j] = 0; fa[j] = 0;
            if (bl < nb_tile) {
                act |= 1u << j;
                const uint32_t flag = (S.bsig[cur][2 * bl + (j & 1)] >> lane) & 1u;
                if ((S.bcopy[cur][bl >> 5] >> (bl & 31)) & 1u) raw |= 1u << j;
                else if (flag) {
                    if (!WONLY) {
                        hk[j] = val[j];
                        fa[j] = S.tab[hk[j]];
                        if (fa[j] != 0 || bit_test(S.vbit, hk[j])) tch |= 1u << j;
                    }
                } else {
                    wr |= 1u << j;
                    const uint32_t p = hash_prod(val[j]);
                    hk[j] = prod_hash(p); fa[j] = prod_fp(p, val[j]);
                }
            }
        }
        __syncthreads();  // S1: readers have read the pre-tile dictionary; next tile's signatures staged
        if (t + 1 < t_end)   // payload of the next tile: in flight during the rest of this tile
            fetch_payload7<WONLY>(S, cur ^ 1, in, (uint32_t)((nblocks - (b0 + 64) < 64) ? (nblocks - (b0 + 64)) : 64), nval);
        if (!WONLY) {
            // ---- B
#pragma unroll
            for (int j = 0; j < D7_QPT; ++j)
                if ((wr >> j) & 1u) { S.tab[hk[j]] = (uint16_t)fa[j]; S.wmark[hk[j] & (D7_MARK_N - 1)] = 1; }
            __syncthreads();  // S2
        }
        // ---- C
        const uint64_t q0 = b0 * 64;   // first output quad of the tile
        uint32_t base = 0;
        uint2* __restrict__ myrec = S.rec + warp * D7_WQ;
#pragma unroll
        for (int j = 0; j < D7_QPT; ++j) {
            const uint32_t pos = warp * D7_WQ + j * 32 + lane;
            const bool active = (act >> j) & 1u, writer = (wr >> j) & 1u;
            bool dirty = writer;
            if (!WONLY && active) {
                if (writer || ((raw >> j) & 1u)) out[q0 + pos] = val[j];
                else if (S.wmark[hk[j] & (D7_MARK_N - 1)] == 0)
                    out[q0 + pos] = ((tch >> j) & 1u) ? quad_from_hf(hk[j], fa[j]) : 0u;      // empty bucket -> 0 (chameleon.rs:41)
                else dirty = true;
            }
            const uint32_t db = __ballot_sync(0xFFFFFFFFu, dirty);
            if (dirty) myrec[base + __popc(db & lanemask_lt())] =
                make_uint2(hk[j] | (writer ? fa[j] << 16 : 0u), pos | (((tch >> j) & 1u) ? D7_TOUCHED : 0u) | (writer ? D7_WRITER : 0u) | (writer ? 0u : fa[j] << 16));
            base += __popc(db);
        }
        if (lane == 0) S.cnt[warp] = base;
        __syncwarp();
        // deposit (writers only)
        uint2 r0 = make_uint2(0, 0);
        #pragma unroll 1
        for (uint32_t i = lane; i < base; i += 32) {
            const uint2 r = myrec[i];
            if (i < 32) r0 = r;
            if (r.y & D7_WRITER) {
                const uint32_t hh = r.x & 0xFFFFu, slot = hh & (D7_MB_SLOTS - 1), sh = (slot & 3u) * 8u;
                const uint32_t k = (atomicAdd(&S.mbcnt[buf][slot >> 2], 1u << sh) >> sh) & 0xFFu;
                if (k < (uint32_t)D7_MB_CAP) S.mb[slot][k] = (uint16_t)(((hh >> 12) << 12) | (warp * D7_WQ + i));
                else {
                    const uint32_t s2 = slot & (D7_SEC_SLOTS - 1);
                    const uint32_t k2 = atomicAdd(&S.seccnt[buf][s2], 1u);
                    if (k2 < (uint32_t)D7_SEC_CAP) S.sec[s2][k2] = (hh << 12) | (warp * D7_WQ + i);
                    else S.overflow = 1;
                }
            }
        }
        __syncthreads();  // S3
        if (S.overflow) {
            if (warp == 0) d7_replay<WONLY>(S, out, q0);
        } else {
            // ---- D
            #pragma unroll 1
            for (uint32_t i0 = 0; i0 < base; i0 += 32) {
                const uint32_t i = i0 + lane;
                const bool valid = i < base;
                uint2 r = r0;
                if (i0) r = valid ? myrec[i] : make_uint2(0, 0);
                if (valid) {
                    const uint32_t hh = r.x & 0xFFFFu, slot = hh & (D7_MB_SLOTS - 1), myidx = warp * D7_WQ + i;
                    const uint32_t n = (S.mbcnt[buf][slot >> 2] >> ((slot & 3u) * 8u)) & 0xFFu;
                    const uint2 e2 = *reinterpret_cast<const uint2*>(&S.mb[slot][0]);
                    const uint32_t me = ((hh >> 12) << 12) | myidx;
                    int best = -1; bool later = false;
#pragma unroll
                    for (int tt = 0; tt < D7_MB_CAP; ++tt) {
                        const uint32_t e = ((tt & 2) ? e2.y : e2.x) >> ((tt & 1) * 16) & 0xFFFFu;
                        if ((uint32_t)tt < n && ((e ^ me) >> 12) == 0) {
                            if (e < me) best = max(best, (int)(e & 0xFFFu));
                            later |= e > me;
                        }
                    }
                    if (n > (uint32_t)D7_MB_CAP) {
                        const uint32_t s2 = slot & (D7_SEC_SLOTS - 1);
                        const uint32_t n2 = S.seccnt[buf][s2];
                        const uint32_t mine = (hh << 12) | myidx;
                        #pragma unroll 1
                        for (uint32_t t4 = 0; t4 < n2; t4 += 4) {
                            const uint4 e4 = *reinterpret_cast<const uint4*>(&S.sec[s2][t4]);
                            const uint32_t ev[4] = {e4.x, e4.y, e4.z, e4.w};
#pragma unroll
                            for (int tt = 0; tt < 4; ++tt) {
                                const uint32_t e = ev[tt];
                                if (t4 + tt < n2 && ((e ^ mine) & 0xFFFFF000u) == 0) {
                                    if (e < mine) best = max(best, (int)(e & 0xFFFu));
                                    later |= e > mine;
                                }
                            }
                        }
                    }
                    if (r.y & D7_WRITER) {
                        if (!later) {      // chameleon.rs:59: the last PLAIN quad of the bucket leaves its value
                            S.tab[hh] = (uint16_t)(r.x >> 16);
                            if ((r.x >> 16) == 0) atomicOr(&S.vbit[hh >> 5], 1u << (hh & 31));
                            S.wmark[hh & (D7_MARK_N - 1)] = 0;
                        }
                    } else if (!WONLY) {
                        uint32_t fv = r.y >> 16; bool have = (r.y & D7_TOUCHED) != 0;
                        if (best >= 0) { fv = S.rec[best].x >> 16; have = true; }
                        out[q0 + (r.y & 0xFFFu)] = have ? quad_from_hf(hh, fv) : 0u;
                    }
                }
            }
        }
        __syncthreads();  // S4
        if (tid == 0) S.overflow = 0;
    }
    for (uint32_t i = tid; i < 65536; i += D7_THREADS) {
        const uint32_t v = S.tab[i];
        const uint32_t tchd = (v != 0 || bit_test(S.vbit, i)) ? 0x10000u : 0u;
        final_tab[(size_t)run * 65536 + i] = v | tchd;
    }
}

// carry-in fold for decode: initial dictionary is all zero values (chameleon.rs:41): nothing touched.
__global__ void dec_carry_scan(const uint32_t* __restrict__ final_tab, uint32_t nruns, uint32_t* __restrict__ carry, uint32_t* __restrict__ dict_out) {
    uint32_t hb = blockIdx.x * blockDim.x + threadIdx.x;
    if (hb >= 65536) return;
    uint32_t c = 0;
    for (uint32_t r = 0; r < nruns; ++r) {
        carry[(size_t)r * 65536 + hb] = c;
        uint32_t v = final_tab[(size_t)r * 65536 + hb];
        if (v & 0x10000u) c = v;
    }
    dict_out[hb] = (c & 0x10000u) ? quad_from_hf(hb, c & 0xFFFFu) : 0u;   // full-width dictionary for the tail loop
}

// ---- 4. tail loop (codec.rs:102-123), one thread, literal control flow ---------------------------------------------------
__global__ void dec_tail(const uint8_t* __restrict__ in, uint64_t n, uint8_t* __restrict__ out, uint64_t cap, uint32_t* __restrict__ dict,
                         DecStatus* __restrict__ st, uint64_t* __restrict__ d_out_size) {
    if (threadIdx.x || blockIdx.x) return;
    if (st->nonquiet) { if (d_out_size) *d_out_size = 0; return; }   // gave up: the caller's in-order fallback (queued behind) produces the result
    if (st->error) { st->out_bytes = 0; if (d_out_size) *d_out_size = 0; return; }
    uint64_t idx = st->tail_off, oidx = st->main_blocks * 256;
    Protection ps; ps.init();
    ps.counter = st->main_blocks;             // quiet so far: penalty 0, start 1 (protection_state.rs:9-16,38-43)
    ps.previous_incompressible = st->last_main_inc;
    if (st->seq) { ps.copy_penalty = st->ps_penalty; ps.copy_penalty_start = st->ps_start; ps.previous_incompressible = st->ps_prev; }
    bool bad = false, overflow = false;
    auto emit = [&](uint32_t q) {
        if (oidx + 4 > cap) { overflow = true; return; }
        out[oidx] = (uint8_t)q; out[oidx + 1] = (uint8_t)(q >> 8); out[oidx + 2] = (uint8_t)(q >> 16); out[oidx + 3] = (uint8_t)(q >> 24);
        oidx += 4;
    };
    while (!bad && !overflow && n - idx > 0) {
        if (ps.revert_to_copy()) {
            const uint64_t rem = n - idx;
            const uint64_t len = rem > 256 ? 256 : rem;
            if (oidx + len > cap) { overflow = true; break; }
            for (uint64_t i = 0; i < len; ++i) out[oidx + i] = in[idx + i];
            oidx += len; idx += len;
            if (rem <= 256) break;
            ps.decay();
        } else {
            const uint64_t mark = idx;
            if (n - idx < 8) { bad = true; break; }
            uint64_t sig = 0;
            for (int i = 0; i < 8; ++i) sig |= (uint64_t)in[idx + i] << (8 * i);
            idx += 8;
            bool end = false;
            for (int u = 0; u < 32 && !end && !bad && !overflow; ++u) {        // 32 units of 2 quads (chameleon.rs:143)
                const bool checked = (n - idx) < 8;
                for (int k = 0; k < 2 && !end; ++k) {
                    const uint32_t fl = (uint32_t)(sig & 1); sig >>= 1;
                    if (checked && fl == 0) {                                    // decode_partial_unit, chameleon.rs:119-129
                        const uint64_t rem = n - idx;
                        if (rem == 0) { end = true; break; }
                        if (rem < 4) {
                            if (oidx + rem > cap) { overflow = true; end = true; break; }
                            for (uint64_t i = 0; i < rem; ++i) out[oidx++] = in[idx++];
                            end = true; break;
                        }
                    }
                    uint32_t q;
                    if (fl) {
                        if (n - idx < 2) { bad = true; break; }
                        q = dict[in[idx] | (in[idx + 1] << 8)]; idx += 2;
                    } else {
                        if (n - idx < 4) { bad = true; break; }
                        q = in[idx] | (in[idx + 1] << 8) | (in[idx + 2] << 16) | ((uint32_t)in[idx + 3] << 24); idx += 4;
                        dict[prod_hash(hash_prod(q))] = q;
                    }
                    emit(q);
                }
            }
            if (end) break;
            ps.update(idx - mark >= 256);
        }
    }
    uint64_t res = oidx;
    if (bad) { st->error = 3; res = 0; }
    else if (overflow) { st->error = 2; res = 0; }
    st->out_bytes = res;
    if (d_out_size) *d_out_size = res;
}

}  // namespace chamdec

using namespace chamdec;

int g_cham_decode_impl = 7;   // 1: round-1 decode pass, 7: write / verify / mailbox (timing comparisons and tests)

struct ChamDecLayout { bounds::BoundsLayout B; size_t final_tab, carry, dict, total; };

static size_t dec_layout(size_t nbytes, size_t cap, int nruns_max, ChamDecLayout* L) {
    size_t off = bounds::bounds_layout<T>(nbytes, cap, &L->B);
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    L->final_tab = take((size_t)nruns_max * 65536 * sizeof(uint32_t));
    L->carry = take((size_t)nruns_max * 65536 * sizeof(uint32_t));
    L->dict = take(65536 * sizeof(uint32_t));
    L->total = off;
    return off;
}

size_t cham_decode_workspace_bytes(size_t nbytes, size_t cap, int nruns_max) { ChamDecLayout L; return dec_layout(nbytes, cap, nruns_max, &L); }

// Enqueues the parallel decode. On return (after the stream drains) *d_nonquiet != 0 means the caller must run the exact
// in-order kernel instead (copy-mode blocks present, or a pathological tile); d_out_size is only written when it is 0.
cudaError_t cham_decode_parallel(const uint8_t* d_in, size_t nbytes, uint8_t* d_out, size_t cap, uint8_t* ws, int num_sms,
                                 uint64_t* d_out_size, uint32_t* d_nonquiet, cudaStream_t stream, uint64_t* launches) {
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e0 = cudaFuncSetAttribute(cham_decode_pass<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DecSmem));
        if (e0 == cudaSuccess) e0 = cudaFuncSetAttribute(cham_decode_pass<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DecSmem));
        if (e0 == cudaSuccess) e0 = cudaFuncSetAttribute(cham_decode_pass7<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Dec7Smem));
        if (e0 == cudaSuccess) e0 = cudaFuncSetAttribute(cham_decode_pass7<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Dec7Smem));
        if (e0 != cudaSuccess) return e0;
        attr_done = true;
    }
    ChamDecLayout L; dec_layout(nbytes, cap, num_sms, &L);
    DecStatus* st = reinterpret_cast<DecStatus*>(ws + L.B.status);
    uint64_t* blk_off = reinterpret_cast<uint64_t*>(ws + L.B.blk_off);
    cudaError_t e = bounds::bounds_launch<T>(d_in, nbytes, cap, ws, L.B, stream, launches);
    if (e != cudaSuccess) return e;
    const uint64_t maxblocks = L.B.maxblocks;
    // run count from an upper bound of the block count (the kernel reads the real one from the status block)
    uint64_t tiles_ub = (maxblocks + 63) / 64;
    uint32_t nruns = (uint32_t)(tiles_ub / 16); if (nruns < 1) nruns = 1; if (nruns > (uint32_t)num_sms) nruns = num_sms;
    uint32_t* final_tab = reinterpret_cast<uint32_t*>(ws + L.final_tab);
    uint32_t* carry = reinterpret_cast<uint32_t*>(ws + L.carry);
    if (g_cham_decode_impl == 1) cham_decode_pass<true><<<nruns, DP_THREADS, sizeof(DecSmem), stream>>>(d_in, blk_off, st, nruns, nullptr, nullptr, final_tab);
    else cham_decode_pass7<true><<<nruns, D7_THREADS, sizeof(Dec7Smem), stream>>>(d_in, blk_off, st, nruns, nullptr, nullptr, final_tab);
    ++*launches;
    dec_carry_scan<<<65536 / 256, 256, 0, stream>>>(final_tab, nruns, carry, reinterpret_cast<uint32_t*>(ws + L.dict)); ++*launches;
    if (g_cham_decode_impl == 1) cham_decode_pass<false><<<nruns, DP_THREADS, sizeof(DecSmem), stream>>>(d_in, blk_off, st, nruns, reinterpret_cast<uint32_t*>(d_out), carry, final_tab);
    else cham_decode_pass7<false><<<nruns, D7_THREADS, sizeof(Dec7Smem), stream>>>(d_in, blk_off, st, nruns, reinterpret_cast<uint32_t*>(d_out), carry, final_tab);
    ++*launches;
    dec_tail<<<1, 32, 0, stream>>>(d_in, nbytes, d_out, cap, reinterpret_cast<uint32_t*>(ws + L.dict), st, d_out_size); ++*launches;
    e = cudaMemcpyAsync(d_nonquiet, &st->nonquiet, sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

}  // namespace dns
