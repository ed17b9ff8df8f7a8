// decode_bounds.cuh — block boundaries of a density stream, in parallel, for all three algorithms (sm_100a).
//
// Replaces the cursor of /root/reference/src/codec/codec.rs:88-100 (the main decode loop): nothing in the stream says where block
// b + 1 starts; an encoded block is `SIG + payload(signature)` bytes, a copy-mode block (protection_state.rs) BS raw bytes.
//
//  1. The stream is cut into chunks of T::CH bytes; for each chunk and each of the T::NCAND possible (even) entry offsets in its
//     first T::MAXBLK bytes, `dec_chunk_walk` walks the chunk in shared memory and records where that walk leaves the chunk and how
//     many blocks it saw. Composing these maps (per group of GROUP chunks, then over the groups, then back down) yields every chunk's
//     true entry point and block index; `dec_block_offsets` re-walks each chunk from its true entry and writes one offset per block.
//  2. `dec_quiet_check`: if no two consecutive blocks are incompressible (consumed >= BS, codec.rs:98) the automaton never leaves its
//     initial state and no block is in copy mode. Otherwise the candidate walks are void (a copy-mode block has no signature) and
//     `dec_seq_walk` redoes the boundaries in order with the exact automaton: chunks (and whole groups) in which the automaton
//     provably stays in encoded mode are jumped in O(1) from the candidate table, the others are walked block by block from shared
//     memory and their copy-mode blocks marked.
//
// Traits (one per algorithm): BS block bytes, SIG signature bytes, MAXBLK = SIG + BS, NCAND = MAXBLK / 2, CH chunk bytes,
// NBB = bits of the per-chunk block count in a table row, consumed(sig) = bytes of an encoded block with that signature.
#pragma once
#include "common.cuh"
#include "cl_core.cuh"

namespace dns {
namespace bounds {

struct ChamT {   // chameleon.rs:138-147: 64 one-bit flags, hit -> 2 bytes, miss -> 4 bytes
    static constexpr uint32_t BS = 256, SIG = 8, MAXBLK = 264, NCAND = 132, CH = 16384, NBB = 8;
    static __device__ __forceinline__ uint32_t consumed(uint64_t sig) { return 264u - 2u * (uint32_t)__popcll(sig); }
};
struct CheeT {   // cheetah.rs:188-197
    static constexpr uint32_t BS = 128, SIG = 8, MAXBLK = 136, NCAND = 68, CH = 4096, NBB = 12;
    static __device__ __forceinline__ uint32_t consumed(uint64_t sig) { return cld::cheetah_block_bytes(sig); }
};
struct LionT {   // lion.rs:317-351: 6-byte signature
    static constexpr uint32_t BS = 64, SIG = 6, MAXBLK = 70, NCAND = 35, CH = 4096, NBB = 12;
    static __device__ __forceinline__ uint32_t consumed(uint64_t sig) { return cld::lion_block_bytes(sig & 0x0000FFFFFFFFFFFFull); }
};

constexpr int GROUP = 64;            // chunks per composition group
constexpr uint32_t TERM = 0xFFu;     // exit code: the walk reached the tail region / end of stream
constexpr uint32_t G_SKIP = 0xFEu;   // g_entry: dec_seq_walk handled this group chunk by chunk (c_entry already written)
constexpr unsigned long long BLK_COPY = 1ull << 63;   // blk_off flag: copy-mode block (raw BS bytes, no signature)

struct DecStatus {
    unsigned long long out_bytes;
    unsigned long long main_blocks;      // blocks decoded by the parallel main loop (codec.rs:88-100)
    unsigned long long tail_off;         // stream offset where the tail loop starts
    unsigned int nonquiet, error;        // nonquiet bit 0: copy-mode blocks present (cleared again by dec_seq_walk)
    unsigned int last_main_inc, seq;     // seq: the boundaries come from dec_seq_walk, automaton state below is valid
    unsigned int ps_penalty, ps_start, ps_prev, pad;   // protection state after the main loop (protection_state.rs:9-16)
};

__device__ __forceinline__ uint32_t ldu16(const uint8_t* p) { return *reinterpret_cast<const uint16_t*>(p); }
__device__ __forceinline__ uint64_t ldsig(const uint8_t* p) {   // 8 bytes at a 2-byte aligned address
    return (uint64_t)(ldu16(p) | (ldu16(p + 2) << 16)) | ((uint64_t)(ldu16(p + 4) | (ldu16(p + 6) << 16)) << 32);
}

// table row entry: exit_idx (8) | nblocks (NBB) | x, where x = term_rel when exit_idx == TERM (the walk ended inside this chunk at
// relative offset term_rel because fewer than MAXBLK bytes remain: that is where codec.rs's tail loop takes over), else the flags
// {bit 0: two consecutive incompressible blocks inside, bit 1: first block incompressible, bit 2: last block incompressible}.
template <class T> __device__ __forceinline__ uint32_t row_pack(uint32_t exitc, uint32_t nb, uint32_t x) { return exitc | (nb << 8) | (x << (8 + T::NBB)); }
template <class T> __device__ __forceinline__ uint32_t row_nb(uint32_t r) { return (r >> 8) & ((1u << T::NBB) - 1u); }
template <class T> __device__ __forceinline__ uint32_t row_x(uint32_t r) { return r >> (8 + T::NBB); }

// ---- 1a. candidate walks -----------------------------------------------------------------------------------------------
template <class T>
__global__ void __launch_bounds__(160) dec_chunk_walk(const uint8_t* __restrict__ in, uint64_t n, uint32_t nchunks, uint32_t* __restrict__ res) {
    constexpr uint32_t SM = T::CH + T::MAXBLK + 24;
    __shared__ __align__(16) uint8_t s[SM];
    const uint32_t c = blockIdx.x;
    const uint64_t base = (uint64_t)c * T::CH;
    for (uint32_t i = threadIdx.x * 2; i < SM; i += blockDim.x * 2) {
        const uint64_t g = base + i;
        *reinterpret_cast<uint16_t*>(s + i) = (g + 2 <= n) ? *reinterpret_cast<const uint16_t*>(in + g) : (uint16_t)((g < n) ? in[g] : 0);
    }
    __syncthreads();
    const uint32_t cand = threadIdx.x;
    if (cand >= T::NCAND) return;
    uint32_t off = cand * 2, nb = 0, exitc = TERM, term = 0;
    uint32_t pair = 0, first = 0, prev = 0;
    while (true) {
        if (off >= T::CH) { exitc = (off - T::CH) >> 1; break; }
        if (base + off + T::MAXBLK > n) { term = off; break; }
        const uint32_t consumed = T::consumed(ldsig(s + off));
        const uint32_t inc = consumed >= T::BS ? 1u : 0u;           // codec.rs:98
        if (nb == 0) first = inc;
        pair |= inc & prev;
        prev = inc;
        off += consumed;
        ++nb;
    }
    res[(size_t)c * T::NCAND + cand] = row_pack<T>(exitc, nb, exitc == TERM ? term : (pair | (first << 1) | (prev << 2)));
    (void)nchunks;
}

// ---- 1b. compose the maps of GROUP consecutive chunks ----------------------------------------------------------------------
// gres[g][cand] = {exit_idx (or TERM), blocks, term_chunk, term_rel}; exit_idx != TERM: z = the chunk flags composed along the path
// (bit 0: two consecutive incompressible blocks anywhere inside the group, bit 1: first block, bit 2: last block incompressible,
//  bit 3: short last group, which dec_seq_walk never jumps)
template <class T>
__global__ void dec_group_compose(const uint32_t* __restrict__ res, uint32_t nchunks, uint4* __restrict__ gres) {
    const uint32_t g = blockIdx.x, cand = threadIdx.x;
    if (cand >= T::NCAND) return;
    uint32_t idx = cand, blocks = 0, tchunk = 0, trel = 0;
    uint32_t pair = 0, first = 0, last = 0, have = 0;
    const uint32_t c0 = g * GROUP, c1 = min(nchunks, c0 + GROUP);
    for (uint32_t c = c0; c < c1; ++c) {
        const uint32_t r = res[(size_t)c * T::NCAND + idx];
        const uint32_t nb = row_nb<T>(r);
        blocks += nb;
        idx = r & 0xFFu;
        if (idx == TERM) { tchunk = c; trel = row_x<T>(r); break; }
        if (nb) {
            const uint32_t fl = row_x<T>(r);
            if (!have) { first = (fl >> 1) & 1u; have = 1; } else pair |= last & (fl >> 1) & 1u;
            pair |= fl & 1u;
            last = (fl >> 2) & 1u;
        }
    }
    if (idx != TERM) tchunk = pair | (first << 1) | (last << 2) | ((c1 - c0 < (uint32_t)GROUP) ? 8u : 0u);
    gres[(size_t)g * T::NCAND + cand] = make_uint4(idx, blocks, tchunk, trel);
}

// ---- 1c. walk the groups from the stream start ---------------------------------------------------------------------------
template <class T>
__global__ void dec_top_walk(const uint4* __restrict__ gres, uint32_t ngroups, uint64_t n, uint32_t* __restrict__ g_entry,
                             uint64_t* __restrict__ g_blockbase, DecStatus* __restrict__ st) {
    if (threadIdx.x || blockIdx.x) return;
    uint32_t idx = 0; uint64_t blocks = 0;
    bool done = false;
    for (uint32_t g = 0; g < ngroups; ++g) {
        g_entry[g] = done ? TERM : idx;
        g_blockbase[g] = blocks;
        if (done) continue;
        const uint4 r = gres[(size_t)g * T::NCAND + idx];
        blocks += r.y;
        idx = r.x;
        if (idx == TERM) { done = true; st->tail_off = (unsigned long long)r.z * T::CH + r.w; }
    }
    if (!done) st->tail_off = n;  // cannot happen for n > 0 (the last chunk always terminates); keeps the tail kernel safe
    st->main_blocks = blocks;
}

// ---- 1d. per chunk: true entry + block index -----------------------------------------------------------------------------
template <class T>
__global__ void dec_chunk_entries(const uint32_t* __restrict__ res, uint32_t nchunks, const uint32_t* __restrict__ g_entry,
                                  const uint64_t* __restrict__ g_blockbase, uint32_t ngroups, uint32_t* __restrict__ c_entry,
                                  uint64_t* __restrict__ c_blockbase, const DecStatus* __restrict__ only_if_seq) {
    if (only_if_seq && !only_if_seq->seq) return;
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= ngroups) return;
    uint32_t idx = g_entry[g]; uint64_t blocks = g_blockbase[g];
    if (idx == G_SKIP) return;
    const uint32_t c0 = g * GROUP, c1 = min(nchunks, c0 + GROUP);
    for (uint32_t c = c0; c < c1; ++c) {
        c_entry[c] = idx; c_blockbase[c] = blocks;
        if (idx == TERM) continue;
        const uint32_t r = res[(size_t)c * T::NCAND + idx];
        blocks += row_nb<T>(r);
        idx = r & 0xFFu;
    }
}

// ---- 1e. one offset per block ------------------------------------------------------------------------------------------------
template <class T>
__global__ void dec_block_offsets(const uint8_t* __restrict__ in, uint64_t n, uint32_t nchunks, const uint32_t* __restrict__ c_entry,
                                  const uint64_t* __restrict__ c_blockbase, uint64_t* __restrict__ blk_off, uint64_t maxblocks,
                                  const DecStatus* __restrict__ only_if_seq) {
    if (only_if_seq && !only_if_seq->seq) return;
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nchunks) return;
    const uint32_t e = c_entry[c];
    if (e == TERM) return;
    const uint64_t base = (uint64_t)c * T::CH;
    uint64_t b = c_blockbase[c];
    uint32_t off = e * 2;
    while (off < T::CH && base + off + T::MAXBLK <= n) {
        if (b < maxblocks) blk_off[b] = base + off;
        ++b;
        off += T::consumed(ldsig(in + base + off));
    }
}

// ---- 2. quiet check + capacity check -------------------------------------------------------------------------------------------
template <class T>
__global__ void dec_quiet_check(const uint8_t* __restrict__ in, const uint64_t* __restrict__ blk_off, uint64_t maxblocks, DecStatus* __restrict__ st, uint64_t cap) {
    const uint64_t nb = st->main_blocks;
    const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // DENSITY_B200_ECAPACITY — unless the count is void because copy-mode blocks were misread as signatures: every block in front of the
    // first copy-mode block is read correctly, and that includes the incompressible pair that started the episode, so the blocks that
    // fit (blk_off holds no more) are enough to find it; dec_seq_walk then recounts and judges the capacity again
    if (b == 0 && nb * T::BS > cap) st->error = 2;
    if (b >= nb || b >= maxblocks) return;
    const bool inc = T::consumed(ldsig(in + blk_off[b])) >= T::BS;  // codec.rs:98
    if (b == nb - 1) st->last_main_inc = inc ? 1u : 0u;
    if (inc && b > 0) {
        if (T::consumed(ldsig(in + blk_off[b - 1])) >= T::BS) atomicOr(&st->nonquiet, 1u);
    }
}

// ---- 2b. in-order boundary walk for streams with copy-mode blocks (codec.rs:88-100 with protection_state.rs) ---------------------
// One CTA; thread 0 carries (stream offset, block count, protection state) through the stream chunk by chunk.
//  * A chunk is JUMPED in O(1) from dec_chunk_walk's table when the automaton provably stays in encoded mode inside it (penalty 0
//    on entry, no two consecutive incompressible blocks inside, none across the entry seam): all its blocks are encoded blocks, the
//    table row gives the exit offset and block count, and dec_block_offsets fills in the per-block offsets afterwards in parallel.
//  * Any other chunk is WALKED block by block from a shared-memory copy (the whole CTA stages it), marking copy-mode blocks.
constexpr int SW_THREADS = 256;
constexpr int SW_BATCH = 24;                 // table rows staged at a time
enum : uint32_t { SW_ROWS = 0, SW_DIRTY = 1, SW_DONE = 2 };
__device__ __forceinline__ uint4 sw_load16(const uint8_t* __restrict__ in, uint64_t g, uint64_t n, bool al16) {
    uint4 v = make_uint4(0, 0, 0, 0);
    if (al16 && g + 16 <= n) return *reinterpret_cast<const uint4*>(in + g);
    uint8_t* vb = reinterpret_cast<uint8_t*>(&v);
    for (int k = 0; k < 16; ++k) if (g + k < n) vb[k] = in[g + k];
    return v;
}
// Jump over nb blocks that never enter copy mode (protection_state.rs:18-24,37-47): the penalty start halves on every 16th block.
__device__ __forceinline__ void sw_jump(Protection& ps, uint32_t nb, uint32_t last_inc) {
    const uint64_t k = (ps.counter + nb + 15) / 16 - (ps.counter + 15) / 16;
    if (ps.copy_penalty_start > 1) { const uint32_t sh = k > 8 ? 8u : (uint32_t)k; const uint32_t v = ps.copy_penalty_start >> sh; ps.copy_penalty_start = v ? v : 1u; }
    ps.counter += nb;
    ps.previous_incompressible = last_inc;
}
template <class T>
__global__ void __launch_bounds__(SW_THREADS) dec_seq_walk(const uint8_t* __restrict__ in, uint64_t n, uint64_t cap, uint32_t nchunks,
                                                           const uint32_t* __restrict__ res, const uint4* __restrict__ gres, uint32_t ngroups,
                                                           uint32_t* __restrict__ g_entry, uint64_t* __restrict__ g_blockbase,
                                                           uint32_t* __restrict__ c_entry, uint64_t* __restrict__ c_blockbase,
                                                           uint64_t* __restrict__ blk_off, uint64_t maxblocks, DecStatus* __restrict__ st) {
    if (!(st->nonquiet & 1u)) return;
    constexpr uint32_t SW_LOAD = T::CH + 16;            // + the signature bytes of a block starting at the chunk's last bytes, rounded up to 16
    __shared__ __align__(16) uint8_t win[2][SW_LOAD];   // the chunk being walked + the next one, prefetched during the walk
    __shared__ uint32_t rows[SW_BATCH * T::NCAND];
    __shared__ uint32_t s_cmd, s_chunk;
    const uint32_t tid = threadIdx.x;
    const bool al16 = (reinterpret_cast<uintptr_t>(in) & 15u) == 0;
    Protection ps; ps.init();
    uint64_t idx = 0, b = 0;        // meaningful in thread 0 only
    uint32_t g_next = 0;            // first group not entered yet (thread 0)
    uint32_t cb = 0, cb_valid = 0;  // staged rows: chunks [cb, cb + cb_valid)
    uint32_t wchunk0 = 0xFFFFFFFFu, wchunk1 = 0xFFFFFFFFu;   // which chunk each window buffer holds (uniform over the CTA)
    while (true) {
        if (tid == 0) {
            uint32_t cmd = SW_DONE, c = 0;
            while (n - idx >= T::MAXBLK) {
                c = (uint32_t)(idx / T::CH);
                const uint32_t e = (uint32_t)(idx - (uint64_t)c * T::CH) >> 1;          // < NCAND: a block is at most MAXBLK bytes
                if (c / GROUP == g_next) {
                    // entering a group of 64 chunks: jump over all of it if the automaton provably stays in encoded mode inside
                    const uint32_t g = g_next++;
                    const uint4 gr = gres[(size_t)g * T::NCAND + e];
                    if (ps.copy_penalty == 0 && gr.x != TERM && !(gr.z & 9u) && !(ps.previous_incompressible && (gr.z & 2u))) {
                        g_entry[g] = e; g_blockbase[g] = b;
                        sw_jump(ps, gr.y, (gr.z >> 2) & 1u);
                        b += gr.y;
                        idx = (uint64_t)(g + 1) * GROUP * T::CH + 2 * gr.x;
                        continue;
                    }
                    g_entry[g] = G_SKIP;                                             // chunk by chunk below
                }
                if (c < cb || c >= cb + cb_valid) { cmd = SW_ROWS; break; }
                const uint32_t r = rows[(c - cb) * T::NCAND + e];
                const uint32_t ex = r & 0xFFu, fl = row_x<T>(r);
                if (ps.copy_penalty == 0 && ex != TERM && !(fl & 1u) && !(ps.previous_incompressible && (fl & 2u))) {
                    const uint32_t nb = row_nb<T>(r);
                    c_entry[c] = e; c_blockbase[c] = b;
                    sw_jump(ps, nb, (fl >> 2) & 1u);
                    b += nb;
                    idx = (uint64_t)(c + 1) * T::CH + 2 * ex;
                } else { cmd = SW_DIRTY; break; }
            }
            s_cmd = cmd; s_chunk = c;
        }
        __syncthreads();
        const uint32_t cmd = s_cmd, c = s_chunk;
        if (cmd == SW_DONE) break;
        if (cmd == SW_ROWS) {
            cb = c; cb_valid = (nchunks - c < (uint32_t)SW_BATCH) ? nchunks - c : (uint32_t)SW_BATCH;
            for (uint32_t i = tid; i < cb_valid * T::NCAND; i += SW_THREADS) rows[i] = res[(size_t)cb * T::NCAND + i];
        } else {
            const uint64_t wbase = (uint64_t)c * T::CH;
            int cur = (wchunk0 == c) ? 0 : (wchunk1 == c) ? 1 : -1;
            if (cur < 0) {                                                   // not prefetched: the whole CTA stages it now
                cur = 0; wchunk0 = c;
                for (uint32_t i = tid * 16; i < SW_LOAD; i += SW_THREADS * 16) *reinterpret_cast<uint4*>(win[0] + i) = sw_load16(in, wbase + i, n, al16);
                __syncthreads();
            }
            if (tid >= 32 && c + 1 < nchunks) {                              // the others fetch the next chunk while thread 0 walks this one
                constexpr int PER = (SW_LOAD / 16 + (SW_THREADS - 32) - 1) / (SW_THREADS - 32);
                uint4 v[PER];
#pragma unroll
                for (int t = 0; t < PER; ++t) {
                    const uint32_t i = ((tid - 32) + t * (SW_THREADS - 32)) * 16;
                    v[t] = (i < SW_LOAD) ? sw_load16(in, wbase + T::CH + i, n, al16) : make_uint4(0, 0, 0, 0);
                }
#pragma unroll
                for (int t = 0; t < PER; ++t) {
                    const uint32_t i = ((tid - 32) + t * (SW_THREADS - 32)) * 16;
                    if (i < SW_LOAD) *reinterpret_cast<uint4*>(win[cur ^ 1] + i) = v[t];
                }
            }
            if (c + 1 < nchunks) { if (cur) wchunk0 = c + 1; else wchunk1 = c + 1; }
            if (tid == 0) {
                const uint32_t* w32 = reinterpret_cast<const uint32_t*>(win[cur]);
                const uint64_t wend = wbase + T::CH;
                c_entry[c] = TERM;                                           // dec_block_offsets leaves this chunk alone
                while (idx < wend && n - idx >= T::MAXBLK) {
                    if (ps.revert_to_copy()) {                               // codec.rs:89-92
                        if (b < maxblocks) blk_off[b] = idx | BLK_COPY;
                        ++b; idx += T::BS; ps.decay();
                    } else {
                        const uint32_t o = (uint32_t)(idx - wbase), sh = (o & 2u) * 8;
                        const uint32_t w0 = w32[o >> 2], w1 = w32[(o >> 2) + 1], w2 = w32[(o >> 2) + 2];
                        const uint64_t sig = (uint64_t)__funnelshift_r(w0, w1, sh) | ((uint64_t)__funnelshift_r(w1, w2, sh) << 32);
                        const uint32_t consumed = T::consumed(sig);
                        if (b < maxblocks) blk_off[b] = idx;
                        ++b; idx += consumed; ps.update(consumed >= T::BS);   // codec.rs:94-98
                    }
                }
            }
        }
        __syncthreads();
    }
    // the chunk in which the main loop ended (if it was not walked it has no block either) and everything behind it carry no blocks
    {
        __shared__ uint32_t s_first_free, s_gnext;
        if (tid == 0) { s_first_free = (uint32_t)(idx / T::CH); s_gnext = g_next; }
        __syncthreads();
        for (uint32_t c = s_first_free + tid; c < nchunks; c += SW_THREADS) c_entry[c] = TERM;
        for (uint32_t g = s_gnext + tid; g < ngroups; g += SW_THREADS) g_entry[g] = G_SKIP;
    }
    if (tid == 0) {
        st->main_blocks = b; st->tail_off = idx;
        st->ps_penalty = ps.copy_penalty; st->ps_start = ps.copy_penalty_start; st->ps_prev = ps.previous_incompressible;
        st->seq = 1;
        st->error = (b * T::BS > cap) ? 2u : 0u;     // the candidate walk's block count was void
        st->nonquiet &= ~1u;
    }
}

// ---- host side: workspace layout + launch sequence ---------------------------------------------------------------------------------
struct BoundsLayout { size_t status, res, gres, g_entry, g_blockbase, c_entry, c_blockbase, blk_off, total; uint64_t maxblocks; };

template <class T>
inline size_t bounds_layout(size_t nbytes, size_t cap, BoundsLayout* L) {
    const uint64_t nchunks = (nbytes + T::CH - 1) / T::CH;
    const uint64_t ngroups = (nchunks + GROUP - 1) / GROUP;
    const uint64_t minblk = T::SIG + (T::BS == 256 ? 128 : 0);     // smallest encoded block: Chameleon 8 + 64 * 2, Cheetah 8, Lion 6
    L->maxblocks = nbytes / minblk + 2;                            // what the stream can hold ...
    if (L->maxblocks > cap / T::BS + 2) L->maxblocks = cap / T::BS + 2;   // ... and what the output can take (more is a capacity error)
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    L->status = take(sizeof(DecStatus));
    L->res = take(nchunks * T::NCAND * sizeof(uint32_t));
    L->gres = take(ngroups * T::NCAND * sizeof(uint4));
    L->g_entry = take(ngroups * sizeof(uint32_t));
    L->g_blockbase = take(ngroups * sizeof(uint64_t));
    L->c_entry = take(nchunks * sizeof(uint32_t));
    L->c_blockbase = take(nchunks * sizeof(uint64_t));
    L->blk_off = take(L->maxblocks * sizeof(uint64_t));
    L->total = off;
    return off;
}

// Enqueues the boundary kernels. Afterwards (on the stream): st->main_blocks / tail_off / protection state, blk_off[0 .. main_blocks).
template <class T>
inline cudaError_t bounds_launch(const uint8_t* d_in, size_t nbytes, size_t cap, uint8_t* ws, const BoundsLayout& L, cudaStream_t stream, uint64_t* launches) {
    DecStatus* st = reinterpret_cast<DecStatus*>(ws + L.status);
    cudaError_t e = cudaMemsetAsync(st, 0, sizeof(DecStatus), stream);
    if (e != cudaSuccess) return e;
    const uint32_t nchunks = (uint32_t)((nbytes + T::CH - 1) / T::CH);
    const uint32_t ngroups = (nchunks + GROUP - 1) / GROUP;
    uint32_t* res = reinterpret_cast<uint32_t*>(ws + L.res);
    uint4* gres = reinterpret_cast<uint4*>(ws + L.gres);
    uint32_t* g_entry = reinterpret_cast<uint32_t*>(ws + L.g_entry);
    uint64_t* g_bb = reinterpret_cast<uint64_t*>(ws + L.g_blockbase);
    uint32_t* c_entry = reinterpret_cast<uint32_t*>(ws + L.c_entry);
    uint64_t* c_bb = reinterpret_cast<uint64_t*>(ws + L.c_blockbase);
    uint64_t* blk_off = reinterpret_cast<uint64_t*>(ws + L.blk_off);
    dec_chunk_walk<T><<<nchunks, 160, 0, stream>>>(d_in, nbytes, nchunks, res);
    dec_group_compose<T><<<ngroups, 160, 0, stream>>>(res, nchunks, gres);
    dec_top_walk<T><<<1, 32, 0, stream>>>(gres, ngroups, nbytes, g_entry, g_bb, st);
    dec_chunk_entries<T><<<(ngroups + 127) / 128, 128, 0, stream>>>(res, nchunks, g_entry, g_bb, ngroups, c_entry, c_bb, nullptr);
    dec_block_offsets<T><<<(nchunks + 127) / 128, 128, 0, stream>>>(d_in, nbytes, nchunks, c_entry, c_bb, blk_off, L.maxblocks, nullptr);
    dec_quiet_check<T><<<(unsigned)((L.maxblocks + 255) / 256), 256, 0, stream>>>(d_in, blk_off, L.maxblocks, st, cap);
    // streams with copy-mode blocks only (the three kernels return at once otherwise): in-order walk, then the entries of the chunks of
    // jumped groups and the offsets of the blocks of jumped chunks
    dec_seq_walk<T><<<1, SW_THREADS, 0, stream>>>(d_in, nbytes, cap, nchunks, res, gres, ngroups, g_entry, g_bb, c_entry, c_bb, blk_off, L.maxblocks, st);
    dec_chunk_entries<T><<<(ngroups + 127) / 128, 128, 0, stream>>>(res, nchunks, g_entry, g_bb, ngroups, c_entry, c_bb, st);
    dec_block_offsets<T><<<(nchunks + 127) / 128, 128, 0, stream>>>(d_in, nbytes, nchunks, c_entry, c_bb, blk_off, L.maxblocks, st);
    *launches += 9;
    return cudaGetLastError();
}

}  // namespace bounds
}  // namespace dns
