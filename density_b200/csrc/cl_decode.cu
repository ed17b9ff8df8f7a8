// cl_decode.cu — run-parallel Cheetah decode for sm_100a.
//
// Replaces /root/reference/src/algorithms/cheetah/cheetah.rs:67-103,152-185 (decode_plain / decode_map_a / decode_map_b /
// decode_predicted, decode_unit / decode_partial_unit) driven by /root/reference/src/codec/codec.rs:82-126, bit-exactly.
// The scheme (every stage with the table logic of cl_core.cuh) is checked on the CPU by tests/cl_model.cpp + tests/test_cl_model_cpu.py.
//
//  0. Boundaries        decode_bounds.cuh with 2-bit flags (an encoded block is 8 + 4*plain + 2*map bytes, a copy-mode block 128 raw bytes).
//  1. Unpack            one warp per block: flag planes, the 16-bit hash of every NOT-predicted quad (explicit for MAP_A / MAP_B,
//                       hash of the literal for PLAIN), literals and copy-mode blocks straight to the output.
//  2. Chunk-map values  (cheetah.rs:72-73,80,89-93) A decoder never compares values, so the MRU-2 state of a bucket can be run
//                       SYMBOLICALLY: one warp per run executes PLAIN = push literal / MAP_A = read slot 0 / MAP_B = read slot 1 + swap
//                       on lists whose slots are literals or "slot j of the list carried into the run"; a fold over the runs (one
//                       thread per bucket) makes every run's carried-in list concrete; reads that hit a carried-in slot are patched.
//  3. Predicted values  (cheetah.rs:98-103) pred[ctx] is written by every not-predicted quad at ctx = hash of the previous quad and
//                       read by predicted quads. The hash of a predicted quad comes out of the table, so the context of the quad
//                       after it is not known up front: ROUNDS. In a round every run walks its blocks in order from the context
//                       and the table snapshot the previous round's fold left for it (round 0: snapshot unknown -> a read of an
//                       entry the run has not written yet is UNKNOWN, and so is the context of the next quad, whose write is
//                       skipped), then the fold (one thread per context) recomputes the snapshots and a sweep the run-entry
//                       contexts. When nothing changed and nothing was unknown the round's values are the reference's (induction
//                       over the runs). Text settles in 5-8 rounds for any run count (tests/test_cl_model_cpu.py).
//  4. Tail              (codec.rs:102-123) the last < 136 stream bytes in order by one thread from the folded tables (scalar_codec.cu).
//
// Lion is NOT decoded here: its 5-deep move-to-front lists make the same iteration advance one run per round (a misplaced
// operation desynchronises a whole list; measured in tests/cl_model.cpp), so lion_decode stays on the in-order kernel.
#include <stdlib.h>
#include "common.cuh"
#include "encode_internal.cuh"
#include "decode_bounds.cuh"
#include "cl_core.cuh"

namespace dns {
namespace cheedec {

using bounds::DecStatus;
using bounds::BLK_COPY;
using bounds::ldu16;
using T = bounds::CheeT;
using namespace cld;

constexpr uint32_t CTX_PASS = 0xFFFFFFFEu;        // ctx_out of a run without encoded quads
constexpr int RP_WARPS = 4;                       // warps (runs) per CTA of the walk kernels
constexpr int MAX_ROUNDS = 40;

struct ClStatus {
    unsigned int changed, unknown, done, rounds;
    unsigned int final_ctx, gave_up, pad0, pad1;
};

__device__ __forceinline__ uint64_t run_step_begin(uint32_t r, uint32_t nruns, uint64_t nsteps) { return (uint64_t)r * nsteps / nruns; }

// ---- 1. unpack ------------------------------------------------------------------------------------------------------------------
// flags[b] = {predicted, MAP_A, MAP_B, active} bit per quad (LSB-first signature, read_signature.rs:11-16)
__global__ void cd_unpack(const uint8_t* __restrict__ in, const uint64_t* __restrict__ blk_off, const DecStatus* __restrict__ st,
                          uint4* __restrict__ flags, uint16_t* __restrict__ K, uint32_t* __restrict__ out) {
    if (st->error) return;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t nb = st->main_blocks;
    for (uint64_t b = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); b < nb; b += (uint64_t)gridDim.x * (blockDim.x >> 5)) {
        const unsigned long long o = blk_off[b];
        const uint8_t* p = in + (o & ~BLK_COPY);
        if (o & BLK_COPY) {                                        // codec.rs:89-92: 128 raw bytes
            out[b * 32 + lane] = ldu16(p + 4 * lane) | (ldu16(p + 4 * lane + 2) << 16);
            if (lane == 0) flags[b] = make_uint4(0, 0, 0, 0);
            continue;
        }
        const uint32_t slo = ldu16(p) | (ldu16(p + 2) << 16), shi = ldu16(p + 4) | (ldu16(p + 6) << 16);
        const uint32_t flag = ((lane < 16 ? slo : shi) >> (2 * (lane & 15))) & 3u;
        const uint32_t plain = __ballot_sync(0xFFFFFFFFu, flag == K_PLAIN);
        const uint32_t ma = __ballot_sync(0xFFFFFFFFu, flag == K_MAP_A), mb = __ballot_sync(0xFFFFFFFFu, flag == K_MAP_B);
        const uint8_t* q = p + 8 + 4 * __popc(plain & lanemask_lt()) + 2 * __popc((ma | mb) & lanemask_lt());
        uint32_t k = 0;
        if (flag == K_PLAIN) { const uint32_t v = ldu16(q) | (ldu16(q + 2) << 16); out[b * 32 + lane] = v; k = hash16(v); }   // cheetah.rs:68-70
        else if (flag != K_PRED) k = ldu16(q);                                                                                  // cheetah.rs:78,88
        K[b * 32 + lane] = (uint16_t)k;
        if (lane == 0) flags[b] = make_uint4(~(plain | ma | mb), ma, mb, 0xFFFFFFFFu);
    }
}

// ---- 2. chunk-map values ------------------------------------------------------------------------------------------------------------
// entry per (run, bucket): {a, b, meta, 0}; meta = epoch << 20 | tags (cl_core.cuh). The tables are zeroed per call: epoch 1 = touched.
__global__ void __launch_bounds__(RP_WARPS * 32)
cd_cmap_walk(const DecStatus* __restrict__ st, uint32_t nruns, const uint4* __restrict__ flags, const uint16_t* __restrict__ K,
             uint4* __restrict__ entC_all, uint32_t* __restrict__ out, uint2* __restrict__ usym /* per block: reads of carried-in slot 0 / slot 1 */) {
    if (st->error) return;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t r = blockIdx.x * RP_WARPS + (threadIdx.x >> 5);
    if (r >= nruns) return;
    const uint64_t nsteps = st->main_blocks;
    const uint64_t s0 = run_step_begin(r, nruns, nsteps), s1 = run_step_begin(r + 1, nruns, nsteps);
    uint4* __restrict__ entC = entC_all + (size_t)r * 65536;
    for (uint64_t s = s0; s < s1; ++s) {
        const uint4 fl = flags[s];
        const uint32_t member_mask = fl.w & ~fl.x;                 // encoded and not predicted
        if (member_mask == 0) { if (lane == 0) usym[s] = make_uint2(0, 0); continue; }
        const bool member = (member_mask >> lane) & 1u;
        const uint32_t kind = ((fl.y >> lane) & 1u) ? K_MAP_A : ((fl.z >> lane) & 1u) ? K_MAP_B : K_PLAIN;
        const uint32_t h = K[s * 32 + lane];
        uint32_t val = 0;
        uint4 e = make_uint4(0, 0, 0, 0);
        if (member) { e = __ldcg(&entC[h]); if (kind == K_PLAIN) val = out[s * 32 + lane]; }
        const uint32_t key = member ? h : 0x10000u + lane;
        const uint32_t grp = __match_any_sync(0xFFFFFFFFu, key);
        const uint32_t lower = grp & lanemask_lt();
        const uint32_t rank = __popc(lower);
        const int src = lower ? 31 - __clz(lower) : (int)lane;
        const uint32_t maxrank = __reduce_max_sync(0xFFFFFFFFu, member ? rank : 0u);
        List<2> L;
        if (meta_epoch(e.z) == 1u) { L.v[0] = e.x; L.v[1] = e.y; list_from_meta<2>(L, e.z); } else list_init<2>(L, nullptr);
        L.unk = 0;
        uint32_t sym = 0;                                          // 0: value known; j + 1: value = slot j of the carried-in list
        for (uint32_t rk = 0; rk <= maxrank; ++rk) {
            if (member && rank == rk) {
                if (kind == K_PLAIN) list_push<2>(L, val);                                   // cheetah.rs:72-73
                else {
                    const int sl = kind == K_MAP_A ? 0 : 1;
                    const uint32_t t = L.slot_tag(sl);
                    if (t == TAG_LIT) val = L.v[sl]; else sym = t;                               // cheetah.rs:80 / :90
                    if (sl == 1) list_mtf<2>(L, 1);                                              // cheetah.rs:92-93
                }
            }
            const uint32_t r0 = __shfl_sync(0xFFFFFFFFu, L.v[0], src), r1 = __shfl_sync(0xFFFFFFFFu, L.v[1], src), rt = __shfl_sync(0xFFFFFFFFu, L.tag, src);
            if (member && rank == rk + 1) { L.v[0] = r0; L.v[1] = r1; L.tag = rt; }
        }
        if (member && (grp & lanemask_gt()) == 0) entC[h] = make_uint4(L.v[0], L.v[1], list_meta<2>(L, 1u), 0u);
        if (member && kind != K_PLAIN && sym == 0) out[s * 32 + lane] = val;
        const uint32_t u0 = __ballot_sync(0xFFFFFFFFu, sym == 1), u1 = __ballot_sync(0xFFFFFFFFu, sym == 2);
        if (lane == 0) usym[s] = make_uint2(u0, u1);
        __syncwarp();
    }
}

// one thread per bucket: the list carried into every run, and the chunk map after the main loop (for the tail)
__global__ void cd_cmap_fold(const DecStatus* __restrict__ st, uint32_t nruns, const uint4* __restrict__ entC_all, uint2* __restrict__ cin,
                             uint32_t* __restrict__ chunk_a, uint32_t* __restrict__ chunk_b) {
    if (st->error) return;
    const uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= 65536) return;
    uint32_t c[2] = {0u, 0u};                                      // chunk map starts as (0, 0) (cheetah.rs:52)
    for (uint32_t r0 = 0; r0 < nruns; r0 += 8) {
        uint4 e[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) e[k] = (r0 + k < nruns) ? entC_all[(size_t)(r0 + k) * 65536 + h] : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (r0 + k >= nruns) break;
            cin[(size_t)(r0 + k) * 65536 + h] = make_uint2(c[0], c[1]);
            if (meta_epoch(e[k].z) == 1u) { List<2> L; L.v[0] = e[k].x; L.v[1] = e[k].y; list_from_meta<2>(L, e[k].z); list_carry<2>(c, L); }
        }
    }
    chunk_a[h] = c[0]; chunk_b[h] = c[1];
}

__device__ __forceinline__ uint32_t run_of_step(uint64_t s, uint32_t nruns, uint64_t nsteps) {
    uint32_t r = (uint32_t)((s * nruns) / nsteps);
    if (r >= nruns) r = nruns - 1;
    while (r + 1 < nruns && run_step_begin(r + 1, nruns, nsteps) <= s) ++r;
    while (r > 0 && run_step_begin(r, nruns, nsteps) > s) --r;
    return r;
}

// reads that hit a carried-in slot
__global__ void cd_cmap_resolve(const DecStatus* __restrict__ st, uint32_t nruns, const uint2* __restrict__ usym, const uint16_t* __restrict__ K,
                                const uint2* __restrict__ cin, uint32_t* __restrict__ out) {
    if (st->error) return;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t nsteps = st->main_blocks;
    for (uint64_t s = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); s < nsteps; s += (uint64_t)gridDim.x * (blockDim.x >> 5)) {
        const uint2 u = usym[s];
        if ((u.x | u.y) == 0) continue;
        const uint32_t r = run_of_step(s, nruns, nsteps);
        if (((u.x | u.y) >> lane) & 1u) {
            const uint2 c = cin[(size_t)r * 65536 + K[s * 32 + lane]];
            out[s * 32 + lane] = ((u.x >> lane) & 1u) ? c.x : c.y;
        }
    }
}

// ---- 3. predicted values --------------------------------------------------------------------------------------------------------------
// context of the first encoded quad of every run when the stream says it: the nearest earlier encoded block ends with a quad that is
// not predicted (its hash is in K); otherwise unknown until a round has produced it. last_hash starts as 0 (cheetah.rs:54).
__global__ void cd_ctx_init(const DecStatus* __restrict__ st, uint32_t nruns, const uint4* __restrict__ flags, const uint16_t* __restrict__ K,
                            uint32_t* __restrict__ ctx_in, uint32_t* __restrict__ dirty_cur, uint32_t* __restrict__ dirty_next, uint32_t* __restrict__ run_epoch,
                            ClStatus* __restrict__ cs) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r == 0) { cs->changed = 0; cs->unknown = 0; cs->done = st->error ? 1u : 0u; cs->rounds = 0; cs->final_ctx = 0; cs->gave_up = 0; cs->pad0 = 0; }
    if (r >= nruns || st->error) return;
    dirty_cur[r] = 1; dirty_next[r] = 0; run_epoch[r] = 0;      // round 0 walks every run
    uint64_t s = run_step_begin(r, nruns, st->main_blocks);
    uint32_t c = 0;
    while (s > 0) {
        --s;
        const uint4 fl = flags[s];
        if (fl.w == 0) continue;                                   // copy-mode block: touches nothing (codec.rs:89-92)
        c = (fl.x >> 31) ? H_UNKNOWN : (uint32_t)K[s * 32 + 31];
        break;
    }
    ctx_in[r] = c;
}

// entry per (run, context): {value, epoch << 20}. Only not-predicted quads write (a predicted quad would store back what it read,
// cheetah.rs:98-103), so an entry of the current epoch always holds a literal.
__global__ void __launch_bounds__(RP_WARPS * 32)
cd_pred_walk(const DecStatus* __restrict__ st, ClStatus* __restrict__ cs, uint32_t nruns, uint32_t round, const uint4* __restrict__ flags,
             const uint16_t* __restrict__ K, uint2* __restrict__ entP_all, const uint32_t* __restrict__ snap_all, const uint32_t* __restrict__ ctx_in,
             uint32_t* __restrict__ ctx_out, const uint32_t* __restrict__ dirty_cur, uint32_t* __restrict__ dirty_next, uint32_t* __restrict__ run_epoch,
             uint32_t* __restrict__ rbits_all, uint32_t* __restrict__ out) {
    if (cs->done) return;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t r = blockIdx.x * RP_WARPS + (threadIdx.x >> 5);
    if (r >= nruns) return;
    // A run is walked again only if something it depends on changed: its entry context, or a snapshot entry it read (cd_pred_fold
    // compares against the run's read set), or it met an unknown last time. Otherwise its values, its table entries (validated by
    // run_epoch) and its exit context stand.
    if (!dirty_cur[r]) return;
    uint32_t* __restrict__ rbits = rbits_all + (size_t)r * 2048;
    for (uint32_t i = lane; i < 2048; i += 32) rbits[i] = 0;
    __syncwarp();
    const uint32_t epoch = round + 1;
    const bool has_snap = round > 0 || r == 0;
    const uint64_t nsteps = st->main_blocks;
    const uint64_t s0 = run_step_begin(r, nruns, nsteps), s1 = run_step_begin(r + 1, nruns, nsteps);
    uint2* __restrict__ entP = entP_all + (size_t)r * 65536;
    const uint32_t* __restrict__ snap = snap_all + (size_t)r * 65536;
    uint32_t carry = ctx_in[r];
    bool any_active = false, unknown_seen = false;
    // software pipeline: the streaming loads of the next step are issued before this step's table work
    uint4 fl_n = make_uint4(0, 0, 0, 0); uint32_t k_n = 0, v_n = 0;
    if (s0 < s1) { fl_n = flags[s0]; k_n = K[s0 * 32 + lane]; v_n = out[s0 * 32 + lane]; }
    for (uint64_t s = s0; s < s1; ++s) {
        const uint4 fl = fl_n; const uint32_t kh = k_n; uint32_t v = v_n;
        if (s + 1 < s1) { fl_n = flags[s + 1]; k_n = K[(s + 1) * 32 + lane]; v_n = out[(s + 1) * 32 + lane]; }
        if (fl.w == 0) continue;                                   // copy-mode block
        any_active = true;
        const uint32_t P = fl.x;
        const bool pred = (P >> lane) & 1u;
        // my context: the hash of the previous quad. Known at once unless that quad is predicted.
        const uint32_t kprev = __shfl_up_sync(0xFFFFFFFFu, kh, 1);
        uint32_t ctx = lane == 0 ? carry : (((P >> (lane - 1)) & 1u) ? H_UNKNOWN : kprev);
        bool ctx_ready = lane == 0 || !((P >> (lane - 1)) & 1u);
        // predicted lanes whose context is known up front fetch their entry now (all in flight together)
        uint2 e = make_uint2(0, 0); uint32_t sv = 0;
        const bool pre = pred && ctx_ready && ctx != H_UNKNOWN;
        if (pre) { e = __ldcg(&entP[ctx]); if (has_snap) sv = __ldcg(&snap[ctx]); }
        uint32_t h = pred ? H_UNKNOWN : kh;                        // my own hash
        uint32_t todo = P;
        while (todo) {
            const int p = __ffs(todo) - 1;
            todo &= todo - 1;
            const uint32_t c = __shfl_sync(0xFFFFFFFFu, ctx, p);    // lane p's context is final by now
            uint32_t val = 0; bool unk = false;
            if (c == H_UNKNOWN) unk = true;
            else {
                // the latest earlier writer of this step with the same context (all earlier contexts are final)
                const uint32_t w = __ballot_sync(0xFFFFFFFFu, !pred && (int)lane < p && ctx == c);
                if (w) val = __shfl_sync(0xFFFFFFFFu, v, 31 - __clz(w));
                else {
                    uint2 ee = e; uint32_t ss = sv;
                    const bool had = __shfl_sync(0xFFFFFFFFu, (int)pre, p) != 0;
                    if (!had && (int)lane == p) { ee = __ldcg(&entP[c]); if (has_snap) ss = __ldcg(&snap[c]); }
                    uint32_t mv = ee.x, me = meta_epoch(ee.y), ms = ss;
                    mv = __shfl_sync(0xFFFFFFFFu, mv, p); me = __shfl_sync(0xFFFFFFFFu, me, p); ms = __shfl_sync(0xFFFFFFFFu, ms, p);
                    if (me == epoch) val = mv;                      // written earlier in this run
                    else if (has_snap) {                            // carried in (as of the previous round's fold): remember that I depend on it
                        val = ms;
                        if ((int)lane == p) atomicOr(&rbits[c >> 5], 1u << (c & 31));
                    } else unk = true;                              // round 0: nothing is known about what earlier runs left here
                }
            }
            const uint32_t hp = unk ? H_UNKNOWN : hash16(val);      // cheetah.rs:101
            unknown_seen |= unk;
            // a stretch of predicted quads right behind p reads the same entry as long as the hash maps the context onto itself
            uint32_t span = 1;
            if (!unk && hp == c) {
                const uint32_t rest = p == 31 ? 0u : ~(P >> (p + 1));
                const uint32_t follow = rest ? (uint32_t)(__ffs(rest) - 1) : (uint32_t)(31 - p);
                span += follow;
            }
            if ((int)lane >= p && (uint32_t)lane < (uint32_t)p + span) { v = val; h = hp; if ((int)lane > p) { ctx = c; ctx_ready = true; } }
            if (span > 1) { const uint32_t clr = ((span >= 32 ? 0xFFFFFFFFu : ((1u << span) - 1u)) << p); todo &= ~clr; }
            if ((uint32_t)lane == (uint32_t)p + span) { ctx = hp; ctx_ready = true; }
        }
        carry = __shfl_sync(0xFFFFFFFFu, h, 31);                    // cheetah.rs:102 (last_hash)
        // writers: pred[ctx] <- value (cheetah.rs:74,82,94); the last writer of a context inside the step leaves its value
        const bool writer = !pred && ctx != H_UNKNOWN;
        const uint32_t grp = __match_any_sync(0xFFFFFFFFu, writer ? ctx : 0x10000u + lane);
        if (writer && (grp & lanemask_gt()) == 0) entP[ctx] = make_uint2(v, epoch << META_EPOCH_SHIFT);
        if (pred && h != H_UNKNOWN) out[s * 32 + lane] = v;
        __syncwarp();
    }
    if (lane == 0) {
        ctx_out[r] = any_active ? carry : CTX_PASS;
        run_epoch[r] = epoch;
        if (unknown_seen) dirty_next[r] = 1;
    }
}

// one thread per context: snapshot of the table in front of every run (in place), and the table after the main loop (for the tail).
// A run whose snapshot changed at a context it read has to be walked again.
__global__ void cd_pred_fold(const DecStatus* __restrict__ st, ClStatus* __restrict__ cs, uint32_t nruns, uint32_t round, const uint2* __restrict__ entP_all,
                             uint32_t* __restrict__ snap, const uint32_t* __restrict__ run_epoch, const uint32_t* __restrict__ rbits_all,
                             uint32_t* __restrict__ dirty_next, uint32_t* __restrict__ pred_final) {
    if (cs->done) return;
    const uint32_t ctx = blockIdx.x * blockDim.x + threadIdx.x;
    if (ctx >= 65536) return;
    uint32_t c = 0;                                                // prediction table starts as 0 everywhere (cheetah.rs:53)
    for (uint32_t r0 = 0; r0 < nruns; r0 += 8) {
        uint2 e[8]; uint32_t so[8], ep[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const bool ok = r0 + k < nruns;
            e[k] = ok ? entP_all[(size_t)(r0 + k) * 65536 + ctx] : make_uint2(0, 0);
            so[k] = ok ? snap[(size_t)(r0 + k) * 65536 + ctx] : 0u;
            ep[k] = ok ? run_epoch[r0 + k] : 0u;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (r0 + k >= nruns) break;
            if (so[k] != c) {
                snap[(size_t)(r0 + k) * 65536 + ctx] = c;
                if (round > 0 && ((rbits_all[(size_t)(r0 + k) * 2048 + (ctx >> 5)] >> (ctx & 31)) & 1u)) dirty_next[r0 + k] = 1;
            }
            if (meta_epoch(e[k].y) == ep[k] && ep[k] != 0) c = e[k].x;
        }
    }
    pred_final[ctx] = c;
    (void)st;
}

// context sweep + verdict of the round (one thread)
__global__ void cd_round_end(ClStatus* __restrict__ cs, uint32_t nruns, uint32_t round, uint32_t* __restrict__ ctx_in, const uint32_t* __restrict__ ctx_out,
                             uint32_t* __restrict__ dirty_cur, uint32_t* __restrict__ dirty_next) {
    if (threadIdx.x || blockIdx.x || cs->done) return;
    uint32_t c = 0, ndirty = 0;
    for (uint32_t r = 0; r < nruns; ++r) {
        uint32_t d = dirty_next[r];
        if (round == 0 && r > 0) d = 1;                            // runs > 0 had no snapshot in round 0
        if (ctx_in[r] != c) { d = 1; ctx_in[r] = c; }
        dirty_cur[r] = d; dirty_next[r] = 0;
        ndirty += d;
        const uint32_t o = ctx_out[r];
        if (o != CTX_PASS) c = o;
    }
    cs->final_ctx = c;
    cs->rounds = round + 1;
    cs->pad0 += ndirty;                                            // diagnostic: run walks queued after round 0
    if (ndirty == 0) cs->done = 1;
}

// verdict for the caller: *d_fallback != 0 -> the in-order kernel (queued behind, gated on it) has to produce the result
__global__ void cd_finish(const DecStatus* __restrict__ st, ClStatus* __restrict__ cs, uint32_t* __restrict__ d_fallback, uint64_t* __restrict__ d_out_size) {
    const bool ok = cs->done && !st->error;
    if (!ok && !st->error) cs->gave_up = 1;
    *d_fallback = ok ? 0u : 1u;
    if (!ok && d_out_size) *d_out_size = 0;
}

}  // namespace cheedec

using namespace cheedec;

struct CheeDecLayout { bounds::BoundsLayout B; size_t cs, flags, K, usym, ctx_in, ctx_out, dirty, run_epoch, rbits, cin, snap0, total; };

static uint32_t cd_pick_runs(size_t nbytes, int num_sms) {
    const uint64_t maxblocks = nbytes / 8 + 2;
    (void)maxblocks;
    // a run should decode to >= 64 KiB (512 blocks); the stream is at most 8.5 bytes per block... use the stream size as a proxy
    uint64_t r = nbytes / (48u << 10);
    static const int per_sm = [] { const char* v = getenv("DENSITY_B200_DEC_RUNS_PER_SM"); const int k = v ? atoi(v) : 0; return (k >= 1 && k <= 32) ? k : 8; }();
    const uint64_t cap = (uint64_t)num_sms * per_sm;   // warps (runs) per SM: tuning knob, the result does not depend on it
    if (r > cap) r = cap;
    if (r < 1) r = 1;
    return (uint32_t)r;
}

static size_t cd_layout(size_t nbytes, size_t cap, uint32_t nruns, CheeDecLayout* L) {
    size_t off = bounds::bounds_layout<bounds::CheeT>(nbytes, cap, &L->B);
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const uint64_t mb = L->B.maxblocks;
    L->cs = take(sizeof(ClStatus));
    L->flags = take(mb * sizeof(uint4));
    L->K = take(mb * 32 * sizeof(uint16_t));
    L->usym = take(mb * sizeof(uint2));
    L->ctx_in = take((size_t)nruns * 4 + 64);
    L->ctx_out = take((size_t)nruns * 4 + 64);
    L->dirty = take((size_t)nruns * 8 + 64);
    L->run_epoch = take((size_t)nruns * 4 + 64);
    L->rbits = take((size_t)nruns * 2048 * sizeof(uint32_t));
    L->cin = take((size_t)nruns * 65536 * sizeof(uint2));
    L->snap0 = take((size_t)nruns * 65536 * sizeof(uint32_t));
    L->total = off;
    return off;
}

size_t chee_decode_workspace_bytes(size_t nbytes, size_t cap, int num_sms) { CheeDecLayout L; return cd_layout(nbytes, cap, cd_pick_runs(nbytes, num_sms), &L); }
// the per-run tables (zeroed at the start of every call): chunk map 16 B, prediction 8 B per run and key
size_t chee_decode_tables_bytes(size_t nbytes, int num_sms) { return (size_t)cd_pick_runs(nbytes, num_sms) * 65536 * (sizeof(uint4) + sizeof(uint2)); }

// Enqueues the parallel Cheetah decode. The block count of the main loop is only known on the device; the output needs
// cap >= main_blocks * 128 (checked on the device). *d_fallback != 0 afterwards: the caller's in-order kernel must run instead.
// `tail` = the scalar workspace (scalar_codec.cu layout): receives the folded tables for the tail loop.
cudaError_t chee_decode_parallel(const uint8_t* d_in, size_t nbytes, uint8_t* d_out, size_t cap, uint8_t* ws, uint8_t* tables, uint8_t* tail_ws,
                                 int num_sms, uint64_t* d_out_size, uint32_t* d_fallback, cudaStream_t stream, uint64_t* launches) {
    const uint32_t nruns = cd_pick_runs(nbytes, num_sms);
    CheeDecLayout L; cd_layout(nbytes, cap, nruns, &L);
    cudaError_t e = bounds::bounds_launch<bounds::CheeT>(d_in, nbytes, cap, ws, L.B, stream, launches);
    if (e != cudaSuccess) return e;
    const DecStatus* st = reinterpret_cast<const DecStatus*>(ws + L.B.status);
    ClStatus* cs = reinterpret_cast<ClStatus*>(ws + L.cs);
    const uint64_t* blk_off = reinterpret_cast<const uint64_t*>(ws + L.B.blk_off);
    uint4* flags = reinterpret_cast<uint4*>(ws + L.flags);
    uint16_t* K = reinterpret_cast<uint16_t*>(ws + L.K);
    uint2* usym = reinterpret_cast<uint2*>(ws + L.usym);
    uint32_t* ctx_in = reinterpret_cast<uint32_t*>(ws + L.ctx_in);
    uint32_t* ctx_out = reinterpret_cast<uint32_t*>(ws + L.ctx_out);
    uint2* cin = reinterpret_cast<uint2*>(ws + L.cin);
    uint32_t* snap = reinterpret_cast<uint32_t*>(ws + L.snap0);
    uint32_t* dirty_cur = reinterpret_cast<uint32_t*>(ws + L.dirty);
    uint32_t* dirty_next = dirty_cur + nruns;
    uint32_t* run_epoch = reinterpret_cast<uint32_t*>(ws + L.run_epoch);
    uint32_t* rbits = reinterpret_cast<uint32_t*>(ws + L.rbits);
    uint4* entC = reinterpret_cast<uint4*>(tables);
    uint2* entP = reinterpret_cast<uint2*>(tables + (size_t)nruns * 65536 * sizeof(uint4));
    uint32_t* out32 = reinterpret_cast<uint32_t*>(d_out);
    // scalar_codec.cu workspace: Status (256 B) + chunk_a + chunk_b + pred
    uint32_t* chunk_a = reinterpret_cast<uint32_t*>(tail_ws + 256);
    uint32_t* chunk_b = chunk_a + 65536;
    uint32_t* pred_final = chunk_a + 2 * 65536;
    e = cudaMemsetAsync(tables, 0, chee_decode_tables_bytes(nbytes, num_sms), stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(snap, 0, (size_t)nruns * 65536 * sizeof(uint32_t), stream);   // run 0's snapshot: the zero table
    if (e != cudaSuccess) return e;
    const int wide = num_sms * 8;
    const uint32_t run_ctas = (nruns + RP_WARPS - 1) / RP_WARPS;
    cd_unpack<<<wide, 256, 0, stream>>>(d_in, blk_off, st, flags, K, out32);
    cd_cmap_walk<<<run_ctas, RP_WARPS * 32, 0, stream>>>(st, nruns, flags, K, entC, out32, usym);
    cd_cmap_fold<<<65536 / 128, 128, 0, stream>>>(st, nruns, entC, cin, chunk_a, chunk_b);
    cd_cmap_resolve<<<wide, 256, 0, stream>>>(st, nruns, usym, K, cin, out32);
    cd_ctx_init<<<(nruns + 127) / 128, 128, 0, stream>>>(st, nruns, flags, K, ctx_in, dirty_cur, dirty_next, run_epoch, cs);
    *launches += 5;
    for (int round = 0; round < MAX_ROUNDS; ++round) {
        cd_pred_walk<<<run_ctas, RP_WARPS * 32, 0, stream>>>(st, cs, nruns, (uint32_t)round, flags, K, entP, snap, ctx_in, ctx_out, dirty_cur, dirty_next, run_epoch, rbits, out32);
        cd_pred_fold<<<65536 / 128, 128, 0, stream>>>(st, cs, nruns, (uint32_t)round, entP, snap, run_epoch, rbits, dirty_next, pred_final);
        cd_round_end<<<1, 32, 0, stream>>>(cs, nruns, (uint32_t)round, ctx_in, ctx_out, dirty_cur, dirty_next);
        *launches += 3;
    }
    cd_finish<<<1, 1, 0, stream>>>(st, cs, d_fallback, d_out_size);
    ++*launches;
    return cudaGetLastError();
}

// device addresses the tail kernel needs (scalar_codec.cu): boundary status + iteration status
const void* chee_decode_status_ptr(uint8_t* ws, size_t nbytes, size_t cap, int num_sms, const void** cl_status) {
    CheeDecLayout L; cd_layout(nbytes, cap, cd_pick_runs(nbytes, num_sms), &L);
    if (cl_status) *cl_status = ws + L.cs;
    return ws + L.B.status;
}

}  // namespace dns
