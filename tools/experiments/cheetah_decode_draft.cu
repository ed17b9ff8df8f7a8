// cheetah_decode_draft.cu — DRAFT of the run-parallel Cheetah decoder planned for the next round.
//
// STATUS: compiles for sm_100a (nvcc -c), has NEVER RUN. It is not part of the library build (density_b200/build.py) and nothing
// imports it. The algorithm of every pass is pinned by a Python model that reproduces the reference on dickens:
//   pass 1  chunk-map values          tools/proto_cheetah_decode_chunkmap.py   (transfer functions per run and bucket, fold, concrete pass)
//   pass 2  contexts by iteration     tools/proto_cheetah_decode_keyiter.py / _np.py  (10 rounds on 8.4 M quads, sparse after round 2)
//   why     PREDICTED quads never change the prediction table's values: tools/proto_cheetah_decode_hashchain.py
// Reference semantics: /root/reference/src/algorithms/cheetah/cheetah.rs:67-103 (decode_plain / decode_map_a / decode_map_b /
// decode_predicted) driven by /root/reference/src/codec/codec.rs:82-126.
//
// Input: the block list of the main loop (one entry per 128-byte block: stream offset, bit 63 = copy-mode block), produced by the
// boundary machinery of chameleon_decode.cu once it is templated on the signature width (not in this file).
// Output: 32 quads per block at out[32*b + lane].
//
// Geometry as in cheetah_encode.cu: a step = 32 quads = one block = one warp instruction; the stream of blocks is cut into R runs
// of whole tiles (128 blocks), one warp per run, run tables in global memory as epoch-tagged 16-byte entries.
#include "../../density_b200/csrc/common.cuh"

namespace dns {
namespace cheedec {

constexpr int TILE_B = 128;
constexpr int RP_WARPS = 4;
constexpr unsigned long long BLK_COPY = 1ull << 63;
constexpr uint32_t H_UNKNOWN = 0xFFFFFFFFu;      // hash estimate of a predicted quad that could not be evaluated yet

__device__ __forceinline__ uint32_t ldu16(const uint8_t* p) { return *reinterpret_cast<const uint16_t*>(p); }
__device__ __forceinline__ uint64_t run_block_begin(uint32_t r, uint32_t nruns, uint64_t ntiles) { return ((uint64_t)r * ntiles / nruns) * TILE_B; }

// ---- pass 0: unpack --------------------------------------------------------------------------------------------------------
// One warp per block. flag planes (f0 = bit 0, f1 = bit 1 of the 2-bit flag, LSB-first signature, read_signature.rs:11-16),
// `lit` = the quad (PLAIN), the 16-bit hash (MAP_A / MAP_B) or nothing (PREDICTED); copy-mode blocks go straight to the output.
__global__ void cd_unpack(const uint8_t* __restrict__ in, const unsigned long long* __restrict__ blk_off, uint64_t nblocks,
                          uint32_t* __restrict__ f0, uint32_t* __restrict__ f1, uint32_t* __restrict__ cpm /* bit per block */,
                          uint32_t* __restrict__ lit, uint32_t* __restrict__ out) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t b = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= nblocks) return;
    const unsigned long long o = blk_off[b];
    const uint8_t* p = in + (o & ~BLK_COPY);
    if (o & BLK_COPY) {                                        // codec.rs:89-92: 128 raw bytes
        out[b * 32 + lane] = ldu16(p + 4 * lane) | (ldu16(p + 4 * lane + 2) << 16);
        if (lane == 0) { f0[b] = 0; f1[b] = 0; atomicOr(&cpm[b >> 5], 1u << (b & 31)); }
        return;
    }
    const uint32_t slo = ldu16(p) | (ldu16(p + 2) << 16), shi = ldu16(p + 4) | (ldu16(p + 6) << 16);
    const uint32_t flag = ((lane < 16 ? slo : shi) >> (2 * (lane & 15))) & 3u;
    const uint32_t plain = __ballot_sync(0xFFFFFFFFu, flag == 0), maps = __ballot_sync(0xFFFFFFFFu, flag == 1 || flag == 2);
    const uint8_t* q = p + 8 + 4 * __popc(plain & lanemask_lt()) + 2 * __popc(maps & lanemask_lt());
    uint32_t v = 0;
    if (flag == 0) v = ldu16(q) | (ldu16(q + 2) << 16);        // cheetah.rs:68
    else if (flag != 3) v = ldu16(q);                          // cheetah.rs:78,88
    lit[b * 32 + lane] = v;
    const uint32_t p0 = __ballot_sync(0xFFFFFFFFu, flag & 1u), p1 = __ballot_sync(0xFFFFFFFFu, flag & 2u);
    if (lane == 0) { f0[b] = p0; f1[b] = p1; }
}

// ---- pass 1: chunk-map values ------------------------------------------------------------------------------------------------
// Symbolic MRU-2 state of a bucket inside a run: each of (a, b) is a literal or one of the carried-in symbols A0 / B0.
//   PLAIN(v): (a, b) <- (v, a);   MAP_B: (a, b) <- (b, a);   MAP_A: no change          (cheetah.rs:76-96)
// entry: {a value, b value, epoch << 4 | a tag << 2 | b tag, -}, tag 0 literal, 1 = A0, 2 = B0; an entry of another epoch = (A0, B0).
struct Sym { uint32_t av, bv, at, bt; };
__device__ __forceinline__ Sym sym_apply(Sym s, uint32_t flag, uint32_t v) {
    Sym r = s;
    if (flag == 0) { r.bv = s.av; r.bt = s.at; r.av = v; r.at = 0; }
    else if (flag == 2) { r.av = s.bv; r.at = s.bt; r.bv = s.av; r.bt = s.at; }
    return r;
}

// 1a: transfer function per run and bucket (only PLAIN and MAP_B quads modify the state; MAP_A and PREDICTED are skipped)
__global__ void __launch_bounds__(RP_WARPS * 32)
cd_cmap_transfer(const uint32_t* __restrict__ f0, const uint32_t* __restrict__ f1, const uint32_t* __restrict__ cpm, const uint32_t* __restrict__ lit,
                 uint64_t nblocks, uint32_t nruns, uint64_t ntiles, uint4* __restrict__ ent_all, uint32_t epoch) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t r = blockIdx.x * RP_WARPS + (threadIdx.x >> 5);
    if (r >= nruns) return;
    uint4* __restrict__ ent = ent_all + (size_t)r * 65536;
    const uint64_t b0 = run_block_begin(r, nruns, ntiles);
    uint64_t b1 = run_block_begin(r + 1, nruns, ntiles);
    if (b1 > nblocks) b1 = nblocks;
    for (uint64_t b = b0; b < b1; ++b) {
        if ((cpm[b >> 5] >> (b & 31)) & 1u) continue;
        const uint32_t flag = ((f0[b] >> lane) & 1u) | (((f1[b] >> lane) & 1u) << 1);
        const uint32_t v = lit[b * 32 + lane];
        const bool member = flag == 0 || flag == 2;
        const uint32_t h = flag == 0 ? prod_hash(hash_prod(v)) : v;              // PLAIN: hash of the literal (cheetah.rs:70); MAP: from the stream
        uint4 e = make_uint4(0, 0, 0, 0);
        if (member) e = __ldcg(&ent[h]);
        const uint32_t key = member ? h : 0x10000u + lane;
        const uint32_t grp = __match_any_sync(0xFFFFFFFFu, key);
        const uint32_t lower = grp & lanemask_lt();
        const uint32_t rank = __popc(lower);
        const int src = lower ? 31 - __clz(lower) : (int)lane;
        const uint32_t maxrank = __reduce_max_sync(0xFFFFFFFFu, member ? rank : 0u);
        Sym s;
        if ((e.z >> 4) == epoch) { s.av = e.x; s.bv = e.y; s.at = (e.z >> 2) & 3u; s.bt = e.z & 3u; }
        else { s.av = 0; s.bv = 0; s.at = 1; s.bt = 2; }                        // untouched: (A0, B0)
        Sym ns = s;
        for (uint32_t rk = 0; rk <= maxrank; ++rk) {
            if (member && rank == rk) ns = sym_apply(s, flag, v);
            const uint32_t r0 = __shfl_sync(0xFFFFFFFFu, ns.av, src), r1 = __shfl_sync(0xFFFFFFFFu, ns.bv, src);
            const uint32_t r2 = __shfl_sync(0xFFFFFFFFu, ns.at, src), r3 = __shfl_sync(0xFFFFFFFFu, ns.bt, src);
            if (member && rank == rk + 1) { s.av = r0; s.bv = r1; s.at = r2; s.bt = r3; }
        }
        if (member && (grp & lanemask_gt()) == 0) ent[h] = make_uint4(ns.av, ns.bv, (epoch << 4) | (ns.at << 2) | ns.bt, 0u);
        __syncwarp();
    }
}

// 1b: fold per bucket over the runs; cin[run][bucket] = the concrete (a, b) every run starts from
__global__ void cd_cmap_fold(uint32_t nruns, const uint4* __restrict__ ent_all, uint32_t epoch, uint2* __restrict__ cin) {
    const uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= 65536) return;
    uint32_t a0 = 0, b0 = 0;                                  // chunk map starts as (0, 0) (cheetah.rs:52)
    for (uint32_t r = 0; r < nruns; ++r) {
        cin[(size_t)r * 65536 + h] = make_uint2(a0, b0);
        const uint4 e = ent_all[(size_t)r * 65536 + h];
        if ((e.z >> 4) != epoch) continue;
        const uint32_t at = (e.z >> 2) & 3u, bt = e.z & 3u;
        const uint32_t na = at == 0 ? e.x : at == 1 ? a0 : b0;
        const uint32_t nb = bt == 0 ? e.y : bt == 1 ? a0 : b0;
        a0 = na; b0 = nb;
    }
}

// 1c: the concrete pass: values of the non-predicted quads -> out[], their hashes -> H[] (predicted quads: H_UNKNOWN)
// entry: {a, b, epoch, -}; first touch of a bucket in a run loads the carry-in
__global__ void __launch_bounds__(RP_WARPS * 32)
cd_cmap_values(const uint32_t* __restrict__ f0, const uint32_t* __restrict__ f1, const uint32_t* __restrict__ cpm, const uint32_t* __restrict__ lit,
               uint64_t nblocks, uint32_t nruns, uint64_t ntiles, uint4* __restrict__ ent_all, uint32_t epoch, const uint2* __restrict__ cin,
               uint32_t* __restrict__ out, uint32_t* __restrict__ H) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t r = blockIdx.x * RP_WARPS + (threadIdx.x >> 5);
    if (r >= nruns) return;
    uint4* __restrict__ ent = ent_all + (size_t)r * 65536;
    const uint2* __restrict__ ci = cin + (size_t)r * 65536;
    const uint64_t b0 = run_block_begin(r, nruns, ntiles);
    uint64_t b1 = run_block_begin(r + 1, nruns, ntiles);
    if (b1 > nblocks) b1 = nblocks;
    for (uint64_t b = b0; b < b1; ++b) {
        if ((cpm[b >> 5] >> (b & 31)) & 1u) continue;
        const uint32_t flag = ((f0[b] >> lane) & 1u) | (((f1[b] >> lane) & 1u) << 1);
        const uint32_t v = lit[b * 32 + lane];
        const bool member = flag != 3;                                           // everything but PREDICTED reads or writes the bucket
        const uint32_t h = flag == 0 ? prod_hash(hash_prod(v)) : v;
        uint4 e = make_uint4(0, 0, 0, 0);
        if (member) { e = __ldcg(&ent[h]); if (e.z != epoch) { const uint2 c = ci[h]; e.x = c.x; e.y = c.y; } }
        const uint32_t key = member ? h : 0x10000u + lane;
        const uint32_t grp = __match_any_sync(0xFFFFFFFFu, key);
        const uint32_t lower = grp & lanemask_lt();
        const uint32_t rank = __popc(lower);
        const int src = lower ? 31 - __clz(lower) : (int)lane;
        const uint32_t maxrank = __reduce_max_sync(0xFFFFFFFFu, member ? rank : 0u);
        uint32_t a = e.x, bb = e.y, na = a, nb = bb, val = 0;
        for (uint32_t rk = 0; rk <= maxrank; ++rk) {
            if (member && rank == rk) {
                if (flag == 0) { val = v; na = v; nb = a; }                      // cheetah.rs:67-75
                else if (flag == 1) { val = a; na = a; nb = bb; }                // :76-84
                else { val = bb; na = bb; nb = a; }                              // :86-96
            }
            const uint32_t ra = __shfl_sync(0xFFFFFFFFu, na, src), rb = __shfl_sync(0xFFFFFFFFu, nb, src);
            if (member && rank == rk + 1) { a = ra; bb = rb; na = a; nb = bb; }
        }
        if (member && (grp & lanemask_gt()) == 0) ent[h] = make_uint4(na, nb, epoch, 0u);
        if (member) out[b * 32 + lane] = val;
        H[b * 32 + lane] = member ? h : H_UNKNOWN;
        __syncwarp();
    }
}

// ---- pass 2: predicted values by global iteration on the contexts -------------------------------------------------------------------
// value(predicted quad i) = value of the latest non-predicted quad j < i with context_j == context_i (0 if none), context = hash of the
// previous ENCODED quad (copy-mode blocks are skipped by the chain). Hin[] holds the current hash estimates (exact for non-predicted
// quads, H_UNKNOWN or a guess for predicted ones); one round writes Hout[] for the predicted quads from a previous-occurrence pass
// under those estimates. Entry per (run, context): {value, epoch, 1 + index of the run's first read before any write, -}.
__global__ void cd_ctx0(const uint32_t* __restrict__ cpm, const uint32_t* __restrict__ Hin, uint32_t nruns, uint64_t ntiles, uint32_t* __restrict__ ctx0) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nruns) return;
    uint64_t b = run_block_begin(r, nruns, ntiles);
    uint32_t c = 0;                                            // last_hash starts as 0 (cheetah.rs:54)
    while (b > 0) {
        --b;
        if ((cpm[b >> 5] >> (b & 31)) & 1u) continue;
        c = Hin[b * 32 + 31];                                  // may be H_UNKNOWN: then the run's first context is unknown this round
        break;
    }
    ctx0[r] = c;
}

__global__ void __launch_bounds__(RP_WARPS * 32)
cd_pred_round(const uint32_t* __restrict__ f0, const uint32_t* __restrict__ f1, const uint32_t* __restrict__ cpm, const uint32_t* __restrict__ val,
              const uint32_t* __restrict__ Hin, uint64_t nblocks, uint32_t nruns, uint64_t ntiles, const uint32_t* __restrict__ ctx0,
              uint4* __restrict__ ent_all, uint32_t epoch, uint32_t* __restrict__ pval /* values of the predicted quads */,
              uint32_t* __restrict__ Hout, uint32_t* __restrict__ changed) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t r = blockIdx.x * RP_WARPS + (threadIdx.x >> 5);
    if (r >= nruns) return;
    uint4* __restrict__ ent = ent_all + (size_t)r * 65536;
    const uint64_t b0 = run_block_begin(r, nruns, ntiles);
    uint64_t b1 = run_block_begin(r + 1, nruns, ntiles);
    if (b1 > nblocks) b1 = nblocks;
    uint32_t last_h = ctx0[r];
    uint32_t any_change = 0;
    for (uint64_t b = b0; b < b1; ++b) {
        if ((cpm[b >> 5] >> (b & 31)) & 1u) continue;
        const uint32_t flag = ((f0[b] >> lane) & 1u) | (((f1[b] >> lane) & 1u) << 1);
        const bool pred = flag == 3;
        const uint32_t hi = Hin[b * 32 + lane];
        const uint32_t hp = __shfl_up_sync(0xFFFFFFFFu, hi, 1);
        const uint32_t ctx = lane ? hp : last_h;                                 // H_UNKNOWN propagates: quad inactive this round
        last_h = __shfl_sync(0xFFFFFFFFu, hi, 31);
        const bool active = ctx != H_UNKNOWN;
        const uint32_t v = pred ? 0u : val[b * 32 + lane];
        uint4 e = make_uint4(0, 0, 0, 0);
        if (active) e = __ldcg(&ent[ctx]);
        const uint32_t key = active ? ctx : 0x10000u + lane;
        const uint32_t grp = __match_any_sync(0xFFFFFFFFu, key);
        // latest WRITER (non-predicted lane) below me in my group
        const uint32_t writers = __ballot_sync(0xFFFFFFFFu, active && !pred);
        const uint32_t wl = grp & writers & lanemask_lt();
        const int wsrc = wl ? 31 - __clz(wl) : 0;
        const uint32_t wv = __shfl_sync(0xFFFFFFFFu, v, wsrc);
        const bool touched = e.y == epoch;                                       // the run has written this context before this step
        uint32_t got = 0; bool have = false;
        if (active && pred) {
            if (wl) { got = wv; have = true; }
            else if (touched) { got = e.x; have = true; }
        }
        // first reads of a context the run has not written yet are left to the fold: chain them per context through the entry
        // (z = 1 + index of the first such read; every later unresolved read of the context gets the same carried-in value, so
        //  the fold only needs the context -> value map and a second sweep: see cd_pred_fold)
        if (active && pred && !have) pval[b * 32 + lane] = 0x80000000u | ctx;    // marker: "carry-in of context ctx", patched by the fold sweep
        else if (active && pred) pval[b * 32 + lane] = got;
        // the group's last writer leaves its value
        const uint32_t gw = grp & writers;
        if (active && !pred && (gw & lanemask_gt()) == 0) ent[ctx] = make_uint4(v, epoch, 0u, 0u);
        // new hash estimate of my predicted quad (unknown if it waits for the fold)
        if (pred) {
            const uint32_t nh = (active && have) ? prod_hash(hash_prod(got)) : H_UNKNOWN;
            Hout[b * 32 + lane] = nh;
            any_change |= nh != hi;
        } else {
            Hout[b * 32 + lane] = hi;
        }
        __syncwarp();
    }
    if (__any_sync(0xFFFFFFFFu, any_change) && lane == 0) atomicOr(changed, 1u);
}

// fold: carry[run][ctx] = value the context holds when the run starts (0 at the stream start, cheetah.rs:53)
__global__ void cd_pred_fold(uint32_t nruns, const uint4* __restrict__ ent_all, uint32_t epoch, uint32_t* __restrict__ carry) {
    const uint32_t ctx = blockIdx.x * blockDim.x + threadIdx.x;
    if (ctx >= 65536) return;
    uint32_t c = 0;
    for (uint32_t r = 0; r < nruns; ++r) {
        carry[(size_t)r * 65536 + ctx] = c;
        const uint4 e = ent_all[(size_t)r * 65536 + ctx];
        if (e.y == epoch) c = e.x;
    }
}
// sweep: patch the reads that waited for the carry-in (they can only be reads BEFORE the run's first write of the context, so the
// carried-in value is the answer) and give their successors a hash estimate
__global__ void cd_pred_patch(const uint32_t* __restrict__ f0, const uint32_t* __restrict__ f1, uint64_t nblocks, uint32_t nruns, uint64_t ntiles,
                              const uint32_t* __restrict__ carry, uint32_t* __restrict__ pval, const uint32_t* __restrict__ Hin,
                              uint32_t* __restrict__ Hout, uint32_t* __restrict__ changed) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nblocks * 32) return;
    const uint64_t b = i >> 5; const uint32_t lane = i & 31;
    const uint32_t flag = ((f0[b] >> lane) & 1u) | (((f1[b] >> lane) & 1u) << 1);
    if (flag != 3) return;
    const uint32_t pv = pval[i];
    if (!(pv & 0x80000000u) || Hout[i] != H_UNKNOWN) return;   // NOTE(draft): a real value with bit 31 set collides with the marker —
                                                               // the final version needs a separate "waiting" bit plane
    // which run am I in? runs are whole tiles: binary search over run_block_begin would do; the draft recomputes it linearly
    uint32_t r = (uint32_t)(((b / TILE_B) * (uint64_t)nruns) / ntiles);
    while (r + 1 < nruns && run_block_begin(r + 1, nruns, ntiles) <= b) ++r;
    while (r > 0 && run_block_begin(r, nruns, ntiles) > b) --r;
    const uint32_t v = carry[(size_t)r * 65536 + (pv & 0xFFFFu)];
    pval[i] = v;
    const uint32_t nh = prod_hash(hash_prod(v));
    Hout[i] = nh;
    if (nh != Hin[i]) atomicOr(changed, 1u);
}

// after the last round: out[] of the predicted quads
__global__ void cd_pred_commit(const uint32_t* __restrict__ f0, const uint32_t* __restrict__ f1, const uint32_t* __restrict__ cpm, uint64_t nblocks,
                               const uint32_t* __restrict__ pval, uint32_t* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nblocks * 32) return;
    const uint64_t b = i >> 5; const uint32_t lane = i & 31;
    if ((cpm[b >> 5] >> (b & 31)) & 1u) return;
    const uint32_t flag = ((f0[b] >> lane) & 1u) | (((f1[b] >> lane) & 1u) << 1);
    if (flag == 3) out[i] = pval[i];
}

// Host driver sketch (rounds until `changed` stays 0; the verdict is read by the host between batches of rounds, or the round
// kernels are gated on it as the encoder's stages are):
//   cd_unpack -> cd_cmap_transfer -> cd_cmap_fold -> cd_cmap_values
//   repeat: cd_ctx0 -> cd_pred_round(Hin -> Hout) -> cd_pred_fold -> cd_pred_patch -> swap(Hin, Hout)   [fresh epoch per round]
//   cd_pred_commit; tail loop (codec.rs:102-123) in order from the folded tables as in chameleon_decode.cu::dec_tail
// Open points: (1) the 0x80000000 marker (see NOTE); (2) in cd_pred_round a writer whose own context is unknown is skipped this
// round (round 1 of the model) — correct at the fixed point, but the convergence check must also cover "no H_UNKNOWN left";
// (3) later rounds touch < 0.1 % of the quads (model): restrict them to the runs that saw a change.

}  // namespace cheedec
}  // namespace dns
