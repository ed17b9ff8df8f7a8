// cheetah_decode_draft.cu — DRAFT of the run-parallel Cheetah decoder planned for the next round.
//
// STATUS: compiles for sm_100a (nvcc -c), has NEVER RUN. It is not part of the library build (density_b200/build.py) and nothing
// imports it. The algorithm of every pass is pinned by a Python model that reproduces the reference on dickens:
//   pass 1  chunk-map values          tools/proto_cheetah_decode_chunkmap.py   (transfer functions per run and bucket, fold, concrete pass)
//   pass 2  predicted values          tools/proto_cheetah_decode_runs_gs.py   (in order inside a run + previous-round snapshots + unknown
//                                     propagation + carry sweep: 3-5 rounds on text / zeros / mixed for any run count)
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

// ---- pass 2: predicted values -----------------------------------------------------------------------------------------------------
// The variant to build (tools/proto_cheetah_decode_runs_gs.py: 4-5 rounds on text / zeros / mixed for any run count):
//   * a warp walks its run IN ORDER, so a chain of consecutive predicted quads advances through the whole run within one round;
//   * value(predicted quad) = value of the latest non-predicted quad with the same context: in-warp predecessor, else the run's
//     table (context written earlier in this run), else the SNAPSHOT of the context folded in the previous round, else UNKNOWN;
//   * unknown propagates and never writes: an unknown hash makes the next context unknown, a non-predicted quad with an unknown
//     context skips its table write this round;
//   * after the round: snapshots refolded from the runs' tables on top of the zero table (cheetah.rs:53), the first context of every
//     run = last hash of the run before it (cd_pred_carry, which also chases through runs of predicted quads only).
// Entry per (run, context): {value, epoch, -, -}. snap[run][ctx] = value the context holds when the run starts; snap_valid[run].
__global__ void __launch_bounds__(RP_WARPS * 32)
cd_pred_round(const uint32_t* __restrict__ f0, const uint32_t* __restrict__ f1, const uint32_t* __restrict__ cpm, const uint32_t* __restrict__ val,
              const uint32_t* __restrict__ Hn /* hashes of the non-predicted quads (pass 1) */, uint64_t nblocks, uint32_t nruns, uint64_t ntiles,
              const uint32_t* __restrict__ ctx0 /* first context per run or H_UNKNOWN */, const uint32_t* __restrict__ snap,
              const uint32_t* __restrict__ snap_valid, uint4* __restrict__ ent_all, uint32_t epoch,
              uint32_t* __restrict__ pval, uint32_t* __restrict__ pknown /* bit per quad: predicted value known */,
              uint32_t* __restrict__ run_last_h, uint32_t* __restrict__ changed) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t r = blockIdx.x * RP_WARPS + (threadIdx.x >> 5);
    if (r >= nruns) return;
    uint4* __restrict__ ent = ent_all + (size_t)r * 65536;
    const uint32_t* __restrict__ sn = snap + (size_t)r * 65536;
    const bool have_snap = snap_valid[r] != 0;
    const uint64_t b0 = run_block_begin(r, nruns, ntiles);
    uint64_t b1 = run_block_begin(r + 1, nruns, ntiles);
    if (b1 > nblocks) b1 = nblocks;
    uint32_t last_h = ctx0[r];
    uint32_t any_change = 0;
    for (uint64_t b = b0; b < b1; ++b) {
        if ((cpm[b >> 5] >> (b & 31)) & 1u) continue;                            // copy-mode block: the context chain skips it
        const uint32_t flag = ((f0[b] >> lane) & 1u) | (((f1[b] >> lane) & 1u) << 1);
        const bool pred = flag == 3;
        uint32_t h = pred ? H_UNKNOWN : Hn[b * 32 + lane];                       // my hash; predicted lanes fill it in below
        uint32_t v = pred ? 0u : val[b * 32 + lane];
        bool known = !pred;
        // predicted lanes in order: everything below the current one already has its hash (or is unknown for good this round)
        uint32_t todo = __ballot_sync(0xFFFFFFFFu, pred);
        while (todo) {
            const int p = __ffs(todo) - 1; todo &= todo - 1;
            const uint32_t hprev = __shfl_sync(0xFFFFFFFFu, h, p ? p - 1 : 0);
            const uint32_t cp = p ? hprev : last_h;                              // context of lane p
            // contexts of the lanes below p (all final by now)
            const uint32_t hb = __shfl_up_sync(0xFFFFFFFFu, h, 1);
            const uint32_t myc = lane ? hb : last_h;
            const uint32_t wr = __ballot_sync(0xFFFFFFFFu, (int)lane < p && !pred && myc != H_UNKNOWN && myc == cp);
            uint32_t got = 0; bool ok = false;
            if (cp != H_UNKNOWN) {
                if (wr) { got = __shfl_sync(0xFFFFFFFFu, v, 31 - __clz(wr)); ok = true; }
                else {
                    uint4 e = make_uint4(0, 0, 0, 0);
                    if ((int)lane == p) e = __ldcg(&ent[cp]);
                    const uint32_t ey = __shfl_sync(0xFFFFFFFFu, e.y, p), ex = __shfl_sync(0xFFFFFFFFu, e.x, p);
                    if (ey == epoch) { got = ex; ok = true; }
                    else if (have_snap) { uint32_t sv = 0; if ((int)lane == p) sv = sn[cp]; got = __shfl_sync(0xFFFFFFFFu, sv, p); ok = true; }
                }
            } else {
                (void)__shfl_sync(0xFFFFFFFFu, v, 0);                            // keep the shuffle pattern uniform
            }
            if ((int)lane == p) { known = ok; if (ok) { v = got; h = prod_hash(hash_prod(got)); } }
        }
        // table writes: the last non-predicted lane of every KNOWN context leaves its value (cheetah.rs:75,84,96: pred[last_hash] = quad)
        const uint32_t hb = __shfl_up_sync(0xFFFFFFFFu, h, 1);
        const uint32_t myc = lane ? hb : last_h;
        const bool writer = !pred && myc != H_UNKNOWN;
        const uint32_t key = writer ? myc : 0x10000u + lane;
        const uint32_t grp = __match_any_sync(0xFFFFFFFFu, key);
        if (writer && (grp & lanemask_gt()) == 0) ent[myc] = make_uint4(v, epoch, 0u, 0u);
        // results of the predicted lanes
        const uint32_t kn = __ballot_sync(0xFFFFFFFFu, pred && known);
        if (pred) {
            const uint32_t old = pval[b * 32 + lane];
            const bool was = (pknown[b] >> lane) & 1u;
            if (known) { pval[b * 32 + lane] = v; any_change |= (!was || old != v); }
            else any_change |= was;
        }
        if (lane == 0) pknown[b] = kn;
        last_h = __shfl_sync(0xFFFFFFFFu, h, 31);
        __syncwarp();
    }
    if (lane == 0) run_last_h[r] = last_h;
    if (__any_sync(0xFFFFFFFFu, any_change) && lane == 0) atomicOr(changed, 1u);
}

// fold: snap[run][ctx] for the NEXT round = value the context holds when the run starts (zero table at the stream start, cheetah.rs:53)
__global__ void cd_pred_fold(uint32_t nruns, const uint4* __restrict__ ent_all, uint32_t epoch, uint32_t* __restrict__ snap,
                             uint32_t* __restrict__ snap_valid) {
    const uint32_t ctx = blockIdx.x * blockDim.x + threadIdx.x;
    if (ctx >= 65536) return;
    uint32_t c = 0;
    for (uint32_t r = 0; r < nruns; ++r) {
        snap[(size_t)r * 65536 + ctx] = c;
        const uint4 e = ent_all[(size_t)r * 65536 + ctx];
        if (e.y == epoch) c = e.x;
    }
    if (ctx == 0) for (uint32_t r = 0; r < nruns; ++r) snap_valid[r] = 1;       // from round 2 on every run has a snapshot
}

// first context of every run for the next round: last hash of the run before it; a run of predicted quads only is a chase of
// (run length) links x -> hash(snap[x]) on a static table, cut short at a fixed point / cycle
// (one thread: R steps; `allpred[r]` = number of encoded quads of run r if they are all predicted, else 0 — from the flag planes)
__global__ void cd_pred_carry(uint32_t nruns, const uint32_t* __restrict__ run_last_h, const uint32_t* __restrict__ allpred,
                              const uint32_t* __restrict__ snap, uint32_t* __restrict__ ctx0) {
    if (threadIdx.x || blockIdx.x) return;
    uint32_t c = 0;                                            // last_hash starts as 0 (cheetah.rs:54)
    for (uint32_t r = 0; r < nruns; ++r) {
        ctx0[r] = c;
        uint32_t e = run_last_h[r];
        if (e == H_UNKNOWN && c != H_UNKNOWN && allpred[r]) {
            const uint32_t* __restrict__ sn = snap + (size_t)r * 65536;
            uint32_t x = c, steps = allpred[r], k = 0, tort = c, lam = 0, power = 1;      // Brent's cycle detection
            while (k < steps) {
                x = prod_hash(hash_prod(sn[x])); ++k; ++lam;
                if (x == tort) { const uint32_t left = steps - k; k = steps - (left % lam); lam = 0; tort = x; power = 1; continue; }
                if (lam == power) { tort = x; power <<= 1; lam = 0; }
            }
            e = x;
        }
        c = e;
    }
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
//   round 1: ctx0 = {0, H_UNKNOWN, ...}, snap_valid = {1, 0, ...} (run 0 starts from the zero table), pknown = 0
//   repeat: cd_pred_round -> cd_pred_fold -> cd_pred_carry     [fresh epoch per round; stop when `changed` stays 0 and no quad is unknown]
//   cd_pred_commit; tail loop (codec.rs:102-123) in order from the folded tables as in chameleon_decode.cu::dec_tail
// Open points: (1) the Brent loop above is untested (the model uses a dictionary of visited states); (2) later rounds touch < 0.1 % of
// the quads (model): restrict them to the runs whose snapshot or first context changed; (3) Lion: 5 values per context
// (tools/proto_lion_decode_full.py), same structure.

}  // namespace cheedec
}  // namespace dns
