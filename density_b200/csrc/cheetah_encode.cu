// cheetah_encode.cu — run-parallel Cheetah encode for sm_100a.
//
// Replaces /root/reference/src/algorithms/cheetah/cheetah.rs:121-150 (encode_quad) driven by
// /root/reference/src/codec/codec.rs:34-80, bit-exactly. Decomposition (validated against the oracle by
// tools/proto_cheetah_runs.py):
//
//  1. PREDICTED_i <=> quad_i == the quad that followed the previous occurrence of the same CONTEXT, where the context is the
//     hash of the previous *encoded* quad (0 at the stream start) and the prediction table starts as "0 everywhere"
//     (cheetah.rs:53,125,144,148: pred[ctx] always ends up holding the last quad seen in that context).
//  2. On the subsequence of NON-predicted quads each hash bucket keeps an MRU-2 of values (a, b) (cheetah.rs:127-143):
//        v == a -> MAP_A, nothing changes;  v == b -> MAP_B, (a, b) <- (v, a);  else PLAIN, (a, b) <- (v, a).
//     Values are compared as the 16-bit in-bucket fingerprints of common.cuh.
//  3. Copy-mode blocks (codec/protection_state.rs) touch neither table nor the context chain; the copy map is the fixed point
//     of  M -> automaton(incompressible bits under M)  (prot_iterate, shared with the Chameleon encoder). Cheetah needs the
//     iteration on every input: a cold dictionary makes the first blocks incompressible (11 copied blocks on dickens).
//
// Parallelisation: the stream is cut into R contiguous runs (thousands), ONE WARP PER RUN walks its run block by block
// (a Cheetah block is 32 quads = one quad per lane), with the run's tables in global memory (L2/HBM resident; 768 KiB do not fit
// an SM). In-warp predecessors come from __match_any_sync; what a run cannot know — the tables carried in from earlier runs —
// is left "unresolved": per run and context at most one PREDICTED decision, per run and bucket at most two map decisions (the
// first touch, and the first access that differs from it). One fold kernel per table then walks the runs in order per
// context / bucket, resolves those accesses and carries the state on. Pass P (predictions) must be completely resolved before
// pass C (chunk map) starts, because only non-predicted quads take part in it.
#include "common.cuh"
#include "encode_internal.cuh"

namespace dns {
namespace chee {

constexpr int TILE_B = 128;                 // blocks per tile (4096 quads = 16 KiB), the unit of run geometry and of the emit grid
constexpr uint32_t FP_INVALID = 0x10000u;   // a fingerprint value that matches nothing (bucket h != 0 initially "holds quad 0")

__device__ __forceinline__ uint64_t run_block_begin(uint32_t r, uint32_t nruns, uint64_t ntiles) { return ((uint64_t)r * ntiles / nruns) * TILE_B; }

// context of the first encoded quad of every run: hash of the last quad of the last encoded block before the run (0 if none)
__global__ void chee_ctx0(const uint32_t* __restrict__ in, uint64_t nquads, const uint8_t* __restrict__ copymap, uint32_t nruns, uint64_t ntiles,
                          const Status* __restrict__ gate, uint32_t* __restrict__ ctx0) {
    if (gate && !(gate->nonquiet && !gate->converged)) return;
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nruns) return;
    uint64_t b = run_block_begin(r, nruns, ntiles);
    uint32_t c = 0;
    while (b > 0) {
        --b;
        if (copymap && copymap[b]) continue;
        const uint64_t qi = b * 32 + 31;          // blocks before a run start are full blocks
        if (qi < nquads) c = prod_hash(hash_prod(in[qi]));
        break;
    }
    ctx0[r] = c;
}

constexpr int RP_WARPS = 4;   // warps (runs) per CTA

// ---- pass P: PREDICTED flags ----------------------------------------------------------------------------------------------
// Table entry per (run, context): {last quad seen, epoch, 1 + index of the run's first (unresolved) access, 0}. An entry whose epoch
// differs from the current round's is "untouched": no memset between rounds / calls (the workspace is zeroed once, epochs start at 1).
__global__ void __launch_bounds__(RP_WARPS * 32)
chee_pass_p(const uint32_t* __restrict__ in, uint64_t nquads, uint64_t nblocks, const uint8_t* __restrict__ copymap, uint32_t nruns, uint64_t ntiles,
            const Status* __restrict__ gate, const uint32_t* __restrict__ ctx0, uint4* __restrict__ entP_all, uint32_t epoch,
            uint32_t* __restrict__ Pbits) {
    if (gate && !(gate->nonquiet && !gate->converged)) return;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t r = blockIdx.x * RP_WARPS + (threadIdx.x >> 5);
    if (r >= nruns) return;
    uint4* __restrict__ entP = entP_all + (size_t)r * 65536;
    const uint64_t b0 = run_block_begin(r, nruns, ntiles);
    uint64_t b1 = run_block_begin(r + 1, nruns, ntiles);
    if (b1 > nblocks) b1 = nblocks;
    uint32_t last_h = ctx0[r];
    uint32_t qn = (b0 < b1 && b0 * 32 + lane < nquads) ? ld_stream_u32(in + b0 * 32 + lane) : 0u;   // software prefetch, one block ahead
    for (uint64_t b = b0; b < b1; ++b) {
        const uint32_t q = qn;
        if (b + 1 < b1) qn = ((b + 1) * 32 + lane < nquads) ? ld_stream_u32(in + (b + 1) * 32 + lane) : 0u;
        if (copymap && copymap[b]) { if (lane == 0) Pbits[b] = 0; continue; }
        const uint64_t q0 = b * 32;
        const uint32_t nq = (q0 >= nquads) ? 0u : (uint32_t)((nquads - q0 < 32) ? (nquads - q0) : 32);
        if (nq == 0) { if (lane == 0) Pbits[b] = 0; continue; }
        const bool active = lane < nq;
        const uint32_t h = prod_hash(hash_prod(q));
        const uint32_t hp = __shfl_up_sync(0xFFFFFFFFu, h, 1);
        const uint32_t ctx = lane ? hp : last_h;
        last_h = __shfl_sync(0xFFFFFFFFu, h, nq - 1);
        uint4 e = make_uint4(0, 0, 0, 0);
        if (active) e = __ldcg(&entP[ctx]);                           // issued for every lane up front: one 16-byte access
        const uint32_t key = active ? ctx : 0x10000u + lane;
        const uint32_t grp = __match_any_sync(0xFFFFFFFFu, key);
        const uint32_t lower = grp & lanemask_lt();
        const int src = lower ? 31 - __clz(lower) : 0;
        const uint32_t qprev = __shfl_sync(0xFFFFFFFFu, q, src);
        const bool touched = e.y == epoch;
        const bool known = lower != 0 || touched;
        const uint32_t pv = lower ? qprev : e.x;
        const bool P = active && known && pv == q;
        // the group's first lane owns the "first access" slot, its last lane leaves the value (cheetah.rs:144; :148 is implied)
        uint32_t u1 = touched ? e.z : 0u;
        if (active && !lower && !touched) u1 = (uint32_t)(q0 + lane) + 1u;
        const uint32_t u1g = __shfl_sync(0xFFFFFFFFu, u1, __ffs(grp) - 1);
        if (active && (grp & lanemask_gt()) == 0) entP[ctx] = make_uint4(q, epoch, u1g, 0u);
        const uint32_t pm = __ballot_sync(0xFFFFFFFFu, P);
        if (lane == 0) Pbits[b] = pm;
        __syncwarp();
    }
}

// walk the runs in order per context: resolve each run's first access from the carried-in value, carry the run's last value on
__global__ void chee_fold_p(const uint32_t* __restrict__ in, uint32_t nruns, const Status* __restrict__ gate, const uint4* __restrict__ entP_all,
                            uint32_t epoch, uint32_t* __restrict__ Pbits) {
    if (gate && !(gate->nonquiet && !gate->converged)) return;
    const uint32_t ctx = blockIdx.x * blockDim.x + threadIdx.x;
    if (ctx >= 65536) return;
    uint32_t c = 0;                                                   // prediction table starts as 0 everywhere (cheetah.rs:53)
    for (uint32_t r0 = 0; r0 < nruns; r0 += 8) {
        uint4 e[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) e[k] = (r0 + k < nruns) ? entP_all[(size_t)(r0 + k) * 65536 + ctx] : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (e[k].y != epoch) continue;
            const uint32_t i = e[k].z - 1;
            if (in[i] == c) atomicOr(&Pbits[i >> 5], 1u << (i & 31));
            c = e[k].x;
        }
    }
}

// ---- pass C: chunk map (MRU-2) on the non-predicted quads -----------------------------------------------------------------
// Table entry per (run, bucket): {a | b << 16, epoch << 2 | T, 1 + index of the first access, 1 + index of the first access that differs}
// T: 0 untouched, 1 a known, 2 a and b known.
__global__ void __launch_bounds__(RP_WARPS * 32)
chee_pass_c(const uint32_t* __restrict__ in, uint64_t nquads, uint64_t nblocks, const uint8_t* __restrict__ copymap, uint32_t nruns, uint64_t ntiles,
            const Status* __restrict__ gate, const uint32_t* __restrict__ Pbits, uint4* __restrict__ entC_all, uint32_t epoch,
            uint32_t* __restrict__ Abits, uint32_t* __restrict__ Bbits) {
    if (gate && !(gate->nonquiet && !gate->converged)) return;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t r = blockIdx.x * RP_WARPS + (threadIdx.x >> 5);
    if (r >= nruns) return;
    uint4* __restrict__ entC = entC_all + (size_t)r * 65536;
    const uint64_t b0 = run_block_begin(r, nruns, ntiles);
    uint64_t b1 = run_block_begin(r + 1, nruns, ntiles);
    if (b1 > nblocks) b1 = nblocks;
    uint32_t qn = (b0 < b1 && b0 * 32 + lane < nquads) ? ld_stream_u32(in + b0 * 32 + lane) : 0u;
    uint32_t pmn = (b0 < b1) ? __ldcg(&Pbits[b0]) : 0u;
    for (uint64_t b = b0; b < b1; ++b) {
        const uint32_t q = qn, pm = pmn;
        if (b + 1 < b1) { qn = ((b + 1) * 32 + lane < nquads) ? ld_stream_u32(in + (b + 1) * 32 + lane) : 0u; pmn = __ldcg(&Pbits[b + 1]); }
        if (copymap && copymap[b]) { if (lane == 0) { Abits[b] = 0; Bbits[b] = 0; } continue; }
        const uint64_t q0 = b * 32;
        const uint32_t nq = (q0 >= nquads) ? 0u : (uint32_t)((nquads - q0 < 32) ? (nquads - q0) : 32);
        if (nq == 0) { if (lane == 0) { Abits[b] = 0; Bbits[b] = 0; } continue; }
        const bool member = lane < nq && !((pm >> lane) & 1u);
        const uint32_t p = hash_prod(q);
        const uint32_t h = prod_hash(p), v = prod_fp(p, q);
        uint4 e = make_uint4(0, 0, 0, 0);
        if (member) e = __ldcg(&entC[h]);
        const uint32_t key = member ? h : 0x10000u + lane;
        const uint32_t grp = __match_any_sync(0xFFFFFFFFu, key);
        const uint32_t lower = grp & lanemask_lt();
        const uint32_t rank = __popc(lower);
        const int src = lower ? 31 - __clz(lower) : (int)lane;
        const uint32_t maxrank = __reduce_max_sync(0xFFFFFFFFu, member ? rank : 0u);
        // state before my access: the first lane of a group takes it from the table, the others from their in-warp predecessor
        const bool fresh = (e.y >> 2) == epoch;
        uint32_t a = fresh ? (e.x & 0xFFFFu) : 0u, bb = fresh ? (e.x >> 16) : 0u, T = fresh ? (e.y & 3u) : 0u;
        uint32_t u1 = fresh ? e.z : 0u, u2 = fresh ? e.w : 0u;
        uint32_t code = 0;          // 0 plain (or not yet decidable), 1 MAP_A, 2 MAP_B
        uint32_t na = 0, nb = 0, nT = 0, nu1 = 0, nu2 = 0;
        for (uint32_t rk = 0; rk <= maxrank; ++rk) {
            if (member && rank == rk) {
                nu1 = u1; nu2 = u2;
                if (T == 0) {                       // first touch of the bucket in this run: decided later from the carry-in
                    nu1 = (uint32_t)(q0 + lane) + 1u; na = v; nb = 0; nT = 1;
                } else if (T == 1) {
                    if (v == a) { code = 1; na = a; nb = 0; nT = 1; }
                    else { nu2 = (uint32_t)(q0 + lane) + 1u; na = v; nb = a; nT = 2; }        // MAP_B iff v == (unknown) b: decided later
                } else {
                    if (v == a) { code = 1; na = a; nb = bb; }
                    else { code = (v == bb) ? 2u : 0u; na = v; nb = a; }                       // cheetah.rs:137-142
                    nT = 2;
                }
            }
            const uint32_t ra = __shfl_sync(0xFFFFFFFFu, na, src), rb = __shfl_sync(0xFFFFFFFFu, nb, src), rT = __shfl_sync(0xFFFFFFFFu, nT, src);
            const uint32_t r1 = __shfl_sync(0xFFFFFFFFu, nu1, src), r2 = __shfl_sync(0xFFFFFFFFu, nu2, src);
            if (member && rank == rk + 1) { a = ra; bb = rb; T = rT; u1 = r1; u2 = r2; }
        }
        if (member && (grp & lanemask_gt()) == 0) entC[h] = make_uint4(na | (nb << 16), (epoch << 2) | nT, nu1, nu2);
        const uint32_t am = __ballot_sync(0xFFFFFFFFu, code == 1), bm = __ballot_sync(0xFFFFFFFFu, code == 2);
        if (lane == 0) { Abits[b] = am; Bbits[b] = bm; }
        __syncwarp();
    }
}

// walk the runs in order per bucket: resolve the (at most two) undecided accesses of each run, carry (a, b) on
__global__ void chee_fold_c(const uint32_t* __restrict__ in, uint32_t nruns, const Status* __restrict__ gate, const uint4* __restrict__ entC_all,
                            uint32_t epoch, uint32_t* __restrict__ Abits, uint32_t* __restrict__ Bbits) {
    if (gate && !(gate->nonquiet && !gate->converged)) return;
    const uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= 65536) return;
    // chunk map starts as (quad 0, quad 0) (cheetah.rs:52): only bucket 0 can ever match that
    uint32_t a0 = h ? FP_INVALID : 0u, b0 = a0;
    for (uint32_t r0 = 0; r0 < nruns; r0 += 8) {
        uint4 e[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) e[k] = (r0 + k < nruns) ? entC_all[(size_t)(r0 + k) * 65536 + h] : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if ((e[k].y >> 2) != epoch) continue;
            const uint32_t T = e[k].y & 3u;
            const uint32_t i1 = e[k].z - 1;
            const uint32_t q1 = in[i1];
            const uint32_t v1 = prod_fp(hash_prod(q1), q1);
            uint32_t bafter;
            if (v1 == a0) { atomicOr(&Abits[i1 >> 5], 1u << (i1 & 31)); bafter = b0; }
            else { if (v1 == b0) atomicOr(&Bbits[i1 >> 5], 1u << (i1 & 31)); bafter = a0; }
            if (T == 2) {
                const uint32_t i2 = e[k].w - 1;
                const uint32_t q2 = in[i2];
                const uint32_t v2 = prod_fp(hash_prod(q2), q2);
                if (v2 == bafter) atomicOr(&Bbits[i2 >> 5], 1u << (i2 & 31));
                a0 = e[k].x & 0xFFFFu; b0 = e[k].x >> 16;
            } else if (v1 != a0) {                  // the run accessed the bucket with one value only
                b0 = a0; a0 = v1;
            }
        }
    }
}

// ---- sizes + incompressible bits ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t chee_block_bytes(uint64_t b, uint64_t nbytes, uint32_t pm, uint32_t am, uint32_t bm, bool copied) {
    const uint64_t boff = b * 128;
    const uint32_t blen = (uint32_t)((nbytes - boff < 128) ? (nbytes - boff) : 128);
    if (copied) return blen;
    const uint32_t nq = blen >> 2;
    const uint32_t act = nq >= 32 ? 0xFFFFFFFFu : ((1u << nq) - 1u);
    const uint32_t maps = (am | bm) & act;
    const uint32_t plain = act & ~(pm | am | bm);
    return 8 + 4 * __popc(plain) + 2 * __popc(maps) + (blen & 3u);
}

// one warp per tile of 128 blocks (4 per lane): tile byte counts; refreshes the incompressible bits of the encoded blocks
__global__ void chee_tile_sizes(const uint32_t* __restrict__ Pbits, const uint32_t* __restrict__ Abits, const uint32_t* __restrict__ Bbits,
                                const uint8_t* __restrict__ copymap, uint64_t nbytes, uint64_t nblocks, uint32_t ntiles, int final_pass,
                                const Status* __restrict__ st, uint8_t* __restrict__ inc, uint32_t* __restrict__ tile_bytes) {
    if (!final_pass && !(st->nonquiet && !st->converged)) return;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (tile >= ntiles) return;
    uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint64_t b = (uint64_t)tile * TILE_B + k * 32 + lane;
        if (b < nblocks) {
            const bool copied = copymap && copymap[b];
            const uint32_t sz = chee_block_bytes(b, nbytes, Pbits[b], Abits[b], Bbits[b], copied);
            sum += sz;
            if (!final_pass && !copied) inc[b] = (nbytes - b * 128 >= 128) && sz >= 128;     // codec.rs:68
        }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, d);
    if (lane == 0) tile_bytes[tile] = sum;
}

// ---- emit: one CTA (256 threads) per tile of 128 blocks, warp w handles blocks w, w+8, ... ----------------------------------------
__device__ __forceinline__ uint32_t part1by1(uint32_t x) {   // spread the low 16 bits to the even bit positions
    x &= 0xFFFFu; x = (x | (x << 8)) & 0x00FF00FFu; x = (x | (x << 4)) & 0x0F0F0F0Fu; x = (x | (x << 2)) & 0x33333333u; x = (x | (x << 1)) & 0x55555555u;
    return x;
}
__global__ void __launch_bounds__(256)
chee_emit(const uint32_t* __restrict__ in, uint64_t nbytes, uint64_t nblocks, const uint32_t* __restrict__ Pbits, const uint32_t* __restrict__ Abits,
          const uint32_t* __restrict__ Bbits, const uint8_t* __restrict__ copymap, const Status* __restrict__ status,
          const uint32_t* __restrict__ tile_local, const uint64_t* __restrict__ group_off, uint32_t scan_group, uint8_t* __restrict__ out) {
    if (status->error || !status->converged) return;
    __shared__ uint32_t s_off[TILE_B + 1];
    __shared__ uint32_t s_wsum[4];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tile = blockIdx.x;
    const uint64_t tile_off = group_off[tile / scan_group] + tile_local[tile];
    const uint64_t nquads = nbytes / 4;
    if (tid < TILE_B) {
        const uint64_t b = (uint64_t)tile * TILE_B + tid;
        uint32_t sz = 0;
        if (b < nblocks) sz = chee_block_bytes(b, nbytes, Pbits[b], Abits[b], Bbits[b], copymap && copymap[b]);
        uint32_t incl = sz;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= (uint32_t)d) incl += u; }
        if (lane == 31) s_wsum[warp] = incl;
        s_off[tid + 1] = incl;
    }
    __syncthreads();
    if (tid < TILE_B && warp > 0) { uint32_t add = 0; for (uint32_t w = 0; w < warp; ++w) add += s_wsum[w]; s_off[tid + 1] += add; }
    if (tid == 0) s_off[0] = 0;
    __syncthreads();
    const uint8_t* in_b = reinterpret_cast<const uint8_t*>(in);
    for (uint32_t bl = warp; bl < TILE_B; bl += 8) {
        const uint64_t b = (uint64_t)tile * TILE_B + bl;
        if (b >= nblocks) break;
        const uint64_t boff = b * 128;
        const uint32_t blen = (uint32_t)((nbytes - boff < 128) ? (nbytes - boff) : 128);
        const uint32_t nq = blen >> 2;
        uint8_t* const bout = out + tile_off + s_off[bl];
        const uint32_t q = (lane < nq && boff / 4 + lane < nquads) ? in[boff / 4 + lane] : 0u;
        if (copymap && copymap[b]) {                        // copy-mode block: raw bytes (codec.rs:36)
            if (lane < nq) { st_u16(bout + 4 * lane, q & 0xFFFFu); st_u16(bout + 4 * lane + 2, q >> 16); }
            if (lane < (blen & 3u)) bout[(blen & ~3u) + lane] = in_b[boff + (blen & ~3u) + lane];
            continue;
        }
        const uint32_t act = nq >= 32 ? 0xFFFFFFFFu : ((1u << nq) - 1u);
        const uint32_t pm = Pbits[b] & act, am = Abits[b] & act, bm = Bbits[b] & act;
        const uint32_t maps = am | bm, plain = act & ~(pm | maps);
        if (lane < 4) {
            // signature: 2 bits per quad, MAP_A=1 MAP_B=2 PREDICTED=3 (cheetah.rs:18-21), LSB first (write_signature.rs:13-16)
            const uint32_t lowbits = am | pm, highbits = bm | pm;
            const uint32_t lo = part1by1(lowbits) | (part1by1(highbits) << 1);
            const uint32_t hi = part1by1(lowbits >> 16) | (part1by1(highbits >> 16) << 1);
            st_u16(bout + 2 * lane, ((lane < 2 ? lo : hi) >> (16 * (lane & 1))) & 0xFFFFu);
        }
        if (lane < nq) {
            uint8_t* p = bout + 8 + 4 * __popc(plain & lanemask_lt()) + 2 * __popc(maps & lanemask_lt());
            if ((plain >> lane) & 1u) { st_u16(p, q & 0xFFFFu); st_u16(p + 2, q >> 16); }       // cheetah.rs:131-132
            else if ((maps >> lane) & 1u) st_u16(p, prod_hash(hash_prod(q)));                     // cheetah.rs:134-135,140-141
        }
        if (lane < (blen & 3u)) {                           // 1..3 raw tail bytes of the last block (codec.rs:58-61)
            uint8_t* p = bout + 8 + 4 * __popc(plain) + 2 * __popc(maps);
            p[lane] = in_b[boff + (blen & ~3u) + lane];
        }
    }
}

__global__ void chee_open_gate(Status* __restrict__ st) { st->nonquiet = 1; }   // Cheetah always runs the copy-map iteration

}  // namespace chee

using namespace chee;

struct CheeLayout {
    size_t status, status2, Pbits, Abits, Bbits, copymap, copymap2, incb, seg_state, ctx0, tile_bytes, tile_local, group_total, group_off, total;
};

constexpr uint32_t PREFIX_TILES = 64;           // stage A settles the copy map of the first MiB on its own (cold-dictionary blocks)

static uint32_t chee_pick_runs(size_t nbytes, int num_sms) {
    const uint64_t nblocks = (nbytes + 127) / 128;
    const uint64_t ntiles = (nblocks + TILE_B - 1) / TILE_B;
    uint64_t r = ntiles / 2;                    // >= 32 KiB per run
    const uint64_t cap = (uint64_t)num_sms * 8; // 8 warps (runs) per SM
    if (r > cap) r = cap;
    if (r < PREFIX_TILES && ntiles >= PREFIX_TILES) r = PREFIX_TILES;
    if (r < 1) r = 1;
    return (uint32_t)r;
}

static size_t chee_layout(size_t nbytes, uint32_t nruns, CheeLayout* L) {
    const uint64_t nblocks = (nbytes + 127) / 128;
    const uint64_t ntiles = (nblocks + TILE_B - 1) / TILE_B;
    const uint64_t ngroups = (ntiles + 4095) / 4096;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    L->status = take(sizeof(Status));
    L->status2 = take(sizeof(Status));
    L->Pbits = take((ntiles * TILE_B + 32) * 4);
    L->Abits = take((ntiles * TILE_B + 32) * 4);
    L->Bbits = take((ntiles * TILE_B + 32) * 4);
    L->copymap = take(ntiles * TILE_B + 64);
    L->copymap2 = take(ntiles * TILE_B + 64);
    L->incb = take(ntiles * TILE_B + 64);
    L->seg_state = take((2 * (ntiles * TILE_B / 256 + 2) + 64) * 4);
    L->ctx0 = take((size_t)nruns * 4 + 64);
    L->tile_bytes = take((ntiles + 1) * 4);
    L->tile_local = take((ntiles + 1) * 4);
    L->group_total = take((ngroups + 1) * 8);
    L->group_off = take((ngroups + 1) * 8);
    L->total = off;
    return off;
}

size_t chee_workspace_bytes(size_t nbytes, int num_sms) {
    CheeLayout L;
    return chee_layout(nbytes, chee_pick_runs(nbytes, num_sms), &L);
}
// The per-run tables live in a buffer of their own: it must be zero when allocated and only ever be written by these kernels
// (entries are validated by epoch tags instead of being cleared).
size_t chee_tables_bytes(size_t nbytes, int num_sms) { return (size_t)chee_pick_runs(nbytes, num_sms) * 65536 * sizeof(uint4) * 2; }

// Enqueue the parallel Cheetah encode. *d_converged (device u32) != 0 afterwards means d_out / d_out_size hold the result; otherwise
// the caller's in-order kernel (queued behind, gated on that flag) produces it. `epoch_base`: the caller hands out 16 fresh epochs
// (values in 1 .. 2^30) per call and clears `tables` if it ever has to reuse one.
cudaError_t chee_encode_parallel(const uint8_t* d_in, size_t nbytes, uint8_t* d_out, size_t cap, uint8_t* ws, uint8_t* tables, uint32_t epoch_base,
                                 int num_sms, uint64_t* d_out_size, uint32_t* d_converged, cudaStream_t stream, uint64_t* launches) {
    const uint32_t nruns = chee_pick_runs(nbytes, num_sms);
    CheeLayout L; chee_layout(nbytes, nruns, &L);
    const uint64_t nblocks = (nbytes + 127) / 128;
    const uint32_t ntiles = (uint32_t)((nblocks + TILE_B - 1) / TILE_B);
    const uint32_t ngroups = (ntiles + 4095) / 4096;
    Status* stA = reinterpret_cast<Status*>(ws + L.status);
    Status* stB = reinterpret_cast<Status*>(ws + L.status2);
    const uint32_t* in32 = reinterpret_cast<const uint32_t*>(d_in);
    uint32_t* Pb = reinterpret_cast<uint32_t*>(ws + L.Pbits); uint32_t* Ab = reinterpret_cast<uint32_t*>(ws + L.Abits); uint32_t* Bb = reinterpret_cast<uint32_t*>(ws + L.Bbits);
    uint8_t* cm = ws + L.copymap; uint8_t* cm2 = ws + L.copymap2; uint8_t* incb = ws + L.incb;
    uint32_t* seg = reinterpret_cast<uint32_t*>(ws + L.seg_state);
    uint32_t* ctx0 = reinterpret_cast<uint32_t*>(ws + L.ctx0);
    uint32_t* tile_bytes = reinterpret_cast<uint32_t*>(ws + L.tile_bytes);
    uint4* entP = reinterpret_cast<uint4*>(tables);
    uint4* entC = entP + (size_t)nruns * 65536;
    cudaError_t e = cudaMemsetAsync(ws + L.status, 0, L.Pbits - L.status, stream);     // both status blocks
    if (e != cudaSuccess) return e;

    // one fixed-point round over the first `nb` bytes cut into `runs` runs: flags under the current map, incompressible bits, automaton
    auto round = [&](Status* st, int it, size_t nb, uint32_t runs, uint32_t epoch) -> cudaError_t {
        const uint64_t nq = nb / 4, nblk = (nb + 127) / 128;
        const uint32_t nt = (uint32_t)((nblk + TILE_B - 1) / TILE_B);
        const uint32_t nseg = (uint32_t)((nblk + 255) / 256);
        const uint32_t run_ctas = (runs + RP_WARPS - 1) / RP_WARPS;
        const uint8_t* mask = it ? cm : nullptr;
        chee_ctx0<<<(runs + 127) / 128, 128, 0, stream>>>(in32, nq, mask, runs, nt, st, ctx0);
        chee_pass_p<<<run_ctas, RP_WARPS * 32, 0, stream>>>(in32, nq, nblk, mask, runs, nt, st, ctx0, entP, epoch, Pb);
        chee_fold_p<<<65536 / 128, 128, 0, stream>>>(in32, runs, st, entP, epoch, Pb);
        chee_pass_c<<<run_ctas, RP_WARPS * 32, 0, stream>>>(in32, nq, nblk, mask, runs, nt, st, Pb, entC, epoch, Ab, Bb);
        chee_fold_c<<<65536 / 128, 128, 0, stream>>>(in32, runs, st, entC, epoch, Ab, Bb);
        chee_tile_sizes<<<(nt + 7) / 8, 256, 0, stream>>>(Pb, Ab, Bb, mask, nb, nblk, nt, 0, st, incb, tile_bytes);
        *launches += 7;
        return prot_iterate_launch(nullptr, nb, nblk, nseg, st, it, incb, cm, cm2, seg, seg + (nseg + 1), 128, num_sms, stream);
    };

    Status* st = stA;
    int first_it = 0;
    chee_open_gate<<<1, 1, 0, stream>>>(stA); ++*launches;
    if (ntiles > 2 * PREFIX_TILES) {
        // Stage A: a cold dictionary makes the first blocks incompressible on every input, and settling that takes ~4 rounds: run
        // them on the first MiB alone (the copy map of a prefix does not depend on what follows). Stage B then starts from that map
        // and normally confirms it in one round over the whole input.
        const size_t nbA = (size_t)PREFIX_TILES * TILE_B * 128;
        e = cudaMemsetAsync(cm, 0, (size_t)ntiles * TILE_B, stream);
        if (e != cudaSuccess) return e;
        for (int it = 0; it <= 6; ++it) {
            e = round(stA, it, nbA, PREFIX_TILES, epoch_base + (uint32_t)it);
            if (e != cudaSuccess) return e;
        }
        chee_open_gate<<<1, 1, 0, stream>>>(stB); ++*launches;
        st = stB;
        first_it = 1;
    }
    for (int it = first_it; it <= 7; ++it) {
        e = round(st, it, nbytes, nruns, epoch_base + 8u + (uint32_t)it);
        if (e != cudaSuccess) return e;
    }
    // final sizes under the committed copy map (valid only if converged), scan, emit
    chee_tile_sizes<<<(ntiles + 7) / 8, 256, 0, stream>>>(Pb, Ab, Bb, cm, nbytes, nblocks, ntiles, 1, st, incb, tile_bytes);
    e = scan_tiles_launch(tile_bytes, ntiles, reinterpret_cast<uint32_t*>(ws + L.tile_local),
                          reinterpret_cast<uint64_t*>(ws + L.group_total), reinterpret_cast<uint64_t*>(ws + L.group_off), ngroups, st, cap, d_out_size, stream);
    if (e != cudaSuccess) return e;
    chee_emit<<<ntiles, 256, 0, stream>>>(in32, nbytes, nblocks, Pb, Ab, Bb, cm, st, reinterpret_cast<uint32_t*>(ws + L.tile_local),
                                          reinterpret_cast<uint64_t*>(ws + L.group_off), 4096, d_out);
    *launches += 4;
    e = cudaMemcpyAsync(d_converged, &st->converged, sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

}  // namespace dns
