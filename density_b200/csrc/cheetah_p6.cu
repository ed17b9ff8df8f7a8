// cheetah_p6.cu — Cheetah pass P (PREDICTED flags) with the prediction table in shared memory (sm_100a).
//
// Replaces the prediction half of /root/reference/src/algorithms/cheetah/cheetah.rs:121-150, bit-exactly; same outputs as
// chee_ctx0 + chee_pass_p + chee_fold_p (cheetah_encode.cu), which walk 1 MiB-per-run tables in HBM at ~2000 cycles per 32 quads.
//
//   PREDICTED_i  <=>  quad_i == the quad of the previous access with the same CONTEXT (context = hash of the previous encoded
//   quad, 0 at the stream start; the table starts as "0 everywhere"), and every access leaves its quad in the table
//   (cheetah.rs:125,144,148).
//
// That is the Chameleon flag semantics with key = context and value = quad, so the kernel is the write / verify / mailbox tile
// protocol of cham_flag_pass6 (chameleon_encode.cu) with a 32-bit value table. 65536 contexts x 4 B do not fit
// one SM; the table is split by the top context bit and the input is swept once per half: CTA c owns run c >> 1 of the stream
// (74 runs on 148 SMs) and the contexts with top bit c & 1; a quad takes part in the sweep its context belongs to and is invisible
// in the other one. Reading the input twice costs 0.3 ms per GiB of HBM time; the table accesses cost shared-memory latency.
// What a run cannot know — the table carried in from earlier runs — is handled as in the Chameleon encoder: the first access of a
// context inside a run goes to an unresolved list, a fold over the runs' last-value tables gives every run's carry-in table, a
// small kernel patches the unresolved flags (run 0 starts from the real initial table: all zero, nothing unresolved).
#include "common.cuh"
#include "encode_internal.cuh"

namespace dns {
namespace p6 {

constexpr int THREADS = 512, QPT = 8, NW = THREADS / 32, WQ = 32 * QPT, TILE_Q = THREADS * QPT;   // tile = 4096 quads = 128 Cheetah blocks
constexpr int MB_SLOTS = 4096, MB_CAP = 4, SEC_SLOTS = 64, SEC_CAP = 16;
constexpr int HALF = 32768;
// record.y: key (15) | pos << 15 (12) | touched << 27 | old == value << 28 | dropped << 29
constexpr uint32_t R_TOUCHED = 1u << 27, R_OLDEQ = 1u << 28, R_DROPPED = 1u << 29;

struct Smem {
    uint32_t tab[HALF];           // quad of the last access of each context of this half
    uint32_t tbit[HALF / 32];     // context accessed in this run (run 0: all set, the table really starts as zeros)
    uint2 rec[TILE_Q];            // warp w: records [256 w, 256 w + cnt[w]) in stream order. x = quad, y see above
    union {
        uint16_t mb[MB_SLOTS][MB_CAP];      // (key >> 12) << 12 | record index
        uint2 dense[TILE_Q];                // replay only
    };
    uint32_t mbcnt[2][MB_SLOTS / 4];
    __align__(16) uint32_t sec[SEC_SLOTS][SEC_CAP];   // key << 12 | record index
    uint32_t seccnt[2][SEC_SLOTS];
    uint32_t pbits[2][TILE_Q / 32];
    uint32_t cnt[32];
    uint32_t unres_count;
    uint32_t overflow;
};
static_assert(sizeof(Smem) <= 227 * 1024, "pass P shared memory");

__device__ __forceinline__ bool bit_test(const uint32_t* bm, uint32_t i) { return (bm[i >> 5] >> (i & 31)) & 1u; }
__device__ __forceinline__ bool gate_closed(const Status* g) { return g && !(g->nonquiet && !g->converged); }

__global__ void p6_zero(uint32_t* __restrict__ p, uint64_t n, const Status* __restrict__ gate) {
    if (gate_closed(gate)) return;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) p[i] = 0;
}

__device__ __forceinline__ void append_unres(bool pred, uint32_t qidx, uint32_t key, uint32_t* s_count, uint2* __restrict__ unres) {
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, pred);
    if (m == 0) return;
    uint32_t base = 0;
    if ((threadIdx.x & 31) == 0) base = atomicAdd(s_count, (uint32_t)__popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, 0);
    if (pred) {
        const uint32_t idx = base + __popc(m & lanemask_lt());
        if (idx < (uint32_t)HALF) unres[idx] = make_uint2(qidx, key);
    }
}

// Overflow fallback: the dirty members of the tile in stream order by one warp (the pre-tile values were restored by their owners).
__device__ __noinline__ void replay(Smem& S, uint32_t buf, uint64_t tile_q0, uint2* __restrict__ unres) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t c = lane < (uint32_t)NW ? S.cnt[lane] : 0u;
    uint32_t incl = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d); if ((int)lane >= d) incl += t; }
    const uint32_t excl = incl - c;
    const uint32_t n = __shfl_sync(0xFFFFFFFFu, incl, 31);
    #pragma unroll 1
    for (uint32_t i0 = 0; i0 < n; i0 += 32) {
        const uint32_t i = i0 + lane;
        const bool valid = i < n;
        uint32_t w = 0;
#pragma unroll
        for (int b = 16; b >= 1; b >>= 1) { const uint32_t t = __shfl_sync(0xFFFFFFFFu, incl, (w + b - 1) & 31); if (t <= i) w += b; }
        const uint32_t e = __shfl_sync(0xFFFFFFFFu, excl, w & 31);
        uint2 r = make_uint2(0, 0);
        if (valid) r = S.rec[w * WQ + (i - e)];
        const uint32_t key = r.y & 0x7FFFu, pos = (r.y >> 15) & 0xFFFu, v = r.x;
        uint32_t cur = 0; bool touched = false;
        if (valid) { cur = S.tab[key]; touched = bit_test(S.tbit, key); }
        const uint32_t grp = __match_any_sync(0xFFFFFFFFu, valid ? key : 0x10000u + lane);
        const uint32_t lower = grp & lanemask_lt();
        const uint32_t vprev = __shfl_sync(0xFFFFFFFFu, v, lower ? 31 - __clz(lower) : 0);
        bool hit;
        if (lower) { hit = vprev == v; touched = true; }
        else hit = touched && cur == v;
        if (valid && (grp & lanemask_gt()) == 0) {
            S.tab[key] = v;
            atomicOr(&S.tbit[key >> 5], 1u << (key & 31));
        }
        if (valid && hit) atomicOr(&S.pbits[buf][pos >> 5], 1u << (pos & 31));
        append_unres(valid && !touched, (uint32_t)(tile_q0 + pos), key, &S.unres_count, unres);
        __syncwarp();
    }
}

// grid = 2 * nruns CTAs: CTA c = (run c >> 1, half c & 1).
__global__ void __launch_bounds__(THREADS, 1)
chee_pass_p6(const uint32_t* __restrict__ in, uint64_t nquads, uint64_t ntiles, uint32_t nruns, const uint8_t* __restrict__ copymap,
             const Status* __restrict__ gate, uint32_t* __restrict__ Pbits, uint32_t* __restrict__ final_val /* nruns x 65536 */,
             uint32_t* __restrict__ final_tbit /* nruns x 2048 */, uint2* __restrict__ unres_all /* 2 nruns x HALF */,
             uint32_t* __restrict__ unres_count /* 2 nruns */) {
    if (gate_closed(gate)) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& S = *reinterpret_cast<Smem*>(smem_raw);
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t run = blockIdx.x >> 1, half = blockIdx.x & 1u;
    const uint64_t t_begin = (uint64_t)run * ntiles / nruns, t_end = (uint64_t)(run + 1) * ntiles / nruns;
    uint2* __restrict__ unres = unres_all + (size_t)blockIdx.x * HALF;
    for (uint32_t i = tid; i < (uint32_t)HALF; i += THREADS) S.tab[i] = 0;
    for (uint32_t i = tid; i < (uint32_t)HALF / 32; i += THREADS) S.tbit[i] = run == 0 ? 0xFFFFFFFFu : 0u;   // cheetah.rs:53: the table starts as zeros
    for (uint32_t i = tid; i < (uint32_t)MB_SLOTS / 4; i += THREADS) { S.mbcnt[0][i] = 0; S.mbcnt[1][i] = 0; }
    if (tid < SEC_SLOTS) { S.seccnt[0][tid] = 0; S.seccnt[1][tid] = 0; }
    if (tid == 0) { S.unres_count = 0; S.overflow = 0; }
    __syncthreads();

    const uint32_t pos0 = warp * WQ + lane;            // my sub-row j quad: pos0 + 32 j; sub-row j of warp w = block w * 8 + j of the tile
    #pragma unroll 1
    for (uint64_t t = t_begin; t < t_end; ++t) {
        const uint64_t tile_q0 = t * TILE_Q;
        const uint32_t buf = (uint32_t)(t - t_begin) & 1u;
#pragma unroll
        for (int k = 0; k < MB_SLOTS / 4 / THREADS; ++k) S.mbcnt[buf ^ 1u][tid + k * THREADS] = 0;
        if (tid < SEC_SLOTS) S.seccnt[buf ^ 1u][tid] = 0;
        // ---- loads: my quads and the quad before each of them
        uint32_t q[QPT], key[QPT], old[QPT];
        uint32_t validm = 0;
#pragma unroll
        for (int j = 0; j < QPT; ++j) {
            const uint64_t gq = tile_q0 + pos0 + 32 * j;
            const uint64_t blk = gq >> 5;
            q[j] = 0; key[j] = 0;
            const bool enc = gq < nquads && !(copymap && copymap[blk]);
            if (enc) {
                q[j] = ld_stream_u32(in + gq);
                uint32_t ctx = 0;
                if (lane) ctx = prod_hash(hash_prod(__ldg(in + gq - 1)));
                else {
                    // first quad of a block: the last quad of the nearest earlier encoded block (copy-mode episodes are at most 255 blocks long)
                    uint64_t pb = blk;
                    bool found = false;
                    while (pb > 0) { --pb; if (!(copymap && copymap[pb])) { found = true; break; } }
                    if (found) ctx = prod_hash(hash_prod(__ldg(in + pb * 32 + 31)));
                }
                if ((ctx >> 15) == half) { validm |= 1u << j; key[j] = ctx & 0x7FFFu; }
            }
        }
        // ---- A
        uint32_t missm = 0, vmin = 0xFFFFFFFFu;
#pragma unroll
        for (int j = 0; j < QPT; ++j) {
            old[j] = 0;
            if ((validm >> j) & 1u) { old[j] = S.tab[key[j]]; vmin = min(vmin, q[j]); if (old[j] != q[j]) missm |= 1u << j; }
        }
        if (vmin == 0) {        // value 0 is also what a never-accessed context shows
#pragma unroll
            for (int j = 0; j < QPT; ++j)
                if (((validm >> j) & 1u) && q[j] == 0 && old[j] == 0 && !bit_test(S.tbit, key[j])) missm |= 1u << j;
        }
        __syncthreads();   // S1
        // ---- B
#pragma unroll
        for (int j = 0; j < QPT; ++j)
            if ((missm >> j) & 1u) S.tab[key[j]] = q[j];
        __syncthreads();   // S2
        // ---- C
        uint32_t clean[QPT], base = 0;
        uint2* __restrict__ myrec = S.rec + warp * WQ;
#pragma unroll
        for (int j = 0; j < QPT; ++j) {
            const bool valid = (validm >> j) & 1u;
            bool dirty = (missm >> j) & 1u;
            if (valid && !dirty) dirty = S.tab[key[j]] != q[j];
            const uint32_t db = __ballot_sync(0xFFFFFFFFu, dirty);
            clean[j] = __ballot_sync(0xFFFFFFFFu, valid && !dirty);
            if (dirty) myrec[base + __popc(db & lanemask_lt())] = make_uint2(q[j], key[j] | ((pos0 + 32 * j) << 15) | (old[j] == q[j] ? R_OLDEQ : 0u));
            base += __popc(db);
        }
        if (lane == 0) {
#pragma unroll
            for (int j = 0; j < QPT; j += 4)
                *reinterpret_cast<uint4*>(&S.pbits[buf][warp * QPT + j]) = make_uint4(clean[j], clean[j + 1], clean[j + 2], clean[j + 3]);
            S.cnt[warp] = base;
        }
        __syncwarp();
        // deposit
        uint2 r0 = make_uint2(0, 0);
        bool drop0 = false;
        {
            uint32_t carry_x = 0, carry_y = 0xFFFFFFFFu;
            #pragma unroll 1
            for (uint32_t i0 = 0; i0 < base; i0 += 32) {
                const uint32_t i = i0 + lane;
                const bool valid = i < base;
                uint2 r = make_uint2(0, 0xFFFFFFFFu);
                if (valid) {
                    r = myrec[i];
                    if (bit_test(S.tbit, r.y & 0x7FFFu)) r.y |= R_TOUCHED;      // stable until phase D
                }
                uint32_t px = __shfl_up_sync(0xFFFFFFFFu, r.x, 1), py = __shfl_up_sync(0xFFFFFFFFu, r.y, 1);
                if (lane == 0) { px = carry_x; py = carry_y; }
                // same context, same quad, and the quad right before mine in the stream: a hit on that access, changes nothing
                const bool drop = valid && px == r.x && ((py ^ r.y) & 0x7FFFu) == 0 && (((py >> 15) & 0xFFFu) + 1u == ((r.y >> 15) & 0xFFFu)) && py != 0xFFFFFFFFu;
                carry_x = __shfl_sync(0xFFFFFFFFu, r.x, 31); carry_y = __shfl_sync(0xFFFFFFFFu, r.y, 31);
                if (i0 == 0) { r0 = r; drop0 = drop; }
                if (drop) {
                    const uint32_t pos = (r.y >> 15) & 0xFFFu;
                    atomicOr(&S.pbits[buf][pos >> 5], 1u << (pos & 31));
                    if (i0) myrec[i].y = r.y | R_DROPPED;
                } else if (valid) {
                    if (i0 && (r.y & R_TOUCHED)) myrec[i].y = r.y;
                    const uint32_t kk = r.y & 0x7FFFu, slot = kk & (MB_SLOTS - 1), sh = (slot & 3u) * 8u;
                    const uint32_t k = (atomicAdd(&S.mbcnt[buf][slot >> 2], 1u << sh) >> sh) & 0xFFu;
                    if (k < (uint32_t)MB_CAP) S.mb[slot][k] = (uint16_t)(((kk >> 12) << 12) | (warp * WQ + i));
                    else {
                        const uint32_t s2 = slot & (SEC_SLOTS - 1);
                        const uint32_t k2 = atomicAdd(&S.seccnt[buf][s2], 1u);
                        if (k2 < (uint32_t)SEC_CAP) S.sec[s2][k2] = (kk << 12) | (warp * WQ + i);
                        else S.overflow = 1;
                    }
                }
            }
        }
        __syncthreads();   // S3
        if (S.overflow) {
            // restore the pre-tile values (their owners still hold them), then one warp replays the dirty members in order
#pragma unroll
            for (int j = 0; j < QPT; ++j)
                if ((missm >> j) & 1u) S.tab[key[j]] = old[j];
            __syncthreads();
            if (warp == 0) replay(S, buf, tile_q0, unres);
        } else {
            // ---- D
            #pragma unroll 1
            for (uint32_t i0 = 0; i0 < base; i0 += 32) {
                const uint32_t i = i0 + lane;
                bool valid = i < base;
                uint2 r = r0;
                if (i0) r = valid ? myrec[i] : make_uint2(0, 0);
                if (i0 ? (r.y & R_DROPPED) != 0 : drop0) valid = false;
                const uint32_t kk = r.y & 0x7FFFu, pos = (r.y >> 15) & 0xFFFu, slot = kk & (MB_SLOTS - 1), myidx = warp * WQ + i;
                bool unresd = false;
                if (valid) {
                    const uint32_t n = (S.mbcnt[buf][slot >> 2] >> ((slot & 3u) * 8u)) & 0xFFu;
                    const uint2 e2 = *reinterpret_cast<const uint2*>(&S.mb[slot][0]);
                    const uint32_t me = ((kk >> 12) << 12) | myidx;
                    int best = -1; bool later = false, hit = false;
#pragma unroll
                    for (int tt = 0; tt < MB_CAP; ++tt) {
                        const uint32_t e = ((tt & 2) ? e2.y : e2.x) >> ((tt & 1) * 16) & 0xFFFFu;
                        if ((uint32_t)tt < n && ((e ^ me) >> 12) == 0) {
                            if (e < me) best = max(best, (int)(e & 0xFFFu));
                            later |= e > me;
                        }
                    }
                    if (n > (uint32_t)MB_CAP) {
                        const uint32_t s2 = slot & (SEC_SLOTS - 1);
                        const uint32_t n2 = S.seccnt[buf][s2];
                        const uint32_t mine = (kk << 12) | myidx;
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
                    if (best >= 0) hit = S.rec[best].x == r.x;
                    else if (r.y & R_TOUCHED) hit = (r.y & R_OLDEQ) != 0;
                    else unresd = true;
                    if (!later) {
                        S.tab[kk] = r.x;
                        if (!(r.y & R_TOUCHED)) atomicOr(&S.tbit[kk >> 5], 1u << (kk & 31));
                    }
                    if (hit) atomicOr(&S.pbits[buf][pos >> 5], 1u << (pos & 31));
                }
                append_unres(unresd, (uint32_t)(tile_q0 + pos), kk, &S.unres_count, unres);
            }
        }
        __syncthreads();   // S4
        if (tid < TILE_Q / 32) {
            const uint32_t w = S.pbits[buf][tid];
            const uint64_t blk = t * (TILE_Q / 32) + tid;
            if (w && blk * 32 < nquads) atomicOr(&Pbits[blk], w);      // the other half's CTA owns the other bits of the word
        }
        if (tid == 0) S.overflow = 0;
    }
    // export: last value and "accessed in this run" per context of my half
    for (uint32_t i = tid; i < (uint32_t)HALF; i += THREADS) final_val[(size_t)run * 65536 + half * HALF + i] = S.tab[i];
    for (uint32_t i = tid; i < (uint32_t)HALF / 32; i += THREADS) final_tbit[(size_t)run * 2048 + half * (HALF / 32) + i] = S.tbit[i];
    if (tid == 0) unres_count[blockIdx.x] = S.unres_count < (uint32_t)HALF ? S.unres_count : (uint32_t)HALF;
}

// carry[r][ctx] = table entry before run r (the table starts as zeros, cheetah.rs:53)
__global__ void p6_carry_scan(const uint32_t* __restrict__ final_val, const uint32_t* __restrict__ final_tbit, uint32_t nruns, const Status* __restrict__ gate,
                              uint32_t* __restrict__ carry) {
    if (gate_closed(gate)) return;
    const uint32_t ctx = blockIdx.x * blockDim.x + threadIdx.x;
    if (ctx >= 65536) return;
    uint32_t c = 0;
    for (uint32_t r = 0; r < nruns; ++r) {
        carry[(size_t)r * 65536 + ctx] = c;
        // run 0's bitmap is all ones by construction (its table is exact from the start): an entry counts only if it was written,
        // which for run 0 is indistinguishable from "still the initial 0" — and that is the same value.
        if ((final_tbit[(size_t)r * 2048 + (ctx >> 5)] >> (ctx & 31)) & 1u) c = final_val[(size_t)r * 65536 + ctx];
    }
}

__global__ void p6_resolve(const uint32_t* __restrict__ in, const uint2* __restrict__ unres_all, const uint32_t* __restrict__ unres_count,
                           const uint32_t* __restrict__ carry, const Status* __restrict__ gate, uint32_t* __restrict__ Pbits) {
    if (gate_closed(gate)) return;
    const uint32_t cta = blockIdx.y, run = cta >> 1, half = cta & 1u;
    const uint32_t n = unres_count[cta];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint2 e = unres_all[(size_t)cta * HALF + i];
        if (carry[(size_t)run * 65536 + half * HALF + e.y] == in[e.x]) atomicOr(&Pbits[e.x >> 5], 1u << (e.x & 31));
    }
}

}  // namespace p6

// bytes of scratch the pass needs behind `base` (the Cheetah pass-P table region is far larger)
size_t chee_p6_scratch_bytes(uint32_t nruns) {
    return (size_t)nruns * 65536 * 4 * 2 + (size_t)nruns * 2048 * 4 + (size_t)2 * nruns * p6::HALF * sizeof(uint2) + (size_t)2 * nruns * 4 + 1024;
}
uint32_t chee_p6_runs(uint64_t ntiles, int num_sms) {
    uint64_t r = (uint64_t)(num_sms > 1 ? num_sms / 2 : 1);
    if (r > ntiles) r = ntiles;
    if (r < 1) r = 1;
    return (uint32_t)r;
}

// Enqueue pass P for the first `nquads` quads (ntiles tiles of 4096 quads): Pbits[b] = PREDICTED bits of block b.
cudaError_t chee_pass_p6_launch(const uint32_t* in, uint64_t nquads, uint64_t ntiles, const uint8_t* copymap, const Status* gate, uint32_t* Pbits,
                                uint8_t* scratch, int num_sms, cudaStream_t stream, uint64_t* launches) {
    using namespace p6;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e0 = cudaFuncSetAttribute(chee_pass_p6, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
        if (e0 != cudaSuccess) return e0;
        attr_done = true;
    }
    const uint32_t nruns = chee_p6_runs(ntiles, num_sms);
    uint32_t* final_val = reinterpret_cast<uint32_t*>(scratch);
    uint32_t* carry = final_val + (size_t)nruns * 65536;
    uint32_t* final_tbit = carry + (size_t)nruns * 65536;
    uint2* unres = reinterpret_cast<uint2*>(final_tbit + (size_t)nruns * 2048);
    uint32_t* unres_count = reinterpret_cast<uint32_t*>(unres + (size_t)2 * nruns * HALF);
    p6_zero<<<num_sms * 2, 256, 0, stream>>>(Pbits, ntiles * (TILE_Q / 32), gate);
    chee_pass_p6<<<2 * nruns, THREADS, sizeof(Smem), stream>>>(in, nquads, ntiles, nruns, copymap, gate, Pbits, final_val, final_tbit, unres, unres_count);
    p6_carry_scan<<<65536 / 256, 256, 0, stream>>>(final_val, final_tbit, nruns, gate, carry);
    p6_resolve<<<dim3(16, 2 * nruns), 256, 0, stream>>>(in, unres, unres_count, carry, gate, Pbits);
    *launches += 4;
    return cudaGetLastError();
}

}  // namespace dns
