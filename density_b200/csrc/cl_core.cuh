// cl_core.cuh — the table logic of the run-parallel Cheetah / Lion DECODERS, written once for device and host.
//
// Reference semantics: /root/reference/src/algorithms/cheetah/cheetah.rs:67-103 (decode_plain / decode_map_a / decode_map_b /
// decode_predicted) and lion/lion.rs:84-186 (decode_plain / decode_map_a / decode_map_b / decode_predicted_a..e, shift_predictions
// :50-57), driven by codec/codec.rs:82-126.
//
// A decoder never COMPARES values: every table operation is positional.
//   chunk map bucket (a, b):       PLAIN(v): (v, a)      MAP_A: read a       MAP_B: read b, then (b, a)
//   prediction list of a context:  not predicted (value v): push v in front (Cheetah: the list has one slot)
//                                  predicted at depth k: read slot k, move it to the front (Cheetah: k = 0, nothing moves)
// so a run (a contiguous piece of the block list, walked by one warp) can execute them on SYMBOLIC lists: a slot holds either a
// literal the run produced itself or "slot j of the list carried into the run". A fold over the runs (one thread per key) turns the
// symbolic lists into each run's concrete carried-in list. What is not positional is WHICH list a quad uses: the context is the
// hash of the previous quad, and for a predicted quad that hash comes out of the table. That part is iterated (cl_decode.cu).
//
// This header is included by cl_decode.cu (nvcc) and by the host-side model tests/cl_model.cpp (g++), which checks the whole
// scheme against the oracle on the CPU.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define CLD_HD __host__ __device__ __forceinline__
#else
#define CLD_HD inline
#endif

namespace dns {
namespace cld {

constexpr uint32_t HASH_MULT = 0x9D6EF916u;                 // cheetah.rs:15, lion.rs:15
CLD_HD uint32_t hash16(uint32_t q) { return (q * HASH_MULT) >> 16; }

constexpr uint32_t H_UNKNOWN = 0xFFFFFFFFu;                  // a context / hash that cannot be known in this round
constexpr uint32_t TAG_LIT = 0;                             // slot tag: a value the run produced; tag j + 1 = carried-in slot j

// A list of N slots in registers. tag: 3 bits per slot; unk: bit per slot = "the value in v[] is not known in this round".
template <int N>
struct List {
    uint32_t v[N];
    uint32_t tag;
    uint32_t unk;
    CLD_HD uint32_t slot_tag(int s) const { return (tag >> (3 * s)) & 7u; }
};

// first touch of a key inside a run: slot j = carried-in slot j, value taken from the run's snapshot (or unknown)
template <int N>
CLD_HD void list_init(List<N>& L, const uint32_t* snap /* N values or nullptr */) {
    L.tag = 0; L.unk = 0;
#pragma unroll
    for (int s = 0; s < N; ++s) {
        L.v[s] = snap ? snap[s] : 0u;
        L.tag |= (uint32_t)(s + 1) << (3 * s);
        if (!snap) L.unk |= 1u << s;
    }
}
// push a literal in front, the last slot falls out (shift_predictions lion.rs:50-57; chunk map (a, b) <- (v, a) cheetah.rs:72-73)
template <int N>
CLD_HD void list_push(List<N>& L, uint32_t v) {
#pragma unroll
    for (int s = N - 1; s > 0; --s) L.v[s] = L.v[s - 1];
    L.v[0] = v;
    const uint32_t mask = (N * 3 >= 32) ? 0xFFFFFFFFu : ((1u << (3 * N)) - 1u);
    L.tag = (L.tag << 3) & mask;                            // TAG_LIT in slot 0
    L.unk = (L.unk << 1) & ((1u << N) - 1u);
}
// move slot k to the front (lion.rs:133-186: predicted b..e rotate entries [0..k]; MAP_B swaps (a, b): cheetah.rs:92-93)
template <int N>
CLD_HD void list_mtf(List<N>& L, int k) {
    const uint32_t vk = L.v[k < N ? k : 0];
    const uint32_t tk = L.slot_tag(k), uk = (L.unk >> k) & 1u;
#pragma unroll
    for (int s = N - 1; s > 0; --s) if (s <= k) L.v[s] = L.v[s - 1];
    L.v[0] = vk;
    const uint32_t low = (1u << (3 * k)) - 1u;              // tags of slots [0, k)
    const uint32_t keep = ~((1u << (3 * (k + 1))) - 1u);    // tags of slots above k stay
    L.tag = (L.tag & keep) | ((L.tag & low) << 3) | tk;
    const uint32_t ulow = (1u << k) - 1u, ukeep = ~((1u << (k + 1)) - 1u);
    L.unk = (L.unk & ukeep) | ((L.unk & ulow) << 1) | uk;
}
// fold step: the list carried OUT of a run, given the list carried INTO it (`c`) and the run's symbolic final list
template <int N>
CLD_HD void list_carry(uint32_t (&c)[N], const List<N>& L) {
    uint32_t n[N];
#pragma unroll
    for (int s = 0; s < N; ++s) {
        const uint32_t t = L.slot_tag(s);
        uint32_t x = L.v[s];
#pragma unroll
        for (int j = 0; j < N; ++j) if (t == (uint32_t)(j + 1)) x = c[j];
        n[s] = x;
    }
#pragma unroll
    for (int s = 0; s < N; ++s) c[s] = n[s];
}

// ---- entry formats in global memory (per run and key; an entry whose epoch is not the current one counts as untouched) ----------
// N = 1 (Cheetah prediction): uint2  {v0, meta}
// N = 2 (chunk map):          uint4  {a, b, meta, 0}
// N = 5 (Lion prediction):    2 x uint4 {v0, v1, v2, v3} {v4, meta, 0, 0}
// meta = epoch << 20 | unk << 15 | tags (15 bits)
constexpr uint32_t META_EPOCH_SHIFT = 20;
constexpr uint32_t EPOCH_MAX = 0xFFEu;
template <int N> CLD_HD uint32_t list_meta(const List<N>& L, uint32_t epoch) { return (epoch << META_EPOCH_SHIFT) | (L.unk << 15) | L.tag; }
template <int N> CLD_HD void list_from_meta(List<N>& L, uint32_t meta) { L.tag = meta & 0x7FFFu; L.unk = (meta >> 15) & 31u; }
CLD_HD uint32_t meta_epoch(uint32_t meta) { return meta >> META_EPOCH_SHIFT; }

// ---- signature decoding (read_signature.rs:11-16: flag k sits at bits [k*F, (k+1)*F), LSB first) -----------------------------------
// Cheetah: 2-bit flags, 32 per 8-byte signature: 0 plain (4 B), 1 MAP_A (2 B), 2 MAP_B (2 B), 3 predicted (0 B)   cheetah.rs:18-21
// Lion:    3-bit flags, 16 per 6-byte signature: 0 plain (4 B), 1..5 predicted a..e (0 B), 6 MAP_A, 7 MAP_B (2 B)  lion.rs:18-25
CLD_HD uint32_t cheetah_block_bytes(uint64_t sig) {          // 8 + 4 * plain + 2 * map   (codec.rs:94-98)
    const uint64_t M = 0x5555555555555555ull;
    const uint64_t lo = sig & M, hi = (sig >> 1) & M;
    const uint64_t plain = ~(lo | hi) & M, map = lo ^ hi;
#if defined(__CUDA_ARCH__)
    return 8u + 4u * (uint32_t)__popcll(plain) + 2u * (uint32_t)__popcll(map);
#else
    return 8u + 4u * (uint32_t)__builtin_popcountll(plain) + 2u * (uint32_t)__builtin_popcountll(map);
#endif
}
CLD_HD uint32_t lion_block_bytes(uint64_t sig48) {           // 6 + 4 * plain + 2 * map
    const uint64_t M = 0x0000249249249249ull;                // bit 0 of each 3-bit flag
    const uint64_t b0 = sig48 & M, b1 = (sig48 >> 1) & M, b2 = (sig48 >> 2) & M;
    const uint64_t plain = ~(b0 | b1 | b2) & M, map = b1 & b2;
#if defined(__CUDA_ARCH__)
    return 6u + 4u * (uint32_t)__popcll(plain) + 2u * (uint32_t)__popcll(map);
#else
    return 6u + 4u * (uint32_t)__builtin_popcountll(plain) + 2u * (uint32_t)__builtin_popcountll(map);
#endif
}

// quad kinds after unpacking (algorithm independent)
enum : uint32_t { K_PLAIN = 0, K_MAP_A = 1, K_MAP_B = 2, K_PRED = 3 };
CLD_HD uint32_t cheetah_kind(uint32_t flag) { return flag; }                                       // cheetah.rs:18-21
CLD_HD uint32_t lion_kind(uint32_t flag) { return flag == 0 ? K_PLAIN : flag == 6 ? K_MAP_A : flag == 7 ? K_MAP_B : K_PRED; }
CLD_HD uint32_t lion_depth(uint32_t flag) { return flag - 1u; }                                    // predicted a..e -> 0..4

}  // namespace cld
}  // namespace dns
