// common.cuh — shared device/host helpers for the density_b200 CUDA path (sm_100a only).
//
// Reference semantics being reproduced (paths relative to /root/reference/src):
//   hash            algorithms/chameleon/chameleon.rs:89 (= cheetah.rs:124, lion.rs:212)
//   block geometry  chameleon.rs:138-147, cheetah.rs:188-197, lion.rs:317-326
//   protection      codec/protection_state.rs:9-47
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace dns {

constexpr uint32_t HASH_MULT = 0x9D6EF916u;       // chameleon.rs:15
constexpr uint32_t HASH_MULT_HALF = HASH_MULT >> 1; // odd: 0x4EB77C8B

// Multiplicative inverse of HASH_MULT_HALF modulo 2^32 (Newton iteration, evaluated at compile time).
constexpr uint32_t inv_odd_u32(uint32_t a) {
    uint32_t x = a;  // correct to 3 bits
    for (int i = 0; i < 5; ++i) x *= 2u - a * x;
    return x;
}
constexpr uint32_t HASH_MULT_HALF_INV = inv_odd_u32(HASH_MULT_HALF);
static_assert(HASH_MULT_HALF * HASH_MULT_HALF_INV == 1u, "inverse");

// The multiplier is even (2 * odd), so p = quad * M (mod 2^32) is always even and determines quad mod 2^31.
// Inside one 16-bit hash bucket h = p >> 16 a quad is therefore identified by the 16-bit *fingerprint*
//     f = (p & 0xFFFE) | (quad >> 31)
// which lets a whole 65,536-entry Chameleon dictionary live in 128 KiB of shared memory.
__host__ __device__ __forceinline__ uint32_t hash_prod(uint32_t quad) { return quad * HASH_MULT; }
__host__ __device__ __forceinline__ uint32_t prod_hash(uint32_t p) { return p >> 16; }
__host__ __device__ __forceinline__ uint32_t prod_fp(uint32_t p, uint32_t quad) { return (p & 0xFFFEu) | (quad >> 31); }
// Inverse of (h, f) -> quad.
__host__ __device__ __forceinline__ uint32_t quad_from_hf(uint32_t h, uint32_t f) {
    uint32_t p = (h << 16) | (f & 0xFFFEu);
    uint32_t q31 = ((p >> 1) * HASH_MULT_HALF_INV) & 0x7FFFFFFFu;
    return q31 | (f << 31);
}

enum Alg : int { ALG_CHAMELEON = 0, ALG_CHEETAH = 1, ALG_LION = 2 };

__host__ __device__ __forceinline__ uint32_t alg_block_bytes(int alg) { return alg == ALG_CHAMELEON ? 256u : alg == ALG_CHEETAH ? 128u : 64u; }
__host__ __device__ __forceinline__ uint32_t alg_sig_bytes(int alg) { return alg == ALG_LION ? 6u : 8u; }
__host__ __device__ __forceinline__ uint32_t alg_flag_bits(int alg) { return alg == ALG_CHAMELEON ? 1u : alg == ALG_CHEETAH ? 2u : 3u; }

// codec/protection_state.rs:1-47, restated as a POD usable on host and device.
struct Protection {
    uint32_t copy_penalty;        // u8 in the reference
    uint32_t copy_penalty_start;  // u8 in the reference
    uint32_t previous_incompressible;
    uint64_t counter;
    __host__ __device__ void init() { copy_penalty = 0; copy_penalty_start = 1; previous_incompressible = 0; counter = 0; }
    __host__ __device__ bool revert_to_copy() {  // :18-27
        if ((counter & 0xf) == 0 && copy_penalty_start > 1) copy_penalty_start >>= 1;
        counter++;
        return copy_penalty > 0;
    }
    __host__ __device__ void decay() {  // :29-35
        copy_penalty = (copy_penalty - 1) & 0xff;
        if (copy_penalty == 0) copy_penalty_start = (copy_penalty_start + 1) & 0xff;
    }
    __host__ __device__ void update(bool incompressible) {  // :37-47
        if (incompressible) {
            if (previous_incompressible) copy_penalty = copy_penalty_start;
            previous_incompressible = 1;
        } else {
            previous_incompressible = 0;
        }
    }
};

// Device status block shared by the kernels of one encode/decode call.
struct Status {
    unsigned long long out_bytes;   // result size (0 on error)
    unsigned int nonquiet;          // fast path invalid: protection automaton would have fired
    unsigned int error;             // 0 ok; see density_b200.h DENSITY_B200_E*
    unsigned long long first_nonquiet_block;
    unsigned int converged;         // copy map reached its fixed point (parallel protection iteration)
    unsigned int iter_changed;      // scratch of the iteration's compare step
    unsigned int barrier[8];        // grid barriers of the protection iteration kernels (one per launch)
    unsigned int relax_changed[2];  // ping-pong "some segment was re-evaluated in this round"
    unsigned int pad2[2];
};

__device__ __forceinline__ uint32_t lanemask_lt() { uint32_t m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }
__device__ __forceinline__ uint32_t lanemask_gt() { uint32_t m; asm("mov.u32 %0, %%lanemask_gt;" : "=r"(m)); return m; }

// Streaming loads/stores: input is read once per pass, output written once.
__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t* p) {
    uint32_t v; asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v;
}
__device__ __forceinline__ void st_u16(uint8_t* p, uint32_t v) { *reinterpret_cast<uint16_t*>(p) = (uint16_t)v; }

}  // namespace dns
