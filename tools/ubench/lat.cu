// Single-warp latency microbenchmarks (sm_100a): dependent chains of warp primitives and shared-memory ops.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N 256
__global__ void k(uint32_t* out, int mode, uint32_t seed) {
    __shared__ uint32_t sm[4096];
    __shared__ uint16_t sh16[8192];
    const uint32_t lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) { sm[i] = (i * 2654435761u) >> 20; }
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sh16[i] = (uint16_t)(i * 40503u);
    __syncthreads();
    if (threadIdx.x >= 32) { __syncthreads(); return; }
    uint32_t x = lane * 2654435761u + seed, acc = 0;
    long long t0 = clock64();
    if (mode == 0) { for (int i = 0; i < N; ++i) { x = __shfl_sync(0xFFFFFFFFu, x, (x + i) & 31) + 1; } }
    else if (mode == 1) { for (int i = 0; i < N; ++i) { uint32_t m = __match_any_sync(0xFFFFFFFFu, x & 0xFFFF); x = x * 1664525u + m; } }            // ~32 distinct
    else if (mode == 2) { for (int i = 0; i < N; ++i) { uint32_t m = __match_any_sync(0xFFFFFFFFu, x & 3); x = x * 1664525u + m; } }                 // 4 distinct
    else if (mode == 3) { for (int i = 0; i < N; ++i) { uint32_t m = __match_any_sync(0xFFFFFFFFu, 7u + (x & 0)); x = x * 1664525u + m; } }          // 1 distinct
    else if (mode == 4) { for (int i = 0; i < N; ++i) { x = sm[x & 4095] + i; } }                                                                       // LDS chain
    else if (mode == 5) { for (int i = 0; i < N; ++i) { x = atomicOr(&sm[x & 4095], 1u) + i; } }                                                        // ATOMS chain (returning)
    else if (mode == 6) { for (int i = 0; i < N; ++i) { x = __ballot_sync(0xFFFFFFFFu, x & 1) + x * 3u; } }                                            // VOTE chain
    else if (mode == 7) { for (int i = 0; i < N; ++i) { sm[(x & 1023)] = x; __syncwarp(); x = sm[(x & 1023)] * 3u + i; } }                              // STS + syncwarp + LDS
    else if (mode == 8) { for (int i = 0; i < N; ++i) { x = __reduce_or_sync(0xFFFFFFFFu, x) + lane + i; } }                                            // REDUX
    else if (mode == 9) { for (int i = 0; i < N; ++i) { x = sh16[x & 8191] * 7u + i; } }                                                                // LDS.U16 chain
    else if (mode == 10) { for (int i = 0; i < N; ++i) { atomicOr(&sm[x & 4095], 1u); __syncwarp(); x = x * 1664525u + i; } }                           // ATOMS (no return) + syncwarp
    else if (mode == 11) { for (int i = 0; i < N; ++i) { x = x * 1664525u + 12345u; } }                                                                 // IMAD chain
    else if (mode == 12) { for (int i = 0; i < N; ++i) { x = __popc(x) + x * 3u; } }
    else if (mode == 13) { for (int i = 0; i < N; ++i) { x = __shfl_up_sync(0xFFFFFFFFu, x, 1) + 1; } }
    long long t1 = clock64();
    acc += x;
    if (lane == 0) { out[0] = (uint32_t)((t1 - t0) / N); out[1] = acc; }
    __syncthreads();
}
int main() {
    uint32_t* d; cudaMalloc(&d, 8);
    const char* names[] = {"shfl.idx", "match.any 32 distinct", "match.any 4 distinct", "match.any 1 distinct", "lds.32 chain", "atoms.or chain", "vote.ballot", "sts+syncwarp+lds", "redux.or", "lds.u16 chain", "atoms(no ret)+syncwarp", "imad", "popc+imad", "shfl.up"};
    for (int nthreads : {32, 1024}) for (int m = 0; m < 14; ++m) {
        uint32_t h[2];
        k<<<1, nthreads>>>(d, m, 12345); cudaDeviceSynchronize();
        k<<<1, nthreads>>>(d, m, 999); cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost);
        printf("threads %4d  %-28s %u cycles/iter\n", nthreads, names[m], h[0]);
    }
    return 0;
}
