// ubench_walk.cu -- isolates the walker of csrc/cm_dec.cuh: cycles per byte of cm_walk_byte on a static table, alone in the
// block, next to eight parked warps, and next to eight warps that run the hand-off protocol with trivial work.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/variants/ubench_walk tools/ubench_walk.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../bzip3_b200/csrc/cm_dec.cuh"
using namespace bz3;

__device__ unsigned long long g_res[16];
__device__ u32 g_sink;

__device__ void fill(u32* ptab, u8* scode, u32 seed) {
    for (int k = threadIdx.x; k < 512; k += blockDim.x) {
        u32 x = (k + 1) * 2654435761u + seed;
        x ^= x >> 15;
        ptab[k] = 0x70000000u + (x & 0x1FFFFFFFu);   // P between 0.44 and 0.56: ~1 bit per decision, a shift per byte
        if (k & 1) ptab[k] = 0xF8000000u + (x & 0x03FFFFFFu);   // every other node: very likely 1 (few shifts)
    }
    for (int k = threadIdx.x; k < 2048; k += blockDim.x) scode[k] = (u8)((k * 40503u + seed) >> 7);
}

template <int MODE>   // 0: walker alone; 1: + 8 parked warps; 2: + 8 warps in the B/S protocol (always "hit")
__global__ void __launch_bounds__(288, 1) k_walk(int nbytes, u32 seed, int slot) {
    __shared__ __align__(16) u32 ptab[512];
    __shared__ u8 scode[2048];
    __shared__ volatile u32 vbyte[2];
    fill(ptab, scode, seed);
    __syncthreads();
    const int tid = threadIdx.x;
    if (tid >= 32) {
        if (MODE == 1) { nb_sync<7>(); }
        if (MODE == 2) {
            for (int i = 0; i < nbytes; i += 2) {
                nb_arrive<kBarS + 1>(); nb_sync<kBarB + 0>(); g_sink = vbyte[0];
                nb_arrive<kBarS + 0>(); nb_sync<kBarB + 1>(); g_sink = vbyte[1];
            }
        }
        return;
    }
    CmWalk W;
    W.insize = 1 << 30; W.scode = scode; W.ip = 4; W.absolute = false; W.low = 0; W.r = 0xFFFFFFFFu; W.code = 0x12345678u; W.d = W.code;
    u32 acc = 0;
    const unsigned long long t0 = clock64();
    for (int i = 0; i < nbytes; i += 2) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const u32 byte = cm_walk_byte(ptab + h * 256, W) & 255u;
            acc += byte;
            if (MODE == 2) {
                vbyte[h] = byte;
                if (h == 0) { nb_arrive<kBarB + 0>(); nb_sync<kBarS + 1>(); } else { nb_arrive<kBarB + 1>(); nb_sync<kBarS + 0>(); }
            }
            W.ip &= 1023;   // endless payload
        }
    }
    const unsigned long long t1 = clock64();
    if (MODE == 1) nb_arrive<7>();
    if (tid == 0) { g_res[slot] = t1 - t0; g_sink = acc; }
}

// the bare recurrence: 8 x (mul.hi, compare, two selects) per byte, no table, no shift test
__global__ void k_chain(int nbytes, u32 m0, int slot) {
    u32 r = 0xFFFFFFFFu, d = 0x12345678u, m = m0, acc = 0;
    const unsigned long long t0 = clock64();
    for (int i = 0; i < nbytes; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const u32 x = __umulhi(r, m);
            const bool bit = d <= x;
            const u32 nx = ~x;
            r = bit ? x : r + nx;
            d = bit ? d : d + nx;
            m = bit ? m + 12345u : m - 54321u;
            acc = acc * 2 + bit;
        }
        r |= 0xFF000000u;
    }
    const unsigned long long t1 = clock64();
    if (threadIdx.x == 0) { g_res[slot] = t1 - t0; g_sink = acc + r + d; }
}

int main() {
    const int nb = 20000;
    k_walk<0><<<1, 32>>>(nb, 1, 0);
    k_walk<0><<<1, 288>>>(nb, 1, 1);
    k_walk<1><<<1, 288>>>(nb, 1, 2);
    k_walk<2><<<1, 288>>>(nb, 1, 3);
    k_chain<<<1, 32>>>(nb, 0x80000000u, 4);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long r[16];
    cudaMemcpyFromSymbol(r, g_res, sizeof r);
    const char* names[] = {"walker alone, block of 1 warp", "walker alone, 8 other warps exited", "walker + 8 warps parked on a barrier",
                           "walker + 8 warps in the B/S hand-off (always hit)", "bare recurrence (no table, no shift test)"};
    for (int k = 0; k < 5; k++) printf("%-52s: %8.1f cycles per byte\n", names[k], (double)r[k] / nb);
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
