// ubench.cu -- latency / issue micro-benchmarks for the single-warp serial kernels of the entropy stage.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/variants/ubench tools/ubench.cu
// Every test times REPS back-to-back copies of one operation with clock64 inside the kernel (one CTA, the
// stated number of warps) and prints cycles per operation.  The numbers calibrate the cost model in
// DESIGN.md (dependent-issue latency of the instructions on the coder recurrences, barrier and shuffle
// round trips, taken branches).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

typedef uint32_t u32;
typedef uint64_t u64;

constexpr int REPS = 512;

#define REP8(x) x x x x x x x x
#define REP64(x) REP8(REP8(x))
#define REP512(x) REP8(REP64(x))

__device__ u64 g_out[64];
__device__ u32 g_sink;

__global__ void k_clock_overhead() {
    u64 t0 = clock64();
    u64 acc = 0;
#pragma unroll
    for (int i = 0; i < 64; i++) acc += clock64();
    u64 t1 = clock64();
    if (threadIdx.x == 0) { g_out[0] = t1 - t0; g_sink = (u32)acc; }
}

__global__ void k_dep_imad_wide(u32 seed, u32 m) {
    u32 r = seed;
    u64 t0 = clock64();
    REP512(asm volatile("{.reg .u64 w; mul.wide.u32 w, %0, %1; shr.u64 w, w, 32; cvt.u32.u64 %0, w;}" : "+r"(r) : "r"(m));)
    u64 t1 = clock64();
    if (threadIdx.x == 0) { g_out[1] = t1 - t0; g_sink = r; }
}

__global__ void k_dep_mulhi(u32 seed, u32 m) {
    u32 r = seed;
    u64 t0 = clock64();
    REP512(asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(r) : "r"(m));)
    u64 t1 = clock64();
    if (threadIdx.x == 0) { g_out[2] = t1 - t0; g_sink = r; }
}

__global__ void k_dep_iadd(u32 seed, u32 m) {
    u32 r = seed;
    u64 t0 = clock64();
    REP512(asm volatile("add.u32 %0, %0, %1;" : "+r"(r) : "r"(m));)
    u64 t1 = clock64();
    if (threadIdx.x == 0) { g_out[3] = t1 - t0; g_sink = r; }
}

__global__ void k_dep_setp_selp(u32 seed, u32 m) {
    u32 r = seed;
    u64 t0 = clock64();
    REP512(asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %1, %0, p; add.u32 %0, %0, 1;}" : "+r"(r) : "r"(m));)
    u64 t1 = clock64();
    if (threadIdx.x == 0) { g_out[4] = t1 - t0; g_sink = r; }   // 3 dependent instructions per rep
}

__global__ void k_dep_lds(u32 seed) {
    __shared__ u32 tab[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) tab[i] = ((i * 37 + 11) & 1023) * 4;
    __syncthreads();
    u32 base = (u32)__cvta_generic_to_shared(tab);
    u32 a = (seed & 1023) * 4;
    u64 t0 = clock64();
    REP512(asm volatile("{.reg .u32 t; add.u32 t, %0, %1; ld.shared.u32 %0, [t];}" : "+r"(a) : "r"(base) : "memory");)
    u64 t1 = clock64();
    if (threadIdx.x == 0) { g_out[5] = t1 - t0; g_sink = a; }   // add + lds per rep
}

__global__ void k_dep_shfl(u32 seed) {
    u32 r = seed + threadIdx.x;
    u64 t0 = clock64();
    REP512(r = __shfl_sync(0xFFFFFFFFu, r, (r + 1) & 31);)
    u64 t1 = clock64();
    if (threadIdx.x == 0) { g_out[6] = t1 - t0; g_sink = r; }   // lop/add + shfl per rep
}

__global__ void k_dep_ballot(u32 seed) {
    u32 r = seed + threadIdx.x;
    u64 t0 = clock64();
    REP512(r = __ballot_sync(0xFFFFFFFFu, (r >> (threadIdx.x & 31)) & 1u) + threadIdx.x;)
    u64 t1 = clock64();
    if (threadIdx.x == 0) { g_out[7] = t1 - t0; g_sink = r; }
}

__global__ void k_barrier(int slot) {
    u64 t0 = clock64();
    REP512(__syncthreads();)
    u64 t1 = clock64();
    if (threadIdx.x == 0) g_out[slot] = t1 - t0;
}

// independent work: 8 independent mul.hi chains interleaved -> issue rate of one warp
__global__ void k_indep_mulhi(u32 seed, u32 m) {
    u32 r0 = seed, r1 = seed + 1, r2 = seed + 2, r3 = seed + 3, r4 = seed + 4, r5 = seed + 5, r6 = seed + 6, r7 = seed + 7;
    u64 t0 = clock64();
    REP64(asm volatile("mul.hi.u32 %0, %0, %8; mul.hi.u32 %1, %1, %8; mul.hi.u32 %2, %2, %8; mul.hi.u32 %3, %3, %8;"
                       "mul.hi.u32 %4, %4, %8; mul.hi.u32 %5, %5, %8; mul.hi.u32 %6, %6, %8; mul.hi.u32 %7, %7, %8;"
                       : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7) : "r"(m));)
    u64 t1 = clock64();
    if (threadIdx.x == 0) { g_out[12] = t1 - t0; g_sink = r0 + r1 + r2 + r3 + r4 + r5 + r6 + r7; }   // 512 instructions
}
__global__ void k_indep_iadd(u32 seed, u32 m) {
    u32 r0 = seed, r1 = seed + 1, r2 = seed + 2, r3 = seed + 3, r4 = seed + 4, r5 = seed + 5, r6 = seed + 6, r7 = seed + 7;
    u64 t0 = clock64();
    REP64(asm volatile("add.u32 %0, %0, %8; add.u32 %1, %1, %8; add.u32 %2, %2, %8; add.u32 %3, %3, %8;"
                       "add.u32 %4, %4, %8; add.u32 %5, %5, %8; add.u32 %6, %6, %8; add.u32 %7, %7, %8;"
                       : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7) : "r"(m));)
    u64 t1 = clock64();
    if (threadIdx.x == 0) { g_out[13] = t1 - t0; g_sink = r0 + r1 + r2 + r3 + r4 + r5 + r6 + r7; }
}
// mixed pipes: mul.hi (FMA-heavy pipe) interleaved with add/lop (ALU pipe), all independent
__global__ void k_indep_mixed(u32 seed, u32 m) {
    u32 r0 = seed, r1 = seed + 1, r2 = seed + 2, r3 = seed + 3, r4 = seed + 4, r5 = seed + 5, r6 = seed + 6, r7 = seed + 7;
    u64 t0 = clock64();
    REP64(asm volatile("mul.hi.u32 %0, %0, %8; add.u32 %1, %1, %8; mul.hi.u32 %2, %2, %8; xor.b32 %3, %3, %8;"
                       "mul.hi.u32 %4, %4, %8; add.u32 %5, %5, %8; mul.hi.u32 %6, %6, %8; xor.b32 %7, %7, %8;"
                       : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7) : "r"(m));)
    u64 t1 = clock64();
    if (threadIdx.x == 0) { g_out[14] = t1 - t0; g_sink = r0 + r1 + r2 + r3 + r4 + r5 + r6 + r7; }
}

// taken branch: a loop whose body is one dependent add; compare with k_dep_iadd
__global__ void k_branch(u32 seed, u32 m, int iters) {
    u32 r = seed;
    u64 t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; i++) asm volatile("add.u32 %0, %0, %1;" : "+r"(r) : "r"(m));
    u64 t1 = clock64();
    if (threadIdx.x == 0) { g_out[15] = t1 - t0; g_sink = r; }
}

// flag ping-pong through shared memory between warp 0 and warp 1 (volatile spin): one-way latency
__global__ void k_pingpong(int iters) {
    __shared__ volatile u32 flag;
    if (threadIdx.x == 0) flag = 0;
    __syncthreads();
    const int w = threadIdx.x >> 5;
    u64 t0 = clock64();
    if (w == 0) {
        for (int i = 0; i < iters; i++) {
            if ((threadIdx.x & 31) == 0) flag = 2 * i + 1;
            while (flag != (u32)(2 * i + 2)) {}
        }
    } else if (w == 1) {
        for (int i = 0; i < iters; i++) {
            while (flag != (u32)(2 * i + 1)) {}
            if ((threadIdx.x & 31) == 0) flag = 2 * i + 2;
        }
    }
    u64 t1 = clock64();
    if (threadIdx.x == 0) g_out[16] = t1 - t0;
}

// named-barrier handshake between two warps: producer bar.arrive, consumer bar.sync (and back)
__global__ void k_named_bar(int iters) {
    const int w = threadIdx.x >> 5;
    u64 t0 = clock64();
    for (int i = 0; i < iters; i++) {
        if (w == 0) {
            asm volatile("bar.arrive 1, 64;");
            asm volatile("bar.sync 2, 64;");
        } else {
            asm volatile("bar.sync 1, 64;");
            asm volatile("bar.arrive 2, 64;");
        }
    }
    u64 t1 = clock64();
    if (threadIdx.x == 0) g_out[17] = t1 - t0;
}

static u64 fetch(int slot) {
    u64 h[64];
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(h, g_out, sizeof(h));
    return h[slot];
}

int main() {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) { printf("no device\n"); return 1; }
    printf("device %s, %d SMs, clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    for (int rep = 0; rep < 2; rep++) {   // second pass = warm
        k_clock_overhead<<<1, 32>>>();
        u64 c = fetch(0);
        if (rep) printf("clock64 read                       : %6.1f cycles each\n", c / 65.0);
        k_dep_imad_wide<<<1, 32>>>(0xF1234567u, 0xFFF00001u);
        c = fetch(1);
        if (rep) printf("dependent mul.wide.u32 (+hi extract): %6.2f cycles/op\n", (double)c / REPS);
        k_dep_mulhi<<<1, 32>>>(0xF1234567u, 0xFFF00001u);
        c = fetch(2);
        if (rep) printf("dependent mul.hi.u32               : %6.2f cycles/op\n", (double)c / REPS);
        k_dep_iadd<<<1, 32>>>(1, 3);
        c = fetch(3);
        if (rep) printf("dependent add.u32                  : %6.2f cycles/op\n", (double)c / REPS);
        k_dep_setp_selp<<<1, 32>>>(1, 77);
        c = fetch(4);
        if (rep) printf("dependent setp+selp+add            : %6.2f cycles/3 ops\n", (double)c / REPS);
        k_dep_lds<<<1, 32>>>(5);
        c = fetch(5);
        if (rep) printf("dependent add+ld.shared            : %6.2f cycles/op\n", (double)c / REPS);
        k_dep_shfl<<<1, 32>>>(5);
        c = fetch(6);
        if (rep) printf("dependent shfl.idx (+index math)   : %6.2f cycles/op\n", (double)c / REPS);
        k_dep_ballot<<<1, 32>>>(5);
        c = fetch(7);
        if (rep) printf("dependent ballot (+pred math)      : %6.2f cycles/op\n", (double)c / REPS);
        const int warps[4] = {1, 2, 8, 9};
        for (int k = 0; k < 4; k++) {
            k_barrier<<<1, 32 * warps[k]>>>(8 + k);
            c = fetch(8 + k);
            if (rep) printf("__syncthreads, %d warp(s)            : %6.2f cycles each\n", warps[k], (double)c / REPS);
        }
        k_indep_mulhi<<<1, 32>>>(0xF1234567u, 0xFFF00001u);
        c = fetch(12);
        if (rep) printf("independent mul.hi x8 interleaved  : %6.2f cycles/instr (1 warp)\n", (double)c / 512);
        k_indep_iadd<<<1, 32>>>(1, 3);
        c = fetch(13);
        if (rep) printf("independent add x8 interleaved     : %6.2f cycles/instr (1 warp)\n", (double)c / 512);
        k_indep_mixed<<<1, 32>>>(0xF1234567u, 0xFFF00001u);
        c = fetch(14);
        if (rep) printf("independent mul.hi/add/xor mix     : %6.2f cycles/instr (1 warp)\n", (double)c / 512);
        k_branch<<<1, 32>>>(1, 3, 512);
        c = fetch(15);
        if (rep) printf("loop: add + taken branch           : %6.2f cycles/iteration\n", (double)c / 512);
        k_pingpong<<<1, 64>>>(256);
        c = fetch(16);
        if (rep) printf("smem flag ping-pong, 2 warps       : %6.2f cycles per one-way hand-off\n", (double)c / 512);
        k_named_bar<<<1, 64>>>(256);
        c = fetch(17);
        if (rep) printf("bar.arrive/bar.sync handshake      : %6.2f cycles per one-way hand-off\n", (double)c / 512);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e == cudaSuccess ? 0 : 1;
}
