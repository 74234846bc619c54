// common.cuh -- shared types, error handling and small device helpers for the bz3 B200 block codec.
#pragma once
#include <cstdint>
#include <cstddef>
#include <cstdio>

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define BZ_HD __host__ __device__ __forceinline__
#define BZ_D __device__ __forceinline__
#define BZ_DEVICE_CODE 1
// dynamic shared memory of a kernel, and a hint inside spin-waits (both have a CPU-emulator meaning, below)
#define BZ_DYN_SMEM(type, name) extern __shared__ __align__(16) type name[]
#define BZ_SPIN_HINT()
// kernel launch: BZ_LAUNCH(grid, block, smem, stream, kernel<...>)(args...).  The kernel name comes last so that
// template argument lists may contain commas.  (The CPU emulator of the tests has its own definition.)
#define BZ_LAUNCH(grid, block, smem, stream, ...) __VA_ARGS__<<<(grid), (block), (smem), (stream)>>>
#else
#define BZ_HD inline
#define BZ_D inline
// BZ_EMU: the CPU test-suite compiles the kernel bodies with g++ on top of tests/native/cta_emu.h (one
// fiber per CUDA thread), which must be included first and supplies the CUDA vocabulary.  Inline PTX has
// a plain C++ twin under BZ_EMU.  Test infrastructure only -- the library itself is always built by nvcc.
#if defined(BZ_EMU)
#define BZ_DEVICE_CODE 1
#endif
#endif

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int8_t s8;
typedef int32_t s32;
typedef int64_t s64;

namespace bz3 {

// block-format constants (reference: src/libbz3.c:84-87, include/common.h:23-25)
constexpr int kLzpSlotsLog2 = 18;
constexpr int kLzpSlots = 1 << kLzpSlotsLog2;
constexpr int kLzpMinMatch = 40;
constexpr int kLzpEscape = 0xF2;
constexpr u32 kCrcPoly = 0x82F63B78u;  // reflected CRC-32C polynomial (src/libbz3.c:37-67)

BZ_HD size_t block_bound(size_t n) { return n + n / 50 + 32; }  // src/libbz3.c:510

#if defined(__CUDACC__)
#define BZ_CUDA_TRY(expr)                                                                         \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            fprintf(stderr, "[bz3_b200] CUDA error %s at %s:%d: %s\n", cudaGetErrorName(_e),      \
                    __FILE__, __LINE__, cudaGetErrorString(_e));                                  \
            return _e;                                                                            \
        }                                                                                         \
    } while (0)
#endif  // __CUDACC__

#if defined(BZ_DEVICE_CODE)
// every kernel launch of the library is counted per host thread (bench.py reports it as gpu_launches)
inline u64& launch_counter() {
    static thread_local u64 c = 0;
    return c;
}
#define BZ_NOTE_LAUNCH() (++::bz3::launch_counter())
#endif

#if defined(BZ_DEVICE_CODE)

constexpr u32 kFullMask = 0xFFFFFFFFu;

BZ_D u32 lane_id() { return threadIdx.x & 31; }
BZ_D u32 warp_id() { return threadIdx.x >> 5; }
BZ_D u32 lanemask_lt() {
#if defined(BZ_EMU)
    return (1u << (threadIdx.x & 31)) - 1u;
#else
    u32 m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
#endif
}

template <typename T>
BZ_D T warp_reduce_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
    return v;
}

// inclusive warp scan (sum)
BZ_D u32 warp_scan_incl(u32 v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        u32 t = __shfl_up_sync(kFullMask, v, o);
        if (lane_id() >= (u32)o) v += t;
    }
    return v;
}

// streaming 128-bit load that does not pollute L1 (guide: Guideline 13)
BZ_D uint4 ld_stream_u4(const uint4* p) {
#if defined(BZ_EMU)
    return *p;
#else
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
#endif
}

#endif  // BZ_DEVICE_CODE

}  // namespace bz3
