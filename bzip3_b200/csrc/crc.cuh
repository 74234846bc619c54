// crc.cuh -- block checksum on the GPU.
//
// Restates crc32sum (reference src/libbz3.c:37-72): reflected CRC-32C (poly 0x82F63B78), register
// initialised by the caller (the block codec uses 1), no final xor, one table step per byte:
//     s' = tab[(s ^ b) & 0xff] ^ (s >> 8)
// The update is GF(2)-linear in (s, b), so for a buffer cut into chunks
//     crc(init, buf) = init * x^(8n)  xor  XOR_j  raw_j * x^(8 * bytes_after_chunk_j)      (mod P)
// where raw_j is the register after running chunk j from a zero register and '*' is carry-less
// multiplication modulo the polynomial in the reflected representation (bit 31 = x^0).
// One thread per chunk computes raw_j with 16-byte loads, multiplies by its power of x and the
// warp/CTA xor-reduces into a single 32-bit word.  Algorithmic traffic: n bytes read.
#pragma once
#include "common.cuh"

namespace bz3 {

// a(x) * b(x) mod P, reflected bit order
BZ_HD u32 gf2_mulmod(u32 a, u32 b) {
    u32 p = 0;
    for (int i = 0; i < 32; i++) {
        if (a & 0x80000000u) p ^= b;
        a <<= 1;
        b = (b & 1u) ? (b >> 1) ^ kCrcPoly : (b >> 1);
    }
    return p;
}

// table of x^(8 * 2^k) mod P for k = 0..31, filled by crc_init_tables() on the host
struct CrcTables {
    u32 xpow8[32];  // x^(8*2^k)
    u32 byte_tab[256];
};

inline void crc_fill_tables(CrcTables& t) {
    for (u32 b = 0; b < 256; b++) {
        u32 r = b;
        for (int k = 0; k < 8; k++) r = (r >> 1) ^ ((r & 1u) ? kCrcPoly : 0u);
        t.byte_tab[b] = r;
    }
    u32 x = 0x40000000u;                                 // x^1
    for (int k = 0; k < 3; k++) x = gf2_mulmod(x, x);    // x^8
    for (int k = 0; k < 32; k++) {
        t.xpow8[k] = x;
        x = gf2_mulmod(x, x);
    }
}

// x^(8*nbytes) mod P
BZ_HD u32 crc_xpow_bytes(const u32* xpow8, u64 nbytes) {
    u32 p = 0x80000000u;  // x^0
    for (int k = 0; nbytes && k < 32; k++, nbytes >>= 1)
        if (nbytes & 1) p = gf2_mulmod(xpow8[k], p);
    return p;
}

#if defined(BZ_DEVICE_CODE)
__constant__ CrcTables c_crc;

constexpr int kCrcChunk = 2048;   // bytes per thread
constexpr int kCrcThreads = 128;

BZ_D u32 crc_step4(u32 s, u32 w, const u32* tab) {
    s = tab[(s ^ w) & 0xff] ^ (s >> 8);
    s = tab[(s ^ (w >> 8)) & 0xff] ^ (s >> 8);
    s = tab[(s ^ (w >> 16)) & 0xff] ^ (s >> 8);
    s = tab[(s ^ (w >> 24)) & 0xff] ^ (s >> 8);
    return s;
}

// buf must be 16-byte aligned.  *acc must be zeroed before launch; afterwards
// result = *acc xor init * x^(8n)  (folded in by thread 0 of block 0).
__global__ void __launch_bounds__(kCrcThreads) crc_kernel(const u8* __restrict__ buf, u32 n, u32 init, u32* acc) {
    __shared__ u32 tab[256];
    for (int i = threadIdx.x; i < 256; i += kCrcThreads) tab[i] = c_crc.byte_tab[i];
    __syncthreads();
    const u32 chunk = blockIdx.x * kCrcThreads + threadIdx.x;
    const u64 start = (u64)chunk * kCrcChunk;
    u32 contrib = 0;
    if (start < n) {
        const u32 len = (n - start) < (u64)kCrcChunk ? (u32)(n - start) : (u32)kCrcChunk;
        u32 s = 0;
        const uint4* p4 = reinterpret_cast<const uint4*>(buf + start);
        u32 i = 0;
        for (; i + 16 <= len; i += 16) {
            uint4 v = p4[i >> 4];
            s = crc_step4(s, v.x, tab);
            s = crc_step4(s, v.y, tab);
            s = crc_step4(s, v.z, tab);
            s = crc_step4(s, v.w, tab);
        }
        for (; i < len; i++) s = tab[(s ^ buf[start + i]) & 0xff] ^ (s >> 8);
        contrib = gf2_mulmod(s, crc_xpow_bytes(c_crc.xpow8, (u64)n - start - len));
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) contrib ^= gf2_mulmod(init, crc_xpow_bytes(c_crc.xpow8, n));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) contrib ^= __shfl_xor_sync(kFullMask, contrib, o);
    if (lane_id() == 0 && contrib) atomicXor(acc, contrib);
}

inline cudaError_t crc_upload_tables() {
    CrcTables t;
    crc_fill_tables(t);
    return cudaMemcpyToSymbol(c_crc, &t, sizeof(t));
}

// acc: device u32.  Result is valid after the stream reaches this point.
inline cudaError_t crc_launch(cudaStream_t st, const u8* buf, u32 n, u32 init, u32* acc) {
    BZ_CUDA_TRY(cudaMemsetAsync(acc, 0, sizeof(u32), st));
    u32 chunks = (n + kCrcChunk - 1) / kCrcChunk;
    u32 blocks = (chunks + kCrcThreads - 1) / kCrcThreads;
    if (blocks == 0) blocks = 1;  // n == 0: only the init term
    BZ_LAUNCH(blocks, kCrcThreads, 0, st, crc_kernel)(buf, n, init, acc); BZ_NOTE_LAUNCH();
    return cudaGetLastError();
}
#endif

}  // namespace bz3
