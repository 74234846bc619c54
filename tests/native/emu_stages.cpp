// emu_stages.cpp -- the multi-kernel stages (CRC, mRLE, LZP encoder, suffix-array BWT, inverse BWT) on the CPU thread-block
// emulator: the REAL host-side launch sequences of bzip3_b200/csrc/*.cuh (scan / radix sort / prefix doubling ...)
// compiled by g++ against tests/native/cta_emu.h, where a launch runs the grid to completion, "device" memory is
// host memory and a stream is synchronous.  Buffers are carved like bz3_api.cu carves its arena.
// Test infrastructure only (tests/test_emu_stages.py).
#define BZ_EMU 1
#include "cta_emu.h"

#include <vector>

#include "../../bzip3_b200/csrc/common.cuh"
#include "../../bzip3_b200/csrc/scan.cuh"
#include "../../bzip3_b200/csrc/radix_sort.cuh"
#include "../../bzip3_b200/csrc/crc.cuh"
#include "../../bzip3_b200/csrc/mrle.cuh"
#include "../../bzip3_b200/csrc/sufsort.cuh"
#include "../../bzip3_b200/csrc/unbwt.cuh"
#include "../../bzip3_b200/csrc/lzp_scan.cuh"

#define EXPORT extern "C" __attribute__((visibility("default")))

using namespace bz3;

namespace {
struct Pool {   // bump allocator over zero-initialised host memory, 256-byte granules like the device arena
    std::vector<std::vector<unsigned char>> blocks;
    template <class T>
    T* take(size_t count) {
        blocks.emplace_back((count * sizeof(T) + 511) & ~(size_t)255, 0);
        return reinterpret_cast<T*>(blocks.back().data());
    }
};
}  // namespace

EXPORT uint32_t emu_stage_crc(const uint8_t* in, uint32_t n, uint32_t init) {
    Pool P;
    u8* buf = P.take<u8>((size_t)n + 64);   // 16-byte aligned copy, like the device buffer
    memcpy(buf, in, n);
    u32* acc = P.take<u32>(4);
    if (crc_upload_tables() != cudaSuccess) return 0xDEADBEEFu;
    if (crc_launch(nullptr, buf, n, init, acc) != cudaSuccess) return 0xDEADBEEFu;
    return *acc;
}

EXPORT int32_t emu_stage_rle_encode(const uint8_t* in, uint32_t n, uint8_t* out) {
    Pool P;
    MrleScratch S;
    S.heads = P.take<u32>((size_t)n + 2);
    S.temp = P.take<u32>(scan_temp_elems(n));
    S.gain = P.take<int>(256);
    S.flagged = P.take<u8>(256);
    S.d_count = P.take<u32>(256);
    S.h_count = P.take<u32>(1024);
    u8* src = P.take<u8>((size_t)n + 64);
    memcpy(src, in, n);
    s32 size = -12345;
    if (mrle_encode(nullptr, src, n, out, S, &size) != cudaSuccess) return -777;
    return size;
}

EXPORT int emu_stage_rle_decode(const uint8_t* in, uint32_t maxin, uint8_t* out, uint32_t outlen) {
    Pool P;
    MrleDecScratch S;
    S.state = P.take<u8>((size_t)maxin + 8);
    S.temp = P.take<u32>(3 * scan_temp_elems(maxin));
    S.flagged = P.take<u8>(256);
    S.d_count = P.take<u32>(256);
    S.h_count = P.take<u32>(1024);
    u8* src = P.take<u8>((size_t)maxin + 64);
    memcpy(src, in, maxin);
    int err = -12345;
    if (mrle_decode(nullptr, src, maxin, out, outlen, S, &err) != cudaSuccess) return -777;
    return err;
}

EXPORT int32_t emu_stage_bwt(const uint8_t* in, uint32_t n, uint8_t* out) {
    Pool P;
    SufsortBuffers B;
    B.sa = P.take<u32>(n);
    B.isa = P.take<u32>((size_t)n + 1);
    for (int i = 0; i < 2; i++) B.key[i] = P.take<u64>(n);
    for (int i = 0; i < 2; i++) B.val[i] = P.take<u32>(n);
    for (int i = 0; i < 2; i++) B.pos[i] = P.take<u32>(n);
    for (int i = 0; i < 2; i++) B.grp[i] = P.take<u32>(n);
    B.temp = P.take<u32>(sufsort_temp_elems(n));
    B.d_count = P.take<u32>(256);
    B.h_count = P.take<u32>(1024);
    u8* src = P.take<u8>((size_t)n + 64);   // zero padding read by the 7-byte key kernel
    memcpy(src, in, n);
    s32 idx = -12345;
    if (suffix_bwt(nullptr, src, n, out, B, &idx) != cudaSuccess) return -777;
    return idx;
}

EXPORT int emu_stage_unbwt(const uint8_t* in, uint32_t n, int32_t idx, uint8_t* out) {
    Pool P;
    UnbwtBuffers B;
    int lg;
    u32 K;
    unbwt_geometry(n, &lg, &K);
    B.psi = P.take<u32>((size_t)n + 2);
    for (int i = 0; i < 2; i++) B.nxt[i] = P.take<u32>((size_t)K + 2);
    for (int i = 0; i < 2; i++) B.dist[i] = P.take<u32>((size_t)K + 2);
    B.len = P.take<u32>((size_t)K + 2);
    B.hist = P.take<u32>(256);
    B.start = P.take<u32>(257);
    B.big = P.take<u32>(65536);
    B.temp = P.take<u32>(rs_temp_elems<u8>(n));
    B.d_count = P.take<u32>(256);
    B.h_count = P.take<u32>(1024);
    u8* src = P.take<u8>((size_t)n + 64);
    memcpy(src, in, n);
    int status = -12345;
    if (unbwt(nullptr, src, n, idx, out, B, &status) != cudaSuccess) return -777;
    return status;
}

EXPORT int32_t emu_stage_lzp_encode(const uint8_t* in, int32_t n, uint8_t* out) {
    if (n < kLzpMinMatch + 32) return -1;
    Pool P;
    LzpScanBuffers B;
    const u32 m = (u32)n - 4u;
    for (int i = 0; i < 2; i++) B.key[i] = P.take<u32>(m);
    for (int i = 0; i < 2; i++) B.idx[i] = P.take<u32>(m);
    B.P = P.take<u32>((size_t)n + 8);
    B.code = P.take<u8>((size_t)n + 8);
    B.lut = P.take<s32>((size_t)kLzpSlots);
    B.temp = P.take<u32>(rs_temp_elems<u32>(m));
    u8* src = P.take<u8>((size_t)n + 64);
    memcpy(src, in, (size_t)n);
    s32* res = P.take<s32>(4);
    *res = -12345;
    if (lzp_scan_encode(nullptr, src, n, out, B, res) != cudaSuccess) return -777;
    return *res;
}
