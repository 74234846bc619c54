// emu_check.cpp -- runs the real kernel bodies of bzip3_b200/csrc on the CPU thread-block emulator.
// Test infrastructure only (tests/test_emu_kernels.py); see cta_emu.h for what this can and cannot show.
#define BZ_EMU 1
#include "cta_emu.h"

#include "../../bzip3_b200/csrc/common.cuh"
#include "../../bzip3_b200/csrc/cm.cuh"
#include "../../bzip3_b200/csrc/lzp_parallel.cuh"
#include "../../bzip3_b200/csrc/cm_dec.cuh"
#include "../../bzip3_b200/csrc/cm_enc.cuh"

#define EXPORT extern "C" __attribute__((visibility("default")))

using namespace bz3;

EXPORT void emu_set_schedule(int mode, unsigned long long seed) { emu::set_schedule(mode, seed); }

// variant numbers follow bz3_b200_set_variant(BZ3_STAGE_CM, v) of the library
EXPORT int32_t emu_cm_encode(int variant, const uint8_t* in, int32_t n, uint8_t* out) {
    s32 res = -12345;
    emu::Dim3 g, b;
    switch (variant) {
        case 1:
            b.x = kCmThreads;
            emu::launch(g, b, kCmSmemBytes, [&] { cm_encode_single_kernel(in, n, out, &res); });
            break;
        case 2:
            b.x = kCmEncThreads;
            emu::launch(g, b, kCmEncSmemBytes, [&] { cm_encode_chunked_kernel<1>(in, n, out, &res); });
            break;
        case 0:
            b.x = kCmEncThreads;
            emu::launch(g, b, kCmEncSmemBytes, [&] { cm_encode_chunked_kernel<0>(in, n, out, &res); });
            break;
        case 4:
            b.x = kCmEncThreads;
            emu::launch(g, b, kCmEncSmemBytes, [&] { cm_encode_chunked_kernel<2>(in, n, out, &res); });
            break;
        case 6:
            b.x = kCmEncThreads;
            emu::launch(g, b, kCmEncSmemBytes, [&] { cm_encode_chunked_kernel<3>(in, n, out, &res); });
            break;
        case 10:
            b.x = kCmE2Threads;
            emu::launch(g, b, kCmE2SmemBytes, [&] { cm_encode_kernel(in, n, out, &res); });
            break;
        default:
            return -777;
    }
    return res;
}

EXPORT int emu_cm_decode(int variant, const uint8_t* in, int32_t insize, uint8_t* out, int32_t n) {
    emu::Dim3 g, b;
    switch (variant) {
        case 1:
            b.x = kCmThreads;
            emu::launch(g, b, kCmSmemBytes, [&] { cm_decode_single_kernel(in, insize, out, n); });
            break;
        case 3:
            b.x = kCmDecPathsThreads;
            emu::launch(g, b, kCmDecSmemBytes, [&] { cm_decode_paths_kernel(in, insize, out, n); });
            break;
        case 0:
            b.x = kCmDecThreads;
            emu::launch(g, b, kCmDecSmemBytes, [&] { cm_decode_tree_kernel(in, insize, out, n); });
            break;
        case 4:
            b.x = kCmDecThreads;
            emu::launch(g, b, kCmDecLanesSmemBytes, [&] { cm_decode_lanes_kernel(in, insize, out, n); });
            break;
        case 6:
            b.x = kCmDecW6Threads;
            emu::launch(g, b, kCmDecW6SmemBytes, [&] { cm_decode_walkers_kernel<0, 0>(in, insize, out, n); });
            break;
        case 7:
            b.x = kCmDecW6Threads;
            emu::launch(g, b, kCmDecW6SmemBytes, [&] { cm_decode_walkers_kernel<1, 0>(in, insize, out, n); });
            break;
        case 8:
            b.x = kCmDecW6Threads;
            emu::launch(g, b, kCmDecW6SmemBytes, [&] { cm_decode_walkers_kernel<1, 1>(in, insize, out, n); });
            break;
        case 9:
            b.x = kCmDecW6Threads;
            emu::launch(g, b, kCmDecW6SmemBytes, [&] { cm_decode_walkers_kernel<2, 1>(in, insize, out, n); });
            break;
        case 10:
            b.x = kCmD2Threads;
            emu::launch(g, b, kCmD2SmemBytes, [&] { cm_decode_kernel(in, insize, out, n); });
            break;
        case 5:
            b.x = kCmDecP2Threads;
            emu::launch(g, b, kCmDecP2SmemBytes, [&] { cm_decode_paths2_kernel(in, insize, out, n); });
            break;
        default:
            return -777;
    }
    return 0;
}

EXPORT int32_t emu_lzp_encode(const uint8_t* in, int32_t n, uint8_t* out, int32_t* lut) {
    s32 res = -12345;
    emu::Dim3 g, b;
    b.x = 32;
    memset(lut, 0, sizeof(int32_t) << kLzpSlotsLog2);
    emu::launch(g, b, 0, [&] { lzp_encode_warp_kernel(in, n, out, lut, &res); });
    return res;
}

EXPORT int32_t emu_lzp_encode_pf(const uint8_t* in, int32_t n, uint8_t* out, int32_t* lut) {
    s32 res = -12345;
    emu::Dim3 g, b;
    b.x = 32;
    memset(lut, 0, sizeof(int32_t) << kLzpSlotsLog2);
    emu::launch(g, b, 0, [&] { lzp_encode_warp_pf_kernel(in, n, out, lut, &res); });
    return res;
}

EXPORT int32_t emu_lzp_decode(const uint8_t* in, int32_t n, uint8_t* out, int32_t max, int32_t* lut) {
    s32 res = -12345;
    emu::Dim3 g, b;
    b.x = 32;
    memset(lut, 0, sizeof(int32_t) << kLzpSlotsLog2);
    emu::launch(g, b, 0, [&] { lzp_decode_warp_kernel(in, n, out, max, lut, &res); });
    return res;
}

EXPORT int32_t emu_lzp_decode_bulk(const uint8_t* in, int32_t n, uint8_t* out, int32_t max, int32_t* lut) {
    s32 res = -12345;
    emu::Dim3 g, b;
    b.x = kLzpBulkThreads;
    memset(lut, 0, sizeof(int32_t) << kLzpSlotsLog2);
    emu::launch(g, b, 0, [&] { lzp_decode_bulk_kernel(in, n, out, max, lut, &res); });
    return res;
}
