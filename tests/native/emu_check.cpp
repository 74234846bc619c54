// emu_check.cpp -- runs the real kernel bodies of bzip3_b200/csrc on the CPU thread-block emulator.
// Test infrastructure only (tests/test_emu_kernels.py); see cta_emu.h for what this can and cannot show.
#define BZ_EMU 1
#include "cta_emu.h"

#include "../../bzip3_b200/csrc/common.cuh"
#include "../../bzip3_b200/csrc/cm.cuh"
#include "../../bzip3_b200/csrc/lzp_parallel.cuh"

#define EXPORT extern "C" __attribute__((visibility("default")))

using namespace bz3;

EXPORT void emu_set_schedule(int mode, unsigned long long seed) { emu::set_schedule(mode, seed); }

EXPORT int32_t emu_cm_encode(const uint8_t* in, int32_t n, uint8_t* out) {
    s32 res = -12345;
    emu::Dim3 g, b;
    b.x = kCmEncThreads;
    emu::launch(g, b, kCmEncSmemBytes, [&] { cm_encode_kernel(in, n, out, &res); });
    return res;
}

EXPORT int emu_cm_decode(const uint8_t* in, int32_t insize, uint8_t* out, int32_t n) {
    emu::Dim3 g, b;
    b.x = kCmDecThreads;
    emu::launch(g, b, kCmDecSmemBytes, [&] { cm_decode_kernel(in, insize, out, n); });
    return 0;
}

EXPORT int32_t emu_lzp_decode(const uint8_t* in, int32_t n, uint8_t* out, int32_t max, int32_t* lut) {
    s32 res = -12345;
    emu::Dim3 g, b;
    b.x = kLzpBulkThreads;
    memset(lut, 0, sizeof(int32_t) << kLzpSlotsLog2);
    emu::launch(g, b, 0, [&] { lzp_decode_bulk_kernel(in, n, out, max, lut, &res); });
    return res;
}
