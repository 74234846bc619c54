// host_check.cpp -- compiles the host+device ("BZ_HD") logic of bzip3_b200/csrc with plain g++ so the
// exact code the single-lane kernels execute can be checked against the oracle without a GPU.
// Test infrastructure: built into tests/_build/libhostcheck.so by tests/test_csrc_host.py.
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../bzip3_b200/csrc/common.cuh"
#include "../../bzip3_b200/csrc/crc.cuh"
#include "../../bzip3_b200/csrc/lzp.cuh"
#include "../../bzip3_b200/csrc/cm.cuh"

using namespace bz3;
#define EXPORT extern "C" __attribute__((visibility("default")))

// crc assembled exactly like crc_kernel does: per-chunk raw registers times powers of x
EXPORT uint32_t hc_crc_chunked(const uint8_t* buf, uint32_t n, uint32_t init, uint32_t chunk) {
    CrcTables t;
    crc_fill_tables(t);
    uint32_t acc = gf2_mulmod(init, crc_xpow_bytes(t.xpow8, n));
    for (uint64_t start = 0; start < n; start += chunk) {
        uint32_t len = (n - start) < chunk ? (uint32_t)(n - start) : chunk;
        uint32_t s = 0;
        for (uint32_t i = 0; i < len; i++) s = t.byte_tab[(s ^ buf[start + i]) & 0xff] ^ (s >> 8);
        acc ^= gf2_mulmod(s, crc_xpow_bytes(t.xpow8, (uint64_t)n - start - len));
    }
    return acc;
}
EXPORT int32_t hc_lzp_encode(const uint8_t* in, int32_t n, uint8_t* out) {
    std::vector<int32_t> lut(kLzpSlots, 0);
    return lzp_encode_serial(in, n, out, lut.data());
}
EXPORT int32_t hc_lzp_decode(const uint8_t* in, int32_t n, uint8_t* out, int32_t max) {
    std::vector<int32_t> lut(kLzpSlots, 0);
    return lzp_decode_serial(in, n, out, max, lut.data());
}
static void tables_init(std::vector<uint16_t>& tab) {
    tab.resize(kCmTableU16);
    for (int k = 0; k < kCmTableU16; k++) tab[k] = cm_initial(k);
}
EXPORT int32_t hc_cm_encode(const uint8_t* in, int32_t n, uint8_t* out) {
    std::vector<uint16_t> tab;
    tables_init(tab);
    return cm_encode_serial(cm_tables_at(tab.data()), in, n, out);
}
EXPORT void hc_cm_decode(const uint8_t* in, int32_t insize, uint8_t* out, int32_t n) {
    std::vector<uint16_t> tab;
    tables_init(tab);
    cm_decode_serial(cm_tables_at(tab.data()), in, insize, out, n);
}
