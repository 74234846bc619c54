// cm_enc.cuh -- entropy encoder of the block codec: encode_bytes (reference src/libbz3.c:360-433) for one block in
// one thread block.
//
// The payload of a block is ONE carry-less range-coder stream, so the coder recurrence is serial per block.  What is
// not serial on the encode side is the model: the symbols are known, and a tree node of depth d is only ever touched by
// bit position d of a byte, so eight lanes (one per bit position) own disjoint counter sets.  Three-stage pipeline over
// chunks of the input, one __syncthreads per chunk, no polling:
//   warp 0, lanes 0..7  stage 1: c0 / c1 counters of tree depth `lane`  -> mixed probability p (16 bit)
//   warp 2, lanes 0..7  stage 2: SSE rows c2 of depth `lane`            -> M = P << 14
//   warp 1, lane 0      stage 3: range coder
// Stage s works on chunk it - s in iteration `it`.
//
// Coder lane (written against the latencies measured on B200, profiles/r01_ubench_b200.log: dependent mul.hi 10
// cycles, add / setp ~3.4, taken branch ~23).  (low, range) form:  x = hi32(range * M)  ( == (range * P) >> 18 ),
//     bit 1: range = x            bit 0: low += x + 1, range -= x + 1
// i.e. one multiply and at most one add on the chain of a decision.  The top bytes of low and low + range can only
// agree when range < 2^24, so that is the only test on the straight-line path; a decision that fails it jumps out to
// the reference's shift loop (:405-409) and the byte is finished in a second unrolled copy entered at the next
// decision -- no redo of anything, and no taken branch in a byte that shifts nothing.
#pragma once
#include "cm.cuh"

namespace bz3 {

#if defined(BZ_DEVICE_CODE)

constexpr int kCmE2Chunk = 768;
constexpr int kCmE2Threads = 96;
constexpr size_t kCmE2SmemBytes = (size_t)kCmTableU16 * 2 + 2 * (size_t)kCmE2Chunk * 8 * 4 + 2 * (size_t)kCmE2Chunk * 8 * 2 + 3 * (size_t)kCmE2Chunk + 64;

struct CmCoder {
    u32 low, range;
    s32 op;
    u8* out;
};

#define BZ_CM_ENC_DECISION(J, MJ)                                      \
    {                                                                  \
        const u32 x = __umulhi(r, (MJ));                               \
        if (sym & (0x80u >> (J))) {                                    \
            r = x;                                                     \
        } else {                                                       \
            l += x + 1u;                                               \
            r -= x + 1u;                                               \
        }                                                              \
    }

// the reference's shift loop (:405-409)
BZ_D void cm_coder_shift(u32& l, u32& r, s32& op, u8* __restrict__ out) {
    u32 high = l + r;
    while ((l ^ high) < (1u << 24)) {
        out[op++] = (u8)(l >> 24);
        l <<= 8;
        high = (high << 8) | 0xFFu;
    }
    r = high - l;
}

// one byte: multipliers of its eight decisions in (a, b)
BZ_D void cm_coder_byte(CmCoder& C, const u32 sym, const uint4 a, const uint4 b) {
    u32 l = C.low, r = C.range;
    int k;
    BZ_CM_ENC_DECISION(0, a.x) if (__builtin_expect(r < (1u << 24), 0)) { k = 0; goto shift; }
    BZ_CM_ENC_DECISION(1, a.y) if (__builtin_expect(r < (1u << 24), 0)) { k = 1; goto shift; }
    BZ_CM_ENC_DECISION(2, a.z) if (__builtin_expect(r < (1u << 24), 0)) { k = 2; goto shift; }
    BZ_CM_ENC_DECISION(3, a.w) if (__builtin_expect(r < (1u << 24), 0)) { k = 3; goto shift; }
    BZ_CM_ENC_DECISION(4, b.x) if (__builtin_expect(r < (1u << 24), 0)) { k = 4; goto shift; }
    BZ_CM_ENC_DECISION(5, b.y) if (__builtin_expect(r < (1u << 24), 0)) { k = 5; goto shift; }
    BZ_CM_ENC_DECISION(6, b.z) if (__builtin_expect(r < (1u << 24), 0)) { k = 6; goto shift; }
    BZ_CM_ENC_DECISION(7, b.w) if (__builtin_expect(r < (1u << 24), 0)) { k = 7; goto shift; }
    C.low = l;
    C.range = r;
    return;
shift:
    for (;;) {   // decision k left range < 2^24
        cm_coder_shift(l, r, C.op, C.out);
        switch (k) {
            case 0: BZ_CM_ENC_DECISION(1, a.y) if (r < (1u << 24)) { k = 1; continue; }
            case 1: BZ_CM_ENC_DECISION(2, a.z) if (r < (1u << 24)) { k = 2; continue; }
            case 2: BZ_CM_ENC_DECISION(3, a.w) if (r < (1u << 24)) { k = 3; continue; }
            case 3: BZ_CM_ENC_DECISION(4, b.x) if (r < (1u << 24)) { k = 4; continue; }
            case 4: BZ_CM_ENC_DECISION(5, b.y) if (r < (1u << 24)) { k = 5; continue; }
            case 5: BZ_CM_ENC_DECISION(6, b.z) if (r < (1u << 24)) { k = 6; continue; }
            case 6: BZ_CM_ENC_DECISION(7, b.w) if (r < (1u << 24)) { k = 7; continue; }
            default: break;
        }
        C.low = l;
        C.range = r;
        return;
    }
}

__global__ void __launch_bounds__(kCmE2Threads, 1) cm_encode_kernel(const u8* __restrict__ in, s32 n, u8* __restrict__ out, s32* out_size) {
    BZ_DYN_SMEM(u16, cm_smem);
    u32* pbuf = reinterpret_cast<u32*>(cm_smem + kCmTableU16);                // [2][chunk * 8]  M = P << 14
    u16* pmid = reinterpret_cast<u16*>(pbuf + 2 * kCmE2Chunk * 8);            // [2][chunk * 8]  p
    u8* sbytes = reinterpret_cast<u8*>(pmid + 2 * kCmE2Chunk * 8);            // [3][chunk]
    cm_tables_init_smem(cm_smem);
    __syncthreads();
    const u32 lane = lane_id();
    const u32 warp = warp_id();
    const s32 nchunks = (n + kCmE2Chunk - 1) / kCmE2Chunk;
    u16* const c0 = cm_smem;
    u16* const c1 = cm_smem + kCmC0;
    u16* const c2 = cm_smem + kCmC0 + kCmC1;
    const int sh_node = 8 - (int)lane, sh_bit = 7 - (int)lane, top = 1 << lane;
    int prev1 = 0, prev2 = 0;   // stage-private copies of the byte context
    u32 run = 0;
    CmCoder C;
    C.low = 0;
    C.range = 0xFFFFFFFFu;
    C.op = 0;
    C.out = out;
#ifdef BZ_CM_PROFILE
    unsigned long long _busy = 0;
#endif
    for (s32 it = 0; it < nchunks + 2; it++) {
#ifdef BZ_CM_PROFILE
        const unsigned long long _tb = clock64();
#endif
        if (warp == 0) {
            if (it < nchunks) {
                const s32 base = it * kCmE2Chunk;
                const s32 len = (n - base) < kCmE2Chunk ? (n - base) : kCmE2Chunk;
                u8* sb = sbytes + (it % 3) * kCmE2Chunk;
                for (s32 k = lane; k < len; k += 32) sb[k] = in[base + k];
                __syncwarp();
                if (lane < 8) {
                    u16* pm = pmid + (it & 1) * (kCmE2Chunk * 8) + lane;
                    // the next symbol is read one byte ahead (its load cannot pass the stores below); two bytes per trip
                    int symn = sb[0];
#pragma unroll 2
                    for (s32 k = 0; k < len; k++) {
                        const int sym = symn;
                        symn = sb[(k + 1 < len) ? k + 1 : k];
                        const int node = top | (sym >> sh_node);
                        const u32 ones = ((sym >> sh_bit) & 1) ? 0xFFFFu : 0u;
                        u16* q0 = c0 + node;
                        u16* q1 = c1 + prev1 * 256 + node;
                        const int a = *q0, b = *q1, d = c1[prev2 * 256 + node];
                        pm[k * 8] = (u16)(((a + b) * 7 + d + d) >> 4);
                        *q0 = (u16)cm_adapt_bf((u32)a, ones, 2);
                        *q1 = (u16)cm_adapt_bf((u32)b, ones, 4);
                        prev2 = prev1;
                        prev1 = sym;
                    }
                }
            }
        } else if (warp == 2) {
            if (lane < 8 && it >= 1 && it - 1 < nchunks) {
                const s32 ch = it - 1;
                const s32 base = ch * kCmE2Chunk;
                const s32 len = (n - base) < kCmE2Chunk ? (n - base) : kCmE2Chunk;
                const u8* sb = sbytes + (ch % 3) * kCmE2Chunk;
                const u16* pm = pmid + (ch & 1) * (kCmE2Chunk * 8) + lane;
                u32* pb = pbuf + (ch & 1) * (kCmE2Chunk * 8) + lane;
                // symbol and mixed probability are read one byte ahead; two bytes per trip
                int symn = sb[0], pn = pm[0];
#pragma unroll 2
                for (s32 k = 0; k < len; k++) {
                    const int sym = symn, p = pn;
                    {
                        const s32 kn = (k + 1 < len) ? k + 1 : k;
                        symn = sb[kn];
                        pn = pm[kn * 8];
                    }
                    run = (prev1 == prev2) ? run + 1 : 0;           // run flag of this byte (src/libbz3.c:367-372)
                    const int flag = run > 2;
                    const int node = top | (sym >> sh_node);
                    const u32 ones = ((sym >> sh_bit) & 1) ? 0xFFFFu : 0u;
                    u16* cell = c2 + (2 * node + flag) * 17 + (p >> 12);
                    const int lo = cell[0], hi = cell[1];
                    const int sse = lo + (((hi - lo) * (p & 4095)) >> 12);
                    pb[k * 8] = (u32)(sse * 3 + p) << 14;
                    cell[0] = (u16)cm_adapt_bf((u32)lo, ones, 6);
                    cell[1] = (u16)cm_adapt_bf((u32)hi, ones, 6);
                    prev2 = prev1;
                    prev1 = sym;
                }
            }
        } else if (warp == 1) {
            if (lane == 0 && it >= 2) {
                const s32 ch = it - 2;
                const s32 base = ch * kCmE2Chunk;
                const s32 len = (n - base) < kCmE2Chunk ? (n - base) : kCmE2Chunk;
                const uint4* pv = reinterpret_cast<const uint4*>(pbuf + (ch & 1) * (kCmE2Chunk * 8));
                const u8* sb = sbytes + (ch % 3) * kCmE2Chunk;
                uint4 a = pv[0], b = pv[1];   // multipliers of byte k, read one byte ahead
                u32 sym = sb[0];
#pragma unroll 2
                for (s32 k = 0; k < len; k++) {
                    const uint4 ca = a, cb = b;
                    const u32 cs = sym;
                    const s32 kn = (k + 1 < len) ? k + 1 : k;   // the last byte re-reads itself
                    a = pv[2 * kn];
                    b = pv[2 * kn + 1];
                    sym = sb[kn];
                    cm_coder_byte(C, cs, ca, cb);
                }
            }
        }
#ifdef BZ_CM_PROFILE
        _busy += clock64() - _tb;
#endif
        __syncthreads();
    }
#ifdef BZ_CM_PROFILE
    if (lane == 0) g_cm_prof[13 + (warp == 0 ? 0 : warp == 2 ? 1 : 2)] = _busy;   // stage1, stage2, coder
#endif
    if (threadIdx.x == 32) {
        for (int k = 0; k < 4; k++) {  // flush (reference src/libbz3.c:425-432)
            C.out[C.op++] = (u8)(C.low >> 24);
            C.low <<= 8;
        }
        *out_size = C.op;
    }
}

#endif  // BZ_DEVICE_CODE

}  // namespace bz3
