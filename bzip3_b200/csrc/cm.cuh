// cm.cuh -- the context-mixing bit model and the 32-bit carry-less binary range coder.
//
// Restates begin / encode_bytes / decode_bytes (reference src/libbz3.c:333-494).  Everything here
// is format-defining integer arithmetic:
//   model    c0[node] (rate 2), c1[prev byte][node] (rate 4), c2[2*node+runflag][0..16] (SSE, rate 6)
//            p   = ((c0 + c1[prev1]) * 7 + 2 * c1[prev2]) >> 4
//            sse = lo + (((hi - lo) * (p & 4095)) >> 12)       lo,hi = c2 row cells p>>12, (p>>12)+1
//            P   = 3 * sse + p        (18-bit probability that the bit is 1)
//   coder    split = low + ((high - low) * P >> 18); bit ? high = split : low = split + 1;
//            while top bytes of low and high agree: emit / shift.
// The payload of a block is ONE such stream, so the coder recurrence is serial per block (SURVEY.md
// section 7.2 H1).  What can be spread over lanes is the model on the encode side: a tree node of
// depth d is only ever touched by bit position d of a byte, so eight lanes (one per bit position) own
// disjoint counter sets and run the model for a whole byte at once; a second warp drains their
// probabilities through the range coder.  Decoding is one chain (the next context depends on the
// decoded bit); it runs out of shared memory with the 149 KB of tables resident.
#pragma once
#include "common.cuh"

namespace bz3 {

constexpr int kCmC0 = 256;
constexpr int kCmC1 = 256 * 256;
constexpr int kCmC2 = 512 * 17;
constexpr int kCmTableU16 = kCmC0 + kCmC1 + kCmC2;  // 74 496 counters = 148 992 bytes (src/libbz3.c:341)

struct CmTables {
    u16* c0;
    u16* c1;
    u16* c2;
};
BZ_HD CmTables cm_tables_at(u16* base) {
    CmTables t;
    t.c0 = base;
    t.c1 = base + kCmC0;
    t.c2 = base + kCmC0 + kCmC1;
    return t;
}
// initial value of counter #k of the flat table (src/libbz3.c:350-358)
BZ_HD u16 cm_initial(int k) {
    if (k < kCmC0 + kCmC1) return 32768;
    int cell = (k - kCmC0 - kCmC1) % 17;
    return (u16)((cell << 12) - (cell == 16));
}

BZ_HD u32 cm_adapt(u32 v, int bit, int rate) {  // update0 / update1 (src/libbz3.c:347-348)
    return bit ? v + ((v ^ 65535u) >> rate) : v - (v >> rate);
}

struct CmCtx {
    int prev1, prev2, flag;
    u32 run;
};
BZ_HD void cm_ctx_begin_byte(CmCtx& c) {  // run detection, evaluated before each byte (src/libbz3.c:367-372)
    c.run = (c.prev1 == c.prev2) ? c.run + 1 : 0;
    c.flag = c.run > 2;
}
BZ_HD void cm_ctx_end_byte(CmCtx& c, int byte) {
    c.prev2 = c.prev1;
    c.prev1 = byte;
}

// Predict-and-learn for one binary decision whose outcome is already known (encoder side).
BZ_HD u32 cm_code_known_bit(const CmTables& t, int node, const CmCtx& c, int bit) {
    u16* q0 = t.c0 + node;
    u16* q1 = t.c1 + c.prev1 * 256 + node;
    const int a = *q0, b = *q1, d = t.c1[c.prev2 * 256 + node];
    const int p = ((a + b) * 7 + d + d) >> 4;
    u16* row = t.c2 + (2 * node + c.flag) * 17 + (p >> 12);
    const int lo = row[0], hi = row[1];
    const int sse = lo + (((hi - lo) * (p & 4095)) >> 12);
    *q0 = (u16)cm_adapt((u32)a, bit, 2);
    *q1 = (u16)cm_adapt((u32)b, bit, 4);
    row[0] = (u16)cm_adapt((u32)lo, bit, 6);
    row[1] = (u16)cm_adapt((u32)hi, bit, 6);
    return (u32)(sse * 3 + p);
}

struct RangeCoder {
    u32 low, high;
};

BZ_HD s32 cm_encode_serial(const CmTables& t, const u8* in, s32 n, u8* out) {
    RangeCoder rc{0u, 0xFFFFFFFFu};
    CmCtx c{0, 0, 0, 0u};
    s32 op = 0;
    for (s32 i = 0; i < n; i++) {
        cm_ctx_begin_byte(c);
        const int sym = in[i];
        int node = 1;
        for (int k = 7; k >= 0; k--) {
            const int bit = (sym >> k) & 1;
            const u32 P = cm_code_known_bit(t, node, c, bit);
            const u32 split = rc.low + (u32)(((u64)(rc.high - rc.low) * P) >> 18);
            if (bit) rc.high = split; else rc.low = split + 1;
            while ((rc.low ^ rc.high) < (1u << 24)) {
                out[op++] = (u8)(rc.low >> 24);
                rc.low <<= 8;
                rc.high = (rc.high << 8) | 0xFFu;
            }
            node = node * 2 + bit;
        }
        cm_ctx_end_byte(c, sym);
    }
    for (int k = 0; k < 4; k++) {  // flush (src/libbz3.c:425-432)
        out[op++] = (u8)(rc.low >> 24);
        rc.low <<= 8;
    }
    return op;
}

// Reads past the end of the payload behave like read_in(): the int -1 is added, i.e. 0xFFFFFFFF.
BZ_HD u32 cm_next_code_byte(const u8* in, s32& ip, s32 insize) { return ip < insize ? (u32)in[ip++] : 0xFFFFFFFFu; }

BZ_HD void cm_decode_serial(const CmTables& t, const u8* in, s32 insize, u8* out, s32 n) {
    RangeCoder rc{0u, 0xFFFFFFFFu};
    CmCtx c{0, 0, 0, 0u};
    s32 ip = 0;
    u32 code = 0;
    for (int k = 0; k < 4; k++) code = (code << 8) + cm_next_code_byte(in, ip, insize);
    for (s32 i = 0; i < n; i++) {
        cm_ctx_begin_byte(c);
        int node = 1;
        while (node < 256) {
            u16* q0 = t.c0 + node;
            u16* q1 = t.c1 + c.prev1 * 256 + node;
            const int a = *q0, b = *q1, d = t.c1[c.prev2 * 256 + node];
            const int p = ((a + b) * 7 + d + d) >> 4;
            u16* row = t.c2 + (2 * node + c.flag) * 17 + (p >> 12);
            const int lo = row[0], hi = row[1];
            const int sse = lo + (((hi - lo) * (p & 4095)) >> 12);
            const u32 P = (u32)(sse * 3 + p);
            const u32 split = rc.low + (u32)(((u64)(rc.high - rc.low) * P) >> 18);
            const int bit = code <= split;
            if (bit) rc.high = split; else rc.low = split + 1;
            while ((rc.low ^ rc.high) < (1u << 24)) {
                rc.low <<= 8;
                rc.high = (rc.high << 8) | 0xFFu;
                code = (code << 8) + cm_next_code_byte(in, ip, insize);
            }
            *q0 = (u16)cm_adapt((u32)a, bit, 2);
            *q1 = (u16)cm_adapt((u32)b, bit, 4);
            row[0] = (u16)cm_adapt((u32)lo, bit, 6);
            row[1] = (u16)cm_adapt((u32)hi, bit, 6);
            node = node * 2 + bit;
        }
        out[i] = (u8)node;
        cm_ctx_end_byte(c, node & 255);
    }
}

#if defined(__CUDACC__)

constexpr int kCmThreads = 64;
constexpr int kCmRing = 2048;  // probability ring entries (power of two)
constexpr size_t kCmSmemBytes = (size_t)kCmTableU16 * 2 + (size_t)kCmRing * 4 + 64;

BZ_D void cm_tables_init_smem(u16* tab) {
    for (int k = threadIdx.x; k < kCmTableU16; k += blockDim.x) tab[k] = cm_initial(k);
}

// ---- single-lane kernels: the literal chain, tables in shared memory (used as on-device cross-check)
__global__ void __launch_bounds__(kCmThreads) cm_encode_single_kernel(const u8* in, s32 n, u8* out, s32* out_size) {
    extern __shared__ __align__(16) u16 cm_smem[];
    cm_tables_init_smem(cm_smem);
    __syncthreads();
    if (threadIdx.x == 0) *out_size = cm_encode_serial(cm_tables_at(cm_smem), in, n, out);
}
__global__ void __launch_bounds__(kCmThreads) cm_decode_single_kernel(const u8* in, s32 insize, u8* out, s32 n) {
    extern __shared__ __align__(16) u16 cm_smem[];
    cm_tables_init_smem(cm_smem);
    __syncthreads();
    if (threadIdx.x == 0) cm_decode_serial(cm_tables_at(cm_smem), in, insize, out, n);
}

// ---- two-warp encoder: warp 0 = model (8 lanes, one per bit position), warp 1 lane 0 = range coder.
// Ring entry = (P << 1) | bit.  head/tail count entries and only ever grow.
__global__ void __launch_bounds__(kCmThreads) cm_encode_pipelined_kernel(const u8* __restrict__ in, s32 n,
                                                                         u8* __restrict__ out, s32* out_size) {
    extern __shared__ __align__(16) u16 cm_smem[];
    u32* ring = reinterpret_cast<u32*>(cm_smem + kCmTableU16);
    volatile u32* vring = ring;
    volatile u32* head_p = ring + kCmRing;      // entries produced
    volatile u32* tail_p = ring + kCmRing + 1;  // entries consumed
    cm_tables_init_smem(cm_smem);
    if (threadIdx.x == 0) { *head_p = 0; *tail_p = 0; }
    __syncthreads();
    const u32 lane = lane_id();
    const u32 total = (u32)n * 8u;  // n < 2^29
    if (warp_id() == 0) {
        const CmTables t = cm_tables_at(cm_smem);
        CmCtx c{0, 0, 0, 0u};
        u32 head = 0;
        for (s32 i0 = 0; i0 < n; i0 += 32) {
            const int mine = (i0 + (s32)lane < n) ? in[i0 + lane] : 0;
            const int cnt = (n - i0) < 32 ? (n - i0) : 32;
            for (int k = 0; k < cnt; k++) {
                const int sym = __shfl_sync(kFullMask, mine, k);
                cm_ctx_begin_byte(c);
                while (head + 8u - *tail_p > (u32)kCmRing) { /* ring full: wait for the coder */ }
                if (lane < 8) {
                    const int node = (1 << lane) | (sym >> (8 - lane));
                    const int bit = (sym >> (7 - lane)) & 1;
                    const u32 P = cm_code_known_bit(t, node, c, bit);
                    vring[(head + lane) & (kCmRing - 1)] = (P << 1) | (u32)bit;
                    __threadfence_block();
                }
                __syncwarp();
                head += 8;
                if (lane == 0) *head_p = head;
                cm_ctx_end_byte(c, sym);
            }
        }
    } else if (threadIdx.x == 32) {
        RangeCoder rc{0u, 0xFFFFFFFFu};
        s32 op = 0;
        u32 done = 0;
        while (done < total) {
            u32 avail;
            while ((avail = *head_p) == done) { /* wait for the model */ }
            __threadfence_block();
            for (; done < avail; done++) {
                const u32 e = vring[done & (kCmRing - 1)];
                const u32 P = e >> 1;
                const u32 split = rc.low + (u32)(((u64)(rc.high - rc.low) * P) >> 18);
                if (e & 1u) rc.high = split; else rc.low = split + 1;
                while ((rc.low ^ rc.high) < (1u << 24)) {
                    out[op++] = (u8)(rc.low >> 24);
                    rc.low <<= 8;
                    rc.high = (rc.high << 8) | 0xFFu;
                }
            }
            *tail_p = done;
        }
        for (int k = 0; k < 4; k++) {
            out[op++] = (u8)(rc.low >> 24);
            rc.low <<= 8;
        }
        *out_size = op;
    }
}

inline cudaError_t cm_set_smem_attrs() {
    BZ_CUDA_TRY(cudaFuncSetAttribute(cm_encode_single_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCmSmemBytes));
    BZ_CUDA_TRY(cudaFuncSetAttribute(cm_decode_single_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCmSmemBytes));
    BZ_CUDA_TRY(cudaFuncSetAttribute(cm_encode_pipelined_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCmSmemBytes));
    return cudaSuccess;
}

#endif  // __CUDACC__

}  // namespace bz3
