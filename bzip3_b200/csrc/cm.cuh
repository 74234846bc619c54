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
constexpr size_t kCmSmemBytes = (size_t)kCmTableU16 * 2 + 64;

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

// ---- chunked two-warp encoder -----------------------------------------------------------------
// Warp 0 runs the model: lane d (0..7) owns tree depth d, i.e. bit position d of every byte.  A node
// of depth d is only ever coded at bit position d, so the eight lanes touch disjoint counters and
// need no synchronisation with each other.  For a chunk of 1024 bytes they leave 8192 entries
// ((P << 14) | bit) in a shared-memory buffer.  Warp 1, lane 0 is the range coder: it drains the
// previous chunk's buffer with 128-bit shared loads while warp 0 fills the other buffer; one
// __syncthreads per chunk swaps them (no polling, no fences in the hot loops).
// Coder recurrence in (low, range) form:  x = umulhi(range, P << 14)  (== (range * P) >> 18)
//     bit 1: range = x            bit 0: low += x + 1, range -= x + 1
// The top bytes of low and low+range can only agree when range < 2^24, which is the cheap test on
// the critical path; the exact test and the byte output live in the rare slow path.
constexpr int kCmEncChunk = 1024;
constexpr int kCmEncThreads = 64;
constexpr size_t kCmEncSmemBytes = (size_t)kCmTableU16 * 2 + 2 * (size_t)kCmEncChunk * 8 * 4 + 2 * (size_t)kCmEncChunk + 64;

struct RcEnc {
    u32 low, range;
    s32 op;
};
BZ_D void rc_encode_entry(RcEnc& rc, u32 e, u8* __restrict__ out) {
    const u32 x = __umulhi(rc.range, e & 0xFFFFC000u);
    if (e & 1u) {
        rc.range = x;
    } else {
        rc.low += x + 1u;
        rc.range -= x + 1u;
    }
    if (rc.range < (1u << 24)) {
        u32 high = rc.low + rc.range;
        while ((rc.low ^ high) < (1u << 24)) {
            out[rc.op++] = (u8)(rc.low >> 24);
            rc.low <<= 8;
            high = (high << 8) | 0xFFu;
        }
        rc.range = high - rc.low;
    }
}

__global__ void __launch_bounds__(kCmEncThreads) cm_encode_chunked_kernel(const u8* __restrict__ in, s32 n,
                                                                         u8* __restrict__ out, s32* out_size) {
    extern __shared__ __align__(16) u16 cm_smem[];
    u32* pbuf = reinterpret_cast<u32*>(cm_smem + kCmTableU16);           // [2][kCmEncChunk * 8]
    u8* sbytes = reinterpret_cast<u8*>(pbuf + 2 * kCmEncChunk * 8);      // [2][kCmEncChunk]
    cm_tables_init_smem(cm_smem);
    __syncthreads();
    const u32 lane = lane_id();
    const u32 warp = warp_id();
    const s32 nchunks = (n + kCmEncChunk - 1) / kCmEncChunk;
    const CmTables t = cm_tables_at(cm_smem);
    CmCtx c{0, 0, 0, 0u};
    RcEnc rc{0u, 0xFFFFFFFFu, 0};
    for (s32 it = 0; it <= nchunks; it++) {
        if (warp == 0 && it < nchunks) {
            const s32 base = it * kCmEncChunk;
            const s32 len = (n - base) < kCmEncChunk ? (n - base) : kCmEncChunk;
            u8* sb = sbytes + (it & 1) * kCmEncChunk;
            for (s32 k = lane; k < len; k += 32) sb[k] = in[base + k];
            __syncwarp();
            if (lane < 8) {
                u32* pb = pbuf + (it & 1) * (kCmEncChunk * 8) + lane;
                const int sh_node = 8 - (int)lane, sh_bit = 7 - (int)lane, top = 1 << lane;
                for (s32 k = 0; k < len; k++) {
                    const int sym = sb[k];
                    cm_ctx_begin_byte(c);
                    const int node = top | (sym >> sh_node);
                    const int bit = (sym >> sh_bit) & 1;
                    const u32 P = cm_code_known_bit(t, node, c, bit);
                    pb[k * 8] = (P << 14) | (u32)bit;
                    cm_ctx_end_byte(c, sym);
                }
            }
        } else if (warp == 1 && lane == 0 && it > 0) {
            const s32 base = (it - 1) * kCmEncChunk;
            const s32 len = (n - base) < kCmEncChunk ? (n - base) : kCmEncChunk;
            const uint4* pv = reinterpret_cast<const uint4*>(pbuf + ((it - 1) & 1) * (kCmEncChunk * 8));
            uint4 a = pv[0], b = pv[1];
            for (s32 k = 0; k < len; k++) {
                const uint4 ca = a, cb = b;
                if (k + 1 < len) { a = pv[2 * k + 2]; b = pv[2 * k + 3]; }
                rc_encode_entry(rc, ca.x, out);
                rc_encode_entry(rc, ca.y, out);
                rc_encode_entry(rc, ca.z, out);
                rc_encode_entry(rc, ca.w, out);
                rc_encode_entry(rc, cb.x, out);
                rc_encode_entry(rc, cb.y, out);
                rc_encode_entry(rc, cb.z, out);
                rc_encode_entry(rc, cb.w, out);
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 32) {
        for (int k = 0; k < 4; k++) {  // flush (reference src/libbz3.c:425-432)
            out[rc.op++] = (u8)(rc.low >> 24);
            rc.low <<= 8;
        }
        *out_size = rc.op;
    }
}

// ---- tree-parallel decoder ---------------------------------------------------------------------
// Decoding is one dependent chain: the next context depends on the bit just decoded.  What does not
// depend on the bits of the current byte is the probability of every one of the 255 tree nodes (no node
// is visited twice within a byte, and prev1/prev2/runflag are fixed at the byte boundary).  So, per byte:
//   A  128 threads compute P for all 255 nodes (thread t owns nodes t and t+128 and is the only
//      thread that ever reads or writes their counters)                       -> ptab[node]
//   B  every thread walks the same 8-step chain down the tree, fetching the two child probabilities
//      of the next node one step ahead (no table lookups or model arithmetic left on the chain)
//   C  the 8 owners of the visited nodes update their counters from registers.
// One __syncthreads per byte (ptab is double buffered).  The compressed bytes are staged through a
// 2 KiB shared window so the renormalisation never waits on global memory.
constexpr int kCmDecThreads = 128;
constexpr size_t kCmDecSmemBytes = (size_t)kCmTableU16 * 2 + 2 * 256 * 4 + 2048 + 64;

struct NodeCalc {
    int a, b, lo, hi;
    u16* q1;
    u16* row;
};
BZ_D u32 cm_node_predict(const CmTables& t, int node, int prev1, int prev2, int flag, NodeCalc& k) {
    k.q1 = t.c1 + prev1 * 256 + node;
    k.a = t.c0[node];
    k.b = *k.q1;
    const int d = t.c1[prev2 * 256 + node];
    const int p = ((k.a + k.b) * 7 + d + d) >> 4;
    k.row = t.c2 + (2 * node + flag) * 17 + (p >> 12);
    k.lo = k.row[0];
    k.hi = k.row[1];
    const int sse = k.lo + (((k.hi - k.lo) * (p & 4095)) >> 12);
    return (u32)(sse * 3 + p);
}
BZ_D void cm_node_learn(const CmTables& t, int node, const NodeCalc& k, int bit) {
    t.c0[node] = (u16)cm_adapt((u32)k.a, bit, 2);
    *k.q1 = (u16)cm_adapt((u32)k.b, bit, 4);
    k.row[0] = (u16)cm_adapt((u32)k.lo, bit, 6);
    k.row[1] = (u16)cm_adapt((u32)k.hi, bit, 6);
}

__global__ void __launch_bounds__(kCmDecThreads) cm_decode_tree_kernel(const u8* __restrict__ in, s32 insize,
                                                                      u8* __restrict__ out, s32 n) {
    extern __shared__ __align__(16) u16 cm_smem[];
    u32* ptab = reinterpret_cast<u32*>(cm_smem + kCmTableU16);  // [2][256]
    u8* scode = reinterpret_cast<u8*>(ptab + 512);              // [2048] window of the compressed stream
    cm_tables_init_smem(cm_smem);
    const int tid = threadIdx.x;
    for (int k = tid; k < 2048; k += kCmDecThreads) scode[k] = (k < insize) ? in[k] : 0;
    __syncthreads();
    const CmTables t = cm_tables_at(cm_smem);
    const int nodeA = tid, nodeB = tid + 128;
    const int depthA = tid ? 31 - __clz(tid) : 0;
    s32 wlo = 0;  // the window holds stream bytes [wlo, wlo + 2048)
    s32 ip = 0;
    u32 low = 0, range = 0xFFFFFFFFu, code = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const u32 add = (ip < insize) ? (u32)scode[ip & 2047] : 0xFFFFFFFFu;  // read_in() past the end adds -1
        ip += (ip < insize);
        code = (code << 8) + add;
    }
    int prev1 = 0, prev2 = 0;
    u32 run = 0;
    for (s32 i = 0; i < n; i++) {
        run = (prev1 == prev2) ? run + 1 : 0;
        const int flag = run > 2;
        u32* pt = ptab + (i & 1) * 256;
        NodeCalc ka, kb;
        if (tid) pt[nodeA] = cm_node_predict(t, nodeA, prev1, prev2, flag, ka) << 14;
        pt[nodeB] = cm_node_predict(t, nodeB, prev1, prev2, flag, kb) << 14;
        __syncthreads();
        // ---- B: the serial chain, identical in every thread
        int node = 1;
        u32 pcur = pt[1];
        uint2 kids = *reinterpret_cast<const uint2*>(pt + 2);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const u32 x = __umulhi(range, pcur);
            const u32 mid = low + x;
            const bool bit = code <= mid;
            node = node * 2 + (bit ? 1 : 0);
            pcur = bit ? kids.y : kids.x;
            if (k < 6) kids = *reinterpret_cast<const uint2*>(pt + 2 * node);
            range = bit ? x : range - x - 1u;
            low = bit ? low : mid + 1u;
            if (range < (1u << 24)) {
                u32 high = low + range;
                while ((low ^ high) < (1u << 24)) {
                    low <<= 8;
                    high = (high << 8) | 0xFFu;
                    const u32 add = (ip < insize) ? (u32)scode[ip & 2047] : 0xFFFFFFFFu;
                    ip += (ip < insize);
                    code = (code << 8) + add;
                }
                range = high - low;
            }
        }
        const int byte = node & 255;
        // ---- C: owners of the visited nodes learn
        if (tid && nodeA == ((256 | byte) >> (8 - depthA))) cm_node_learn(t, nodeA, ka, (byte >> (7 - depthA)) & 1);
        if (nodeB == (128 | (byte >> 1))) cm_node_learn(t, nodeB, kb, byte & 1);
        if (tid == 0) out[i] = (u8)byte;
        prev2 = prev1;
        prev1 = byte;
        if (ip - wlo >= 1024) {  // uniform: every thread follows the same chain
            __syncthreads();
            for (int k = tid; k < 1024; k += kCmDecThreads) {
                const s32 src = wlo + 2048 + k;
                scode[src & 2047] = (src < insize) ? in[src] : 0;
            }
            wlo += 1024;
            __syncthreads();
        }
    }
}

inline cudaError_t cm_set_smem_attrs() {
    BZ_CUDA_TRY(cudaFuncSetAttribute(cm_encode_single_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCmSmemBytes));
    BZ_CUDA_TRY(cudaFuncSetAttribute(cm_decode_single_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCmSmemBytes));
    BZ_CUDA_TRY(cudaFuncSetAttribute(cm_encode_chunked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCmEncSmemBytes));
    BZ_CUDA_TRY(cudaFuncSetAttribute(cm_decode_tree_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCmDecSmemBytes));
    return cudaSuccess;
}

#endif  // __CUDACC__

}  // namespace bz3
