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

#if defined(BZ_DEVICE_CODE)

constexpr int kCmThreads = 64;
constexpr size_t kCmSmemBytes = (size_t)kCmTableU16 * 2 + 64;

BZ_D void cm_tables_init_smem(u16* tab) {
    for (int k = threadIdx.x; k < kCmTableU16; k += blockDim.x) tab[k] = cm_initial(k);
}

// ---- single-lane kernels: the literal chain, tables in shared memory (used as on-device cross-check)
__global__ void __launch_bounds__(kCmThreads) cm_encode_single_kernel(const u8* in, s32 n, u8* out, s32* out_size) {
    BZ_DYN_SMEM(u16, cm_smem);
    cm_tables_init_smem(cm_smem);
    __syncthreads();
    if (threadIdx.x == 0) *out_size = cm_encode_serial(cm_tables_at(cm_smem), in, n, out);
}
__global__ void __launch_bounds__(kCmThreads) cm_decode_single_kernel(const u8* in, s32 insize, u8* out, s32 n) {
    BZ_DYN_SMEM(u16, cm_smem);
    cm_tables_init_smem(cm_smem);
    __syncthreads();
    if (threadIdx.x == 0) cm_decode_serial(cm_tables_at(cm_smem), in, insize, out, n);
}

// branch-free counter update: bit ? v + ((v ^ 65535) >> rate) : v - (v >> rate)
BZ_D u32 cm_adapt_bf(u32 v, u32 ones /* bit ? 0xFFFF : 0 */, int rate) {
    const u32 u = (v ^ ones) >> rate;
    return ones ? v + u : v - u;
}

#ifdef BZ_CM_PROFILE
__device__ unsigned long long g_cm_prof[48];
#define BZ_PROF_DECL unsigned long long _t0 = clock64(), _t1
#define BZ_PROF(slot) do { _t1 = clock64(); _acc[slot] += _t1 - _t0; _t0 = _t1; } while (0)
// after a barrier: BAR.SYNC does not block at issue, so make the clock read depend on a post-barrier load
#define BZ_PROF_AFTER_BAR(slot, ptr) do { unsigned _v = *(ptr); asm volatile("mov.u64 %0, %%clock64; // %1" : "=l"(_t1) : "r"(_v)); _acc[slot] += _t1 - _t0; _t0 = _t1; } while (0)
#else
#define BZ_PROF_AFTER_BAR(slot, ptr)
#define BZ_PROF_DECL
#define BZ_PROF(slot)
#endif

// ---- chunked pipelined encoder -----------------------------------------------------------------
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
constexpr int kCmEncChunk = 768;
constexpr int kCmEncThreads = 96;
constexpr size_t kCmEncSmemBytes = (size_t)kCmTableU16 * 2 + 2 * (size_t)kCmEncChunk * 8 * 4 + 2 * (size_t)kCmEncChunk * 8 * 2 + 3 * (size_t)kCmEncChunk + 64;

// mul.hi that the compiler may neither sink into a branch nor drop: it is issued speculatively for the
// NEXT decision before the (rare) renormalisation test of the current one has resolved, which takes the
// test off the critical recurrence  range -> mul.hi -> range.
BZ_D u32 mulhi_pinned(u32 a, u32 b) {
#if defined(BZ_EMU)
    return __umulhi(a, b);
#else
    u32 r;
    asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
#endif
}

// ---- range coder lane -------------------------------------------------------------------------
// Measured on B200 (profiles/r01_ncu_source_cm_*): for one in-order thread a TAKEN branch costs ~28
// cycles and every dependent ALU instruction ~4-5, so the byte is coded in two tiers:
//   fast tier   eight decisions with no branch at all.  No renormalisation is applied; instead the
//               minimum over the eight steps of  low ^ (low + range)  is kept.  If it never dropped below
//               2^24 no byte had to be shifted out and the result is exact.
//   exact tier  otherwise the byte is redone from its saved start state with the reference loop.
// For BWT output most bytes take the fast tier (a byte is shifted out every ~40 decisions).
// Recurrence in (low, range) form:  x = umulhi(range, P << 14)  ( == (range * P) >> 18 ),
//     bit 1: range = x            bit 0: low += x + 1, range -= x + 1
BZ_D void rc_fast_step(u32& low, u32& range, u32& x, u32& tmin, u32 bit, u32 mnext) {
#if defined(BZ_EMU)
    if (bit) range = x; else { low += x + 1u; range -= x + 1u; }
    x = __umulhi(range, mnext);
    const u32 t = low ^ (low + range);
    tmin = t < tmin ? t : tmin;
#else
    asm volatile(
        "{\n\t"
        ".reg .pred pb;\n\t"
        ".reg .u32 hi, t, r0;\n\t"
        "setp.ne.u32 pb, %4, 0;\n\t"
        "sub.u32 r0, %1, %2;\n\t"          // range - x - 1 (one 3-input add)
        "add.u32 r0, r0, -1;\n\t"
        "selp.u32 %1, %2, r0, pb;\n\t"     // bit 1: range = x    bit 0: range -= x + 1
        "mul.hi.u32 r0, %1, %5;\n\t"       // product for the next decision
        "@!pb add.u32 %0, %0, %2;\n\t"     // bit 0: low += x + 1
        "@!pb add.u32 %0, %0, 1;\n\t"
        "mov.u32 %2, r0;\n\t"
        "add.u32 hi, %0, %1;\n\t"
        "xor.b32 t, %0, hi;\n\t"
        "min.u32 %3, %3, t;\n\t"
        "}"
        : "+r"(low), "+r"(range), "+r"(x), "+r"(tmin)
        : "r"(bit), "r"(mnext));
#endif
}

// exact tier, one decision: same recurrence, then the renormalisation of the reference.  range < 2^24 is
// necessary for the top bytes of low and low+range to agree, so that cheap test guards the loop.
BZ_D void rc_exact_step(u32& low, u32& range, u32& x, s32& op, u32 bit, u32 mnext, u8* __restrict__ out) {
    u32 slow;
#if defined(BZ_EMU)
    if (bit) range = x; else { low += x + 1u; range -= x + 1u; }
    x = __umulhi(range, mnext);
    slow = range < 0x1000000u;
#else
    asm volatile(
        "{\n\t"
        ".reg .pred pb, ps;\n\t"
        ".reg .u32 nx;\n\t"
        "setp.ne.u32 pb, %4, 0;\n\t"
        "not.b32 nx, %2;\n\t"
        "@pb mov.u32 %1, %2;\n\t"
        "@!pb add.u32 %1, %1, nx;\n\t"
        "mul.hi.u32 %2, %1, %5;\n\t"
        "@!pb sub.u32 %0, %0, nx;\n\t"
        "setp.lt.u32 ps, %1, 0x1000000;\n\t"
        "selp.u32 %3, 1, 0, ps;\n\t"
        "}"
        : "+r"(low), "+r"(range), "+r"(x), "=r"(slow)
        : "r"(bit), "r"(mnext));
#endif
    if (slow) {
        u32 high = low + range;
        while ((low ^ high) < (1u << 24)) {
            out[op++] = (u8)(low >> 24);
            low <<= 8;
            high = (high << 8) | 0xFFu;
        }
        range = high - low;
        x = mulhi_pinned(range, mnext);
    }
}

// exact tier, whole byte from memory (cross-check variant only)
__device__ __noinline__ uint4 rc_exact_byte(u32 low, u32 range, s32 op, u32 sym, const u32* m, u8* out) {
    u32 high = low + range;
    for (int j = 0; j < 8; j++) {
        const u32 x = __umulhi(high - low, m[j]);
        if ((sym << j) & 0x80u) high = low + x; else low += x + 1u;
        while ((low ^ high) < (1u << 24)) {
            out[op++] = (u8)(low >> 24);
            low <<= 8;
            high = (high << 8) | 0xFFu;
        }
    }
    return make_uint4(low, high - low, (u32)op, 0u);
}

// 32 x 32 -> 64 bit product (one IMAD.WIDE): high half = new range, low half = exactness check
BZ_D u64 cm_mul_wide(u32 a, u32 b) {
#if defined(BZ_EMU)
    return (u64)a * b;
#else
    u64 d;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(d) : "r"(a), "r"(b));
    return d;
#endif
}

// MODE 2 coder lane: one multiply per decision.  The SSE stage stored  m = bit ? M : -M  (M = P << 14), and
//     bit 1:  new range = x = hi32(range * M)
//     bit 0:  new range = range - x - 1 = hi32(range * -M)        unless lo32(range * M) == 0
// so the recurrence is eight dependent IMAD.WIDE per byte; low moves by (range_k - range_{k+1}) at 0-bits.
// Ranges only shrink, so "no decision of this byte needed a shift" is implied by  range_8 >= 2^24  (necessary
// condition for a shift: range < 2^24); otherwise, or when a low half was zero, the byte is redone from its
// start state by the reference loop.
BZ_D void rc_byte2(u32& low, u32& range, s32& op, const u32 sym, const uint4 ca, const uint4 cb, u8* __restrict__ out
#ifdef BZ_CM_PROFILE
                   , unsigned long long* _ex
#endif
) {
    const u32 m[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
    u32 rk[9];
    u32 r = range, l = low, zmin = 0xFFFFFFFFu;
    rk[0] = r;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const u64 w = cm_mul_wide(r, m[k]);
        const u32 rn = (u32)(w >> 32);
        zmin = min(zmin, (u32)w);
        if (!(sym & (0x80u >> k))) l += r - rn;
        r = rn;
        rk[k + 1] = rn;
    }
    if (r >= (1u << 24) && zmin != 0u) {
        low = l;
        range = r;
        return;
    }
    if (zmin != 0u) {
        // range < 2^24 is only necessary for a shift (low and high may straddle a top-byte boundary for a
        // while): test the reference's condition after every decision before giving up on the fast result
        u32 a = low, tmin = 0xFFFFFFFFu;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (!(sym & (0x80u >> k))) a += rk[k] - rk[k + 1];
            tmin = min(tmin, a ^ (a + rk[k + 1]));
        }
        if (tmin >= (1u << 24)) {
            low = l;
            range = r;
            return;
        }
    }
#ifdef BZ_CM_PROFILE
    const unsigned long long _e0 = clock64();
#endif
    // exact tier (reference src/libbz3.c:388-416)
    u32 high = low + range;
    l = low;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const bool bit = (sym & (0x80u >> k)) != 0;
        const u32 mk = bit ? m[k] : 0u - m[k];
        const u32 x = __umulhi(high - l, mk);
        if (bit) high = l + x; else l += x + 1u;
        while ((l ^ high) < (1u << 24)) {
            out[op++] = (u8)(l >> 24);
            l <<= 8;
            high = (high << 8) | 0xFFu;
        }
    }
    low = l;
    range = high - l;
#ifdef BZ_CM_PROFILE
    _ex[0] += clock64() - _e0;
    _ex[1] += 1;
#endif
}

// MODE 3 coder lane: one multiply per decision, exact shift test per decision, resume after the first event.
// Measured on B200 (profiles/r01_ubench_b200.log, r01_cm_phase_cycles_1MiB.log): a dependent IMAD.WIDE costs
// 10 cycles, a taken branch ~20, and redoing a whole byte in the exact tier ~700.  So: the eight decisions
// of a byte run without any branch -- product of decision k+1 issued from the new range right away, the
// reference's shift test (low ^ high < 2^24) and the exactness test of the -M shortcut (zero low half) of
// decision k folded into a bit mask next to it -- and ONE branch per byte asks whether anything happened.
// If so, the decisions before the first event stand; from that decision on the byte is finished with the
// reference loop (about once per 4-5 bytes of BWT text, ~4 decisions on average).
BZ_D void rc_byte3(u32& low, u32& range, u64& w, s32& op, const u32 sym, const uint4 ca, const uint4 cb, const u32 mfirst_next,
                   const u32* __restrict__ mrow, u8* __restrict__ out) {
    const u32 m[9] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w, mfirst_next};
    u32 rk[9];   // range before decision k (rk[8]: after the byte)
    u32 l = low, tmin = 0xFFFFFFFFu, zmin = 0xFFFFFFFFu;
    u64 ww = w;
    rk[0] = range;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const u32 rn = (u32)(ww >> 32);
        zmin = min(zmin, (u32)ww);
        ww = cm_mul_wide(rn, m[k + 1]);
        if (!(sym & (0x80u >> k))) l += rk[k] - rn;
        rk[k + 1] = rn;
        tmin = min(tmin, l ^ (l + rn));
    }
    if (tmin < (1u << 24) || zmin == 0u) {
        // something happened: find the first decision with an event (cheap: the ranges are known) ...
        l = low;
        int kf = 0;
        for (; kf < 8; kf++) {
            u32 rb, ra;   // rk[kf], rk[kf + 1] without dynamic register indexing
            rb = rk[0]; ra = rk[1];
#pragma unroll
            for (int j = 1; j < 8; j++) {
                rb = (j == kf) ? rk[j] : rb;
                ra = (j == kf) ? rk[j + 1] : ra;
            }
            const bool bit = (sym & (0x80u >> kf)) != 0;
            const u32 ln = bit ? l : l + (rb - ra);
            const u32 mk = mrow[kf];
            if (((ln ^ (ln + ra)) < (1u << 24)) || (u32)(rb * mk) == 0u) break;   // lo32 of the product
            l = ln;
        }
        // ... the decisions before it stand; from it on the byte is finished with the reference loop
        u32 r = rk[0];
#pragma unroll
        for (int j = 1; j < 8; j++) r = (j == kf) ? rk[j] : r;
        u32 high = l + r;
        for (int k = kf; k < 8; k++) {   // src/libbz3.c:388-416
            const bool bit = (sym & (0x80u >> k)) != 0;
            const u32 mk = mrow[k];
            const u32 x = __umulhi(high - l, bit ? mk : 0u - mk);
            if (bit) high = l + x; else l += x + 1u;
            while ((l ^ high) < (1u << 24)) {
                out[op++] = (u8)(l >> 24);
                l <<= 8;
                high = (high << 8) | 0xFFu;
            }
        }
        low = l;
        range = high - l;
        w = cm_mul_wide(range, mfirst_next);
        return;
    }
    low = l;
    range = rk[8];
    w = ww;
}

// The coder lane of MODE 3 over one chunk.  Two nested loops on purpose: the inner loop runs over bytes without an
// event and its only taken branch is the back-edge (a taken branch costs ~20 cycles, a skipped-over block is a
// taken branch); an event leaves it through a rarely taken exit, is handled, and the inner loop is re-entered.
BZ_D void rc_lane3(const u32* __restrict__ pw, const u8* __restrict__ sb, const s32 len, u32& low, u32& range, s32& op,
                   u8* __restrict__ out) {
    const uint4* pv = reinterpret_cast<const uint4*>(pw);
    s32 k = 0;
    uint4 a = pv[0], b = pv[1];   // multipliers of byte k
    u32 sym = sb[0];
    u32 l = low, r = range;
    u64 w = cm_mul_wide(r, a.x);  // product of the first decision of byte k
    u32 rk[9];
    for (;;) {
        bool event = false;
#pragma unroll 2
        while (k < len) {
            const s32 kn = (k + 1 < len) ? k + 1 : k;   // the last byte re-reads itself; that product is unused
            const uint4 na = pv[2 * kn], nb = pv[2 * kn + 1];
            const u32 nsym = sb[kn];
            const u32 m[9] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, na.x};
            u32 ln = l, tmin = 0xFFFFFFFFu, zmin = 0xFFFFFFFFu;
            u64 ww = w;
            rk[0] = r;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const u32 rn = (u32)(ww >> 32);
                zmin = min(zmin, (u32)ww);
                ww = cm_mul_wide(rn, m[j + 1]);
                if (!(sym & (0x80u >> j))) ln += rk[j] - rn;
                rk[j + 1] = rn;
                tmin = min(tmin, ln ^ (ln + rn));
            }
            if (tmin < (1u << 24) || zmin == 0u) {
                event = true;
                break;
            }
            l = ln;
            r = rk[8];
            w = ww;
            a = na;
            b = nb;
            sym = nsym;
            k++;
        }
        if (!event) break;
        // byte k had an event: find the first decision with one (cheap: the ranges are known) ...
        int kf = 0;
        for (; kf < 8; kf++) {
            u32 rb = rk[0], ra = rk[1];   // rk[kf], rk[kf + 1] without dynamic register indexing
#pragma unroll
            for (int j = 1; j < 8; j++) {
                rb = (j == kf) ? rk[j] : rb;
                ra = (j == kf) ? rk[j + 1] : ra;
            }
            const bool bit = (sym & (0x80u >> kf)) != 0;
            const u32 ln = bit ? l : l + (rb - ra);
            if (((ln ^ (ln + ra)) < (1u << 24)) || (u32)(rb * pw[8 * k + kf]) == 0u) break;   // lo32 of the product
            l = ln;
        }
        // ... the decisions before it stand; from it on the byte is finished with the reference loop
        u32 rr = rk[0];
#pragma unroll
        for (int j = 1; j < 8; j++) rr = (j == kf) ? rk[j] : rr;
        u32 high = l + rr;
        for (int j = kf; j < 8; j++) {   // src/libbz3.c:388-416
            const bool bit = (sym & (0x80u >> j)) != 0;
            const u32 mk = pw[8 * k + j];
            const u32 x = __umulhi(high - l, bit ? mk : 0u - mk);
            if (bit) high = l + x; else l += x + 1u;
            while ((l ^ high) < (1u << 24)) {
                out[op++] = (u8)(l >> 24);
                l <<= 8;
                high = (high << 8) | 0xFFu;
            }
        }
        r = high - l;
        k++;
        if (k >= len) break;
        a = pv[2 * k];
        b = pv[2 * k + 1];
        sym = sb[k];
        w = cm_mul_wide(r, a.x);
    }
    low = l;
    range = r;
}

// Three-stage chunk pipeline (one __syncthreads per chunk, no polling):
//   warp 0, lanes 0..7  stage 1: c0 / c1 counters of tree depth `lane`  -> mixed probability p (16 bit)
//   warp 2, lanes 0..7  stage 2: SSE rows c2 of depth `lane`            -> P << 14
//   warp 1, lane 0      stage 3: range coder
// Stage s works on chunk it-s in iteration `it`.  A node of depth d is only ever coded at bit position d,
// so the eight lanes of a stage own disjoint counters and never synchronise with each other.
template <int MODE>
__global__ void __launch_bounds__(kCmEncThreads, (MODE == 3) ? 1 : 0) cm_encode_chunked_kernel(const u8* __restrict__ in, s32 n,
                                                                         u8* __restrict__ out, s32* out_size) {
    BZ_DYN_SMEM(u16, cm_smem);
    u32* pbuf = reinterpret_cast<u32*>(cm_smem + kCmTableU16);                 // [2][chunk * 8]  P << 14
    u16* pmid = reinterpret_cast<u16*>(pbuf + 2 * kCmEncChunk * 8);            // [2][chunk * 8]  p
    u8* sbytes = reinterpret_cast<u8*>(pmid + 2 * kCmEncChunk * 8);            // [3][chunk]
    cm_tables_init_smem(cm_smem);
    __syncthreads();
    const u32 lane = lane_id();
    const u32 warp = warp_id();
    const s32 nchunks = (n + kCmEncChunk - 1) / kCmEncChunk;
    u16* const c0 = cm_smem;
    u16* const c1 = cm_smem + kCmC0;
    u16* const c2 = cm_smem + kCmC0 + kCmC1;
    const int sh_node = 8 - (int)lane, sh_bit = 7 - (int)lane, top = 1 << lane;
    int prev1 = 0, prev2 = 0;   // stage-private copies of the byte context
    u32 run = 0;
    u32 low = 0, range = 0xFFFFFFFFu;
    s32 op = 0;
#ifdef BZ_CM_PROFILE
    unsigned long long _busy = 0;
    unsigned long long _ex[2] = {0, 0};
#endif
    for (s32 it = 0; it < nchunks + 2; it++) {
#ifdef BZ_CM_PROFILE
        const unsigned long long _tb = clock64();
#endif
        if (warp == 0) {
            if (it < nchunks) {
                const s32 base = it * kCmEncChunk;
                const s32 len = (n - base) < kCmEncChunk ? (n - base) : kCmEncChunk;
                u8* sb = sbytes + (it % 3) * kCmEncChunk;
                for (s32 k = lane; k < len; k += 32) sb[k] = in[base + k];
                __syncwarp();
                if (lane < 8) {
                    u16* pm = pmid + (it & 1) * (kCmEncChunk * 8) + lane;
                    if (MODE == 3) {
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
                    } else {
                        for (s32 k = 0; k < len; k++) {
                            const int sym = sb[k];
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
            }
        } else if (warp == 2) {
            if (lane < 8 && it >= 1 && it - 1 < nchunks) {
                const s32 ch = it - 1;
                const s32 base = ch * kCmEncChunk;
                const s32 len = (n - base) < kCmEncChunk ? (n - base) : kCmEncChunk;
                const u8* sb = sbytes + (ch % 3) * kCmEncChunk;
                const u16* pm = pmid + (ch & 1) * (kCmEncChunk * 8) + lane;
                u32* pb = pbuf + (ch & 1) * (kCmEncChunk * 8) + lane;
                if (MODE == 3) {
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
                        const u32 m = (u32)(sse * 3 + p) << 14;
                        pb[k * 8] = ones ? m : 0u - m;   // the coder lane multiplies by M for a 1-bit and by -M for a 0-bit
                        cell[0] = (u16)cm_adapt_bf((u32)lo, ones, 6);
                        cell[1] = (u16)cm_adapt_bf((u32)hi, ones, 6);
                        prev2 = prev1;
                        prev1 = sym;
                    }
                } else {
                    for (s32 k = 0; k < len; k++) {
                        const int sym = sb[k];
                        const int p = pm[k * 8];
                        run = (prev1 == prev2) ? run + 1 : 0;           // run flag of this byte (src/libbz3.c:367-372)
                        const int flag = run > 2;
                        const int node = top | (sym >> sh_node);
                        const u32 ones = ((sym >> sh_bit) & 1) ? 0xFFFFu : 0u;
                        u16* cell = c2 + (2 * node + flag) * 17 + (p >> 12);
                        const int lo = cell[0], hi = cell[1];
                        const int sse = lo + (((hi - lo) * (p & 4095)) >> 12);
                        {
                            const u32 m = (u32)(sse * 3 + p) << 14;
                            // MODE 2: the coder lane multiplies by M for a 1-bit and by -M for a 0-bit (see rc_byte2)
                            pb[k * 8] = (MODE == 2 && !ones) ? 0u - m : m;
                        }
                        cell[0] = (u16)cm_adapt_bf((u32)lo, ones, 6);
                        cell[1] = (u16)cm_adapt_bf((u32)hi, ones, 6);
                        prev2 = prev1;
                        prev1 = sym;
                    }
                }
            }
        } else if (warp == 1) {
            if (lane == 0 && it >= 2) {
                const s32 ch = it - 2;
                const s32 base = ch * kCmEncChunk;
                const s32 len = (n - base) < kCmEncChunk ? (n - base) : kCmEncChunk;
                const u32* pw = pbuf + (ch & 1) * (kCmEncChunk * 8);
                const uint4* pv = reinterpret_cast<const uint4*>(pw);
                const u8* sb = sbytes + (ch % 3) * kCmEncChunk;
                if (MODE == 3) {
                    rc_lane3(pw, sb, len, low, range, op, out);
                }
                uint4 a = pv[0], b = pv[1];
                u32 sym = sb[0];
                u32 x = mulhi_pinned(range, a.x);
                for (s32 k = 0; MODE != 3 && k < len; k++) {
                    const uint4 ca = a, cb = b;
                    const u32 cs = sym;
                    const s32 kn = (k + 1 < len) ? k + 1 : k;  // the last byte re-reads itself; that product is unused
                    {   // pinned prefetch of the next byte's entries: issued before this byte's decisions, not after
#if defined(BZ_EMU)
                        a = pv[2 * kn];
                        b = pv[2 * kn + 1];
                        sym = sb[kn];
#else
                        const u32 ap = (u32)__cvta_generic_to_shared(pv + 2 * kn);
                        const u32 sp = (u32)__cvta_generic_to_shared(sb + kn);
                        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "r"(ap));
                        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+16];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "r"(ap));
                        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(sym) : "r"(sp));
#endif
                    }
                    if (MODE == 2) {
#ifdef BZ_CM_PROFILE
                        rc_byte2(low, range, op, cs, ca, cb, out, _ex);
#else
                        rc_byte2(low, range, op, cs, ca, cb, out);
#endif
                    } else if (MODE == 0) {
                        const u32 low0 = low, range0 = range, x0 = x;
                        u32 tmin = 0xFFFFFFFFu;
                        rc_fast_step(low, range, x, tmin, cs & 0x80u, ca.y);
                        rc_fast_step(low, range, x, tmin, cs & 0x40u, ca.z);
                        rc_fast_step(low, range, x, tmin, cs & 0x20u, ca.w);
                        rc_fast_step(low, range, x, tmin, cs & 0x10u, cb.x);
                        rc_fast_step(low, range, x, tmin, cs & 0x08u, cb.y);
                        rc_fast_step(low, range, x, tmin, cs & 0x04u, cb.z);
                        rc_fast_step(low, range, x, tmin, cs & 0x02u, cb.w);
                        rc_fast_step(low, range, x, tmin, cs & 0x01u, a.x);
                        if (tmin < (1u << 24)) {  // some decision needed a shift: redo this byte exactly
                            low = low0;
                            range = range0;
                            x = x0;
                            rc_exact_step(low, range, x, op, cs & 0x80u, ca.y, out);
                            rc_exact_step(low, range, x, op, cs & 0x40u, ca.z, out);
                            rc_exact_step(low, range, x, op, cs & 0x20u, ca.w, out);
                            rc_exact_step(low, range, x, op, cs & 0x10u, cb.x, out);
                            rc_exact_step(low, range, x, op, cs & 0x08u, cb.y, out);
                            rc_exact_step(low, range, x, op, cs & 0x04u, cb.z, out);
                            rc_exact_step(low, range, x, op, cs & 0x02u, cb.w, out);
                            rc_exact_step(low, range, x, op, cs & 0x01u, a.x, out);
                        }
                    } else {
                        const uint4 r = rc_exact_byte(low, range, op, cs, pw + 8 * k, out);
                        low = r.x;
                        range = r.y;
                        op = (s32)r.z;
                    }
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
    if (threadIdx.x == 32) {
        g_cm_prof[32] = _ex[0];   // MODE 2: cycles and bytes in the exact tier of the coder lane
        g_cm_prof[33] = _ex[1];
    }
#endif
    if (threadIdx.x == 32) {
        for (int k = 0; k < 4; k++) {  // flush (reference src/libbz3.c:425-432)
            out[op++] = (u8)(low >> 24);
            low <<= 8;
        }
        *out_size = op;
    }
}

// probability table of one byte step.  LAYOUT 0: u32 M = P << 14 per node (serial walk).  LAYOUT 1: {M, -M}
// per node, so that a lane of the lane-parallel walk fetches the multiplier of ITS branch directly.
template <int LAYOUT>
BZ_D void cm_ptab_put(u32* ptab, int idx, u32 m) {
    if (LAYOUT == 0) ptab[idx] = m;
    else reinterpret_cast<uint2*>(ptab)[idx] = make_uint2(m, 0u - m);
}

// model thread of the tree decoders: owner of tree node `node` (0 is a dummy).
// PROTO 0: "byte ready" is a plain barrier.  PROTO 1 (walker kernel): it is a barrier-OR that tells whether a
// walker published the byte; if not, the walkers run up to two more barriers (exact per-path test, serial redo)
// in which the model threads simply take part.
template <int LAYOUT, int PROTO = 0>
BZ_D void cm_dec_model_thread(u16* cm_smem, u32* ptab, volatile u32* vbyte, s32 n, const int node) {
    // ------------------------------------------------------------------ model thread
    // Owns one node; its counters are carried in registers (only this thread writes them).  While the
    // chain warp walks byte i the thread (1) computes both outcomes of its pending update and
    // (2) SPECULATES that byte i repeats byte i-1 -- the common case in BWT output -- and predicts
    // byte i+1 under that hypothesis into the other half of ptab.  On a hit the chain continues at
    // once (no predict phase, no second barrier); on a miss the speculation is simply overwritten.
    const int sh = node ? 8 - (31 - __clz(node)) : 8;                 // (256|byte) >> sh == node <=> on the path
    u16* const q0 = cm_smem + node;
    u16* const c1col = cm_smem + kCmC0 + node;                        // + prev * 256
    u16* const rows = cm_smem + kCmC0 + kCmC1 + (2 * node) * 17;      // + flag * 17 + cell
    int prev1 = 0, prev2 = 0;
    u32 run = 0;
    u16* q1 = c1col;
    u32 a = *q0, b = *q1, d = *q1;
    u32 lo = 0, hi = 0;
    u16* cell = rows;
    bool have = false;   // ptab of the current byte was already produced by the speculation
    for (s32 i = 0; i < n; i++) {
        if (!have) {
            run = (prev1 == prev2) ? run + 1 : 0;
            const int flag = run > 2;
            // (A) predict byte i
            const u32 p = ((a + b) * 7 + d + d) >> 4;
            cell = rows + flag * 17 + (p >> 12);
            lo = cell[0];
            hi = cell[1];
            const int sse = (int)lo + ((((int)hi - (int)lo) * (int)(p & 4095)) >> 12);
            cm_ptab_put<LAYOUT>(ptab, (i & 1) * 256 + node, (u32)(sse * 3 + (int)p) << 14);   // slot 0 is never read
            __syncthreads();   // ptab ready
        }
        // both outcomes of the update of byte i
        const u32 a0 = cm_adapt_bf(a, 0u, 2), a1 = cm_adapt_bf(a, 0xFFFFu, 2);
        const u32 b0 = cm_adapt_bf(b, 0u, 4), b1 = cm_adapt_bf(b, 0xFFFFu, 4);
        const u32 l0 = cm_adapt_bf(lo, 0u, 6), l1 = cm_adapt_bf(lo, 0xFFFFu, 6);
        const u32 h0 = cm_adapt_bf(hi, 0u, 6), h1 = cm_adapt_bf(hi, 0xFFFFu, 6);
        // speculation: byte i == prev1.  Then prev1' = prev2' = prev1, both order-1 inputs of byte i+1 are
        // this thread's current order-1 counter (updated if the node is on the path of prev1).
        const u32 hyp = (u32)prev1;
        const bool on_h = node != 0 && ((256u | hyp) >> sh) == (u32)node;
        const bool one_h = ((hyp >> (sh - 1)) & 1u) != 0;
        const u32 a_s = on_h ? (one_h ? a1 : a0) : a;
        const u32 b_s = on_h ? (one_h ? b1 : b0) : b;
        const u32 run_s = run + 1u;   // run rule (src/libbz3.c:367-370) applied to (prev1, prev1)
        const int flag_s = run_s > 2;
        const u32 p_s = ((a_s + b_s) * 7 + b_s + b_s) >> 4;
        u16* const cell_s = rows + flag_s * 17 + (p_s >> 12);
        u32 lo_s = cell_s[0], hi_s = cell_s[1];
        {
            const u32 nl = one_h ? l1 : l0, nh = one_h ? h1 : h0;   // what byte i would leave in cell[0], cell[1]
            const bool same = on_h && cell_s == cell, up = on_h && cell_s == cell + 1, dn = on_h && cell_s + 1 == cell;
            lo_s = same ? nl : (up ? nh : lo_s);
            hi_s = same ? nh : (dn ? nl : hi_s);
        }
        {
            const int sse = (int)lo_s + ((((int)hi_s - (int)lo_s) * (int)(p_s & 4095)) >> 12);
            cm_ptab_put<LAYOUT>(ptab, ((i + 1) & 1) * 256 + node, (u32)(sse * 3 + (int)p_s) << 14);
        }
        if (PROTO == 0) {
            __syncthreads();   // byte ready
        } else {
            if (!__syncthreads_or(0))
                if (!__syncthreads_or(0)) __syncthreads();
        }
        const u32 byte = vbyte[i & 1];
        const bool on = node != 0 && ((256u | byte) >> sh) == (u32)node;
        const bool one = ((byte >> (sh - 1)) & 1u) != 0;
        const u32 na = one ? a1 : a0, nb = one ? b1 : b0;
        if (on) {   // (C) learn byte i
            *q0 = (u16)na;
            *q1 = (u16)nb;
            cell[0] = (u16)(one ? l1 : l0);
            cell[1] = (u16)(one ? h1 : h0);
        }
        have = byte == hyp;   // uniform across the CTA
        if (have) {
            a = a_s;
            b = b_s;
            d = b_s;
            lo = lo_s;
            hi = hi_s;
            cell = cell_s;
            run = run_s;
            prev2 = prev1;   // == byte
        } else {
            a = on ? na : a;
            d = on ? nb : b;              // this byte's order-1 counter is the next byte's prev2 counter
            prev2 = prev1;
            prev1 = (int)byte;
            q1 = c1col + prev1 * 256;
            b = *q1;                      // after the store above in program order
        }
    }
}

// Same model thread with fewer instructions per byte (variant 7).  On B200 the walker kernel is bound by
// instruction issue, not by its latency chain: every scheduler hosts two walker warps and two model warps, and
// the loop above is ~110 SASS instructions per byte.  Here the update outcome is computed only for the
// hypothesised byte and only by the threads on its path (eight of 255, so half of the warps skip the block
// entirely); the outcome for any other byte is computed after the fact, on a miss, again only on the path.
template <int LAYOUT, int PROTO>
BZ_D void cm_dec_model_thread_slim(u16* cm_smem, u32* ptab, volatile u32* vbyte, s32 n, const int node) {
    const int sh = node ? 8 - (31 - __clz(node)) : 8;                 // (256|byte) >> sh == node <=> on the path
    u16* const q0 = cm_smem + node;
    u16* const c1col = cm_smem + kCmC0 + node;                        // + prev * 256
    u16* const rows = cm_smem + kCmC0 + kCmC1 + (2 * node) * 17;      // + flag * 17 + cell
    int prev1 = 0, prev2 = 0;
    u32 run = 0;
    u16* q1 = c1col;
    u32 a = *q0, b = *q1, d = *q1;
    u32 lo = 0, hi = 0;
    u16* cell = rows;
    bool have = false;   // ptab of the current byte was already produced by the speculation
    for (s32 i = 0; i < n; i++) {
        if (!have) {
            run = (prev1 == prev2) ? run + 1 : 0;
            const int flag = run > 2;
            const u32 p = ((a + b) * 7 + d + d) >> 4;
            cell = rows + flag * 17 + (p >> 12);
            lo = cell[0];
            hi = cell[1];
            const int sse = (int)lo + ((((int)hi - (int)lo) * (int)(p & 4095)) >> 12);
            cm_ptab_put<LAYOUT>(ptab, (i & 1) * 256 + node, (u32)(sse * 3 + (int)p) << 14);
            __syncthreads();   // ptab ready
        }
        // speculation: byte i == prev1
        const u32 hyp = (u32)prev1;
        const bool on_h = node != 0 && ((256u | hyp) >> sh) == (u32)node;
        u32 a_s = a, b_s = b, nl = lo, nh = hi;   // counters as byte i == hyp would leave them
        if (on_h) {
            const u32 ones = ((hyp >> (sh - 1)) & 1u) ? 0xFFFFu : 0u;
            a_s = cm_adapt_bf(a, ones, 2);
            b_s = cm_adapt_bf(b, ones, 4);
            nl = cm_adapt_bf(lo, ones, 6);
            nh = cm_adapt_bf(hi, ones, 6);
        }
        const u32 run_s = run + 1u;   // run rule (src/libbz3.c:367-370) applied to (prev1, prev1)
        const int flag_s = run_s > 2;
        const u32 p_s = ((a_s + b_s) * 7 + b_s + b_s) >> 4;
        u16* const cell_s = rows + flag_s * 17 + (p_s >> 12);
        u32 lo_s = cell_s[0], hi_s = cell_s[1];
        if (on_h) {   // the pending update of byte i is not in shared memory yet
            const bool same = cell_s == cell, up = cell_s == cell + 1, dn = cell_s + 1 == cell;
            lo_s = same ? nl : (up ? nh : lo_s);
            hi_s = same ? nh : (dn ? nl : hi_s);
        }
        {
            const int sse = (int)lo_s + ((((int)hi_s - (int)lo_s) * (int)(p_s & 4095)) >> 12);
            cm_ptab_put<LAYOUT>(ptab, ((i + 1) & 1) * 256 + node, (u32)(sse * 3 + (int)p_s) << 14);
        }
        if (PROTO == 0) {
            __syncthreads();   // byte ready
        } else {
            if (!__syncthreads_or(0))
                if (!__syncthreads_or(0)) __syncthreads();
        }
        const u32 byte = vbyte[i & 1];
        have = byte == hyp;   // uniform across the CTA
        if (have) {
            if (on_h) {   // (C) learn byte i
                *q0 = (u16)a_s;
                *q1 = (u16)b_s;
                cell[0] = (u16)nl;
                cell[1] = (u16)nh;
            }
            a = a_s;
            b = b_s;
            d = b_s;
            lo = lo_s;
            hi = hi_s;
            cell = cell_s;
            run = run_s;
            prev2 = prev1;   // == byte
        } else {
            u32 na = a, nb = b;
            if (node != 0 && ((256u | byte) >> sh) == (u32)node) {   // (C) learn byte i
                const u32 ones = ((byte >> (sh - 1)) & 1u) ? 0xFFFFu : 0u;
                na = cm_adapt_bf(a, ones, 2);
                nb = cm_adapt_bf(b, ones, 4);
                *q0 = (u16)na;
                *q1 = (u16)nb;
                cell[0] = (u16)cm_adapt_bf(lo, ones, 6);
                cell[1] = (u16)cm_adapt_bf(hi, ones, 6);
            }
            a = na;
            d = nb;                       // this byte's order-1 counter is the next byte's prev2 counter
            prev2 = prev1;
            prev1 = (int)byte;
            q1 = c1col + prev1 * 256;
            b = *q1;                      // after the store above in program order
        }
    }
}

// Third edition of the model thread (variant 9), written against the measured cost of a TAKEN branch (~20
// cycles, as much as two dependent multiplies).  In the slim loop above a thread of an off-path warp runs into
// five of them per byte on a speculation hit (skip the predict block, skip the outcome block, skip the extra
// barriers, skip the learn stores, loop back-edge): ~250 cycles per byte, more than the walkers need.  Here
//   * the loop is unrolled by the parity of the byte index (ptab / byte-slot offsets become immediates, the
//     loop-carried register shuffles disappear, half a back-edge per byte),
//   * the real prediction after a miss sits at the END of the step that missed (no "if (!have)" at the top),
//   * the barrier protocol and the learn stores are predicated instead of branched over.
#if defined(BZ_EMU)
BZ_D void cm_bar_byte_ready_or0() {
    if (!__syncthreads_or(0))
        if (!__syncthreads_or(0)) __syncthreads();
}
BZ_D void cm_learn_stores(bool on, u16* q0, u32 a, u16* q1, u32 b, u16* cell, u32 lo, u32 hi) {
    if (on) {
        *q0 = (u16)a;
        *q1 = (u16)b;
        cell[0] = (u16)lo;
        cell[1] = (u16)hi;
    }
}
#else
// "byte ready" of the walker kernel seen from a model thread: barrier-OR with a false vote; if no walker
// published, take part in the exact-test barrier-OR; if still nobody, in the barrier after the serial redo.
BZ_D void cm_bar_byte_ready_or0() {
    asm volatile(
        "{\n\t"
        ".reg .pred pf, p1, p2;\n\t"
        "setp.ne.u32 pf, 0, 0;\n\t"
        "bar.red.or.pred p1, 0, pf;\n\t"
        "mov.pred p2, p1;\n\t"
        "@!p1 bar.red.or.pred p2, 0, pf;\n\t"
        "@!p2 bar.sync 0;\n\t"
        "}" ::: "memory");
}
BZ_D void cm_learn_stores(bool on, u16* q0, u32 a, u16* q1, u32 b, u16* cell, u32 lo, u32 hi) {
    const u32 s0 = (u32)__cvta_generic_to_shared(q0), s1 = (u32)__cvta_generic_to_shared(q1);
    const u32 s2 = (u32)__cvta_generic_to_shared(cell);
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.u32 p, %0, 0;\n\t"
        "@p st.shared.u16 [%1], %2;\n\t"
        "@p st.shared.u16 [%3], %4;\n\t"
        "@p st.shared.u16 [%5], %6;\n\t"
        "@p st.shared.u16 [%5+2], %7;\n\t"
        "}" ::"r"((u32)on), "r"(s0), "h"((u16)a), "r"(s1), "h"((u16)b), "r"(s2), "h"((u16)lo), "h"((u16)hi)
        : "memory");
}
#endif

struct CmModelState {
    int prev1, prev2;
    u32 run, a, b, d, lo, hi;
    u16* q1;
    u16* cell;
};

// real prediction of the next byte from the registers (after a miss, and for byte 0) into half H of ptab
template <int LAYOUT, int H>
BZ_D void cm_model_predict(CmModelState& M, u32* ptab, u16* rows, const int node) {
    M.run = (M.prev1 == M.prev2) ? M.run + 1 : 0;
    const int flag = M.run > 2;
    const u32 p = ((M.a + M.b) * 7 + M.d + M.d) >> 4;
    M.cell = rows + flag * 17 + (p >> 12);
    M.lo = M.cell[0];
    M.hi = M.cell[1];
    const int sse = (int)M.lo + ((((int)M.hi - (int)M.lo) * (int)(p & 4095)) >> 12);
    cm_ptab_put<LAYOUT>(ptab, H * 256 + node, (u32)(sse * 3 + (int)p) << 14);
}

template <int LAYOUT, int HALF>
BZ_D void cm_model_step(CmModelState& M, u16* cm_smem, u32* ptab, volatile u32* vbyte, const bool last, const int node,
                        const int sh, u16* q0, u16* c1col, u16* rows) {
    // speculation: this byte == prev1
    const u32 hyp = (u32)M.prev1;
    const bool on_h = node != 0 && ((256u | hyp) >> sh) == (u32)node;
    u32 a_s = M.a, b_s = M.b, nl = M.lo, nh = M.hi;   // counters as (this byte == hyp) would leave them
    if (on_h) {
        const u32 ones = ((hyp >> (sh - 1)) & 1u) ? 0xFFFFu : 0u;
        a_s = cm_adapt_bf(M.a, ones, 2);
        b_s = cm_adapt_bf(M.b, ones, 4);
        nl = cm_adapt_bf(M.lo, ones, 6);
        nh = cm_adapt_bf(M.hi, ones, 6);
    }
    const u32 run_s = M.run + 1u;   // run rule (src/libbz3.c:367-370) applied to (prev1, prev1)
    const int flag_s = run_s > 2;
    const u32 p_s = ((a_s + b_s) * 7 + b_s + b_s) >> 4;
    u16* const cell_s = rows + flag_s * 17 + (p_s >> 12);
    u32 lo_s = cell_s[0], hi_s = cell_s[1];
    {   // the pending update of this byte is not in shared memory yet (predicated, no branch)
        const bool same = on_h && cell_s == M.cell, up = on_h && cell_s == M.cell + 1, dn = on_h && cell_s + 1 == M.cell;
        lo_s = same ? nl : (up ? nh : lo_s);
        hi_s = same ? nh : (dn ? nl : hi_s);
    }
    {
        const int sse = (int)lo_s + ((((int)hi_s - (int)lo_s) * (int)(p_s & 4095)) >> 12);
        cm_ptab_put<LAYOUT>(ptab, (HALF ^ 1) * 256 + node, (u32)(sse * 3 + (int)p_s) << 14);
    }
    cm_bar_byte_ready_or0();
    const u32 byte = vbyte[HALF];
    if (__builtin_expect(byte != hyp, 0)) {   // uniform across the CTA
        // miss: learn the byte that really came, then predict the next one for real
        u32 na = M.a, nb = M.b;
        if (node != 0 && ((256u | byte) >> sh) == (u32)node) {
            const u32 ones = ((byte >> (sh - 1)) & 1u) ? 0xFFFFu : 0u;
            na = cm_adapt_bf(M.a, ones, 2);
            nb = cm_adapt_bf(M.b, ones, 4);
            *q0 = (u16)na;
            *M.q1 = (u16)nb;
            M.cell[0] = (u16)cm_adapt_bf(M.lo, ones, 6);
            M.cell[1] = (u16)cm_adapt_bf(M.hi, ones, 6);
        }
        M.a = na;
        M.d = nb;                       // this byte's order-1 counter is the next byte's prev2 counter
        M.prev2 = M.prev1;
        M.prev1 = (int)byte;
        M.q1 = c1col + M.prev1 * 256;
        M.b = *M.q1;                    // after the store above in program order
        if (!last) {
            cm_model_predict<LAYOUT, HALF ^ 1>(M, ptab, rows, node);
            __syncthreads();   // ptab ready
        }
        return;
    }
    cm_learn_stores(on_h, q0, a_s, M.q1, b_s, M.cell, nl, nh);
    M.a = a_s;
    M.b = b_s;
    M.d = b_s;
    M.lo = lo_s;
    M.hi = hi_s;
    M.cell = cell_s;
    M.run = run_s;
    M.prev2 = M.prev1;   // == byte
}

template <int LAYOUT>
BZ_D void cm_dec_model_thread_slim2(u16* cm_smem, u32* ptab, volatile u32* vbyte, s32 n, const int node) {
    const int sh = node ? 8 - (31 - __clz(node)) : 8;                 // (256|byte) >> sh == node <=> on the path
    u16* const q0 = cm_smem + node;
    u16* const c1col = cm_smem + kCmC0 + node;                        // + prev * 256
    u16* const rows = cm_smem + kCmC0 + kCmC1 + (2 * node) * 17;      // + flag * 17 + cell
    CmModelState M;
    M.prev1 = 0;
    M.prev2 = 0;
    M.run = 0;
    M.q1 = c1col;
    M.a = *q0;
    M.b = *M.q1;
    M.d = M.b;
    M.lo = 0;
    M.hi = 0;
    M.cell = rows;
    if (n <= 0) return;
    cm_model_predict<LAYOUT, 0>(M, ptab, rows, node);
    __syncthreads();   // ptab of byte 0 ready
    for (s32 i = 0; i < n; i += 2) {
        cm_model_step<LAYOUT, 0>(M, cm_smem, ptab, vbyte, i + 1 >= n, node, sh, q0, c1col, rows);
        if (i + 1 < n) cm_model_step<LAYOUT, 1>(M, cm_smem, ptab, vbyte, i + 2 >= n, node, sh, q0, c1col, rows);
    }
}

// ---- tree-parallel decoder ---------------------------------------------------------------------
// Decoding is one dependent chain: the next context depends on the bit just decoded.  What does not
// depend on the bits of the current byte is the probability of every one of the 255 tree nodes (no node
// is visited twice within a byte, and prev1/prev2/runflag are fixed at the byte boundary).  Roles:
//   warps 1..8  256 model threads, thread owns ONE node and is the only one that ever touches its
//               counters: (C) learn the previous byte if the node was on its path, (A) predict -> ptab
//   warp 0      the chain: walks the 8 levels using ptab, two-tier like the encoder (branch-free fast
//               byte, exact redo when a renormalisation was needed), publishes the byte.
// Two __syncthreads per byte (ptab ready / byte ready).  The compressed bytes are staged through a
// 2 KiB shared window owned by warp 0, so the renormalisation never waits on global memory.
constexpr int kCmDecThreads = 288;
constexpr size_t kCmDecSmemBytes = (size_t)kCmTableU16 * 2 + 2 * 256 * 4 + 2048 + 64;

__global__ void __launch_bounds__(kCmDecThreads) cm_decode_tree_kernel(const u8* __restrict__ in, s32 insize,
                                                                      u8* __restrict__ out, s32 n) {
    BZ_DYN_SMEM(u16, cm_smem);
    u32* ptab = reinterpret_cast<u32*>(cm_smem + kCmTableU16);  // [2][256]  P << 14 per node; byte i uses half i&1
    u8* scode = reinterpret_cast<u8*>(ptab + 512);              // [2048] window of the compressed stream
    // decoded byte of step i lives in slot i&1: after a speculation hit the chain warp goes straight on to
    // byte i+1 and must not overwrite what the model threads are about to read
    volatile u32* vbyte = reinterpret_cast<volatile u32*>(scode + 2048);
    cm_tables_init_smem(cm_smem);
    const int tid = threadIdx.x;
    if (tid < 32)
        for (int k = tid; k < 2048; k += 32) scode[k] = (k < insize) ? in[k] : 0;
    __syncthreads();
    if (tid >= 32) {
        cm_dec_model_thread<0>(cm_smem, ptab, vbyte, n, tid - 32);
        return;
    }
    // ---------------------------------------------------------------------- chain warp (all lanes identical)
    s32 wlo = 0;  // the window holds stream bytes [wlo, wlo + 2048)
    s32 ip = 0;
    u32 low = 0, range = 0xFFFFFFFFu, code = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const u32 add = (ip < insize) ? (u32)scode[ip & 2047] : 0xFFFFFFFFu;  // read_in() past the end adds -1
        ip += (ip < insize);
        code = (code << 8) + add;
    }
    bool have = false;
    u32 prevb = 0;
#ifdef BZ_CM_PROFILE
    unsigned long long _acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, _t0 = clock64(), _t1;
#endif
    for (s32 i = 0; i < n; i++) {
        if (!have) __syncthreads();   // ptab ready (skipped when the speculation of the model threads hit)
        BZ_PROF_AFTER_BAR(0, ptab + (i & 1) * 256 + 1);
        const u32* pt = ptab + (i & 1) * 256;
        const uint4 g0 = *reinterpret_cast<const uint4*>(pt);
        uint4 gk = *reinterpret_cast<const uint4*>(pt + 4);
        u32 node = 1;
        {
            // fast tier: 8 branch-free steps on copies of the state, assuming no renormalisation is needed
            u32 flow = low, frange = range;
            u32 pcur = g0.y, kid0 = g0.z, kid1 = g0.w;
            u32 x = mulhi_pinned(frange, pcur);
            u32 tmin = 0xFFFFFFFFu;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                u32 bit;
#if defined(BZ_EMU)
                {
                    const u32 mid = flow + x;
                    bit = code <= mid;
                    frange = bit ? x : frange - x - 1u;
                    pcur = bit ? kid1 : kid0;
                    x = __umulhi(frange, pcur);
                    if (!bit) flow = mid + 1u;
                    const u32 t = flow ^ (flow + frange);
                    tmin = t < tmin ? t : tmin;
                }
#else
                asm volatile(
                    "{\n\t"
                    ".reg .pred pb;\n\t"
                    ".reg .u32 mid, nx, r0, hi, t;\n\t"
                    "add.u32 mid, %0, %2;\n\t"
                    "not.b32 nx, %2;\n\t"
                    "setp.le.u32 pb, %6, mid;\n\t"        // bit = code <= low + x
                    "add.u32 r0, %1, nx;\n\t"             // range - x - 1
                    "selp.u32 %1, %2, r0, pb;\n\t"        // bit ? x : range - x - 1
                    "selp.u32 %3, %8, %7, pb;\n\t"        // P of the chosen child
                    "mul.hi.u32 %2, %1, %3;\n\t"          // product for the next step
                    "@!pb add.u32 %0, mid, 1;\n\t"        // bit 0: low = mid + 1
                    "selp.u32 %4, 1, 0, pb;\n\t"
                    "add.u32 hi, %0, %1;\n\t"
                    "xor.b32 t, %0, hi;\n\t"
                    "min.u32 %5, %5, t;\n\t"
                    "}"
                    : "+r"(flow), "+r"(frange), "+r"(x), "+r"(pcur), "=r"(bit), "+r"(tmin)
                    : "r"(code), "r"(kid0), "r"(kid1));
#endif
                node = node * 2 + bit;
                kid0 = bit ? gk.z : gk.x;
                kid1 = bit ? gk.w : gk.y;
                if (k < 5) gk = *reinterpret_cast<const uint4*>(pt + 4 * node);
            }
            BZ_PROF(1);
            if (tmin >= (1u << 24)) {
                low = flow;
                range = frange;
            } else {
#ifdef BZ_CM_PROFILE
                _acc[4]++;
#endif
                // exact tier: same walk, renormalising after every step like the reference
                node = 1;
                gk = *reinterpret_cast<const uint4*>(pt + 4);
                pcur = g0.y;
                kid0 = g0.z;
                kid1 = g0.w;
                x = mulhi_pinned(range, pcur);
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    u32 bit, slow;
#if defined(BZ_EMU)
                    {
                        const u32 mid = low + x;
                        bit = code <= mid;
                        range = bit ? x : range - x - 1u;
                        pcur = bit ? kid1 : kid0;
                        x = __umulhi(range, pcur);
                        if (!bit) low = mid + 1u;
                        slow = range < 0x1000000u;
                    }
#else
                    asm volatile(
                        "{\n\t"
                        ".reg .pred pb, ps;\n\t"
                        ".reg .u32 mid, nx, r0;\n\t"
                        "add.u32 mid, %0, %2;\n\t"
                        "not.b32 nx, %2;\n\t"
                        "setp.le.u32 pb, %6, mid;\n\t"
                        "add.u32 r0, %1, nx;\n\t"
                        "selp.u32 %1, %2, r0, pb;\n\t"
                        "selp.u32 %3, %8, %7, pb;\n\t"
                        "mul.hi.u32 %2, %1, %3;\n\t"
                        "@!pb add.u32 %0, mid, 1;\n\t"
                        "selp.u32 %4, 1, 0, pb;\n\t"
                        "setp.lt.u32 ps, %1, 0x1000000;\n\t"
                        "selp.u32 %5, 1, 0, ps;\n\t"
                        "}"
                        : "+r"(low), "+r"(range), "+r"(x), "+r"(pcur), "=r"(bit), "=r"(slow)
                        : "r"(code), "r"(kid0), "r"(kid1));
#endif
                    node = node * 2 + bit;
                    kid0 = bit ? gk.z : gk.x;
                    kid1 = bit ? gk.w : gk.y;
                    if (k < 5) gk = *reinterpret_cast<const uint4*>(pt + 4 * node);
                    if (slow) {
                        u32 high = low + range;
                        while ((low ^ high) < (1u << 24)) {
                            low <<= 8;
                            high = (high << 8) | 0xFFu;
                            const u32 add = (ip < insize) ? (u32)scode[ip & 2047] : 0xFFFFFFFFu;
                            ip += (ip < insize);
                            code = (code << 8) + add;
                        }
                        range = high - low;
                        x = mulhi_pinned(range, pcur);
                    }
                }
            }
        }
        BZ_PROF(2);
        const u32 byte = node & 255u;
        // every lane holds the same byte: unconditional (convergent) stores of one value to one address
        vbyte[i & 1] = byte;
        out[i] = (u8)byte;
        if (ip - wlo >= 1024) {  // uniform in the warp; the window belongs to this warp alone
            __syncwarp();
            for (int k = tid; k < 1024; k += 32) {
                const s32 src = wlo + 2048 + k;
                scode[src & 2047] = (src < insize) ? in[src] : 0;
            }
            wlo += 1024;
            __syncwarp();
        }
        __syncthreads();   // byte ready
        BZ_PROF_AFTER_BAR(3, ptab);
        have = byte == prevb;
        prevb = byte;
    }
#ifdef BZ_CM_PROFILE
    if (tid == 0)
        for (int k = 0; k < 5; k++) g_cm_prof[8 + k] = _acc[k];   // wait ptab, fast tier, exact tier, publish+wait byte, #redo
#endif
}

// ---- lane-parallel walk (variant 4) -------------------------------------------------------------------
// The chain warp of the tree decoder executes every instruction for 32 lanes that all hold the same
// values.  Here the lanes hold the 32 possible five-bit prefixes of the byte instead.  A lane's branches
// are fixed, so its walk needs no compare/select/lookup chain -- one multiply per level:
//     bit 1:  new range = x = (range * P) >> 18         = hi32(range * M),    M = P << 14
//     bit 0:  new range = range - x - 1                 = hi32(range * -M)    unless lo32(range * M) == 0
// (range * (2^32 - M) = range * 2^32 - range * M, and floor of that over 2^32 is range - ceil(range*M / 2^32)).
// The model threads store {M, -M} per node and a lane loads the one of its branch; a lane that meets a zero
// low half anywhere disqualifies itself (about once per 2^18 decisions; that also covers range == 0).
// Without a renormalisation the 32 sub-intervals partition the current interval, so exactly one lane ends
// with  code - low' <= range'  -- found with a ballot, its (low', range') broadcast with two shuffles.
// low' needs no per-level work either: low only moves at 0-bits, by range_k - range_{k+1}, which telescopes
// into a lane-constant +1/0/-1 combination of the ranges.  A second round does the last three levels with
// 8 suffixes.  A lane whose range dropped below 2^24 (necessary for a shift) also disqualifies itself; if
// that was the true path nobody wins and the round is redone by the reference loop (cm_dec_exact_levels).

// reference loop for `nlev` tree levels from `node` on (src/libbz3.c:452-476); uniform across the caller's lanes.
// Serial, so written for the dependent-issue latencies measured on B200 (multiply 10, shared load 23, add/select
// ~3.4): one mul.hi per level, both outcomes of the level formed next to the compare, and the multipliers of both
// possible next nodes' children loaded one level ahead (the {M, -M} pairs of the two children of a node share one
// 16-byte word), so no load sits on the recurrence  mul.hi -> add -> compare -> select.
template <int NLEV>
BZ_D u32 cm_dec_exact_levels(const u32* __restrict__ pt, u32 node, u32& low, u32& range, u32& code,
                             s32& ip, const s32 insize, const u8* __restrict__ scode) {
    const uint4* pt4 = reinterpret_cast<const uint4*>(pt);   // pt4[n] = {M, -M} of nodes 2n and 2n+1
    u32 m = pt[2 * node];
    uint4 kids = pt4[node];
    u32 r = range, lo = low;   // absolute low, not code - low: a hostile stream may put the code below low
#pragma unroll
    for (int k = 0; k < NLEV; k++) {   // unrolled: a taken branch costs as much as the whole recurrence of a level
        uint4 g0 = kids, g1 = kids;
        if (node < 64) {   // children of both possible next nodes
            g0 = pt4[2 * node];
            g1 = pt4[2 * node + 1];
        }
        const u32 x = __umulhi(r, m);
        const u32 mid = lo + x;
        const bool bit = code <= mid;
        r = bit ? x : r + ~x;        // x  |  range - x - 1
        lo = bit ? lo : mid + 1u;
        node = node * 2 + (bit ? 1u : 0u);
        m = bit ? kids.z : kids.x;
        kids = bit ? g1 : g0;
        if (r < (1u << 24)) {   // necessary for the top bytes of low and high to agree
            u32 high = lo + r;
            while ((lo ^ high) < (1u << 24)) {
                lo <<= 8;
                high = (high << 8) | 0xFFu;
                const u32 add = (ip < insize) ? (u32)scode[ip & 2047] : 0xFFFFFFFFu;   // read_in() past the end adds -1
                ip += (ip < insize);
                code = (code << 8) + add;
            }
            r = high - lo;
        }
    }
    low = lo;
    range = r;
    return node;
}

// Shared-memory access by 32-bit shared-window address with an immediate displacement (no generic-address
// arithmetic in the loop).  The emulator keeps ordinary pointers.
#if defined(BZ_EMU)
typedef const u8* SmemAddr;
BZ_D SmemAddr smem_addr_of(const volatile void* p) { return reinterpret_cast<const u8*>(const_cast<const void*>(p)); }
template <int DISP> BZ_D u32 lds_u32(SmemAddr a) { return *reinterpret_cast<const volatile u32*>(a + DISP); }
template <int DISP> BZ_D void sts_u32(SmemAddr a, u32 v) { *reinterpret_cast<volatile u32*>(const_cast<u8*>(a) + DISP) = v; }
#else
typedef u32 SmemAddr;
BZ_D SmemAddr smem_addr_of(const volatile void* p) { return (u32)__cvta_generic_to_shared(const_cast<const void*>(p)); }
template <int DISP> BZ_D u32 lds_u32(SmemAddr a) {
    u32 v;
    asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(DISP) : "memory");
    return v;
}
template <int DISP> BZ_D void sts_u32(SmemAddr a, u32 v) {
    asm volatile("st.shared.u32 [%0+%1], %2;" :: "r"(a), "n"(DISP), "r"(v) : "memory");
}
#endif

// ptxas likes to rematerialise thread-constant values from the thread id inside the hot loop (a dozen extra
// instructions per byte).  Values that went through shared memory once are opaque to it and stay in registers.
template <int N>
BZ_D void launder_u32(u32 (&x)[N], volatile u32* scratch) {
#pragma unroll
    for (int k = 0; k < N; k++) scratch[k] = x[k];
#pragma unroll
    for (int k = 0; k < N; k++) x[k] = scratch[k];
}

#if defined(BZ_EMU_STATS)
static unsigned long long g_emu_stats[8];
#endif

// One byte of the lane-parallel walk.  HALF (the parity of the byte index) is a template parameter so that
// every shared-memory displacement in the loop is an immediate.
struct CmLaneConsts {
    SmemAddr a1[5];   // round 1: shared address of this lane's {M | -M} word at levels 0..4 (half 0)
    SmemAddr a2[3];   // round 2: ptab base + the lane-constant part of the word offset at levels 5..7
    u32 cs1[6];       // +1 / 0 / -1 (mod 2^32): coefficient of range_k in  low_5 - low_0
    u32 cs2[4];
    u32 z1[5], z2[3]; // all-ones where this lane takes the 0-branch at that level (slow path only)
};
struct CmChainState {
    u32 low, range, code, prevb;
    s32 ip, wlo;
    bool have;
#ifdef BZ_CM_PROFILE
    unsigned long long _acc[8], _t0, _t1;
#endif
};
#ifdef BZ_CM_PROFILE
#define BZ_SPROF(S, slot) do { (S)._t1 = clock64(); (S)._acc[slot] += (S)._t1 - (S)._t0; (S)._t0 = (S)._t1; } while (0)
#define BZ_SPROF_AFTER_BAR(S, slot, ptr) do { unsigned _v = *(ptr); asm volatile("mov.u64 %0, %%clock64; // %1" : "=l"((S)._t1) : "r"(_v)); (S)._acc[slot] += (S)._t1 - (S)._t0; (S)._t0 = (S)._t1; } while (0)
#define BZ_SCOUNT(S, slot) ((S)._acc[slot]++)
#else
#define BZ_SPROF(S, slot)
#define BZ_SPROF_AFTER_BAR(S, slot, ptr)
#define BZ_SCOUNT(S, slot)
#endif

template <int HALF>
BZ_D void cm_dec_lanes_step(const CmLaneConsts& K, CmChainState& S, const s32 i, u32* ptab, const u8* __restrict__ scode,
                            SmemAddr vbyte_a, const u8* __restrict__ in, const s32 insize, u8* __restrict__ out,
                            const int tid) {
    if (!S.have) __syncthreads();   // ptab ready (skipped when the speculation of the model threads hit)
    BZ_SPROF_AFTER_BAR(S, 0, ptab + HALF * 512 + 2);
    constexpr int HB = HALF * 2048;   // byte offset of this byte's half of ptab
    const u32* pt = ptab + HALF * 512;
    u32 node;
    {   // round 1: levels 0..4, 32 prefixes
        const u32 m0 = lds_u32<HB>(K.a1[0]), m1 = lds_u32<HB>(K.a1[1]), m2 = lds_u32<HB>(K.a1[2]);
        const u32 m3 = lds_u32<HB>(K.a1[3]), m4 = lds_u32<HB>(K.a1[4]);
        const u64 w1 = cm_mul_wide(S.range, m0);
        const u32 r1 = (u32)(w1 >> 32);
        const u64 w2 = cm_mul_wide(r1, m1);
        const u32 r2 = (u32)(w2 >> 32);
        const u64 w3 = cm_mul_wide(r2, m2);
        const u32 r3 = (u32)(w3 >> 32);
        const u64 w4 = cm_mul_wide(r3, m3);
        const u32 r4 = (u32)(w4 >> 32);
        const u64 w5 = cm_mul_wide(r4, m4);
        const u32 r5 = (u32)(w5 >> 32);
        const u32 acc = K.cs1[0] * S.range + K.cs1[1] * r1 + K.cs1[2] * r2 + K.cs1[3] * r3 + K.cs1[4] * r4 + K.cs1[5] * r5;
        const u32 zmin = min(min(min((u32)w1, (u32)w2), min((u32)w3, (u32)w4)), (u32)w5);   // 0 <=> some low half was 0
        const u32 rchk = zmin ? r5 : 0u;   // ranges only shrink along a path: the last one is the smallest
        const u32 d5 = S.code - S.low - acc;
        u32 win = __ballot_sync(kFullMask, d5 <= r5 && rchk >= (1u << 24));
        if (!win) {
            // range < 2^24 is only NECESSARY for a shift (low and high may straddle a top-byte boundary for a
            // while).  Before giving up, every lane applies the reference's test to its own path.
            u32 a = S.low, tmin;
            a += K.z1[0] & (S.range - r1);
            tmin = a ^ (a + r1);
            a += K.z1[1] & (r1 - r2);
            tmin = min(tmin, a ^ (a + r2));
            a += K.z1[2] & (r2 - r3);
            tmin = min(tmin, a ^ (a + r3));
            a += K.z1[3] & (r3 - r4);
            tmin = min(tmin, a ^ (a + r4));
            a += K.z1[4] & (r4 - r5);
            tmin = min(tmin, a ^ (a + r5));
            win = __ballot_sync(kFullMask, d5 <= r5 && zmin != 0u && tmin >= (1u << 24));
            BZ_SCOUNT(S, 6);
        }
        if (win) {
            const int w = 31 - __clz((int)win);   // exactly one lane wins
            S.range = __shfl_sync(kFullMask, r5, w);
            S.low = S.code - __shfl_sync(kFullMask, d5, w);
            node = 32u | (u32)w;
            BZ_SPROF(S, 1);
        } else {
            BZ_SPROF(S, 1);
#if defined(BZ_EMU_STATS)
            if (tid == 0) g_emu_stats[0]++;
#endif
            node = cm_dec_exact_levels<5>(pt, 1u, S.low, S.range, S.code, S.ip, insize, scode);
            BZ_SPROF(S, 2);
        }
    }
    {   // round 2: levels 5..7, 8 suffixes (four copies each)
        const u32 m5 = lds_u32<HB>(K.a2[0] + node * 8), m6 = lds_u32<HB>(K.a2[1] + node * 16), m7 = lds_u32<HB>(K.a2[2] + node * 32);
        const u64 w6 = cm_mul_wide(S.range, m5);
        const u32 r6 = (u32)(w6 >> 32);
        const u64 w7 = cm_mul_wide(r6, m6);
        const u32 r7 = (u32)(w7 >> 32);
        const u64 w8 = cm_mul_wide(r7, m7);
        const u32 r8 = (u32)(w8 >> 32);
        const u32 acc = K.cs2[0] * S.range + K.cs2[1] * r6 + K.cs2[2] * r7 + K.cs2[3] * r8;
        const u32 zmin = min(min((u32)w6, (u32)w7), (u32)w8);
        const u32 rchk = zmin ? r8 : 0u;
        const u32 d8 = S.code - S.low - acc;
        u32 win = __ballot_sync(kFullMask, d8 <= r8 && rchk >= (1u << 24));
        if (!win) {   // exact per-path shift test, as in round 1
            u32 a = S.low, tmin;
            a += K.z2[0] & (S.range - r6);
            tmin = a ^ (a + r6);
            a += K.z2[1] & (r6 - r7);
            tmin = min(tmin, a ^ (a + r7));
            a += K.z2[2] & (r7 - r8);
            tmin = min(tmin, a ^ (a + r8));
            win = __ballot_sync(kFullMask, d8 <= r8 && zmin != 0u && tmin >= (1u << 24));
            BZ_SCOUNT(S, 7);
        }
        if (win) {
            const int w = (31 - __clz((int)win)) & 7;   // lanes j, j+8, j+16, j+24 hold suffix j
            S.range = __shfl_sync(kFullMask, r8, w);
            S.low = S.code - __shfl_sync(kFullMask, d8, w);
            node = node * 8 + (u32)w;
            BZ_SPROF(S, 3);
        } else {
            BZ_SPROF(S, 3);
#if defined(BZ_EMU_STATS)
            if (tid == 0) g_emu_stats[1]++;
#endif
            node = cm_dec_exact_levels<3>(pt, node, S.low, S.range, S.code, S.ip, insize, scode);
            BZ_SPROF(S, 4);
        }
    }
    const u32 byte = node & 255u;
    // every lane holds the same byte: unconditional (convergent) stores of one value to one address
    sts_u32<HALF * 4>(vbyte_a, byte);
    out[i] = (u8)byte;
    if (S.ip - S.wlo >= 1024) {  // uniform in the warp; the window belongs to this warp alone
        __syncwarp();
        u8* sc = const_cast<u8*>(scode);
        for (int k = tid; k < 1024; k += 32) {
            const s32 src = S.wlo + 2048 + k;
            sc[src & 2047] = (src < insize) ? in[src] : 0;
        }
        S.wlo += 1024;
        __syncwarp();
    }
    __syncthreads();   // byte ready
    BZ_SPROF_AFTER_BAR(S, 5, ptab);
#if defined(BZ_EMU_STATS)
    if (tid == 0) { g_emu_stats[2]++; g_emu_stats[3] += (byte == S.prevb); }
#endif
    S.have = byte == S.prevb;
    S.prevb = byte;
}

BZ_D void cm_dec_lanes_chain(u32* ptab, u8* scode, volatile u32* vbyte, const u8* __restrict__ in, const s32 insize,
                             u8* __restrict__ out, const s32 n) {
    const int tid = threadIdx.x;
    const u32 L = (u32)tid;
    CmLaneConsts K;
    {
        // Round 1: the branch at level k is bit 4-k of the lane id.  Round 2: the branch at level 5+k is bit
        // 2-k of j = lane & 7.  Word offset of a node's {M, -M} pair is 2*node; +1 selects -M (a 0-branch).
        u32 zprev = 0;
#pragma unroll
        for (int k = 0; k < 5; k++) {
            const u32 b = (L >> (4 - k)) & 1u;
            const u32 nodek = (1u << k) | (L >> (5 - k));
            K.a1[k] = smem_addr_of(ptab) + 4 * (nodek * 2 + (b ? 0u : 1u));
            K.z1[k] = b ? 0u : 0xFFFFFFFFu;
            const u32 z = b ? 0u : 1u;
            K.cs1[k] = z - zprev;
            zprev = z;
        }
        K.cs1[5] = 0u - zprev;
        const u32 j = L & 7u;
        zprev = 0;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const u32 b = (j >> (2 - k)) & 1u;
            K.a2[k] = smem_addr_of(ptab) + 4 * ((j >> (3 - k)) * 2 + (b ? 0u : 1u));   // + (node << (3 + k)) at run time
            K.z2[k] = b ? 0u : 0xFFFFFFFFu;
            const u32 z = b ? 0u : 1u;
            K.cs2[k] = z - zprev;
            zprev = z;
        }
        K.cs2[3] = 0u - zprev;
    }
    {
        volatile u32* scr = vbyte + 16 + 32 * L;   // private scratch behind the byte slots
#if !defined(BZ_EMU)
        launder_u32(K.a1, scr);
        launder_u32(K.a2, scr + 5);
#endif
        launder_u32(K.cs1, scr + 8);
        launder_u32(K.cs2, scr + 14);
    }
    CmChainState S;
    S.wlo = 0;  // the window holds stream bytes [wlo, wlo + 2048)
    S.ip = 0;
    S.low = 0;
    S.range = 0xFFFFFFFFu;
    S.code = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const u32 add = (S.ip < insize) ? (u32)scode[S.ip & 2047] : 0xFFFFFFFFu;  // read_in() past the end adds -1
        S.ip += (S.ip < insize);
        S.code = (S.code << 8) + add;
    }
    S.have = false;
    S.prevb = 0;
    const SmemAddr vbyte_a = smem_addr_of(vbyte);
#ifdef BZ_CM_PROFILE
    for (int k = 0; k < 8; k++) S._acc[k] = 0;
    S._t0 = clock64();
#endif
    for (s32 i = 0; i < n; i += 2) {
        cm_dec_lanes_step<0>(K, S, i, ptab, scode, vbyte_a, in, insize, out, tid);
        if (i + 1 < n) cm_dec_lanes_step<1>(K, S, i + 1, ptab, scode, vbyte_a, in, insize, out, tid);
    }
#ifdef BZ_CM_PROFILE
    if (tid == 0)
        for (int k = 0; k < 8; k++) g_cm_prof[16 + k] = S._acc[k];   // wait ptab, r1 fast, r1 exact, r2 fast, r2 exact, publish+wait, #fb1, #fb2
#endif
}

constexpr size_t kCmDecLanesSmemBytes = (size_t)kCmTableU16 * 2 + 2 * 256 * 8 + 2048 + 64 + 32 * 32 * 4;

__global__ void __launch_bounds__(kCmDecThreads, 1) cm_decode_lanes_kernel(const u8* __restrict__ in, s32 insize,
                                                                       u8* __restrict__ out, s32 n) {
    BZ_DYN_SMEM(u16, cm_smem);
    u32* ptab = reinterpret_cast<u32*>(cm_smem + kCmTableU16);  // [2][256] {M, -M}; byte i uses half i&1
    u8* scode = reinterpret_cast<u8*>(ptab + 1024);             // [2048] window of the compressed stream
    volatile u32* vbyte = reinterpret_cast<volatile u32*>(scode + 2048);   // decoded byte of step i in slot i&1
    cm_tables_init_smem(cm_smem);
    const int tid = threadIdx.x;
    if (tid < 32)
        for (int k = tid; k < 2048; k += 32) scode[k] = (k < insize) ? in[k] : 0;
    __syncthreads();
    if (tid >= 32) {
        cm_dec_model_thread<1>(cm_smem, ptab, vbyte, n, tid - 32);
        return;
    }
    // ---------------------------------------------------------------------- chain warp
    cm_dec_lanes_chain(ptab, scode, vbyte, in, insize, out, n);
}

// ---- all-paths decoder ---------------------------------------------------------------------------
// Same model phase as above, but no chain warp: after the 255 node probabilities of a byte are known,
// thread v (0..255) walks the root-to-leaf path of byte value v with its own, compile-time-known-per-
// thread bits: 8 x (mul.hi, compare, update, predicated renormalisation shift).  A path is "alive"
// while every decision the coder would take on it (code <= split) equals the path's bit; exactly one
// path stays alive to the leaf -- that thread publishes the byte and the new coder state.  Nothing on
// the walk is selected by a decoded bit, so the serial select chain of the chain-warp kernel disappears
// and the walk costs ~8 x 35 cycles.  Warps whose 32 leaves are all dead leave early (a warp shares the
// top three bits).  Two situations fall back to the exact serial byte decoder on thread 0: a step that
// needs a second shift, and the last 8 bytes of the stream (the reference adds -1 past the end).
constexpr int kCmDecPathsThreads = 256;

__global__ void __launch_bounds__(kCmDecPathsThreads) cm_decode_paths_kernel(const u8* __restrict__ in, s32 insize,
                                                                            u8* __restrict__ out, s32 n) {
    BZ_DYN_SMEM(u16, cm_smem);
    u32* ptab = reinterpret_cast<u32*>(cm_smem + kCmTableU16);  // [256]  P << 14 per node
    u8* scode = reinterpret_cast<u8*>(ptab + 256);              // [2048] window of the compressed stream
    volatile u32* st = reinterpret_cast<volatile u32*>(scode + 2048);  // 2 slots of 8: [0]=byte [1]=low [2]=range [3]=code [4]=ip
    cm_tables_init_smem(cm_smem);
    const int tid = threadIdx.x;
    for (int k = tid; k < 2048; k += kCmDecPathsThreads) scode[k] = (k < insize) ? in[k] : 0;
    __syncthreads();
    if (tid == 0) {
        s32 ip = 0;
        u32 code = 0;
        for (int k = 0; k < 4; k++) {
            const u32 add = (ip < insize) ? (u32)scode[ip & 2047] : 0xFFFFFFFFu;  // read_in() past the end adds -1
            ip += (ip < insize);
            code = (code << 8) + add;
        }
        st[1] = 0u;
        st[2] = 0xFFFFFFFFu;
        st[3] = code;
        st[4] = (u32)ip;
    }
    // model role: owner of node `tid` (0 is a dummy)
    const int node = tid;
    const int sh = node ? 8 - (31 - __clz(node)) : 8;
    u16* const q0 = cm_smem + node;
    u16* const c1col = cm_smem + kCmC0 + node;
    u16* const rows = cm_smem + kCmC0 + kCmC1 + (2 * node) * 17;
    int prev1 = 0, prev2 = 0;
    u32 run = 0;
    u16* q1 = c1col;
    u32 a = *q0, b = *q1, d = *q1;
    // path role: leaf value v = tid; node visited at level k is (256 | v) >> (8 - k)
    const u32 v = (u32)tid;
    const u32* const scode32 = reinterpret_cast<const u32*>(scode);
    s32 wlo = 0;
    for (s32 i = 0; i < n; i++) {
        run = (prev1 == prev2) ? run + 1 : 0;
        const int flag = run > 2;
        // (A) predict byte i
        const u32 p = ((a + b) * 7 + d + d) >> 4;
        u16* const cell = rows + flag * 17 + (p >> 12);
        const u32 lo = cell[0], hi = cell[1];
        const int sse = (int)lo + ((((int)hi - (int)lo) * (int)(p & 4095)) >> 12);
        ptab[node] = (u32)(sse * 3 + (int)p) << 14;
        __syncthreads();   // S1: ptab and coder state ready
        // ---- walk (coder state is double buffered: read slot i&1, write slot (i+1)&1)
        volatile u32* const sr = st + 8 * (i & 1);
        volatile u32* const sw = st + 8 * ((i + 1) & 1);
        u32 low = sr[1], range = sr[2], code = sr[3];
        const s32 ip = (s32)sr[4];
        const bool tail = ip + 8 > insize;             // uniform
        u32 P0 = ptab[1], P1 = ptab[(256u | v) >> 7], P2 = ptab[(256u | v) >> 6], P3 = ptab[(256u | v) >> 5];
        u32 P4 = ptab[(256u | v) >> 4], P5 = ptab[(256u | v) >> 3], P6 = ptab[(256u | v) >> 2], P7 = ptab[(256u | v) >> 1];
        // next 8 stream bytes, big-endian in (chi, clo)
        u32 chi, clo;
        {
            const u32 w = ((u32)ip >> 2) & 511u, sft = ((u32)ip & 3u) * 8u;
            const u32 a0 = scode32[w], a1 = scode32[(w + 1) & 511u], a2 = scode32[(w + 2) & 511u];
            chi = __byte_perm(__funnelshift_r(a0, a1, sft), 0u, 0x0123);
            clo = __byte_perm(__funnelshift_r(a1, a2, sft), 0u, 0x0123);
        }
        bool ok = true, dbl = false;
        u32 nsh = 0;
#define BZ_PATH_STEP(K, PK)                                                            \
        {   /* branch-free: lanes of a warp take different bits and shift at different steps */ \
            const bool bk = ((v >> (7 - (K))) & 1u) != 0;                              \
            const u32 x = __umulhi(range, (PK));                                       \
            const u32 mid = low + x;                                                   \
            ok = ok && ((code <= mid) == bk);                                          \
            low = bk ? low : mid + 1u;                                                 \
            range = bk ? x : range - x - 1u;                                           \
            const bool s_ = ((low ^ (low + range)) < (1u << 24));                      \
            const u32 ncode = __funnelshift_l(chi, code, 8);                           \
            const u32 nchi = __funnelshift_l(clo, chi, 8);                             \
            low = s_ ? (low << 8) : low;                                               \
            range = s_ ? ((range << 8) | 0xFFu) : range;                               \
            code = s_ ? ncode : code;                                                  \
            chi = s_ ? nchi : chi;                                                     \
            clo = s_ ? (clo << 8) : clo;                                               \
            nsh += s_ ? 1u : 0u;                                                       \
            dbl = dbl || (ok && s_ && ((low ^ (low + range)) < (1u << 24)));           \
        }
        BZ_PATH_STEP(0, P0)
        BZ_PATH_STEP(1, P1)
        BZ_PATH_STEP(2, P2)
        if (__any_sync(kFullMask, ok)) {          // a warp shares the top three bits: 7 of 8 warps stop here
            BZ_PATH_STEP(3, P3)
            BZ_PATH_STEP(4, P4)
            BZ_PATH_STEP(5, P5)
            BZ_PATH_STEP(6, P6)
            BZ_PATH_STEP(7, P7)
        } else {
            ok = false;
        }
#undef BZ_PATH_STEP
        if (ok && !dbl && !tail) {
            sw[0] = v;
            sw[1] = low;
            sw[2] = range;
            sw[3] = code;
            sw[4] = (u32)ip + nsh;
        }
        const int fallback = __syncthreads_or((dbl || tail) ? 1 : 0);   // S2: byte and state ready (or nobody won)
        if (fallback) {
            if (tid == 0) {   // exact serial decoder for this byte (reference loop)
                u32 flow = sr[1], fhigh = sr[1] + sr[2], fcode = sr[3];
                s32 fip = (s32)sr[4];
                u32 nd = 1;
                for (int k = 0; k < 8; k++) {
                    const u32 mid = flow + __umulhi(fhigh - flow, ptab[nd]);
                    const bool bit = fcode <= mid;
                    if (bit) fhigh = mid; else flow = mid + 1u;
                    nd = nd * 2 + (bit ? 1u : 0u);
                    while ((flow ^ fhigh) < (1u << 24)) {
                        flow <<= 8;
                        fhigh = (fhigh << 8) | 0xFFu;
                        const u32 add = (fip < insize) ? (u32)scode[fip & 2047] : 0xFFFFFFFFu;
                        fip += (fip < insize);
                        fcode = (fcode << 8) + add;
                    }
                }
                sw[0] = nd & 255u;
                sw[1] = flow;
                sw[2] = fhigh - flow;
                sw[3] = fcode;
                sw[4] = (u32)fip;
            }
            __syncthreads();
        }
        const u32 byte = sw[0];
        if (tid == 0) out[i] = (u8)byte;
        // both outcomes of the update were not precomputed here (no idle phase): learn directly
        const bool on = node != 0 && ((256u | byte) >> sh) == (u32)node;
        const u32 ones = ((byte >> (sh - 1)) & 1u) ? 0xFFFFu : 0u;
        const u32 na = cm_adapt_bf(a, ones, 2), nb = cm_adapt_bf(b, ones, 4);
        if (on) {   // (C) learn byte i
            *q0 = (u16)na;
            *q1 = (u16)nb;
            cell[0] = (u16)cm_adapt_bf(lo, ones, 6);
            cell[1] = (u16)cm_adapt_bf(hi, ones, 6);
        }
        a = on ? na : a;
        d = on ? nb : b;
        prev2 = prev1;
        prev1 = (int)byte;
        q1 = c1col + prev1 * 256;
        b = *q1;
        // keep the stream window ahead of the read position (uniform)
        const s32 nip = (s32)sw[4];
        if (nip - wlo >= 1024) {
            __syncthreads();
            for (int k = tid; k < 1024; k += kCmDecPathsThreads) {
                const s32 src = wlo + 2048 + k;
                scode[src & 2047] = (src < insize) ? in[src] : 0;
            }
            wlo += 1024;
        }
    }
}

// ---- all-paths decoder, second edition (variant 5) ------------------------------------------------------
// 256 threads, no chain warp.  Thread t owns tree node t (model role, as above) AND walks the root-to-leaf
// path of byte value t with the one-multiply-per-level step of the lane-parallel walk: its eight branches
// are fixed, so the walk is 8 shared loads at thread-constant addresses, 8 dependent IMAD.WIDE, one
// lane-constant +1/0/-1 combination of the nine ranges for the new low, and one final test.  Without a
// renormalisation the 256 leaf intervals partition the current interval: exactly one thread passes
//   code - low' <= range'   and   range' >= 2^24   and   no zero low half on the way
// and publishes (byte, low', range').  When the true path needs a shift nobody passes; the barrier
// reduction reports that and thread 0 redoes the byte with the reference loop.  Per byte: predict -> barrier
// -> walk -> barrier(+or) -> learn.  Every thread carries the (uniform) coder state in registers.
constexpr int kCmDecP2Threads = 256;
constexpr size_t kCmDecP2SmemBytes = (size_t)kCmTableU16 * 2 + 256 * 8 + 2048 + 64 + 256 * 20 * 4;

__global__ void __launch_bounds__(kCmDecP2Threads, 1) cm_decode_paths2_kernel(const u8* __restrict__ in, s32 insize,
                                                                          u8* __restrict__ out, s32 n) {
    BZ_DYN_SMEM(u16, cm_smem);
    u32* ptab = reinterpret_cast<u32*>(cm_smem + kCmTableU16);  // [256] {M, -M} per node
    u8* scode = reinterpret_cast<u8*>(ptab + 512);              // [2048] window of the compressed stream
    volatile u32* pub = reinterpret_cast<volatile u32*>(scode + 2048);   // [0..3] byte, low, range, code  [4] ip
    cm_tables_init_smem(cm_smem);
    const int tid = threadIdx.x;
    for (int k = tid; k < 2048; k += kCmDecP2Threads) scode[k] = (k < insize) ? in[k] : 0;
    // model role: owner of node `tid` (0 is a dummy)
    const int node = tid;
    const int sh = node ? 8 - (31 - __clz(node)) : 8;
    u16* const q0 = cm_smem + node;
    u16* const c1col = cm_smem + kCmC0 + node;
    u16* const rows = cm_smem + kCmC0 + kCmC1 + (2 * node) * 17;
    // path role: leaf v = tid.  Level k visits node (1 << k) | (v >> (8 - k)) and takes branch bit 7-k of v.
    const u32 v = (u32)tid;
    SmemAddr pa[8];
    u32 cs[9];   // +1 / 0 / -1 (mod 2^32): coefficient of range_k in  low_8 - low_0
    u32 zk[8];   // all-ones where this path takes the 0-branch (slow path only)
    {
        u32 zprev = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const u32 b = (v >> (7 - k)) & 1u;
            const u32 nodek = (1u << k) | (v >> (8 - k));
            pa[k] = smem_addr_of(ptab) + 4 * (nodek * 2 + (b ? 0u : 1u));
            zk[k] = b ? 0u : 0xFFFFFFFFu;
            const u32 z = b ? 0u : 1u;
            cs[k] = z - zprev;
            zprev = z;
        }
        cs[8] = 0u - zprev;
    }
    {
        volatile u32* scr = pub + 16 + 20 * tid;   // private scratch behind the publication slots
#if !defined(BZ_EMU)
        launder_u32(pa, scr);
#endif
        launder_u32(cs, scr + 8);
    }
    const SmemAddr pub_a = smem_addr_of(pub);
    __syncthreads();
    int prev1 = 0, prev2 = 0;
    u32 run = 0;
    u16* q1 = c1col;
    u32 a = *q0, b = *q1, d = *q1;
    // coder state, identical in every thread
    s32 wlo = 0, ip = 0;
    u32 low = 0, range = 0xFFFFFFFFu, code = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const u32 add = (ip < insize) ? (u32)scode[ip & 2047] : 0xFFFFFFFFu;  // read_in() past the end adds -1
        ip += (ip < insize);
        code = (code << 8) + add;
    }
#ifdef BZ_CM_PROFILE
    unsigned long long _acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, _t0 = clock64(), _t1;
#endif
    for (s32 i = 0; i < n; i++) {
        run = (prev1 == prev2) ? run + 1 : 0;
        const int flag = run > 2;
        // (A) predict byte i
        const u32 p = ((a + b) * 7 + d + d) >> 4;
        u16* const cell = rows + flag * 17 + (p >> 12);
        const u32 lo = cell[0], hi = cell[1];
        const int sse = (int)lo + ((((int)hi - (int)lo) * (int)(p & 4095)) >> 12);
        cm_ptab_put<1>(ptab, node, (u32)(sse * 3 + (int)p) << 14);
        BZ_PROF(0);
        __syncthreads();   // S1: ptab ready
        BZ_PROF_AFTER_BAR(1, ptab + 2);
        // (B) walk my path
        bool ok, ok2;
        {
            const u32 m0 = lds_u32<0>(pa[0]), m1 = lds_u32<0>(pa[1]), m2 = lds_u32<0>(pa[2]), m3 = lds_u32<0>(pa[3]);
            const u32 m4 = lds_u32<0>(pa[4]), m5 = lds_u32<0>(pa[5]), m6 = lds_u32<0>(pa[6]), m7 = lds_u32<0>(pa[7]);
            u32 rk[9];
            u32 zmin = 0xFFFFFFFFu;
            rk[0] = range;
            const u32 mm[8] = {m0, m1, m2, m3, m4, m5, m6, m7};
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const u64 w = cm_mul_wide(rk[k], mm[k]);
                rk[k + 1] = (u32)(w >> 32);
                zmin = min(zmin, (u32)w);
            }
            u32 acc = 0;
#pragma unroll
            for (int k = 0; k < 9; k++) acc += cs[k] * rk[k];
            const u32 r8 = rk[8];
            const u32 d8 = code - low - acc;
            const bool cand = d8 <= r8 && zmin != 0u;
            ok = cand && r8 >= (1u << 24);   // ranges only shrink along a path: r8 is the smallest
            if (ok) {
                sts_u32<0>(pub_a, v);
                sts_u32<4>(pub_a, code - d8);
                sts_u32<8>(pub_a, r8);
            }
            BZ_PROF(2);
            int won = __syncthreads_or(ok ? 1 : 0);   // S2: byte and state published (or nobody won)
            BZ_PROF_AFTER_BAR(3, pub);
            if (!won) {
#ifdef BZ_CM_PROFILE
                _acc[6]++;
#endif
                // range < 2^24 is only NECESSARY for a shift (low and high may straddle a top-byte boundary):
                // every candidate applies the reference's test to its own path before the byte is redone serially
                ok2 = false;
                if (cand) {
                    u32 al = low, tmin = 0xFFFFFFFFu;
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        al += zk[k] & (rk[k] - rk[k + 1]);
                        tmin = min(tmin, al ^ (al + rk[k + 1]));
                    }
                    ok2 = tmin >= (1u << 24);
                    if (ok2) {
                        sts_u32<0>(pub_a, v);
                        sts_u32<4>(pub_a, code - d8);
                        sts_u32<8>(pub_a, r8);
                    }
                }
                won = __syncthreads_or(ok2 ? 1 : 0);
                if (!won) {
#ifdef BZ_CM_PROFILE
                    _acc[7]++;
#endif
                    if (tid == 0) {   // exact serial decoder for this byte (reference loop)
                        u32 flow = low, frange = range, fcode = code;
                        s32 fip = ip;
                        const u32 nd = cm_dec_exact_levels<8>(ptab, 1u, flow, frange, fcode, fip, insize, scode);
                        pub[0] = nd & 255u;
                        pub[1] = flow;
                        pub[2] = frange;
                        pub[3] = fcode;
                        pub[4] = (u32)fip;
                    }
                    __syncthreads();
                    code = pub[3];
                    ip = (s32)pub[4];
                }
                BZ_PROF_AFTER_BAR(4, pub);
            }
        }
        const u32 byte = lds_u32<0>(pub_a);
        low = lds_u32<4>(pub_a);
        range = lds_u32<8>(pub_a);
        if (tid == 0) out[i] = (u8)byte;
        // (C) learn byte i
        const bool on = node != 0 && ((256u | byte) >> sh) == (u32)node;
        const u32 ones = ((byte >> (sh - 1)) & 1u) ? 0xFFFFu : 0u;
        const u32 na = cm_adapt_bf(a, ones, 2), nb = cm_adapt_bf(b, ones, 4);
        if (on) {
            *q0 = (u16)na;
            *q1 = (u16)nb;
            cell[0] = (u16)cm_adapt_bf(lo, ones, 6);
            cell[1] = (u16)cm_adapt_bf(hi, ones, 6);
        }
        a = on ? na : a;
        d = on ? nb : b;
        prev2 = prev1;
        prev1 = (int)byte;
        q1 = c1col + prev1 * 256;
        b = *q1;
        // keep the stream window ahead of the read position (uniform)
        if (ip - wlo >= 1024) {
            for (int k = tid; k < 1024; k += kCmDecP2Threads) {
                const s32 src = wlo + 2048 + k;
                scode[src & 2047] = (src < insize) ? in[src] : 0;
            }
            wlo += 1024;
        }
        BZ_PROF(5);
    }
#ifdef BZ_CM_PROFILE
    if (tid == 0)
        for (int k = 0; k < 8; k++) g_cm_prof[24 + k] = _acc[k];   // predict, wait S1, walk, wait S2, fallback, learn, #slow, #serial
#endif
}

// ---- walker warps + model threads (variant 6) -------------------------------------------------------------
// Measured on B200 (profiles/r01_cm_phase_cycles_1MiB.log): the all-paths walk of variant 5 takes 188 + 47
// cycles per byte (walk + publish barrier) against 443 + 228 for the serial chain warp of variant 0, but
// variant 5 pays for the model (predict 161 + learn 234) on the critical path because the same threads do
// both.  Here the two are separated again: threads 0..255 are walkers (one root-to-leaf path each, as in
// variant 5), threads 256..511 are the model threads of variant 0/4 (one tree node each, counters in
// registers, speculating that the byte repeats while the walkers walk).  Per byte:
//     model    [predict -> B1]  precompute both update outcomes, predict byte i+1 under "byte repeats"  B2  learn
//     walkers  [B1]  8 loads, 8 dependent IMAD.WIDE, final test, winner publishes               B2  read state
// B1 is skipped on a speculation hit.  B2 is a barrier-OR: when no walker passed the (sufficient) fast test the
// candidates apply the reference's exact shift test to their own paths (second barrier-OR) and only if the true
// path really needs a shift thread 0 redoes the byte with the reference loop (third barrier).
constexpr int kCmDecW6Threads = 512;
constexpr size_t kCmDecW6SmemBytes = (size_t)kCmTableU16 * 2 + 2 * 256 * 8 + 2048 + 128 + 256 * 28 * 4;

struct CmWalkConsts {
    SmemAddr pa[8];   // shared address of this path's {M | -M} word at level k (half 0)
    u32 cs[9];        // +1 / 0 / -1 (mod 2^32): coefficient of range_k in  low_8 - low_0
    u32 zk[8];        // all-ones where the path takes the 0-branch (exact test only)
};
struct CmWalkState {
    u32 low, range, code, prevb;
    s32 ip, wlo;
    bool have;
#ifdef BZ_CM_PROFILE
    unsigned long long _acc[8], _t0, _t1;
#endif
};

template <int HALF, int PRUNE>
BZ_D void cm_dec_walk_step(const CmWalkConsts& K, CmWalkState& S, const s32 i, u32* ptab, u8* scode, volatile u32* pub,
                           SmemAddr pub_a, const u8* __restrict__ in, const s32 insize, u8* __restrict__ out,
                           const u32 v) {
    if (!S.have) __syncthreads();   // B1: ptab ready (skipped when the speculation of the model threads hit)
    BZ_SPROF_AFTER_BAR(S, 0, ptab + HALF * 512 + 2);
    constexpr int HB = HALF * 2048;   // byte offset of this byte's half of ptab
    constexpr int PB = HALF * 32;     // byte offset of this byte's publication slot
    u32 rk[9];
    u32 zmin = 0xFFFFFFFFu;
    rk[0] = S.range;
    {
        const u32 mm[8] = {lds_u32<HB>(K.pa[0]), lds_u32<HB>(K.pa[1]), lds_u32<HB>(K.pa[2]), lds_u32<HB>(K.pa[3]),
                           lds_u32<HB>(K.pa[4]), lds_u32<HB>(K.pa[5]), lds_u32<HB>(K.pa[6]), lds_u32<HB>(K.pa[7])};
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const u64 w = cm_mul_wide(rk[k], mm[k]);
            rk[k + 1] = (u32)(w >> 32);
            zmin = min(zmin, (u32)w);
        }
        // PRUNE: the 32 lanes of a walker warp share the first three branches, so after three levels the whole
        // warp knows (uniformly, no vote) whether the code lies in its eighth of the interval at all; seven of
        // the eight walker warps stop here and free their issue slots.  (If a shift was due in these levels
        // the test may fail for every warp -- then nobody wins and the byte is redone serially, as it would be
        // anyway.)
        bool alive = true;
        if (PRUNE) {
            const u32 a3 = S.low + (K.zk[0] & (rk[0] - rk[1])) + (K.zk[1] & (rk[1] - rk[2])) + (K.zk[2] & (rk[2] - rk[3]));
            alive = (S.code - a3) <= rk[3];
        }
        if (alive) {
#pragma unroll
            for (int k = 3; k < 8; k++) {
                const u64 w = cm_mul_wide(rk[k], mm[k]);
                rk[k + 1] = (u32)(w >> 32);
                zmin = min(zmin, (u32)w);
            }
        } else {
            zmin = 0u;   // not a candidate
#pragma unroll
            for (int k = 3; k < 8; k++) rk[k + 1] = 0u;
        }
    }
    u32 acc = 0;
#pragma unroll
    for (int k = 0; k < 9; k++) acc += K.cs[k] * rk[k];
    const u32 r8 = rk[8];
    const u32 d8 = S.code - S.low - acc;
    const bool cand = d8 <= r8 && zmin != 0u;
    const bool ok = cand && r8 >= (1u << 24);   // ranges only shrink along a path: r8 is the smallest
    if (ok) {
        sts_u32<PB + 0>(pub_a, v);
        sts_u32<PB + 4>(pub_a, S.code - d8);
        sts_u32<PB + 8>(pub_a, r8);
        sts_u32<64 + HALF * 4>(pub_a, v);   // the byte slot the model threads read
    }
    BZ_SPROF(S, 1);
    const int won1 = __syncthreads_or(ok ? 1 : 0);   // B2: byte and state published -- or nobody passed the fast test
    BZ_SPROF_AFTER_BAR(S, 2, pub + HALF * 8);
    if (!won1) {
        BZ_SCOUNT(S, 5);
        bool ok2 = false;
        if (cand) {   // reference's shift test on my own path
            u32 al = S.low, tmin = 0xFFFFFFFFu;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                al += K.zk[k] & (rk[k] - rk[k + 1]);
                tmin = min(tmin, al ^ (al + rk[k + 1]));
            }
            ok2 = tmin >= (1u << 24);
            if (ok2) {
                sts_u32<PB + 0>(pub_a, v);
                sts_u32<PB + 4>(pub_a, S.code - d8);
                sts_u32<PB + 8>(pub_a, r8);
                sts_u32<64 + HALF * 4>(pub_a, v);
            }
        }
        if (!__syncthreads_or(ok2 ? 1 : 0)) {
            BZ_SCOUNT(S, 6);
            if (v == 0) {   // exact serial decoder for this byte (reference loop)
                u32 flow = S.low, frange = S.range, fcode = S.code;
                s32 fip = S.ip;
                const u32 nd = cm_dec_exact_levels<8>(ptab + HALF * 512, 1u, flow, frange, fcode, fip, insize, scode);
                pub[HALF * 8 + 0] = nd & 255u;
                pub[HALF * 8 + 1] = flow;
                pub[HALF * 8 + 2] = frange;
                pub[HALF * 8 + 3] = fcode;
                pub[HALF * 8 + 4] = (u32)fip;
                pub[16 + HALF] = nd & 255u;
            }
            __syncthreads();
            S.code = pub[HALF * 8 + 3];
            S.ip = (s32)pub[HALF * 8 + 4];
            // the stream position only moves here (shifts happen in the serial redo), so this is also the only
            // place where the window can need a refill -- not a test on the common path
            if (S.ip - S.wlo >= 1024) {   // uniform; visibility to the serial reader is ordered by the next barrier
                for (int k = (int)v; k < 1024; k += 256) {
                    const s32 src = S.wlo + 2048 + k;
                    scode[src & 2047] = (src < insize) ? in[src] : 0;
                }
                S.wlo += 1024;
            }
        }
        BZ_SPROF_AFTER_BAR(S, 3, pub + HALF * 8);
    }
    const u32 byte = lds_u32<PB + 0>(pub_a);
    S.low = lds_u32<PB + 4>(pub_a);
    S.range = lds_u32<PB + 8>(pub_a);
    if (v == 0) out[i] = (u8)byte;
    S.have = byte == S.prevb;
    S.prevb = byte;
    BZ_SPROF(S, 4);
}

template <int SLIM, int PRUNE>
__global__ void __launch_bounds__(kCmDecW6Threads, 1) cm_decode_walkers_kernel(const u8* __restrict__ in, s32 insize,
                                                                              u8* __restrict__ out, s32 n) {
    BZ_DYN_SMEM(u16, cm_smem);
    u32* ptab = reinterpret_cast<u32*>(cm_smem + kCmTableU16);  // [2][256] {M, -M}; byte i uses half i&1
    u8* scode = reinterpret_cast<u8*>(ptab + 1024);             // [2048] window of the compressed stream
    // [0..7], [8..15]: publication slot of even / odd bytes (byte, low, range, code, ip); [16], [17]: the byte
    // again, where the model threads look for it
    volatile u32* pub = reinterpret_cast<volatile u32*>(scode + 2048);
    cm_tables_init_smem(cm_smem);
    const int tid = threadIdx.x;
    for (int k = tid; k < 2048; k += kCmDecW6Threads) scode[k] = (k < insize) ? in[k] : 0;
    if (tid >= 256) {
        __syncthreads();
        if (SLIM == 2) cm_dec_model_thread_slim2<1>(cm_smem, ptab, pub + 16, n, tid - 256);
        else if (SLIM == 1) cm_dec_model_thread_slim<1, 1>(cm_smem, ptab, pub + 16, n, tid - 256);
        else cm_dec_model_thread<1, 1>(cm_smem, ptab, pub + 16, n, tid - 256);
        return;
    }
    // ---------------------------------------------------------------------- walker: leaf v = tid
    const u32 v = (u32)tid;
    CmWalkConsts K;
    {
        // level k visits node (1 << k) | (v >> (8 - k)) and takes branch bit 7-k of v
        u32 zprev = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const u32 b = (v >> (7 - k)) & 1u;
            const u32 nodek = (1u << k) | (v >> (8 - k));
            K.pa[k] = smem_addr_of(ptab) + 4 * (nodek * 2 + (b ? 0u : 1u));
            K.zk[k] = b ? 0u : 0xFFFFFFFFu;
            const u32 z = b ? 0u : 1u;
            K.cs[k] = z - zprev;
            zprev = z;
        }
        K.cs[8] = 0u - zprev;
        volatile u32* scr = pub + 32 + 28 * v;   // private scratch behind the publication slots
#if !defined(BZ_EMU)
        launder_u32(K.pa, scr);
#endif
        launder_u32(K.cs, scr + 8);
        launder_u32(K.zk, scr + 17);
    }
    __syncthreads();
    CmWalkState S;
    S.wlo = 0;
    S.ip = 0;
    S.low = 0;
    S.range = 0xFFFFFFFFu;
    S.code = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const u32 add = (S.ip < insize) ? (u32)scode[S.ip & 2047] : 0xFFFFFFFFu;  // read_in() past the end adds -1
        S.ip += (S.ip < insize);
        S.code = (S.code << 8) + add;
    }
    S.have = false;
    S.prevb = 0;
    const SmemAddr pub_a = smem_addr_of(pub);
#ifdef BZ_CM_PROFILE
    for (int k = 0; k < 8; k++) S._acc[k] = 0;
    S._t0 = clock64();
#endif
    for (s32 i = 0; i < n; i += 2) {
        cm_dec_walk_step<0, PRUNE>(K, S, i, ptab, scode, pub, pub_a, in, insize, out, v);
        if (i + 1 < n) cm_dec_walk_step<1, PRUNE>(K, S, i + 1, ptab, scode, pub, pub_a, in, insize, out, v);
    }
#ifdef BZ_CM_PROFILE
    if (v == 0)
        for (int k = 0; k < 8; k++) g_cm_prof[36 + k] = S._acc[k];   // wait B1, walk, wait B2, slow path, tail, #slow, #serial
#endif
}

#if defined(__CUDACC__) || defined(BZ_EMU)
inline cudaError_t cm_set_smem_attrs() {
    BZ_CUDA_TRY(cudaFuncSetAttribute(cm_encode_single_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCmSmemBytes));
    BZ_CUDA_TRY(cudaFuncSetAttribute(cm_decode_single_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCmSmemBytes));
    BZ_CUDA_TRY(cudaFuncSetAttribute(cm_encode_chunked_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCmEncSmemBytes));
    BZ_CUDA_TRY(cudaFuncSetAttribute(cm_encode_chunked_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCmEncSmemBytes));
    BZ_CUDA_TRY(cudaFuncSetAttribute(cm_encode_chunked_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCmEncSmemBytes));
    BZ_CUDA_TRY(cudaFuncSetAttribute(cm_encode_chunked_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCmEncSmemBytes));
    BZ_CUDA_TRY(cudaFuncSetAttribute(cm_decode_tree_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCmDecSmemBytes));
    BZ_CUDA_TRY(cudaFuncSetAttribute(cm_decode_paths_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCmDecSmemBytes));
    BZ_CUDA_TRY(cudaFuncSetAttribute(cm_decode_lanes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCmDecLanesSmemBytes));
    BZ_CUDA_TRY(cudaFuncSetAttribute(cm_decode_paths2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCmDecP2SmemBytes));
    BZ_CUDA_TRY((cudaFuncSetAttribute(cm_decode_walkers_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCmDecW6SmemBytes)));
    BZ_CUDA_TRY((cudaFuncSetAttribute(cm_decode_walkers_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCmDecW6SmemBytes)));
    BZ_CUDA_TRY((cudaFuncSetAttribute(cm_decode_walkers_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCmDecW6SmemBytes)));
    BZ_CUDA_TRY((cudaFuncSetAttribute(cm_decode_walkers_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCmDecW6SmemBytes)));
    return cudaSuccess;
}
#endif

#endif  // BZ_DEVICE_CODE

}  // namespace bz3
