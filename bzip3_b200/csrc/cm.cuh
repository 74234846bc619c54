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

BZ_D void cm_tables_init_smem(u16* tab) {
    for (int k = threadIdx.x; k < kCmTableU16; k += blockDim.x) tab[k] = cm_initial(k);
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

// 32 x 32 -> 64 bit product as two registers (one IMAD.WIDE; no 64-bit shift for the compiler to lower into an extra add on
// the chain: -5 % (Zipf) / -8 % (source) encode time, profiles/r02_call17_cm_encoder_branch_free_exact_byte.log)
BZ_D void cm_mul_wide_halves(u32 a, u32 b, u32& lo, u32& hi) {
#if defined(BZ_EMU)
    const u64 w = (u64)a * (u64)b;
    lo = (u32)w;
    hi = (u32)(w >> 32);
#else
    asm("{\n\t.reg .u64 w;\n\tmul.wide.u32 w, %2, %3;\n\tmov.b64 {%0, %1}, w;\n\t}" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
#endif
}

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
// Measured on B200 (profiles/r01_ncu_source_cm_*, r02_call3_ubench_walk.log): one in-order thread pays ~2.5-5 cycles per
// dependent instruction and ~25-50 per conditional branch, taken or not, so a byte is coded in two tiers:
//   fast tier   eight decisions with no branch at all and ONE multiply each: with m = bit ? M : -M (M = P << 14) the new
//               range is hi32(range * m) -- for a 1-bit that is x = (range * P) >> 18, for a 0-bit range - x - 1 unless
//               lo32(range * M) == 0.  low moves by the range lost at 0-bits.  No shift is applied; the intervals are
//               nested, so low ^ (low + range) after the eighth decision tells whether one was due anywhere in the byte
//               (most bytes of BWT output: none).
//   exact tier  otherwise the byte is redone from its start state, still without a branch: one PREDICATED one-byte shift
//               per decision; only a decision that needs a second shift (probability < 2^-8) falls back to the
//               reference's loop (src/libbz3.c:388-416).
// Measured steps (profiles/r02_call16_*, r02_call17_*): one multiply instead of multiply + select chain -3 %, predicated
// exact byte instead of a loop with a branch per decision -11 % (Zipf) / -8 % (source).

// entries (m = bit ? M : -M, eight per byte) and the byte itself of position k of the chunk: pinned loads, issued where written
BZ_D void cm_enc_fetch(const uint4* pv, const u8* sb, s32 k, uint4& a, uint4& b, u32& sym) {
#if defined(BZ_EMU)
    a = pv[2 * k];
    b = pv[2 * k + 1];
    sym = sb[k];
#else
    const u32 ap = (u32)__cvta_generic_to_shared(pv + 2 * k);
    const u32 sp = (u32)__cvta_generic_to_shared(sb + k);
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "r"(ap));
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+16];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "r"(ap));
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(sym) : "r"(sp));
#endif
}

// One byte through the range coder lane (tiers: see above).
BZ_D void cm_code_byte(const uint4 ca, const uint4 cb, const u32 cs, u32& low, u32& range, s32& op, u8* __restrict__ out) {
    // one multiply per decision: bit 1: new range = hi32(range * M) = x; bit 0: range - x - 1 = hi32(range * -M)
    // unless lo32(range * M) == 0 (range * (2^32 - M) = range * 2^32 - range * M).  low moves by the range
    // lost at 0-bits.  No shift is applied; the minimum of the low halves tells whether the shortcut was exact.
    const u32 m[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
    u32 l = low, r = range, zmin = 0xFFFFFFFFu;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        u32 rn, wl;
        cm_mul_wide_halves(r, m[j], wl, rn);
        zmin = min(zmin, wl);
        if (!(cs & (0x80u >> j))) l += r - rn;
        r = rn;
    }
    // The intervals are nested and no shift was applied, so once the top bytes of low and low + range agree they agree
    // for the rest of the byte: ONE test after the eighth decision tells whether a shift was due anywhere in the byte
    // (three instructions per decision less than a running minimum).
    if (((l ^ (l + r)) >= (1u << 24)) && zmin != 0u) {
        low = l;
        range = r;
        return;
    }
    // A shift was due somewhere in this byte: redo it exactly -- branch-free, one predicated one-byte shift per decision
    // (the common case: a decision shifts at most one byte out); only a decision that needs a second shift sends the
    // byte to the reference's loop below.
    {
        u32 lo2 = low, rg = range;
        s32 o2 = op;
        bool multi = false;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const bool bit = (cs & (0x80u >> j)) != 0;
            const u32 xx = __umulhi(rg, bit ? m[j] : 0u - m[j]);
            rg = bit ? xx : rg - xx - 1u;
            lo2 = bit ? lo2 : lo2 + xx + 1u;
            const bool sh = ((lo2 ^ (lo2 + rg)) < (1u << 24));
            if (sh) out[o2] = (u8)(lo2 >> 24);
            o2 += sh ? 1 : 0;
            lo2 = sh ? lo2 << 8 : lo2;
            rg = sh ? (rg << 8) | 0xFFu : rg;
            multi = multi || (sh && ((lo2 ^ (lo2 + rg)) < (1u << 24)));
        }
        if (!multi) {
            low = lo2;
            range = rg;
            op = o2;
            return;
        }
    }
    {   // the reference's loop (src/libbz3.c:388-416)
        u32 lo2 = low, high = low + range;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const bool bit = (cs & (0x80u >> j)) != 0;
            const u32 xx = __umulhi(high - lo2, bit ? m[j] : 0u - m[j]);
            if (bit) high = lo2 + xx; else lo2 += xx + 1u;
            while ((lo2 ^ high) < (1u << 24)) {
                out[op++] = (u8)(lo2 >> 24);
                lo2 <<= 8;
                high = (high << 8) | 0xFFu;
            }
        }
        low = lo2;
        range = high - lo2;
    }
}


// Three-stage chunk pipeline (one __syncthreads per chunk, no polling):
//   warp 0, lanes 0..7  stage 1: c0 / c1 counters of tree depth `lane`  -> mixed probability p (16 bit)
//   warp 2, lanes 0..7  stage 2: SSE rows c2 of depth `lane`            -> P << 14
//   warp 1, lane 0      stage 3: range coder
// Stage s works on chunk it-s in iteration `it`.  A node of depth d is only ever coded at bit position d,
// so the eight lanes of a stage own disjoint counters and never synchronise with each other.
__global__ void __launch_bounds__(kCmEncThreads) cm_encode_kernel(const u8* __restrict__ in, s32 n, u8* __restrict__ out, s32* out_size) {
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
        } else if (warp == 2) {
            if (lane < 8 && it >= 1 && it - 1 < nchunks) {
                const s32 ch = it - 1;
                const s32 base = ch * kCmEncChunk;
                const s32 len = (n - base) < kCmEncChunk ? (n - base) : kCmEncChunk;
                const u8* sb = sbytes + (ch % 3) * kCmEncChunk;
                const u16* pm = pmid + (ch & 1) * (kCmEncChunk * 8) + lane;
                u32* pb = pbuf + (ch & 1) * (kCmEncChunk * 8) + lane;
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
                        pb[k * 8] = ones ? m : 0u - m;   // the coder lane multiplies by M for a 1-bit and by -M for a 0-bit
                    }
                    cell[0] = (u16)cm_adapt_bf((u32)lo, ones, 6);
                    cell[1] = (u16)cm_adapt_bf((u32)hi, ones, 6);
                    prev2 = prev1;
                    prev1 = sym;
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
                // Two bytes per trip with two register sets: the entries of byte k + 1 are requested before byte k is coded and
                // nothing has to be copied from "next" to "current" at the back edge (eight moves per byte in a one-byte loop).
                uint4 a0 = pv[0], b0 = pv[1], a1, b1;
                u32 sym0 = sb[0], sym1;
                for (s32 k = 0; k < len; k += 2) {
                    const s32 k1 = (k + 1 < len) ? k + 1 : k;   // past the end: re-read the last byte, the values are unused
                    cm_enc_fetch(pv, sb, k1, a1, b1, sym1);
                    cm_code_byte(a0, b0, sym0, low, range, op, out);
                    if (k + 1 >= len) break;
                    const s32 k2 = (k + 2 < len) ? k + 2 : k + 1;
                    cm_enc_fetch(pv, sb, k2, a0, b0, sym0);
                    cm_code_byte(a1, b1, sym1, low, range, op, out);
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
            out[op++] = (u8)(low >> 24);
            low <<= 8;
        }
        *out_size = op;
    }
}

// Model thread of the decoder: owner of tree node `node` (0 is a dummy); its counters are carried in registers (only this
// thread writes them).  While the chain warp walks byte i the thread SPECULATES that byte i repeats byte i-1 -- the common
// case in BWT output -- and predicts byte i+1 under that hypothesis into the other half of ptab.  On a hit the chain
// continues at once (no predict phase, no second barrier); on a miss the speculation is simply overwritten.
// Model threads and chain warp are co-bottlenecks of the decoder (eight model warps share four schedulers with the chain
// warp), so the loop is written for instruction count and against the cost of a TAKEN branch (~20 cycles): the update
// outcome is computed only for the hypothesised byte and only by the threads on its path (eight of 255) and for any other
// byte after the fact, on a miss; the loop is unrolled by the parity of the byte index (ptab / byte-slot offsets become
// immediates); the real prediction after a miss sits at the END of the step that missed; the learn stores are predicated
// instead of branched over.  Measured against computing both outcomes in every thread: -8.5 % decode time on Zipf text,
// -8.8 % on the source corpus (profiles/r02_call12_cm_slim_model_threads.log).
#if defined(BZ_EMU)
BZ_D void cm_learn_stores(bool on, u16* q0, u32 a, u16* q1, u32 b, u16* cell, u32 lo, u32 hi) {
    if (on) {
        *q0 = (u16)a;
        *q1 = (u16)b;
        cell[0] = (u16)lo;
        cell[1] = (u16)hi;
    }
}
#else
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
template <int H>
BZ_D void cm_model_predict(CmModelState& M, u32* ptab, u16* rows, const int node) {
    M.run = (M.prev1 == M.prev2) ? M.run + 1 : 0;
    const int flag = M.run > 2;
    const u32 p = ((M.a + M.b) * 7 + M.d + M.d) >> 4;
    M.cell = rows + flag * 17 + (p >> 12);
    M.lo = M.cell[0];
    M.hi = M.cell[1];
    const int sse = (int)M.lo + ((((int)M.hi - (int)M.lo) * (int)(p & 4095)) >> 12);
    ptab[H * 256 + node] = (u32)(sse * 3 + (int)p) << 14;
}

template <int HALF>
BZ_D void cm_model_step(CmModelState& M, u16* cm_smem, u32* ptab, volatile u32* vbyte, const int node,
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
        ptab[(HALF ^ 1) * 256 + node] = (u32)(sse * 3 + (int)p_s) << 14;
    }
    __syncthreads();   // byte ready
    const u32 byte = vbyte[HALF];
    if (__builtin_expect(byte != hyp, 0)) {   // uniform across the CTA
        // miss: learn the byte that really came, then predict the next one for real.  No further branch on this path (the
        // chain warp is waiting for it): the learn step is predicated, and the prediction after the last byte of the block is
        // simply made and never used (the chain warp takes the matching barrier after its loop).
        const bool on_b = node != 0 && ((256u | byte) >> sh) == (u32)node;
        const u32 ones = ((byte >> (sh - 1)) & 1u) ? 0xFFFFu : 0u;
        const u32 na = on_b ? cm_adapt_bf(M.a, ones, 2) : M.a;
        const u32 nb = on_b ? cm_adapt_bf(M.b, ones, 4) : M.b;
        cm_learn_stores(on_b, q0, na, M.q1, nb, M.cell, cm_adapt_bf(M.lo, ones, 6), cm_adapt_bf(M.hi, ones, 6));
        M.a = na;
        M.d = nb;                       // this byte's order-1 counter is the next byte's prev2 counter
        M.prev2 = M.prev1;
        M.prev1 = (int)byte;
        M.q1 = c1col + M.prev1 * 256;
        M.b = *M.q1;                    // after the store above in program order
        cm_model_predict<HALF ^ 1>(M, ptab, rows, node);
        __syncthreads();   // ptab ready
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

BZ_D void cm_dec_model_thread(u16* cm_smem, u32* ptab, volatile u32* vbyte, s32 n, const int node) {
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
    cm_model_predict<0>(M, ptab, rows, node);
    __syncthreads();   // ptab of byte 0 ready
    for (s32 i = 0; i < n; i += 2) {
        cm_model_step<0>(M, cm_smem, ptab, vbyte, node, sh, q0, c1col, rows);
        if (i + 1 < n) cm_model_step<1>(M, cm_smem, ptab, vbyte, node, sh, q0, c1col, rows);
    }
}

// ---- tree-parallel decoder ---------------------------------------------------------------------
// Decoding is one dependent chain: the next context depends on the bit just decoded.  What does not
// depend on the bits of the current byte is the probability of every one of the 255 tree nodes (no node
// is visited twice within a byte, and prev1/prev2/runflag are fixed at the byte boundary).  Roles:
//   8 warps     256 model threads, thread owns ONE node and is the only one that ever touches its
//               counters: (C) learn the previous byte if the node was on its path, (A) predict -> ptab
//   warp 0      the chain: walks the 8 levels using ptab, two-tier like the encoder (branch-free fast
//               byte, exact redo when a renormalisation was needed), publishes the byte.
// Two __syncthreads per byte (ptab ready / byte ready).  The compressed bytes are staged through a
// 2 KiB shared window owned by warp 0, so the renormalisation never waits on global memory.
constexpr int kCmDecThreads = 384;   // 12 warps: chain warp 0, model warps 1-3, 5-7, 9-10; warps 4, 8, 11 idle (see below)
constexpr size_t kCmDecSmemBytes = (size_t)kCmTableU16 * 2 + 2 * 256 * 4 + 2048 + 64;

__global__ void __launch_bounds__(kCmDecThreads) cm_decode_kernel(const u8* __restrict__ in, s32 insize,
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
        // A warp is issued by scheduler (warp id mod 4).  The chain warp is the critical path of the whole block, so the
        // warps that would share its scheduler (4, 8) stay idle and the eight model warps live on the other three
        // schedulers (measured: -11 % decode time on Zipf text, -14 % on source, profiles/r02_call14_cm_decoder_warp_spread.log).
        const int w = tid >> 5;
        if (w == 4 || w == 8 || w == 11) return;
        const int slot = w < 4 ? w - 1 : (w < 8 ? w - 2 : w - 3);
        cm_dec_model_thread(cm_smem, ptab, vbyte, n, slot * 32 + (tid & 31));
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
                }
#else
                asm volatile(
                    "{\n\t"
                    ".reg .pred pb;\n\t"
                    ".reg .u32 mid, nx, r0;\n\t"
                    "add.u32 mid, %0, %2;\n\t"
                    "not.b32 nx, %2;\n\t"
                    "setp.le.u32 pb, %5, mid;\n\t"        // bit = code <= low + x
                    "add.u32 r0, %1, nx;\n\t"             // range - x - 1
                    "selp.u32 %1, %2, r0, pb;\n\t"        // bit ? x : range - x - 1
                    "selp.u32 %3, %7, %6, pb;\n\t"        // P of the chosen child
                    "mul.hi.u32 %2, %1, %3;\n\t"          // product for the next step
                    "@!pb add.u32 %0, mid, 1;\n\t"        // bit 0: low = mid + 1
                    "selp.u32 %4, 1, 0, pb;\n\t"
                    "}"
                    : "+r"(flow), "+r"(frange), "+r"(x), "+r"(pcur), "=r"(bit)
                    : "r"(code), "r"(kid0), "r"(kid1));
#endif
                node = node * 2 + bit;
                kid0 = bit ? gk.z : gk.x;
                kid1 = bit ? gk.w : gk.y;
                if (k < 5) gk = *reinterpret_cast<const uint4*>(pt + 4 * node);
            }
            BZ_PROF(1);
            // The intervals of the eight steps are nested and nothing was shifted, so the top bytes of low and low + range,
            // once equal, stay equal: one test after the byte tells whether a renormalisation was due at ANY step.  Nested
            // needs low <= code <= high, which every decision preserves and only read_in() past the end of the payload
            // (the int -1, src/libbz3.c:345, :473) can break: an exhausted stream decodes in the exact tier.
            if (((flow ^ (flow + frange)) >= (1u << 24)) && ip < insize) {
                low = flow;
                range = frange;
            } else {
#ifdef BZ_CM_PROFILE
                _acc[4]++;
#endif
                bool done = false;
                // exact tier A: the same walk with ONE predicated one-byte renormalisation per step and no branch (a branch per
                // step costs more than the step).  The next four payload bytes wait in a register; with t = low ^ high a step
                // that shifts once needs a second shift iff t < 2^16 (the new low ^ high is (t << 8) | 0xFF): such a byte, or one
                // that shifts more than four times, or one within four bytes of the end of the payload, goes to tier B.
                if (ip + 4 <= insize) {
                    u32 W;
#if defined(BZ_EMU)
                    W = ((u32)scode[ip & 2047] << 24) | ((u32)scode[(ip + 1) & 2047] << 16) | ((u32)scode[(ip + 2) & 2047] << 8) |
                        (u32)scode[(ip + 3) & 2047];
#else
                    W = __byte_perm(*reinterpret_cast<const u32*>(scode + (ip & 2044)),
                                    *reinterpret_cast<const u32*>(scode + ((ip + 4) & 2044)), 0x0123u + (u32)(ip & 3) * 0x1111u);
#endif
                    u32 plow = low, prange = range, pcode = code, nsh = 0, multi = 0, pnode = 1;
                    uint4 pg = *reinterpret_cast<const uint4*>(pt + 4);
                    u32 pp = g0.y, pk0 = g0.z, pk1 = g0.w;
                    u32 px = mulhi_pinned(prange, pp);
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        u32 bit;
#if defined(BZ_EMU)
                        {
                            const u32 mid = plow + px;
                            bit = pcode <= mid;
                            prange = bit ? px : prange - px - 1u;
                            pp = bit ? pk1 : pk0;
                            if (!bit) plow = mid + 1u;
                            const u32 t = plow ^ (plow + prange);
                            if (t < (1u << 24)) {
                                plow <<= 8;
                                prange = (prange << 8) | 0xFFu;
                                pcode = (pcode << 8) | (W >> 24);
                                W <<= 8;
                                nsh++;
                                if (t < (1u << 16)) multi = 1;
                            }
                            px = __umulhi(prange, pp);
                        }
#else
                        asm volatile(
                            "{\n\t"
                            ".reg .pred pb, ps, pm;\n\t"
                            ".reg .u32 mid, nx, r0, hi, t;\n\t"
                            "add.u32 mid, %0, %2;\n\t"
                            "not.b32 nx, %2;\n\t"
                            "setp.le.u32 pb, %5, mid;\n\t"
                            "add.u32 r0, %1, nx;\n\t"
                            "selp.u32 %1, %2, r0, pb;\n\t"
                            "selp.u32 %3, %10, %9, pb;\n\t"
                            "@!pb add.u32 %0, mid, 1;\n\t"
                            "selp.u32 %4, 1, 0, pb;\n\t"
                            "add.u32 hi, %0, %1;\n\t"
                            "xor.b32 t, %0, hi;\n\t"
                            "setp.lt.u32 ps, t, 0x1000000;\n\t"
                            "setp.lt.u32 pm, t, 0x10000;\n\t"
                            "@ps shl.b32 %0, %0, 8;\n\t"
                            "@ps mad.lo.u32 %1, %1, 256, 255;\n\t"          // (range << 8) | 0xFF
                            "@ps shf.l.clamp.b32 %5, %6, %5, 8;\n\t"        // code = (code << 8) | next payload byte
                            "@ps shl.b32 %6, %6, 8;\n\t"
                            "@ps add.u32 %7, %7, 1;\n\t"
                            "@pm mov.u32 %8, 1;\n\t"
                            "mul.hi.u32 %2, %1, %3;\n\t"
                            "}"
                            : "+r"(plow), "+r"(prange), "+r"(px), "+r"(pp), "=r"(bit), "+r"(pcode), "+r"(W), "+r"(nsh), "+r"(multi)
                            : "r"(pk0), "r"(pk1));
#endif
                        pnode = pnode * 2 + bit;
                        pk0 = bit ? pg.z : pg.x;
                        pk1 = bit ? pg.w : pg.y;
                        if (k < 5) pg = *reinterpret_cast<const uint4*>(pt + 4 * pnode);
                    }
                    if (multi == 0 && nsh <= 4) {
                        low = plow;
                        range = prange;
                        code = pcode;
                        ip += (s32)nsh;
                        node = pnode;
                        done = true;
                    }
                }
                if (!done) {
                    // exact tier B (rare: a step that shifts twice, more than four shifts in the byte, the end of the payload):
                    // same walk, renormalising after every step with the reference's loop (src/libbz3.c:464-474).
                    // range < 2^24 is necessary for equal top bytes only while low <= high, i.e. until the payload is exhausted
                    // (see above); from then on every step takes the reference's test itself.
                    // (Kept unrolled although it is rare: with a rolled loop here ptxas allocates the fast tier's table rows to
                    // the registers it has just read, the loads issue late and the fast tier loses 100 cycles per byte --
                    // profiles/r02_call26_cm_tier_b_rolled_loses.log.)
                    node = 1;
                    gk = *reinterpret_cast<const uint4*>(pt + 4);
                    pcur = g0.y;
                    kid0 = g0.z;
                    kid1 = g0.w;
                    x = mulhi_pinned(range, pcur);
                    u32 exhausted = ip >= insize;
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
                        if (slow | exhausted) {
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
                            exhausted = ip >= insize;
                        }
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
    if (n > 0 && !have) __syncthreads();   // the model threads predicted once more after a last byte that missed
#ifdef BZ_CM_PROFILE
    if (tid == 0)
        for (int k = 0; k < 5; k++) g_cm_prof[8 + k] = _acc[k];   // wait ptab, fast tier, exact tier, publish+wait byte, #redo
#endif
}

#if defined(__CUDACC__) || defined(BZ_EMU)
inline cudaError_t cm_set_smem_attrs() {
    BZ_CUDA_TRY(cudaFuncSetAttribute(cm_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCmEncSmemBytes));
    BZ_CUDA_TRY(cudaFuncSetAttribute(cm_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCmDecSmemBytes));
    return cudaSuccess;
}
#endif

#endif  // BZ_DEVICE_CODE

}  // namespace bz3
