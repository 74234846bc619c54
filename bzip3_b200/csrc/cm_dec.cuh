// cm_dec.cuh -- entropy decoder of the block codec: decode_bytes (reference src/libbz3.c:435-494) for one block in
// one thread block.
//
// Decoding is one dependent chain per block: the probability of a decision depends on the bits decoded before it
// (src/libbz3.c:453-489).  What does not depend on the bits of the CURRENT byte is the probability of every one of
// the 255 tree nodes (no node is visited twice within a byte; prev1 / prev2 / run flag are fixed at the byte
// boundary), so the work is split into
//   warp 0        the walker: the 8 decisions of a byte, one multiply per decision, reading the node multipliers
//                 M = P << 14 from a 255-entry table in shared memory (ptab),
//   warps 1..8    256 model threads; thread t owns tree node t and is the only one that ever touches that node's
//                 counters (c0 in a register, its column of c1 and its SSE rows in shared memory).  After every byte
//                 it learns the byte (if the node was on its path) and predicts the next one into ptab.
// Latencies this is written against (measured on B200, profiles/r01_ubench_b200.log): dependent mul.hi 10 cycles,
// dependent add / setp / selp ~3.4, shared load 23, taken branch ~23, barrier hand-off 30-40.
//
// Walker.  The chain of a decision is  x = hi32(range * M) -> bit = (code - low <= x) -> select range / M of the
// child: ~17 cycles.  The table lookups are NOT on that chain: the multipliers of the four grandchildren of a node
// are contiguous (heap order: 4*node .. 4*node+3), so one 128-bit shared load issued when a node becomes known
// brings in every candidate of the level three below it, two decisions before it is needed; the candidates are
// narrowed by selects.  Renormalisation is handled in place: a shift can only be due when range < 2^24, which is one
// compare per decision and a rarely taken branch to the reference's loop (:473-477) -- no second tier, no redo.
// The walker keeps  d = code - low  instead of low; code never leaves [low, high] as long as the stream lasts
// (`bit = code <= mid` keeps it inside, a shift preserves it), so the two forms agree.  They differ once read_in()
// runs past the end of the payload (:345: the int -1 is added); from that decision on the walker switches to the
// reference's absolute form for the rest of the block.
//
// Model threads speculate: while the walker works on byte i they predict byte i+1 under the hypothesis "byte i
// repeats byte i-1" (the common case in BWT output) into the other half of ptab.  On a hit the walker continues
// at once; on a miss the model threads learn the real byte, predict again and the walker waits for that.
//
// Hand-offs are named barriers used as producer / consumer pairs (bar.arrive by the producer, bar.sync by the
// consumer), all indexed by the parity of the byte so that a fast party can never arrive at a barrier generation
// the slow party has not left yet:
//   B[h]   walker -> model threads   "byte i is in vbyte[h]"
//   S[h]   model threads -> walker   "the speculative table for byte i is in ptab[h]"
//   R[h]   model threads -> walker   "the real table for byte i is in ptab[h]" (byte 0 and after a miss)
#pragma once
#include "cm.cuh"

namespace bz3 {

#if defined(BZ_DEVICE_CODE)

constexpr int kCmD2Threads = 288;   // warp 0: walker; warps 1..8: model threads (node = tid - 32)
constexpr size_t kCmD2SmemBytes = (size_t)kCmTableU16 * 2 + 2 * 256 * 4 + 2048 + 64;

// named barriers (0 is __syncthreads)
constexpr int kBarB = 1, kBarS = 3, kBarR = 5;

template <int ID>
BZ_D void nb_sync() {
#if defined(BZ_EMU)
    ::emu::bar_sync(ID, kCmD2Threads);
#else
    asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(kCmD2Threads) : "memory");
#endif
}
template <int ID>
BZ_D void nb_arrive() {
#if defined(BZ_EMU)
    ::emu::bar_arrive(ID, kCmD2Threads);
#else
    asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(kCmD2Threads) : "memory");
#endif
}

#ifdef BZ_CM_PROFILE
#define BZ_DPROF_DECL unsigned long long _acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, _t0 = clock64(), _t1
#define BZ_DPROF(slot) do { _t1 = clock64(); _acc[slot] += _t1 - _t0; _t0 = _t1; } while (0)
#define BZ_DPROF_COUNT(slot) (_acc[slot]++)
#else
#define BZ_DPROF_DECL
#define BZ_DPROF(slot)
#define BZ_DPROF_COUNT(slot)
#endif

// ------------------------------------------------------------------------------------------------ model thread
struct CmPredState {
    int prev1, prev2;
    u32 run, a, b, d, lo, hi;
    u16* q1;     // this node's cell of row prev1 of c1
    u16* cell;   // SSE cell pair used for the current byte
};

// real prediction of the next byte from the registers into half H of ptab (byte 0, and after a miss)
template <int H>
BZ_D void cm_pred_real(CmPredState& M, u32* ptab, u16* rows, const int node) {
    M.run = (M.prev1 == M.prev2) ? M.run + 1 : 0;   // src/libbz3.c:367-372
    const int flag = M.run > 2;
    const u32 p = ((M.a + M.b) * 7 + M.d + M.d) >> 4;
    M.cell = rows + flag * 17 + (p >> 12);
    M.lo = M.cell[0];
    M.hi = M.cell[1];
    const int sse = (int)M.lo + ((((int)M.hi - (int)M.lo) * (int)(p & 4095)) >> 12);
    ptab[H * 256 + node] = (u32)(sse * 3 + (int)p) << 14;
}

// One byte (index parity H) as seen by the owner of `node`.
template <int H>
BZ_D void cm_pred_step(CmPredState& M, u32* ptab, volatile u32* vbyte, const int node, const int sh, u16* c1col, u16* rows) {
    // ---- speculation: this byte == prev1.  Then prev1' = prev2' = prev1, both order-1 inputs of the next byte are
    // this thread's current order-1 counter (updated if the node is on the path of prev1).
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
    {   // the pending update of this byte is not in shared memory yet
        const bool same = on_h && cell_s == M.cell, up = on_h && cell_s == M.cell + 1, dn = on_h && cell_s + 1 == M.cell;
        lo_s = same ? nl : (up ? nh : lo_s);
        hi_s = same ? nh : (dn ? nl : hi_s);
    }
    {
        const int sse = (int)lo_s + ((((int)hi_s - (int)lo_s) * (int)(p_s & 4095)) >> 12);
        ptab[(H ^ 1) * 256 + node] = (u32)(sse * 3 + (int)p_s) << 14;
    }
    nb_arrive<kBarS + (H ^ 1)>();
    nb_sync<kBarB + H>();
    const u32 byte = vbyte[H];
    if (byte != hyp) {   // uniform across the block
        // miss: learn the byte that really came, then predict the next one for real
        u32 na = M.a, nb = M.b;
        if (node != 0 && ((256u | byte) >> sh) == (u32)node) {
            const u32 ones = ((byte >> (sh - 1)) & 1u) ? 0xFFFFu : 0u;
            na = cm_adapt_bf(M.a, ones, 2);
            nb = cm_adapt_bf(M.b, ones, 4);
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
        cm_pred_real<H ^ 1>(M, ptab, rows, node);
        nb_arrive<kBarR + (H ^ 1)>();
        return;
    }
    if (on_h) {   // hit: the speculative outcome is the real one
        *M.q1 = (u16)b_s;
        M.cell[0] = (u16)nl;
        M.cell[1] = (u16)nh;
    }
    M.a = a_s;
    M.b = b_s;
    M.d = b_s;
    M.lo = lo_s;
    M.hi = hi_s;
    M.cell = cell_s;
    M.run = run_s;
    M.prev2 = M.prev1;   // == byte
}

BZ_D void cm_pred_thread(u16* cm_smem, u32* ptab, volatile u32* vbyte, s32 n, const int node) {
    const int sh = node ? 8 - (31 - __clz(node)) : 8;                 // (256|byte) >> sh == node <=> on the path
    u16* const c1col = cm_smem + kCmC0 + node;                        // + prev * 256
    u16* const rows = cm_smem + kCmC0 + kCmC1 + (2 * node) * 17;      // + flag * 17 + cell
    CmPredState M;
    M.prev1 = 0;
    M.prev2 = 0;
    M.run = 0;
    M.q1 = c1col;
    M.a = 32768u;   // c0[node]: lives in this register only
    M.b = *M.q1;
    M.d = M.b;
    M.lo = 0;
    M.hi = 0;
    M.cell = rows;
    if (n <= 0) return;
    cm_pred_real<0>(M, ptab, rows, node);
    nb_arrive<kBarR + 0>();
    for (s32 i = 0; i < n; i += 2) {
        cm_pred_step<0>(M, ptab, vbyte, node, sh, c1col, rows);
        if (i + 1 < n) cm_pred_step<1>(M, ptab, vbyte, node, sh, c1col, rows);
    }
}

// ------------------------------------------------------------------------------------------------ walker
struct CmWalk {
    u32 r;        // high - low
    u32 d;        // code - low            (fast form)
    u32 code;
    u32 low;      // only maintained in the absolute form
    s32 ip;       // next payload byte to read
    s32 insize;
    bool absolute;   // read_in() has run past the end: reference form from here on
    const u8* scode; // 2 KiB window of the payload in shared memory
};

// the reference's decision loop in its own (absolute) form, levels k0..7 of a byte; node multipliers looked up as they
// are needed.  Only used once the payload has run out (truncated / hostile blocks).
BZ_D u32 cm_walk_absolute(const u32* __restrict__ pt, int k0, u32 node, CmWalk& W) {
    u32 low = W.low, high = W.low + W.r, code = W.code;
    for (int k = k0; k < 8; k++) {
        const u32 x = __umulhi(high - low, pt[node]);
        const u32 mid = low + x;
        const u32 bit = code <= mid;
        if (bit) high = mid; else low = mid + 1u;
        while ((low ^ high) < (1u << 24)) {
            low <<= 8;
            high = (high << 8) | 0xFFu;
            const u32 add = (W.ip < W.insize) ? (u32)W.scode[W.ip & 2047] : 0xFFFFFFFFu;   // read_in(), :345
            W.ip += (W.ip < W.insize);
            code = (code << 8) + add;
        }
        node = node * 2 + bit;
    }
    W.low = low;
    W.r = high - low;
    W.code = code;
    return node;
}

// range < 2^24 after a decision: apply the reference's shift loop (:473-477).  Returns true when a byte was read
// past the end of the payload (the caller then continues in the absolute form).
BZ_D bool cm_walk_shift(CmWalk& W) {
    u32 low = W.code - W.d, high = low + W.r, code = W.code;
    bool past = false;
    while ((low ^ high) < (1u << 24)) {
        low <<= 8;
        high = (high << 8) | 0xFFu;
        const bool in = W.ip < W.insize;
        const u32 add = in ? (u32)W.scode[W.ip & 2047] : 0xFFFFFFFFu;
        W.ip += in;
        past = past || !in;
        code = (code << 8) + add;
    }
    W.r = high - low;
    W.code = code;
    W.d = code - low;
    W.low = low;
    return past;
}

// One decision of the walk at tree depth K (see the header of this file).  mc = multiplier of the current node,
// (ma, mb) = multipliers of its two children, qp = the four grandchildren candidates, loaded one decision ago.
#define BZ_CM_WALK_LEVEL(K)                                                         \
    {                                                                               \
        const u32 x = __umulhi(r, mc);                                              \
        const bool bit = d <= x;                                                    \
        const u32 nx = ~x;                                                          \
        r = bit ? x : r + nx; /* bit 1: high = mid; bit 0: low = mid + 1 */         \
        d = bit ? d : d + nx;                                                       \
        node = node * 2 + (bit ? 1u : 0u);                                          \
        if ((K) < 7) mc = bit ? mb : ma;                                            \
        if ((K) < 6) {                                                              \
            ma = bit ? qp.z : qp.x;                                                 \
            mb = bit ? qp.w : qp.y;                                                 \
        }                                                                           \
        if ((K) < 5) qp = pt4[node]; /* candidates of level K + 3 */                \
    }

// One byte.  pt = this byte's table (u32 M per node, heap order).  The eight decisions are unrolled twice: the first
// copy is the straight-line path of a byte without a shift (its only branches are not taken); a decision that leaves
// range < 2^24 jumps out to the shift loop, and the byte is finished in the second copy, entered at the next level.
BZ_D u32 cm_walk_byte(const u32* __restrict__ pt, CmWalk& W) {
    if (__builtin_expect(W.absolute, 0)) return cm_walk_absolute(pt, 0, 1u, W);
    const uint4* pt4 = reinterpret_cast<const uint4*>(pt);
    const uint4 q0 = pt4[0];   // -, M1, M2, M3
    uint4 qp = pt4[1];         // M4 .. M7: candidates of level 2
    u32 mc = q0.y, ma = q0.z, mb = q0.w;
    u32 node = 1;
    u32 r = W.r, d = W.d;
    int k;
    BZ_CM_WALK_LEVEL(0) if (__builtin_expect(r < (1u << 24), 0)) { k = 0; goto shift; }
    BZ_CM_WALK_LEVEL(1) if (__builtin_expect(r < (1u << 24), 0)) { k = 1; goto shift; }
    BZ_CM_WALK_LEVEL(2) if (__builtin_expect(r < (1u << 24), 0)) { k = 2; goto shift; }
    BZ_CM_WALK_LEVEL(3) if (__builtin_expect(r < (1u << 24), 0)) { k = 3; goto shift; }
    BZ_CM_WALK_LEVEL(4) if (__builtin_expect(r < (1u << 24), 0)) { k = 4; goto shift; }
    BZ_CM_WALK_LEVEL(5) if (__builtin_expect(r < (1u << 24), 0)) { k = 5; goto shift; }
    BZ_CM_WALK_LEVEL(6) if (__builtin_expect(r < (1u << 24), 0)) { k = 6; goto shift; }
    BZ_CM_WALK_LEVEL(7) if (__builtin_expect(r < (1u << 24), 0)) { k = 7; goto shift; }
    W.r = r;
    W.d = d;
    return node;
shift:
    for (;;) {   // decision k left range < 2^24
        W.r = r;
        W.d = d;
        if (cm_walk_shift(W)) {
            W.absolute = true;
            return cm_walk_absolute(pt, k + 1, node, W);
        }
        r = W.r;
        d = W.d;
        switch (k) {
            case 0: BZ_CM_WALK_LEVEL(1) if (r < (1u << 24)) { k = 1; continue; }
            case 1: BZ_CM_WALK_LEVEL(2) if (r < (1u << 24)) { k = 2; continue; }
            case 2: BZ_CM_WALK_LEVEL(3) if (r < (1u << 24)) { k = 3; continue; }
            case 3: BZ_CM_WALK_LEVEL(4) if (r < (1u << 24)) { k = 4; continue; }
            case 4: BZ_CM_WALK_LEVEL(5) if (r < (1u << 24)) { k = 5; continue; }
            case 5: BZ_CM_WALK_LEVEL(6) if (r < (1u << 24)) { k = 6; continue; }
            case 6: BZ_CM_WALK_LEVEL(7) if (r < (1u << 24)) { k = 7; continue; }
            default: break;
        }
        W.r = r;
        W.d = d;
        return node;
    }
}

__global__ void __launch_bounds__(kCmD2Threads, 1) cm_decode_kernel(const u8* __restrict__ in, s32 insize, u8* __restrict__ out, s32 n) {
    BZ_DYN_SMEM(u16, cm_smem);
    u32* ptab = reinterpret_cast<u32*>(cm_smem + kCmTableU16);  // [2][256]  M = P << 14 per node; byte i uses half i & 1
    u8* scode = reinterpret_cast<u8*>(ptab + 512);              // [2048] window of the payload
    volatile u32* vbyte = reinterpret_cast<volatile u32*>(scode + 2048);   // [2] decoded byte of step i in slot i & 1
    cm_tables_init_smem(cm_smem);
    const int tid = threadIdx.x;
    if (tid < 32)
        for (int k = tid; k < 2048; k += 32) scode[k] = (k < insize) ? in[k] : 0;
    __syncthreads();
    if (tid >= 32) {
        cm_pred_thread(cm_smem, ptab, vbyte, n, tid - 32);
        return;
    }
    if (n <= 0) return;
    // ---------------------------------------------------------------------- walker warp (all lanes identical)
    CmWalk W;
    W.insize = insize;
    W.scode = scode;
    W.ip = 0;
    W.absolute = false;
    W.low = 0;
    W.r = 0xFFFFFFFFu;
    W.code = 0;
    s32 wlo = 0;  // the window holds payload bytes [wlo, wlo + 2048)
    bool past = false;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const bool inb = W.ip < insize;
        const u32 add = inb ? (u32)scode[W.ip & 2047] : 0xFFFFFFFFu;  // read_in() past the end adds -1
        W.ip += inb;
        past = past || !inb;
        W.code = (W.code << 8) + add;
    }
    W.d = W.code;           // low = 0
    W.absolute = past;      // a payload shorter than 4 bytes: the reference form from the start
    u32 prevb = 0;
    BZ_DPROF_DECL;
    nb_sync<kBarR + 0>();   // table of byte 0
    BZ_DPROF(0);
    for (s32 i = 0; i < n; i += 2) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            if (i + h < n) {
                const u32 byte = cm_walk_byte(ptab + h * 256, W) & 255u;
                BZ_DPROF(1);
                // every lane holds the same byte: convergent stores of one value to one address
                vbyte[h] = byte;
                out[i + h] = (u8)byte;
                if (h == 0) nb_arrive<kBarB + 0>(); else nb_arrive<kBarB + 1>();
                if (W.ip - wlo >= 1024) {  // uniform in the warp; the window belongs to this warp alone
                    __syncwarp();
                    for (int k = tid; k < 1024; k += 32) {
                        const s32 src = wlo + 2048 + k;
                        scode[src & 2047] = (src < insize) ? in[src] : 0;
                    }
                    wlo += 1024;
                    __syncwarp();
                }
                const bool hit = byte == prevb;
                prevb = byte;
                if (h == 0) nb_sync<kBarS + 1>(); else nb_sync<kBarS + 0>();
                BZ_DPROF(2);
                if (!hit) {
                    if (h == 0) nb_sync<kBarR + 1>(); else nb_sync<kBarR + 0>();
                    BZ_DPROF_COUNT(4);
                }
                BZ_DPROF(3);
            }
        }
    }
#ifdef BZ_CM_PROFILE
    if (tid == 0)
        for (int k = 0; k < 5; k++) g_cm_prof[8 + k] = _acc[k];   // first table, walk, publish + wait S, wait R, #misses
#endif
}

#endif  // BZ_DEVICE_CODE

}  // namespace bz3
