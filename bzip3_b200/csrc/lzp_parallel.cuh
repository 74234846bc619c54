// lzp_parallel.cuh -- the LZP decoder (one thread block per block of data); the encoder is lzp_scan.cuh.
//
// Same format and same sequential semantics as the single-lane code in lzp.cuh (reference src/libbz3.c:200-241).
#pragma once
#include "common.cuh"
#include "lzp.cuh"

namespace bz3 {

BZ_D u32 lzp_ld32(const u8* __restrict__ p) {
    return (u32)p[0] | ((u32)p[1] << 8) | ((u32)p[2] << 16) | ((u32)p[3] << 24);
}

// ---- bulk decoder ---------------------------------------------------------------------------------------------
// The sequential decoder reads the table only when the input byte is 0xF2; everything between two such bytes
// is a run of literals whose only side effect is  table[hash(context)] = position  for every position -- and
// positions only grow, so those updates are an order-free scatter-max.  One CTA therefore alternates
//   A  find the first 0xF2 in the next (up to) 8 KiB of input: 16 bytes per thread, block-wide min
//   B  copy the literals before it and scatter-max their positions into the table, all threads at once
//   C  one thread resolves the 0xF2 against the (now complete) table exactly like the reference -- plain
//      literal when the slot is empty, escaped literal, or a match, which all threads copy (periodic when it
//      overlaps itself).
// Text with few matches decodes at copy speed instead of one L2/atomic round per 32 bytes; a match costs a few
// barriers and one table round trip.
constexpr int kLzpBulkThreads = 512;
constexpr int kLzpBulkChunk = kLzpBulkThreads * 16;

__global__ void __launch_bounds__(kLzpBulkThreads) lzp_decode_bulk_kernel(const u8* __restrict__ in, s32 n, u8* __restrict__ out,
                                                                          s32 max, s32* __restrict__ lut,
                                                                          s32* __restrict__ result) {
    __shared__ u32 s_last4[kLzpBulkThreads];   // last four bytes of every thread's slice (little end = oldest)
    __shared__ s32 s_first[kLzpBulkThreads / 32];
    __shared__ s32 s_ctl[8];                   // [0] first, [1] kind, [2] count, [3] ref, [4] new ip, [5] status
    __shared__ u32 s_tail;
    const int t = threadIdx.x;
    const u32 lane = lane_id();
    if (n < 4) {
        if (t == 0) *result = -1;
        return;
    }
    if (t < 4) out[t] = in[t];
    s32 ip = 4, op = 4;
    // the four most recent output bytes, most recent in the low byte (uniform across the CTA)
    u32 tail = (u32)in[3] | ((u32)in[2] << 8) | ((u32)in[1] << 16) | ((u32)in[0] << 24);
    s32 status = 0;
    while (ip < n && op < max) {
        s32 lim = n - ip;
        if (max - op < lim) lim = max - op;
        if (lim > kLzpBulkChunk) lim = kLzpBulkChunk;
        // ---- A: my 16 bytes, first escape among them
        const s32 j0 = 16 * t;
        u8 by[16];
        s32 mine = kLzpBulkChunk;
#pragma unroll
        for (int j = 15; j >= 0; j--) {
            by[j] = (j0 + j < lim) ? in[ip + j0 + j] : (u8)0;
            if (j0 + j < lim && by[j] == kLzpEscape) mine = j0 + j;
        }
        s_last4[t] = (u32)by[12] | ((u32)by[13] << 8) | ((u32)by[14] << 16) | ((u32)by[15] << 24);
        s32 m = mine;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const s32 x = __shfl_xor_sync(kFullMask, m, o);
            m = x < m ? x : m;
        }
        if (lane == 0) s_first[t >> 5] = m;
        __syncthreads();
        if (t < 32) {
            s32 v = (t < kLzpBulkThreads / 32) ? s_first[t] : kLzpBulkChunk;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const s32 x = __shfl_xor_sync(kFullMask, v, o);
                v = x < v ? x : v;
            }
            if (t == 0) s_ctl[0] = v;
        }
        __syncthreads();
        const s32 first = s_ctl[0];
        const s32 nlit = first < lim ? first : lim;
        // ---- B: literals before the escape: copy + scatter-max of their positions
        {
            // context of my byte j = the four bytes before it: previous slice's last four, or the carried tail
            u32 prev;   // byte k of `prev` is the byte 4-k positions before my slice start... (oldest in the low byte)
            if (t == 0) prev = ((tail >> 24) & 0xFFu) | (((tail >> 16) & 0xFFu) << 8) | (((tail >> 8) & 0xFFu) << 16) | ((tail & 0xFFu) << 24);
            else prev = s_last4[t - 1];
            // sliding window: w holds the four bytes before the current byte, most recent in the low byte
            u32 w = ((prev & 0xFFu) << 24) | (((prev >> 8) & 0xFFu) << 16) | (((prev >> 16) & 0xFFu) << 8) | ((prev >> 24) & 0xFFu);
#pragma unroll
            for (int j = 0; j < 16; j++) {
                if (j0 + j < nlit) {
                    out[op + j0 + j] = by[j];
                    atomicMax(&lut[lzp_hash(w)], op + j0 + j);   // visited positions only grow
                }
                w = (w << 8) | (u32)by[j];
            }
            // new tail after nlit literals: owned by the thread holding byte nlit-1
            if (nlit > 0 && (nlit - 1) / 16 == t) {
                u32 nt = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const s32 idx = nlit - 1 - k;   // chunk index of the k-th most recent byte
                    u32 b;
                    if (idx >= j0) {
                        b = 0;
#pragma unroll
                        for (int j = 0; j < 16; j++) b = (j0 + j == idx) ? (u32)by[j] : b;
                    } else if (idx >= 0) {
                        b = (prev >> (8 * (idx - (j0 - 4)))) & 0xFFu;   // one of the four bytes before my slice
                    } else {
                        b = (tail >> (8 * (-idx - 1))) & 0xFFu;         // still the old tail
                    }
                    nt |= b << (8 * k);
                }
                s_tail = nt;
            }
        }
        __syncthreads();   // orders the global writes above for every thread of the block (reads below bypass L1)
        if (nlit > 0) tail = s_tail;
        op += nlit;
        ip += nlit;
        if (!(first < lim)) continue;   // no escape in this chunk (uniform)
        // ---- C: the byte at `ip` is 0xF2 (uniform: ip < n and op < max hold because first < lim)
        if (t == 0) {
            s32 kind = 0, count = 0, nip = ip;   // kind 0: literal 0xF2, 1: match, -1: truncated
            const s32 ref = atomicMax(&lut[lzp_hash(tail)], op);
            if (ref <= 0) {   // empty slot: 0xF2 is an ordinary literal (src/libbz3.c:212, :236)
                nip = ip + 1;
            } else {
                nip = ip + 1;
                if (nip == n) {
                    kind = -1;
                } else if (in[nip] == 255) {   // escaped literal
                    nip++;
                } else {
                    u32 ulen = kLzpMinMatch;  // wraps like the reference's signed 32-bit accumulator
                    bool truncated = false;
                    for (;;) {
                        if (nip == n) { truncated = true; break; }
                        const u32 c = in[nip++];
                        ulen += c;
                        if (c != 254) break;
                    }
                    if (truncated) {
                        kind = -1;
                    } else {
                        s64 stop64 = (s64)op + (s64)(s32)ulen;
                        if (stop64 > max) stop64 = max;
                        count = stop64 > op ? (s32)(stop64 - op) : 0;
                        kind = 1;
                    }
                }
            }
            if (kind == 0) out[op] = (u8)kLzpEscape;
            s_ctl[1] = kind;
            s_ctl[2] = count;
            s_ctl[3] = ref;
            s_ctl[4] = nip;
        }
        __syncthreads();
        const s32 kind = s_ctl[1], count = s_ctl[2], ref = s_ctl[3];
        ip = s_ctl[4];
        if (kind < 0) { status = -1; break; }
        if (kind == 0) {
            tail = (tail << 8) | (u32)kLzpEscape;
            op++;
        } else if (count > 0) {
            const s32 dist = op - ref;  // > 0; the source [ref, op) is final, the copy repeats it with period dist
            for (s32 k = t; k < count; k += kLzpBulkThreads) out[op + k] = __ldcg(out + ref + (dist >= count ? k : k % dist));
            __syncthreads();
            op += count;
            tail = (u32)__ldcg(out + op - 1) | ((u32)__ldcg(out + op - 2) << 8) | ((u32)__ldcg(out + op - 3) << 16) |
                   ((u32)__ldcg(out + op - 4) << 24);
        }
        __syncthreads();   // s_ctl / s_tail are reused by the next round
    }
    if (t == 0) *result = status < 0 ? -1 : op;
}

}  // namespace bz3
