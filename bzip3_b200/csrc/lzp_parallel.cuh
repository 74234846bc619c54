// lzp_parallel.cuh -- warp-windowed LZP encoder and decoder (one warp per block).
//
// Same format and same sequential semantics as the single-lane code in lzp.cuh (reference
// src/libbz3.c:124-257); the warp processes 32 consecutive positions per step and commits them in
// order, so the result is bit-identical:
//
// encode  every lane hashes the context of its position and probes the table as it was BEFORE the
//         window; a lane whose hash equals that of a lower lane takes that lane's position instead
//         (the lower lane would have overwritten the slot first).  This is exact as long as every
//         lower position is really visited, i.e. up to and including the first match that is taken;
//         the window restarts right behind that match.  Lanes whose 8-byte quick check passes are the
//         only ones that can become matches; they are examined in order with the `heur` veto state,
//         their length measured by the whole warp.  Literals are emitted with a warp prefix sum.
// decode  a window of 32 input bytes is copied as literals up to the first 0xF2 byte; contexts come
//         from register shuffles, table updates are atomicMax (positions only grow).  The 0xF2 byte is
//         then resolved against the table exactly like the reference (plain literal when the slot is
//         empty, escaped literal, or a match copied by the whole warp; overlapping copies are periodic).
#pragma once
#include "common.cuh"
#include "lzp.cuh"

namespace bz3 {

BZ_D u32 lzp_ld32(const u8* __restrict__ p) {
    return (u32)p[0] | ((u32)p[1] << 8) | ((u32)p[2] << 16) | ((u32)p[3] << 24);
}

// number of bytes (multiple of 4, starting at 4) that in+ip and in+ref share, measured in 32-bit words
// while ip+len < scan_end -- the loop of src/libbz3.c:147-150, executed 32 words at a time
BZ_D s32 lzp_warp_match_words(const u8* __restrict__ in, s32 ip, s32 ref, s32 scan_end, u32 lane) {
    s32 base = 4;
    for (;;) {
        const s32 off = base + 4 * (s32)lane;
        bool stop = true;
        if (ip + off < scan_end) stop = lzp_ld32(in + ip + off) != lzp_ld32(in + ref + off);
        const u32 m = __ballot_sync(kFullMask, stop);
        if (m) return base + 4 * (__ffs(m) - 1);
        base += 128;
    }
}

__global__ void __launch_bounds__(32) lzp_encode_warp_kernel(const u8* __restrict__ in, s32 n, u8* __restrict__ out,
                                                             s32* __restrict__ lut, s32* __restrict__ result) {
    const u32 lane = lane_id();
    const u32 lt = lanemask_lt();
    if (n < kLzpMinMatch + 32) {
        if (lane == 0) *result = -1;
        return;
    }
    const s32 out_stop = n - 8;
    const s32 scan_end = n - kLzpMinMatch - 32;
    if (lane < 4) out[lane] = in[lane];
    s32 ip = 4, op = 4, veto_until = 0;
    // phase 0: positions that may start a match; phase 1: the literal-only tail (src/libbz3.c:187-195)
    for (int phase = 0; phase < 2; phase++) {
        const s32 limit = phase == 0 ? scan_end : n;
        while (ip < limit && op < out_stop) {
            const s32 W = (limit - ip) < 32 ? (limit - ip) : 32;
            const bool active = (s32)lane < W;
            const s32 p = ip + (s32)lane;
            u32 h = 0xFFFFFFFFu - lane;  // inactive lanes never collide
            s32 ref = 0;
            u8 b = 0;
            if (active) {
                h = lzp_hash(lzp_context(in, p));
                ref = __ldcg(&lut[h]);  // L2: the table is rewritten every step, never trust L1
                b = in[p];
            }
            const u32 peers = __match_any_sync(kFullMask, h);
            const u32 lower = peers & lt;
            if (lower) ref = ip + (31 - __clz(lower));  // the nearest lower lane with the same hash wrote the slot last
            s32 match_lane = -1, mlen = 0;
            if (phase == 0) {
                bool qc = false;
                if (active && ref > 0)
                    qc = lzp_ld32(in + p + kLzpMinMatch - 4) == lzp_ld32(in + ref + kLzpMinMatch - 4) &&
                         lzp_ld32(in + p) == lzp_ld32(in + ref);
                u32 cand = __ballot_sync(kFullMask, qc);
                while (cand) {
                    const int l = __ffs(cand) - 1;
                    cand &= cand - 1;
                    const s32 pl = ip + l;
                    const s32 rl = __shfl_sync(kFullMask, ref, l);
                    if (veto_until > pl && lzp_ld32(in + veto_until) != lzp_ld32(in + rl + (veto_until - pl))) continue;
                    s32 len = lzp_warp_match_words(in, pl, rl, scan_end, lane);
                    if (len < kLzpMinMatch) {
                        if (veto_until < pl + len) veto_until = pl + len;
                        continue;
                    }
                    len += in[pl + len] == in[rl + len];
                    len += in[pl + len] == in[rl + len];
                    len += in[pl + len] == in[rl + len];
                    match_lane = l;
                    mlen = len;
                    break;
                }
            }
            // positions ip .. ip+nvis-1 are visited (the match start included): they own their table slot
            const s32 nvis = match_lane >= 0 ? match_lane + 1 : W;
            const bool visited = (s32)lane < nvis;
            const u32 vis_peers = peers & __ballot_sync(kFullMask, visited);
            if (visited && (31 - __clz(vis_peers)) == (int)lane) __stcg(&lut[h], p);  // last visited lane of a hash wins
            // literals: lanes below the match (or the whole window)
            const s32 nlit = match_lane >= 0 ? match_lane : W;
            const bool lit = (s32)lane < nlit;
            const bool esc = lit && b == kLzpEscape && ref > 0;
            const u32 cnt = lit ? (esc ? 2u : 1u) : 0u;
            const u32 incl = warp_scan_incl(cnt);
            if (lit) {
                u8* o = out + op + (incl - cnt);
                o[0] = b;
                if (esc) o[1] = 255;
            }
            op += (s32)__shfl_sync(kFullMask, incl, 31);
            if (match_lane >= 0) {
                if (lane == 0) {
                    s32 o = op;
                    out[o++] = (u8)kLzpEscape;
                    s32 code = mlen - kLzpMinMatch;
                    while (code >= 254) {
                        code -= 254;
                        out[o++] = 254;
                        if (o >= out_stop) break;
                    }
                    out[o++] = (u8)code;
                    op = o;
                }
                op = __shfl_sync(kFullMask, op, 0);
                ip += match_lane + mlen;
            } else {
                ip += W;
            }
            __syncwarp();
        }
    }
    if (lane == 0) *result = op >= out_stop ? -1 : op;
}

__global__ void __launch_bounds__(32) lzp_decode_warp_kernel(const u8* __restrict__ in, s32 n, u8* __restrict__ out,
                                                             s32 max, s32* __restrict__ lut, s32* __restrict__ result) {
    const u32 lane = lane_id();
    if (n < 4) {
        if (lane == 0) *result = -1;
        return;
    }
    if (lane < 4) out[lane] = in[lane];
    s32 ip = 4, op = 4;
    // the four most recent output bytes, most recent in the low byte (uniform across the warp)
    u32 tail = (u32)in[3] | ((u32)in[2] << 8) | ((u32)in[1] << 16) | ((u32)in[0] << 24);
    s32 status = 0;
    while (ip < n && op < max) {
        s32 W = n - ip;
        if (max - op < W) W = max - op;
        if (W > 32) W = 32;
        const bool active = (s32)lane < W;
        const u32 b = active ? (u32)in[ip + lane] : 0u;
        const u32 escmask = __ballot_sync(kFullMask, active && b == (u32)kLzpEscape);
        const s32 nlit = escmask ? (__ffs(escmask) - 1) : W;
        // context of lane l = output bytes op+l-4 .. op+l-1: lower lanes' bytes, or the carried tail
        const u32 b1 = __shfl_up_sync(kFullMask, b, 1), b2 = __shfl_up_sync(kFullMask, b, 2);
        const u32 b3 = __shfl_up_sync(kFullMask, b, 3), b4 = __shfl_up_sync(kFullMask, b, 4);
        u32 ctx;
        {
            // byte j-1-lane of `tail` is the output byte j positions before op+lane when lane < j
            const u32 c1 = lane >= 1 ? b1 : (tail >> (8 * ((0 - lane) & 3))) & 0xFF;
            const u32 c2 = lane >= 2 ? b2 : (tail >> (8 * ((1 - lane) & 3))) & 0xFF;
            const u32 c3 = lane >= 3 ? b3 : (tail >> (8 * ((2 - lane) & 3))) & 0xFF;
            const u32 c4 = lane >= 4 ? b4 : (tail >> (8 * ((3 - lane) & 3))) & 0xFF;
            ctx = c1 | (c2 << 8) | (c3 << 16) | (c4 << 24);
        }
        if ((s32)lane < nlit) {
            out[op + lane] = (u8)b;
            atomicMax(&lut[lzp_hash(ctx)], op + (s32)lane);  // visited positions only grow
        }
        // new tail after nlit literals (uniform): bytes at op+nlit-1 .. op+nlit-4
        if (nlit > 0) {
            u32 nt = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int src = nlit - 1 - k;  // lane holding the k-th most recent byte, or the old tail
                const u32 v = __shfl_sync(kFullMask, b, src < 0 ? 0 : src);
                const u32 byte = src >= 0 ? v : (tail >> (8 * ((-src - 1) & 3))) & 0xFF;
                nt |= byte << (8 * k);
            }
            tail = nt;
        }
        op += nlit;
        ip += nlit;
        __syncwarp();
        if (!(escmask && nlit < W)) continue;
        // ---- the byte at `ip` is 0xF2: resolve it against the table (all lanes run this uniformly)
        s32 ref = 0;
        if (lane == 0) {
            // earlier atomicMax updates of this warp must be visible: same-thread ordering does not cover
            // other lanes, so go through the L2 with an atomic exchange-like max
            ref = atomicMax(&lut[lzp_hash(tail)], op);
        }
        ref = __shfl_sync(kFullMask, ref, 0);
        if (ref <= 0) {  // empty slot: 0xF2 is an ordinary literal (src/libbz3.c:212, :236)
            if (lane == 0) out[op] = (u8)kLzpEscape;
            tail = (tail << 8) | (u32)kLzpEscape;
            op++;
            ip++;
            continue;
        }
        ip++;
        if (ip == n) { status = -1; break; }
        u32 c = in[ip];
        if (c == 255) {  // escaped literal
            ip++;
            if (lane == 0) out[op] = (u8)kLzpEscape;
            tail = (tail << 8) | (u32)kLzpEscape;
            op++;
            continue;
        }
        u32 ulen = kLzpMinMatch;  // wraps like the reference's signed 32-bit accumulator
        bool truncated = false;
        for (;;) {
            if (ip == n) { truncated = true; break; }
            c = in[ip++];
            ulen += c;
            if (c != 254) break;
        }
        if (truncated) { status = -1; break; }
        s64 stop64 = (s64)op + (s64)(s32)ulen;
        if (stop64 > max) stop64 = max;
        const s32 count = stop64 > op ? (s32)(stop64 - op) : 0;
        if (count > 0) {
            __threadfence_block();
            const s32 dist = op - ref;  // > 0; the source [ref, op) is final, the copy repeats it with period dist
            for (s32 k = lane; k < count; k += 32) out[op + k] = __ldcg(out + ref + (dist >= count ? k : k % dist));
            __threadfence_block();
            __syncwarp();
            op += count;
            tail = (u32)__ldcg(out + op - 1) | ((u32)__ldcg(out + op - 2) << 8) | ((u32)__ldcg(out + op - 3) << 16) |
                   ((u32)__ldcg(out + op - 4) << 24);
        }
    }
    if (lane == 0) *result = status < 0 ? -1 : op;
}

}  // namespace bz3
