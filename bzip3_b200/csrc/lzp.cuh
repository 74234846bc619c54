// lzp.cuh -- the LZP (hashed order-4 Lempel-Ziv prediction) pre-pass.
//
// Restates lzp_encode_block / lzp_decode_block (reference src/libbz3.c:124-257).  The format is
// defined by the exact sequential semantics of a 2^18-slot table that remembers the last VISITED
// position of every context hash; positions covered by a match are never inserted.  Constants are
// format-defining: min match 40, escape byte 0xF2, length code (len-40) as 254-runs.
//
// This header holds the single-lane form of both directions (host+device callable, so the same
// code is unit-tested on the CPU build in tests/); lzp_parallel.cuh holds the windowed CTA form.
#pragma once
#include "common.cuh"

namespace bz3 {

BZ_HD u32 lzp_hash(u32 ctx) { return ((ctx >> 15) ^ ctx ^ (ctx >> 3)) & (u32)(kLzpSlots - 1); }

BZ_HD u32 lzp_load32(const u8* p) {  // unaligned native-endian (little) 32-bit read
    return (u32)p[0] | ((u32)p[1] << 8) | ((u32)p[2] << 16) | ((u32)p[3] << 24);
}
// context word: the four bytes before `pos`, most recent byte in the low 8 bits
BZ_HD u32 lzp_context(const u8* base, s32 pos) {
    return (u32)base[pos - 1] | ((u32)base[pos - 2] << 8) | ((u32)base[pos - 3] << 16) | ((u32)base[pos - 4] << 24);
}

// Length of the match between in+ip and in+ref as the encoder measures it: 32-bit steps while
// ip+len < scan_end.  Starts from len = 4 (the first word was already compared).
BZ_HD s32 lzp_match_words(const u8* in, s32 ip, s32 ref, s32 scan_end) {
    s32 len = 4;
    while (ip + len < scan_end && lzp_load32(in + ip + len) == lzp_load32(in + ref + len)) len += 4;
    return len;
}

// Single-lane encoder.  lut must be zero-filled (kLzpSlots entries).  Returns the encoded size or
// -1 (input shorter than 72 bytes, or output not at least 8 bytes smaller than the input).
BZ_HD s32 lzp_encode_serial(const u8* in, s32 n, u8* out, s32* lut) {
    if (n < kLzpMinMatch + 32) return -1;
    const s32 out_stop = n - 8;
    const s32 scan_end = n - kLzpMinMatch - 32;
    s32 ip = 4, op = 4, veto_until = 0;
    out[0] = in[0]; out[1] = in[1]; out[2] = in[2]; out[3] = in[3];
    u32 ctx = lzp_context(in, ip);
    while (ip < scan_end && op < out_stop) {
        const u32 slot = lzp_hash(ctx);
        const s32 ref = lut[slot];
        lut[slot] = ip;
        s32 mlen = 0;
        if (ref > 0) {
            if (lzp_load32(in + ip + kLzpMinMatch - 4) == lzp_load32(in + ref + kLzpMinMatch - 4) &&
                lzp_load32(in + ip) == lzp_load32(in + ref)) {
                bool veto = veto_until > ip && lzp_load32(in + veto_until) != lzp_load32(in + ref + (veto_until - ip));
                if (!veto) {
                    s32 len = lzp_match_words(in, ip, ref, scan_end);
                    if (len >= kLzpMinMatch) {
                        len += in[ip + len] == in[ref + len];
                        len += in[ip + len] == in[ref + len];
                        len += in[ip + len] == in[ref + len];
                        mlen = len;
                    } else if (veto_until < ip + len) {
                        veto_until = ip + len;
                    }
                }
            }
        }
        if (mlen) {
            ip += mlen;
            ctx = lzp_context(in, ip);
            out[op++] = (u8)kLzpEscape;
            s32 code = mlen - kLzpMinMatch;
            while (code >= 254) {
                code -= 254;
                out[op++] = 254;
                if (op >= out_stop) break;
            }
            out[op++] = (u8)code;
        } else {
            const u8 b = in[ip++];
            out[op++] = b;
            ctx = (ctx << 8) | b;
            if (ref > 0 && b == kLzpEscape) out[op++] = 255;
        }
    }
    ctx = lzp_context(in, ip);
    while (ip < n && op < out_stop) {
        const u32 slot = lzp_hash(ctx);
        const s32 ref = lut[slot];
        lut[slot] = ip;
        const u8 b = in[ip++];
        out[op++] = b;
        ctx = (ctx << 8) | b;
        if (ref > 0 && b == kLzpEscape) out[op++] = 255;
    }
    return op >= out_stop ? -1 : op;
}

// Single-lane decoder.  lut zero-filled.  Returns decoded size, or -1 for a truncated token.
BZ_HD s32 lzp_decode_serial(const u8* in, s32 n, u8* out, s32 max, s32* lut) {
    if (n < 4) return -1;
    s32 ip = 4, op = 4;
    out[0] = in[0]; out[1] = in[1]; out[2] = in[2]; out[3] = in[3];
    u32 ctx = lzp_context(out, op);
    while (ip < n && op < max) {
        const u32 slot = lzp_hash(ctx);
        const s32 ref = lut[slot];
        lut[slot] = op;
        const u8 b = in[ip];
        if (b != kLzpEscape || ref <= 0) {
            ip++;
            out[op++] = b;
            ctx = (ctx << 8) | b;
            continue;
        }
        if (++ip == n) return -1;
        if (in[ip] == 255) {
            ip++;
            out[op++] = (u8)kLzpEscape;
            ctx = (ctx << 8) | (u32)kLzpEscape;
            continue;
        }
        u32 ulen = kLzpMinMatch;  // the reference accumulates in a signed 32-bit int; mirror the wrap
        for (;;) {
            if (ip == n) return -1;
            const u8 c = in[ip++];
            ulen += c;
            if (c != 254) break;
        }
        s64 stop = (s64)op + (s64)(s32)ulen;
        if (stop > max) stop = max;
        s32 src = ref;
        while (op < stop) out[op++] = out[src++];
        ctx = lzp_context(out, op);
    }
    return op;
}


}  // namespace bz3
