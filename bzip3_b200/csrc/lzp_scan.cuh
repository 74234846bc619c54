// lzp_scan.cuh -- the LZP encoder: grid-parallel hash-match scan + an in-order commit engine.
//
// Restates lzp_encode_block (reference src/libbz3.c:124-198).  The format is defined by the sequential semantics of a
// 2^18-slot table that remembers the last VISITED position of every context hash (positions covered by a match are
// never inserted), so the bytes written depend on the order of events.  What does not depend on it is, for every
// position i, the chain of EARLIER positions with the same context hash; the table entry the reference reads at i is
// the nearest element of that chain that was visited.  So the work is split:
//
//   grid-parallel (all SMs, streaming reads)
//     lzp_hash_keys_kernel   context hash of every position, 16 positions per thread from two 128-bit loads
//     radix sort             (hash, position), stable, 18 key bits = 3 passes of radix_sort.cuh
//     lzp_link_kernel        P[i] = previous position with the hash of i (0 = none): neighbours in sorted order
//     lzp_code_kernel        for every i: the reference's 8-byte quick check (:143-144) and its word-wise match length
//                            (:147-150), capped at 40 bytes, against P[i] -- i.e. against what the table would hold if
//                            nothing had been skipped; one byte per position: 0 = no candidate, k = 4k bytes match
//                            (10 = "at least LZP_MIN_MATCH")
//   in order (one thread block, 1024 positions per step)
//     lzp_commit_kernel      walks the input once and keeps the reference's table (last visited position per hash) for
//                            everything BEFORE the current literal run.  A step loads P / code / the byte of 1024
//                            consecutive positions.  A lane whose predecessor P[i] lies inside the run sees exactly P[i]
//                            in the reference's table (every position of the run was visited); any other lane reads the
//                            table, and only if that differs from P[i] (the predecessor was skipped by an earlier match)
//                            re-does quick check and length against it.  One warp then runs the reference's candidate
//                            logic (`heur` veto, :145, :152-155) over the lanes that passed the quick check, literals
//                            are emitted with a block-wide prefix sum and the visited positions enter the table
//                            (order-free: positions only grow, atomicMax).  An accepted match is measured to its full
//                            length by all threads, its token written, and the walk restarts behind it.
// Text without long repeats never leaves the streaming path: the commit engine then moves ~1 byte per clock.  Data that
// is one match after the other pays one step (a few dependent memory round trips) per match, like the reference pays a
// cache miss per match.
#pragma once
#include "common.cuh"
#include "lzp.cuh"
#include "radix_sort.cuh"

namespace bz3 {

#if defined(BZ_DEVICE_CODE)

constexpr int kLzpEngThreads = 1024;
constexpr int kLzpEngWarps = kLzpEngThreads / 32;

BZ_D u32 lzp_word(const u8* __restrict__ p) {
    return (u32)p[0] | ((u32)p[1] << 8) | ((u32)p[2] << 16) | ((u32)p[3] << 24);
}

// keys[j] = hash of the context of position j + 4 (the four bytes j .. j+3, oldest byte in the top 8 bits: :134, :163),
// j in [0, m), m = n - 4.  `in` is 16-byte aligned and readable up to n + 32.
__global__ void __launch_bounds__(256) lzp_hash_keys_kernel(const u8* __restrict__ in, u32 m, u32* __restrict__ keys) {
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 j0 = t * 16u;
    if (j0 >= m) return;
    const uint4 a = *reinterpret_cast<const uint4*>(in + j0);
    const uint4 b = *reinterpret_cast<const uint4*>(in + j0 + 16);
    const u32 w[5] = {a.x, a.y, a.z, a.w, b.x};   // little-endian words: byte j0 + 4k in the low 8 bits of w[k]
    u32 out[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const u32 lo = w[k >> 2], hi = w[(k >> 2) + 1];
        const u32 le = (k & 3) ? __funnelshift_r(lo, hi, 8 * (k & 3)) : lo;   // bytes j..j+3, byte j in the low 8 bits
        const u32 ctx = __byte_perm(le, 0, 0x0123);                           // byte j in the top 8 bits
        out[k] = lzp_hash(ctx);
    }
    if (j0 + 16 <= m) {
        uint4* o = reinterpret_cast<uint4*>(keys + j0);
        o[0] = make_uint4(out[0], out[1], out[2], out[3]);
        o[1] = make_uint4(out[4], out[5], out[6], out[7]);
        o[2] = make_uint4(out[8], out[9], out[10], out[11]);
        o[3] = make_uint4(out[12], out[13], out[14], out[15]);
    } else {
        for (u32 k = 0; k < 16 && j0 + k < m; k++) keys[j0 + k] = out[k];
    }
}

// sorted (key, index) records, stable: within one hash the indices ascend.  P[index + 4] = predecessor position.
__global__ void __launch_bounds__(256) lzp_link_kernel(const u32* __restrict__ skey, const u32* __restrict__ sidx, u32 m,
                                                       u32* __restrict__ P) {
    const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const u32 k = skey[j];
    const u32 prev = (j > 0 && skey[j - 1] == k) ? sidx[j - 1] + 4u : 0u;
    P[sidx[j] + 4u] = prev;
}

// the ten 32-bit words at p, p+4, .. p+36 (any alignment) from eleven aligned loads that are all in flight together
BZ_D void lzp_load10(const u8* __restrict__ in, s32 p, u32 (&w)[10]) {
    const u32* a = reinterpret_cast<const u32*>(in + (p & ~3));
    const int sh = 8 * (p & 3);
    u32 x[11];
#pragma unroll
    for (int k = 0; k < 11; k++) x[k] = a[k];
#pragma unroll
    for (int k = 0; k < 10; k++) w[k] = sh ? __funnelshift_r(x[k], x[k + 1], sh) : x[k];
}

// quick check and capped word-wise length of position p against reference position r (:143-150).  0 = the quick check
// fails; k in 1..10 = 4k bytes match (10: at least LZP_MIN_MATCH).  p < scan_end.  All loads are issued before the first
// compare (the word loop of the reference is a chain of dependent loads otherwise).
BZ_D u32 lzp_candidate_code(const u8* __restrict__ in, s32 p, s32 r, s32 scan_end) {
    if (lzp_word(in + p) != lzp_word(in + r)) return 0u;
    u32 a[10], b[10];
    lzp_load10(in, p, a);
    lzp_load10(in, r, b);
    if (a[9] != b[9]) return 0u;   // bytes 36..39
    s32 len = 4;
#pragma unroll
    for (int k = 1; k < 10; k++)
        if (len == 4 * k && p + len < scan_end && a[k] == b[k]) len += 4;
    return (u32)len >> 2;
}

__global__ void __launch_bounds__(256) lzp_code_kernel(const u8* __restrict__ in, const u32* __restrict__ P, s32 scan_end,
                                                       u8* __restrict__ code) {
    const s32 p = 4 + (s32)(blockIdx.x * blockDim.x + threadIdx.x);
    if (p >= scan_end) return;
    const u32 r = P[p];
    code[p] = r ? (u8)lzp_candidate_code(in, p, (s32)r, scan_end) : (u8)0;
}

// ---- the in-order engine ------------------------------------------------------------------------------------
struct LzpEngShared {
    u32 wmask[kLzpEngWarps];     // ballot of the lanes that passed the quick check, per warp
    u32 wsum[kLzpEngWarps];      // literal bytes per warp (prefix sum)
    s32 sval[kLzpEngThreads];    // table value seen by the lane's position
    u8 scode[kLzpEngThreads];    // candidate code of the lane
    s32 match_lane, match_len;
    s32 heur;
    u32 first_stop;
};

// in: n bytes (readable, zero or not, up to n + 32).  P, code: from the kernels above (code zero outside [4, scan_end)).
// lut: the 2^18-entry table, zeroed.
__global__ void __launch_bounds__(kLzpEngThreads, 1) lzp_commit_kernel(const u8* __restrict__ in, s32 n, const u32* __restrict__ P,
                                                                      const u8* __restrict__ code, s32* __restrict__ lut,
                                                                      u8* __restrict__ out, s32* __restrict__ result) {
    __shared__ LzpEngShared S;
    const int t = threadIdx.x;
    const u32 lane = lane_id(), warp = warp_id();
    const s32 out_stop = n - 8;
    const s32 scan_end = n - kLzpMinMatch - 32;
    if (t < 4) out[t] = in[t];
    if (t == 0) S.heur = 0;
    __syncthreads();
    s32 ip = 4, op = 4;
    s32 run_start = 4;   // first position behind the last match: everything in [run_start, ip) was visited
    s32 pf_ip = -1, pf_val = 0;   // prefetched window (start position; this lane's values)
    u32 pf_c = 0, pf_h = 0;
    u8 pf_b = 0;
    s32 wcap = kLzpEngThreads;   // positions per step: shrinks where one match follows the other (the lanes behind a match are
                                 // thrown away), doubles again while no match is found
    for (int phase = 0; phase < 2; phase++) {   // 0: positions that may start a match; 1: the literal-only tail (:187-195)
        const s32 limit = phase == 0 ? scan_end : n;
        while (ip < limit && op < out_stop) {
            const s32 W = (limit - ip) < wcap ? (limit - ip) : wcap;
            const bool active = t < W;
            // Behind a match the step is narrow (128 positions = 4 of the 32 warps).  The other warps only keep the barriers
            // company: measured with ncu on a source block, the block executed 34 warp instructions per input position, most
            // of them in warps without a single active lane (profiles/r02_call28_ncu_lzp_commit_source16.md).
            const bool wact = (s32)(warp * 32u) < W;
            const int nwact = (W + 31) >> 5;
            const s32 p = ip + t;
            s32 val = 0;
            u32 c = 0, h = 0;
            u8 b = 0;
            if (active) {
                if (pf_ip == ip) {   // this window was requested one step ago
                    b = pf_b;
                    val = pf_val;
                    c = pf_c;
                    h = pf_h;
                } else {
                    b = in[p];
                    val = (s32)P[p];
                    c = phase == 0 ? code[p] : 0u;
                    h = lzp_hash(lzp_context(in, p));
                }
            }
            {   // request the next full window now (the common case: no match in this one); it is consumed a step later
                const s32 np = ip + W + t;
                if (W == kLzpEngThreads && np < limit) {
                    pf_b = in[np];
                    pf_val = (s32)P[np];
                    pf_c = phase == 0 ? code[np] : 0u;
                    pf_h = lzp_hash(lzp_context(in, np));
                    pf_ip = ip + W;
                } else {
                    pf_ip = -1;
                }
            }
            if (active && val > 0 && val < run_start) {
                // predecessor from before the run: the table knows whether it was ever inserted
                const s32 tv = __ldcg(&lut[h]);
                if (tv != val) {   // it was skipped by a match: the table holds an older visited position (or none)
                    val = tv;
                    c = (phase == 0 && val > 0) ? lzp_candidate_code(in, p, val, scan_end) : 0u;
                }
            }
            u32 qb = 0;
            if (wact) {
                S.sval[t] = val;
                S.scode[t] = (u8)c;
                qb = __ballot_sync(kFullMask, c != 0u);
            }
            if (lane == 0) S.wmask[warp] = qb;
            if (t == 0) { S.match_lane = -1; S.match_len = 0; }
            __syncthreads();
            if (phase == 0 && warp == 0) {   // the reference's candidate logic, in order (:145-155); warp 0, all lanes alike
                s32 heur = S.heur;
                const u32 mine = S.wmask[lane];
                u32 nz = __ballot_sync(kFullMask, mine != 0u);
                s32 found = -1;
                while (nz && found < 0) {
                    const int w = __ffs(nz) - 1;
                    nz &= nz - 1;
                    u32 mask = __shfl_sync(kFullMask, mine, w);
                    while (mask) {
                        const int l = w * 32 + (__ffs(mask) - 1);
                        mask &= mask - 1;
                        const s32 pl = ip + l, rl = S.sval[l];
                        if (heur > pl && lzp_word(in + heur) != lzp_word(in + rl + (heur - pl))) continue;
                        const s32 len = 4 * (s32)S.scode[l];
                        if (len < kLzpMinMatch) {
                            if (heur < pl + len) heur = pl + len;
                            continue;
                        }
                        found = l;
                        break;
                    }
                }
                if (lane == 0) {
                    S.heur = heur;
                    S.match_lane = found;
                }
            }
            __syncthreads();
            const s32 match_lane = S.match_lane;
            // literals: the lanes below the match (or the whole window)
            const s32 nlit = match_lane >= 0 ? match_lane : W;
            const bool lit = t < nlit;
            const bool esc = lit && b == kLzpEscape && val > 0;   // :176-178, :181, :194
            const u32 cnt = lit ? (esc ? 2u : 1u) : 0u;
            u32 incl = 0;
            if (wact) {
                incl = warp_scan_incl(cnt);
                if (lane == 31) S.wsum[warp] = incl;
            }
            __syncthreads();
            u32 before = 0, total = 0;
            if (nwact == kLzpEngWarps) {   // full step (text without matches): unrolled
#pragma unroll
                for (int k = 0; k < kLzpEngWarps; k++) {
                    const u32 v = S.wsum[k];
                    if ((u32)k < warp) before += v;
                    total += v;
                }
            } else {
                for (int k = 0; k < nwact; k++) {   // uniform bound: the warps that hold positions of this step
                    const u32 v = S.wsum[k];
                    if ((u32)k < warp) before += v;
                    total += v;
                }
            }
            if (lit) {
                const s32 o = op + (s32)(before + incl - cnt);
                if (o < n) out[o] = b;               // past out_stop the result is -1 anyway: never write past the input size
                if (esc && o + 1 < n) out[o + 1] = 255;
            }
            op += (s32)total;
            // the visited positions (literals and the match position itself) enter the table; the last one of a hash wins
            if (active && t <= (match_lane >= 0 ? match_lane : W - 1)) atomicMax(&lut[h], p);
            if (match_lane >= 0) {
                const s32 m = ip + match_lane, r = S.sval[match_lane];
                // full length, word-wise while m + len < scan_end (:147-150); the first 40 bytes are known to match
                s32 len = kLzpMinMatch;
                for (;;) {
                    __syncthreads();
                    if (t == 0) S.first_stop = 0xFFFFFFFFu;
                    __syncthreads();
                    const s32 off = len + 4 * t;
                    bool stop = true;
                    if (m + off < scan_end) stop = lzp_word(in + m + off) != lzp_word(in + r + off);
                    const u32 sb = __ballot_sync(kFullMask, stop);
                    if (sb && lane == 0) atomicMin(&S.first_stop, warp * 32u + (u32)(__ffs(sb) - 1));
                    __syncthreads();
                    const u32 fs = S.first_stop;
                    if (fs != 0xFFFFFFFFu) { len += 4 * (s32)fs; break; }
                    len += 4 * kLzpEngThreads;
                }
                if (t == 0) {
                    // :157-159 -- three byte compares, each at the current len: together they add the number of equal leading
                    // bytes, at most 3.  One round trip to memory instead of three dependent ones.
                    const u32 x = lzp_word(in + m + len) ^ lzp_word(in + r + len);
                    const s32 e = x ? ((__ffs((int)x) - 1) >> 3) : 4;
                    len += e < 3 ? e : 3;
                    S.match_len = len;
                }
                __syncthreads();
                len = S.match_len;
                // token: 0xF2, (len - 40) as a run of 254s and a last byte (:163-173)
                const s32 codev = len - kLzpMinMatch;
                const s32 q = codev / 254;
                if (t == 0 && op < n) out[op] = (u8)kLzpEscape;
                for (s32 k = t; k < q; k += kLzpEngThreads)
                    if (op + 1 + k < n) out[op + 1 + k] = 254;
                if (t == 0 && op + 1 + q < n) out[op + 1 + q] = (u8)(codev - 254 * q);
                op += 2 + q;
                ip = m + len;
                run_start = ip;
                wcap = 128;
                while (wcap < 2 * (match_lane + 1) && wcap < kLzpEngThreads) wcap *= 2;
            } else {
                ip += W;
                if (wcap < kLzpEngThreads) wcap *= 2;
            }
            __syncthreads();
        }
    }
    if (t == 0) *result = op >= out_stop ? -1 : op;
}

#endif  // BZ_DEVICE_CODE

#if defined(__CUDACC__) || defined(BZ_EMU)
struct LzpScanBuffers {
    u32* key[2];   // [m] each, m = n - 4
    u32* idx[2];   // [m] each
    u32* P;        // [n + 8]
    u8* code;      // [n + 8]
    s32* lut;      // [2^18], zeroed by lzp_scan_encode
    u32* temp;     // rs_temp_elems<u32>(m)
};

// in: device, n bytes, 16-byte aligned, readable up to n + 32.  *result (device) = encoded size or -1.
inline cudaError_t lzp_scan_encode(cudaStream_t st, const u8* in, s32 n, u8* out, const LzpScanBuffers& B, s32* d_result) {
    const u32 m = (u32)n - 4u;
    const s32 scan_end = n - kLzpMinMatch - 32;
    BZ_LAUNCH((m / 16 + 256) / 256, 256, 0, st, lzp_hash_keys_kernel)(in, m, B.key[0]); BZ_NOTE_LAUNCH();
    BZ_CUDA_TRY(cudaGetLastError());
    bool in_b = false;
    BZ_CUDA_TRY(rs_sort_pairs<u32>(st, B.key[0], B.idx[0], B.key[1], B.idx[1], m, kLzpSlotsLog2, B.temp, &in_b, true));
    const int sc = in_b ? 1 : 0;
    BZ_CUDA_TRY(cudaMemsetAsync(B.P, 0, sizeof(u32) * 8, st));   // positions 0..3 have no context
    BZ_LAUNCH((m + 255) / 256, 256, 0, st, lzp_link_kernel)(B.key[sc], B.idx[sc], m, B.P); BZ_NOTE_LAUNCH();
    BZ_CUDA_TRY(cudaMemsetAsync(B.code, 0, (size_t)n + 8, st));
    BZ_CUDA_TRY(cudaMemsetAsync(B.lut, 0, sizeof(s32) * kLzpSlots, st));
    if (scan_end > 4) {
        BZ_LAUNCH((u32)(scan_end - 4 + 255) / 256, 256, 0, st, lzp_code_kernel)(in, B.P, scan_end, B.code); BZ_NOTE_LAUNCH();
    }
    BZ_CUDA_TRY(cudaGetLastError());
    BZ_LAUNCH(1, kLzpEngThreads, 0, st, lzp_commit_kernel)(in, n, B.P, B.code, B.lut, out, d_result); BZ_NOTE_LAUNCH();
    BZ_CUDA_TRY(cudaGetLastError());
    return cudaSuccess;
}
#endif

}  // namespace bz3
