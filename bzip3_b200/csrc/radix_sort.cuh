// radix_sort.cuh -- stable LSD radix sort passes for (key, value) records, 8-bit digits.
//
// This is the engine behind the suffix sorter (64-bit keys = (rank[i], rank[i+h]), 32-bit values =
// suffix index) and the psi construction of the inverse BWT (8-bit keys, generated values).
//
// One sort = one histogram launch + ONE launch per pass ("one sweep"):
//   hist_all    : the digit histograms of EVERY pass in one read of the keys (they do not depend on the order of the
//                 records), then their exclusive scans over the digits -> gbase[pass][digit]
//   pass        : every CTA takes the next tile (atomic ticket), brings it into shared memory with one bulk async copy
//                 per array (TMA: cp.async.bulk + mbarrier), ranks the records stably inside the tile (warp match +
//                 per-warp counters), publishes its per-digit counts and finds the counts of the tiles before it by
//                 decoupled look-back (descriptor word = 2 flag bits + 30-bit count: "tile total" first, "inclusive
//                 prefix" once known), stages the tile in shared memory in sorted order and writes each digit's records
//                 as one contiguous, coalesced run.
// Algorithmic HBM bytes per pass over m records: (sizeof(K)+4)*m read + (sizeof(K)+4)*m written -- 24 B per record per
// pass for the 64-bit-key sort of the suffix sorter -- plus sizeof(K)*m once per sort for the histograms and 1 KiB of
// descriptors per tile of 3072 records.
#pragma once
#include "common.cuh"
#include "scan.cuh"

namespace bz3 {

constexpr int kRsThreads = 256;
constexpr int kRsWarps = kRsThreads / 32;

template <typename K>
struct RsCfg;
#ifndef BZ_RS_ITEMS_U64
#define BZ_RS_ITEMS_U64 12
#endif
#ifndef BZ_RS_MIN_BLOCKS
#define BZ_RS_MIN_BLOCKS 3
#endif
template <>
struct RsCfg<u64> {
    static constexpr int kItems = BZ_RS_ITEMS_U64;
};
template <>
struct RsCfg<u32> {
    static constexpr int kItems = 18;
};
template <>
struct RsCfg<u8> {
    static constexpr int kItems = 32;
};

template <typename K>
BZ_D u32 rs_digit(K k, int shift, u32 mask) {
    return (u32)(k >> shift) & mask;
}

// value generators for sorts whose payload is implicit
struct ValFromArray {
    const u32* v;
    BZ_D u32 operator()(u32 i) const { return v[i]; }
};
struct ValIdentity {
    BZ_D u32 operator()(u32 i) const { return i; }
};

constexpr u32 kRsFlagAgg = 1u << 30;   // descriptor holds the tile's own count of the digit
constexpr u32 kRsFlagInc = 2u << 30;   // descriptor holds the inclusive prefix over tiles 0..t
constexpr u32 kRsValMask = (1u << 30) - 1u;   // n < 2^30 (blocks are at most 511 MiB + 2 %)
constexpr int kRsMaxPasses = 8;

// digit histograms of all passes in one read of the keys: ghist[pass][256] (zeroed by the caller).  Every warp keeps
// private counters for all passes in shared memory across all the chunks its block walks (grid-stride); one reduction
// and one round of global atomics per block at the end.
constexpr int kRsHistItems = 8;
inline size_t rs_hist_smem(int npass) { return (size_t)kRsWarps * npass * 256 * sizeof(u32); }

template <typename K>
__global__ void __launch_bounds__(kRsThreads) rs_hist_all_kernel(const K* __restrict__ keys, u32 n, int nbits, u32* __restrict__ ghist) {
    constexpr int ITEMS = kRsHistItems;
    BZ_DYN_SMEM(u32, cnt);   // [kRsWarps][npass][256]
    const u32 w = warp_id(), l = lane_id();
    const int npass = (nbits + 7) / 8;
    for (int i = threadIdx.x; i < kRsWarps * npass * 256; i += kRsThreads) cnt[i] = 0;
    __syncthreads();
    u32* mine = cnt + (size_t)w * npass * 256;
    for (u32 tile = blockIdx.x; (u64)tile * (kRsThreads * ITEMS) < n; tile += gridDim.x) {
        const u32 base = tile * (kRsThreads * ITEMS) + w * (32 * ITEMS);
        K key[ITEMS];
#pragma unroll
        for (int it = 0; it < ITEMS; it++) {
            const u32 i = base + it * 32 + l;
            key[it] = (i < n) ? keys[i] : (K)0;
        }
        for (int p = 0; p < npass; p++) {
            const int shift = 8 * p;
            const u32 mask = (nbits - shift < 8) ? ((1u << (nbits - shift)) - 1u) : 255u;
#pragma unroll
            for (int it = 0; it < ITEMS; it++) {
                const u32 i = base + it * 32 + l;
                if (i < n) atomicAdd(&mine[p * 256 + rs_digit(key[it], shift, mask)], 1u);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < npass * 256; i += kRsThreads) {
        u32 sum = 0;
#pragma unroll
        for (int k = 0; k < kRsWarps; k++) sum += cnt[(size_t)k * npass * 256 + i];
        if (sum) atomicAdd(&ghist[i], sum);
    }
}

// ghist[pass][256] -> exclusive scan over the digits of each pass, in place (one warp per pass)
__global__ void __launch_bounds__(32 * kRsMaxPasses) rs_digit_scan_kernel(u32* __restrict__ ghist, int npass) {
    const u32 p = warp_id(), l = lane_id();
    if ((int)p >= npass) return;
    u32* row = ghist + p * 256;
    u32 carry = 0;
    for (int base = 0; base < 256; base += 32) {
        const u32 v = row[base + l];
        const u32 incl = warp_scan_incl(v);
        row[base + l] = carry + incl - v;
        carry += __shfl_sync(kFullMask, incl, 31);
    }
}

// ---- bulk async copy (TMA) of one contiguous run of global memory into shared memory, completion on an mbarrier
BZ_D void rs_mbar_init(u64* mbar) {
#if !defined(BZ_EMU)
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((u32)__cvta_generic_to_shared(mbar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
}
BZ_D void rs_bulk_load(void* smem_dst, const void* gmem_src, u32 bytes, u64* mbar) {   // bytes: multiple of 16, both 16-byte aligned
#if !defined(BZ_EMU)
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((u32)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src), "r"(bytes), "r"((u32)__cvta_generic_to_shared(mbar))
                 : "memory");
#endif
}
BZ_D void rs_mbar_expect(u64* mbar, u32 bytes) {
#if !defined(BZ_EMU)
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((u32)__cvta_generic_to_shared(mbar)), "r"(bytes) : "memory");
#endif
}
BZ_D void rs_mbar_wait(u64* mbar, u32 parity) {
#if !defined(BZ_EMU)
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "RS_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra RS_DONE;\n\t"
        "bra RS_WAIT;\n\t"
        "RS_DONE:\n\t"
        "}" ::"r"((u32)__cvta_generic_to_shared(mbar)), "r"(parity) : "memory");
#endif
}

BZ_D u32 rs_ld_desc(const u32* p) {
#if defined(BZ_EMU)
    return *reinterpret_cast<const volatile u32*>(p);
#else
    u32 v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
#endif
}
BZ_D void rs_st_desc(u32* p, u32 v) {
#if defined(BZ_EMU)
    *reinterpret_cast<volatile u32*>(p) = v;
#else
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#endif
}

// One stable pass on digit bits [shift, shift+bits).  KOUT: write keys;  VOUT: write values.  ValGen produces the value of
// input record i; VTMA: the values are an array (ValFromArray) and come in by bulk copy like the keys.
// desc[tile][256] (zeroed), ticket (zeroed), gbase[256] = first output slot of every digit.
template <typename K, bool KOUT, bool VOUT, bool VTMA, bool TMA, typename ValGen>
__global__ void __launch_bounds__(kRsThreads, BZ_RS_MIN_BLOCKS)
rs_onesweep_kernel(const K* __restrict__ kin, ValGen vgen, const u32* __restrict__ vin, K* __restrict__ kout, u32* __restrict__ vout, u32 n,
                   int shift, u32 mask, const u32* __restrict__ gbase, u32* __restrict__ desc, u32* __restrict__ ticket) {
    constexpr int ITEMS = RsCfg<K>::kItems;
    constexpr u32 TILE = kRsThreads * ITEMS;
    BZ_DYN_SMEM(unsigned char, rs_smem);
    u32* warp_cnt = reinterpret_cast<u32*>(rs_smem);              // [kRsWarps][256]
    u32* lbase = warp_cnt + kRsWarps * 256;                       // [256] first sorted slot of digit in tile
    u32* gdelta = lbase + 256;                                    // [256] global index minus tile slot
    u32* svals = gdelta + 256;                                    // [TILE]  (bulk copy target, then sorted staging)
    K* skeys = reinterpret_cast<K*>(svals + TILE);                // [TILE]  (16-byte aligned: TILE * 4 is a multiple of 16)
    __shared__ u32 s_tile;
    __shared__ u64 s_mbar;   // 8-byte aligned by its type
    const u32 w = warp_id(), l = lane_id();

    if (threadIdx.x == 0) {
        s_tile = atomicAdd(ticket, 1u);   // tiles start in ticket order: every tile before this one is running or done
        rs_mbar_init(&s_mbar);
    }
    for (int i = threadIdx.x; i < kRsWarps * 256; i += kRsThreads) warp_cnt[i] = 0;
    __syncthreads();
    const u32 tile = s_tile;
    const u32 tile_base = tile * TILE;
    const u32 count = min(TILE, n - tile_base);

    K key[ITEMS];
    u32 val[VOUT ? ITEMS : 1];
    u16 rank[ITEMS];
    const u32 base = tile_base + w * (32 * ITEMS);
    const u32 lt = lanemask_lt();
#if !defined(BZ_EMU)
    if (TMA) {   // the tile comes in by bulk async copies (the arrays are padded to 256 bytes, so rounding the last tile up to 16 is safe)
        if (threadIdx.x == 0) {
            const u32 kb = (count * (u32)sizeof(K) + 15u) & ~15u;
            const u32 vb = (VOUT && VTMA) ? ((count * 4u + 15u) & ~15u) : 0u;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            rs_mbar_expect(&s_mbar, kb + vb);
            rs_bulk_load(skeys, kin + tile_base, kb, &s_mbar);
            if (VOUT && VTMA) rs_bulk_load(svals, vin + tile_base, vb, &s_mbar);
        }
        rs_mbar_wait(&s_mbar, 0u);
#pragma unroll
        for (int it = 0; it < ITEMS; it++) {
            const u32 j = w * (32 * ITEMS) + it * 32 + l;
            key[it] = (j < count) ? skeys[j] : (K)0;
            if (VOUT) val[it] = (j < count) ? (VTMA ? svals[j] : vgen(tile_base + j)) : 0u;
        }
    } else
#endif
    {
#pragma unroll
        for (int it = 0; it < ITEMS; it++) {
            u32 i = base + it * 32 + l;
            key[it] = (i < n) ? kin[i] : (K)0;
            if (VOUT) val[it] = (i < n) ? vgen(i) : 0u;
        }
    }
#pragma unroll
    for (int it = 0; it < ITEMS; it++) {
        u32 i = base + it * 32 + l;
        bool valid = i < n;
        // records past n get digit 255; they are last in tile order, so their stable rank puts them
        // behind every real record and the write-out loop (j < count) never emits them
        u32 d = valid ? rs_digit(key[it], shift, mask) : 255u;
        u32 peers = __match_any_sync(kFullMask, d);
        u32 leader = __ffs(peers) - 1;
        u32 before = 0;
        if (l == leader) {
            before = warp_cnt[w * 256 + d];
            warp_cnt[w * 256 + d] = before + __popc(peers);
        }
        before = __shfl_sync(kFullMask, before, leader);
        rank[it] = (u16)(before + __popc(peers & lt));
        __syncwarp();
    }
    __syncthreads();
    // per digit: exclusive scan over warps, the tile's count -> descriptor ("tile total"), exclusive scan over digits
    u32 my_real;
    {
        const u32 d = threadIdx.x;
        u32 s = 0;
#pragma unroll
        for (int k = 0; k < kRsWarps; k++) {
            u32 t = warp_cnt[k * 256 + d];
            warp_cnt[k * 256 + d] = s;
            s += t;
        }
        // digit 255 also holds the padding records of a partial tile; exclude them from the real count
        u32 real = s;
        if (d == 255) real -= (TILE - count);
        my_real = real;
        rs_st_desc(desc + (size_t)tile * 256 + d, real | (tile == 0 ? kRsFlagInc : kRsFlagAgg));
        u32 incl = warp_scan_incl(real);
        __shared__ u32 wsum[kRsWarps];
        if (l == 31) wsum[w] = incl;
        __syncthreads();
        u32 wp = 0;
#pragma unroll
        for (int k = 0; k < kRsWarps; k++)
            if ((u32)k < w) wp += wsum[k];
        lbase[d] = wp + incl - real;
    }
    __syncthreads();
    // the tile goes to shared memory in sorted order (local offsets only) ...
#pragma unroll
    for (int it = 0; it < ITEMS; it++) {
        u32 i = base + it * 32 + l;
        u32 d = (i < n) ? rs_digit(key[it], shift, mask) : 255u;
        u32 pos = lbase[d] + warp_cnt[w * 256 + d] + rank[it];
        if (pos < TILE) {
            if (KOUT) skeys[pos] = key[it];
            if (VOUT) svals[pos] = val[it];
            if (!KOUT) reinterpret_cast<u8*>(skeys)[pos] = (u8)d;  // digit is still needed for the write-out
        }
    }
    // ... while the tiles before this one finish counting: decoupled look-back, eight predecessors per round trip (their
    // descriptors are requested together, then consumed nearest first; a descriptor that is not there yet is polled; an
    // inclusive prefix ends the walk)
    {
        const u32 d = threadIdx.x;
        u32* mine = desc + (size_t)tile * 256 + d;
        u32 before = 0;   // records with this digit in the tiles before this one
        if (tile > 0) {
            constexpr int LB = 8;
            s32 t = (s32)tile - 1;
            bool done = false;
            while (!done) {
                u32 v[LB];
#pragma unroll
                for (int k = 0; k < LB; k++) v[k] = (t - k >= 0) ? rs_ld_desc(desc + (size_t)(t - k) * 256 + d) : kRsFlagInc;
#pragma unroll
                for (int k = 0; k < LB; k++) {
                    if (!done) {
                        u32 x = v[k];
                        while ((x >> 30) == 0u) {
                            BZ_SPIN_HINT();
                            x = rs_ld_desc(desc + (size_t)(t - k) * 256 + d);
                        }
                        before += x & kRsValMask;
                        done = (x & kRsFlagInc) != 0u;
                    }
                }
                t -= LB;
            }
            rs_st_desc(mine, (before + my_real) | kRsFlagInc);
        }
        gdelta[d] = gbase[d] + before - lbase[d];
    }
    __syncthreads();
    for (u32 j = threadIdx.x; j < count; j += kRsThreads) {
        u32 d;
        if (KOUT) {
            K k = skeys[j];
            d = rs_digit(k, shift, mask);
            kout[gdelta[d] + j] = k;
        } else {
            d = reinterpret_cast<u8*>(skeys)[j];
        }
        if (VOUT) vout[gdelta[d] + j] = svals[j];
    }
}

template <typename K>
inline u32 rs_num_tiles(u32 n) {
    constexpr u32 TILE = kRsThreads * RsCfg<K>::kItems;
    return (n + TILE - 1) / TILE;
}
template <typename K>
inline size_t rs_scatter_smem() {
    constexpr u32 TILE = kRsThreads * RsCfg<K>::kItems;
    return (size_t)(kRsWarps * 256 + 512) * 4 + (size_t)TILE * 4 + (size_t)TILE * sizeof(K);
}
// scratch (in u32 elements) for a sort of n records: descriptors of one pass, histograms of all passes, ticket
template <typename K>
inline size_t rs_temp_elems(u32 n) {
    return (size_t)256 * rs_num_tiles<K>(n) + 256 * kRsMaxPasses + 64;
}

// temp layout
template <typename K>
struct RsTemp {
    u32* desc;
    u32* ghist;
    u32* ticket;
    RsTemp(u32* temp, u32 n) : desc(temp), ghist(temp + (size_t)256 * rs_num_tiles<K>(n)), ticket(ghist + 256 * kRsMaxPasses) {}
};

// histograms of all passes of a sort on key bits [0, nbits) -> gbase[pass][digit] in T.ghist
template <typename K>
cudaError_t rs_histograms(cudaStream_t st, const K* keys, u32 n, int nbits, const RsTemp<K>& T) {
    const int npass = (nbits + 7) / 8;
    BZ_CUDA_TRY(cudaMemsetAsync(T.ghist, 0, sizeof(u32) * 256 * kRsMaxPasses, st));
    const u32 chunks = (n + kRsThreads * kRsHistItems - 1) / (kRsThreads * kRsHistItems);
    const u32 grid = chunks < 148u * 3u ? chunks : 148u * 3u;   // persistent: 3 blocks per SM (64 KiB of counters each)
    const size_t smem = rs_hist_smem(npass);
    // The limit is a property of the function on the device, shared by every host thread that codes a block on it: always
    // the SAME value (the largest launch), never this launch's own size -- two threads with different pass counts would
    // otherwise lower it under each other's launches ("too many resources requested for launch").
    BZ_CUDA_TRY(cudaFuncSetAttribute(rs_hist_all_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)rs_hist_smem(kRsMaxPasses)));
    BZ_LAUNCH(grid, kRsThreads, smem, st, rs_hist_all_kernel<K>)(keys, n, nbits, T.ghist); BZ_NOTE_LAUNCH();
    BZ_LAUNCH(1, 32 * kRsMaxPasses, 0, st, rs_digit_scan_kernel)(T.ghist, npass); BZ_NOTE_LAUNCH();
    BZ_CUDA_TRY(cudaGetLastError());
    return cudaSuccess;
}

// One stable pass on digit bits [shift, shift+bits); T.ghist + 256 * pass_index must hold the digit bases of this pass.
template <typename K, bool KOUT, bool VOUT, bool VTMA, typename ValGen>
cudaError_t rs_pass_sweep(cudaStream_t st, const K* kin, ValGen vgen, const u32* vin, K* kout, u32* vout, u32 n, int shift, int bits,
                          int pass_index, const RsTemp<K>& T) {
    const u32 ntiles = rs_num_tiles<K>(n);
    const u32 mask = (1u << bits) - 1u;
    BZ_CUDA_TRY(cudaMemsetAsync(T.desc, 0, sizeof(u32) * 256 * (size_t)ntiles, st));
    BZ_CUDA_TRY(cudaMemsetAsync(T.ticket, 0, sizeof(u32), st));
    static const bool use_tma = !(getenv("BZ3_B200_RS_TMA") && getenv("BZ3_B200_RS_TMA")[0] == '0');   // A/B switch (tuning)
    auto kern = use_tma ? rs_onesweep_kernel<K, KOUT, VOUT, VTMA, true, ValGen> : rs_onesweep_kernel<K, KOUT, VOUT, VTMA, false, ValGen>;
    const size_t smem = rs_scatter_smem<K>();
    // set every time (the attribute is per device and a process may drive several GPUs), always to the same value
    BZ_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    BZ_LAUNCH(ntiles, kRsThreads, smem, st, kern)(kin, vgen, vin, kout, vout, n, shift, mask, T.ghist + 256 * pass_index, T.desc, T.ticket);
    BZ_NOTE_LAUNCH();
    BZ_CUDA_TRY(cudaGetLastError());
    return cudaSuccess;
}

// A single stable pass (the psi construction of the inverse BWT): histogram + one sweep.
template <typename K, bool KOUT, bool VOUT, typename ValGen>
cudaError_t rs_pass(cudaStream_t st, const K* kin, ValGen vgen, K* kout, u32* vout, u32 n, int shift, int bits, u32* temp) {
    if (n == 0) return cudaSuccess;
    RsTemp<K> T(temp, n);
    BZ_CUDA_TRY(rs_histograms<K>(st, kin, n, shift + bits, T));
    return rs_pass_sweep<K, KOUT, VOUT, false, ValGen>(st, kin, vgen, nullptr, kout, vout, n, shift, bits, shift / 8, T);
}

// Full LSD sort of (key, value) records on key bits [0, nbits).  Ping-pongs between the A and B
// buffers; *result_in_b tells where the sorted records ended up.
// If vals_are_index the values of the input records are their positions 0..n-1 and `va` is not read.
template <typename K>
cudaError_t rs_sort_pairs(cudaStream_t st, K* ka, u32* va, K* kb, u32* vb, u32 n, int nbits, u32* temp,
                          bool* result_in_b, bool vals_are_index = false) {
    bool in_b = false;
    if (n == 0) { *result_in_b = false; return cudaSuccess; }
    RsTemp<K> T(temp, n);
    BZ_CUDA_TRY(rs_histograms<K>(st, ka, n, nbits, T));
    for (int shift = 0; shift < nbits; shift += 8) {
        int bits = nbits - shift < 8 ? nbits - shift : 8;
        K* src_k = in_b ? kb : ka;
        u32* src_v = in_b ? vb : va;
        K* dst_k = in_b ? ka : kb;
        u32* dst_v = in_b ? va : vb;
        if (shift == 0 && vals_are_index) {
            BZ_CUDA_TRY((rs_pass_sweep<K, true, true, false, ValIdentity>(st, src_k, ValIdentity{}, nullptr, dst_k, dst_v, n, shift, bits,
                                                                        shift / 8, T)));
        } else {
            BZ_CUDA_TRY((rs_pass_sweep<K, true, true, true, ValFromArray>(st, src_k, ValFromArray{src_v}, src_v, dst_k, dst_v, n, shift,
                                                                        bits, shift / 8, T)));
        }
        in_b = !in_b;
    }
    *result_in_b = in_b;
    return cudaSuccess;
}

}  // namespace bz3
