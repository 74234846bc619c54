// radix_sort.cuh -- stable LSD radix sort passes for (key, value) records, 8-bit digits.
//
// This is the engine behind the suffix sorter (64-bit keys = (rank[i], rank[i+h]), 32-bit values =
// suffix index) and the psi construction of the inverse BWT (8-bit keys, generated values).
//
// One pass = three launches and no inter-CTA spinning:
//   tile_hist   : every CTA counts the digits of its tile        -> hist[digit][tile]   (reads keys)
//   device_scan : exclusive scan of hist in digit-major order    -> global base of each (digit, tile)
//   scatter     : every CTA re-reads its tile, ranks the records stably inside the tile
//                 (warp match + per-warp counters), stages the tile in shared memory in sorted order
//                 and writes each digit's records as one contiguous, coalesced run.
// Algorithmic HBM bytes per pass over m records: sizeof(K)*m (hist) + (sizeof(K)+4)*m read
// + (sizeof(K)+4)*m written; for the 64-bit-key sort that is 32 B per record per pass.
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

template <typename K>
__global__ void __launch_bounds__(kRsThreads)
rs_tile_hist_kernel(const K* __restrict__ keys, u32 n, int shift, u32 mask, u32* __restrict__ hist, u32 ntiles) {
    constexpr int ITEMS = RsCfg<K>::kItems;
    constexpr u32 TILE = kRsThreads * ITEMS;
    __shared__ u32 cnt[kRsWarps][256];
    for (int i = threadIdx.x; i < kRsWarps * 256; i += kRsThreads) (&cnt[0][0])[i] = 0;
    __syncthreads();
    const u32 w = warp_id(), l = lane_id();
    const u32 base = blockIdx.x * TILE + w * (32 * ITEMS);
#pragma unroll
    for (int it = 0; it < ITEMS; it++) {
        u32 i = base + it * 32 + l;
        if (i < n) atomicAdd(&cnt[w][rs_digit(keys[i], shift, mask)], 1u);
    }
    __syncthreads();
    {
        u32 d = threadIdx.x, s = 0;
#pragma unroll
        for (int k = 0; k < kRsWarps; k++) s += cnt[k][d];
        hist[d * ntiles + blockIdx.x] = s;
    }
}

// KOUT: write keys;  VOUT: write values.  ValGen produces the value of input record i.
template <typename K, bool KOUT, bool VOUT, typename ValGen>
__global__ void __launch_bounds__(kRsThreads, BZ_RS_MIN_BLOCKS)
rs_scatter_kernel(const K* __restrict__ kin, ValGen vgen, K* __restrict__ kout, u32* __restrict__ vout, u32 n,
                  int shift, u32 mask, const u32* __restrict__ bases, u32 ntiles) {
    constexpr int ITEMS = RsCfg<K>::kItems;
    constexpr u32 TILE = kRsThreads * ITEMS;
    BZ_DYN_SMEM(unsigned char, rs_smem);
    u32* warp_cnt = reinterpret_cast<u32*>(rs_smem);              // [kRsWarps][256]
    u32* lbase = warp_cnt + kRsWarps * 256;                       // [256] first sorted slot of digit in tile
    u32* gdelta = lbase + 256;                                    // [256] global index minus tile slot
    u32* svals = gdelta + 256;                                    // [TILE]
    K* skeys = reinterpret_cast<K*>(svals + TILE);                // [TILE]
    const u32 w = warp_id(), l = lane_id();
    const u32 tile_base = blockIdx.x * TILE;
    const u32 count = min(TILE, n - tile_base);

    for (int i = threadIdx.x; i < kRsWarps * 256; i += kRsThreads) warp_cnt[i] = 0;
    __syncthreads();

    K key[ITEMS];
    u32 val[VOUT ? ITEMS : 1];
    u16 rank[ITEMS];
    const u32 base = tile_base + w * (32 * ITEMS);
    const u32 lt = lanemask_lt();
    // all loads of the tile are issued before anything consumes them (memory-level parallelism: the
    // v0 profile showed 60 % long-scoreboard stalls with the value loads serialised behind shared stores)
#pragma unroll
    for (int it = 0; it < ITEMS; it++) {
        u32 i = base + it * 32 + l;
        key[it] = (i < n) ? kin[i] : (K)0;
        if (VOUT) val[it] = (i < n) ? vgen(i) : 0u;
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
    // per digit: exclusive scan over warps, then exclusive scan over digits
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
        u32 incl = warp_scan_incl(real);
        __shared__ u32 wsum[kRsWarps];
        if (l == 31) wsum[w] = incl;
        __syncthreads();
        u32 wp = 0;
#pragma unroll
        for (int k = 0; k < kRsWarps; k++)
            if ((u32)k < w) wp += wsum[k];
        u32 excl = wp + incl - real;
        lbase[d] = excl;
        gdelta[d] = bases[d * ntiles + blockIdx.x] - excl;
    }
    __syncthreads();
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

// Exclusive scan of hist[256][ntiles] in digit-major order in two launches (one CTA per digit row):
// row totals first, then every row scans itself starting from the sum of the rows before it.
__global__ void __launch_bounds__(256) rs_row_total_kernel(const u32* __restrict__ hist, u32 ntiles, u32* __restrict__ row_total) {
    __shared__ u32 red[8];
    const u32* row = hist + (size_t)blockIdx.x * ntiles;
    u32 s = 0;
    for (u32 i = threadIdx.x; i < ntiles; i += 256) s += row[i];
    s = warp_reduce_sum(s);
    if (lane_id() == 0) red[warp_id()] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 t = 0;
        for (int k = 0; k < 8; k++) t += red[k];
        row_total[blockIdx.x] = t;
    }
}
__global__ void __launch_bounds__(256) rs_row_scan_kernel(u32* __restrict__ hist, u32 ntiles, const u32* __restrict__ row_total) {
    __shared__ u32 wsum[8];
    __shared__ u32 carry_s;
    const u32 d = blockIdx.x;
    {   // sum of the totals of rows 0..d-1 (256 values)
        u32 v = threadIdx.x < d ? row_total[threadIdx.x] : 0u;
        v = warp_reduce_sum(v);
        if (lane_id() == 0) wsum[warp_id()] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            u32 t = 0;
            for (int k = 0; k < 8; k++) t += wsum[k];
            carry_s = t;
        }
        __syncthreads();
    }
    u32* row = hist + (size_t)d * ntiles;
    u32 carry = carry_s;
    for (u32 base = 0; base < ntiles; base += 256) {
        const u32 i = base + threadIdx.x;
        const u32 v = i < ntiles ? row[i] : 0u;
        const u32 incl = warp_scan_incl(v);
        __syncthreads();  // wsum reuse
        if (lane_id() == 31) wsum[warp_id()] = incl;
        __syncthreads();
        u32 wp = 0, tot = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const u32 t = wsum[k];
            if ((u32)k < warp_id()) wp += t;
            tot += t;
        }
        if (i < ntiles) row[i] = carry + wp + incl - v;
        carry += tot;
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
// scratch (in u32 elements) for one pass over n records
template <typename K>
inline size_t rs_temp_elems(u32 n) {
    size_t h = (size_t)256 * rs_num_tiles<K>(n);
    return h + 256 /* row totals */ + scan_temp_elems((u32)h) + 16;
}

// One stable pass on digit bits [shift, shift+bits).
template <typename K, bool KOUT, bool VOUT, typename ValGen>
cudaError_t rs_pass(cudaStream_t st, const K* kin, ValGen vgen, K* kout, u32* vout, u32 n, int shift, int bits,
                    u32* temp) {
    if (n == 0) return cudaSuccess;
    const u32 ntiles = rs_num_tiles<K>(n);
    const u32 mask = (1u << bits) - 1u;
    u32* hist = temp;
    BZ_LAUNCH(ntiles, kRsThreads, 0, st, rs_tile_hist_kernel<K>)(kin, n, shift, mask, hist, ntiles); BZ_NOTE_LAUNCH();
    BZ_CUDA_TRY(cudaGetLastError());
    const u32 hn = 256 * ntiles;
    BZ_LAUNCH(256, 256, 0, st, rs_row_total_kernel)(hist, ntiles, hist + hn); BZ_NOTE_LAUNCH();
    BZ_LAUNCH(256, 256, 0, st, rs_row_scan_kernel)(hist, ntiles, hist + hn); BZ_NOTE_LAUNCH();
    BZ_CUDA_TRY(cudaGetLastError());
    auto kern = rs_scatter_kernel<K, KOUT, VOUT, ValGen>;
    const size_t smem = rs_scatter_smem<K>();
    // set every time: the attribute is per device and a process may drive several GPUs
    BZ_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    BZ_LAUNCH(ntiles, kRsThreads, smem, st, kern)(kin, vgen, kout, vout, n, shift, mask, hist, ntiles); BZ_NOTE_LAUNCH();
    BZ_CUDA_TRY(cudaGetLastError());
    return cudaSuccess;
}

// Full LSD sort of (key, value) records on key bits [0, nbits).  Ping-pongs between the A and B
// buffers; *result_in_b tells where the sorted records ended up.
// If vals_are_index the values of the input records are their positions 0..n-1 and `va` is not read.
template <typename K>
cudaError_t rs_sort_pairs(cudaStream_t st, K* ka, u32* va, K* kb, u32* vb, u32 n, int nbits, u32* temp,
                          bool* result_in_b, bool vals_are_index = false) {
    bool in_b = false;
    for (int shift = 0; shift < nbits; shift += 8) {
        int bits = nbits - shift < 8 ? nbits - shift : 8;
        K* src_k = in_b ? kb : ka;
        u32* src_v = in_b ? vb : va;
        K* dst_k = in_b ? ka : kb;
        u32* dst_v = in_b ? va : vb;
        if (shift == 0 && vals_are_index) {
            BZ_CUDA_TRY((rs_pass<K, true, true, ValIdentity>(st, src_k, ValIdentity{}, dst_k, dst_v, n, shift, bits, temp)));
        } else {
            BZ_CUDA_TRY((rs_pass<K, true, true, ValFromArray>(st, src_k, ValFromArray{src_v}, dst_k, dst_v, n, shift,
                                                              bits, temp)));
        }
        in_b = !in_b;
    }
    *result_in_b = in_b;
    return cudaSuccess;
}

}  // namespace bz3
