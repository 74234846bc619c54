// scan.cuh -- device-wide ordered prefix scan with an arbitrary associative (not necessarily
// commutative) operator.  Used for compaction offsets, run/token offsets in the mRLE codec, the
// 2-state token automaton of the mRLE decoder, and group-head propagation in the suffix sorter.
//
// Structure: reduce-per-tile -> (recursive) scan of tile sums -> rescan-per-tile with carry-in.
// Three launches, no inter-CTA spinning (so it can never hang the device).  Inputs are produced by a
// functor (usually fused flag computation), outputs are consumed by a functor, so no flag arrays are
// materialised in HBM: algorithmic traffic is 2 reads of the source + 1 write of whatever Out writes.
#pragma once
#include "common.cuh"

namespace bz3 {

template <typename T>
BZ_D T shfl_up_any(T v, int delta) {
    static_assert(sizeof(T) % 4 == 0, "scan element must be a multiple of 4 bytes");
    union { T t; u32 w[sizeof(T) / 4]; } u;
    u.t = v;
#pragma unroll
    for (int i = 0; i < (int)(sizeof(T) / 4); i++) u.w[i] = __shfl_up_sync(kFullMask, u.w[i], delta);
    return u.t;
}

constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

// Ordered block-wide exclusive scan of one aggregate per thread.  smem must hold kScanThreads/32 T's.
template <typename T, typename Op>
BZ_D T block_exclusive_scan(T agg, T identity, Op op, T* smem, T& block_total) {
    const u32 lane = lane_id(), warp = warp_id();
    T incl = agg;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T t = shfl_up_any(incl, o);
        if (lane >= (u32)o) incl = op(t, incl);
    }
    if (lane == 31) smem[warp] = incl;
    __syncthreads();
    T warp_prefix = identity, total = identity;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; w++) {
        T s = smem[w];
        if ((u32)w < warp) warp_prefix = op(warp_prefix, s);
        total = op(total, s);
    }
    T excl = shfl_up_any(incl, 1);
    if (lane == 0) excl = identity;
    block_total = total;
    __syncthreads();
    return op(warp_prefix, excl);
}

template <typename T, typename Op, typename In>
__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(In in, u32 n, T* tile_sums, T identity, Op op) {
    __shared__ T smem[kScanThreads / 32];
    const u32 base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    T acc = identity;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        u32 i = base + k;
        if (i < n) acc = op(acc, in(i));
    }
    T total;
    block_exclusive_scan(acc, identity, op, smem, total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

template <typename T, typename Op, typename In, typename Out>
__global__ void __launch_bounds__(kScanThreads)
scan_down_kernel(In in, Out out, u32 n, const T* tile_prefix, T identity, Op op) {
    __shared__ T smem[kScanThreads / 32];
    const u32 base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    T v[kScanItems];
    T acc = identity;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        u32 i = base + k;
        v[k] = (i < n) ? in(i) : identity;
        acc = op(acc, v[k]);
    }
    T total;
    T prefix = block_exclusive_scan(acc, identity, op, smem, total);
    if (tile_prefix) prefix = op(tile_prefix[blockIdx.x], prefix);
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        u32 i = base + k;
        T incl = op(prefix, v[k]);
        if (i < n) out(i, prefix, incl);
        prefix = incl;
    }
}

template <typename T>
struct PtrIn {
    const T* p;
    BZ_D T operator()(u32 i) const { return p[i]; }
};
template <typename T>
struct PtrOutExcl {
    T* p;
    BZ_D void operator()(u32 i, const T& excl, const T&) const { p[i] = excl; }
};

inline u32 scan_num_tiles(u32 n) { return (n + kScanTile - 1) / kScanTile; }

// Number of T elements of scratch needed by device_scan for n inputs.
inline size_t scan_temp_elems(u32 n) {
    size_t total = 0;
    u32 t = scan_num_tiles(n);
    while (true) {
        total += t + 1;
        if (t <= 1) break;
        t = scan_num_tiles(t);
    }
    return total + 8;
}

// Scans n functor-produced elements.  `temp` holds scan_temp_elems(n) T's.  If total_out != nullptr
// the reduction of all n elements is written there (device pointer).
template <typename T, typename Op, typename In, typename Out>
cudaError_t device_scan(cudaStream_t st, In in, Out out, u32 n, T identity, Op op, T* temp, T* total_out = nullptr) {
    if (n == 0) {
        if (total_out) BZ_CUDA_TRY(cudaMemcpyAsync(total_out, &identity, sizeof(T), cudaMemcpyHostToDevice, st));
        return cudaSuccess;
    }
    const u32 tiles = scan_num_tiles(n);
    T* sums = temp;
    BZ_LAUNCH(tiles, kScanThreads, 0, st, scan_reduce_kernel<T, Op, In>)(in, n, sums, identity, op); BZ_NOTE_LAUNCH();
    BZ_CUDA_TRY(cudaGetLastError());
    if (tiles > 1) {
        // exclusive scan of the tile sums in place; its grand total lands in sums[tiles]
        BZ_CUDA_TRY((device_scan<T, Op, PtrIn<T>, PtrOutExcl<T>>(st, PtrIn<T>{sums}, PtrOutExcl<T>{sums}, tiles, identity,
                                                                 op, temp + tiles + 1, total_out)));
        BZ_LAUNCH(tiles, kScanThreads, 0, st, scan_down_kernel<T, Op, In, Out>)(in, out, n, sums, identity, op); BZ_NOTE_LAUNCH();
    } else {
        if (total_out) BZ_CUDA_TRY(cudaMemcpyAsync(total_out, sums, sizeof(T), cudaMemcpyDeviceToDevice, st));
        BZ_LAUNCH(1, kScanThreads, 0, st, scan_down_kernel<T, Op, In, Out>)(in, out, n, (const T*)nullptr, identity, op); BZ_NOTE_LAUNCH();
    }
    BZ_CUDA_TRY(cudaGetLastError());
    return cudaSuccess;
}

struct SumU32 {
    BZ_D u32 operator()(u32 a, u32 b) const { return a + b; }
};
struct MaxU32 {
    BZ_D u32 operator()(u32 a, u32 b) const { return a > b ? a : b; }
};

}  // namespace bz3
