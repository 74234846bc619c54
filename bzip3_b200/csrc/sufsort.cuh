// sufsort.cuh -- suffix array construction and forward BWT on the GPU.
//
// Replaces the libsais SA-IS call of the reference (include/libsais.h:4095-4121 via
// src/libbz3.c:623).  The transform is a pure function of the text, so the construction is free to
// differ: this is prefix doubling with discarding, built on the radix passes of radix_sort.cuh.
//
//   round 0   key = (first 7 bytes, number of valid bytes) of every suffix, one 64-bit LSD sort.
//             The valid-byte count plays the role of the unique end marker: a suffix shorter than 7
//             bytes sorts before every longer suffix it is a prefix of (libsais order).
//   round r   only suffixes whose group is still ambiguous are kept ("unresolved list", in SA order);
//             key = (group rank, rank of suffix i+h), 2*ceil(log2(n+1)) bits; sort; regroup; compact.
//             h = 7, 14, 28, ...   Work shrinks with the list; worst case ceil(log2(n/7)) rounds.
//   finish    U[0] = T[n-1]; U[k] = T[SA'[k]-1] skipping the row with SA == 0; idx = rank of suffix 0.
//
// HBM layout (n = bytes entering the stage): T[n+16] u8, SA[n] u32, ISA[n+1] u32 (rank = 1 + first
// SA slot of the group, ISA[n] = 0 = "past the end"), two (u64 key, u32 value) record buffers, two
// (position, group) list buffers.
#pragma once
#include "common.cuh"
#include "scan.cuh"
#include "radix_sort.cuh"

namespace bz3 {

struct SufsortBuffers {
    u32* sa;      // [n]
    u32* isa;     // [n+1]
    u64* key[2];  // [n] each
    u32* val[2];  // [n] each
    u32* pos[2];  // [n] each   SA slot of list element
    u32* grp[2];  // [n] each   current group rank of list element
    u32* temp;    // radix / scan scratch
    u32* d_count; // device scalars (>= 4 u32)
    u32* h_count; // pinned mirror
    cudaEvent_t* ev = nullptr;  // optional: 2*max_ev events bracketing every radix sort (statistics)
    int max_ev = 0;
    int* used_ev = nullptr;
};

inline size_t sufsort_temp_elems(u32 n) {
    size_t a = rs_temp_elems<u64>(n);
    size_t b = 2 * scan_temp_elems(n) + 16;  // scan of 8-byte elements: counted in u32 units
    return (a > b ? a : b) + 64;
}

__global__ void sa_init_keys_kernel(const u8* __restrict__ T, u32 n, u64* __restrict__ keys) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // T is zero padded past n, so the 7 bytes can always be read
    u64 k = 0;
#pragma unroll
    for (int b = 0; b < 7; b++) k = (k << 8) | T[i + b];
    u32 valid = n - i < 7 ? n - i : 7;
    // bytes past the end are forced to zero even if the caller's padding is not
    if (valid < 7) k &= ~((1ull << (8 * (7 - valid))) - 1ull);
    keys[i] = (k << 8) | valid;
}

__global__ void sa_build_keys_kernel(const u32* __restrict__ val, const u32* __restrict__ grp,
                                     const u32* __restrict__ isa, u32 m, u32 h, int rank_bits,
                                     u64* __restrict__ keys) {
    u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    u32 s = val[j];
    keys[j] = ((u64)grp[j] << rank_bits) | (u64)isa[s + h];  // s + h <= n for every unresolved suffix
}

struct RegroupElem {
    u32 head_rank;  // max-scan: 1 + SA slot of the most recent group head
    u32 keep;       // sum-scan: number of still-ambiguous elements so far
};
struct RegroupOp {
    BZ_D RegroupElem operator()(const RegroupElem& a, const RegroupElem& b) const {
        RegroupElem r;
        r.head_rank = a.head_rank > b.head_rank ? a.head_rank : b.head_rank;
        r.keep = a.keep + b.keep;
        return r;
    }
};
// pos == nullptr means the list is the whole SA (round 0): slot j.
struct RegroupIn {
    const u64* keys;
    const u32* pos;
    u32 m;
    BZ_D bool head(u32 j) const { return j == 0 || keys[j] != keys[j - 1]; }
    BZ_D RegroupElem operator()(u32 j) const {
        RegroupElem e;
        bool hd = head(j);
        bool single = hd && (j + 1 == m || keys[j + 1] != keys[j]);
        e.head_rank = hd ? (pos ? pos[j] : j) + 1 : 0;
        e.keep = single ? 0 : 1;
        return e;
    }
};
struct RegroupOut {
    const u64* keys;
    const u32* val;
    const u32* pos;
    u32 m;
    u32* sa;
    u32* isa;
    u32* nval;
    u32* npos;
    u32* ngrp;
    BZ_D void operator()(u32 j, const RegroupElem& excl, const RegroupElem& incl) const {
        u32 slot = pos ? pos[j] : j;
        u32 s = val[j];
        sa[slot] = s;
        isa[s] = incl.head_rank;
        if (incl.keep != excl.keep) {
            u32 d = excl.keep;
            nval[d] = s;
            npos[d] = slot;
            ngrp[d] = incl.head_rank;
        }
    }
};

__global__ void bwt_gather_kernel(const u8* __restrict__ T, const u32* __restrict__ sa, u32 n, u32 idx,
                                  u8* __restrict__ U) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i == 0) U[0] = T[n - 1];
    u32 s = sa[i];
    if (i + 1 < idx) U[i + 1] = T[s - 1];
    else if (i + 1 > idx) U[i] = T[s - 1];
}

inline int bit_width_u32(u32 v) {
    int b = 0;
    while (v) { b++; v >>= 1; }
    return b;
}

// T must be readable (and zero) for 8 bytes past n.  Returns the primary index in *idx_out (host).
inline cudaError_t suffix_bwt(cudaStream_t st, const u8* T, u32 n, u8* U, const SufsortBuffers& B, s32* idx_out,
                              int* rounds_out = nullptr, u64* record_passes_out = nullptr) {
    if (n <= 1) {  // include/libsais.h:4098-4108
        if (n == 1) BZ_CUDA_TRY(cudaMemcpyAsync(U, T, 1, cudaMemcpyDeviceToDevice, st));
        *idx_out = (s32)n;
        return cudaSuccess;
    }
    const int TPB = 256;
    BZ_CUDA_TRY(cudaMemsetAsync(B.isa + n, 0, sizeof(u32), st));
    BZ_LAUNCH((n + TPB - 1) / TPB, TPB, 0, st, sa_init_keys_kernel)(T, n, B.key[0]); BZ_NOTE_LAUNCH();
    BZ_CUDA_TRY(cudaGetLastError());
    bool in_b = false;
    int nev = 0;
    auto mark = [&](int which) {
        if (B.ev && nev < B.max_ev) cudaEventRecord(B.ev[2 * nev + which], st);
        if (which == 1 && B.ev && nev < B.max_ev) nev++;
    };
    mark(0);
    BZ_CUDA_TRY(rs_sort_pairs<u64>(st, B.key[0], B.val[0], B.key[1], B.val[1], n, 64, B.temp, &in_b, true));
    mark(1);
    int kc = in_b ? 1 : 0;  // buffer holding the sorted records
    int lc = 0;             // list buffer to write
    RegroupElem ident{0u, 0u};
    RegroupElem* stemp = reinterpret_cast<RegroupElem*>(B.temp);
    RegroupElem* d_total = reinterpret_cast<RegroupElem*>(B.d_count);
    BZ_CUDA_TRY((device_scan<RegroupElem, RegroupOp, RegroupIn, RegroupOut>(
        st, RegroupIn{B.key[kc], nullptr, n},
        RegroupOut{B.key[kc], B.val[kc], nullptr, n, B.sa, B.isa, B.val[kc ^ 1], B.pos[lc], B.grp[lc]}, n, ident,
        RegroupOp{}, stemp, d_total)));
    BZ_CUDA_TRY(cudaMemcpyAsync(B.h_count, B.d_count, 2 * sizeof(u32), cudaMemcpyDeviceToHost, st));
    BZ_CUDA_TRY(cudaStreamSynchronize(st));
    u32 m = B.h_count[1];
    int vc = kc ^ 1;  // val buffer holding the list
    const int rank_bits = bit_width_u32(n);
    int rounds = 0;
    u64 record_passes = (u64)n * 8;  // round 0: eight 8-bit passes over n records
    for (u64 h = 7; m > 0; h *= 2) {
        if (h >= n) return cudaErrorUnknown;  // cannot happen: every suffix is unique within n symbols
        rounds++;
        record_passes += (u64)m * (u64)((2 * rank_bits + 7) / 8);
        BZ_LAUNCH((m + TPB - 1) / TPB, TPB, 0, st, sa_build_keys_kernel)(B.val[vc], B.grp[lc], B.isa, m, (u32)h, rank_bits,
                                                                  B.key[0]); BZ_NOTE_LAUNCH();
        BZ_CUDA_TRY(cudaGetLastError());
        // sort (key[0], val[vc]) <-> (key[1], val[vc^1])
        mark(0);
        BZ_CUDA_TRY(rs_sort_pairs<u64>(st, B.key[0], B.val[vc], B.key[1], B.val[vc ^ 1], m, 2 * rank_bits, B.temp, &in_b));
        mark(1);
        int sk = in_b ? 1 : 0;
        int sv = in_b ? (vc ^ 1) : vc;
        BZ_CUDA_TRY((device_scan<RegroupElem, RegroupOp, RegroupIn, RegroupOut>(
            st, RegroupIn{B.key[sk], B.pos[lc], m},
            RegroupOut{B.key[sk], B.val[sv], B.pos[lc], m, B.sa, B.isa, B.val[sv ^ 1], B.pos[lc ^ 1], B.grp[lc ^ 1]}, m,
            ident, RegroupOp{}, stemp, d_total)));
        BZ_CUDA_TRY(cudaMemcpyAsync(B.h_count, B.d_count, 2 * sizeof(u32), cudaMemcpyDeviceToHost, st));
        BZ_CUDA_TRY(cudaStreamSynchronize(st));
        m = B.h_count[1];
        vc = sv ^ 1;
        lc ^= 1;
    }
    if (rounds_out) *rounds_out = rounds;
    if (B.used_ev) *B.used_ev = nev;
    if (record_passes_out) *record_passes_out = record_passes;
    // primary index = rank of suffix 0 (1-based SA slot)
    BZ_CUDA_TRY(cudaMemcpyAsync(B.h_count, B.isa, sizeof(u32), cudaMemcpyDeviceToHost, st));
    BZ_CUDA_TRY(cudaStreamSynchronize(st));
    u32 idx = B.h_count[0];
    BZ_LAUNCH((n + TPB - 1) / TPB, TPB, 0, st, bwt_gather_kernel)(T, B.sa, n, idx, U); BZ_NOTE_LAUNCH();
    BZ_CUDA_TRY(cudaGetLastError());
    *idx_out = (s32)idx;
    return cudaSuccess;
}

}  // namespace bz3
