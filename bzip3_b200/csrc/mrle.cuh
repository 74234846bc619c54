// mrle.cuh -- the "mRLE" pre-pass of the block codec, encode and decode, as parallel scans.
//
// Restates mrlec / mrled (reference src/libbz3.c:264-329):
//   * a 32-byte bitmap (LSB first) flags every byte value whose runs are worth collapsing:
//       flagged(c)  <=>  sum over maximal runs of c of  (L-1) - floor((L-1)/255) - 1  >  0
//   * a run of L bytes of a flagged value becomes  c, 255 x floor((L-1)/255), (L-1) mod 255
//   * unflagged values are copied.
// Encode: run heads -> compacted head list -> per-run gain (shared-memory histogram) -> bitmap ->
// per-run output size -> exclusive scan -> one thread per run writes its token.
// Decode: the token grammar is a 2-state automaton (expect-symbol / inside-count); composing the
// per-byte transition maps with an ordered scan yields every byte's role without a serial parse;
// a second scan gives output offsets and the owning symbol; one thread per input byte expands.
#pragma once
#include "common.cuh"
#include "scan.cuh"

namespace bz3 {

// ------------------------------------------------------------------------------------ encode
struct HeadFlagIn {
    const u8* t;
    BZ_D u32 operator()(u32 i) const { return (i == 0 || t[i] != t[i - 1]) ? 1u : 0u; }
};
struct HeadCompactOut {
    u32* heads;
    BZ_D void operator()(u32 i, const u32& excl, const u32& incl) const {
        if (incl != excl) heads[excl] = i;
    }
};

// gain[c] += (L-1) - (L-1)/255 - 1 for every run; heads[nruns] must hold n
__global__ void __launch_bounds__(256) mrle_gain_kernel(const u8* __restrict__ t, const u32* __restrict__ heads,
                                                        u32 nruns, int* __restrict__ gain) {
    __shared__ int sg[256];
    sg[threadIdx.x] = 0;
    __syncthreads();
    for (u32 k = blockIdx.x * blockDim.x + threadIdx.x; k < nruns; k += gridDim.x * blockDim.x) {
        u32 h = heads[k];
        int L = (int)(heads[k + 1] - h);
        int g = (L - 1) - (L - 1) / 255 - 1;
        if (g) atomicAdd(&sg[t[h]], g);
    }
    __syncthreads();
    if (sg[threadIdx.x]) atomicAdd(&gain[threadIdx.x], sg[threadIdx.x]);
}

// writes the 32-byte bitmap at out[0..32) and a 256-entry 0/1 table
__global__ void mrle_bitmap_kernel(const int* __restrict__ gain, u8* __restrict__ out, u8* __restrict__ flagged) {
    u32 c = threadIdx.x;  // 256 threads
    u32 f = gain[c] > 0 ? 1u : 0u;
    flagged[c] = (u8)f;
    u32 bits = __ballot_sync(kFullMask, f);  // lane k of warp w <-> value 32*w + k
    if ((c & 31) == 0) {
        out[(c >> 3) + 0] = (u8)bits;
        out[(c >> 3) + 1] = (u8)(bits >> 8);
        out[(c >> 3) + 2] = (u8)(bits >> 16);
        out[(c >> 3) + 3] = (u8)(bits >> 24);
    }
}

struct RunSizeIn {
    const u8* t;
    const u32* heads;
    const u8* flagged;
    BZ_D u32 operator()(u32 k) const {
        u32 h = heads[k];
        u32 L = heads[k + 1] - h;
        return flagged[t[h]] ? 2u + (L - 1) / 255u : L;
    }
};
struct RunEmitOut {
    const u8* t;
    const u32* heads;
    const u8* flagged;
    u8* out;  // already advanced past the bitmap
    BZ_D void operator()(u32 k, const u32& excl, const u32& incl) const {
        u32 h = heads[k];
        u32 L = heads[k + 1] - h;
        u8 c = t[h];
        u8* o = out + excl;
        if (flagged[c]) {
            u32 q = (L - 1) / 255u;
            o[0] = c;
            for (u32 i = 0; i < q; i++) o[1 + i] = 255;
            o[1 + q] = (u8)((L - 1) % 255u);
        } else {
            for (u32 i = 0; i < L; i++) o[i] = c;
        }
    }
};

struct MrleScratch {
    u32* heads;    // [n+1]
    u32* temp;     // scan scratch, scan_temp_elems(n) u32
    int* gain;     // [256]
    u8* flagged;   // [256]
    u32* d_count;  // device scalars
    u32* h_count;  // pinned mirror
};

// in: n bytes; out: capacity >= 32 + 2n.  Returns encoded size in *out_size (host).
inline cudaError_t mrle_encode(cudaStream_t st, const u8* in, u32 n, u8* out, const MrleScratch& S, s32* out_size) {
    BZ_CUDA_TRY(cudaMemsetAsync(S.gain, 0, 256 * sizeof(int), st));
    u32 nruns = 0;
    if (n > 0) {
        BZ_CUDA_TRY((device_scan<u32, SumU32, HeadFlagIn, HeadCompactOut>(st, HeadFlagIn{in}, HeadCompactOut{S.heads}, n, 0u,
                                                                         SumU32{}, S.temp, S.d_count)));
        BZ_CUDA_TRY(cudaMemcpyAsync(S.h_count, S.d_count, sizeof(u32), cudaMemcpyDeviceToHost, st));
        BZ_CUDA_TRY(cudaStreamSynchronize(st));
        nruns = S.h_count[0];
        BZ_CUDA_TRY(cudaMemcpyAsync(S.heads + nruns, &n, sizeof(u32), cudaMemcpyHostToDevice, st));
        u32 blocks = (nruns + 255) / 256;
        if (blocks > 148 * 8) blocks = 148 * 8;
        BZ_LAUNCH(blocks, 256, 0, st, mrle_gain_kernel)(in, S.heads, nruns, S.gain); BZ_NOTE_LAUNCH();
        BZ_CUDA_TRY(cudaGetLastError());
    }
    BZ_LAUNCH(1, 256, 0, st, mrle_bitmap_kernel)(S.gain, out, S.flagged); BZ_NOTE_LAUNCH();
    BZ_CUDA_TRY(cudaGetLastError());
    u32 total = 0;
    if (nruns > 0) {
        BZ_CUDA_TRY((device_scan<u32, SumU32, RunSizeIn, RunEmitOut>(st, RunSizeIn{in, S.heads, S.flagged},
                                                                    RunEmitOut{in, S.heads, S.flagged, out + 32}, nruns,
                                                                    0u, SumU32{}, S.temp, S.d_count)));
        BZ_CUDA_TRY(cudaMemcpyAsync(S.h_count, S.d_count, sizeof(u32), cudaMemcpyDeviceToHost, st));
        BZ_CUDA_TRY(cudaStreamSynchronize(st));
        total = S.h_count[0];
    }
    *out_size = (s32)(32 + total);
    return cudaSuccess;
}

// ------------------------------------------------------------------------------------ decode
// automaton states: 0 = next byte is a symbol, 1 = next byte is a count byte.
// A transition map is packed as bits: bit0 = image of state 0, bit1 = image of state 1.
struct MapOp {
    BZ_D u32 operator()(u32 a, u32 b) const {  // apply a, then b
        u32 a0 = a & 1u, a1 = (a >> 1) & 1u;
        return ((b >> a0) & 1u) | (((b >> a1) & 1u) << 1);
    }
};
struct TokenMapIn {
    const u8* in;       // stream after the bitmap
    const u8* flagged;  // [256]
    BZ_D u32 operator()(u32 i) const {
        u8 b = in[i];
        u32 from0 = flagged[b] ? 1u : 0u;  // symbol: flagged -> counts follow
        u32 from1 = (b == 255) ? 1u : 0u;  // count byte: 255 continues
        return from0 | (from1 << 1);
    }
};
struct StateOut {
    u8* state;  // state BEFORE byte i
    BZ_D void operator()(u32 i, const u32& excl, const u32&) const { state[i] = (u8)(excl & 1u); }
};

struct ExpandElem {
    u32 count;    // saturating sum of output bytes
    u32 sympos1;  // 1 + index of the most recent symbol byte
    u32 term1;    // 1 + index of the most recent terminating (non-255) count byte
};
constexpr u32 kSat = 0x7FFFFFFFu;
struct ExpandOp {
    BZ_D ExpandElem operator()(const ExpandElem& a, const ExpandElem& b) const {
        ExpandElem r;
        u32 s = a.count + b.count;
        r.count = (s < a.count || s > kSat) ? kSat : s;
        r.sympos1 = a.sympos1 > b.sympos1 ? a.sympos1 : b.sympos1;
        r.term1 = a.term1 > b.term1 ? a.term1 : b.term1;
        return r;
    }
};
struct ExpandIn {
    const u8* in;
    const u8* state;
    const u8* flagged;
    BZ_D ExpandElem operator()(u32 i) const {
        ExpandElem e;
        u8 b = in[i];
        if (state[i] == 0) {
            e.count = flagged[b] ? 0u : 1u;
            e.sympos1 = i + 1;
            e.term1 = 0;
        } else {
            e.count = (b == 255) ? 255u : (u32)b + 1u;
            e.sympos1 = 0;
            e.term1 = (b == 255) ? 0u : i + 1;
        }
        return e;
    }
};
struct ExpandOut {
    const u8* in;
    const u8* state;
    u8* out;
    u32 outlen;
    BZ_D void operator()(u32 i, const ExpandElem& excl, const ExpandElem& incl) const {
        u32 from = excl.count, to = incl.count;
        if (from >= outlen || to == from) return;
        if (to > outlen) to = outlen;
        u8 c = (state[i] == 0) ? in[i] : in[incl.sympos1 - 1];
        for (u32 k = from; k < to; k++) out[k] = c;
    }
};

// Reference quirk (src/libbz3.c:320-322): when the input ends inside a token the loop variable `pc`
// keeps an older value and  run += pc + 1  is still executed.  Only reachable with malformed data.
__global__ void mrle_tail_fix_kernel(const u8* in, u32 m, u32 final_state, const ExpandElem* total, u8* out,
                                     u32 outlen, u32* total_out) {
    u32 produced = total->count;
    if (final_state == 1 && m > 0) {
        u32 extra;
        u8 last = in[m - 1];
        if (total->sympos1 == m) {
            // the flagged symbol was the last byte: pc is the previous terminator (or -1)
            extra = total->term1 ? (u32)in[total->term1 - 1] + 1u : 0u;
        } else {
            extra = (u32)last + 1u;  // last byte was a 255 count byte: pc == 255
        }
        u8 c = in[total->sympos1 - 1];
        u32 from = produced, to = produced + extra;
        if (to > outlen) to = outlen;
        for (u32 k = from; k < to && k < outlen; k++) out[k] = c;
        u32 s = produced + extra;
        produced = (s > kSat) ? kSat : s;
    }
    *total_out = produced;
}

struct MrleDecScratch {
    u8* state;     // [maxin]
    u32* temp;     // scan scratch: 3 * scan_temp_elems(maxin) u32
    u8* flagged;   // [256]
    u32* d_count;  // >= 8 u32
    u32* h_count;
};

__global__ void mrle_unpack_bitmap_kernel(const u8* in, u8* flagged) {
    u32 c = threadIdx.x;
    flagged[c] = (in[c >> 3] >> (c & 7)) & 1;
}

// Mirrors mrled(in, out, outlen, maxin): *err = 1 when the output does not come to exactly outlen.
inline cudaError_t mrle_decode(cudaStream_t st, const u8* in, u32 maxin, u8* out, u32 outlen, const MrleDecScratch& S,
                               int* err) {
    if (maxin < 32) { *err = 1; return cudaSuccess; }  // :310
    const u32 m = maxin - 32;
    const u8* body = in + 32;
    BZ_LAUNCH(1, 256, 0, st, mrle_unpack_bitmap_kernel)(in, S.flagged); BZ_NOTE_LAUNCH();
    BZ_CUDA_TRY(cudaGetLastError());
    u32* d_final_map = S.d_count;                                              // [0]
    ExpandElem* d_total = reinterpret_cast<ExpandElem*>(S.d_count + 1);       // [1..3]
    u32* d_produced = S.d_count + 4;                                           // [4]
    BZ_CUDA_TRY((device_scan<u32, MapOp, TokenMapIn, StateOut>(st, TokenMapIn{body, S.flagged}, StateOut{S.state}, m,
                                                              2u /* identity map: 0->0, 1->1 */, MapOp{}, S.temp,
                                                              d_final_map)));
    ExpandElem ident{0u, 0u, 0u};
    BZ_CUDA_TRY((device_scan<ExpandElem, ExpandOp, ExpandIn, ExpandOut>(
        st, ExpandIn{body, S.state, S.flagged}, ExpandOut{body, S.state, out, outlen}, m, ident, ExpandOp{},
        reinterpret_cast<ExpandElem*>(S.temp), d_total)));
    BZ_CUDA_TRY(cudaMemcpyAsync(S.h_count, S.d_count, sizeof(u32), cudaMemcpyDeviceToHost, st));
    BZ_CUDA_TRY(cudaStreamSynchronize(st));
    u32 final_state = S.h_count[0] & 1u;  // image of state 0 under the whole stream
    BZ_LAUNCH(1, 1, 0, st, mrle_tail_fix_kernel)(body, m, final_state, d_total, out, outlen, d_produced); BZ_NOTE_LAUNCH();
    BZ_CUDA_TRY(cudaGetLastError());
    BZ_CUDA_TRY(cudaMemcpyAsync(S.h_count, d_produced, sizeof(u32), cudaMemcpyDeviceToHost, st));
    BZ_CUDA_TRY(cudaStreamSynchronize(st));
    *err = (S.h_count[0] >= outlen) ? 0 : 1;
    return cudaSuccess;
}

}  // namespace bz3
