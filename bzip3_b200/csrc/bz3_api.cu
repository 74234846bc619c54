// bz3_api.cu -- the C ABI of the B200 block codec (include/libbz3.h, include/bz3_b200.h).
//
// Host-side orchestration only: every byte of block data is transformed by the CUDA kernels in the
// headers included below.  Stage order, header layout, validation order and error numbers restate
// bz3_encode_block / bz3_decode_block (reference src/libbz3.c:585-809); the batch entry points restate
// bz3_encode_blocks / bz3_decode_blocks (:813-872) with one host thread + one CUDA stream per block.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/bz3_b200.h"
#include "common.cuh"
#include "scan.cuh"
#include "radix_sort.cuh"
#include "crc.cuh"
#include "mrle.cuh"
#include "lzp.cuh"
#include "lzp_parallel.cuh"
#include "lzp_scan.cuh"
#include "sufsort.cuh"
#include "unbwt.cuh"
#include "cm.cuh"
#include "stream.h"

using namespace bz3;

// One stream per block: with more blocks in flight than hardware queues (CUDA_DEVICE_MAX_CONNECTIONS, default 8) streams
// alias and a copy queued behind one block's long coder kernel stalls other blocks' launches.  The library does not touch
// the process environment; bench.py and the bz3b200 tool set CUDA_DEVICE_MAX_CONNECTIONS=32 themselves before the CUDA
// context exists, and a host program that keeps many blocks in flight should do the same.

#ifndef BZ3_VERSION_STRING
#define BZ3_VERSION_STRING "1.5.2-b200"
#endif

namespace {

constexpr size_t kAlign = 256;
inline size_t align_up(size_t v) { return (v + kAlign - 1) & ~(kAlign - 1); }

struct Arena {
    u8* base = nullptr;
    size_t size = 0, used = 0;
    void reset() { used = 0; }
    template <typename T>
    T* take(size_t count) {
        size_t bytes = align_up(count * sizeof(T));
        if (used + bytes > size) return nullptr;
        T* p = reinterpret_cast<T*>(base + used);
        used += bytes;
        return p;
    }
};

struct StageClock {
    cudaEvent_t a[BZ3_STAGE_COUNT], b[BZ3_STAGE_COUNT];
    bool used[BZ3_STAGE_COUNT];
};

}  // namespace

struct bz3_state {
    s32 block_size;
    s8 last_error;
    int device;
    cudaStream_t stream;
    size_t cap;         // capacity of each data buffer
    u8* d_buf[3];       // ping-pong buffers + payload buffer
    s32* d_lut;         // LZP table
    u32* d_scal;        // device scalars
    u32* h_scal;        // pinned mirror (1024 u32)
    Arena arena;        // stage workspace: a lease from the device's pool, valid inside one stage call (ArenaLease)
    size_t arena_need;  // workspace bytes the largest stage of this state needs
    bool pool_attached;
    size_t device_bytes;
    // data staged by bz3_b200_upload / produced by *_resident
    int resident_buf;   // index of the buffer holding resident data
    s32 resident_size;
    // statistics
    StageClock clk;
    double stage_ms[2][BZ3_STAGE_COUNT];
    u64 launches;
    u64 sort_records;
    s32 sort_rounds;
    double sort_ms;
    cudaEvent_t sort_ev[2 * 40];
};

namespace {

int env_int(const char* name, int fallback) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : fallback;
}

// ------------------------------------------------------------------------------- stage workspace pool
// mRLE, the suffix sort and the inverse BWT need a workspace of up to 48 bytes per input byte (12.9 GB for a 256 MiB
// block, 25.7 GB at 511 MiB) for the fraction of a second a block spends in them, while the entropy stage then holds
// the block for seconds on ONE thread block.  Throughput on a B200 is blocks in flight, so the workspace must not be a
// per-state cost: the states of a device share kArenaSlots workspaces (BZ3_B200_ARENAS, default 2: one block sorts
// while the next one's launches are queued), leased for the duration of one stage call.  What a state owns is its
// three data buffers and the LZP table: ~3.06 bytes per byte of block size instead of ~51, i.e. >150 blocks of
// 256 MiB resident in the 180 GB of one B200 instead of 13.
constexpr int kMaxDevices = 64;
constexpr int kMaxArenaSlots = 8;

struct ArenaPool {
    std::mutex m;
    std::condition_variable cv;
    struct Slot {
        u8* base = nullptr;
        size_t size = 0;
        bool busy = false;
    } slot[kMaxArenaSlots];
    int nslots = 0;   // fixed at the first attach
    int states = 0;   // live states of this device
};
ArenaPool g_pool[kMaxDevices];

// Makes a free slot at least `need` bytes large (caller holds the pool's mutex).  The new workspace is allocated before
// the old one is released when both fit; if neither order works the old size is restored so that the states which
// rely on it keep working.
bool pool_grow_slot(ArenaPool::Slot& sl, size_t need) {
    if (sl.size >= need) return true;
    u8* fresh = nullptr;
    if (cudaMalloc(&fresh, need) != cudaSuccess) {
        cudaGetLastError();   // reported through the return value, not left behind as a sticky error
        fresh = nullptr;
        const size_t old = sl.size;
        if (sl.base) cudaFree(sl.base);
        sl.base = nullptr;
        sl.size = 0;
        if (cudaMalloc(&fresh, need) != cudaSuccess) {
            cudaGetLastError();
            if (old && cudaMalloc(&sl.base, old) == cudaSuccess) sl.size = old;
            else { sl.base = nullptr; cudaGetLastError(); }
            return false;
        }
    } else if (sl.base) {
        cudaFree(sl.base);
    }
    sl.base = fresh;
    sl.size = need;
    return true;
}

// bz3_new: registers a state and sizes the workspaces for it, so that running out of device memory is reported where
// the reference reports it (bz3_new returns NULL, src/libbz3.c:553-561) and never in the middle of a block.  Slot 0 is
// always large enough for every live state; the further slots are sized from the second state on (a lone state
// cannot use two) and are a bonus: failing to grow one of them is not an error.
bool pool_attach(int dev, size_t need) {
    ArenaPool& P = g_pool[dev];
    std::unique_lock<std::mutex> lk(P.m);
    if (P.nslots == 0) {
        int k = env_int("BZ3_B200_ARENAS", 2);
        P.nslots = k < 1 ? 1 : (k > kMaxArenaSlots ? kMaxArenaSlots : k);
    }
    const int want = std::min(P.nslots, P.states + 1);
    for (int k = 0; k < want; k++) {
        P.cv.wait(lk, [&] { return !P.slot[k].busy; });
        if (!pool_grow_slot(P.slot[k], need) && k == 0) return false;
    }
    P.states++;
    return true;
}

void pool_detach(int dev) {
    ArenaPool& P = g_pool[dev];
    std::unique_lock<std::mutex> lk(P.m);
    if (--P.states > 0) return;
    P.states = 0;
    for (int k = 0; k < P.nslots; k++) {   // the last state of the device takes the workspaces with it
        P.cv.wait(lk, [&] { return !P.slot[k].busy; });
        if (P.slot[k].base) cudaFree(P.slot[k].base);
        P.slot[k] = ArenaPool::Slot();
    }
    P.nslots = 0;
}

// One stage call's hold on a workspace.  The constructor waits for a free one; the destructor gives it back only when
// the state's stream has drained, so no kernel of this stage can still be using it (every stage driver below has
// synchronised by then on its good path; this covers the early returns).
struct ArenaLease {
    bz3_state* s;
    int k = -1;
    ArenaLease(bz3_state* st, int stage) : s(st) {
        ArenaPool& P = g_pool[s->device];
        {
            std::unique_lock<std::mutex> lk(P.m);
            bool any = false;
            P.cv.wait(lk, [&] {   // a free workspace that is large enough for this state (slot 0 always is, see pool_attach)
                any = false;
                for (int i = 0; i < P.nslots; i++)
                    if (P.slot[i].size >= s->arena_need) {
                        any = true;
                        if (!P.slot[i].busy) { k = i; return true; }
                    }
                return !any;
            });
            if (!any || k < 0) { k = -1; return; }   // only if a later bz3_new lost slot 0 while running out of memory
            P.slot[k].busy = true;
            s->arena.base = P.slot[k].base;
            s->arena.size = P.slot[k].size;
            s->arena.used = 0;
        }
        if (s->clk.a[stage]) cudaEventRecord(s->clk.a[stage], s->stream);   // the stage's clock starts once the workspace is there
    }
    ~ArenaLease() {
        if (k < 0) return;
        cudaStreamSynchronize(s->stream);
        give_back();
    }
    bool ok() const { return k >= 0; }
    ArenaLease(const ArenaLease&) = delete;
    ArenaLease& operator=(const ArenaLease&) = delete;

private:
    void give_back() {
        ArenaPool& P = g_pool[s->device];
        s->arena = Arena();
        {
            std::lock_guard<std::mutex> lk(P.m);
            P.slot[k].busy = false;
        }
        k = -1;
        P.cv.notify_all();
    }
};

size_t sufsort_arena_bytes(size_t n) {
    return align_up(4 * n) + align_up(4 * (n + 1)) + 2 * align_up(8 * n) + 6 * align_up(4 * n) +
           align_up(4 * sufsort_temp_elems((u32)n)) + 16 * kAlign;
}
size_t lzp_arena_bytes(size_t n) {
    return 4 * align_up(4 * n) + align_up(4 * (n + 8)) + align_up(n + 8) +
           align_up(4 * rs_temp_elems<u32>((u32)n)) + 16 * kAlign;
}
size_t other_arena_bytes(size_t n) {
    size_t mr = align_up(4 * (n + 2)) + align_up(12 * scan_temp_elems((u32)n)) + align_up(n) + 8 * kAlign;
    size_t ub = align_up(4 * (n + 2)) + 5 * align_up(4 * ((n >> 5) + 8)) + align_up(4 * rs_temp_elems<u8>((u32)n)) +
                align_up(65536 * 4) + 8 * kAlign;
    const size_t lz = lzp_arena_bytes(n);
    return std::max(lz, mr > ub ? mr : ub);
}

bool carve_sufsort(bz3_state* s, u32 n, SufsortBuffers& B) {   // inside an ArenaLease
    Arena& A = s->arena;
    A.reset();
    B.sa = A.take<u32>(n);
    B.isa = A.take<u32>((size_t)n + 1);
    for (int i = 0; i < 2; i++) B.key[i] = A.take<u64>(n);
    for (int i = 0; i < 2; i++) B.val[i] = A.take<u32>(n);
    for (int i = 0; i < 2; i++) B.pos[i] = A.take<u32>(n);
    for (int i = 0; i < 2; i++) B.grp[i] = A.take<u32>(n);
    B.temp = A.take<u32>(sufsort_temp_elems(n));
    B.d_count = s->d_scal;
    B.h_count = s->h_scal;
    return B.temp != nullptr;
}

struct Timer {
    bz3_state* s;
    int stage;
    Timer(bz3_state* st, int sg) : s(st), stage(sg) {
        cudaEventRecord(s->clk.a[stage], s->stream);
    }
    ~Timer() {
        cudaEventRecord(s->clk.b[stage], s->stream);
        s->clk.used[stage] = true;
    }
};
void clocks_begin(bz3_state*) {}  // entries reset themselves when collected
void clocks_collect(bz3_state* s, int decode) {
    cudaStreamSynchronize(s->stream);
    for (int i = 0; i < BZ3_STAGE_COUNT; i++)
        if (s->clk.used[i]) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, s->clk.a[i], s->clk.b[i]) == cudaSuccess) s->stage_ms[decode][i] += ms;
            s->clk.used[i] = false;  // collected once
        }
}

struct LaunchScope {  // folds this thread's launch count into the state
    bz3_state* s;
    u64 before;
    explicit LaunchScope(bz3_state* st) : s(st), before(launch_counter()) {}
    ~LaunchScope() { s->launches += launch_counter() - before; }
};

inline void put32(u8* p, u32 v) { p[0] = (u8)v; p[1] = (u8)(v >> 8); p[2] = (u8)(v >> 16); p[3] = (u8)(v >> 24); }
inline u32 get32(const u8* p) { return (u32)p[0] | ((u32)p[1] << 8) | ((u32)p[2] << 16) | ((u32)p[3] << 24); }

// ---------------------------------------------------------------------------------- stage drivers
cudaError_t run_crc(bz3_state* s, const u8* d_in, u32 n, u32* crc_out) {
    BZ_CUDA_TRY(crc_launch(s->stream, d_in, n, 1u, s->d_scal + 32));
    BZ_CUDA_TRY(cudaMemcpyAsync(s->h_scal + 32, s->d_scal + 32, 4, cudaMemcpyDeviceToHost, s->stream));
    BZ_CUDA_TRY(cudaStreamSynchronize(s->stream));
    *crc_out = s->h_scal[32];
    return cudaSuccess;
}

cudaError_t run_rle_encode(bz3_state* s, const u8* d_in, u32 n, u8* d_out, s32* out_size) {
    ArenaLease lease(s, BZ3_STAGE_RLE);
    if (!lease.ok()) return cudaErrorMemoryAllocation;
    Arena& A = s->arena;
    A.reset();
    MrleScratch S;
    S.heads = A.take<u32>((size_t)n + 2);
    S.temp = A.take<u32>(scan_temp_elems(n));
    S.gain = A.take<int>(256);
    S.flagged = A.take<u8>(256);
    S.d_count = s->d_scal;
    S.h_count = s->h_scal;
    if (!S.flagged) return cudaErrorMemoryAllocation;
    return mrle_encode(s->stream, d_in, n, d_out, S, out_size);
}

cudaError_t run_rle_decode(bz3_state* s, const u8* d_in, u32 maxin, u8* d_out, u32 outlen, int* err) {
    ArenaLease lease(s, BZ3_STAGE_RLE);
    if (!lease.ok()) return cudaErrorMemoryAllocation;
    Arena& A = s->arena;
    A.reset();
    MrleDecScratch S;
    S.state = A.take<u8>((size_t)maxin + 8);
    S.temp = A.take<u32>(3 * scan_temp_elems(maxin));
    S.flagged = A.take<u8>(256);
    S.d_count = s->d_scal;
    S.h_count = s->h_scal;
    if (!S.flagged) return cudaErrorMemoryAllocation;
    return mrle_decode(s->stream, d_in, maxin, d_out, outlen, S, err);
}

cudaError_t run_lzp_encode(bz3_state* s, const u8* d_in, s32 n, u8* d_out, s32* result) {
    if (n < kLzpMinMatch + 32) { *result = -1; return cudaSuccess; }   // src/libbz3.c:244
    ArenaLease lease(s, BZ3_STAGE_LZP);
    if (!lease.ok()) return cudaErrorMemoryAllocation;
    Arena& A = s->arena;
    A.reset();
    LzpScanBuffers B;
    const u32 m = (u32)n - 4u;
    for (int i = 0; i < 2; i++) B.key[i] = A.take<u32>(m);
    for (int i = 0; i < 2; i++) B.idx[i] = A.take<u32>(m);
    B.P = A.take<u32>((size_t)n + 8);
    B.code = A.take<u8>((size_t)n + 8);
    B.lut = s->d_lut;
    B.temp = A.take<u32>(rs_temp_elems<u32>(m));
    if (!B.temp) return cudaErrorMemoryAllocation;
    s32* d_res = reinterpret_cast<s32*>(s->d_scal + 8);
    BZ_CUDA_TRY(lzp_scan_encode(s->stream, d_in, n, d_out, B, d_res));
    BZ_CUDA_TRY(cudaStreamSynchronize(s->stream));   // see run_cm_encode: nothing waits in a queue behind a long kernel
    BZ_CUDA_TRY(cudaMemcpyAsync(s->h_scal + 8, d_res, 4, cudaMemcpyDeviceToHost, s->stream));
    BZ_CUDA_TRY(cudaStreamSynchronize(s->stream));
    *result = (s32)s->h_scal[8];
    return cudaSuccess;
}

cudaError_t run_lzp_decode(bz3_state* s, const u8* d_in, s32 n, u8* d_out, s32 max, s32* result) {
    if (n < 4) { *result = -1; return cudaSuccess; }
    BZ_CUDA_TRY(cudaMemsetAsync(s->d_lut, 0, sizeof(s32) * kLzpSlots, s->stream));
    BZ_LAUNCH(1, kLzpBulkThreads, 0, s->stream, lzp_decode_bulk_kernel)(d_in, n, d_out, max, s->d_lut, reinterpret_cast<s32*>(s->d_scal + 8));
    BZ_NOTE_LAUNCH();
    BZ_CUDA_TRY(cudaGetLastError());
    BZ_CUDA_TRY(cudaStreamSynchronize(s->stream));   // see run_cm_encode: nothing waits in a queue behind a long kernel
    BZ_CUDA_TRY(cudaMemcpyAsync(s->h_scal + 8, s->d_scal + 8, 4, cudaMemcpyDeviceToHost, s->stream));
    BZ_CUDA_TRY(cudaStreamSynchronize(s->stream));
    *result = (s32)s->h_scal[8];
    return cudaSuccess;
}

cudaError_t run_bwt(bz3_state* s, u8* d_in, u32 n, u8* d_out, s32* idx) {
    ArenaLease lease(s, BZ3_STAGE_BWT);
    if (!lease.ok()) return cudaErrorMemoryAllocation;
    SufsortBuffers B;
    if (!carve_sufsort(s, n, B)) return cudaErrorMemoryAllocation;
    BZ_CUDA_TRY(cudaMemsetAsync(d_in + n, 0, 16, s->stream));  // zero padding read by the 7-byte key kernel
    int rounds = 0, nev = 0;
    u64 rp = 0;
    B.ev = s->sort_ev;
    B.max_ev = 40;
    B.used_ev = &nev;
    cudaError_t e = suffix_bwt(s->stream, d_in, n, d_out, B, idx, &rounds, &rp);
    s->sort_rounds = rounds;
    s->sort_records += rp;
    if (e == cudaSuccess && cudaStreamSynchronize(s->stream) == cudaSuccess)
        for (int k = 0; k < nev; k++) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, s->sort_ev[2 * k], s->sort_ev[2 * k + 1]) == cudaSuccess) s->sort_ms += ms;
        }
    return e;
}

cudaError_t run_unbwt(bz3_state* s, const u8* d_in, u32 n, s32 idx, u8* d_out, int* status) {
    ArenaLease lease(s, BZ3_STAGE_BWT);
    if (!lease.ok()) return cudaErrorMemoryAllocation;
    Arena& A = s->arena;
    A.reset();
    UnbwtBuffers B;
    int lg;
    u32 K;
    unbwt_geometry(n, &lg, &K);
    B.psi = A.take<u32>((size_t)n + 2);
    for (int i = 0; i < 2; i++) B.nxt[i] = A.take<u32>((size_t)K + 2);
    for (int i = 0; i < 2; i++) B.dist[i] = A.take<u32>((size_t)K + 2);
    B.len = A.take<u32>((size_t)K + 2);
    B.hist = A.take<u32>(256);
    B.start = A.take<u32>(257);
    B.big = A.take<u32>(65536);
    B.temp = A.take<u32>(rs_temp_elems<u8>(n));
    B.d_count = s->d_scal;
    B.h_count = s->h_scal;
    if (!B.temp) return cudaErrorMemoryAllocation;
    return unbwt(s->stream, d_in, n, idx, d_out, B, status);
}

// Entropy stage: one thread block per block of data (cm.cuh).
cudaError_t run_cm_encode(bz3_state* s, const u8* d_in, s32 n, u8* d_out, s32* out_size) {
    s32* d_res = reinterpret_cast<s32*>(s->d_scal + 12);
    BZ_LAUNCH(1, kCmEncThreads, kCmEncSmemBytes, s->stream, cm_encode_kernel)(d_in, n, d_out, d_res);
    BZ_NOTE_LAUNCH();
    BZ_CUDA_TRY(cudaGetLastError());
    // Nothing is queued behind a long single-CTA kernel: with more streams than hardware queues (32 at most,
    // CUDA_DEVICE_MAX_CONNECTIONS) a copy or launch waiting in a queue for THIS block's coder would hold up the
    // launches of every other block that shares the queue -- and blocks in flight are the throughput of this library.
    BZ_CUDA_TRY(cudaStreamSynchronize(s->stream));
    BZ_CUDA_TRY(cudaMemcpyAsync(s->h_scal + 12, d_res, 4, cudaMemcpyDeviceToHost, s->stream));
    BZ_CUDA_TRY(cudaStreamSynchronize(s->stream));
    *out_size = (s32)s->h_scal[12];
    return cudaSuccess;
}

cudaError_t run_cm_decode(bz3_state* s, const u8* d_in, s32 insize, u8* d_out, s32 n) {
    BZ_LAUNCH(1, kCmDecThreads, kCmDecSmemBytes, s->stream, cm_decode_kernel)(d_in, insize, d_out, n);
    BZ_NOTE_LAUNCH();
    BZ_CUDA_TRY(cudaGetLastError());
    return cudaStreamSynchronize(s->stream);   // see run_cm_encode: the inverse BWT's launches must not queue behind the decoder
}

// ---------------------------------------------------------------------------------- block encode
struct EncodeResult {
    u32 crc;
    s32 bwt_idx, lzp_size, rle_size, payload;
    int model;
    int payload_buf;  // d_buf index holding the payload
};

// input: `size` bytes in d_buf[in_buf] (size >= 64).  Fills R; returns BZ3 error code.
int encode_core(bz3_state* s, int in_buf, s32 size, EncodeResult& R) {
    int cur = in_buf, other = (in_buf + 1) % 3, third = (in_buf + 2) % 3;
    s32 cur_size = size;
    R.model = 0;
    R.lzp_size = R.rle_size = -1;
    {
        Timer t(s, BZ3_STAGE_CRC);
        if (run_crc(s, s->d_buf[cur], (u32)size, &R.crc) != cudaSuccess) return BZ3_ERR_INIT;
    }
    {
        Timer t(s, BZ3_STAGE_RLE);
        if (run_rle_encode(s, s->d_buf[cur], (u32)cur_size, s->d_buf[other], &R.rle_size) != cudaSuccess) return BZ3_ERR_INIT;
    }
    if (R.rle_size < cur_size) {  // src/libbz3.c:610
        int t = cur; cur = other; other = t;
        cur_size = R.rle_size;
        R.model |= 4;
    }
    {
        Timer t(s, BZ3_STAGE_LZP);
        if (run_lzp_encode(s, s->d_buf[cur], cur_size, s->d_buf[other], &R.lzp_size) != cudaSuccess) return BZ3_ERR_INIT;
    }
    if (R.lzp_size > 0 && R.lzp_size < cur_size) {  // :617
        int t = cur; cur = other; other = t;
        cur_size = R.lzp_size;
        R.model |= 2;
    }
    {
        Timer t(s, BZ3_STAGE_BWT);
        if (run_bwt(s, s->d_buf[cur], (u32)cur_size, s->d_buf[other], &R.bwt_idx) != cudaSuccess) return BZ3_ERR_BWT;
    }
    if (R.bwt_idx < 0) return BZ3_ERR_BWT;
    R.payload_buf = third;
    {
        Timer t(s, BZ3_STAGE_CM);
        if (run_cm_encode(s, s->d_buf[other], cur_size, s->d_buf[third], &R.payload) != cudaSuccess) return BZ3_ERR_INIT;
    }
    return BZ3_OK;
}

int header_bytes(int model) { return 9 + ((model & 2) ? 4 : 0) + ((model & 4) ? 4 : 0); }

void write_header(u8* p, const EncodeResult& R) {  // :641-647
    put32(p, R.crc);
    put32(p + 4, (u32)R.bwt_idx);
    p[8] = (u8)R.model;
    int at = 9;
    if (R.model & 2) { put32(p + at, (u32)R.lzp_size); at += 4; }
    if (R.model & 4) { put32(p + at, (u32)R.rle_size); at += 4; }
}

// ---------------------------------------------------------------------------------- block decode
struct DecodeHeader {
    u32 crc;
    s32 bwt_idx, lzp_size, rle_size, n, payload;
    int model, hdr;
};

// Validation that needs only the first bytes of the block (reference :658-737).  Returns 1 when the
// block is a raw (<64 byte) block, 0 for a coded block, or a negative error.
int parse_header(bz3_state* s, const u8* head, size_t head_avail, size_t buffer_size, s32 compressed_size,
                 s32 orig_size, DecodeHeader& H) {
    (void)head_avail;
    const s64 bound = (s64)block_bound((size_t)s->block_size);
    if (buffer_size < 9 || buffer_size < (size_t)(s64)compressed_size) return BZ3_ERR_DATA_SIZE_TOO_SMALL;
    H.crc = get32(head);
    H.bwt_idx = (s32)get32(head + 4);
    if (compressed_size < 0 || (s64)compressed_size > bound) return BZ3_ERR_MALFORMED_HEADER;
    if (H.bwt_idx == -1) {
        if (compressed_size - 8 > 64 || compressed_size < 8) return BZ3_ERR_MALFORMED_HEADER;
        if ((size_t)(compressed_size - 8) > buffer_size) return BZ3_ERR_DATA_SIZE_TOO_SMALL;
        return 1;
    }
    H.model = (s8)head[8];
    size_t need = 9 + (size_t)((H.model & 2) * 4) + (size_t)((H.model & 4) * 4);  // sic, :697
    if (buffer_size < need) return BZ3_ERR_DATA_SIZE_TOO_SMALL;
    H.lzp_size = H.rle_size = -1;
    int at = 9;
    if (H.model & 2) { H.lzp_size = (s32)get32(head + at); at += 4; }
    if (H.model & 4) { H.rle_size = (s32)get32(head + at); at += 4; }
    H.hdr = at;
    H.payload = compressed_size - at;
    if (((H.model & 2) && (H.lzp_size < 0 || H.lzp_size > bound)) || ((H.model & 4) && (H.rle_size < 0 || H.rle_size > bound)))
        return BZ3_ERR_MALFORMED_HEADER;
    if (orig_size < 0 || orig_size > bound) return BZ3_ERR_MALFORMED_HEADER;
    H.n = (H.model & 2) ? H.lzp_size : (H.model & 4) ? H.rle_size : orig_size;
    size_t l = H.lzp_size < 0 ? 0 : (size_t)H.lzp_size, r = H.rle_size < 0 ? 0 : (size_t)H.rle_size;
    if (l > buffer_size || r > buffer_size || (size_t)orig_size > buffer_size) return BZ3_ERR_DATA_SIZE_TOO_SMALL;
    return 0;
}

// payload: H.payload bytes at d_buf[pay_buf] + pay_off.  On success *out_buf holds *out_size bytes and
// *crc_ok tells whether the checksum matched.  Mirrors :739-809; sets nothing on the state.
int decode_core(bz3_state* s, int pay_buf, size_t pay_off, const DecodeHeader& H, size_t buffer_size, s32 orig_size,
                int* out_buf, s32* out_size, bool* crc_ok) {
    int a = (pay_buf + 1) % 3, b = (pay_buf + 2) % 3;
    const s32 bound = (s32)block_bound((size_t)s->block_size);
    {
        Timer t(s, BZ3_STAGE_CM);
        if (run_cm_decode(s, s->d_buf[pay_buf] + pay_off, H.payload, s->d_buf[a], H.n) != cudaSuccess) return BZ3_ERR_INIT;
    }
    if (H.bwt_idx > H.n) return BZ3_ERR_MALFORMED_HEADER;  // :750
    {
        Timer t(s, BZ3_STAGE_BWT);
        int status = 0;
        if (run_unbwt(s, s->d_buf[a], (u32)H.n, H.bwt_idx, s->d_buf[b], &status) != cudaSuccess) return BZ3_ERR_INIT;
        if (status < 0) return BZ3_ERR_BWT;
    }
    int cur = b, other = a;
    s32 cur_size = H.n;
    if (H.model & 2) {
        Timer t(s, BZ3_STAGE_LZP);
        s32 r = -1;
        if (run_lzp_decode(s, s->d_buf[cur], H.lzp_size, s->d_buf[other], bound, &r) != cudaSuccess) return BZ3_ERR_INIT;
        if (r == -1) return BZ3_ERR_CRC;                                   // :769
        if ((size_t)r > buffer_size) return BZ3_ERR_DATA_SIZE_TOO_SMALL;    // :776
        cur_size = r;
        int t2 = cur; cur = other; other = t2;
    }
    if (H.model & 4) {
        Timer t(s, BZ3_STAGE_RLE);
        int err = 0;
        if (run_rle_decode(s, s->d_buf[cur], (u32)cur_size, s->d_buf[other], (u32)orig_size, &err) != cudaSuccess)
            return BZ3_ERR_INIT;
        if (err) return BZ3_ERR_CRC;  // :786
        cur_size = orig_size;
        int t2 = cur; cur = other; other = t2;
    }
    if (cur_size > s->block_size || cur_size < 0) return BZ3_ERR_MALFORMED_HEADER;  // :796
    {
        Timer t(s, BZ3_STAGE_CRC);
        u32 crc = 0;
        if (run_crc(s, s->d_buf[cur], (u32)cur_size, &crc) != cudaSuccess) return BZ3_ERR_INIT;
        *crc_ok = crc == H.crc;
    }
    *out_buf = cur;
    *out_size = cur_size;
    return BZ3_OK;
}

bool use_device(bz3_state* s) { return cudaSetDevice(s->device) == cudaSuccess; }

// Where bz3_new() puts a state.  Default: the calling thread's current device (one process per GPU, as bench.py runs).
// With BZ3_B200_DEVICES=N|all, or after bz3_b200_set_devices(N), states are dealt round-robin over N visible GPUs
// starting at the current one: state i -> GPU i mod N -- the reference's "n states, n threads" batch
// (bz3_encode_blocks, src/libbz3.c:845-856; src/main.c:336-363) then runs block i on GPU i mod N with nothing
// exchanged between the devices (SURVEY 8b "GPU mapping", 8e).
std::atomic<int> g_devices{0};        // 0: not decided yet
std::atomic<unsigned> g_next_device{0};
int visible_devices() {
    int n = 0;
    return (cudaGetDeviceCount(&n) == cudaSuccess && n > 0) ? n : 0;
}
int placement_devices() {
    int d = g_devices.load();
    if (d > 0) return d;
    d = 1;
    if (const char* v = getenv("BZ3_B200_DEVICES")) {
        const int vis = visible_devices();
        d = (v[0] == 'a' || v[0] == 'A') ? vis : atoi(v);
        d = d < 1 ? 1 : (vis > 0 && d > vis ? vis : d);
    }
    g_devices.store(d);
    return d;
}

// the calling thread's current device is left as it was found by the entry points of the reference ABI
struct DeviceScope {
    int before = -1;
    explicit DeviceScope(int dev) {
        if (cudaGetDevice(&before) != cudaSuccess) before = -1;
        if (before != dev) cudaSetDevice(dev); else before = -1;
    }
    ~DeviceScope() { if (before >= 0) cudaSetDevice(before); }
};

}  // namespace

// =================================================================================== public ABI
extern "C" {

BZIP3_API const char* bz3_version(void) { return BZ3_VERSION_STRING; }
BZIP3_API int8_t bz3_last_error(struct bz3_state* state) { return state->last_error; }
BZIP3_API size_t bz3_bound(size_t input_size) { return block_bound(input_size); }

BZIP3_API const char* bz3_strerror(struct bz3_state* state) {  // reference src/libbz3.c:512-533
    switch (state->last_error) {
        case BZ3_OK: return "No error";
        case BZ3_ERR_OUT_OF_BOUNDS: return "Data index out of bounds";
        case BZ3_ERR_BWT: return "Burrows-Wheeler transform failed";
        case BZ3_ERR_CRC: return "CRC32 check failed";
        case BZ3_ERR_MALFORMED_HEADER: return "Malformed header";
        case BZ3_ERR_TRUNCATED_DATA: return "Truncated data";
        case BZ3_ERR_DATA_TOO_BIG: return "Too much data";
        case BZ3_ERR_DATA_SIZE_TOO_SMALL:
            return "Size of buffer `buffer_size` passed to the block decoder (bz3_decode_block) is too small. See "
                   "function docs for details.";
        default: return "Unknown error";
    }
}

namespace {

// Per-device constants (CRC tables in constant memory, the kernels' shared-memory opt-in): once per device, not once per
// state -- states are also created while other blocks are running (stream.h), and there is no reason to rewrite a
// constant bank that running kernels read.
bool device_setup(int dev) {
    static std::once_flag once[kMaxDevices];
    static bool good[kMaxDevices];
    std::call_once(once[dev], [dev] { good[dev] = crc_upload_tables() == cudaSuccess && cm_set_smem_attrs() == cudaSuccess; });
    return good[dev];
}

}  // namespace

BZIP3_API struct bz3_state* bz3_new(int32_t block_size) {
    if (block_size < 65 * 1024 || block_size > 511 * 1024 * 1024) return nullptr;  // :536
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) {
        fprintf(stderr, "[bz3_b200] no CUDA device: this library has no CPU path\n");
        return nullptr;
    }
    if (const int nd = placement_devices(); nd > 1) {   // deal the states over the GPUs (see placement_devices)
        const int vis = visible_devices();
        if (vis > 1) dev = (dev + (int)(g_next_device.fetch_add(1) % (unsigned)nd)) % vis;
        if (dev >= kMaxDevices) dev = 0;
    }
    DeviceScope scope(dev);
    bz3_state* s = new (std::nothrow) bz3_state();
    if (!s) return nullptr;
    memset(static_cast<void*>(s), 0, sizeof(*s));   // every field of the state is plain data
    s->block_size = block_size;
    s->device = dev;
    s->last_error = BZ3_OK;
    const size_t n = block_bound((size_t)block_size) + 64;
    s->cap = align_up(n + 256);
    bool ok = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && device_setup(dev);
    size_t total = 0;
    for (int i = 0; i < 3 && ok; i++) {
        ok = cudaMalloc(&s->d_buf[i], s->cap) == cudaSuccess;
        total += s->cap;
    }
    ok = ok && cudaMalloc(&s->d_lut, sizeof(s32) * kLzpSlots) == cudaSuccess;
    ok = ok && cudaMalloc(&s->d_scal, 256 * sizeof(u32)) == cudaSuccess;
    ok = ok && cudaMallocHost(&s->h_scal, 1024 * sizeof(u32)) == cudaSuccess;
    total += sizeof(s32) * kLzpSlots;
    if (ok) {   // the stage workspace is shared by the states of the device (ArenaPool above)
        size_t a = sufsort_arena_bytes(n), b = other_arena_bytes(n);
        s->arena_need = a > b ? a : b;
        ok = s->pool_attached = pool_attach(dev, s->arena_need);
    }
    for (int i = 0; i < BZ3_STAGE_COUNT && ok; i++)
        ok = cudaEventCreate(&s->clk.a[i]) == cudaSuccess && cudaEventCreate(&s->clk.b[i]) == cudaSuccess;
    for (int i = 0; i < 80 && ok; i++) ok = cudaEventCreate(&s->sort_ev[i]) == cudaSuccess;
    s->device_bytes = total;
    if (!ok) {
        fprintf(stderr, "[bz3_b200] bz3_new(%d): device setup failed: %s\n", block_size,
                cudaGetErrorString(cudaGetLastError()));
        bz3_free(s);
        return nullptr;
    }
    return s;
}

BZIP3_API void bz3_free(struct bz3_state* s) {
    if (!s) return;
    int caller_device = -1;   // a state may live on another device than the calling thread's current one (bz3_b200_*_fd2)
    if (cudaGetDevice(&caller_device) != cudaSuccess) caller_device = -1;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    for (int i = 0; i < 3; i++)
        if (s->d_buf[i]) cudaFree(s->d_buf[i]);
    if (s->d_lut) cudaFree(s->d_lut);
    if (s->d_scal) cudaFree(s->d_scal);
    if (s->h_scal) cudaFreeHost(s->h_scal);
    if (s->pool_attached) pool_detach(s->device);
    for (int i = 0; i < BZ3_STAGE_COUNT; i++) {
        if (s->clk.a[i]) cudaEventDestroy(s->clk.a[i]);
        if (s->clk.b[i]) cudaEventDestroy(s->clk.b[i]);
    }
    for (int i = 0; i < 80; i++)
        if (s->sort_ev[i]) cudaEventDestroy(s->sort_ev[i]);
    if (s->stream) cudaStreamDestroy(s->stream);
    if (caller_device >= 0 && caller_device != s->device) cudaSetDevice(caller_device);
    delete s;
}

BZIP3_API size_t bz3_min_memory_needed(int32_t block_size) {  // host-equivalent figure of the reference, :999-1022
    if (block_size < 65 * 1024 || block_size > 511 * 1024 * 1024) return 0;
    // sizeof(struct bz3_state) = 48 and sizeof(state) = 149024 in the reference on LP64 (five pointers + s32 + s8,
    // padded; 148992 bytes of counters + two pointers + four 32-bit fields): checked against the compiled reference
    // in tests/test_abi.py
    size_t total = 48 + 149024;
    total += block_bound((size_t)block_size);
    total += (block_bound((size_t)block_size) + 128) * sizeof(s32);
    total += (size_t)kLzpSlots * sizeof(s32);
    return total;
}

// ------------------------------------------------------------------ resident (device) operation
BZIP3_API int bz3_b200_upload(struct bz3_state* s, const uint8_t* host, int32_t size) {
    if (!use_device(s) || size < 0 || (size_t)size > s->cap - 64) return -1;
    Timer t(s, BZ3_STAGE_H2D);
    if (cudaMemcpyAsync(s->d_buf[0], host, (size_t)size, cudaMemcpyHostToDevice, s->stream) != cudaSuccess) return -1;
    s->resident_buf = 0;
    s->resident_size = size;
    return 0;
}

BZIP3_API int bz3_b200_download(struct bz3_state* s, uint8_t* host, int32_t size) {
    if (!use_device(s) || size < 0 || size > s->resident_size) return -1;
    Timer t(s, BZ3_STAGE_D2H);
    if (cudaMemcpyAsync(host, s->d_buf[s->resident_buf], (size_t)size, cudaMemcpyDeviceToHost, s->stream) != cudaSuccess)
        return -1;
    return cudaStreamSynchronize(s->stream) == cudaSuccess ? 0 : -1;
}

// Encodes the resident data; the encoded block (header + payload) becomes the resident data.
BZIP3_API int32_t bz3_b200_encode_resident(struct bz3_state* s, int32_t size) {
    if (!use_device(s)) { s->last_error = BZ3_ERR_INIT; return -1; }
    LaunchScope ls(s);
    if (size > s->block_size) { s->last_error = BZ3_ERR_DATA_TOO_BIG; return -1; }  // :588
    clocks_begin(s);
    const int in_buf = s->resident_buf;
    if (size < 64) {  // raw block, :596-601 (last_error untouched)
        u32 crc = 0;
        { Timer t(s, BZ3_STAGE_CRC); if (run_crc(s, s->d_buf[in_buf], (u32)size, &crc) != cudaSuccess) { s->last_error = BZ3_ERR_INIT; return -1; } }
        int ob = (in_buf + 1) % 3;
        u8 hdr[8];
        put32(hdr, crc);
        put32(hdr + 4, 0xFFFFFFFFu);
        cudaMemcpyAsync(s->d_buf[ob], hdr, 8, cudaMemcpyHostToDevice, s->stream);
        cudaMemcpyAsync(s->d_buf[ob] + 8, s->d_buf[in_buf], (size_t)size, cudaMemcpyDeviceToDevice, s->stream);
        s->resident_buf = ob;
        s->resident_size = size + 8;
        clocks_collect(s, 0);
        return size + 8;
    }
    EncodeResult R;
    int e = encode_core(s, in_buf, size, R);
    if (e != BZ3_OK) { s->last_error = (s8)e; clocks_collect(s, 0); return -1; }
    // assemble header in front of the payload: payload buffer has 32 spare bytes? no: copy into a free buffer
    const int hb = header_bytes(R.model);
    int ob = (R.payload_buf + 1) % 3;
    u8 hdr[17];
    write_header(hdr, R);
    cudaMemcpyAsync(s->d_buf[ob], hdr, (size_t)hb, cudaMemcpyHostToDevice, s->stream);
    cudaMemcpyAsync(s->d_buf[ob] + hb, s->d_buf[R.payload_buf], (size_t)R.payload, cudaMemcpyDeviceToDevice, s->stream);
    s->resident_buf = ob;
    s->resident_size = R.payload + hb;
    s->last_error = BZ3_OK;
    clocks_collect(s, 0);
    return R.payload + hb;
}

BZIP3_API int32_t bz3_b200_decode_resident(struct bz3_state* s, int32_t compressed_size, int32_t orig_size) {
    if (!use_device(s)) { s->last_error = BZ3_ERR_INIT; return -1; }
    LaunchScope ls(s);
    clocks_begin(s);
    const int in_buf = s->resident_buf;
    const size_t buffer_size = s->cap - 64;
    u8 head[20];
    memset(head, 0, sizeof head);
    size_t hn = compressed_size < 0 ? 0 : ((size_t)compressed_size < sizeof head ? (size_t)compressed_size : sizeof head);
    if (hn) {
        cudaMemcpyAsync(head, s->d_buf[in_buf], hn, cudaMemcpyDeviceToHost, s->stream);
        cudaStreamSynchronize(s->stream);
    }
    DecodeHeader H;
    int k = parse_header(s, head, hn, buffer_size, compressed_size, orig_size, H);
    if (k < 0) { s->last_error = (s8)k; return -1; }
    if (k == 1) {
        const s32 len = compressed_size - 8;
        int ob = (in_buf + 1) % 3;
        cudaMemcpyAsync(s->d_buf[ob], s->d_buf[in_buf] + 8, (size_t)len, cudaMemcpyDeviceToDevice, s->stream);
        u32 crc = 0;
        if (run_crc(s, s->d_buf[ob], (u32)len, &crc) != cudaSuccess) { s->last_error = BZ3_ERR_INIT; return -1; }
        s->resident_buf = ob;
        s->resident_size = len;
        if (crc != H.crc) { s->last_error = BZ3_ERR_CRC; return -1; }
        return len;  // last_error untouched, :691
    }
    int ob = in_buf;
    s32 osz = 0;
    bool crc_ok = false;
    int e = decode_core(s, in_buf, (size_t)H.hdr, H, buffer_size, orig_size, &ob, &osz, &crc_ok);
    clocks_collect(s, 1);
    if (e != BZ3_OK) { s->last_error = (s8)e; return -1; }
    s->resident_buf = ob;
    s->resident_size = osz;
    if (!crc_ok) { s->last_error = BZ3_ERR_CRC; return -1; }
    s->last_error = BZ3_OK;
    return osz;
}

// ------------------------------------------------------------------ reference block API (host buffers)
BZIP3_API int32_t bz3_encode_block(struct bz3_state* s, uint8_t* buffer, int32_t size) {
    DeviceScope scope(s->device);
    if (size > s->block_size) { s->last_error = BZ3_ERR_DATA_TOO_BIG; return -1; }
    if (size < 0 || bz3_b200_upload(s, buffer, size) != 0) { s->last_error = BZ3_ERR_INIT; return -1; }
    int32_t r = bz3_b200_encode_resident(s, size);
    if (r < 0) return -1;
    if (bz3_b200_download(s, buffer, r) != 0) { s->last_error = BZ3_ERR_INIT; return -1; }
    clocks_collect(s, 0);
    return r;
}

BZIP3_API int32_t bz3_decode_block(struct bz3_state* s, uint8_t* buffer, size_t buffer_size, int32_t compressed_size,
                                   int32_t orig_size) {
    DeviceScope scope(s->device);
    if (!use_device(s)) { s->last_error = BZ3_ERR_INIT; return -1; }
    LaunchScope ls(s);
    clocks_begin(s);
    DecodeHeader H;
    // the reference reads header bytes straight from the caller's buffer after the size checks
    if (buffer_size < 9 || buffer_size < (size_t)(s64)compressed_size) { s->last_error = BZ3_ERR_DATA_SIZE_TOO_SMALL; return -1; }
    u8 head[20];
    memset(head, 0, sizeof head);
    memcpy(head, buffer, buffer_size < sizeof head ? buffer_size : sizeof head);
    int k = parse_header(s, head, sizeof head, buffer_size, compressed_size, orig_size, H);
    if (k < 0) { s->last_error = (s8)k; return -1; }
    if (k == 1) {  // raw block :672-692
        const s32 len = compressed_size - 8;
        memmove(buffer, buffer + 8, (size_t)len);
        u32 crc = 0;
        if (cudaMemcpyAsync(s->d_buf[0], buffer, (size_t)len, cudaMemcpyHostToDevice, s->stream) != cudaSuccess ||
            run_crc(s, s->d_buf[0], (u32)len, &crc) != cudaSuccess) { s->last_error = BZ3_ERR_INIT; return -1; }
        if (crc != H.crc) { s->last_error = BZ3_ERR_CRC; return -1; }
        return len;
    }
    // stage only the payload
    const s32 pay = H.payload > 0 ? H.payload : 0;
    {
        Timer t(s, BZ3_STAGE_H2D);
        if (pay && cudaMemcpyAsync(s->d_buf[0], buffer + H.hdr, (size_t)pay, cudaMemcpyHostToDevice, s->stream) != cudaSuccess) {
            s->last_error = BZ3_ERR_INIT; return -1;
        }
    }
    int ob = 0;
    s32 osz = 0;
    bool crc_ok = false;
    int e = decode_core(s, 0, 0, H, buffer_size, orig_size, &ob, &osz, &crc_ok);
    if (e != BZ3_OK) { s->last_error = (s8)e; clocks_collect(s, 1); return -1; }
    s->last_error = BZ3_OK;
    {
        Timer t(s, BZ3_STAGE_D2H);
        if (osz && cudaMemcpyAsync(buffer, s->d_buf[ob], (size_t)osz, cudaMemcpyDeviceToHost, s->stream) != cudaSuccess) {
            s->last_error = BZ3_ERR_INIT; return -1;
        }
    }
    clocks_collect(s, 1);
    s->resident_buf = ob;
    s->resident_size = osz;
    if (!crc_ok) { s->last_error = BZ3_ERR_CRC; return -1; }  // :803 (output already copied back, like the reference)
    return osz;
}

// ------------------------------------------------------------------ batch API: one thread + stream per block
BZIP3_API void bz3_encode_blocks(struct bz3_state* states[], uint8_t* buffers[], int32_t sizes[], int32_t n) {
    std::vector<std::thread> th;
    std::vector<int32_t> res((size_t)(n > 0 ? n : 0));
    for (int32_t i = 0; i < n; i++)
        th.emplace_back([&, i] { res[(size_t)i] = bz3_encode_block(states[i], buffers[i], sizes[i]); });
    for (auto& t : th) t.join();
    for (int32_t i = 0; i < n; i++) sizes[i] = res[(size_t)i];
}

BZIP3_API void bz3_decode_blocks(struct bz3_state* states[], uint8_t* buffers[], size_t buffer_sizes[], int32_t sizes[],
                                 int32_t orig_sizes[], int32_t n) {
    std::vector<std::thread> th;
    for (int32_t i = 0; i < n; i++)
        th.emplace_back([&, i] { bz3_decode_block(states[i], buffers[i], buffer_sizes[i], sizes[i], orig_sizes[i]); });
    for (auto& t : th) t.join();
}

BZIP3_API void bz3_b200_encode_resident_many(struct bz3_state* states[], int32_t sizes[], int32_t results[], int32_t n) {
    std::vector<std::thread> th;
    for (int32_t i = 0; i < n; i++) th.emplace_back([&, i] { results[i] = bz3_b200_encode_resident(states[i], sizes[i]); });
    for (auto& t : th) t.join();
}
BZIP3_API void bz3_b200_decode_resident_many(struct bz3_state* states[], int32_t csizes[], int32_t osizes[],
                                             int32_t results[], int32_t n) {
    std::vector<std::thread> th;
    for (int32_t i = 0; i < n; i++)
        th.emplace_back([&, i] { results[i] = bz3_b200_decode_resident(states[i], csizes[i], osizes[i]); });
    for (auto& t : th) t.join();
}

// ------------------------------------------------------------------ frame API (reference :876-997)
BZIP3_API int bz3_compress(uint32_t block_size, const uint8_t* in, uint8_t* out, size_t in_size, size_t* out_size) {
    if (block_size > in_size) block_size = (uint32_t)block_bound(in_size);
    if (block_size <= 65 * 1024) block_size = 65 * 1024;
    bz3_state* s = bz3_new((int32_t)block_size);
    if (!s) return BZ3_ERR_INIT;
    std::vector<u8> scratch;
    scratch.resize(block_bound(block_size));
    const size_t buf_max = *out_size;
    *out_size = 0;
    u32 n_blocks = (u32)(in_size / block_size);
    if (in_size % block_size) n_blocks++;
    if (buf_max < 13 || buf_max < block_bound(in_size)) { bz3_free(s); return BZ3_ERR_DATA_TOO_BIG; }
    memcpy(out, "BZ3v1", 5);
    put32(out + 5, block_size);
    put32(out + 9, n_blocks);
    *out_size = 13;
    size_t in_off = 0;
    for (u32 i = 0; i < n_blocks; i++) {
        // like the reference, the last block takes in_size % block_size bytes -- 0 when the input is an exact
        // multiple of the block size (known reference behaviour; kept for format parity)
        s32 size = (s32)block_size;
        if (i == n_blocks - 1) size = (s32)(in_size % block_size);
        memcpy(scratch.data(), in + in_off, (size_t)size);
        s32 enc = bz3_encode_block(s, scratch.data(), size);
        if (bz3_last_error(s) != BZ3_OK) { int e = s->last_error; bz3_free(s); return e; }
        memcpy(out + *out_size + 8, scratch.data(), (size_t)enc);
        put32(out + *out_size, (u32)enc);
        put32(out + *out_size + 4, (u32)size);
        *out_size += (size_t)enc + 8;
        in_off += (size_t)size;
    }
    bz3_free(s);
    return BZ3_OK;
}

BZIP3_API int bz3_decompress(const uint8_t* in, uint8_t* out, size_t in_size, size_t* out_size) {
    if (in_size < 13) return BZ3_ERR_MALFORMED_HEADER;
    if (memcmp(in, "BZ3v1", 5) != 0) return BZ3_ERR_MALFORMED_HEADER;
    u32 block_size = get32(in + 5);
    u32 n_blocks = get32(in + 9);
    in_size -= 13;
    in += 13;
    bz3_state* s = bz3_new((int32_t)block_size);
    if (!s) return BZ3_ERR_INIT;
    const size_t cap = block_bound(block_size);
    std::vector<u8> scratch(cap);
    const size_t buf_max = *out_size;
    *out_size = 0;
    for (u32 i = 0; i < n_blocks; i++) {
        if (in_size < 8) { bz3_free(s); return BZ3_ERR_MALFORMED_HEADER; }
        s32 size = (s32)get32(in);
        if (size < 0 || (u32)size > block_size) { bz3_free(s); return BZ3_ERR_MALFORMED_HEADER; }
        if (in_size < (size_t)size + 8) { bz3_free(s); return BZ3_ERR_TRUNCATED_DATA; }
        s32 orig = (s32)get32(in + 4);
        if (orig < 0) { bz3_free(s); return BZ3_ERR_MALFORMED_HEADER; }
        if (buf_max < *out_size + (size_t)orig) { bz3_free(s); return BZ3_ERR_DATA_TOO_BIG; }
        memcpy(scratch.data(), in + 8, (size_t)size);
        bz3_decode_block(s, scratch.data(), cap, size, orig);
        if (bz3_last_error(s) != BZ3_OK) { int e = s->last_error; bz3_free(s); return e; }
        memcpy(out + *out_size, scratch.data(), (size_t)orig);
        *out_size += (size_t)orig;
        in += size + 8;
        in_size -= (size_t)size + 8;
    }
    bz3_free(s);
    return BZ3_OK;
}

BZIP3_API int bz3_orig_size_sufficient_for_decode(const uint8_t* block, size_t block_size, int32_t orig_size) {  // :1025-1055
    if (block_size < 9) return -1;
    if ((s32)get32(block + 4) == -1) return 1;
    int model = (s8)block[8];
    size_t need = 9 + (size_t)((model & 2) * 4) + (size_t)((model & 4) * 4);
    if (block_size < need) return -1;
    s32 lzp = -1, rle = -1;
    size_t at = 9;
    if (model & 2) { lzp = (s32)get32(block + at); at += 4; }
    if (model & 4) rle = (s32)get32(block + at);
    size_t bs = (size_t)orig_size;
    size_t l = lzp < 0 ? 0 : (size_t)lzp, r = rle < 0 ? 0 : (size_t)rle, o = orig_size < 0 ? 0 : (size_t)orig_size;
    return (l <= bs) && (r <= bs) && (o <= bs);
}

// ------------------------------------------------------------------ introspection / statistics
BZIP3_API int bz3_b200_device_count(void) {
    int n = 0;
    return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}
BZIP3_API int bz3_b200_state_device(struct bz3_state* s) { return s->device; }
BZIP3_API int bz3_b200_set_devices(int devices) {
    const int vis = visible_devices();
    int d = devices <= 0 ? vis : devices;
    d = d < 1 ? 1 : (vis > 0 && d > vis ? vis : d);
    g_devices.store(d);
    return d;
}
BZIP3_API size_t bz3_b200_device_bytes(struct bz3_state* s) { return s->device_bytes; }
BZIP3_API size_t bz3_b200_workspace_bytes(struct bz3_state* s) {   // the device's shared stage workspaces, all slots
    ArenaPool& P = g_pool[s->device];
    std::lock_guard<std::mutex> lk(P.m);
    size_t total = 0;
    for (int k = 0; k < P.nslots; k++) total += P.slot[k].size;
    return total;
}
BZIP3_API void bz3_b200_stats_reset(struct bz3_state* s) {
    memset(s->stage_ms, 0, sizeof s->stage_ms);
    s->launches = 0;
    s->sort_records = 0;
    s->sort_ms = 0;
}
BZIP3_API double bz3_b200_stage_ms(struct bz3_state* s, int stage, int decode) {
    if (stage < 0 || stage >= BZ3_STAGE_COUNT) return 0.0;
    return s->stage_ms[decode ? 1 : 0][stage];
}
BZIP3_API uint64_t bz3_b200_kernel_launches(struct bz3_state* s) { return s->launches; }
BZIP3_API void bz3_b200_last_sort_stats(struct bz3_state* s, uint64_t* records, int32_t* rounds, double* ms) {
    if (records) *records = s->sort_records;
    if (rounds) *rounds = s->sort_rounds;
    if (ms) *ms = s->sort_ms;
}
#ifdef BZ_CM_PROFILE
extern "C" BZIP3_API void bz3_b200_debug_cm_profile(unsigned long long* out16) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out16, g_cm_prof, sizeof(unsigned long long) * 48);
}
#endif
// ------------------------------------------------------------------ the CLI's container with a deep block queue (stream.h)
namespace {
void* stream_host_alloc(size_t n) {
    void* p = nullptr;
    if (cudaMallocHost(&p, n) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void stream_host_free(void* p) { cudaFreeHost(p); }
void stream_thread_init(int dev) { cudaSetDevice(dev); }   // a worker's state lives on the device the worker was dealt
int stream_device() {
    int dev = 0;
    return cudaGetDevice(&dev) == cudaSuccess ? dev : 0;
}

int stream_depth(int32_t block_size, int in_flight) {
    if (in_flight > 256) in_flight = 256;
    if (in_flight > 0) return in_flight;
    int depth = 64;   // fallback if the device cannot be asked
#if !defined(BZ_EMU)
    // A block spends most of its time in a single-CTA coder kernel whose tables fill an SM's shared memory: one block per
    // SM is the useful depth (148 on a B200; the reference's -j stops at 64, src/main.c:213).  Blocks beyond that only wait.
    int dev = 0, sms = 0;
    size_t free_b = 0, total_b = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0)
        depth = sms;
    const size_t n = block_bound((size_t)block_size) + 64;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
        int slots = env_int("BZ3_B200_ARENAS", 2);   // the workspaces the states of this device will share (pool_attach)
        slots = slots < 1 ? 1 : (slots > kMaxArenaSlots ? kMaxArenaSlots : slots);
        const size_t ws = (size_t)slots * std::max(sufsort_arena_bytes(n), other_arena_bytes(n)) + (size_t(1) << 30);
        const size_t per_state = 3 * align_up(n + 256) + (size_t(2) << 20);
        depth = free_b > ws + per_state ? (int)std::min<size_t>((size_t)depth, (free_b - ws) / per_state) : 1;
    }
    depth = (int)std::min<size_t>((size_t)depth, std::max<size_t>(1, (size_t(16) << 30) / n));   // pinned host memory: 16 GiB at most
#else
    (void)block_size;
    depth = 4;
#endif
    return depth < 1 ? 1 : depth;
}

// Workers are dealt round-robin over `devices` GPUs starting with the caller's current one (1: that one only, the
// one-process-per-GPU model of bench.py; <= 0: every visible GPU): block i -> worker i mod in_flight -> GPU, the static
// work queue of SURVEY 8(e) inside one process.  Blocks are independent, so nothing is exchanged between the GPUs.
std::vector<int> stream_worker_devices(int depth, int devices) {
    int count = 1;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count < 1) count = 1;
    if (devices <= 0 || devices > count) devices = count;
    const int first = stream_device();
    std::vector<int> dev((size_t)depth);
    for (int k = 0; k < depth; k++) dev[(size_t)k] = (first + k % devices) % count;
    return dev;
}
}  // namespace

BZIP3_API int bz3_b200_encode_fd2(int in_fd, int out_fd, int32_t block_size, int in_flight, int devices, uint64_t* bytes_in,
                                  uint64_t* bytes_out) {
    if (block_size < 65 * 1024 || block_size > 511 * 1024 * 1024) return BZ3_B200_ERR_BLOCK_SIZE;
    u8 head[9] = {'B', 'Z', '3', 'v', '1'};
    put32(head + 5, (u32)block_size);
    if (out_fd >= 0 && !bz3stream::write_full(out_fd, head, 9)) return BZ3_B200_ERR_IO;
    uint64_t out = 0;
    const int depth = stream_depth(block_size, in_flight);
    const std::vector<int> dev = stream_worker_devices(depth, devices);
    int r = bz3stream::run(in_fd, out_fd, block_size, depth, false, bytes_in, &out, stream_host_alloc, stream_host_free,
                           stream_thread_init, dev.data());
    if (bytes_out) *bytes_out = out + 9;
    return r;
}

BZIP3_API int bz3_b200_decode_fd2(int in_fd, int out_fd, int in_flight, int devices, uint64_t* bytes_in, uint64_t* bytes_out) {
    u8 head[9];
    if (bz3stream::read_full(in_fd, head, 9) != 9 || memcmp(head, "BZ3v1", 5) != 0) return BZ3_B200_ERR_SIGNATURE;
    const s32 block_size = (s32)get32(head + 5);
    if (block_size < 65 * 1024 || block_size > 511 * 1024 * 1024) return BZ3_B200_ERR_BLOCK_SIZE;
    uint64_t in = 0;
    const int depth = stream_depth(block_size, in_flight);
    const std::vector<int> dev = stream_worker_devices(depth, devices);
    int r = bz3stream::run(in_fd, out_fd, block_size, depth, true, &in, bytes_out, stream_host_alloc, stream_host_free,
                           stream_thread_init, dev.data());
    if (bytes_in) *bytes_in = in + 9;
    return r;
}

BZIP3_API int bz3_b200_encode_fd(int in_fd, int out_fd, int32_t block_size, int in_flight, uint64_t* bytes_in, uint64_t* bytes_out) {
    return bz3_b200_encode_fd2(in_fd, out_fd, block_size, in_flight, 1, bytes_in, bytes_out);
}
BZIP3_API int bz3_b200_decode_fd(int in_fd, int out_fd, int in_flight, uint64_t* bytes_in, uint64_t* bytes_out) {
    return bz3_b200_decode_fd2(in_fd, out_fd, in_flight, 1, bytes_in, bytes_out);
}

// ------------------------------------------------------------------ single stages on host buffers (tests)
namespace {
bool stage_in(bz3_state* s, const u8* in, size_t n, int buf) {
    if (!use_device(s) || n > s->cap - 64) return false;
    return cudaMemcpyAsync(s->d_buf[buf], in, n, cudaMemcpyHostToDevice, s->stream) == cudaSuccess;
}
bool stage_out(bz3_state* s, u8* out, size_t n, int buf) {
    if (n && cudaMemcpyAsync(out, s->d_buf[buf], n, cudaMemcpyDeviceToHost, s->stream) != cudaSuccess) return false;
    return cudaStreamSynchronize(s->stream) == cudaSuccess;
}
}  // namespace

BZIP3_API uint32_t bz3_b200_stage_crc(struct bz3_state* s, const uint8_t* in, int32_t n) {
    u32 crc = 0;
    if (!stage_in(s, in, (size_t)n, 0) || run_crc(s, s->d_buf[0], (u32)n, &crc) != cudaSuccess) return 0xDEADBEEF;
    return crc;
}
BZIP3_API int32_t bz3_b200_stage_rle_encode(struct bz3_state* s, const uint8_t* in, int32_t n, uint8_t* out) {
    s32 r = -1;
    if (!stage_in(s, in, (size_t)n, 0) || run_rle_encode(s, s->d_buf[0], (u32)n, s->d_buf[1], &r) != cudaSuccess) return -100;
    if (r >= 0 && (size_t)r <= s->cap && !stage_out(s, out, (size_t)r, 1)) return -100;
    return r;
}
BZIP3_API int bz3_b200_stage_rle_decode(struct bz3_state* s, const uint8_t* in, int32_t maxin, uint8_t* out, int32_t outlen) {
    int err = 1;
    cudaMemsetAsync(s->d_buf[1], 0, (size_t)outlen, s->stream);
    if (!stage_in(s, in, (size_t)maxin, 0) || run_rle_decode(s, s->d_buf[0], (u32)maxin, s->d_buf[1], (u32)outlen, &err) != cudaSuccess)
        return -100;
    if (!stage_out(s, out, (size_t)outlen, 1)) return -100;
    return err;
}
BZIP3_API int32_t bz3_b200_stage_lzp_encode(struct bz3_state* s, const uint8_t* in, int32_t n, uint8_t* out) {
    s32 r = -1;
    if (!stage_in(s, in, (size_t)n, 0) || run_lzp_encode(s, s->d_buf[0], n, s->d_buf[1], &r) != cudaSuccess) return -100;
    if (r > 0 && !stage_out(s, out, (size_t)r, 1)) return -100;
    return r;
}
BZIP3_API int32_t bz3_b200_stage_lzp_decode(struct bz3_state* s, const uint8_t* in, int32_t n, uint8_t* out, int32_t max) {
    s32 r = -1;
    if ((size_t)max > s->cap - 64) return -100;
    if (!stage_in(s, in, (size_t)n, 0) || run_lzp_decode(s, s->d_buf[0], n, s->d_buf[1], max, &r) != cudaSuccess) return -100;
    if (r > 0 && !stage_out(s, out, (size_t)r, 1)) return -100;
    return r;
}
BZIP3_API int32_t bz3_b200_stage_bwt(struct bz3_state* s, const uint8_t* in, int32_t n, uint8_t* out) {
    s32 idx = -1;
    LaunchScope ls(s);
    if (!stage_in(s, in, (size_t)n, 0) || run_bwt(s, s->d_buf[0], (u32)n, s->d_buf[1], &idx) != cudaSuccess) return -100;
    if (!stage_out(s, out, (size_t)n, 1)) return -100;
    return idx;
}
BZIP3_API int32_t bz3_b200_stage_unbwt(struct bz3_state* s, const uint8_t* in, int32_t n, int32_t idx, uint8_t* out) {
    int status = 0;
    if (!stage_in(s, in, (size_t)n, 0) || run_unbwt(s, s->d_buf[0], (u32)n, idx, s->d_buf[1], &status) != cudaSuccess) return -100;
    if (status == 0 && !stage_out(s, out, (size_t)n, 1)) return -100;
    return status;
}
BZIP3_API int32_t bz3_b200_stage_cm_encode(struct bz3_state* s, const uint8_t* in, int32_t n, uint8_t* out) {
    s32 r = -1;
    if (!stage_in(s, in, (size_t)n, 0) || run_cm_encode(s, s->d_buf[0], n, s->d_buf[1], &r) != cudaSuccess) return -100;
    if (r > 0 && !stage_out(s, out, (size_t)r, 1)) return -100;
    return r;
}
BZIP3_API int bz3_b200_stage_cm_decode(struct bz3_state* s, const uint8_t* in, int32_t insize, uint8_t* out, int32_t n) {
    if (!stage_in(s, in, (size_t)(insize > 0 ? insize : 0), 0) || run_cm_decode(s, s->d_buf[0], insize, s->d_buf[1], n) != cudaSuccess)
        return -100;
    return stage_out(s, out, (size_t)n, 1) ? 0 : -100;
}

}  // extern "C"
