// cta_emu.h -- a small CPU emulator for ONE CUDA thread block, used only by the CPU test-suite.
//
// Test infrastructure, not product code: nothing under bzip3_b200/ uses it.  It lets `pytest -m "not gpu"`
// execute the real kernel bodies of bzip3_b200/csrc/*.cuh (compiled by g++ with -DBZ_EMU) against the
// oracle in a container that has no GPU.  Every CUDA thread is a fiber (own stack, hand-written x86-64
// context switch); fibers run cooperatively and switch only at synchronisation points:
//     __syncthreads / __syncthreads_or, named barriers (bar_sync / bar_arrive), warp collectives
//     (__shfl*_sync, __ballot_sync, __any_sync, __all_sync, __match_any_sync, __syncwarp), BZ_SPIN_HINT()
// so a kernel whose cross-thread communication is correctly fenced by those primitives computes here
// exactly what it computes on the device.  What the emulator can NOT show: data races that the
// cooperative schedule happens to hide, memory-ordering bugs, and anything about performance.  To widen
// the net a little the scheduler order can be reversed or randomised per launch (emu::set_schedule).
#pragma once
#if !defined(__x86_64__)
#error "cta_emu.h needs x86-64"
#endif
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <vector>
#include <sys/mman.h>
// system headers that product code includes later must come before the CUDA vocabulary macros below (__noinline__ ...)
#include <dlfcn.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <time.h>
#include <poll.h>
#include <signal.h>
#include <spawn.h>
#include <sys/wait.h>
#include <unistd.h>
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>

// ThreadSanitizer build (tools/emu_tsan.py): every CUDA thread is announced to TSan as a fiber and every barrier /
// warp collective as a release-acquire edge, so a shared- or global-memory access pair of two CUDA threads with no
// barrier, collective or atomic between them is reported as a data race -- a race check without a GPU.
#if defined(__SANITIZE_THREAD__)
extern "C" {
void* __tsan_get_current_fiber(void);
void* __tsan_create_fiber(unsigned flags);
void __tsan_destroy_fiber(void* fiber);
void __tsan_switch_to_fiber(void* fiber, unsigned flags);
void __tsan_acquire(void* addr);
void __tsan_release(void* addr);
void AnnotateIgnoreReadsBegin(const char* file, int line);
void AnnotateIgnoreReadsEnd(const char* file, int line);
void AnnotateIgnoreWritesBegin(const char* file, int line);
void AnnotateIgnoreWritesEnd(const char* file, int line);
}
// the emulator's own bookkeeping (scheduler state, collective slots) is shared by all fibers on purpose
struct BzTsanIgnore {
    BzTsanIgnore() { AnnotateIgnoreReadsBegin(__FILE__, __LINE__); AnnotateIgnoreWritesBegin(__FILE__, __LINE__); }
    ~BzTsanIgnore() { AnnotateIgnoreWritesEnd(__FILE__, __LINE__); AnnotateIgnoreReadsEnd(__FILE__, __LINE__); }
};
#define BZ_TSAN_INTERNAL() BzTsanIgnore _bz_tsan_ignore
#define BZ_TSAN_SWITCH(f) __tsan_switch_to_fiber((f), 1) /* 1 = no implied synchronisation between the fibers */
#define BZ_TSAN_RELEASE(a) __tsan_release((void*)(a))
#define BZ_TSAN_ACQUIRE(a) __tsan_acquire((void*)(a))
#else
#define BZ_TSAN_SWITCH(f) ((void)0)
#define BZ_TSAN_RELEASE(a) ((void)0)
#define BZ_TSAN_ACQUIRE(a) ((void)0)
#define BZ_TSAN_INTERNAL() ((void)0)
#endif

namespace emu {

struct Dim3 {
    unsigned x = 1, y = 1, z = 1;
};

constexpr int kMaxThreads = 1024;
constexpr size_t kStackBytes = 256 * 1024;

struct WarpSlot {
    uint64_t val[32];
    uint32_t arrived = 0, consumed = 0;
};
struct WarpState {
    std::map<uint32_t, WarpSlot> slots[2];  // keyed by participation mask, double buffered
};
struct NamedBarrier {
    unsigned arrived = 0;
    unsigned gen = 0;
    unsigned orv = 0, last_or = 0;
    char tsan_token[2] = {0, 0};   // release/acquire object of the even / odd generations (see bar_sync_impl)
};

struct Fiber {
    void* sp = nullptr;
    char* stack = nullptr;
    Dim3 tid;
    bool done = false;
    void* tsan = nullptr;                // ThreadSanitizer's handle of this fiber
    const void* waiting_on = nullptr;    // barrier / collective slot this fiber sleeps on (not scheduled meanwhile)
    std::map<uint32_t, unsigned> phase;  // per-mask count of warp collectives executed
};

struct Cta {
    std::vector<Fiber> fibers;
    std::vector<WarpState> warps;
    NamedBarrier bars[16];
    unsigned live = 0;
    std::function<void()> body;
    void* main_sp = nullptr;
    void* main_tsan = nullptr;
    int cur = -1;
    unsigned long long switches = 0;
};

extern Cta* g_cta;
extern Fiber* g_cur;
extern Dim3 g_blockIdx, g_blockDim, g_gridDim;
extern unsigned char* g_dyn_smem;
extern char g_tsan_token;   // host -> thread, thread -> host and CTA -> next CTA edges (CTAs of a grid run one after another
                            // here and share the static "shared" arrays, so races BETWEEN CTAs are out of this check's reach)
extern int g_schedule;  // 0 round robin ascending, 1 descending, 2 pseudo random
extern uint64_t g_rng;

extern "C" void emu_switch(void** save_sp, void* load_sp);

inline void set_schedule(int mode, uint64_t seed = 1) {
    g_schedule = mode;
    g_rng = seed * 0x9E3779B97F4A7C15ull + 1;
}

inline int pick_next() {
    Cta& c = *g_cta;
    const int n = (int)c.fibers.size();
    int start = c.cur;
    if (g_schedule == 2) {
        g_rng ^= g_rng << 13;
        g_rng ^= g_rng >> 7;
        g_rng ^= g_rng << 17;
        start = (int)(g_rng % (uint64_t)n);
    }
    for (int k = 1; k <= n; k++) {
        int i = g_schedule == 1 ? ((start - k) % n + n) % n : (start + k) % n;
        if (!c.fibers[i].done && !c.fibers[i].waiting_on) return i;
    }
    return -1;
}

inline void wake_all(const void* obj) {
    for (auto& f : g_cta->fibers)
        if (f.waiting_on == obj) f.waiting_on = nullptr;
}

[[noreturn]] inline void deadlock(const char* what) {
    fprintf(stderr, "[cta_emu] deadlock: every live thread sleeps (%s)\n", what);
    abort();
}

// sleep until somebody calls wake_all(obj)
inline void sleep_on(const void* obj, const char* what);

inline void yield() {
    BZ_TSAN_INTERNAL();
    Cta& c = *g_cta;
    const int nxt = pick_next();
    if (nxt < 0 || nxt == c.cur) return;
    Fiber* from = g_cur;
    c.cur = nxt;
    g_cur = &c.fibers[nxt];
    c.switches++;
    BZ_TSAN_SWITCH(g_cur->tsan);
    emu_switch(&from->sp, g_cur->sp);
}

inline void sleep_on(const void* obj, const char* what) {
    Cta& c = *g_cta;
    g_cur->waiting_on = obj;
    const int nxt = pick_next();
    if (nxt < 0) deadlock(what);
    Fiber* from = g_cur;
    c.cur = nxt;
    g_cur = &c.fibers[nxt];
    c.switches++;
    BZ_TSAN_SWITCH(g_cur->tsan);
    emu_switch(&from->sp, g_cur->sp);
}

[[noreturn]] inline void fiber_exit() {
    void* next_sp;
    void* next_tsan;
    {
        BZ_TSAN_INTERNAL();   // must end before the last switch: this fiber never comes back
        Cta& c = *g_cta;
        BZ_TSAN_RELEASE(&g_tsan_token);
        g_cur->done = true;
        c.live--;
        for (auto& f : c.fibers) f.waiting_on = nullptr;   // an exit can complete a barrier: let sleepers re-check
        const int nxt = pick_next();
        if (nxt < 0) {
            next_sp = c.main_sp;
            next_tsan = c.main_tsan;
        } else {
            c.cur = nxt;
            g_cur = &c.fibers[nxt];
            next_sp = g_cur->sp;
            next_tsan = g_cur->tsan;
        }
    }
    (void)next_tsan;
    void* dummy;
    BZ_TSAN_SWITCH(next_tsan);
    emu_switch(&dummy, next_sp);
    abort();
}

extern "C" inline void emu_fiber_main() {
    BZ_TSAN_ACQUIRE(&g_tsan_token);
    {
        std::function<void()>* body;
        {
            BZ_TSAN_INTERNAL();
            body = &g_cta->body;
        }
        (*body)();
    }
    fiber_exit();
}

// ---- barriers -------------------------------------------------------------------------------------
// bar id 0 with count 0 == __syncthreads: all live (not yet exited) threads.  A watchdog aborts when a
// barrier can never complete (every live thread is waiting on something).
inline unsigned bar_sync_impl(int id, unsigned count, unsigned pred, bool wait) {
    BZ_TSAN_INTERNAL();
    Cta& c = *g_cta;
    NamedBarrier& b = c.bars[id];
    const unsigned gen = b.gen;
    b.arrived++;
    b.orv |= pred;
    // the edge belongs to THIS generation: a thread that already left and arrived at the next barrier must not
    // lend its later writes to a thread that is only now waking up from this one
    BZ_TSAN_RELEASE(&b.tsan_token[gen & 1]);
    for (;;) {
        const unsigned need = count ? count : c.live;
        if (b.gen != gen) break;
        if (b.arrived >= need) {
            b.arrived = 0;
            b.last_or = b.orv;
            b.orv = 0;
            b.gen++;
            wake_all(&b);
            break;
        }
        if (!wait) return 0;
        sleep_on(&b, "barrier");
    }
    BZ_TSAN_ACQUIRE(&b.tsan_token[gen & 1]);
    return b.last_or;
}

inline void syncthreads() { bar_sync_impl(0, 0, 0, true); }
inline int syncthreads_or(int p) { return bar_sync_impl(0, 0, p ? 1u : 0u, true) != 0; }
inline void bar_sync(int id, unsigned count) { bar_sync_impl(id, count, 0, true); }
inline void bar_arrive(int id, unsigned count) { bar_sync_impl(id, count, 0, false); }

// ---- warp collectives -------------------------------------------------------------------------------
inline unsigned lane_of(const Fiber* f) { return f->tid.x & 31u; }
inline unsigned warp_of(const Fiber* f) { return f->tid.x >> 5; }

template <class R>
inline auto warp_collective(uint32_t mask, uint64_t mine, R reader) -> decltype(reader((const uint64_t*)nullptr)) {
    BZ_TSAN_INTERNAL();
    Fiber* f = g_cur;
    const unsigned lane = lane_of(f);
    const uint32_t bit = 1u << lane;
    if (!(mask & bit)) {
        fprintf(stderr, "[cta_emu] lane %u calls a collective with mask %08x that excludes it\n", lane, mask);
        abort();
    }
    const unsigned ph = f->phase[mask]++;
    WarpSlot& s = g_cta->warps[warp_of(f)].slots[ph & 1][mask];
    unsigned long long spins = 0;
    while (s.arrived & bit) {  // slot still in use by the collective two back
        yield();
        if (++spins > 50000000ull) { fprintf(stderr, "[cta_emu] deadlock in warp collective (slot busy)\n"); abort(); }
    }
    s.val[lane] = mine;
    s.arrived |= bit;
    BZ_TSAN_RELEASE(&s);
    if ((s.arrived & mask) == mask) wake_all(&s);
    // lanes of the mask that already exited can never arrive: that is a bug and ends in the deadlock report
    while ((s.arrived & mask) != mask) sleep_on(&s, "warp collective");
    BZ_TSAN_ACQUIRE(&s);
    auto r = reader(s.val);
    s.consumed |= bit;
    if ((s.consumed & mask) == mask) {
        s.arrived &= ~mask;
        s.consumed &= ~mask;
    }
    return r;
}

// ---- launch -------------------------------------------------------------------------------------------
void launch(Dim3 grid, Dim3 block, size_t dyn_smem_bytes, const std::function<void()>& body);

}  // namespace emu

// ======================================================================================================
// CUDA vocabulary for the kernel sources
// ======================================================================================================
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) alignas(n)
#define threadIdx (::emu::g_cur->tid)
#define blockIdx (::emu::g_blockIdx)
#define blockDim (::emu::g_blockDim)
#define gridDim (::emu::g_gridDim)

struct uint2 { uint32_t x, y; };
struct alignas(16) uint4 { uint32_t x, y, z, w; };
inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }

inline void __syncthreads() { ::emu::syncthreads(); }
inline int __syncthreads_or(int p) { return ::emu::syncthreads_or(p); }
inline void __syncwarp(unsigned mask = 0xFFFFFFFFu) {
    ::emu::warp_collective(mask, 0, [](const uint64_t*) { return 0; });
}
inline void __threadfence_block() {}
inline void __threadfence() {}

template <class T>
inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    static_assert(sizeof(T) <= 8, "shfl payload");
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    const unsigned lane = ::emu::lane_of(::emu::g_cur);
    const int base = (int)(lane & ~(unsigned)(width - 1));
    const int s = base + (src & (width - 1));
    uint64_t r = ::emu::warp_collective(mask, raw, [&](const uint64_t* a) { return (mask >> s) & 1u ? a[s] : raw; });
    T out;
    memcpy(&out, &r, sizeof(T));
    return out;
}
template <class T>
inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    const unsigned lane = ::emu::lane_of(::emu::g_cur);
    const int base = (int)(lane & ~(unsigned)(width - 1));
    const int s = (int)lane - (int)delta;
    uint64_t r = ::emu::warp_collective(mask, raw, [&](const uint64_t* a) { return (s >= base && ((mask >> s) & 1u)) ? a[s] : raw; });
    T out;
    memcpy(&out, &r, sizeof(T));
    return out;
}
template <class T>
inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    const unsigned lane = ::emu::lane_of(::emu::g_cur);
    const int base = (int)(lane & ~(unsigned)(width - 1));
    const int s = (int)lane + (int)delta;
    uint64_t r = ::emu::warp_collective(mask, raw, [&](const uint64_t* a) { return (s < base + width && ((mask >> s) & 1u)) ? a[s] : raw; });
    T out;
    memcpy(&out, &r, sizeof(T));
    return out;
}
template <class T>
inline T __shfl_xor_sync(unsigned mask, T v, int x, int width = 32) {
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    const unsigned lane = ::emu::lane_of(::emu::g_cur);
    const int s = (int)(lane ^ (unsigned)x);
    (void)width;
    uint64_t r = ::emu::warp_collective(mask, raw, [&](const uint64_t* a) { return ((mask >> s) & 1u) ? a[s] : raw; });
    T out;
    memcpy(&out, &r, sizeof(T));
    return out;
}
inline unsigned __ballot_sync(unsigned mask, int pred) {
    return ::emu::warp_collective(mask, pred ? 1u : 0u, [&](const uint64_t* a) {
        unsigned r = 0;
        for (int l = 0; l < 32; l++)
            if (((mask >> l) & 1u) && a[l]) r |= 1u << l;
        return r;
    });
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == mask; }
template <class T>
inline unsigned __match_any_sync(unsigned mask, T v) {
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    return ::emu::warp_collective(mask, raw, [&](const uint64_t* a) {
        unsigned r = 0;
        for (int l = 0; l < 32; l++)
            if (((mask >> l) & 1u) && a[l] == raw) r |= 1u << l;
        return r;
    });
}

inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline unsigned __brev(unsigned v) {
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= ((v >> i) & 1u) << (31 - i);
    return r;
}
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s) {
    s &= 31;
    return s ? (hi << s) | (lo >> (32 - s)) : hi;
}
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s) {
    s &= 31;
    return s ? (lo >> s) | (hi << (32 - s)) : lo;
}
inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel) {
    const uint64_t v = ((uint64_t)b << 32) | a;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        const unsigned s = (sel >> (4 * i)) & 0xF;
        unsigned byte = (unsigned)(v >> (8 * (s & 7))) & 0xFF;
        if (s & 8) byte = (byte & 0x80) ? 0xFF : 0x00;
        r |= byte << (8 * i);
    }
    return r;
}
template <class T> inline T __ldcg(const T* p) { return *p; }
template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline void __stcg(T* p, T v) { *p = v; }
template <class T> inline T atomicMax(T* p, T v) {
    T o = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (v > o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return o;
}
template <class T> inline T atomicMin(T* p, T v) {
    T o = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (v < o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return o;
}
template <class T> inline T atomicAdd(T* p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
template <class T> inline T atomicOr(T* p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
template <class T> inline T atomicAnd(T* p, T v) { return __atomic_fetch_and(p, v, __ATOMIC_SEQ_CST); }
template <class T> inline T atomicXor(T* p, T v) { return __atomic_fetch_xor(p, v, __ATOMIC_SEQ_CST); }
template <class T> inline T atomicExch(T* p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
template <class T> inline T atomicCAS(T* p, T c, T v) {
    __atomic_compare_exchange_n(p, &c, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return c;
}
inline long long clock64() { return (long long)::emu::g_cta->switches; }
template <class T> inline T min(T a, T b) { return a < b ? a : b; }
template <class T> inline T max(T a, T b) { return a > b ? a : b; }

// ---- the slice of the CUDA runtime that the host-side launch sequences in bzip3_b200/csrc/*.cuh use -------------
// "Device" memory is host memory here, a stream is synchronous, a launch runs the grid to completion.
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
constexpr cudaError_t cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorUnknown = 999;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) {
    memmove(d, s, n);
    return cudaSuccess;
}
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) {
    memset(d, v, n);
    return cudaSuccess;
}
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
inline const char* cudaGetErrorName(cudaError_t) { return "emu"; }
inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
template <class T>
inline cudaError_t cudaMemcpyToSymbol(T& sym, const void* src, size_t n) {
    memcpy(&sym, src, n);
    return cudaSuccess;
}
template <class F>
inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
// memory, devices, streams, events: one "device", everything synchronous
constexpr unsigned cudaStreamNonBlocking = 1;
// test hook BZ_EMU_MALLOC_TOTAL: a "device" with that many bytes in all (live allocations are tracked)
struct EmuDeviceHeap {
    std::mutex m;
    std::unordered_map<void*, size_t> live;
    size_t used = 0;
    static EmuDeviceHeap& get() { static EmuDeviceHeap h; return h; }
};
template <class T>
inline cudaError_t cudaMalloc(T** p, size_t n) {
    if (const char* cap = getenv("BZ_EMU_MALLOC_MAX"))   // test hook: a "device" that cannot serve larger requests
        if (n > strtoull(cap, nullptr, 10)) { *p = nullptr; return cudaErrorMemoryAllocation; }
    const char* total = getenv("BZ_EMU_MALLOC_TOTAL");
    if (total) {
        EmuDeviceHeap& H = EmuDeviceHeap::get();
        std::lock_guard<std::mutex> lk(H.m);
        if (H.used + n > strtoull(total, nullptr, 10)) { *p = nullptr; return cudaErrorMemoryAllocation; }
        H.used += n;
    }
    *p = static_cast<T*>(aligned_alloc(256, (n + 255) & ~(size_t)255));
    if (*p) memset(*p, 0xCD, n);   // device memory is not zeroed
    if (total && *p) {
        EmuDeviceHeap& H = EmuDeviceHeap::get();
        std::lock_guard<std::mutex> lk(H.m);
        H.live[(void*)*p] = n;
    }
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
template <class T>
inline cudaError_t cudaMallocHost(T** p, size_t n) {
    *p = static_cast<T*>(aligned_alloc(256, (n + 255) & ~(size_t)255));
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
inline cudaError_t cudaFree(void* p) {
    if (p && getenv("BZ_EMU_MALLOC_TOTAL")) {
        EmuDeviceHeap& H = EmuDeviceHeap::get();
        std::lock_guard<std::mutex> lk(H.m);
        auto it = H.live.find(p);
        if (it != H.live.end()) { H.used -= it->second; H.live.erase(it); }
    }
    free(p);
    return cudaSuccess;
}
inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.0f; return cudaSuccess; }
template <class T>
inline cudaError_t cudaMemcpyFromSymbol(void* dst, const T& sym, size_t n) {
    memcpy(dst, &sym, n);
    return cudaSuccess;
}
#define __constant__ static

namespace emu {
// BZ_LAUNCH(grid, block, smem, stream, kernel)(args...)
template <class K>
struct Launcher {
    unsigned grid, block;
    size_t smem;
    K kern;
    template <class... A>
    void operator()(A&&... a) {
        Dim3 g, b;
        g.x = grid;
        b.x = block;
        launch(g, b, smem, [&] { kern(a...); });
    }
};
template <class K>
inline Launcher<K> make_launcher(unsigned grid, unsigned block, size_t smem, K k) { return Launcher<K>{grid, block, smem, k}; }
}  // namespace emu
#define BZ_LAUNCH(grid, block, smem, stream, ...) \
    ::emu::make_launcher((unsigned)(grid), (unsigned)(block), (size_t)(smem), [&](auto&&... _a) { __VA_ARGS__(_a...); })
#define BZ_CUDA_TRY(expr)                  \
    do {                                   \
        cudaError_t _e = (expr);           \
        if (_e != cudaSuccess) return _e;  \
    } while (0)

// dynamic shared memory: kernels declare it with BZ_DYN_SMEM(type, name)
#define BZ_DYN_SMEM(type, name) type* const name = reinterpret_cast<type*>(::emu::g_dyn_smem)
// a spin-wait on shared/global memory must give the other fibers a chance to run
#define BZ_SPIN_HINT() ::emu::yield()
