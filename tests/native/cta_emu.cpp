// cta_emu.cpp -- scheduler and context switch of the CPU thread-block emulator (see cta_emu.h).
// Test infrastructure only.
#include "cta_emu.h"

#include <mutex>

#if defined(__SANITIZE_ADDRESS__)
#include <sanitizer/asan_interface.h>
#endif

namespace emu {
Cta* g_cta = nullptr;
Fiber* g_cur = nullptr;
Dim3 g_blockIdx, g_blockDim, g_gridDim;
unsigned char* g_dyn_smem = nullptr;
int g_schedule = 0;
uint64_t g_rng = 88172645463325252ull;
char g_tsan_token = 0;
}  // namespace emu

// void emu_switch(void** save_sp, void* load_sp): save the callee-saved registers of the running fiber
// on its stack, publish its stack pointer, adopt the other stack and restore from it.
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size emu_switch,.-emu_switch
)");

namespace emu {

static std::vector<char*> g_stack_pool;

static char* get_stack(size_t i) {
    while (g_stack_pool.size() <= i) {
        void* p = mmap(nullptr, kStackBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (p == MAP_FAILED) { perror("mmap"); abort(); }
        g_stack_pool.push_back((char*)p);
    }
    return g_stack_pool[i];
}

static unsigned char* g_smem_buf = nullptr;
static size_t g_smem_cap = 0;

static std::mutex g_launch_mutex;   // one grid at a time: the emulator state and the kernels' static "shared" arrays are global

void launch(Dim3 grid, Dim3 block, size_t dyn_smem_bytes, const std::function<void()>& body) {
    std::lock_guard<std::mutex> lock(g_launch_mutex);
    BZ_TSAN_INTERNAL();
    const unsigned nthreads = block.x * block.y * block.z;
    if (nthreads == 0 || nthreads > (unsigned)kMaxThreads) { fprintf(stderr, "[cta_emu] bad block size\n"); abort(); }
    if (dyn_smem_bytes + 64 > g_smem_cap) {
        free(g_smem_buf);
        g_smem_cap = dyn_smem_bytes + 64;
        g_smem_buf = (unsigned char*)aligned_alloc(128, (g_smem_cap + 127) & ~(size_t)127);
    }
    g_dyn_smem = g_smem_buf;
    g_blockDim = block;
    g_gridDim = grid;
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                Cta cta;
                cta.body = body;
                cta.fibers.resize(nthreads);
                cta.warps.resize((nthreads + 31) / 32);
                cta.live = nthreads;
                for (unsigned t = 0; t < nthreads; t++) {
                    Fiber& f = cta.fibers[t];
                    f.tid.x = t % block.x;
                    f.tid.y = (t / block.x) % block.y;
                    f.tid.z = t / (block.x * block.y);
                    f.stack = get_stack(t);
#if defined(__SANITIZE_THREAD__)
                    f.tsan = __tsan_create_fiber(0);
#endif
#if defined(__SANITIZE_ADDRESS__)
                    // fibers of the previous launch left by switching away, never unwinding: their redzones are stale
                    __asan_unpoison_memory_region(f.stack, kStackBytes);
#endif
                    uintptr_t top = ((uintptr_t)f.stack + kStackBytes) & ~(uintptr_t)15;
                    uintptr_t* a = (uintptr_t*)(top - 16);  // 16-byte aligned slot of the first return address
                    a[0] = (uintptr_t)&emu_fiber_main;
                    a[1] = 0;
                    uintptr_t* sp = a - 6;
                    for (int k = 0; k < 6; k++) sp[k] = 0;
                    f.sp = sp;
                }
                memset(g_smem_buf, 0xA5, dyn_smem_bytes);  // shared memory is not zero-initialised on the device either
                g_blockIdx.x = bx;
                g_blockIdx.y = by;
                g_blockIdx.z = bz;
                g_cta = &cta;
                cta.cur = 0;
                g_cur = &cta.fibers[0];
#if defined(__SANITIZE_THREAD__)
                cta.main_tsan = __tsan_get_current_fiber();
                __tsan_release(&g_tsan_token);      // everything the host did so far happens before the grid
#endif
                BZ_TSAN_SWITCH(g_cur->tsan);
                emu_switch(&cta.main_sp, g_cur->sp);
#if defined(__SANITIZE_THREAD__)
                __tsan_acquire(&g_tsan_token);      // ... and the grid happens before what the host does next
                for (auto& f : cta.fibers) __tsan_destroy_fiber(f.tsan);
#endif
                for (auto& f : cta.fibers)
                    if (!f.done) { fprintf(stderr, "[cta_emu] a thread never finished\n"); abort(); }
                g_cta = nullptr;
                g_cur = nullptr;
            }
}

}  // namespace emu
