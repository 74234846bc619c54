// ubench_chain.cu -- formulations of the decoder's 8-decision walk on ONE warp (the chain warp of cm_decode_kernel).
// Every variant decodes the same pseudo-random bytes from the same probability table in shared memory (layout of
// ptab in cm.cuh: P << 14 per tree node, node 0 unused) and must produce the same checksum; printed: cycles per byte.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/variants/ubench_chain tools/ubench_chain.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

typedef uint32_t u32;
typedef uint64_t u64;

__device__ u64 g_cycles[16];
__device__ u32 g_sum[16];

__device__ __forceinline__ u32 mulhi_pinned(u32 a, u32 b) {
    u32 r;
    asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ u32 next_code(u32& s) {
    s = s * 1664525u + 1013904223u;
    return s ^ (s >> 15);
}

// V0: the shipped fast tier (mid = low + x, compare code with mid, select range and the child's P, one multiply)
__device__ __forceinline__ u32 walk_v0(const u32* pt, u32 code, u32 range_in, u32& range_out, u32& tmin_out) {
    const uint4 g0 = *reinterpret_cast<const uint4*>(pt);
    uint4 gk = *reinterpret_cast<const uint4*>(pt + 4);
    u32 node = 1, flow = 0, frange = range_in;
    u32 pcur = g0.y, kid0 = g0.z, kid1 = g0.w;
    u32 x = mulhi_pinned(frange, pcur);
    u32 tmin = 0xFFFFFFFFu;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        u32 bit;
        asm volatile(
            "{\n\t"
            ".reg .pred pb;\n\t"
            ".reg .u32 mid, nx, r0, hi, t;\n\t"
            "add.u32 mid, %0, %2;\n\t"
            "not.b32 nx, %2;\n\t"
            "setp.le.u32 pb, %6, mid;\n\t"
            "add.u32 r0, %1, nx;\n\t"
            "selp.u32 %1, %2, r0, pb;\n\t"
            "selp.u32 %3, %8, %7, pb;\n\t"
            "mul.hi.u32 %2, %1, %3;\n\t"
            "@!pb add.u32 %0, mid, 1;\n\t"
            "selp.u32 %4, 1, 0, pb;\n\t"
            "add.u32 hi, %0, %1;\n\t"
            "xor.b32 t, %0, hi;\n\t"
            "min.u32 %5, %5, t;\n\t"
            "}"
            : "+r"(flow), "+r"(frange), "+r"(x), "+r"(pcur), "=r"(bit), "+r"(tmin)
            : "r"(code), "r"(kid0), "r"(kid1));
        node = node * 2 + bit;
        kid0 = bit ? gk.z : gk.x;
        kid1 = bit ? gk.w : gk.y;
        if (k < 5) gk = *reinterpret_cast<const uint4*>(pt + 4 * node);
    }
    tmin_out = tmin;
    range_out = frange;
    return node;
}

// V1: compare d = code - low with x (no add before the compare), otherwise as V0
__device__ __forceinline__ u32 walk_v1(const u32* pt, u32 code, u32 range_in, u32& range_out, u32& tmin_out) {
    const uint4 g0 = *reinterpret_cast<const uint4*>(pt);
    uint4 gk = *reinterpret_cast<const uint4*>(pt + 4);
    u32 node = 1, d = code, frange = range_in;
    u32 pcur = g0.y, kid0 = g0.z, kid1 = g0.w;
    u32 x = mulhi_pinned(frange, pcur);
    u32 tmin = 0xFFFFFFFFu;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        u32 bit;
        asm volatile(
            "{\n\t"
            ".reg .pred pb;\n\t"
            ".reg .u32 nx, r0, lo, hi, t;\n\t"
            "not.b32 nx, %2;\n\t"
            "setp.le.u32 pb, %0, %2;\n\t"        // bit = d <= x
            "add.u32 r0, %1, nx;\n\t"
            "selp.u32 %1, %2, r0, pb;\n\t"
            "selp.u32 %3, %8, %7, pb;\n\t"
            "@!pb add.u32 %0, %0, nx;\n\t"       // bit 0: d -= x + 1
            "mul.hi.u32 %2, %1, %3;\n\t"
            "selp.u32 %4, 1, 0, pb;\n\t"
            "sub.u32 lo, %6, %0;\n\t"            // low = code - d
            "add.u32 hi, lo, %1;\n\t"
            "xor.b32 t, lo, hi;\n\t"
            "min.u32 %5, %5, t;\n\t"
            "}"
            : "+r"(d), "+r"(frange), "+r"(x), "+r"(pcur), "=r"(bit), "+r"(tmin)
            : "r"(code), "r"(kid0), "r"(kid1));
        node = node * 2 + bit;
        kid0 = bit ? gk.z : gk.x;
        kid1 = bit ? gk.w : gk.y;
        if (k < 5) gk = *reinterpret_cast<const uint4*>(pt + 4 * node);
    }
    tmin_out = tmin;
    range_out = frange;
    return node;
}

// V2: V1 + both next products are formed before the bit is known (x1 = hi(x * kid1), x0 = hi((range - x - 1) * kid0))
__device__ __forceinline__ u32 walk_v2(const u32* pt, u32 code, u32 range_in, u32& range_out, u32& tmin_out) {
    const uint4 g0 = *reinterpret_cast<const uint4*>(pt);
    uint4 gk = *reinterpret_cast<const uint4*>(pt + 4);
    u32 node = 1, d = code, frange = range_in;
    u32 kid0 = g0.z, kid1 = g0.w;
    u32 x = mulhi_pinned(frange, g0.y);
    u32 tmin = 0xFFFFFFFFu;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        u32 bit;
        asm volatile(
            "{\n\t"
            ".reg .pred pb;\n\t"
            ".reg .u32 nx, r0, x0, x1, lo, hi, t;\n\t"
            "not.b32 nx, %2;\n\t"
            "mul.hi.u32 x1, %2, %7;\n\t"
            "add.u32 r0, %1, nx;\n\t"
            "setp.le.u32 pb, %0, %2;\n\t"
            "mul.hi.u32 x0, r0, %6;\n\t"
            "@!pb add.u32 %0, %0, nx;\n\t"
            "selp.u32 %1, %2, r0, pb;\n\t"
            "selp.u32 %2, x1, x0, pb;\n\t"
            "selp.u32 %3, 1, 0, pb;\n\t"
            "sub.u32 lo, %5, %0;\n\t"
            "add.u32 hi, lo, %1;\n\t"
            "xor.b32 t, lo, hi;\n\t"
            "min.u32 %4, %4, t;\n\t"
            "}"
            : "+r"(d), "+r"(frange), "+r"(x), "=r"(bit), "+r"(tmin)
            : "r"(code), "r"(kid0), "r"(kid1));
        node = node * 2 + bit;
        kid0 = bit ? gk.z : gk.x;
        kid1 = bit ? gk.w : gk.y;
        if (k < 5) gk = *reinterpret_cast<const uint4*>(pt + 4 * node);
    }
    tmin_out = tmin;
    range_out = frange;
    return node;
}

// V3: V2 with the table rows of BOTH children requested before the bit is known (two 128-bit loads per level, the
// chosen one kept by selects): takes the shared-memory round trip off the path bit -> address -> load -> kid select
__device__ __forceinline__ u32 walk_v3(const u32* pt, u32 code, u32 range_in, u32& range_out, u32& tmin_out) {
    const uint4 g0 = *reinterpret_cast<const uint4*>(pt);
    u32 node = 1, d = code, frange = range_in;
    u32 kid0 = g0.z, kid1 = g0.w;
    u32 x = mulhi_pinned(frange, g0.y);
    u32 tmin = 0xFFFFFFFFu;
    // rows of node 2 and node 3 (grandchildren of the root by child)
    uint4 ga = *reinterpret_cast<const uint4*>(pt + 4);   // row of the root
    // row(n) = p[4n .. 4n+3] = P of the grandchildren of n.  At level k (node n) we need, after bit(n), kids of child c:
    // row(n)[2 bit], row(n)[2 bit + 1].  Rows of both children of n are requested at the start of level n.
    uint4 ra, rb;   // rows of children 2n and 2n+1
#pragma unroll
    for (int k = 0; k < 8; k++) {
        if (k < 5) {
            ra = *reinterpret_cast<const uint4*>(pt + 8 * node);
            rb = *reinterpret_cast<const uint4*>(pt + 8 * node + 4);
        }
        u32 bit;
        asm volatile(
            "{\n\t"
            ".reg .pred pb;\n\t"
            ".reg .u32 nx, r0, x0, x1, lo, hi, t;\n\t"
            "not.b32 nx, %2;\n\t"
            "mul.hi.u32 x1, %2, %7;\n\t"
            "add.u32 r0, %1, nx;\n\t"
            "setp.le.u32 pb, %0, %2;\n\t"
            "mul.hi.u32 x0, r0, %6;\n\t"
            "@!pb add.u32 %0, %0, nx;\n\t"
            "selp.u32 %1, %2, r0, pb;\n\t"
            "selp.u32 %2, x1, x0, pb;\n\t"
            "selp.u32 %3, 1, 0, pb;\n\t"
            "sub.u32 lo, %5, %0;\n\t"
            "add.u32 hi, lo, %1;\n\t"
            "xor.b32 t, lo, hi;\n\t"
            "min.u32 %4, %4, t;\n\t"
            "}"
            : "+r"(d), "+r"(frange), "+r"(x), "=r"(bit), "+r"(tmin)
            : "r"(code), "r"(kid0), "r"(kid1));
        node = node * 2 + bit;
        kid0 = bit ? ga.z : ga.x;
        kid1 = bit ? ga.w : ga.y;
        if (k < 5) {
            ga.x = bit ? rb.x : ra.x;
            ga.y = bit ? rb.y : ra.y;
            ga.z = bit ? rb.z : ra.z;
            ga.w = bit ? rb.w : ra.w;
        }
    }
    tmin_out = tmin;
    range_out = frange;
    return node;
}

// V4: V2 with a borrow mask instead of a predicate: m = (d <= x) ? ~0 : 0 from a 64-bit subtract, selects by LOP3
__device__ __forceinline__ u32 walk_v4(const u32* pt, u32 code, u32 range_in, u32& range_out, u32& tmin_out) {
    const uint4 g0 = *reinterpret_cast<const uint4*>(pt);
    uint4 gk = *reinterpret_cast<const uint4*>(pt + 4);
    u32 node = 1, d = code, frange = range_in;
    u32 kid0 = g0.z, kid1 = g0.w;
    u32 x = mulhi_pinned(frange, g0.y);
    u32 tmin = 0xFFFFFFFFu;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const u32 nx = ~x;
        const u32 x1 = mulhi_pinned(x, kid1);
        const u32 r0 = frange + nx;
        const u32 x0 = mulhi_pinned(r0, kid0);
        const u32 m = (u32)(((u64)x - (u64)d) >> 32) ^ 0xFFFFFFFFu;   // x >= d: no borrow -> high word 0 -> m = ~0
        const u32 d0 = d + nx;
        d = (d & m) | (d0 & ~m);
        frange = (x & m) | (r0 & ~m);
        x = (x1 & m) | (x0 & ~m);
        const u32 bit = m & 1u;
        const u32 lo = code - d;
        const u32 t = lo ^ (lo + frange);
        tmin = t < tmin ? t : tmin;
        node = node * 2 + bit;
        kid0 = (gk.z & m) | (gk.x & ~m);
        kid1 = (gk.w & m) | (gk.y & ~m);
        if (k < 5) gk = *reinterpret_cast<const uint4*>(pt + 4 * node);
    }
    tmin_out = tmin;
    range_out = frange;
    return node;
}

// V5: V4 with the rows of both children requested before the bit is known (as V3), all selects by LOP3
__device__ __forceinline__ u32 walk_v5(const u32* pt, u32 code, u32 range_in, u32& range_out, u32& tmin_out) {
    const uint4 g0 = *reinterpret_cast<const uint4*>(pt);
    uint4 ga = *reinterpret_cast<const uint4*>(pt + 4);
    u32 node = 1, d = code, frange = range_in;
    u32 kid0 = g0.z, kid1 = g0.w;
    u32 x = mulhi_pinned(frange, g0.y);
    u32 tmin = 0xFFFFFFFFu;
    uint4 ra, rb;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        if (k < 5) {
            ra = *reinterpret_cast<const uint4*>(pt + 8 * node);
            rb = *reinterpret_cast<const uint4*>(pt + 8 * node + 4);
        }
        const u32 nx = ~x;
        const u32 x1 = mulhi_pinned(x, kid1);
        const u32 r0 = frange + nx;
        const u32 x0 = mulhi_pinned(r0, kid0);
        const u32 m = (u32)(((u64)x - (u64)d) >> 32) ^ 0xFFFFFFFFu;
        const u32 d0 = d + nx;
        d = (d & m) | (d0 & ~m);
        frange = (x & m) | (r0 & ~m);
        x = (x1 & m) | (x0 & ~m);
        const u32 lo = code - d;
        const u32 t = lo ^ (lo + frange);
        tmin = t < tmin ? t : tmin;
        node = node * 2 + (m & 1u);
        kid0 = (ga.z & m) | (ga.x & ~m);
        kid1 = (ga.w & m) | (ga.y & ~m);
        if (k < 5) {
            ga.x = (rb.x & m) | (ra.x & ~m);
            ga.y = (rb.y & m) | (ra.y & ~m);
            ga.z = (rb.z & m) | (ra.z & ~m);
            ga.w = (rb.w & m) | (ra.w & ~m);
        }
    }
    tmin_out = tmin;
    range_out = frange;
    return node;
}

// V6: V1 (one multiply, no speculation) with the borrow mask instead of the predicate
__device__ __forceinline__ u32 walk_v6(const u32* pt, u32 code, u32 range_in, u32& range_out, u32& tmin_out) {
    const uint4 g0 = *reinterpret_cast<const uint4*>(pt);
    uint4 gk = *reinterpret_cast<const uint4*>(pt + 4);
    u32 node = 1, d = code, frange = range_in;
    u32 kid0 = g0.z, kid1 = g0.w;
    u32 x = mulhi_pinned(frange, g0.y);
    u32 tmin = 0xFFFFFFFFu;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const u32 nx = ~x;
        const u32 r0 = frange + nx;
        const u32 m = (u32)(((u64)x - (u64)d) >> 32) ^ 0xFFFFFFFFu;
        const u32 d0 = d + nx;
        d = (d & m) | (d0 & ~m);
        frange = (x & m) | (r0 & ~m);
        const u32 pc = (kid1 & m) | (kid0 & ~m);
        x = mulhi_pinned(frange, pc);
        const u32 lo = code - d;
        const u32 t = lo ^ (lo + frange);
        tmin = t < tmin ? t : tmin;
        node = node * 2 + (m & 1u);
        kid0 = (gk.z & m) | (gk.x & ~m);
        kid1 = (gk.w & m) | (gk.y & ~m);
        if (k < 5) gk = *reinterpret_cast<const uint4*>(pt + 4 * node);
    }
    tmin_out = tmin;
    range_out = frange;
    return node;
}

// V7: V0 without the running minimum (one renormalisation test per byte: nested intervals)
__device__ __forceinline__ u32 walk_v7(const u32* pt, u32 code, u32 range_in, u32& range_out, u32& tmin_out) {
    const uint4 g0 = *reinterpret_cast<const uint4*>(pt);
    uint4 gk = *reinterpret_cast<const uint4*>(pt + 4);
    u32 node = 1, flow = 0, frange = range_in;
    u32 pcur = g0.y, kid0 = g0.z, kid1 = g0.w;
    u32 x = mulhi_pinned(frange, pcur);
#pragma unroll
    for (int k = 0; k < 8; k++) {
        u32 bit;
        asm volatile(
            "{\n\t"
            ".reg .pred pb;\n\t"
            ".reg .u32 mid, nx, r0;\n\t"
            "add.u32 mid, %0, %2;\n\t"
            "not.b32 nx, %2;\n\t"
            "setp.le.u32 pb, %5, mid;\n\t"
            "add.u32 r0, %1, nx;\n\t"
            "selp.u32 %1, %2, r0, pb;\n\t"
            "selp.u32 %3, %7, %6, pb;\n\t"
            "mul.hi.u32 %2, %1, %3;\n\t"
            "@!pb add.u32 %0, mid, 1;\n\t"
            "selp.u32 %4, 1, 0, pb;\n\t"
            "}"
            : "+r"(flow), "+r"(frange), "+r"(x), "+r"(pcur), "=r"(bit)
            : "r"(code), "r"(kid0), "r"(kid1));
        node = node * 2 + bit;
        kid0 = bit ? gk.z : gk.x;
        kid1 = bit ? gk.w : gk.y;
        if (k < 5) gk = *reinterpret_cast<const uint4*>(pt + 4 * node);
    }
    tmin_out = flow ^ (flow + frange);
    range_out = frange;
    return node;
}

// V8: the predicated exact walk (tier A of cm_decode_kernel): one predicated one-byte shift per step, payload bytes from a
// register.  Here no shift is ever due (wide range), so the result equals V0's; the cost is that of the instruction stream.
__device__ __forceinline__ u32 walk_v8(const u32* pt, u32 code, u32 range_in, u32& range_out, u32& tmin_out) {
    const uint4 g0 = *reinterpret_cast<const uint4*>(pt);
    uint4 pg = *reinterpret_cast<const uint4*>(pt + 4);
    u32 pnode = 1, plow = 0, prange = range_in, pcode = code, W = code * 2654435761u, nsh = 0, multi = 0;
    u32 pp = g0.y, pk0 = g0.z, pk1 = g0.w;
    u32 px = mulhi_pinned(prange, pp);
#pragma unroll
    for (int k = 0; k < 8; k++) {
        u32 bit;
        asm volatile(
            "{\n\t"
            ".reg .pred pb, ps, pm;\n\t"
            ".reg .u32 mid, nx, r0, hi, t;\n\t"
            "add.u32 mid, %0, %2;\n\t"
            "not.b32 nx, %2;\n\t"
            "setp.le.u32 pb, %5, mid;\n\t"
            "add.u32 r0, %1, nx;\n\t"
            "selp.u32 %1, %2, r0, pb;\n\t"
            "selp.u32 %3, %10, %9, pb;\n\t"
            "@!pb add.u32 %0, mid, 1;\n\t"
            "selp.u32 %4, 1, 0, pb;\n\t"
            "add.u32 hi, %0, %1;\n\t"
            "xor.b32 t, %0, hi;\n\t"
            "setp.lt.u32 ps, t, 0x1000000;\n\t"
            "setp.lt.u32 pm, t, 0x10000;\n\t"
            "@ps shl.b32 %0, %0, 8;\n\t"
            "@ps mad.lo.u32 %1, %1, 256, 255;\n\t"
            "@ps shf.l.clamp.b32 %5, %6, %5, 8;\n\t"
            "@ps shl.b32 %6, %6, 8;\n\t"
            "@ps add.u32 %7, %7, 1;\n\t"
            "@pm mov.u32 %8, 1;\n\t"
            "mul.hi.u32 %2, %1, %3;\n\t"
            "}"
            : "+r"(plow), "+r"(prange), "+r"(px), "+r"(pp), "=r"(bit), "+r"(pcode), "+r"(W), "+r"(nsh), "+r"(multi)
            : "r"(pk0), "r"(pk1));
        pnode = pnode * 2 + bit;
        pk0 = bit ? pg.z : pg.x;
        pk1 = bit ? pg.w : pg.y;
        if (k < 5) pg = *reinterpret_cast<const uint4*>(pt + 4 * pnode);
    }
    tmin_out = (plow ^ (plow + prange)) | (multi << 31) | (nsh << 28);
    range_out = prange;
    return pnode;
}

// V9: V7 with the row ADDRESS carried instead of the node: both children's row addresses are formed before the bit is known,
// one select after it (instead of bit -> node -> address: three dependent instructions before the load can issue)
__device__ __forceinline__ u32 walk_v9(const u32* pt, u32 code, u32 range_in, u32& range_out, u32& tmin_out) {
    const u32 pbase = (u32)__cvta_generic_to_shared(pt);
    uint4 g0, gk;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(g0.x), "=r"(g0.y), "=r"(g0.z), "=r"(g0.w) : "r"(pbase));
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+16];" : "=r"(gk.x), "=r"(gk.y), "=r"(gk.z), "=r"(gk.w) : "r"(pbase));
    u32 ra = pbase + 16u, flow = 0, frange = range_in;   // ra = address of row(node) = pbase + 16 * node
    u32 pcur = g0.y, kid0 = g0.z, kid1 = g0.w;
    u32 x = mulhi_pinned(frange, pcur);
#pragma unroll
    for (int k = 0; k < 8; k++) {
        u32 bit;
        const u32 c0 = 2u * ra - pbase, c1 = c0 + 16u;
        asm volatile(
            "{\n\t"
            ".reg .pred pb;\n\t"
            ".reg .u32 mid, nx, r0;\n\t"
            "add.u32 mid, %0, %2;\n\t"
            "not.b32 nx, %2;\n\t"
            "setp.le.u32 pb, %6, mid;\n\t"
            "add.u32 r0, %1, nx;\n\t"
            "selp.u32 %1, %2, r0, pb;\n\t"
            "selp.u32 %3, %8, %7, pb;\n\t"
            "selp.u32 %5, %10, %9, pb;\n\t"
            "mul.hi.u32 %2, %1, %3;\n\t"
            "@!pb add.u32 %0, mid, 1;\n\t"
            "selp.u32 %4, 1, 0, pb;\n\t"
            "}"
            : "+r"(flow), "+r"(frange), "+r"(x), "+r"(pcur), "=r"(bit), "=r"(ra)
            : "r"(code), "r"(kid0), "r"(kid1), "r"(c0), "r"(c1));
        kid0 = bit ? gk.z : gk.x;
        kid1 = bit ? gk.w : gk.y;
        if (k < 5) asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(gk.x), "=r"(gk.y), "=r"(gk.z), "=r"(gk.w) : "r"(ra));
    }
    tmin_out = flow ^ (flow + frange);
    range_out = frange;
    return (ra - pbase) >> 4;
}

template <int V>
__global__ void k_walk(int nbytes, u32 seed, int slot) {
    __shared__ __align__(16) u32 ptab[512];
    for (int k = threadIdx.x; k < 512; k += blockDim.x) {
        u32 s = (u32)k * 2654435761u + 12345u;
        s ^= s >> 13;
        const u32 P = 2048u + (s % 258000u);   // 18-bit probability, away from the ends
        ptab[k] = P << 14;
    }
    __syncthreads();
    u32 s = seed, sum = 0, tacc = 0xFFFFFFFFu, range = 0xFFFFFFFFu;
    const u64 t0 = clock64();
    for (int i = 0; i < nbytes; i++) {
        const u32 code = next_code(s);
        asm volatile("" ::: "memory");   // the table is rewritten between bytes in the real kernel: no hoisted loads
        const u32* pt = ptab + (i & 1) * 256;
        u32 tm, node;
        if (V == 0) node = walk_v0(pt, code, range, range, tm);
        else if (V == 1) node = walk_v1(pt, code, range, range, tm);
        else if (V == 2) node = walk_v2(pt, code, range, range, tm);
        else if (V == 3) node = walk_v3(pt, code, range, range, tm);
        else if (V == 4) node = walk_v4(pt, code, range, range, tm);
        else if (V == 5) node = walk_v5(pt, code, range, range, tm);
        else if (V == 6) node = walk_v6(pt, code, range, range, tm);
        else if (V == 7) node = walk_v7(pt, code, range, range, tm);
        else if (V == 8) node = walk_v8(pt, code, range, range, tm);
        else node = walk_v9(pt, code, range, range, tm);
        range |= 0xFF000000u;   // stands for the renormalisation: the range stays wide, the dependency stays
        if (tm < (1u << 24)) sum += 0x9E3779B9u;   // the tier test of the real kernel (here: rare or never)
        tacc = tm < tacc ? tm : tacc;
        sum = sum * 31u + node;
        s += node;   // the next code depends on this byte: bytes cannot overlap, as in the real chain
    }
    const u64 t1 = clock64();
    if (threadIdx.x == 0) {
        g_cycles[slot] = t1 - t0;
        g_sum[slot] = sum;
        if (tacc == 0x12345u) g_sum[15] = tacc;   // keep the minimum alive
    }
}

int main() {
    const int n = 1 << 16;
    const char* names[10] = {"V0 shipped fast tier (add, compare, 2 selects, 1 multiply)",
                            "V1 compare d = code - low with x (no add before the compare)",
                            "V2 V1 + both next products before the bit is known",
                            "V3 V2 + rows of both children requested before the bit",
                            "V4 V2 with a borrow mask and LOP3 selects (no predicate)",
                            "V5 V4 + rows of both children requested before the bit",
                            "V6 V1 with the borrow mask (one multiply, no predicate)",
                            "V7 V0 without the running minimum (one test per byte)",
                            "V8 predicated exact walk (tier A), no shift ever due",
                            "V9 V7 with the row address carried (one select between bit and load)"};
    for (int rep = 0; rep < 2; rep++) {
        k_walk<0><<<1, 32>>>(n, 777u, 0);
        k_walk<1><<<1, 32>>>(n, 777u, 1);
        k_walk<2><<<1, 32>>>(n, 777u, 2);
        k_walk<3><<<1, 32>>>(n, 777u, 3);
        k_walk<4><<<1, 32>>>(n, 777u, 4);
        k_walk<5><<<1, 32>>>(n, 777u, 5);
        k_walk<6><<<1, 32>>>(n, 777u, 6);
        k_walk<7><<<1, 32>>>(n, 777u, 7);
        k_walk<8><<<1, 32>>>(n, 777u, 8);
        k_walk<9><<<1, 32>>>(n, 777u, 9);
        cudaDeviceSynchronize();
    }
    u64 cyc[16];
    u32 sum[16];
    cudaMemcpyFromSymbol(cyc, g_cycles, sizeof(cyc));
    cudaMemcpyFromSymbol(sum, g_sum, sizeof(sum));
    for (int v = 0; v < 10; v++)
        printf("%-66s: %7.1f cycles per byte  (checksum %08x%s)\n", names[v], (double)cyc[v] / n, sum[v],
               sum[v] == sum[0] ? "" : "  MISMATCH");
    printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
