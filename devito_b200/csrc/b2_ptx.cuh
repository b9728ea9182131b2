// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), proxy fences.
#pragma once
#include <cuda.h>
#include <cstdint>

namespace b2ptx {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(
                     smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile(
        "{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(
            smem_u32(bar)),
        "r"(bytes)
        : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// TMA tiled loads global -> shared, completion signalled on an mbarrier (complete_tx bytes)
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *tm, uint64_t *bar,
                                            int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3)
        : "memory");
}

__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *tm, uint64_t *bar,
                                            int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}

// generic-proxy writes (ours or a peer GPU's, observed through an acquire) -> async-proxy (TMA) reads of global memory
__device__ __forceinline__ void fence_proxy_async_global() {
    asm volatile("fence.proxy.async.global;" ::: "memory");
}

// system-scope flag handshake between GPUs (peer-mapped memory)
__device__ __forceinline__ int ld_acquire_sys(const int *p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void st_release_sys(int *p, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---- packed fp32x2 arithmetic (sm_100: FFMA2 / FADD2 / FMUL2) -----------------------------------------
// One instruction does two fp32 operations on an aligned register pair. The FP32 pipe rate is unchanged
// (profiles/r2b_micro_*: 127 FMA/clk/SM either way); what it halves is ISSUE SLOTS — the limiter of the
// TTI kernel (ncu r1g: 169 instructions/point, 66 % issue utilisation, HBM at 51 %).
typedef unsigned long long u64;

__device__ __forceinline__ u64 pk2(float a, float b) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void upk2(u64 v, float &a, float &b) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
    u64 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
    u64 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// a float4 as two packed pairs (x,y) (z,w): what LDS.128 / LDG.128 deliver in aligned register quads
struct F4 {
    u64 a, b;
};
__device__ __forceinline__ F4 f4pack(const float4 &v) { return F4{pk2(v.x, v.y), pk2(v.z, v.w)}; }
__device__ __forceinline__ float4 f4unpack(const F4 &p) {
    float4 v;
    upk2(p.a, v.x, v.y);
    upk2(p.b, v.z, v.w);
    return v;
}
__device__ __forceinline__ F4 f4zero() { return F4{0ull, 0ull}; }
__device__ __forceinline__ void f4fma2(F4 &acc, float2 w, const F4 &v) {     // w = {w, w}
    const u64 ww = pk2(w.x, w.y);
    acc.a = fma2(ww, v.a, acc.a);
    acc.b = fma2(ww, v.b, acc.b);
}
__device__ __forceinline__ F4 f4add2(const F4 &x, const F4 &y) { return F4{add2(x.a, y.a), add2(x.b, y.b)}; }
__device__ __forceinline__ F4 f4sub2(const F4 &x, const F4 &y) { return F4{sub2(x.a, y.a), sub2(x.b, y.b)}; }
__device__ __forceinline__ F4 f4mul2(float2 w, const F4 &v) {
    const u64 ww = pk2(w.x, w.y);
    return F4{mul2(ww, v.a), mul2(ww, v.b)};
}

__device__ __forceinline__ float4 lds128(const float *p) {
    return *reinterpret_cast<const float4 *>(p);
}

}  // namespace b2ptx
