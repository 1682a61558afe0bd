// simt.h -- the handful of warp-level primitives the decode kernels use.
//
// Device build (nvcc, sm_100a): thin wrappers over the CUDA intrinsics and the
// cp.async (LDGSTS) PTX used to stage compressed bytes into shared memory.
//
// Host build (-DDBG_SIMT_EMU, used ONLY by tests/simt_emu): the same names are
// provided by a single-threaded 32-lane coroutine emulator so the kernel bodies
// in *_core.h can be exercised on a CPU-only box before spending GPU minutes.
// The emulator is test scaffolding for the kernel sources; it is never built
// into, linked with, or reachable from the product library.
#pragma once
#include <stdint.h>

#ifndef DBG_SIMT_EMU
// ----------------------------------------------------------------- device ----
#include <cuda_runtime.h>

#define DBG_DEV __device__ __forceinline__
#define DBG_DEVM __device__ __forceinline__
#define DBG_DEV_NOINLINE __device__ __noinline__
#define DBG_FULL 0xffffffffu

namespace simt {

DBG_DEV int lane() { return (int)(threadIdx.x & 31); }
DBG_DEV void syncwarp() { __syncwarp(); }
DBG_DEV uint32_t ballot(bool p) { return __ballot_sync(DBG_FULL, p); }
DBG_DEV bool any(bool p) { return __any_sync(DBG_FULL, p); }
DBG_DEV uint32_t match_any(uint32_t v) { return __match_any_sync(DBG_FULL, v); }
DBG_DEV uint32_t shfl(uint32_t v, int src) { return __shfl_sync(DBG_FULL, v, src); }
DBG_DEV uint32_t shfl_up(uint32_t v, int d) { return __shfl_up_sync(DBG_FULL, v, d); }
DBG_DEV uint32_t shfl_down(uint32_t v, int d) { return __shfl_down_sync(DBG_FULL, v, d); }
DBG_DEV uint32_t shfl_xor(uint32_t v, int m) { return __shfl_xor_sync(DBG_FULL, v, m); }

DBG_DEV uint32_t brev(uint32_t v) { return __brev(v); }
DBG_DEV int popc(uint32_t v) { return __popc(v); }
DBG_DEV int clz(uint32_t v) { return __clz(v); }
DBG_DEV int ffs(uint32_t v) { return __ffs(v); }  // 1-based, 0 if none
DBG_DEV uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) { return __funnelshift_r(lo, hi, sh); }

// 16-byte global -> shared asynchronous copy (LDGSTS). `src_bytes` is 16 or 0;
// with 0 the destination is zero-filled and the source is not read.
DBG_DEV void cp_async16(void *smem_dst, const void *gsrc, int src_bytes)
{
    uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(src_bytes)
                 : "memory");
}
// cross-warp hand-off through global memory (band pipeline of the PNG un-filter)
DBG_DEV uint64_t ld_acquire_u64(const uint64_t *p)
{
    uint64_t v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    return v;
}
DBG_DEV void st_release_u64(uint64_t *p, uint64_t v)
{
    asm volatile("st.release.gpu.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
DBG_DEV void threadfence() { __threadfence(); }
DBG_DEV void backoff() { __nanosleep(64); }
DBG_DEV uint32_t ldg_u32(const uint32_t *p) { return __ldg(p); }
DBG_DEV void ldg_u32x4(const uint32_t *p, uint32_t &a, uint32_t &b, uint32_t &c, uint32_t &d)  // p 16-byte aligned
{
    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
    a = v.x, b = v.y, c = v.z, d = v.w;
}
DBG_DEV uint32_t atomic_inc_shared(uint32_t *p) { return atomicAdd(p, 1u); }
DBG_DEV uint32_t ldcg_u32(const uint32_t *p) { return __ldcg(p); }
DBG_DEV uint32_t ldcg_u8(const uint8_t *p) { return __ldcg(p); }

// Same, for data that is read exactly once (the compressed input): the L2 line is marked evict-first so
// that the stream does not push the recently written output -- which LZ77 matches read back -- out of L2.
DBG_DEV void cp_async16_stream(void *smem_dst, const void *gsrc, int src_bytes)
{
    uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(pol));
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2, %3;\n" ::"r"(d), "l"(gsrc), "r"(src_bytes), "l"(pol)
                 : "memory");
}
DBG_DEV void st_u32x4(uint32_t *p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) { *reinterpret_cast<uint4 *>(p) = make_uint4(a, b, c, d); }  // p 16-byte aligned
DBG_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
DBG_DEV void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

}  // namespace simt

#else
// --------------------------------------------------------------- emulator ----
#include <string.h>

#define DBG_DEV static inline
#define DBG_DEVM inline
#define DBG_DEV_NOINLINE static
#define DBG_FULL 0xffffffffu

namespace simt {

// implemented in tests/simt_emu/emu.cpp
extern int g_lane;
extern uint32_t g_slot[32];
void emu_barrier();  // every lane must call it the same number of times

DBG_DEV int lane() { return g_lane; }
DBG_DEV void syncwarp() { emu_barrier(); }

DBG_DEV uint32_t shfl(uint32_t v, int src)
{
    g_slot[g_lane] = v;
    emu_barrier();
    uint32_t r = g_slot[src & 31];
    emu_barrier();
    return r;
}
DBG_DEV uint32_t shfl_up(uint32_t v, int d)
{
    g_slot[g_lane] = v;
    emu_barrier();
    uint32_t r = (g_lane - d >= 0) ? g_slot[g_lane - d] : v;
    emu_barrier();
    return r;
}
DBG_DEV uint32_t shfl_down(uint32_t v, int d)
{
    g_slot[g_lane] = v;
    emu_barrier();
    uint32_t r = (g_lane + d < 32) ? g_slot[g_lane + d] : v;
    emu_barrier();
    return r;
}
DBG_DEV uint32_t shfl_xor(uint32_t v, int m)
{
    g_slot[g_lane] = v;
    emu_barrier();
    uint32_t r = g_slot[(g_lane ^ m) & 31];
    emu_barrier();
    return r;
}
DBG_DEV uint32_t ballot(bool p)
{
    g_slot[g_lane] = p ? 1u : 0u;
    emu_barrier();
    uint32_t r = 0;
    for (int i = 0; i < 32; i++) r |= (g_slot[i] & 1u) << i;
    emu_barrier();
    return r;
}
DBG_DEV bool any(bool p) { return ballot(p) != 0; }
DBG_DEV uint32_t match_any(uint32_t v)
{
    g_slot[g_lane] = v;
    emu_barrier();
    uint32_t r = 0;
    for (int i = 0; i < 32; i++) r |= (uint32_t)(g_slot[i] == v) << i;
    emu_barrier();
    return r;
}

DBG_DEV uint32_t brev(uint32_t v)
{
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
    v = ((v >> 8) & 0x00FF00FFu) | ((v & 0x00FF00FFu) << 8);
    return (v >> 16) | (v << 16);
}
DBG_DEV int popc(uint32_t v) { return __builtin_popcount(v); }
DBG_DEV int clz(uint32_t v) { return v ? __builtin_clz(v) : 32; }
DBG_DEV int ffs(uint32_t v) { return __builtin_ffs((int)v); }
DBG_DEV uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh)
{
    sh &= 31;
    return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
}

DBG_DEV void cp_async16(void *smem_dst, const void *gsrc, int src_bytes)
{
    if (src_bytes)
        memcpy(smem_dst, gsrc, 16);
    else
        memset(smem_dst, 0, 16);
}
DBG_DEV uint64_t ld_acquire_u64(const uint64_t *p) { return *p; }
DBG_DEV void st_release_u64(uint64_t *p, uint64_t v) { *p = v; }
DBG_DEV void threadfence() {}
DBG_DEV void backoff() {}
DBG_DEV uint32_t ldg_u32(const uint32_t *p) { return *p; }
DBG_DEV void ldg_u32x4(const uint32_t *p, uint32_t &a, uint32_t &b, uint32_t &c, uint32_t &d) { a = p[0], b = p[1], c = p[2], d = p[3]; }
DBG_DEV uint32_t atomic_inc_shared(uint32_t *p) { return (*p)++; }
DBG_DEV uint32_t ldcg_u32(const uint32_t *p) { return *p; }
DBG_DEV uint32_t ldcg_u8(const uint8_t *p) { return *p; }

DBG_DEV void cp_async16_stream(void *smem_dst, const void *gsrc, int src_bytes) { cp_async16(smem_dst, gsrc, src_bytes); }
DBG_DEV void st_u32x4(uint32_t *p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) { p[0] = a; p[1] = b; p[2] = c; p[3] = d; }
DBG_DEV void cp_async_commit() {}
DBG_DEV void cp_async_wait_all() {}

}  // namespace simt
#endif
