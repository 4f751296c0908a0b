// Shared device helpers for libgsat_b200 (sm_100a only).
#pragma once
#ifdef GSATB_HOST_SIM
// tests/simt/: the same kernel source compiled by g++ against a host SIMT emulator (threads = fibers, real
// __syncthreads / *_sync rendezvous) for the CPU-side kernel-logic tests.  Never defined in the product build.
#include "simt.h"
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>
#include "../../include/gsat_b200.h"

#define GSATB_NUM_SMS 148

#define GSATB_CHECK_LAUNCH()                                   \
    do {                                                       \
        cudaError_t e__ = cudaPeekAtLastError();               \
        if (e__ != cudaSuccess) return GSATB_ELAUNCH;          \
    } while (0)

// Optional device-resident step counter (gsatb_set_step_counter): added to every counter-based random stream (dropout
// hashes, the sampler's Philox offset) INSIDE the kernels, so that a CUDA graph of a whole training step draws fresh
// randomness on every replay although its kernel arguments are frozen.  Host-side accessor, defined in api.cu.
const unsigned long long*& gsatb_step_counter_ref();

// Kernel launch on `stream` (no dynamic shared memory).  KERNEL must be a plain identifier: name a template
// instantiation through a local function pointer first (`auto k = k_foo<4, true>;`).
#ifdef GSATB_HOST_SIM
#define GSATB_LAUNCH(KERNEL, GRID, BLOCK, STREAM, ...) \
    simt::launch(dim3(GRID), dim3(BLOCK), [&]() { KERNEL(__VA_ARGS__); })
#else
#define GSATB_LAUNCH(KERNEL, GRID, BLOCK, STREAM, ...) KERNEL<<<(GRID), (BLOCK), 0, (STREAM)>>>(__VA_ARGS__)
#endif

static inline bool gsatb_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// 128-bit streaming load through the read-only path without L1 allocation (data touched once).
__device__ __forceinline__ float4 ldg_stream_f4(const float4* p) {
#ifdef GSATB_HOST_SIM
    return *p;
#else
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
#endif
}
// 64-bit variant (four bf16)
__device__ __forceinline__ uint2 ldg_stream_u2(const uint2* p) {
#ifdef GSATB_HOST_SIM
    return *p;
#else
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
#endif
}
// 128-bit gather load through the read-only path, L1-allocating (neighbour rows are re-used inside a CTA).
__device__ __forceinline__ float4 ldg_f4(const float4* p) { return __ldg(p); }

__device__ __forceinline__ void fma4(float4& acc, float a, const float4& v) {
    acc.x = fmaf(a, v.x, acc.x);
    acc.y = fmaf(a, v.y, acc.y);
    acc.z = fmaf(a, v.z, acc.z);
    acc.w = fmaf(a, v.w, acc.w);
}
__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
    return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}

// Philox4x32-10 counter-based generator (Salmon et al. 2011), used for in-kernel noise / dropout masks so that
// the backward pass can regenerate what the forward pass drew instead of storing it.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}
// uniform in [1e-10, 1-1e-10] as the reference draws it (src/run_gsat.py:880), from 32 random bits
__device__ __forceinline__ float u01_clamped(uint32_t bits) {
    // 23 random bits + 0.5 is exact in fp32, so u lies in [2^-24, 1 - 2^-24]: log(u) and log(1-u) stay finite
    return (static_cast<float>(bits >> 9) + 0.5f) * (1.0f / 8388608.0f);
}
