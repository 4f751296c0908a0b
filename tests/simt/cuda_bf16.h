// Host stand-in for <cuda_bf16.h> in the SIMT-emulator build (tests/simt): only what csrc/aggregate.cu uses.
#pragma once
#include <cstdint>
#include <cstring>
struct __nv_bfloat16 { uint16_t bits; };
struct alignas(4) __nv_bfloat162 { __nv_bfloat16 x, y; };
static inline __nv_bfloat16 __float2bfloat16_rn(float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return __nv_bfloat16{(uint16_t)0x7fff};        // NaN
    u += 0x7fffu + ((u >> 16) & 1u);                                                    // round to nearest even
    return __nv_bfloat16{(uint16_t)(u >> 16)};
}
static inline __nv_bfloat162 __floats2bfloat162_rn(float a, float b) {
    return __nv_bfloat162{__float2bfloat16_rn(a), __float2bfloat16_rn(b)};
}
static inline float __bfloat162float(__nv_bfloat16 h) {
    uint32_t u = (uint32_t)h.bits << 16;
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}
