// Helpers shared by the tensor-core ops (operand packing, dropout hash, segment tables, uniform tiling).
#pragma once
#include <cuda_bf16.h>
#include "tc_gemm.cuh"

namespace tcg {

// 8 consecutive floats of a row.  Contract of the tensor-core ops: K % 8 == 0, row stride % 4 == 0 and 16-byte
// aligned bases, so the chunk is two aligned 128-bit loads with no tail (checked at the C entry points).
__device__ __forceinline__ void load8_f32(const float* __restrict__ src, int k, int, float v[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src + k));
    const float4 b = __ldg(reinterpret_cast<const float4*>(src + k + 4));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void pack8(const float v[8], uint32_t o[4]) {
    o[0] = tc::pack_bf16(v[0], v[1]);
    o[1] = tc::pack_bf16(v[2], v[3]);
    o[2] = tc::pack_bf16(v[4], v[5]);
    o[3] = tc::pack_bf16(v[6], v[7]);
}
// 8 bf16 (one 16-byte load) -> 8 floats
__device__ __forceinline__ void unpack8(const uint4& q, float v[8]) {
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[2 * i] = __uint_as_float(w[i] << 16);
        v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
}
__device__ __forceinline__ float bf16_bits_to_float(uint16_t b) { return __uint_as_float((uint32_t)b << 16); }
__device__ __forceinline__ uint16_t float_to_bf16_bits(float f) {
    __nv_bfloat16 h = __float2bfloat16_rn(f);
    return *reinterpret_cast<uint16_t*>(&h);
}

// Dropout keep decision for element `idx` of a tensor: counter-based (regenerable in backward), 24-bit threshold.
// mask (uint8, 1 = keep) overrides the hash when given (parity tests inject masks).
struct Dropout {
    const uint8_t* mask;   // nullable
    uint32_t seed;
    uint32_t thr24;        // drop when (hash >> 8) < thr24 ;  thr24 = p * 2^24
    float scale;           // 1 / (1 - p)   (1 when disabled; word mode: 256 / (256 - thr8), the exact keep rate)
    int enabled;
    uint32_t thr8;         // word mode: drop when the element's 8-bit uniform < thr8 ;  thr8 = round(p * 256)
    int word;              // 1: "word" scheme (dropout_word), 0: per-element hash (hash_keep)
    const unsigned long long* step;   // nullable device step counter folded into the seed (gsatb_set_step_counter)
};
__device__ __forceinline__ uint32_t mix32(uint32_t h) {
    h ^= h >> 16;
    h *= 0x85EBCA6Bu;
    h ^= h >> 13;
    h *= 0xC2B2AE35u;
    h ^= h >> 16;
    return h;
}
// element (row, ch) of a [rows, C] tensor; the hash needs no 64-bit arithmetic (rows < 2^31, ch < 2^16)
// ch_term = ch * 0x7FEB352D + seed (hoisted by callers that walk rows of one channel)
__device__ __forceinline__ bool hash_keep(const Dropout& d, uint32_t row, uint32_t ch_term) {
    uint32_t h = row * 0x9E3779B1u + ch_term;
    h ^= h >> 16;
    h *= 0x85EBCA6Bu;
    h ^= h >> 13;
    h *= 0xC2B2AE35u;
    return (h >> 8) >= d.thr24;
}
__device__ __forceinline__ uint32_t hash_ch_term(const Dropout& d, int ch) {
    const uint32_t step = d.step ? (uint32_t)__ldg(d.step) * 0x9E3779B1u : 0u;
    return (uint32_t)ch * 0x7FEB352Du + d.seed + step;
}
__device__ __forceinline__ bool dropout_keep(const Dropout& d, int64_t row, int ch, int C) {
    if (!d.enabled) return true;
    if (d.mask) return __ldg(d.mask + row * C + ch) != 0;
    return hash_keep(d, (uint32_t)row, hash_ch_term(d, ch));
}

// The effective 32-bit seed of this launch (host seed + optional device step counter).
__device__ __forceinline__ uint32_t dropout_seed(const Dropout& d) {
    return d.seed + (d.step ? (uint32_t)__ldg(d.step) * 0x9E3779B1u : 0u);
}
// "Word" scheme: ONE call yields the keep bits of 32 consecutive channels (32*wc .. 32*wc+31) of a row -- bit b = keep
// (row, 32*wc + b).  p = 0.5 is a single hash (every bit of a good hash is a fair coin); any other p compares an 8-bit
// uniform, held bit-sliced in 8 hashes, against thr8 with bitwise ops (p is quantised to 1/256 and the kept values are
// scaled by the exact keep rate, so the estimator stays unbiased).  Per-element hashing cost ~10 instructions per
// element in the issue-bound epilogues; this costs 0.3 (p = 0.5) to 3.
__device__ __forceinline__ uint32_t dropout_word(const Dropout& d, uint32_t row, uint32_t wc, uint32_t seed) {
    const uint32_t base = row * 0x9E3779B1u + wc * 0x7FEB352Du + seed;
    if (d.thr8 == 128u) return mix32(base);
    uint32_t lt = 0u, eq = 0xffffffffu;
#pragma unroll
    for (int i = 7; i >= 0; --i) {
        const uint32_t h = mix32(base + (uint32_t)(i + 1) * 0x632BE5ABu);
        if ((d.thr8 >> i) & 1u) {
            lt |= eq & ~h;
            eq &= h;
        } else {
            eq &= ~h;
        }
    }
    return d.thr8 >= 256u ? 0u : ~lt;      // keep = not (U < thr8)
}
// "Channel word" scheme (fused extractor kernels, whose threads own a CHANNEL and walk rows): ONE call yields the keep
// bits of the 32 rows 32 g .. 32 g + 31 of channel ch (bit b = keep(32 g + b, ch)), thread-local, no warp transpose.
// Same estimator as dropout_word: p = 0.5 is one hash, any other p an 8-bit uniform held bit-sliced in 8 hashes.
__device__ __forceinline__ uint32_t dropout_chan_word(const Dropout& d, uint32_t ch, uint32_t g, uint32_t seed) {
    const uint32_t base = g * 0x9E3779B1u + ch * 0x7FEB352Du + seed;
    if (d.thr8 == 128u) return mix32(base);
    uint32_t lt = 0u, eq = 0xffffffffu;
#pragma unroll
    for (int i = 7; i >= 0; --i) {
        const uint32_t h = mix32(base + (uint32_t)(i + 1) * 0x632BE5ABu);
        if ((d.thr8 >> i) & 1u) {
            lt |= eq & ~h;
            eq &= h;
        } else {
            eq &= ~h;
        }
    }
    return d.thr8 >= 256u ? 0u : ~lt;
}
// keep bits of rows row0 .. row0 + 31 (bit j = keep(row0 + j, ch)) for any row0
__device__ __forceinline__ uint32_t dropout_chan_bits32(const Dropout& d, uint32_t ch, uint32_t row0, uint32_t seed) {
    const uint32_t g = row0 >> 5, sh = row0 & 31u;
    const uint32_t w0 = dropout_chan_word(d, ch, g, seed);
    if (sh == 0) return w0;
    return tc::funnel_r(w0, dropout_chan_word(d, ch, g + 1, seed), sh);
}

// 32 x 32 bit-matrix transpose across a warp: in: lane r holds word r (bit c = element (r, c)); out: lane c holds the
// bits of column c (bit r = element (r, c)).  Five shuffle / mask steps.
__device__ __forceinline__ uint32_t warp_transpose32(uint32_t x, int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const uint32_t mlow = s == 16 ? 0x0000FFFFu : s == 8 ? 0x00FF00FFu : s == 4 ? 0x0F0F0F0Fu : s == 2 ? 0x33333333u : 0x55555555u;
        const uint32_t y = __shfl_xor_sync(0xffffffffu, x, s);
        x = (lane & s) ? ((x & ~mlow) | ((y >> s) & mlow)) : ((x & mlow) | ((y << s) & ~mlow));
    }
    return x;
}
// Keep bits of 32 consecutive ROWS (row0 .. row0+31) for THIS thread's channel, for an epilogue warp whose 32 lanes own
// the 32 channels of word column wc (lane = channel & 31): bit j = keep(row0 + j, channel).  Warp-collective.
__device__ __forceinline__ uint32_t dropout_rows32(const Dropout& d, uint32_t row0, uint32_t wc, uint32_t seed, int lane) {
    return warp_transpose32(dropout_word(d, row0 + (uint32_t)lane, wc, seed), lane);
}

inline Dropout make_dropout(const uint8_t* mask, uint64_t seed, float pdrop, int training, int word = 0) {
    Dropout d;
    d.mask = mask;
    d.seed = (uint32_t)(seed * 0x9E3779B97F4A7C15ull >> 32) ^ (uint32_t)seed;
    d.enabled = training && pdrop > 0.f;
    d.scale = d.enabled ? 1.f / (1.f - pdrop) : 1.f;
    double t = (double)pdrop * 16777216.0;
    d.thr24 = (uint32_t)(t < 0 ? 0 : (t > 16777216.0 ? 16777216.0 : t));
    d.step = gsatb_step_counter_ref();
    d.word = word;
    double t8 = (double)pdrop * 256.0 + 0.5;
    d.thr8 = (uint32_t)(t8 < 0 ? 0 : (t8 > 256.0 ? 256.0 : t8));
    if (word && d.enabled && !mask && d.thr8 < 256u) d.scale = 256.f / (float)(256u - d.thr8);
    return d;
}

inline Tiling uniform_tiling(int64_t rows) {
    Tiling t;
    t.rows = rows;
    t.num_tiles = (int)((rows + TILE_ROWS - 1) / TILE_ROWS);
    t.tile_row = nullptr;
    t.tile_seg = nullptr;
    t.seg_ptr = nullptr;
    t.dbg = nullptr;
    return t;
}

// Stream one accumulator block out as fp32 rows through the group's staging buffer: two buffers of RB rows x 128
// channels x 4 bytes (RB = 32: 32 KiB per group, RB = 16: 16 KiB).  f(col, acc) is the value of element
// (row r0 + col, this thread's channel); it is called, in column order, for all 32 columns of every chunk.
template <int RB, class F>
__device__ __forceinline__ void epi_emit_f32(const EpiCtx& cx, float* out, int ld, F f) {
    constexpr int HALVES = 32 / RB;
    const int nchunks = (cx.cnt + 31) >> 5;
    if (nchunks == 0) epi_release_acc(cx);
#pragma unroll 1
    for (int c = 0; c < nchunks; ++c) {
        float v[32];
        tc::tmem_ld_32x32(cx.taddr + c * 32, v);
        tc::tmem_ld_wait();
        if (c == nchunks - 1) epi_release_acc(cx);
#pragma unroll
        for (int h = 0; h < HALVES; ++h) {
            const int lo = c * 32 + h * RB, n = min(RB, cx.cnt - lo);
            // f runs for every column of the chunk (it may be warp-collective / keep per-chunk state); rows past the
            // tile end are computed and dropped
#pragma unroll
            for (int j = 0; j < RB; ++j) v[h * RB + j] = f(lo + j, v[h * RB + j]);
            if (n > 0) {       // uniform across the group
                float* buf = reinterpret_cast<float*>(cx.stage) + (HALVES == 1 ? (c & 1) : h) * (RB * 128);
#pragma unroll
                for (int j = 0; j < RB; ++j) buf[j * 128 + cx.gtid] = v[h * RB + j];
                epi_sync(cx);  // buffer staged; also orders the previous copy-out of the OTHER buffer before its re-use
                stage_store<4>(cx, reinterpret_cast<const uint8_t*>(buf), out + cx.r0 * ld, ld, lo, n);
            }
        }
    }
}

// Same, bf16 output: two buffers of 32 rows x 128 channels x 2 bytes (16 KiB per group).
template <class F>
__device__ __forceinline__ void epi_emit_bf16(const EpiCtx& cx, uint16_t* out, int ld, F f) {
    const int nchunks = (cx.cnt + 31) >> 5;
    if (nchunks == 0) epi_release_acc(cx);
#pragma unroll 1
    for (int c = 0; c < nchunks; ++c) {
        float v[32];
        tc::tmem_ld_32x32(cx.taddr + c * 32, v);
        tc::tmem_ld_wait();
        if (c == nchunks - 1) epi_release_acc(cx);
        uint16_t* buf = reinterpret_cast<uint16_t*>(cx.stage) + (c & 1) * (32 * 128);
#pragma unroll
        for (int j = 0; j < 32; ++j) buf[j * 128 + cx.gtid] = float_to_bf16_bits(f(c * 32 + j, v[j]));
        epi_sync(cx);
        stage_store<2>(cx, reinterpret_cast<const uint8_t*>(buf), out + cx.r0 * ld, ld, c * 32, min(32, cx.cnt - c * 32));
    }
}

// Per-warp copy of the tile's graph boundaries (local row offsets) in the epilogue scratch: bnd[0..nseg]
__device__ __forceinline__ int load_segments(const Tiling& tl, int tile, int64_t r0, uint8_t* misc, int q, int lane,
                                             const int*& bnd, int& g0) {
    int* mine = reinterpret_cast<int*>(misc) + q * (MAX_SEG + 2);
    g0 = __ldg(tl.tile_seg + tile);
    const int nseg = __ldg(tl.tile_seg + tile + 1) - g0;
    __syncwarp();
    for (int s = lane; s <= nseg; s += 32) mine[s] = __ldg(tl.seg_ptr + g0 + s) - (int)r0;
    __syncwarp();
    bnd = mine;
    return nseg;
}

__device__ __forceinline__ int load_segments(const Tiling& tl, const EpiCtx& cx, const int*& bnd, int& g0) {
    return load_segments(tl, cx.tile, cx.r0, cx.misc, cx.q, cx.lane, bnd, g0);
}

}  // namespace tcg
