// Fused extractor MLP, backward, ONE persistent kernel (autograd of src/run_gsat.py:909-927 + src/utils/get_model.py:57-68
// as reached by loss.backward() at src/run_gsat.py:634), over the same graph-aligned slot tiles as the forward kernel.
//
// Nothing of width 4H was saved by the forward: per tile the kernel
//   head : x^2 (InstanceNorm-2 output, bf16, slot space) arrives by TMA; with d logit and w3 the epilogue threads (one per
//          channel) back-propagate Linear(H,1), Dropout 2, ReLU 2 and InstanceNorm 2 -> dz2 (in place, an MN-major B tile),
//          dw3 partial sums;
//   per 128-channel block cb:
//          GEMM1 is RECOMPUTED from the re-gathered, centred input tile (the tensor pipe is mostly idle) -> z~1,
//          dh1 = W2^T[cb] dz2 on tcgen05 from the MN-major dz2 tile; the epilogue (thread = channel, a whole graph of
//          both accumulators resident in registers) rebuilds x^1, the ReLU / Dropout-1 gates, back-propagates
//          InstanceNorm 1 (sum dy, sum dy x^ thread-local) -> dz1, and re-materialises h1 = Dropout(ReLU(x^1));
//          dz1 goes to an MN-major tile (B operand of dx) and, with h1, to HBM in CHANNEL-major slot space
//          [C1, tiles * 128] -- the operands of the weight-gradient products (gsatb_tc_dw), written as 16-byte vectors;
//   dx   : d f12^T += W1^T[:, cb] dz1 accumulated over the channel blocks in TMEM; its epilogue writes d f12 [rows, Kin]
//          (fp32, coalesced) for the deterministic scatter back to the nodes (gsatb_gather_concat_bwd).
// The gather producers also dump the centred bf16 input tile as xs [tiles * 128, pad64(Kin)]: the B operand of dW1.
// b1 / b2 get exact zeros (they cancel in the InstanceNorms); centring needs no backward (sum_g dz1 = 0 per channel).
//
// Roles (16 warps): 0 weight TMA, 1 MMA issuer, 2 TMEM allocator, 3 x^2-tile TMA, 4-11 two epilogue warpgroups (they
// split every step of a tile by graph / by 128-channel block), 12-15 gather producers.
#include "ext_fused.cuh"

namespace {

using namespace extf;

struct BwdParams {
    const int32_t* seg_ptr;
    const int32_t* tile_seg;
    const int32_t* num_tiles;
    const float* dlogit;         // [rows]
    const float* w3;             // [H]
    const float* rstd2;          // [G, H]
    Dropout drop1, drop2;
    const uint32_t* seeds;       // [2] effective seeds written by the forward (null: injected masks / no dropout)
    uint16_t* dz2t;              // bf16 slot space, tile-major [tiles][pad128(H)][128]
    uint16_t* dz1t;              // bf16 [tiles][pad128(C1)][128]
    uint16_t* h1t;               // bf16 [tiles][pad128(C1)][128]
    float* df12;                 // [rows, Kin] fp32, or bf16 when df12_bf16
    int df12_bf16;
    float* dw3_part;             // [grid, H]
    int64_t ld_slots;
    int ldx;
    int H, Kin, C1, KB1, KBH, NCB, KM, NW;
    int HP, C1P;                 // pad128(H), pad128(C1): channel rows per tile block of the slot-space tensors
    int xkb;
    float eps;
    long long* dbg;
};

struct Smem {
    uint32_t ring, x, dz2, dz1, h1, dl, scr, bars, total;
};
__host__ __device__ inline Smem smem_plan(int KB1, int NW, int xkb) {
    Smem s;
    s.ring = 0;
    s.x = s.ring + (uint32_t)NW * BRICK;
    s.dz2 = s.x + (uint32_t)KB1 * xkb;
    s.dz1 = s.dz2 + 2 * BRICK;
    s.h1 = s.dz1 + 2 * BRICK;                // dz1 tile (B operand of dx + TMA store source), h1 tile (TMA store source)
    s.dl = s.h1 + 2 * BRICK;
    s.scr = s.dl + 2 * 512;                  // d logit per slot, one copy per epilogue warpgroup
    s.bars = s.scr;
    s.total = s.bars + 512 + 1024;
    return s;
}

__device__ __forceinline__ int kblocks_in_cb(int C1, int cb) {      // 64-channel K blocks inside channel block cb
    const int left = ((C1 + 63) / 64) * 64 - cb * 128;
    return left >= 128 ? 2 : (left > 0 ? 1 : 0);
}

constexpr int BAR_EPI = 2;                   // named barriers: BAR_EPI + e (+ 8)
constexpr int BWD_THREADS = 384;             // 12 warps: 4 control, 2 epilogue warpgroups
constexpr int BWD_EPI_REGS = 200;            // 128 x 56 + 256 x 200 = 58368 <= 384 x 168

#define EXT_DISPATCH_NB7(nb, CALL)                          \
    switch (nb) {                                           \
        case 1: { constexpr int NB = 1; CALL; } break;      \
        case 2: { constexpr int NB = 2; CALL; } break;      \
        case 3: { constexpr int NB = 3; CALL; } break;      \
        case 4: { constexpr int NB = 4; CALL; } break;      \
        case 5: { constexpr int NB = 5; CALL; } break;      \
        case 6: { constexpr int NB = 6; CALL; } break;      \
        default: { constexpr int NB = 7; CALL; } break;     \
    }
#define EXT_DISPATCH_NB8(nb, CALL)                          \
    switch (nb) {                                           \
        case 1: { constexpr int NB = 1; CALL; } break;      \
        case 2: { constexpr int NB = 2; CALL; } break;      \
        case 3: { constexpr int NB = 3; CALL; } break;      \
        case 4: { constexpr int NB = 4; CALL; } break;      \
        case 5: { constexpr int NB = 5; CALL; } break;      \
        case 6: { constexpr int NB = 6; CALL; } break;      \
        case 7: { constexpr int NB = 7; CALL; } break;      \
        default: { constexpr int NB = 8; CALL; } break;     \
    }
#define EXT_DISPATCH_NB4(nb, CALL)                          \
    switch (nb) {                                           \
        case 1: { constexpr int NB = 1; CALL; } break;      \
        case 2: { constexpr int NB = 2; CALL; } break;      \
        case 3: { constexpr int NB = 3; CALL; } break;      \
        default: { constexpr int NB = 4; CALL; } break;     \
    }

// ---- head: backward through Linear(H,1), Dropout 2, ReLU 2, InstanceNorm 2, for one graph of one channel -----------
struct HeadCtx {
    uint32_t tile_s;     // shared address of the x^2 / dz2 tile row of this channel: base + ch * 128 handled by offsets
    uint32_t dl_s;       // shared address of d logit per slot [128] floats
    int ch, gtid;
    bool ch_ok;
    float w3s;           // w3[ch] * dropout scale
};
// one graph (NB <= 8 blocks) of one channel, register resident: x^2 (bf16 -> fp32) and g = d xhat2 stay in registers
// between the statistics pass and the emitting pass
template <int NB>
__device__ __forceinline__ void head_graph(const HeadCtx& c, const DropCtx& dc, int n, int row0, int slot0, float r, float& dw) {
    float x[8 * NB], g[8 * NB];
    const int blk0 = slot0 >> 3;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const uint4 xq = tc::lds128(c.tile_s + mn_tile_offset_blk(c.gtid, blk0 + b));
        unpack8(xq, &x[8 * b]);
    }
    const uint32_t k0 = keep_bits32(dc, c.ch, c.ch_ok, row0, n);
    uint32_t k1 = 0xffffffffu;
    if (NB > 4) k1 = keep_bits32(dc, c.ch, c.ch_ok, row0 + 32, n - 32);
    float s1 = 0.f, s2 = 0.f, dwt = 0.f;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const uint4 d0 = tc::lds128(c.dl_s + 32 * (blk0 + b)), d1 = tc::lds128(c.dl_s + 32 * (blk0 + b) + 16);
        const float dl[8] = {__uint_as_float(d0.x), __uint_as_float(d0.y), __uint_as_float(d0.z), __uint_as_float(d0.w),
                             __uint_as_float(d1.x), __uint_as_float(d1.y), __uint_as_float(d1.z), __uint_as_float(d1.w)};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int j = 8 * b + i;
            const uint32_t kb = (j < 32 ? k0 >> j : k1 >> (j - 32)) & 1u;
            // padding slots: xhat2 = 0 and d logit = 0 were written by the forward / the producers -> gate off, g = 0
            const float xr = (x[j] > 0.f && kb != 0u) ? x[j] : 0.f;          // relu(xhat2) * keep
            dwt = fmaf(dl[i], xr, dwt);                                      // d w3 / scale
            const float gv = xr > 0.f ? dl[i] * c.w3s : 0.f;                 // d xhat2
            s1 += gv;
            s2 = fmaf(gv, x[j], s2);
            g[j] = gv;
        }
    }
    dw += dwt;
    const float inv_n = 1.f / (float)n;
    const float c0 = -s1 * inv_n * r, c1 = -s2 * inv_n * r;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int j = 8 * b + i;
            o[i] = fmaf(x[j], c1, fmaf(g[j], r, c0));                         // r (g - m1 - xhat m2)
            if (b == NB - 1) o[i] = i < n - 8 * b ? o[i] : 0.f;               // padding slots stay zero
        }
        tc::sts128(c.tile_s + mn_tile_offset_blk(c.gtid, blk0 + b), tc::pack_bf16(o[0], o[1]), tc::pack_bf16(o[2], o[3]),
                   tc::pack_bf16(o[4], o[5]), tc::pack_bf16(o[6], o[7]));
    }
}
// larger graphs: block by block from shared memory, two passes
__device__ __forceinline__ void head_blk_stats(const HeadCtx& c, int blk, int nv, uint32_t bits, float& s1, float& s2, float& dw) {
    const uint4 xq = tc::lds128(c.tile_s + mn_tile_offset_blk(c.gtid, blk));
    const uint4 d0 = tc::lds128(c.dl_s + 32 * blk), d1 = tc::lds128(c.dl_s + 32 * blk + 16);
    float x[8];
    unpack8(xq, x);
    const float dl[8] = {__uint_as_float(d0.x), __uint_as_float(d0.y), __uint_as_float(d0.z), __uint_as_float(d0.w),
                         __uint_as_float(d1.x), __uint_as_float(d1.y), __uint_as_float(d1.z), __uint_as_float(d1.w)};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const bool gate = j < nv && x[j] > 0.f && ((bits >> j) & 1u);
        const float g = gate ? dl[j] * c.w3s : 0.f;
        s1 += g;
        s2 = fmaf(g, x[j], s2);
        dw = fmaf(gate ? dl[j] : 0.f, x[j], dw);
    }
}
__device__ __forceinline__ void head_blk_emit(const HeadCtx& c, int blk, int nv, uint32_t bits, float r, float c0, float c1) {
    const uint32_t sa = c.tile_s + mn_tile_offset_blk(c.gtid, blk);
    const uint4 xq = tc::lds128(sa);
    const uint4 d0 = tc::lds128(c.dl_s + 32 * blk), d1 = tc::lds128(c.dl_s + 32 * blk + 16);
    float x[8], o[8];
    unpack8(xq, x);
    const float dl[8] = {__uint_as_float(d0.x), __uint_as_float(d0.y), __uint_as_float(d0.z), __uint_as_float(d0.w),
                         __uint_as_float(d1.x), __uint_as_float(d1.y), __uint_as_float(d1.z), __uint_as_float(d1.w)};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const bool gate = x[j] > 0.f && ((bits >> j) & 1u);
        const float g = gate ? dl[j] * c.w3s : 0.f;
        o[j] = j < nv ? fmaf(x[j], c1, fmaf(g, r, c0)) : 0.f;
    }
    tc::sts128(sa, tc::pack_bf16(o[0], o[1]), tc::pack_bf16(o[2], o[3]), tc::pack_bf16(o[4], o[5]), tc::pack_bf16(o[6], o[7]));
}

// ---- epilogue 3: one graph of one channel of a 128-channel block, both accumulators resident in registers ----------
struct Epi3Ctx {
    uint32_t t1, t3;     // TMEM addresses (this warp's lane quarter) of z~1 (GEMM1 recompute) and dh1
    uint32_t dz1_s, h1_s;   // shared addresses of the dz1 and h1 tiles
    int ch, gtid;
    bool ch_ok;
    float eps, dscale;
};
// pass 2 of 8 slots: dz1 = r (dy - m1 - xhat m2), h1 = Dropout(ReLU(xhat)) * scale; xh = xhat, dy = d xhat (0 where the
// gate is off), sg = scale where the gate is on else 0
template <bool LAST>
__device__ __forceinline__ void epi3_emit_blk(const Epi3Ctx& c, const float* xh, const float* dy, const uint32_t kbits, float r,
                                              float c0, float c1, int nv, int blk) {
    uint32_t o[4], hq[4];
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
        float a0 = fmaf(xh[i], c1, fmaf(dy[i], r, c0)), a1 = fmaf(xh[i + 1], c1, fmaf(dy[i + 1], r, c0));
        if (LAST) {
            a0 = i < nv ? a0 : 0.f;
            a1 = i + 1 < nv ? a1 : 0.f;
        }
        o[i >> 1] = tc::pack_bf16(a0, a1);
        // gate = relu on (xhat has the sign of z~: rstd > 0) and kept; padding slots: xhat = 0 -> off
        const float h0 = (xh[i] > 0.f && ((kbits >> i) & 1u)) ? xh[i] * c.dscale : 0.f;
        const float h1v = (xh[i + 1] > 0.f && ((kbits >> (i + 1)) & 1u)) ? xh[i + 1] * c.dscale : 0.f;
        hq[i >> 1] = tc::pack_bf16(h0, h1v);
    }
    tc::sts128(c.dz1_s + mn_tile_offset_blk(c.gtid, blk), o[0], o[1], o[2], o[3]);
    tc::sts128(c.h1_s + mn_tile_offset_blk(c.gtid, blk), hq[0], hq[1], hq[2], hq[3]);
}
template <int NB, class WaitTiles, class Loaded>
__device__ __forceinline__ void epi3_graph(const Epi3Ctx& c, const DropCtx& dc, int n, int row0, int slot0, WaitTiles wait_tiles,
                                           Loaded loaded) {
    float z[8 * NB], d[8 * NB];
    tmem_ld_blocks<NB>(c.t1 + slot0, z);
    tmem_ld_blocks<NB>(c.t3 + slot0, d);
    const uint32_t k0 = keep_bits32(dc, c.ch, c.ch_ok, row0, n);
    uint32_t k1 = 0xffffffffu;
    if (NB > 4) k1 = keep_bits32(dc, c.ch, c.ch_ok, row0 + 32, n - 32);
    tc::tmem_ld_wait();
    loaded();      // both accumulators of this graph now live in registers: the caller may hand them back to the MMA issuer
    const float inv_n = 1.f / (float)n;
    const float r = 1.f / sqrtf(sumsq_blocks<NB>(z) * inv_n + c.eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < 8 * NB; ++j) {
        const uint32_t kb = (j < 32 ? k0 >> j : k1 >> (j - 32)) & 1u;
        const float sg = (z[j] > 0.f && kb != 0u) ? c.dscale : 0.f;      // padding slots: z~ = 0 -> gate off
        const float xh = z[j] * r, dy = d[j] * sg;                        // xhat1, d xhat1
        s1 += dy;
        s2 = fmaf(dy, xh, s2);
        z[j] = xh;
        d[j] = dy;
    }
    const float c0 = -s1 * inv_n * r, c1 = -s2 * inv_n * r;
    wait_tiles();                                                         // the dz1 / h1 tiles of the previous block are free
    const int blk0 = slot0 >> 3;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const uint32_t kbits = b < 4 ? k0 >> (8 * b) : k1 >> (8 * (b - 4));
        if (b == NB - 1) epi3_emit_blk<true>(c, &z[8 * b], &d[8 * b], kbits, r, c0, c1, n - 8 * b, blk0 + b);
        else epi3_emit_blk<false>(c, &z[8 * b], &d[8 * b], kbits, r, c0, c1, 8, blk0 + b);
    }
}
// larger graphs: three passes over chunks of <= 4 blocks (the accumulators are re-read)
template <int NB>
__device__ __forceinline__ float epi3_chunk_sumsq(const Epi3Ctx& c, int slot) {
    float z[8 * NB];
    tmem_ld_blocks<NB>(c.t1 + slot, z);
    tc::tmem_ld_wait();
    return sumsq_blocks<NB>(z);
}
template <int NB>
__device__ __forceinline__ void epi3_chunk_sums(const Epi3Ctx& c, const DropCtx& dc, float r, int nleft, int row, int slot,
                                                float& s1, float& s2) {
    float z[8 * NB], d[8 * NB];
    tmem_ld_blocks<NB>(c.t1 + slot, z);
    tmem_ld_blocks<NB>(c.t3 + slot, d);
    const uint32_t k0 = keep_bits32(dc, c.ch, c.ch_ok, row, nleft);
    tc::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 8 * NB; ++j) {
        const bool gate = z[j] > 0.f && ((k0 >> j) & 1u);
        const float dy = gate ? d[j] * c.dscale : 0.f;
        s1 += dy;
        s2 = fmaf(dy, z[j] * r, s2);
    }
}
template <int NB>
__device__ __forceinline__ void epi3_chunk_emit(const Epi3Ctx& c, const DropCtx& dc, float r, float c0, float c1, int nleft,
                                                int row, int slot) {
    float z[8 * NB], d[8 * NB];
    tmem_ld_blocks<NB>(c.t1 + slot, z);
    tmem_ld_blocks<NB>(c.t3 + slot, d);
    const uint32_t k0 = keep_bits32(dc, c.ch, c.ch_ok, row, nleft);
    tc::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 8 * NB; ++j) {
        const float sg = (z[j] > 0.f && ((k0 >> j) & 1u)) ? c.dscale : 0.f;
        z[j] *= r;
        d[j] *= sg;
    }
#pragma unroll
    for (int b = 0; b < NB; ++b) epi3_emit_blk<true>(c, &z[8 * b], &d[8 * b], k0 >> (8 * b), r, c0, c1, nleft - 8 * b, (slot >> 3) + b);
}

// d f12 rows of one graph for one input channel: accumulator columns -> coalesced fp32 stores (lane = channel)
template <int NB>
__device__ __forceinline__ void dx_chunk_out(uint32_t taddr, float* out, uint16_t* out16, int Kin, int nleft) {
    float v[8 * NB];
    tmem_ld_blocks<NB>(taddr, v);
    tc::tmem_ld_wait();
    if (out) {
#pragma unroll
        for (int j = 0; j < 8 * NB; ++j)
            if (j < nleft) out[(int64_t)j * Kin] = v[j];
    } else if (out16) {          // bf16 d f12 (edge mode of the bf16 precision mode: gsatb_gather_concat_bwd_bf16 reduces it)
#pragma unroll
        for (int j = 0; j < 8 * NB; ++j)
            if (j < nleft) out16[(int64_t)j * Kin] = float_to_bf16_bits(v[j]);
    }
}

__global__ void __launch_bounds__(BWD_THREADS, 1)
k_ext_fused_bwd(const __grid_constant__ CUtensorMap tm_w1, const __grid_constant__ CUtensorMap tm_w2t,
                const __grid_constant__ CUtensorMap tm_w1t, const __grid_constant__ CUtensorMap tm_x2,
                const __grid_constant__ CUtensorMap tm_dz2, const __grid_constant__ CUtensorMap tm_dz1,
                const __grid_constant__ CUtensorMap tm_h1, const __grid_constant__ CUtensorMap tm_xs, const BwdParams p) {
#ifdef GSATB_HOST_SIM
    uint8_t* smem_raw = simt::dyn_smem();
#else
    extern __shared__ uint8_t smem_raw[];
#endif
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const Smem L = smem_plan(p.KB1, p.NW, p.xkb);
    uint8_t* ring = smem + L.ring;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
    uint64_t* w_full = bars;                 // [8]
    uint64_t* w_empty = bars + 8;            // [8]
    uint64_t* x_full = bars + 16;            // TMA tx: the centred input tile (from xs) landed
    uint64_t* x_empty = bars + 17;           // commit: the tile's last GEMM1 has read the x tile
    uint64_t* x2_full = bars + 18;           // TMA tx: x^2 tile landed
    uint64_t* dz2_ready = bars + 19;         // 256 epilogue threads: dz2 tile complete
    uint64_t* dz2_empty = bars + 20;         // commit: the tile's last dh1 GEMM has read the dz2 tile
    uint64_t* acc_full = bars + 21;          // commit: z~1 and dh1 of a channel block complete
    uint64_t* acc_empty = bars + 22;         // 256: both accumulators read
    uint64_t* dz1_full = bars + 23;          // 256: dz1 and h1 tiles complete
    uint64_t* dz1_empty = bars + 25;         // commit: the dx GEMM and the TMA stores of that block have read them
    uint64_t* dx_full = bars + 27;           // commit: d f12^T complete
    uint64_t* dx_empty = bars + 28;          // 256: d f12^T read
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 30);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T = __ldg(p.num_tiles);
    const uint32_t x_s = tc::smem_u32(smem + L.x), dz2_s = tc::smem_u32(smem + L.dz2), dz1_s = tc::smem_u32(smem + L.dz1),
                   h1_s = tc::smem_u32(smem + L.h1), dl_s = tc::smem_u32(smem + L.dl);

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&tm_w1);
        tc::tma_prefetch_desc(&tm_w2t);
        tc::tma_prefetch_desc(&tm_w1t);
        tc::tma_prefetch_desc(&tm_x2);
        tc::tma_prefetch_desc(&tm_dz2);
        tc::tma_prefetch_desc(&tm_dz1);
        tc::tma_prefetch_desc(&tm_h1);
        for (int i = 0; i < 8; ++i) {
            tc::mbar_init(&w_full[i], 1);
            tc::mbar_init(&w_empty[i], 1);
        }
        tc::mbar_init(x_full, 1);
        tc::mbar_init(x_empty, 1);
        tc::mbar_init(x2_full, 1);
        tc::mbar_init(dz2_ready, 256);
        tc::mbar_init(dz2_empty, 1);
        tc::mbar_init(acc_full, 1);
        tc::mbar_init(acc_empty, 256);
        tc::mbar_init(dz1_full, 256);
        tc::mbar_init(dz1_empty, 1);
        tc::mbar_init(dx_full, 1);
        tc::mbar_init(dx_empty, 256);
        tc::fence_barrier_init();
    }
    if (warp == 2) {
        tc::tmem_alloc(tmem_slot, 512);
        tc::tmem_relinquish();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // TMEM columns: [0,128) z~1, [128,256) dh1, [256,512) d f12^T (KM blocks of 128 input channels)

    if (warp < 4) {
        tc::reg_dec<EXT_CTL_REGS>();
        if (warp == 0) {
            // ===================== weight bricks (TMA), in MMA consumption order =====================
            if (lane == 0) {
                uint32_t cw = 0;
                auto put = [&](const CUtensorMap* tm, int k_elem, int row) {
                    const uint32_t s = cw % p.NW, use = cw / p.NW;
                    tc::mbar_wait(&w_empty[s], (use & 1) ^ 1);
                    tc::mbar_arrive_expect_tx(&w_full[s], BRICK);
                    tc::tma_load_2d(ring + s * BRICK, tm, &w_full[s], k_elem, row);
                    ++cw;
                };
                auto put_dx = [&](int cb) {
                    const int nk = kblocks_in_cb(p.C1, cb);
                    for (int mb = 0; mb < p.KM; ++mb)
                        for (int kb = 0; kb < nk; ++kb) put(&tm_w1t, cb * 128 + kb * 64, mb * 128);
                };
                for (int tile = blockIdx.x; tile < T; tile += gridDim.x) {
                    for (int cb = 0; cb < p.NCB; ++cb) {
                        for (int kb = 0; kb < p.KB1; ++kb) put(&tm_w1, kb * 64, cb * 128);
                        for (int kb = 0; kb < p.KBH; ++kb) put(&tm_w2t, kb * 64, cb * 128);
                        if (cb > 0) put_dx(cb - 1);
                    }
                    put_dx(p.NCB - 1);
                }
            }
        } else if (warp == 1) {
            // ===================== MMA issuer =====================
            if (lane == 0) {
                uint32_t cw = 0, n = 0, ti = 0;
                long long t_all = clock64(), w_in = 0, w_acc = 0, w_w = 0, w_dz1 = 0, w_dx = 0, t0;
                for (int tile = blockIdx.x; tile < T; tile += gridDim.x, ++ti) {
                    int N = pad16(tile_total_slots(p.tile_seg, p.seg_ptr, tile));
                    if (N < 16) N = 16;
                    const uint32_t idesc_k = tc::make_idesc_bf16(128, N, 0, 0), idesc_mn = tc::make_idesc_bf16(128, N, 0, 1);
                    auto brick = [&]() -> uint64_t {
                        const uint32_t s = cw % p.NW, usew = cw / p.NW;
                        t0 = clock64();
                        tc::mbar_wait(&w_full[s], usew & 1);
                        w_w += clock64() - t0;
                        tc::tc_fence_after();
                        return tc::make_desc_k_sw128(tc::smem_u32(ring + s * BRICK));
                    };
                    auto release = [&]() {
                        tc::mma_commit(&w_empty[cw % p.NW]);
                        ++cw;
                    };
                    auto dx_step = [&](int cb, uint32_t nn) {      // d f12^T += W1^T[:, cb] dz1(cb); dz1 / h1 tiles -> HBM
                        t0 = clock64();
                        tc::mbar_wait(dz1_full, nn & 1);
                        w_dz1 += clock64() - t0;
                        t0 = clock64();
                        if (cb == 0) tc::mbar_wait(dx_empty, (ti & 1) ^ 1);
                        w_dx += clock64() - t0;
                        tc::tc_fence_after();
                        for (int sl = 0; sl < 2; ++sl) {              // channel-major slot-space operands of dW1 / dW2
                            tc::tma_store_2d(&tm_dz1, smem + L.dz1 + sl * BRICK, sl * 64, tile * p.C1P + cb * 128);
                            tc::tma_store_2d(&tm_h1, smem + L.h1 + sl * BRICK, sl * 64, tile * p.C1P + cb * 128);
                        }
                        tc::tma_store_commit();
                        const int nk = kblocks_in_cb(p.C1, cb);
                        for (int mb = 0; mb < p.KM; ++mb)
                            for (int kb = 0; kb < nk; ++kb) {
                                const uint64_t a_desc = brick();
#pragma unroll
                                for (int k4 = 0; k4 < 4; ++k4)
                                    tc::mma_bf16_ss(tmem_base + 256 + mb * 128, a_desc + (uint64_t)(k4 * 2),
                                                    tc::make_desc_mn_sw128(dz1_s + (kb * 64 + k4 * 16) * 128, BRICK), idesc_mn,
                                                    (cb | kb | k4) != 0);
                                release();
                            }
                        tc::tma_store_wait_read<0>();                 // the stores have read the tiles
                        tc::mma_commit(dz1_empty);
                    };
                    t0 = clock64();
                    tc::mbar_wait(x_full, ti & 1);
                    w_in += clock64() - t0;
                    tc::tc_fence_after();
                    for (int cb = 0; cb < p.NCB; ++cb, ++n) {
                        t0 = clock64();
                        tc::mbar_wait(acc_empty, (n & 1) ^ 1);
                        w_acc += clock64() - t0;
                        tc::tc_fence_after();
                        for (int kb = 0; kb < p.KB1; ++kb) {          // z~1 = W1[cb] x~^T   (recomputed)
                            const uint64_t a_desc = brick();
                            const uint64_t b_desc = tc::make_desc_k_sw128(x_s + kb * p.xkb);
#pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4)
                                tc::mma_bf16_ss(tmem_base, a_desc + (uint64_t)(k4 * 2), b_desc + (uint64_t)(k4 * 2), idesc_k,
                                                (kb | k4) != 0);
                            release();
                        }
                        if (cb == 0) {      // GEMM1 of the first block ran underneath the head epilogue; dh1 needs its dz2 tile
                            t0 = clock64();
                            tc::mbar_wait(dz2_ready, ti & 1);
                            w_in += clock64() - t0;
                            tc::tc_fence_after();
                            tc::tma_store_2d(&tm_dz2, smem + L.dz2, 0, tile * p.HP);
                            tc::tma_store_2d(&tm_dz2, smem + L.dz2 + BRICK, 64, tile * p.HP);
                            tc::tma_store_commit();
                        }
                        for (int kb = 0; kb < p.KBH; ++kb) {          // dh1 = W2^T[cb] dz2
                            const uint64_t a_desc = brick();
#pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4)
                                tc::mma_bf16_ss(tmem_base + 128, a_desc + (uint64_t)(k4 * 2),
                                                tc::make_desc_mn_sw128(dz2_s + (kb * 64 + k4 * 16) * 128, BRICK), idesc_mn,
                                                (kb | k4) != 0);
                            release();
                        }
                        tc::mma_commit(acc_full);
                        if (cb == p.NCB - 1) {
                            tc::mma_commit(x_empty);
                            tc::tma_store_wait_read<0>();             // (the dz2 store has long finished reading the tile)
                            tc::mma_commit(dz2_empty);
                        }
                        if (cb > 0) dx_step(cb - 1, n - 1);
                    }
                    dx_step(p.NCB - 1, n - 1);
                    tc::mma_commit(dx_full);
                }
                if (p.dbg) {
                    long long* d = p.dbg + (size_t)blockIdx.x * 16;
                    d[0] = clock64() - t_all, d[1] = w_in, d[2] = w_acc, d[3] = w_w, d[4] = w_dz1, d[5] = w_dx;
                }
            }
        } else if (warp == 3) {
            // ===================== activations by TMA: the centred input tile (xs, written by the forward: boxes of 64
            // input channels x max_slots rows) and the x^2 tile ([H channels][slots]: boxes of 64 slots x 128 channels) ====
            if (lane == 0) {
                uint32_t ti = 0;
                const int cap = p.xkb >> 7;
                for (int tile = blockIdx.x; tile < T; tile += gridDim.x, ++ti) {
                    tc::mbar_wait(x_empty, (ti & 1) ^ 1);
                    tc::mbar_arrive_expect_tx(x_full, (uint32_t)p.KB1 * p.xkb);
                    for (int kb = 0; kb < p.KB1; ++kb)
                        tc::tma_load_2d(smem + L.x + (size_t)kb * p.xkb, &tm_xs, x_full, kb * 64, tile * TILE_SLOTS);
                    (void)cap;
                    tc::mbar_wait(dz2_empty, (ti & 1) ^ 1);
                    tc::mbar_arrive_expect_tx(x2_full, 2 * BRICK);
                    tc::tma_load_2d(smem + L.dz2, &tm_x2, x2_full, 0, tile * p.HP);
                    tc::tma_load_2d(smem + L.dz2 + BRICK, &tm_x2, x2_full, 64, tile * p.HP);
                }
            }
        }
    } else {
        // ===================== epilogue warpgroups =====================
        tc::reg_inc<BWD_EPI_REGS>();
        const int e = (warp - 4) >> 2, q = warp & 3, gtid = q * 32 + lane;
        const uint32_t seed1 = p.seeds ? __ldg(p.seeds) : 0u, seed2 = p.seeds ? __ldg(p.seeds + 1) : 0u;
        const float w3 = gtid < p.H ? __ldg(p.w3 + gtid) : 0.f;
        float dw3 = 0.f;
        uint32_t n = 0, ti = 0;
        long long w_head = 0, t_head = 0, w_cb = 0, t_cb = 0, w_dxf = 0, t_dx = 0, t0;
        // d f12^T -> d f12 rows of one tile (warpgroup e takes the 128-channel blocks mb % 2 == e).  Deferred: it runs AFTER the
        // head of the NEXT tile, so the wait for the tile's last dx GEMM overlaps that head instead of stalling the epilogue.
        auto dx_out = [&](const SegTable& tq, uint32_t tiq) {
            t0 = clock64();
            tc::group_mbar_wait(gtid == 0, dx_full, tiq & 1, BAR_EPI + e, 128);
            w_dxf += clock64() - t0;
            t0 = clock64();
            tc::tc_fence_after();
            for (int mb = e; mb < p.KM; mb += 2) {
                const int k = mb * 128 + gtid;
                const uint32_t taddr = tmem_base + 256 + mb * 128 + ((uint32_t)(q * 32) << 16);
                for (int s = 0; s < tq.nseg; ++s) {
                    const int ns = __shfl_sync(0xffffffffu, tq.n, s);
                    if (ns == 0) continue;
                    const int nblk = pad8(ns) >> 3;
                    const int slot0 = __shfl_sync(0xffffffffu, tq.slot0, s), row0 = __shfl_sync(0xffffffffu, tq.row0, s);
                    for (int c4 = 0; c4 < nblk; c4 += 4) {
                        const int nbk = nblk - c4 < 4 ? nblk - c4 : 4;
                        const int64_t o = (int64_t)(row0 + 8 * c4) * p.Kin + k;
                        float* out = (k < p.Kin && !p.df12_bf16) ? p.df12 + o : nullptr;
                        uint16_t* out16 = (k < p.Kin && p.df12_bf16) ? reinterpret_cast<uint16_t*>(p.df12) + o : nullptr;
                        EXT_DISPATCH_NB4(nbk, (dx_chunk_out<NB>(taddr + slot0 + 8 * c4, out, out16, p.Kin, ns - 8 * c4)));
                    }
                }
            }
            tc::tc_fence_before();
            tc::mbar_arrive(dx_empty);
            t_dx += clock64() - t0;
        };
        SegTable tb_prev;
        bool have_prev = false;
        for (int tile = blockIdx.x; tile < T; tile += gridDim.x, ++ti) {
            const SegTable tb = load_seg_table(p.tile_seg, p.seg_ptr, tile, lane);
            // ---------- head: x^2, d logit -> dz2 (graphs s % 2 == e) ----------
            {
                const DropCtx dc = make_drop_ctx(p.drop2, seed2, p.H);
                HeadCtx hc;
                hc.tile_s = dz2_s, hc.dl_s = dl_s + e * 512, hc.ch = gtid, hc.gtid = gtid, hc.ch_ok = gtid < p.H;
                hc.w3s = w3 * p.drop2.scale;
                t0 = clock64();
                {      // d logit of this thread's slot (0 in padding slots), one copy per warpgroup
                    int row = -1;
                    for (int s = 0; s < tb.nseg; ++s) {
                        const int ns = __shfl_sync(0xffffffffu, tb.n, s), sl0 = __shfl_sync(0xffffffffu, tb.slot0, s),
                                  r0 = __shfl_sync(0xffffffffu, tb.row0, s);
                        if (gtid >= sl0 && gtid < sl0 + ns) row = r0 + gtid - sl0;
                    }
                    tc::sts_f32(hc.dl_s + 4 * gtid, row >= 0 ? __ldg(p.dlogit + row) : 0.f);
                }
                tc::group_mbar_wait(gtid == 0, x2_full, ti & 1, BAR_EPI + e, 128);      // (also orders the d logit staging)
                w_head += clock64() - t0;
                t0 = clock64();
                float dwt = 0.f;
                for (int s = e; s < tb.nseg; s += 2) {
                    const int ns = __shfl_sync(0xffffffffu, tb.n, s);
                    if (ns == 0) continue;
                    const int nblk = pad8(ns) >> 3;
                    const int slot0 = __shfl_sync(0xffffffffu, tb.slot0, s), row0 = __shfl_sync(0xffffffffu, tb.row0, s);
                    const float r = hc.ch_ok ? __ldg(p.rstd2 + (int64_t)(tb.g0 + s) * p.H + gtid) : 0.f;
                    if (nblk <= 8) {
                        EXT_DISPATCH_NB8(nblk, (head_graph<NB>(hc, dc, ns, row0, slot0, r, dwt)));
                    } else {
                        const int blk0 = slot0 >> 3;
                        float s1 = 0.f, s2 = 0.f;
                        uint32_t kw = 0xffffffffu;
                        for (int b = 0; b < nblk; ++b) {
                            if ((b & 3) == 0) kw = keep_bits32(dc, hc.ch, hc.ch_ok, row0 + 8 * b, ns - 8 * b);
                            head_blk_stats(hc, blk0 + b, ns - 8 * b, kw >> (8 * (b & 3)), s1, s2, dwt);
                        }
                        const float inv_n = 1.f / (float)ns;
                        const float c0 = -s1 * inv_n * r, c1 = -s2 * inv_n * r;
                        for (int b = 0; b < nblk; ++b) {
                            if ((b & 3) == 0) kw = keep_bits32(dc, hc.ch, hc.ch_ok, row0 + 8 * b, ns - 8 * b);
                            head_blk_emit(hc, blk0 + b, ns - 8 * b, kw >> (8 * (b & 3)), r, c0, c1);
                        }
                    }
                }
                dw3 = fmaf(dwt, p.drop2.scale, dw3);
                // slots outside every graph: zero columns (they feed accumulator columns nobody reads, but must be finite)
                if (e == 0) {
                    const int nb_tot = tb.total >> 3;
                    for (int b = nb_tot; b < 16; ++b) tc::sts128(dz2_s + mn_tile_offset_blk(gtid, b), 0u, 0u, 0u, 0u);
                }
                tc::fence_proxy_async_smem();
                tc::mbar_arrive(dz2_ready);
                t_head += clock64() - t0;
            }
            if (have_prev) dx_out(tb_prev, ti - 1);      // the previous tile's input gradient (see dx_out)
            // ---------- per channel block: InstanceNorm-1 backward (graphs s % 2 == e) ----------
            const DropCtx dc = make_drop_ctx(p.drop1, seed1, p.C1);
            int last_s = -1;                   // this warpgroup's last non-empty graph of the tile
            for (int s = e; s < tb.nseg; s += 2)
                if (__shfl_sync(0xffffffffu, tb.n, s) != 0) last_s = s;
            for (int cb = 0; cb < p.NCB; ++cb, ++n) {
                Epi3Ctx c3;
                c3.t1 = tmem_base + ((uint32_t)(q * 32) << 16), c3.t3 = c3.t1 + 128;
                c3.dz1_s = dz1_s, c3.h1_s = h1_s;
                c3.ch = cb * 128 + gtid, c3.gtid = gtid, c3.ch_ok = c3.ch < p.C1;
                c3.eps = p.eps, c3.dscale = p.drop1.scale;
                t0 = clock64();
                tc::group_mbar_wait(gtid == 0, acc_full, n & 1, BAR_EPI + 8 + e, 128);
                w_cb += clock64() - t0;
                t0 = clock64();
                tc::tc_fence_after();
                bool tiles_ready = false;
                auto wait_tiles = [&]() {      // dz1 / h1 tiles: the dx GEMM and the stores of the previous block have read them
                    if (!tiles_ready) {
                        tc::group_mbar_wait(gtid == 0, dz1_empty, (n & 1) ^ 1, BAR_EPI + e, 128);
                        tiles_ready = true;
                    }
                };
                // The accumulators go back to the MMA issuer as soon as this warpgroup's LAST graph of the block has been
                // read into registers (not after its statistics / emitting sweeps): GEMM1 + dh1 of the next channel block
                // then run underneath this block's epilogue arithmetic instead of after it.
                bool acc_released = false;
                for (int s = e; s < tb.nseg; s += 2) {
                    const int ns = __shfl_sync(0xffffffffu, tb.n, s);
                    if (ns == 0) continue;
                    const int nblk = pad8(ns) >> 3;
                    const int slot0 = __shfl_sync(0xffffffffu, tb.slot0, s), row0 = __shfl_sync(0xffffffffu, tb.row0, s);
                    if (nblk <= 7) {
                        const bool last = s == last_s;
                        auto loaded = [&]() {
                            if (last) {
                                tc::tc_fence_before();
                                tc::mbar_arrive(acc_empty);
                                acc_released = true;
                            }
                        };
                        EXT_DISPATCH_NB7(nblk, (epi3_graph<NB>(c3, dc, ns, row0, slot0, wait_tiles, loaded)));
                    } else {
                        float qs = 0.f;
                        for (int c4 = 0; c4 < nblk; c4 += 4) {
                            const int nbk = nblk - c4 < 4 ? nblk - c4 : 4;
                            EXT_DISPATCH_NB4(nbk, (qs += epi3_chunk_sumsq<NB>(c3, slot0 + 8 * c4)));
                        }
                        const float inv_n = 1.f / (float)ns;
                        const float r = 1.f / sqrtf(qs * inv_n + p.eps);
                        float s1 = 0.f, s2 = 0.f;
                        for (int c4 = 0; c4 < nblk; c4 += 4) {
                            const int nbk = nblk - c4 < 4 ? nblk - c4 : 4;
                            EXT_DISPATCH_NB4(nbk, (epi3_chunk_sums<NB>(c3, dc, r, ns - 8 * c4, row0 + 8 * c4, slot0 + 8 * c4, s1, s2)));
                        }
                        const float c0 = -s1 * inv_n * r, c1 = -s2 * inv_n * r;
                        wait_tiles();
                        for (int c4 = 0; c4 < nblk; c4 += 4) {
                            const int nbk = nblk - c4 < 4 ? nblk - c4 : 4;
                            EXT_DISPATCH_NB4(nbk, (epi3_chunk_emit<NB>(c3, dc, r, c0, c1, ns - 8 * c4, row0 + 8 * c4, slot0 + 8 * c4)));
                        }
                    }
                }
                wait_tiles();
                if (e == 0) {      // slots outside every graph: zero columns of both tiles (they are stored as a whole)
                    const int nb_tot = tb.total >> 3;
                    for (int b = nb_tot; b < 16; ++b) {
                        tc::sts128(dz1_s + mn_tile_offset_blk(gtid, b), 0u, 0u, 0u, 0u);
                        tc::sts128(h1_s + mn_tile_offset_blk(gtid, b), 0u, 0u, 0u, 0u);
                    }
                }
                tc::fence_proxy_async_smem();
                if (!acc_released) {      // (multi-pass path of a large last graph, or no graph at all for this warpgroup)
                    tc::tc_fence_before();
                    tc::mbar_arrive(acc_empty);
                }
                tc::mbar_arrive(dz1_full);
                t_cb += clock64() - t0;
            }
            tb_prev = tb;
            have_prev = true;
        }
        if (have_prev) dx_out(tb_prev, ti - 1);
        if (p.dbg && gtid == 0 && e == 0) {
            long long* d = p.dbg + (size_t)blockIdx.x * 16 + 6;
            d[0] = w_head, d[1] = t_head, d[2] = w_cb, d[3] = t_cb, d[4] = w_dxf, d[5] = t_dx;
        }
        // dw3 partial of this CTA: warpgroup 0 and 1 hold disjoint graphs -> two slabs, summed on the host side
        if (gtid < p.H) p.dw3_part[((size_t)blockIdx.x * 2 + e) * p.H + gtid] = dw3;
    }
    if (warp == 1 && lane == 0) tc::tma_store_wait_all<0>();      // outstanding TMA stores before exit
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) tc::tmem_dealloc(tmem_base, 512);
}

}  // namespace

extern "C" int gsatb_ext_fused_bwd(const int32_t* seg_ptr, const int32_t* tile_seg, const int32_t* num_tiles_dev, int max_tiles,
                                   int max_slots, int edge_mode, const void* w1_bf16_padded, const void* w2t_bf16_padded,
                                   const void* w1t_bf16_padded, const float* w3, const float* dlogit, const void* xhat2t,
                                   const float* rstd2, const void* xs, const uint8_t* mask1, const uint8_t* mask2,
                                   const uint32_t* seeds, float pdrop, int training, void* dz2t, void* dz1t, void* h1t,
                                   void* df12, int df12_is_bf16, float* dw3_part, int64_t ld_slots, int64_t rows, int H, int C1,
                                   float eps, gsatb_stream_t stream) {
    if (rows < 0 || H <= 0 || C1 <= 0 || max_tiles < 0) return GSATB_EINVAL;
    if (rows == 0 || max_tiles == 0) return GSATB_OK;
    if (!seg_ptr || !tile_seg || !num_tiles_dev || !w1_bf16_padded || !w2t_bf16_padded || !w1t_bf16_padded || !w3 || !dlogit ||
        !xhat2t || !rstd2 || !xs || !dz2t || !dz1t || !h1t || !df12 || !dw3_part)
        return GSATB_EINVAL;
    const int Kin = edge_mode ? 2 * H : H;
    if (H % 8 != 0 || H > 128 || Kin > 256 || C1 > 512) return GSATB_ESHAPE;
    if (max_slots != gsatb_ext_tile_slots(H, edge_mode)) return GSATB_EINVAL;
    if (ld_slots % TILE_SLOTS != 0) return GSATB_EINVAL;
    if (!gsatb_aligned16(xhat2t) || !gsatb_aligned16(dz2t) || !gsatb_aligned16(dz1t) || !gsatb_aligned16(h1t) || !gsatb_aligned16(xs))
        return GSATB_EALIGN;
    BwdParams p;
    p.seg_ptr = seg_ptr, p.tile_seg = tile_seg, p.num_tiles = num_tiles_dev;
    p.dlogit = dlogit, p.w3 = w3, p.rstd2 = rstd2;
    // masks in backward come from the forward's effective seeds (or the injected masks), never from the step counter
    p.drop1 = make_dropout(mask1, 0, pdrop, training, 1);
    p.drop2 = make_dropout(mask2, 0, pdrop, training, 1);
    p.drop1.step = nullptr, p.drop2.step = nullptr;
    p.seeds = seeds;
    p.dz2t = (uint16_t*)dz2t, p.dz1t = (uint16_t*)dz1t, p.h1t = (uint16_t*)h1t;
    p.df12 = (float*)df12, p.df12_bf16 = df12_is_bf16 != 0, p.dw3_part = dw3_part, p.ld_slots = ld_slots;
    p.H = H, p.Kin = Kin, p.C1 = C1, p.KB1 = (Kin + 63) / 64, p.KBH = (H + 63) / 64, p.NCB = (C1 + 127) / 128;
    p.KM = (Kin + 127) / 128;
    p.HP = pad128(H), p.C1P = pad128(C1);
    p.ldx = p.KB1 * 64;
    p.xkb = max_slots * 128;
    p.eps = eps;
    p.dbg = profile_buffer();
    if (training && pdrop > 0.f && !mask1 && !seeds) return GSATB_EINVAL;
    int nw = 8;
    while (nw > 2 && smem_plan(p.KB1, nw, p.xkb).total > 227 * 1024) --nw;
    p.NW = nw;
    const Smem L = smem_plan(p.KB1, p.NW, p.xkb);
    if (L.total > 227 * 1024) return GSATB_ESHAPE;
    CUtensorMap tm1, tm2t, tm1t, tmx;
    int rc = make_weight_tmap(&tm1, w1_bf16_padded, p.NCB * 128, p.KB1 * 64);
    if (rc != GSATB_OK) return rc;
    rc = make_weight_tmap(&tm2t, w2t_bf16_padded, p.NCB * 128, p.KBH * 64);            // W2^T [C1, H]
    if (rc != GSATB_OK) return rc;
    rc = make_weight_tmap(&tm1t, w1t_bf16_padded, p.KM * 128, ((C1 + 63) / 64) * 64);  // W1^T [Kin, C1]
    if (rc != GSATB_OK) return rc;
    CUtensorMap tdz2, tdz1, th1, txs;
    // slot-space tensors are TILE-major [tiles][pad128(C)][128 slots]: one contiguous block per tile, so the [128 channels x
    // 64 slots] boxes the kernels move are 128-byte pieces 256 bytes apart inside one 32 KiB run (not 128-byte pieces
    // ld_slots * 2 bytes apart).  As a 2-D map: rows = tile * pad128(C) + channel, 128 columns.
    const int64_t tiles = ld_slots / TILE_SLOTS;
    auto slot_map = [&](CUtensorMap* tm, const void* base, int C) -> int {
        PFN_tmapEncodeTiled fn = get_encode_fn();
        if (!fn) return GSATB_ELAUNCH;
        cuuint64_t gdim[2] = {(cuuint64_t)TILE_SLOTS, (cuuint64_t)(tiles * pad128(C))};
        cuuint64_t gstride[1] = {(cuuint64_t)TILE_SLOTS * 2};
        cuuint32_t box[2] = {64, 128}, estr[2] = {1, 1};
        return fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? GSATB_OK : GSATB_EINVAL;
    };
    if ((rc = slot_map(&tmx, xhat2t, H)) != GSATB_OK) return rc;
    if ((rc = slot_map(&tdz2, dz2t, H)) != GSATB_OK) return rc;
    if ((rc = slot_map(&tdz1, dz1t, C1)) != GSATB_OK) return rc;
    if ((rc = slot_map(&th1, h1t, C1)) != GSATB_OK) return rc;
    {      // xs [ld_slots, ldx] row-major: boxes of 64 input channels x max_slots rows
        PFN_tmapEncodeTiled fn = get_encode_fn();
        cuuint64_t gdim[2] = {(cuuint64_t)p.ldx, (cuuint64_t)ld_slots};
        cuuint64_t gstride[1] = {(cuuint64_t)p.ldx * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)max_slots}, estr[2] = {1, 1};
        if (fn(&txs, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(xs), gdim, gstride, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return GSATB_EINVAL;
    }
    if (cudaFuncSetAttribute(k_ext_fused_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
        return GSATB_ELAUNCH;
    const int grid = max_tiles < GSATB_NUM_SMS ? max_tiles : GSATB_NUM_SMS;
    k_ext_fused_bwd<<<grid, BWD_THREADS, L.total, (cudaStream_t)stream>>>(tm1, tm2t, tm1t, tmx, tdz2, tdz1, th1, txs, p);
    GSATB_CHECK_LAUNCH();
    return GSATB_OK;
}
