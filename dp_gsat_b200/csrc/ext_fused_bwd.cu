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
    GatherArgs ga;
    const int32_t* seg_ptr;
    const int32_t* tile_seg;
    const int32_t* num_tiles;
    const float* dlogit;         // [rows]
    const float* w3;             // [H]
    const float* rstd2;          // [G, H]
    Dropout drop1, drop2;
    const uint32_t* seeds;       // [2] effective seeds written by the forward (null: injected masks / no dropout)
    uint16_t* dz2t;              // bf16 [H, ld_slots]
    uint16_t* dz1t;              // bf16 [C1, ld_slots]
    uint16_t* h1t;               // bf16 [C1, ld_slots]
    uint16_t* xs;                // bf16 [ld_slots, ldx]  centred input rows in slot space
    float* df12;                 // [rows, Kin]
    float* dw3_part;             // [grid, H]
    int64_t ld_slots;
    int ldx;
    int H, Kin, C1, KB1, KBH, NCB, KM, NW;
    int xkb;
    float eps;
    long long* dbg;
};

struct Smem {
    uint32_t ring, x, dz2, dz1, dl, scr, bars, total;
};
__host__ __device__ inline Smem smem_plan(int KB1, int NW, int xkb) {
    Smem s;
    s.ring = 0;
    s.x = s.ring + (uint32_t)NW * BRICK;
    s.dz2 = s.x + (uint32_t)KB1 * xkb;
    s.dz1 = s.dz2 + 2 * BRICK;
    s.dl = s.dz1 + 2 * 2 * BRICK;            // two dz1 tiles
    s.scr = s.dl + 512;                      // d logit per slot
    s.bars = s.scr + GATHER_SCRATCH;
    s.total = s.bars + 512 + 1024;
    return s;
}

__device__ __forceinline__ int kblocks_in_cb(int C1, int cb) {      // 64-channel K blocks inside channel block cb
    const int left = ((C1 + 63) / 64) * 64 - cb * 128;
    return left >= 128 ? 2 : (left > 0 ? 1 : 0);
}

constexpr int BAR_EPI = 2, BAR_PRO = 5;      // named barriers: BAR_EPI + e (+ 8), BAR_PRO (+ 8)

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
#define EXT_DISPATCH_NB4(nb, CALL)                          \
    switch (nb) {                                           \
        case 1: { constexpr int NB = 1; CALL; } break;      \
        case 2: { constexpr int NB = 2; CALL; } break;      \
        case 3: { constexpr int NB = 3; CALL; } break;      \
        default: { constexpr int NB = 4; CALL; } break;     \
    }

__device__ __forceinline__ void stg128(uint16_t* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    *reinterpret_cast<uint4*>(p) = make_uint4(a, b, c, d);
}

// ---- head: backward through Linear(H,1), Dropout 2, ReLU 2, InstanceNorm 2, for one graph of one channel -----------
struct HeadCtx {
    uint32_t tile_s;     // shared address of the x^2 / dz2 tile row of this channel: base + ch * 128 handled by offsets
    uint32_t dl_s;       // shared address of d logit per slot [128] floats
    uint16_t* grow;      // dz2t row of this channel at the tile's first slot (null for padded channels)
    int ch, gtid;
    bool ch_ok;
    float w3s;           // w3[ch] * dropout scale
};
// one 8-slot block: load x^2 (bf16) and d logit, accumulate the InstanceNorm backward sums and the dw3 term
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
        const float g = gate ? dl[j] * c.w3s : 0.f;          // d xhat2
        s1 += g;
        s2 = fmaf(g, x[j], s2);
        dw = fmaf(gate ? dl[j] : 0.f, x[j], dw);             // d w3 / scale: sum dl * relu(xhat2) * keep
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
        o[j] = j < nv ? fmaf(x[j], c1, fmaf(g, r, c0)) : 0.f;      // r (g - m1 - xhat m2); padding slots stay zero
    }
    const uint32_t p0 = tc::pack_bf16(o[0], o[1]), p1 = tc::pack_bf16(o[2], o[3]), p2 = tc::pack_bf16(o[4], o[5]),
                   p3 = tc::pack_bf16(o[6], o[7]);
    tc::sts128(sa, p0, p1, p2, p3);
    if (c.grow) stg128(c.grow + 8 * blk, p0, p1, p2, p3);
}

// ---- epilogue 3: one graph of one channel of a 128-channel block, both accumulators resident in registers ----------
struct Epi3Ctx {
    uint32_t t1, t3;     // TMEM addresses (this warp's lane quarter) of z~1 (GEMM1 recompute) and dh1
    uint32_t dz1_s;      // shared address of the dz1 tile being filled
    uint16_t* dz1row;    // dz1t / h1t rows of this channel at the tile's first slot (null for padded channels)
    uint16_t* h1row;
    int ch, gtid;
    bool ch_ok;
    float eps, dscale;
};
template <int NB>
__device__ __forceinline__ void epi3_graph(const Epi3Ctx& c, const DropCtx& dc, int n, int row0, int slot0) {
    float z[8 * NB], d[8 * NB];
    tmem_ld_blocks<NB>(c.t1 + slot0, z);
    tmem_ld_blocks<NB>(c.t3 + slot0, d);
    tc::tmem_ld_wait();
    const float inv_n = 1.f / (float)n;
    const float r = 1.f / sqrtf(sumsq_blocks<NB>(z) * inv_n + c.eps);
    const uint32_t k0 = keep_bits32(dc, c.ch, c.ch_ok, row0, n);
    uint32_t k1 = 0xffffffffu;
    if (NB > 4) k1 = keep_bits32(dc, c.ch, c.ch_ok, row0 + 32, n - 32);
    float s1 = 0.f, s2 = 0.f;
    uint32_t gates[(NB + 3) / 4];
#pragma unroll
    for (int w = 0; w < (NB + 3) / 4; ++w) gates[w] = 0u;
#pragma unroll
    for (int j = 0; j < 8 * NB; ++j) {
        const uint32_t kb = (j < 32 ? k0 >> j : k1 >> (j - 32)) & 1u;
        const bool gate = z[j] > 0.f && kb != 0u;            // padding slots: z~ = 0 -> gate off
        const float xh = z[j] * r;
        const float dy = gate ? d[j] * c.dscale : 0.f;      // d xhat1
        s1 += dy;
        s2 = fmaf(dy, xh, s2);
        z[j] = xh;
        d[j] = dy;
        gates[j >> 5] |= gate ? 1u << (j & 31) : 0u;
    }
    const float c0 = -s1 * inv_n * r, c1 = -s2 * inv_n * r;
    const int blk0 = slot0 >> 3;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const int nv = n - 8 * b;                            // only the last block can be partial
        uint32_t o[4], hq[4];
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
            const int j = 8 * b + i;
            float a0 = fmaf(z[j], c1, fmaf(d[j], r, c0)), a1 = fmaf(z[j + 1], c1, fmaf(d[j + 1], r, c0));
            if (b == NB - 1) {
                a0 = i < nv ? a0 : 0.f;
                a1 = i + 1 < nv ? a1 : 0.f;
            }
            o[i >> 1] = tc::pack_bf16(a0, a1);
            const float h0 = ((gates[j >> 5] >> (j & 31)) & 1u) ? z[j] * c.dscale : 0.f;
            const float h1v = ((gates[(j + 1) >> 5] >> ((j + 1) & 31)) & 1u) ? z[j + 1] * c.dscale : 0.f;
            hq[i >> 1] = tc::pack_bf16(h0, h1v);
        }
        tc::sts128(c.dz1_s + mn_tile_offset_blk(c.gtid, blk0 + b), o[0], o[1], o[2], o[3]);
        if (c.dz1row) {
            stg128(c.dz1row + 8 * (blk0 + b), o[0], o[1], o[2], o[3]);
            stg128(c.h1row + 8 * (blk0 + b), hq[0], hq[1], hq[2], hq[3]);
        }
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
    for (int b = 0; b < NB; ++b) {
        uint32_t o[4], hq[4];
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
            float a[2], hh[2];
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int j = 8 * b + i + t;
                const bool gate = z[j] > 0.f && ((k0 >> j) & 1u);
                const float xh = z[j] * r, dy = gate ? d[j] * c.dscale : 0.f;
                a[t] = j < nleft ? fmaf(xh, c1, fmaf(dy, r, c0)) : 0.f;
                hh[t] = gate ? xh * c.dscale : 0.f;
            }
            o[i >> 1] = tc::pack_bf16(a[0], a[1]);
            hq[i >> 1] = tc::pack_bf16(hh[0], hh[1]);
        }
        const int blk = (slot >> 3) + b;
        tc::sts128(c.dz1_s + mn_tile_offset_blk(c.gtid, blk), o[0], o[1], o[2], o[3]);
        if (c.dz1row) {
            stg128(c.dz1row + 8 * blk, o[0], o[1], o[2], o[3]);
            stg128(c.h1row + 8 * blk, hq[0], hq[1], hq[2], hq[3]);
        }
    }
}

// d f12 rows of one graph for one input channel: accumulator columns -> coalesced fp32 stores (lane = channel)
template <int NB>
__device__ __forceinline__ void dx_chunk_out(uint32_t taddr, float* out, int Kin, int nleft) {
    float v[8 * NB];
    tmem_ld_blocks<NB>(taddr, v);
    tc::tmem_ld_wait();
    if (out) {
#pragma unroll
        for (int j = 0; j < 8 * NB; ++j)
            if (j < nleft) out[(int64_t)j * Kin] = v[j];
    }
}

__global__ void __launch_bounds__(EXT_THREADS, 1)
k_ext_fused_bwd(const __grid_constant__ CUtensorMap tm_w1, const __grid_constant__ CUtensorMap tm_w2t,
                const __grid_constant__ CUtensorMap tm_w1t, const __grid_constant__ CUtensorMap tm_x2, const BwdParams p) {
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
    uint64_t* x_full = bars + 16;            // producers (128): x tile + d logit staged
    uint64_t* x_empty = bars + 17;           // commit: the tile's last GEMM1 has read the x tile
    uint64_t* x2_full = bars + 18;           // TMA tx: x^2 tile landed
    uint64_t* dz2_ready = bars + 19;         // 256 epilogue threads: dz2 tile complete
    uint64_t* dz2_empty = bars + 20;         // commit: the tile's last dh1 GEMM has read the dz2 tile
    uint64_t* acc_full = bars + 21;          // commit: z~1 and dh1 of a channel block complete
    uint64_t* acc_empty = bars + 22;         // 256: both accumulators read
    uint64_t* dz1_full = bars + 23;          // [2] 256: dz1 tile complete
    uint64_t* dz1_empty = bars + 25;         // [2] commit: dx GEMM of that block has read it
    uint64_t* dx_full = bars + 27;           // commit: d f12^T complete
    uint64_t* dx_empty = bars + 28;          // 256: d f12^T read
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 30);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T = __ldg(p.num_tiles);
    const uint32_t x_s = tc::smem_u32(smem + L.x), dz2_s = tc::smem_u32(smem + L.dz2), dz1_s0 = tc::smem_u32(smem + L.dz1),
                   dl_s = tc::smem_u32(smem + L.dl);

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&tm_w1);
        tc::tma_prefetch_desc(&tm_w2t);
        tc::tma_prefetch_desc(&tm_w1t);
        tc::tma_prefetch_desc(&tm_x2);
        for (int i = 0; i < 8; ++i) {
            tc::mbar_init(&w_full[i], 1);
            tc::mbar_init(&w_empty[i], 1);
        }
        tc::mbar_init(x_full, 128);
        tc::mbar_init(x_empty, 1);
        tc::mbar_init(x2_full, 1);
        tc::mbar_init(dz2_ready, 256);
        tc::mbar_init(dz2_empty, 1);
        tc::mbar_init(acc_full, 1);
        tc::mbar_init(acc_empty, 256);
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&dz1_full[i], 256);
            tc::mbar_init(&dz1_empty[i], 1);
        }
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
                for (int tile = blockIdx.x; tile < T; tile += gridDim.x, ++ti) {
                    int N = pad16(tile_total_slots(p.tile_seg, p.seg_ptr, tile));
                    if (N < 16) N = 16;
                    const uint32_t idesc_k = tc::make_idesc_bf16(128, N, 0, 0), idesc_mn = tc::make_idesc_bf16(128, N, 0, 1);
                    auto brick = [&]() -> uint64_t {
                        const uint32_t s = cw % p.NW, usew = cw / p.NW;
                        tc::mbar_wait(&w_full[s], usew & 1);
                        tc::tc_fence_after();
                        return tc::make_desc_k_sw128(tc::smem_u32(ring + s * BRICK));
                    };
                    auto release = [&]() {
                        tc::mma_commit(&w_empty[cw % p.NW]);
                        ++cw;
                    };
                    auto dx_step = [&](int cb, uint32_t nn) {      // d f12^T += W1^T[:, cb] dz1(cb)
                        const uint32_t buf = nn & 1, use = nn >> 1;
                        tc::mbar_wait(&dz1_full[buf], use & 1);
                        if (cb == 0) tc::mbar_wait(dx_empty, (ti & 1) ^ 1);
                        tc::tc_fence_after();
                        const uint32_t b_addr = dz1_s0 + buf * 2 * BRICK;
                        const int nk = kblocks_in_cb(p.C1, cb);
                        for (int mb = 0; mb < p.KM; ++mb)
                            for (int kb = 0; kb < nk; ++kb) {
                                const uint64_t a_desc = brick();
#pragma unroll
                                for (int k4 = 0; k4 < 4; ++k4)
                                    tc::mma_bf16_ss(tmem_base + 256 + mb * 128, a_desc + (uint64_t)(k4 * 2),
                                                    tc::make_desc_mn_sw128(b_addr + (kb * 64 + k4 * 16) * 128, BRICK), idesc_mn,
                                                    (cb | kb | k4) != 0);
                                release();
                            }
                        tc::mma_commit(&dz1_empty[buf]);
                    };
                    tc::mbar_wait(x_full, ti & 1);
                    tc::mbar_wait(dz2_ready, ti & 1);
                    tc::tc_fence_after();
                    for (int cb = 0; cb < p.NCB; ++cb, ++n) {
                        tc::mbar_wait(acc_empty, (n & 1) ^ 1);
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
                            tc::mma_commit(dz2_empty);
                        }
                        if (cb > 0) dx_step(cb - 1, n - 1);
                    }
                    dx_step(p.NCB - 1, n - 1);
                    tc::mma_commit(dx_full);
                }
            }
        } else if (warp == 3) {
            // ===================== x^2 tile (TMA): [H channels][slots] boxes of 64 slots x 128 channels =====================
            if (lane == 0) {
                uint32_t ti = 0;
                for (int tile = blockIdx.x; tile < T; tile += gridDim.x, ++ti) {
                    tc::mbar_wait(dz2_empty, (ti & 1) ^ 1);
                    tc::mbar_arrive_expect_tx(x2_full, 2 * BRICK);
                    tc::tma_load_2d(smem + L.dz2, &tm_x2, x2_full, tile * TILE_SLOTS, 0);
                    tc::tma_load_2d(smem + L.dz2 + BRICK, &tm_x2, x2_full, tile * TILE_SLOTS + 64, 0);
                }
            }
        }
    } else if (warp < 12) {
        // ===================== epilogue warpgroups =====================
        tc::reg_inc<EXT_EPI_REGS>();
        const int e = (warp - 4) >> 2, q = warp & 3, gtid = q * 32 + lane;
        const uint32_t seed1 = p.seeds ? __ldg(p.seeds) : 0u, seed2 = p.seeds ? __ldg(p.seeds + 1) : 0u;
        const float w3 = gtid < p.H ? __ldg(p.w3 + gtid) : 0.f;
        float dw3 = 0.f;
        uint32_t n = 0, ti = 0;
        for (int tile = blockIdx.x; tile < T; tile += gridDim.x, ++ti) {
            const SegTable tb = load_seg_table(p.tile_seg, p.seg_ptr, tile, lane);
            // ---------- head: x^2, d logit -> dz2 (graphs s % 2 == e) ----------
            {
                const DropCtx dc = make_drop_ctx(p.drop2, seed2, p.H);
                HeadCtx hc;
                hc.tile_s = dz2_s, hc.dl_s = dl_s, hc.ch = gtid, hc.gtid = gtid, hc.ch_ok = gtid < p.H;
                hc.grow = hc.ch_ok ? p.dz2t + (int64_t)gtid * p.ld_slots + (int64_t)tile * TILE_SLOTS : nullptr;
                hc.w3s = w3 * p.drop2.scale;
                tc::group_mbar_wait(gtid == 0, x2_full, ti & 1, BAR_EPI + e, 128);
                tc::group_mbar_wait(gtid == 0, x_full, ti & 1, BAR_EPI + 8 + e, 128);      // d logit staged with the x tile
                float dwt = 0.f;
                for (int s = e; s < tb.nseg; s += 2) {
                    const int ns = __shfl_sync(0xffffffffu, tb.n, s);
                    if (ns == 0) continue;
                    const int nblk = pad8(ns) >> 3;
                    const int slot0 = __shfl_sync(0xffffffffu, tb.slot0, s), row0 = __shfl_sync(0xffffffffu, tb.row0, s);
                    const int blk0 = slot0 >> 3;
                    float s1 = 0.f, s2 = 0.f;
                    uint32_t kw = 0xffffffffu;
                    for (int b = 0; b < nblk; ++b) {
                        if ((b & 3) == 0) kw = keep_bits32(dc, hc.ch, hc.ch_ok, row0 + 8 * b, ns - 8 * b);
                        head_blk_stats(hc, blk0 + b, ns - 8 * b, kw >> (8 * (b & 3)), s1, s2, dwt);
                    }
                    const float inv_n = 1.f / (float)ns;
                    const float r = hc.ch_ok ? __ldg(p.rstd2 + (int64_t)(tb.g0 + s) * p.H + gtid) : 0.f;
                    const float c0 = -s1 * inv_n * r, c1 = -s2 * inv_n * r;
                    for (int b = 0; b < nblk; ++b) {
                        if ((b & 3) == 0) kw = keep_bits32(dc, hc.ch, hc.ch_ok, row0 + 8 * b, ns - 8 * b);
                        head_blk_emit(hc, blk0 + b, ns - 8 * b, kw >> (8 * (b & 3)), r, c0, c1);
                    }
                }
                dw3 = fmaf(dwt, p.drop2.scale, dw3);
                // slots outside every graph: zero columns (they feed accumulator columns nobody reads, but must be finite)
                if (e == 0) {
                    const int nb_tot = tb.total >> 3;
                    for (int b = nb_tot; b < 16; ++b) {
                        tc::sts128(dz2_s + mn_tile_offset_blk(gtid, b), 0u, 0u, 0u, 0u);
                        if (hc.grow) stg128(hc.grow + 8 * b, 0u, 0u, 0u, 0u);      // dW operands: zero outside the graphs
                    }
                }
                tc::fence_proxy_async_smem();
                tc::mbar_arrive(dz2_ready);
            }
            // ---------- per channel block: InstanceNorm-1 backward (graphs s % 2 == e) ----------
            const DropCtx dc = make_drop_ctx(p.drop1, seed1, p.C1);
            for (int cb = 0; cb < p.NCB; ++cb, ++n) {
                const uint32_t buf = n & 1, use = n >> 1;
                Epi3Ctx c3;
                c3.t1 = tmem_base + ((uint32_t)(q * 32) << 16), c3.t3 = c3.t1 + 128;
                c3.dz1_s = dz1_s0 + buf * 2 * BRICK;
                c3.ch = cb * 128 + gtid, c3.gtid = gtid, c3.ch_ok = c3.ch < p.C1;
                c3.dz1row = c3.ch_ok ? p.dz1t + (int64_t)c3.ch * p.ld_slots + (int64_t)tile * TILE_SLOTS : nullptr;
                c3.h1row = c3.ch_ok ? p.h1t + (int64_t)c3.ch * p.ld_slots + (int64_t)tile * TILE_SLOTS : nullptr;
                c3.eps = p.eps, c3.dscale = p.drop1.scale;
                tc::group_mbar_wait(gtid == 0, &dz1_empty[buf], (use & 1) ^ 1, BAR_EPI + e, 128);
                tc::group_mbar_wait(gtid == 0, acc_full, n & 1, BAR_EPI + 8 + e, 128);
                tc::tc_fence_after();
                for (int s = e; s < tb.nseg; s += 2) {
                    const int ns = __shfl_sync(0xffffffffu, tb.n, s);
                    if (ns == 0) continue;
                    const int nblk = pad8(ns) >> 3;
                    const int slot0 = __shfl_sync(0xffffffffu, tb.slot0, s), row0 = __shfl_sync(0xffffffffu, tb.row0, s);
                    if (nblk <= 7) {
                        EXT_DISPATCH_NB7(nblk, (epi3_graph<NB>(c3, dc, ns, row0, slot0)));
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
                        for (int c4 = 0; c4 < nblk; c4 += 4) {
                            const int nbk = nblk - c4 < 4 ? nblk - c4 : 4;
                            EXT_DISPATCH_NB4(nbk, (epi3_chunk_emit<NB>(c3, dc, r, c0, c1, ns - 8 * c4, row0 + 8 * c4, slot0 + 8 * c4)));
                        }
                    }
                }
                if (e == 0) {      // slots outside every graph: zero columns of the dz1 tile and of the dW operands
                    const int nb_tot = tb.total >> 3;
                    for (int b = nb_tot; b < 16; ++b) {
                        tc::sts128(c3.dz1_s + mn_tile_offset_blk(gtid, b), 0u, 0u, 0u, 0u);
                        if (c3.dz1row) {
                            stg128(c3.dz1row + 8 * b, 0u, 0u, 0u, 0u);
                            stg128(c3.h1row + 8 * b, 0u, 0u, 0u, 0u);
                        }
                    }
                }
                tc::fence_proxy_async_smem();
                tc::tc_fence_before();
                tc::mbar_arrive(acc_empty);
                tc::mbar_arrive(&dz1_full[buf]);
            }
            // ---------- d f12^T -> d f12 rows (warpgroup e takes the 128-channel blocks mb % 2 == e) ----------
            tc::group_mbar_wait(gtid == 0, dx_full, ti & 1, BAR_EPI + e, 128);
            tc::tc_fence_after();
            for (int mb = e; mb < p.KM; mb += 2) {
                const int k = mb * 128 + gtid;
                const uint32_t taddr = tmem_base + 256 + mb * 128 + ((uint32_t)(q * 32) << 16);
                for (int s = 0; s < tb.nseg; ++s) {
                    const int ns = __shfl_sync(0xffffffffu, tb.n, s);
                    if (ns == 0) continue;
                    const int nblk = pad8(ns) >> 3;
                    const int slot0 = __shfl_sync(0xffffffffu, tb.slot0, s), row0 = __shfl_sync(0xffffffffu, tb.row0, s);
                    for (int c4 = 0; c4 < nblk; c4 += 4) {
                        const int nbk = nblk - c4 < 4 ? nblk - c4 : 4;
                        float* out = k < p.Kin ? p.df12 + (int64_t)(row0 + 8 * c4) * p.Kin + k : nullptr;
                        EXT_DISPATCH_NB4(nbk, (dx_chunk_out<NB>(taddr + slot0 + 8 * c4, out, p.Kin, ns - 8 * c4)));
                    }
                }
            }
            tc::tc_fence_before();
            tc::mbar_arrive(dx_empty);
        }
        // dw3 partial of this CTA: warpgroup 0 and 1 hold disjoint graphs -> two slabs, summed on the host side
        if (gtid < p.H) p.dw3_part[((size_t)blockIdx.x * 2 + e) * p.H + gtid] = dw3;
    } else {
        // ===================== gather producers: x tile (as in the forward) + d logit per slot + xs dump ================
        tc::reg_dec<EXT_PRO_REGS>();
        const int pt = threadIdx.x - 12 * 32;
        uint32_t ti = 0;
        for (int tile = blockIdx.x; tile < T; tile += gridDim.x, ++ti) {
            const SegTable tb = load_seg_table(p.tile_seg, p.seg_ptr, tile, lane);
            tc::group_mbar_wait(pt == 0, x_empty, (ti & 1) ^ 1, BAR_PRO + 8, 128);
            prefetch_next_tile(p.ga, p.tile_seg, p.seg_ptr, tile + (int)gridDim.x, T, pt);
            {      // d logit of this thread's slot
                int row = -1;
                for (int s = 0; s < tb.nseg; ++s) {
                    const int ns = __shfl_sync(0xffffffffu, tb.n, s), sl0 = __shfl_sync(0xffffffffu, tb.slot0, s),
                              r0 = __shfl_sync(0xffffffffu, tb.row0, s);
                    if (pt >= sl0 && pt < sl0 + ns) row = r0 + pt - sl0;
                }
                tc::sts_f32(dl_s + 4 * pt, row >= 0 ? __ldg(p.dlogit + row) : 0.f);
            }
            produce_x_tile(p.ga, tb, x_s, p.xkb, tc::smem_u32(smem + L.scr), pt, lane, BAR_PRO);
            tc::named_bar_sync(BAR_PRO, 128);
            // dump the tile (all 128 slots of this tile's slot range; slots past the MMA width are written as zeros)
            {
                int N = pad16(tb.total);
                if (N < 16) N = 16;
                const int cpr = p.ldx >> 3;                       // 16-byte chunks per row
                for (int i = pt; i < TILE_SLOTS * cpr; i += 128) {
                    const int r = i / cpr, c = (i % cpr) * 8;
                    uint4 v = make_uint4(0u, 0u, 0u, 0u);
                    if (r < N) v = tc::lds128(x_s + (uint32_t)(c >> 6) * p.xkb + tc::sw128_offset(r, c & 63));
                    *reinterpret_cast<uint4*>(p.xs + ((int64_t)tile * TILE_SLOTS + r) * p.ldx + c) = v;
                }
            }
            tc::fence_proxy_async_smem();
            tc::mbar_arrive(x_full);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) tc::tmem_dealloc(tmem_base, 512);
}

}  // namespace

extern "C" int gsatb_ext_fused_bwd(const float* emb, const int32_t* src, const int32_t* dst, const int32_t* node_ptr,
                                   const int32_t* rowptr_src, const int32_t* rowptr_dst, const int32_t* seg_ptr,
                                   const int32_t* tile_seg, const int32_t* num_tiles_dev, int max_tiles, int max_slots,
                                   const void* w1_bf16_padded, const void* w2t_bf16_padded, const void* w1t_bf16_padded,
                                   const float* w3, const float* dlogit, const void* xhat2t, const float* rstd2,
                                   const uint8_t* mask1, const uint8_t* mask2, const uint32_t* seeds, float pdrop,
                                   int training, void* dz2t, void* dz1t, void* h1t, void* xs, int ldx, float* df12,
                                   float* dw3_part, int64_t ld_slots, int64_t rows, int H, int C1, float eps,
                                   gsatb_stream_t stream) {
    if (rows < 0 || H <= 0 || C1 <= 0 || max_tiles < 0) return GSATB_EINVAL;
    if (rows == 0 || max_tiles == 0) return GSATB_OK;
    if (!emb || !seg_ptr || !tile_seg || !num_tiles_dev || !w1_bf16_padded || !w2t_bf16_padded || !w1t_bf16_padded || !w3 ||
        !dlogit || !xhat2t || !rstd2 || !dz2t || !dz1t || !h1t || !xs || !df12 || !dw3_part)
        return GSATB_EINVAL;
    if ((src == nullptr) != (dst == nullptr)) return GSATB_EINVAL;
    if (src && (!node_ptr || !rowptr_src || !rowptr_dst)) return GSATB_EINVAL;
    const int Kin = src ? 2 * H : H;
    if (H % 8 != 0 || H > 128 || Kin > 256 || C1 > 512) return GSATB_ESHAPE;
    if (max_slots != gsatb_ext_tile_slots(H, src != nullptr)) return GSATB_EINVAL;
    if (ldx != ((Kin + 63) / 64) * 64 || ld_slots % 8 != 0) return GSATB_EINVAL;
    if (!gsatb_aligned16(emb) || !gsatb_aligned16(xhat2t) || !gsatb_aligned16(dz2t) || !gsatb_aligned16(dz1t) ||
        !gsatb_aligned16(h1t) || !gsatb_aligned16(xs))
        return GSATB_EALIGN;
    BwdParams p;
    p.ga = GatherArgs{emb, src, dst, node_ptr, rowptr_src, rowptr_dst, H, Kin, (Kin + 63) / 64};
    p.seg_ptr = seg_ptr, p.tile_seg = tile_seg, p.num_tiles = num_tiles_dev;
    p.dlogit = dlogit, p.w3 = w3, p.rstd2 = rstd2;
    // masks in backward come from the forward's effective seeds (or the injected masks), never from the step counter
    p.drop1 = make_dropout(mask1, 0, pdrop, training, 1);
    p.drop2 = make_dropout(mask2, 0, pdrop, training, 1);
    p.drop1.step = nullptr, p.drop2.step = nullptr;
    p.seeds = seeds;
    p.dz2t = (uint16_t*)dz2t, p.dz1t = (uint16_t*)dz1t, p.h1t = (uint16_t*)h1t, p.xs = (uint16_t*)xs;
    p.df12 = df12, p.dw3_part = dw3_part, p.ld_slots = ld_slots, p.ldx = ldx;
    p.H = H, p.Kin = Kin, p.C1 = C1, p.KB1 = (Kin + 63) / 64, p.KBH = (H + 63) / 64, p.NCB = (C1 + 127) / 128;
    p.KM = (Kin + 127) / 128;
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
    {      // x^2 in slot space [H, ld_slots]: boxes of 64 slots x 128 channels (channels past H read as zero)
        PFN_tmapEncodeTiled fn = get_encode_fn();
        if (!fn) return GSATB_ELAUNCH;
        cuuint64_t gdim[2] = {(cuuint64_t)ld_slots, (cuuint64_t)H};
        cuuint64_t gstride[1] = {(cuuint64_t)ld_slots * 2};
        cuuint32_t box[2] = {64, 128}, estr[2] = {1, 1};
        if (fn(&tmx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(xhat2t), gdim, gstride, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return GSATB_EINVAL;
    }
    if (cudaFuncSetAttribute(k_ext_fused_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
        return GSATB_ELAUNCH;
    const int grid = max_tiles < GSATB_NUM_SMS ? max_tiles : GSATB_NUM_SMS;
    k_ext_fused_bwd<<<grid, EXT_THREADS, L.total, (cudaStream_t)stream>>>(tm1, tm2t, tm1t, tmx, p);
    GSATB_CHECK_LAUNCH();
    return GSATB_OK;
}
