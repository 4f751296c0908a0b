// Fused extractor MLP, forward, ONE persistent kernel (reference src/run_gsat.py:909-927 + src/utils/get_model.py:57-68):
//
//   f12 = cat(emb[col], emb[row])  ->  Linear(2H,4H) -> InstanceNorm(batch[col]) -> ReLU -> Dropout
//                                  ->  Linear(4H,H)  -> InstanceNorm            -> ReLU -> Dropout -> Linear(H,1)
//
// Nothing of width 4H (or 2H) ever reaches HBM: per graph-aligned tile (ext_fused.cuh) the kernel
//   * gathers emb[src] | emb[dst] (node mode: emb rows) straight into the swizzled B-operand tile of GEMM1, CENTRED per
//     graph in fp32 before the bf16 rounding.  Linear is linear: W (x - mean_g x) = z - mean_g z, so GEMM1 delivers the
//     centred pre-activation and InstanceNorm 1 only needs sum z~^2 (the biases b1, b2 cancel exactly, as in the norm);
//   * GEMM1 per 128-channel block into TMEM (tcgen05, swap-AB: lane = channel, column = slot), double buffered;
//   * epilogue 1 (thread = channel): rstd from one unmasked sweep, then ReLU (rstd > 0 commutes with it), dropout keep
//     bits, bf16 -> written as 16-byte vectors into an MN-major B tile [channel (K)][slot (N)] -- the layout in which an
//     epilogue thread's 8 consecutive slots are contiguous;
//   * GEMM2 accumulates over the channel blocks straight from that tile (MN-major B operand);
//   * epilogue 2: InstanceNorm 2 (shifted sums), ReLU, dropout, the w3 dot as a transposing warp reduction, one fp32
//     logit per row; optionally x^2 = InstanceNorm-2 output as bf16 in slot space [H, tiles * 128] for the backward.
// Weights stream from L2 through a ring of [128 x 64] TMA bricks in the order the MMA issuer consumes them.
//
// Roles (16 warps): 0 weight TMA, 1 MMA issuer, 2 TMEM allocator, 4-11 two epilogue warpgroups (one per GEMM1
// accumulator; they take the tiles' GEMM2 epilogue in turns), 12-15 gather producers.  setmaxnreg hands the epilogue
// threads 176 registers: a whole <= 64-slot graph of a channel stays in registers between the statistics sweep and the
// emitting sweep (one tcgen05.wait per graph instead of one per 8-column piece).
#include "ext_fused.cuh"

namespace {

using namespace extf;

struct FwdParams {
    GatherArgs ga;               // emb, src / dst (null: node mode), node_ptr, degree pointers, H, Kin, KB1
    const int32_t* seg_ptr;      // [G + 1] rows of every graph
    const int32_t* tile_seg;     // [T + 1] first graph of every tile
    const int32_t* num_tiles;    // [1] device
    const float* w3;
    const float* b3;             // nullable
    Dropout drop1, drop2;
    float* logit;                // [rows]
    uint16_t* xhat2t;            // bf16 slot space, TILE-major [tiles][pad128(H)][128 slots] (one contiguous block per tile), nullable
    int64_t ld_slots;
    float* rstd2;                // [G, H] InstanceNorm-2 reciprocal standard deviations (for the backward), nullable
    int dump_xs;                 // store every centred bf16 input tile to xs [tiles * 128, pad64(Kin)] (TMA), for the backward
    uint32_t* seed_out;          // [2] effective dropout seeds of this launch (for the backward), nullable
    int H, Kin, C1, KB1, NCB, NXB, NW;
    int xkb;                     // bytes of one K-block of the x tile: max slots per tile * 128
    float eps;
    long long* dbg;              // optional per-CTA cycle counters [grid][16] (gsatb_tc_set_profile_buffer)
};

struct Smem {
    uint32_t ring, x, h1, red, scr, bars, total;
};
__host__ __device__ inline Smem smem_plan(int KB1, int NXB, int NW, int xkb) {
    Smem s;
    s.ring = 0;
    s.x = s.ring + (uint32_t)NW * BRICK;
    s.h1 = s.x + (uint32_t)NXB * KB1 * xkb;
    s.red = s.h1 + 2 * BRICK;
    s.scr = s.red + 2 * 4 * 128 * 4;          // epilogue-2 reduction scratch, double buffered
    s.bars = s.scr + GATHER_SCRATCH;          // producer scratch: slot -> node table, partial column sums, segment mean
    s.total = s.bars + 256 + 1024;            // + slack for the manual 1024-byte alignment
    return s;
}

// GEMM steps of one tile in issue order: G1(0) G1(1) | G2(0) G1(2) | G2(1) G1(3) | ... ; f(is_gemm1, channel block)
template <class F>
__device__ __forceinline__ void schedule(int NCB, F f) {
    f(true, 0);
    if (NCB > 1) f(true, 1);
    for (int cb = 0; cb < NCB; ++cb) {
        f(false, cb);
        if (cb + 2 < NCB) f(true, cb + 2);
    }
}
__device__ __forceinline__ int k2_blocks(int C1, int cb) {      // 64-channel K blocks of GEMM2 inside channel block cb
    const int left = ((C1 + 63) / 64) * 64 - cb * 128;
    return left >= 128 ? 2 : (left > 0 ? 1 : 0);
}

// run CALL with a compile-time NB = nb (1..MAXNB)
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

struct Epi1Ctx {
    uint32_t taddr;      // TMEM address of this warp's lane quarter of the accumulator
    uint32_t h1_s;       // shared address of the h1 tile
    int ch, gtid;
    bool ch_ok;
    float eps, dscale;
};
// Epilogue 1 of one graph of NB <= 8 blocks, register resident: one TMEM wait, statistics and emission from registers.
template <int NB, class WaitH1>
__device__ __forceinline__ void epi1_graph(const Epi1Ctx& c, const DropCtx& dc, int n, int row0, int slot0, WaitH1 wait_h1) {
    float v[8 * NB];
    tmem_ld_blocks<NB>(c.taddr + slot0, v);
    tc::tmem_ld_wait();
    const float rs = c.dscale / sqrtf(sumsq_blocks<NB>(v) / (float)n + c.eps);
    wait_h1();
    const uint32_t k0 = keep_bits32(dc, c.ch, c.ch_ok, row0, n);
    uint32_t k1 = 0xffffffffu;
    if (NB > 4) k1 = keep_bits32(dc, c.ch, c.ch_ok, row0 + 32, n - 32);
    const int blk0 = slot0 >> 3;
#pragma unroll
    for (int b = 0; b < NB; ++b)
        emit_h1_blk(&v[8 * b], rs, b < 4 ? k0 >> (8 * b) : k1 >> (8 * (b - 4)), c.h1_s + mn_tile_offset_blk(c.gtid, blk0 + b));
}
// larger graphs, in chunks of <= 4 blocks: statistics pass, then emitting pass (the accumulator is read twice)
template <int NB>
__device__ __forceinline__ float epi1_chunk_sumsq(const Epi1Ctx& c, int slot) {
    float v[8 * NB];
    tmem_ld_blocks<NB>(c.taddr + slot, v);
    tc::tmem_ld_wait();
    return sumsq_blocks<NB>(v);
}
template <int NB>
__device__ __forceinline__ void epi1_chunk_emit(const Epi1Ctx& c, const DropCtx& dc, float rs, int nleft, int row, int slot) {
    float v[8 * NB];
    tmem_ld_blocks<NB>(c.taddr + slot, v);
    const uint32_t k0 = keep_bits32(dc, c.ch, c.ch_ok, row, nleft);
    tc::tmem_ld_wait();
#pragma unroll
    for (int b = 0; b < NB; ++b) emit_h1_blk(&v[8 * b], rs, k0 >> (8 * b), c.h1_s + mn_tile_offset_blk(c.gtid, (slot >> 3) + b));
}

struct Epi2Ctx {
    uint32_t taddr;
    uint32_t redq_s;     // shared address of this warp's row of the reduction scratch [4][128] floats
    uint16_t* xrow;      // xhat2t row of this channel at the tile's first slot, or null
    float* rs_out;       // rstd2 + channel (stride H per graph), or null
    int ch, lane;
    bool ch_ok;
    float eps, w3s;      // w3s = dropout scale * w3[ch]
};
// Epilogue 2, 8 slots of one channel: xhat = (z - mu) * rstd (stored as bf16 in slot space when xrow is given), then
// Dropout(ReLU(xhat)) * scale * w3 reduced over the warp's 32 channels; lanes 0-7 leave the 8 partial dots in redq.
template <bool LAST>
__device__ __forceinline__ void out2_blk(const Epi2Ctx& c, float* v, int nv, float r, float nmr, uint32_t bits, int slot) {
    float a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float xh = fmaf(v[j], r, nmr);
        if (LAST) xh = j < nv ? xh : 0.f;
        v[j] = xh;
        const float h = ((bits >> j) & 1u) ? fmaxf(xh, 0.f) : 0.f;
        a[j] = h * c.w3s;
    }
    if (c.xrow)
        *reinterpret_cast<uint4*>(c.xrow + slot) = make_uint4(tc::pack_bf16(v[0], v[1]), tc::pack_bf16(v[2], v[3]),
                                                              tc::pack_bf16(v[4], v[5]), tc::pack_bf16(v[6], v[7]));
    const float tot = transpose_reduce<8>(a, c.lane);
    if (c.lane < 8) tc::sts_f32(c.redq_s + 4 * (slot + c.lane), tot);
}
template <int NB>
__device__ __forceinline__ void epi2_graph(const Epi2Ctx& c, const DropCtx& dc, int n, int row0, int slot0, float* rs_out) {
    float v[8 * NB];
    tmem_ld_blocks<NB>(c.taddr + slot0, v);
    tc::tmem_ld_wait();
    const float Ksh = v[0], inv_n = 1.f / (float)n;
    float s1 = 0.f, s2 = 0.f;
    shifted_stats<NB>(v, n - 8 * (NB - 1), Ksh, s1, s2);
    const float md = s1 * inv_n;
    const float r = 1.f / sqrtf(fmaxf(s2 * inv_n - md * md, 0.f) + c.eps), nmr = -(Ksh + md) * r;
    if (rs_out) *rs_out = r;
    const uint32_t k0 = keep_bits32(dc, c.ch, c.ch_ok, row0, n);
    uint32_t k1 = 0xffffffffu;
    if (NB > 4) k1 = keep_bits32(dc, c.ch, c.ch_ok, row0 + 32, n - 32);
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const uint32_t bits = b < 4 ? k0 >> (8 * b) : k1 >> (8 * (b - 4));
        if (b == NB - 1) out2_blk<true>(c, &v[8 * b], n - 8 * b, r, nmr, bits, slot0 + 8 * b);
        else out2_blk<false>(c, &v[8 * b], 8, r, nmr, bits, slot0 + 8 * b);
    }
}
template <int NB>
__device__ __forceinline__ void epi2_chunk_stats(const Epi2Ctx& c, int slot, int nv_last, bool first, float& Ksh, float& s1, float& s2) {
    float v[8 * NB];
    tmem_ld_blocks<NB>(c.taddr + slot, v);
    tc::tmem_ld_wait();
    if (first) Ksh = v[0];
    shifted_stats<NB>(v, nv_last, Ksh, s1, s2);
}
template <int NB>
__device__ __forceinline__ void epi2_chunk_out(const Epi2Ctx& c, const DropCtx& dc, float r, float nmr, int nleft, int row, int slot) {
    float v[8 * NB];
    tmem_ld_blocks<NB>(c.taddr + slot, v);
    const uint32_t k0 = keep_bits32(dc, c.ch, c.ch_ok, row, nleft);
    tc::tmem_ld_wait();
#pragma unroll
    for (int b = 0; b < NB; ++b) out2_blk<true>(c, &v[8 * b], nleft - 8 * b, r, nmr, k0 >> (8 * b), slot + 8 * b);
}

constexpr int BAR_EPI1 = 2, BAR_PRO = 5;      // named barriers: BAR_EPI1 + e (+ 8), BAR_PRO (+ 8)

__global__ void __launch_bounds__(EXT_THREADS, 1)
k_ext_fused_fwd(const __grid_constant__ CUtensorMap tm_w1, const __grid_constant__ CUtensorMap tm_w2,
                const __grid_constant__ CUtensorMap tm_xs, const FwdParams p) {
#ifdef GSATB_HOST_SIM
    uint8_t* smem_raw = simt::dyn_smem();
#else
    extern __shared__ uint8_t smem_raw[];
#endif
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const Smem L = smem_plan(p.KB1, p.NXB, p.NW, p.xkb);
    uint8_t* ring = smem + L.ring;
    uint8_t* xt = smem + L.x;
    uint8_t* h1 = smem + L.h1;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
    uint64_t* w_full = bars;                 // [8]
    uint64_t* w_empty = bars + 8;            // [8]
    uint64_t* x_full = bars + 16;            // [2]
    uint64_t* x_empty = bars + 18;           // [2]
    uint64_t* acc1_full = bars + 20;         // [2]
    uint64_t* acc1_empty = bars + 22;        // [2]
    uint64_t* acc2_full = bars + 24;         // [2]
    uint64_t* acc2_empty = bars + 26;        // [2]
    uint64_t* h1_full = bars + 28;
    uint64_t* h1_empty = bars + 29;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 30);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int T = __ldg(p.num_tiles);

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&tm_w1);
        tc::tma_prefetch_desc(&tm_w2);
        for (int i = 0; i < 8; ++i) {
            tc::mbar_init(&w_full[i], 1);
            tc::mbar_init(&w_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&x_full[i], 128);
            tc::mbar_init(&x_empty[i], 1);
            tc::mbar_init(&acc1_full[i], 1);
            tc::mbar_init(&acc1_empty[i], 128);
            tc::mbar_init(&acc2_full[i], 1);
            tc::mbar_init(&acc2_empty[i], 128);
        }
        tc::mbar_init(h1_full, 128);
        tc::mbar_init(h1_empty, 1);
        tc::fence_barrier_init();
        if (blockIdx.x == 0 && p.seed_out) {
            p.seed_out[0] = dropout_seed(p.drop1);
            p.seed_out[1] = dropout_seed(p.drop2);
        }
    }
    if (warp == 2) {
        tc::tmem_alloc(tmem_slot, 512);
        tc::tmem_relinquish();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

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
                for (int tile = blockIdx.x; tile < T; tile += gridDim.x)
                    schedule(p.NCB, [&](bool g1, int cb) {
                        if (g1) {
                            for (int kb = 0; kb < p.KB1; ++kb) put(&tm_w1, kb * 64, cb * 128);
                        } else {
                            const int nk = k2_blocks(p.C1, cb);
                            for (int kb = 0; kb < nk; ++kb) put(&tm_w2, cb * 128 + kb * 64, 0);
                        }
                    });
            }
        } else if (warp == 1) {
            // ===================== MMA issuer =====================
            if (lane == 0) {
                uint32_t cw = 0, n1 = 0, nh = 0, ti = 0;
                long long wx = 0, wa1 = 0, ww = 0, wh = 0, wa2 = 0, t0, t_all = clock64();
                for (int tile = blockIdx.x; tile < T; tile += gridDim.x, ++ti) {
                    int N = pad16(tile_total_slots(p.tile_seg, p.seg_ptr, tile));
                    if (N < 16) N = 16;
                    const uint32_t idesc1 = tc::make_idesc_bf16(128, N, 0, 0), idesc2 = tc::make_idesc_bf16(128, N, 0, 1);
                    const uint32_t xb = ti % p.NXB, xuse = ti / p.NXB, b2 = ti & 1, use2 = ti >> 1;
                    const uint32_t x_addr = tc::smem_u32(xt) + xb * p.KB1 * p.xkb;
                    const uint32_t h1_addr = tc::smem_u32(h1);
                    schedule(p.NCB, [&](bool g1, int cb) {
                        if (g1) {
                            const uint32_t buf = n1 & 1, use = n1 >> 1;
                            t0 = clock64();
                            tc::mbar_wait(&acc1_empty[buf], (use & 1) ^ 1);
                            wa1 += clock64() - t0;
                            if (cb == 0) {
                                t0 = clock64();
                                tc::mbar_wait(&x_full[xb], xuse & 1);
                                wx += clock64() - t0;
                            }
                            tc::tc_fence_after();
                            const uint32_t d = tmem_base + buf * 128;
                            for (int kb = 0; kb < p.KB1; ++kb, ++cw) {
                                const uint32_t s = cw % p.NW, usew = cw / p.NW;
                                t0 = clock64();
                                tc::mbar_wait(&w_full[s], usew & 1);
                                ww += clock64() - t0;
                                tc::tc_fence_after();
                                const uint64_t a_desc = tc::make_desc_k_sw128(tc::smem_u32(ring + s * BRICK));
                                const uint64_t b_desc = tc::make_desc_k_sw128(x_addr + kb * p.xkb);
#pragma unroll
                                for (int k4 = 0; k4 < 4; ++k4)
                                    tc::mma_bf16_ss(d, a_desc + (uint64_t)(k4 * 2), b_desc + (uint64_t)(k4 * 2), idesc1,
                                                    (kb | k4) != 0);
                                tc::mma_commit(&w_empty[s]);
                            }
                            tc::mma_commit(&acc1_full[buf]);
                            if (cb == p.NCB - 1) tc::mma_commit(&x_empty[xb]);
                            ++n1;
                        } else {
                            t0 = clock64();
                            tc::mbar_wait(h1_full, nh & 1);
                            wh += clock64() - t0;
                            if (cb == 0) {
                                t0 = clock64();
                                tc::mbar_wait(&acc2_empty[b2], (use2 & 1) ^ 1);
                                wa2 += clock64() - t0;
                            }
                            tc::tc_fence_after();
                            const uint32_t d = tmem_base + 256 + b2 * 128;
                            const int nk = k2_blocks(p.C1, cb);
                            for (int kb = 0; kb < nk; ++kb, ++cw) {
                                const uint32_t s = cw % p.NW, usew = cw / p.NW;
                                t0 = clock64();
                                tc::mbar_wait(&w_full[s], usew & 1);
                                ww += clock64() - t0;
                                tc::tc_fence_after();
                                const uint64_t a_desc = tc::make_desc_k_sw128(tc::smem_u32(ring + s * BRICK));
#pragma unroll
                                for (int k4 = 0; k4 < 4; ++k4) {
                                    const uint64_t b_desc = tc::make_desc_mn_sw128(h1_addr + (kb * 64 + k4 * 16) * 128, BRICK);
                                    tc::mma_bf16_ss(d, a_desc + (uint64_t)(k4 * 2), b_desc, idesc2, (cb | kb | k4) != 0);
                                }
                                tc::mma_commit(&w_empty[s]);
                            }
                            tc::mma_commit(h1_empty);
                            if (cb == p.NCB - 1) tc::mma_commit(&acc2_full[b2]);
                            ++nh;
                        }
                    });
                }
                if (p.dbg) {
                    long long* d = p.dbg + (size_t)blockIdx.x * 16;
                    d[0] = clock64() - t_all, d[1] = wx, d[2] = wa1, d[3] = ww, d[4] = wh, d[5] = wa2;
                }
            }
        }
    } else if (warp < 12) {
        // ===================== epilogue warpgroups ====================================================================
        //   epilogue 1 (per channel block): InstanceNorm 1 -> ReLU -> Dropout -> h1 tile (B operand of GEMM2)
        //   epilogue 2 (per tile, warpgroup ti % 2): InstanceNorm 2 -> ReLU -> Dropout -> w3 dot -> logit
        tc::reg_inc<EXT_EPI_REGS>();
        const int e = (warp - 4) >> 2, q = warp & 3, gtid = q * 32 + lane;
        const float w3 = gtid < p.H ? __ldg(p.w3 + gtid) : 0.f;
        const float b3 = p.b3 ? __ldg(p.b3) : 0.f;
        uint32_t n1 = 0, ti = 0;
        long long w_acc = 0, w_h1 = 0, t_work = 0, w_acc2 = 0, t_work2 = 0, t0;
        for (int tile = blockIdx.x; tile < T; tile += gridDim.x, ++ti) {
            const SegTable tb = load_seg_table(p.tile_seg, p.seg_ptr, tile, lane);
            for (int cb = 0; cb < p.NCB; ++cb, ++n1) {
                if ((int)(n1 & 1) != e) continue;
                const DropCtx dc = make_drop_ctx(p.drop1, dropout_seed(p.drop1), p.C1);
                const float dscale = p.drop1.scale;
                const uint32_t use = n1 >> 1;
                const int ch = cb * 128 + gtid;
                const bool ch_ok = ch < p.C1;
                t0 = clock64();
                tc::group_mbar_wait(gtid == 0, &acc1_full[e], use & 1, BAR_EPI1 + e, 128);
                w_acc += clock64() - t0;
                tc::tc_fence_after();
                const long long t_begin = clock64();
                const uint32_t taddr = tmem_base + e * 128 + ((uint32_t)(q * 32) << 16);
                const uint32_t h1_s = tc::smem_u32(h1);
                bool h1_ready = false;
                auto wait_h1 = [&]() {      // the single h1 tile: GEMM2 of the previous channel block has read it
                    if (!h1_ready) {
                        t0 = clock64();
                        tc::group_mbar_wait(gtid == 0, h1_empty, (n1 & 1) ^ 1, BAR_EPI1 + 8 + e, 128);
                        w_h1 += clock64() - t0;
                        h1_ready = true;
                    }
                };
                Epi1Ctx c1;
                c1.taddr = taddr, c1.h1_s = h1_s, c1.ch = ch, c1.gtid = gtid, c1.ch_ok = ch_ok, c1.eps = p.eps, c1.dscale = dscale;
                for (int s = 0; s < tb.nseg; ++s) {
                    const int n = __shfl_sync(0xffffffffu, tb.n, s);
                    if (n == 0) continue;
                    const int nblk = pad8(n) >> 3;
                    const int slot0 = __shfl_sync(0xffffffffu, tb.slot0, s), row0 = __shfl_sync(0xffffffffu, tb.row0, s);
                    if (nblk <= 8) {
                        EXT_DISPATCH_NB8(nblk, (epi1_graph<NB>(c1, dc, n, row0, slot0, wait_h1)));
                    } else {
                        float q = 0.f;
                        for (int cb4 = 0; cb4 < nblk; cb4 += 4) {
                            const int nbk = nblk - cb4 < 4 ? nblk - cb4 : 4;
                            EXT_DISPATCH_NB4(nbk, (q += epi1_chunk_sumsq<NB>(c1, slot0 + 8 * cb4)));
                        }
                        const float rs = dscale / sqrtf(q / (float)n + p.eps);
                        wait_h1();
                        for (int cb4 = 0; cb4 < nblk; cb4 += 4) {
                            const int nbk = nblk - cb4 < 4 ? nblk - cb4 : 4;
                            EXT_DISPATCH_NB4(nbk, (epi1_chunk_emit<NB>(c1, dc, rs, n - 8 * cb4, row0 + 8 * cb4, slot0 + 8 * cb4)));
                        }
                    }
                }
                wait_h1();
                tc::fence_proxy_async_smem();
                tc::tc_fence_before();
                tc::mbar_arrive(&acc1_empty[e]);
                tc::mbar_arrive(h1_full);
                t_work += clock64() - t_begin;
            }
            if ((int)(ti & 1) != e) continue;
            // ---------- epilogue 2 of this tile ----------
            const uint32_t b2 = ti & 1, use2 = ti >> 1;
            const int ch = gtid;
            const bool ch_ok = ch < p.H;
            const DropCtx dc = make_drop_ctx(p.drop2, dropout_seed(p.drop2), p.H);
            const float dscale = p.drop2.scale;
            const uint32_t red_s = tc::smem_u32(smem + L.red) + b2 * 2048;
            t0 = clock64();
            tc::group_mbar_wait(gtid == 0, &acc2_full[b2], use2 & 1, BAR_EPI1 + e, 128);
            w_acc2 += clock64() - t0;
            t0 = clock64();
            tc::tc_fence_after();
            const uint32_t taddr = tmem_base + 256 + b2 * 128 + ((uint32_t)(q * 32) << 16);
            uint16_t* xrow = p.xhat2t && ch_ok ? p.xhat2t + ((int64_t)tile * pad128(p.H) + ch) * TILE_SLOTS : nullptr;
            Epi2Ctx c2;
            c2.taddr = taddr, c2.redq_s = red_s + 4 * (q * 128), c2.xrow = xrow, c2.ch = ch, c2.lane = lane, c2.ch_ok = ch_ok;
            c2.eps = p.eps, c2.w3s = dscale * w3;
            for (int s = 0; s < tb.nseg; ++s) {
                const int n = __shfl_sync(0xffffffffu, tb.n, s);
                if (n == 0) continue;
                const int nblk = pad8(n) >> 3;
                const int slot0 = __shfl_sync(0xffffffffu, tb.slot0, s), row0 = __shfl_sync(0xffffffffu, tb.row0, s);
                float* rs_out = p.rstd2 && ch_ok ? p.rstd2 + (int64_t)(tb.g0 + s) * p.H + ch : nullptr;
                if (nblk <= 8) {
                    EXT_DISPATCH_NB8(nblk, (epi2_graph<NB>(c2, dc, n, row0, slot0, rs_out)));
                } else {
                    float Ksh = 0.f, s1 = 0.f, s2 = 0.f;
                    for (int cb4 = 0; cb4 < nblk; cb4 += 4) {
                        const int nbk = nblk - cb4 < 4 ? nblk - cb4 : 4;
                        const int nvl = cb4 + 4 >= nblk ? n - 8 * (nblk - 1) : 8;
                        EXT_DISPATCH_NB4(nbk, (epi2_chunk_stats<NB>(c2, slot0 + 8 * cb4, nvl, cb4 == 0, Ksh, s1, s2)));
                    }
                    const float inv_n = 1.f / (float)n, md = s1 * inv_n;
                    const float r = 1.f / sqrtf(fmaxf(s2 * inv_n - md * md, 0.f) + p.eps), nmr = -(Ksh + md) * r;
                    if (rs_out) *rs_out = r;
                    for (int cb4 = 0; cb4 < nblk; cb4 += 4) {
                        const int nbk = nblk - cb4 < 4 ? nblk - cb4 : 4;
                        EXT_DISPATCH_NB4(nbk, (epi2_chunk_out<NB>(c2, dc, r, nmr, n - 8 * cb4, row0 + 8 * cb4, slot0 + 8 * cb4)));
                    }
                }
            }
            tc::tc_fence_before();
            tc::mbar_arrive(&acc2_empty[b2]);
            tc::named_bar_sync(BAR_EPI1 + 8 + e, 128);
            // one logit per valid slot: the four lane quarters' partial dots + b3
            {
                const int slot = gtid;
                int row = -1;
                for (int s = 0; s < tb.nseg; ++s) {
                    const int n = __shfl_sync(0xffffffffu, tb.n, s), sl0 = __shfl_sync(0xffffffffu, tb.slot0, s),
                              r0 = __shfl_sync(0xffffffffu, tb.row0, s);
                    if (slot >= sl0 && slot < sl0 + n) row = r0 + slot - sl0;
                }
                if (row >= 0)
                    p.logit[row] = tc::lds_f32(red_s + 4 * slot) + tc::lds_f32(red_s + 4 * (128 + slot)) +
                                   tc::lds_f32(red_s + 4 * (256 + slot)) + tc::lds_f32(red_s + 4 * (384 + slot)) + b3;
            }
            t_work2 += clock64() - t0;
        }
        if (p.dbg && gtid == 0) {
            long long* d = p.dbg + (size_t)blockIdx.x * 16 + 6 + e * 4;
            d[0] = w_acc, d[1] = t_work, d[2] = w_acc2, d[3] = t_work2;        // t_work includes the h1_empty waits
        }
    } else {
        // ===================== gather producers: centred bf16 rows of f12 -> swizzled B tile of GEMM1 ================
        tc::reg_dec<EXT_PRO_REGS>();
        const int pt = threadIdx.x - 12 * 32;                  // 0..127
        uint32_t ti = 0;
        long long w_x = 0, t_work = 0, t0;
        if (p.dump_xs)      // whole tiles are stored: rows past a tile's MMA width must hold finite values (0 * NaN = NaN in dW1)
            for (uint32_t o = (uint32_t)pt * 16; o < (uint32_t)(p.NXB * p.KB1) * p.xkb; o += 128 * 16)
                tc::sts128(tc::smem_u32(xt) + o, 0u, 0u, 0u, 0u);
        for (int tile = blockIdx.x; tile < T; tile += gridDim.x, ++ti) {
            const uint32_t xb = ti % p.NXB, xuse = ti / p.NXB;
            const SegTable tb = load_seg_table(p.tile_seg, p.seg_ptr, tile, lane);
            t0 = clock64();
            if (p.dump_xs && pt == 0) tc::tma_store_wait_read<0>();      // earlier xs stores have read their x buffer
            tc::group_mbar_wait(pt == 0, &x_empty[xb], (xuse & 1) ^ 1, BAR_PRO + 8, 128);
            w_x += clock64() - t0;
            t0 = clock64();
            prefetch_next_tile(p.ga, p.tile_seg, p.seg_ptr, tile + (int)gridDim.x, T, pt);
            produce_x_tile(p.ga, tb, tc::smem_u32(xt) + xb * p.KB1 * p.xkb, p.xkb, tc::smem_u32(smem + L.scr), pt, lane, BAR_PRO);
            tc::fence_proxy_async_smem();
            if (p.dump_xs) {
                tc::named_bar_sync(BAR_PRO, 128);
                if (pt == 0) {
                    for (int kb = 0; kb < p.KB1; ++kb)
                        tc::tma_store_2d(&tm_xs, xt + ((size_t)xb * p.KB1 + kb) * p.xkb, kb * 64, tile * TILE_SLOTS);
                    tc::tma_store_commit();
                }
            }
            tc::mbar_arrive(&x_full[xb]);
            t_work += clock64() - t0;
        }
        if (p.dbg && pt == 0) {
            long long* d = p.dbg + (size_t)blockIdx.x * 16 + 14;
            d[0] = w_x, d[1] = t_work;
        }
    }
    if (p.dump_xs && warp == 12 && lane == 0) tc::tma_store_wait_all<0>();
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) tc::tmem_dealloc(tmem_base, 512);
}

// ---- tile plan ---------------------------------------------------------------------------------------------------
// Greedy packing of consecutive graphs into tiles (<= 128 slots with every graph padded to 8, <= 16 graphs), on the
// device: each thread packs a chunk of PLAN_CHUNK consecutive graphs (a tile never spans two chunks), a block scan
// places the chunks' tiles.  out[0] = tiles, out[1] = graphs with more than 128 rows (the caller must not run the
// fused kernels on such a batch).
constexpr int PLAN_CHUNK = 256, PLAN_THREADS = 1024;

template <class F>
__device__ __forceinline__ int plan_chunk(const int32_t* __restrict__ seg_ptr, int64_t g_lo, int64_t g_hi, int cap, int& oversize, F emit) {
    int cnt = 0, used = 0, nsg = 0;
    int prev = g_lo < g_hi ? __ldg(seg_ptr + g_lo) : 0;
    for (int64_t g = g_lo; g < g_hi; ++g) {
        const int nx = __ldg(seg_ptr + g + 1);
        const int len = nx - prev, np = pad8(len);
        prev = nx;
        if (np > cap) ++oversize;
        if (nsg > 0 && (used + np > cap || nsg == MAX_TSEG)) {
            ++cnt;
            used = 0;
            nsg = 0;
        }
        if (nsg == 0) emit(cnt, (int32_t)g);
        used += np;
        ++nsg;
    }
    return cnt + (nsg > 0 ? 1 : 0);
}

__global__ void __launch_bounds__(PLAN_THREADS, 1)
k_ext_tile_plan(const int32_t* __restrict__ seg_ptr, int64_t G, int cap, int32_t* __restrict__ tile_seg, int32_t* __restrict__ out) {
    __shared__ int scan[PLAN_THREADS];
    __shared__ int base, over;
    if (threadIdx.x == 0) base = 0, over = 0;
    __syncthreads();
    const int64_t nchunks = (G + PLAN_CHUNK - 1) / PLAN_CHUNK;
    for (int64_t c0 = 0; c0 < nchunks; c0 += PLAN_THREADS) {
        const int64_t c = c0 + threadIdx.x;
        const int64_t g_lo = c * PLAN_CHUNK < G ? c * PLAN_CHUNK : G, g_hi = (c + 1) * PLAN_CHUNK < G ? (c + 1) * PLAN_CHUNK : G;
        int oversize = 0;
        const int cnt = plan_chunk(seg_ptr, g_lo, g_hi, cap, oversize, [](int, int32_t) {});
        if (oversize) atomicAdd(&over, oversize);
        scan[threadIdx.x] = cnt;
        __syncthreads();
        for (int off = 1; off < PLAN_THREADS; off <<= 1) {
            const int y = (int)threadIdx.x >= off ? scan[threadIdx.x - off] : 0;
            __syncthreads();
            scan[threadIdx.x] += y;
            __syncthreads();
        }
        const int my0 = base + scan[threadIdx.x] - cnt;
        int dummy = 0;
        plan_chunk(seg_ptr, g_lo, g_hi, cap, dummy, [&](int k, int32_t g) { tile_seg[my0 + k] = g; });
        __syncthreads();
        if (threadIdx.x == PLAN_THREADS - 1) base += scan[PLAN_THREADS - 1];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        tile_seg[base] = (int32_t)G;
        out[0] = base;
        out[1] = over;
    }
}

}  // namespace

// Slots per tile for hidden width H: 112 when the GEMM1 operand tile is wider than two 64-blocks (two x buffers of
// 4 x 14 KiB fit beside the weight ring; 112 = two 56-slot graphs of the BA-2Motifs shape), else 128.
extern "C" int gsatb_ext_tile_slots(int H, int edge_mode) {
    const int Kin = edge_mode ? 2 * H : H;
    return Kin > 128 ? 112 : 128;
}

extern "C" int gsatb_ext_tile_plan(const int32_t* seg_ptr, int64_t G, int max_slots, int32_t* tile_seg, int32_t* out2,
                                   gsatb_stream_t stream) {
    if (G < 0 || !tile_seg || !out2 || (G > 0 && !seg_ptr)) return GSATB_EINVAL;
    if (max_slots < 8 || max_slots > TILE_SLOTS || max_slots % 8 != 0) return GSATB_EINVAL;
    GSATB_LAUNCH(k_ext_tile_plan, 1, PLAN_THREADS, (cudaStream_t)stream, seg_ptr, G, max_slots, tile_seg, out2);
    GSATB_CHECK_LAUNCH();
    return GSATB_OK;
}

extern "C" int gsatb_ext_fused_fwd(const float* emb, const int32_t* src, const int32_t* dst, const int32_t* node_ptr,
                                   const int32_t* rowptr_src, const int32_t* rowptr_dst, const int32_t* seg_ptr,
                                   const int32_t* tile_seg, const int32_t* num_tiles_dev, int max_tiles, int max_slots,
                                   const void* w1_bf16_padded, const void* w2_bf16_padded, const float* w3,
                                   const float* b3, const uint8_t* mask1, const uint8_t* mask2, uint64_t seed, float pdrop,
                                   int training, float* logit, void* xhat2t, int64_t ld_slots, float* rstd2, void* xs, uint32_t* seed_out,
                                   int64_t rows, int H, int C1, float eps, gsatb_stream_t stream) {
    if (rows < 0 || H <= 0 || C1 <= 0 || max_tiles < 0) return GSATB_EINVAL;
    if (rows == 0 || max_tiles == 0) return GSATB_OK;
    if (!emb || !seg_ptr || !tile_seg || !num_tiles_dev || !w1_bf16_padded || !w2_bf16_padded || !w3 || !logit)
        return GSATB_EINVAL;
    if ((src == nullptr) != (dst == nullptr)) return GSATB_EINVAL;
    if (src && (!node_ptr || !rowptr_src || !rowptr_dst)) return GSATB_EINVAL;
    const int Kin = src ? 2 * H : H;
    if (H % 8 != 0 || H > 128 || Kin > 256 || C1 > 512) return GSATB_ESHAPE;
    if (!gsatb_aligned16(emb) || (xhat2t && (!gsatb_aligned16(xhat2t) || ld_slots % 8 != 0))) return GSATB_EALIGN;
    FwdParams p;
    p.ga = GatherArgs{emb, src, dst, node_ptr, rowptr_src, rowptr_dst, H, Kin, (Kin + 63) / 64};
    p.seg_ptr = seg_ptr, p.tile_seg = tile_seg, p.num_tiles = num_tiles_dev;
    p.w3 = w3, p.b3 = b3;
    p.drop1 = make_dropout(mask1, seed * 2 + 1, pdrop, training, 1);
    p.drop2 = make_dropout(mask2, seed * 2 + 2, pdrop, training, 1);
    p.logit = logit, p.xhat2t = (uint16_t*)xhat2t, p.ld_slots = ld_slots, p.rstd2 = rstd2, p.seed_out = seed_out;
    p.H = H, p.Kin = Kin, p.C1 = C1, p.KB1 = (Kin + 63) / 64, p.NCB = (C1 + 127) / 128;
    if (max_slots != gsatb_ext_tile_slots(H, src != nullptr)) return GSATB_EINVAL;      // the plan was built for another width
    p.NXB = 2;
    p.xkb = max_slots * 128;
    p.eps = eps;
    p.dbg = profile_buffer();
    int nw = 8;
    while (nw > 2 && smem_plan(p.KB1, p.NXB, nw, p.xkb).total > 227 * 1024) --nw;
    p.NW = nw;
    const Smem L = smem_plan(p.KB1, p.NXB, p.NW, p.xkb);
    if (L.total > 227 * 1024) return GSATB_ESHAPE;
    CUtensorMap tm1, tm2;
    int rc = make_weight_tmap(&tm1, w1_bf16_padded, p.NCB * 128, p.KB1 * 64);
    if (rc != GSATB_OK) return rc;
    rc = make_weight_tmap(&tm2, w2_bf16_padded, 128, ((C1 + 63) / 64) * 64);
    if (rc != GSATB_OK) return rc;
    CUtensorMap txs = tm1;
    p.dump_xs = xs != nullptr;
    if (xs) {      // xs [ld_slots, pad64(Kin)] row-major: boxes of 64 input channels x max_slots rows
        if (!gsatb_aligned16(xs) || ld_slots <= 0) return GSATB_EALIGN;
        PFN_tmapEncodeTiled fn = get_encode_fn();
        if (!fn) return GSATB_ELAUNCH;
        const int ldx = p.KB1 * 64;
        cuuint64_t gdim[2] = {(cuuint64_t)ldx, (cuuint64_t)ld_slots};
        cuuint64_t gstride[1] = {(cuuint64_t)ldx * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)max_slots}, estr[2] = {1, 1};
        if (fn(&txs, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, xs, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return GSATB_EINVAL;
    }
    if (cudaFuncSetAttribute(k_ext_fused_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
        return GSATB_ELAUNCH;
    const int grid = max_tiles < GSATB_NUM_SMS ? max_tiles : GSATB_NUM_SMS;
    k_ext_fused_fwd<<<grid, EXT_THREADS, L.total, (cudaStream_t)stream>>>(tm1, tm2, txs, p);
    GSATB_CHECK_LAUNCH();
    return GSATB_OK;
}
