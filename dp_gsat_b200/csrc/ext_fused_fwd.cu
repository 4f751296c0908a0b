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
// Roles (20 warps): 0 weight TMA, 1 MMA issuer, 2 TMEM allocator, 4-11 epilogue 1 (two warpgroups, one per GEMM1
// accumulator), 12-15 epilogue 2, 16-19 gather producers.
#include "ext_fused.cuh"

namespace {

using namespace extf;

struct FwdParams {
    const float* emb;
    const int32_t* src;          // null: node mode (row r = node r, K = H)
    const int32_t* dst;
    const int32_t* seg_ptr;      // [G + 1] rows of every graph
    const int32_t* tile_seg;     // [T + 1] first graph of every tile
    const int32_t* num_tiles;    // [1] device
    const float* w3;
    const float* b3;             // nullable
    Dropout drop1, drop2;
    float* logit;                // [rows]
    uint16_t* xhat2t;            // bf16 [H, ld_slots] (slot space: tile t owns columns [128 t, 128 t + 128)), nullable
    int64_t ld_slots;
    uint32_t* seed_out;          // [2] effective dropout seeds of this launch (for the backward), nullable
    int H, Kin, C1, KB1, NCB, NXB, NW;
    float eps;
};

struct Smem {
    uint32_t ring, x, h1, red, scr, bars, total;
};
__host__ __device__ inline Smem smem_plan(int KB1, int NXB, int NW) {
    Smem s;
    s.ring = 0;
    s.x = s.ring + (uint32_t)NW * BRICK;
    s.h1 = s.x + (uint32_t)NXB * KB1 * BRICK;
    s.red = s.h1 + 2 * BRICK;
    s.scr = s.red + 2 * 4 * 128 * 4;          // epilogue-2 reduction scratch, double buffered
    s.bars = s.scr + 4096 + 1024;             // producer scratch: partial column sums + the segment mean
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

constexpr int BAR_EPI1 = 2, BAR_EPI2 = 4, BAR_PRO = 5;      // named barriers: BAR_EPI1 + e, BAR_EPI2, BAR_PRO (+8: waits)

__global__ void __launch_bounds__(EXT_THREADS, 1)
k_ext_fused_fwd(const __grid_constant__ CUtensorMap tm_w1, const __grid_constant__ CUtensorMap tm_w2, const FwdParams p) {
#ifdef GSATB_HOST_SIM
    uint8_t* smem_raw = simt::dyn_smem();
#else
    extern __shared__ uint8_t smem_raw[];
#endif
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const Smem L = smem_plan(p.KB1, p.NXB, p.NW);
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
        tc::reg_dec<CTL_REGS>();
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
                for (int tile = blockIdx.x; tile < T; tile += gridDim.x, ++ti) {
                    int N = pad16(tile_total_slots(p.tile_seg, p.seg_ptr, tile));
                    if (N < 16) N = 16;
                    const uint32_t idesc1 = tc::make_idesc_bf16(128, N, 0, 0), idesc2 = tc::make_idesc_bf16(128, N, 0, 1);
                    const uint32_t xb = ti % p.NXB, xuse = ti / p.NXB, b2 = ti & 1, use2 = ti >> 1;
                    const uint32_t x_addr = tc::smem_u32(xt + (size_t)xb * p.KB1 * BRICK);
                    const uint32_t h1_addr = tc::smem_u32(h1);
                    schedule(p.NCB, [&](bool g1, int cb) {
                        if (g1) {
                            const uint32_t buf = n1 & 1, use = n1 >> 1;
                            tc::mbar_wait(&acc1_empty[buf], (use & 1) ^ 1);
                            if (cb == 0) tc::mbar_wait(&x_full[xb], xuse & 1);
                            tc::tc_fence_after();
                            const uint32_t d = tmem_base + buf * 128;
                            for (int kb = 0; kb < p.KB1; ++kb, ++cw) {
                                const uint32_t s = cw % p.NW, usew = cw / p.NW;
                                tc::mbar_wait(&w_full[s], usew & 1);
                                tc::tc_fence_after();
                                const uint64_t a_desc = tc::make_desc_k_sw128(tc::smem_u32(ring + s * BRICK));
                                const uint64_t b_desc = tc::make_desc_k_sw128(x_addr + kb * BRICK);
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
                            tc::mbar_wait(h1_full, nh & 1);
                            if (cb == 0) tc::mbar_wait(&acc2_empty[b2], (use2 & 1) ^ 1);
                            tc::tc_fence_after();
                            const uint32_t d = tmem_base + 256 + b2 * 128;
                            const int nk = k2_blocks(p.C1, cb);
                            for (int kb = 0; kb < nk; ++kb, ++cw) {
                                const uint32_t s = cw % p.NW, usew = cw / p.NW;
                                tc::mbar_wait(&w_full[s], usew & 1);
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
            }
        }
    } else if (warp < 12) {
        // ===================== epilogue 1: InstanceNorm 1 -> ReLU -> Dropout -> h1 tile (B operand of GEMM2) ==========
        tc::reg_inc<EPI4_REGS>();
        const int e = (warp - 4) >> 2, q = warp & 3, gtid = q * 32 + lane;
        DropCtx dc;
        dc.d = p.drop1;
        dc.seed = dropout_seed(p.drop1);
        dc.on = p.drop1.enabled != 0;
        dc.use_mask = dc.on && p.drop1.mask != nullptr;
        dc.C = p.C1;
        const float dscale = p.drop1.scale;
        uint32_t n1 = 0;
        for (int tile = blockIdx.x; tile < T; tile += gridDim.x) {
            const SegTable tb = load_seg_table(p.tile_seg, p.seg_ptr, tile, lane);
            for (int cb = 0; cb < p.NCB; ++cb, ++n1) {
                if ((int)(n1 & 1) != e) continue;
                const uint32_t use = n1 >> 1;
                const int ch = cb * 128 + gtid;
                const bool ch_ok = ch < p.C1;
                tc::group_mbar_wait(gtid == 0, &acc1_full[e], use & 1, BAR_EPI1 + e, 128);
                tc::tc_fence_after();
                const uint32_t taddr = tmem_base + e * 128 + ((uint32_t)(q * 32) << 16);
                bool h1_ready = false;
                for (int s = 0; s < tb.nseg; ++s) {
                    const int n = __shfl_sync(0xffffffffu, tb.n, s);
                    if (n == 0) continue;
                    const int npad = pad8(n);
                    const int slot0 = __shfl_sync(0xffffffffu, tb.slot0, s), row0 = __shfl_sync(0xffffffffu, tb.row0, s);
                    float qa = 0.f, qb = 0.f;
                    for_pieces(npad, [&](auto Wt, int off) {
                        constexpr int W = decltype(Wt)::value;
                        float v[W];
                        tmem_ld_cols<W>(taddr + slot0 + off, v);
                        tc::tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < W; j += 2) {
                            qa = fmaf(v[j], v[j], qa);
                            qb = fmaf(v[j + 1], v[j + 1], qb);
                        }
                    });
                    const float rs = dscale / sqrtf((qa + qb) / (float)n + p.eps);
                    if (!h1_ready) {      // the single h1 tile: GEMM2 of the previous channel block has read it
                        tc::group_mbar_wait(gtid == 0, h1_empty, (n1 & 1) ^ 1, BAR_EPI1 + 8 + e, 128);
                        h1_ready = true;
                    }
                    uint32_t kw = 0xffffffffu;
                    for_pieces(npad, [&](auto Wt, int off) {
                        constexpr int W = decltype(Wt)::value;
                        if (dc.on && !dc.use_mask && (off & 31) == 0) kw = keep_word32(dc, (uint32_t)(row0 + off), ch, lane);
                        const uint32_t bits = !dc.on ? 0xffffffffu
                                              : dc.use_mask ? keep_bits_mask<W>(dc, row0 + off, n - off, ch, ch_ok)
                                                            : (kw >> (off & 31));
                        float v[W];
                        tmem_ld_cols<W>(taddr + slot0 + off, v);
                        tc::tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < W; j += 8) {
                            uint32_t o[4];
#pragma unroll
                            for (int i = 0; i < 8; i += 2) {
                                float a = fmaxf(v[j + i], 0.f) * rs, b = fmaxf(v[j + i + 1], 0.f) * rs;
                                a = ((bits >> (j + i)) & 1u) ? a : 0.f;
                                b = ((bits >> (j + i + 1)) & 1u) ? b : 0.f;
                                o[i >> 1] = tc::pack_bf16(a, b);
                            }
                            *reinterpret_cast<uint4*>(h1 + mn_tile_offset(gtid, slot0 + off + j)) = make_uint4(o[0], o[1], o[2], o[3]);
                        }
                    });
                }
                if (!h1_ready) tc::group_mbar_wait(gtid == 0, h1_empty, (n1 & 1) ^ 1, BAR_EPI1 + 8 + e, 128);
                tc::fence_proxy_async_smem();
                tc::tc_fence_before();
                tc::mbar_arrive(&acc1_empty[e]);
                tc::mbar_arrive(h1_full);
            }
        }
    } else if (warp < 16) {
        // ===================== epilogue 2: InstanceNorm 2 -> ReLU -> Dropout -> w3 dot -> logit =====================
        tc::reg_inc<EPI4_REGS>();
        const int q = warp & 3, gtid = q * 32 + lane;
        const int ch = gtid;
        const bool ch_ok = ch < p.H;
        const float w3 = ch_ok ? __ldg(p.w3 + ch) : 0.f;
        const float b3 = p.b3 ? __ldg(p.b3) : 0.f;
        DropCtx dc;
        dc.d = p.drop2;
        dc.seed = dropout_seed(p.drop2);
        dc.on = p.drop2.enabled != 0;
        dc.use_mask = dc.on && p.drop2.mask != nullptr;
        dc.C = p.H;
        const float dscale = p.drop2.scale;
        uint32_t ti = 0;
        for (int tile = blockIdx.x; tile < T; tile += gridDim.x, ++ti) {
            const uint32_t b2 = ti & 1, use2 = ti >> 1;
            const SegTable tb = load_seg_table(p.tile_seg, p.seg_ptr, tile, lane);
            float* red = reinterpret_cast<float*>(smem + L.red) + b2 * 512;
            tc::group_mbar_wait(gtid == 0, &acc2_full[b2], use2 & 1, BAR_EPI2, 128);
            tc::tc_fence_after();
            const uint32_t taddr = tmem_base + 256 + b2 * 128 + ((uint32_t)(q * 32) << 16);
            uint16_t* xrow = p.xhat2t && ch_ok ? p.xhat2t + (int64_t)ch * p.ld_slots + (int64_t)tile * TILE_SLOTS : nullptr;
            for (int s = 0; s < tb.nseg; ++s) {
                const int n = __shfl_sync(0xffffffffu, tb.n, s);
                if (n == 0) continue;
                const int npad = pad8(n);
                const int slot0 = __shfl_sync(0xffffffffu, tb.slot0, s), row0 = __shfl_sync(0xffffffffu, tb.row0, s);
                // sweep 1: sums of d = z - K and d^2 around a shift K close to the mean (the graph's first row)
                float s1 = 0.f, s2 = 0.f, Ksh = 0.f;
                for_pieces(npad, [&](auto Wt, int off) {
                    constexpr int W = decltype(Wt)::value;
                    float v[W];
                    tmem_ld_cols<W>(taddr + slot0 + off, v);
                    tc::tmem_ld_wait();
                    if (off == 0) Ksh = v[0];
                    const int nv = n - off;
#pragma unroll
                    for (int j = 0; j < W; ++j) {
                        const float d = j < nv ? v[j] - Ksh : 0.f;
                        s1 += d;
                        s2 = fmaf(d, d, s2);
                    }
                });
                const float inv_n = 1.f / (float)n, md = s1 * inv_n;
                const float var = fmaxf(s2 * inv_n - md * md, 0.f);
                const float r = 1.f / sqrtf(var + p.eps), mu = Ksh + md;
                uint32_t kw = 0xffffffffu;
                for_pieces(npad, [&](auto Wt, int off) {
                    constexpr int W = decltype(Wt)::value;
                    if (dc.on && !dc.use_mask && (off & 31) == 0) kw = keep_word32(dc, (uint32_t)(row0 + off), ch, lane);
                    const uint32_t bits = !dc.on ? 0xffffffffu
                                          : dc.use_mask ? keep_bits_mask<W>(dc, row0 + off, n - off, ch, ch_ok)
                                                        : (kw >> (off & 31));
                    float v[W], a[W];
                    tmem_ld_cols<W>(taddr + slot0 + off, v);
                    tc::tmem_ld_wait();
                    const int nv = n - off;
#pragma unroll
                    for (int j = 0; j < W; ++j) {
                        const float xh = j < nv ? (v[j] - mu) * r : 0.f;
                        v[j] = xh;
                        const float h = ((bits >> j) & 1u) ? fmaxf(xh, 0.f) : 0.f;
                        a[j] = h * (dscale * w3);
                    }
                    if (xrow) {
#pragma unroll
                        for (int j = 0; j < W; j += 8)
                            *reinterpret_cast<uint4*>(xrow + slot0 + off + j) =
                                make_uint4(tc::pack_bf16(v[j], v[j + 1]), tc::pack_bf16(v[j + 2], v[j + 3]),
                                           tc::pack_bf16(v[j + 4], v[j + 5]), tc::pack_bf16(v[j + 6], v[j + 7]));
                    }
                    const float tot = transpose_reduce<W>(a, lane);
                    if (lane < W) red[q * 128 + slot0 + off + lane] = tot;
                });
            }
            tc::tc_fence_before();
            tc::mbar_arrive(&acc2_empty[b2]);
            tc::named_bar_sync(BAR_EPI2 + 8, 128);
            // one logit per valid slot: the four lane quarters' partial dots + b3
            {
                const int slot = gtid;
                int row = -1;
                for (int s = 0; s < tb.nseg; ++s) {
                    const int n = __shfl_sync(0xffffffffu, tb.n, s), sl0 = __shfl_sync(0xffffffffu, tb.slot0, s),
                              r0 = __shfl_sync(0xffffffffu, tb.row0, s);
                    if (slot >= sl0 && slot < sl0 + n) row = r0 + slot - sl0;
                }
                if (row >= 0) p.logit[row] = red[slot] + red[128 + slot] + red[256 + slot] + red[384 + slot] + b3;
            }
        }
    } else {
        // ===================== gather producers: centred bf16 rows of f12 -> swizzled B tile of GEMM1 ================
        tc::reg_inc<PRO_REGS>();
        const int pt = threadIdx.x - 16 * 32;                  // 0..127
        const int nck = p.Kin >> 3;                            // 8-element chunks per row
        const int RP = 128 / nck > 0 ? 128 / nck : 1;          // rows processed in parallel
        const int ck = pt % nck, rl = pt / nck;
        const bool active = rl < RP && pt < RP * nck;
        const int k0 = ck * 8;
        const bool second = p.src != nullptr && k0 >= p.H;    // dst half of the concatenation
        const int kk = second ? k0 - p.H : k0;
        const int32_t* idx = p.src == nullptr ? nullptr : (second ? p.dst : p.src);
        float* part = reinterpret_cast<float*>(smem + L.scr);           // [RP][Kin] partial column sums
        float* mean = part + 1024;                                       // [Kin]
        const int kb = k0 >> 6, kin = k0 & 63;
        uint32_t ti = 0;
        for (int tile = blockIdx.x; tile < T; tile += gridDim.x, ++ti) {
            const uint32_t xb = ti % p.NXB, xuse = ti / p.NXB;
            const SegTable tb = load_seg_table(p.tile_seg, p.seg_ptr, tile, lane);
            tc::group_mbar_wait(pt == 0, &x_empty[xb], (xuse & 1) ^ 1, BAR_PRO + 8, 128);
            uint8_t* xbuf = xt + (size_t)xb * p.KB1 * BRICK + (size_t)kb * BRICK;
            for (int s = 0; s < tb.nseg; ++s) {
                const int n = __shfl_sync(0xffffffffu, tb.n, s);
                if (n == 0) continue;
                const int npad = pad8(n);
                const int slot0 = __shfl_sync(0xffffffffu, tb.slot0, s), row0 = __shfl_sync(0xffffffffu, tb.row0, s);
                // pass 1: column sums of the graph's gathered rows
                float acc[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = 0.f;
                if (active) {
#pragma unroll 4
                    for (int r = rl; r < n; r += RP) {
                        const int64_t node = idx ? __ldg(idx + row0 + r) : row0 + r;
                        float v[8];
                        load8_f32(p.emb + node * p.H, kk, p.H, v);
#pragma unroll
                        for (int i = 0; i < 8; ++i) acc[i] += v[i];
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) part[rl * p.Kin + k0 + i] = acc[i];
                }
                tc::named_bar_sync(BAR_PRO, 128);
                if (active && rl == 0) {
                    const float inv_n = 1.f / (float)n;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float t = 0.f;
                        for (int j = 0; j < RP; ++j) t += part[j * p.Kin + k0 + i];
                        mean[k0 + i] = t * inv_n;
                    }
                }
                tc::named_bar_sync(BAR_PRO, 128);
                // pass 2: centre, round to bf16, store swizzled; padding slots of the graph are zero rows
                if (active) {
                    float mu[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) mu[i] = mean[k0 + i];
#pragma unroll 4
                    for (int r = rl; r < npad; r += RP) {
                        uint32_t o[4] = {0u, 0u, 0u, 0u};
                        if (r < n) {
                            const int64_t node = idx ? __ldg(idx + row0 + r) : row0 + r;
                            float v[8];
                            load8_f32(p.emb + node * p.H, kk, p.H, v);
#pragma unroll
                            for (int i = 0; i < 8; ++i) v[i] -= mu[i];
                            pack8(v, o);
                        }
                        *reinterpret_cast<uint4*>(xbuf + tc::sw128_offset(slot0 + r, kin)) = make_uint4(o[0], o[1], o[2], o[3]);
                    }
                }
            }
            // slots between the last graph and the MMA width, and the K padding up to the 64-block, are zero
            {
                int N = pad16(tb.total);
                if (N < 16) N = 16;
                if (active)
                    for (int r = tb.total + rl; r < N; r += RP)
                        *reinterpret_cast<uint4*>(xbuf + tc::sw128_offset(r, kin)) = make_uint4(0u, 0u, 0u, 0u);
                const int kpad = p.KB1 * 64 - p.Kin;         // < 64, multiple of 8
                for (int i = pt; i < (kpad >> 3) * N; i += 128) {
                    const int r = i / (kpad >> 3), c = p.Kin + (i % (kpad >> 3)) * 8;
                    *reinterpret_cast<uint4*>(xt + (size_t)xb * p.KB1 * BRICK + (size_t)(c >> 6) * BRICK +
                                              tc::sw128_offset(r, c & 63)) = make_uint4(0u, 0u, 0u, 0u);
                }
            }
            tc::fence_proxy_async_smem();
            tc::mbar_arrive(&x_full[xb]);
        }
    }
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
__device__ __forceinline__ int plan_chunk(const int32_t* __restrict__ seg_ptr, int64_t g_lo, int64_t g_hi, int& oversize, F emit) {
    int cnt = 0, used = 0, nsg = 0;
    int prev = g_lo < g_hi ? __ldg(seg_ptr + g_lo) : 0;
    for (int64_t g = g_lo; g < g_hi; ++g) {
        const int nx = __ldg(seg_ptr + g + 1);
        const int len = nx - prev, np = pad8(len);
        prev = nx;
        if (len > TILE_SLOTS) ++oversize;
        if (nsg > 0 && (used + np > TILE_SLOTS || nsg == MAX_TSEG)) {
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
k_ext_tile_plan(const int32_t* __restrict__ seg_ptr, int64_t G, int32_t* __restrict__ tile_seg, int32_t* __restrict__ out) {
    __shared__ int scan[PLAN_THREADS];
    __shared__ int base, over;
    if (threadIdx.x == 0) base = 0, over = 0;
    __syncthreads();
    const int64_t nchunks = (G + PLAN_CHUNK - 1) / PLAN_CHUNK;
    for (int64_t c0 = 0; c0 < nchunks; c0 += PLAN_THREADS) {
        const int64_t c = c0 + threadIdx.x;
        const int64_t g_lo = c * PLAN_CHUNK < G ? c * PLAN_CHUNK : G, g_hi = (c + 1) * PLAN_CHUNK < G ? (c + 1) * PLAN_CHUNK : G;
        int oversize = 0;
        const int cnt = plan_chunk(seg_ptr, g_lo, g_hi, oversize, [](int, int32_t) {});
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
        plan_chunk(seg_ptr, g_lo, g_hi, dummy, [&](int k, int32_t g) { tile_seg[my0 + k] = g; });
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

extern "C" int gsatb_ext_tile_plan(const int32_t* seg_ptr, int64_t G, int32_t* tile_seg, int32_t* out2,
                                   gsatb_stream_t stream) {
    if (G < 0 || !tile_seg || !out2 || (G > 0 && !seg_ptr)) return GSATB_EINVAL;
    GSATB_LAUNCH(k_ext_tile_plan, 1, PLAN_THREADS, (cudaStream_t)stream, seg_ptr, G, tile_seg, out2);
    GSATB_CHECK_LAUNCH();
    return GSATB_OK;
}

extern "C" int gsatb_ext_fused_fwd(const float* emb, const int32_t* src, const int32_t* dst, const int32_t* seg_ptr,
                                   const int32_t* tile_seg, const int32_t* num_tiles_dev, int max_tiles,
                                   const void* w1_bf16_padded, const void* w2_bf16_padded, const float* w3,
                                   const float* b3, const uint8_t* mask1, const uint8_t* mask2, uint64_t seed, float pdrop,
                                   int training, float* logit, void* xhat2t, int64_t ld_slots, uint32_t* seed_out,
                                   int64_t rows, int H, int C1, float eps, gsatb_stream_t stream) {
    if (rows < 0 || H <= 0 || C1 <= 0 || max_tiles < 0) return GSATB_EINVAL;
    if (rows == 0 || max_tiles == 0) return GSATB_OK;
    if (!emb || !seg_ptr || !tile_seg || !num_tiles_dev || !w1_bf16_padded || !w2_bf16_padded || !w3 || !logit)
        return GSATB_EINVAL;
    if ((src == nullptr) != (dst == nullptr)) return GSATB_EINVAL;
    const int Kin = src ? 2 * H : H;
    if (H % 8 != 0 || H > 128 || Kin > 256 || C1 > 512) return GSATB_ESHAPE;
    if (!gsatb_aligned16(emb) || (xhat2t && (!gsatb_aligned16(xhat2t) || ld_slots % 8 != 0))) return GSATB_EALIGN;
    FwdParams p;
    p.emb = emb, p.src = src, p.dst = dst, p.seg_ptr = seg_ptr, p.tile_seg = tile_seg, p.num_tiles = num_tiles_dev;
    p.w3 = w3, p.b3 = b3;
    p.drop1 = make_dropout(mask1, seed * 2 + 1, pdrop, training, 1);
    p.drop2 = make_dropout(mask2, seed * 2 + 2, pdrop, training, 1);
    p.logit = logit, p.xhat2t = (uint16_t*)xhat2t, p.ld_slots = ld_slots, p.seed_out = seed_out;
    p.H = H, p.Kin = Kin, p.C1 = C1, p.KB1 = (Kin + 63) / 64, p.NCB = (C1 + 127) / 128;
    p.NXB = p.KB1 <= 2 ? 2 : 1;
    p.eps = eps;
    int nw = 8;
    while (nw > 2 && smem_plan(p.KB1, p.NXB, nw).total > 227 * 1024) --nw;
    p.NW = nw;
    const Smem L = smem_plan(p.KB1, p.NXB, p.NW);
    if (L.total > 227 * 1024) return GSATB_ESHAPE;
    CUtensorMap tm1, tm2;
    int rc = make_weight_tmap(&tm1, w1_bf16_padded, p.NCB * 128, p.KB1 * 64);
    if (rc != GSATB_OK) return rc;
    rc = make_weight_tmap(&tm2, w2_bf16_padded, 128, ((C1 + 63) / 64) * 64);
    if (rc != GSATB_OK) return rc;
    if (cudaFuncSetAttribute(k_ext_fused_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
        return GSATB_ELAUNCH;
    const int grid = max_tiles < GSATB_NUM_SMS ? max_tiles : GSATB_NUM_SMS;
    k_ext_fused_fwd<<<grid, EXT_THREADS, L.total, (cudaStream_t)stream>>>(tm1, tm2, p);
    GSATB_CHECK_LAUNCH();
    return GSATB_OK;
}
