// GIN node MLP (reference src/models/gin.py:55-62 with the ReLU / Dropout of gin.py:50-52, and its autograd) for the
// hidden sizes 64 and 128 in the ROW-OWNER orientation:
//
//      D[row (TMEM lane), channel (TMEM column)]  =  X[row, K] (A operand, a tile of 128 rows)  x  W[channel, K]^T (B operand)
//
// An epilogue thread owns one ROW of the tile and finds that row's channels in consecutive TMEM columns, i.e. in the
// order they lie in the row-major tensors of the step: the output leaves through a small per-warp SWIZZLE_128B staging
// buffer and one TMA store per 32 rows x 128 bytes, a dropout word (32 channels of one row) and a sign-bit word of the
// output are thread-local, and every input tile arrives by TMA through a ring of stages that is always in flight.  (The
// channel-owner orientation of tc_gemm.cuh is what per-channel statistics want; its transposing epilogues and
// register-staged operand producers left the node-MLP kernels at 49-60 % of the HBM peak.)
//
// Roles of the persistent CTA (one per SM, 640 threads):
//   warp 0        TMA loads: the weights once (they stay resident), then the input tile(s) of every tile of the CTA
//   warp 1        MMA issuer (one thread): tcgen05.mma M = 128 rows, N = H, K = 16
//   warp 2        TMEM allocator; with a transform stage: TMA store of the transformed operand tile (a1 / dz1)
//   Op::XF        warps 16-19 transform the input tile IN PLACE in shared memory (BatchNorm + ReLU; BatchNorm backward)
//                 before it becomes the A operand; warps 4-15 = 3 epilogue groups / accumulators
//   Op::SWAP      (first Linear: BatchNorm batch statistics) the MMA issuer ALSO issues the swapped product
//                 D^T[channel, row] = W x X^T into two more accumulators -- the tensor pipe idles anyway -- so that
//                 warps 12-19 own CHANNELS and sum z, z^2 over rows thread-locally (no shuffles, no atomics,
//                 deterministic); warps 4-11 = 2 row-owner epilogue groups writing z1
#include <cstdlib>
#include "tc_ops_common.cuh"

namespace {

using namespace tcg;

constexpr int RW_THREADS = 640;
constexpr int RW_STG = 4096;        // one staging buffer of an epilogue warp: 32 rows x 128 bytes, SWIZZLE_128B
constexpr int RW_BAR_XF = 2;        // named barrier of the transform warps
constexpr int RW_MAX_ST = 8;

struct RowsShape {
    int64_t rows;
    int num_tiles;
    int H, KB;        // width (K = OUT = H) and its 64-column blocks
    int NST;          // input stages
    int NSB;          // staging buffers per storing epilogue warp (1 or 2)
    long long* dbg;
};

struct RowsSmem {
    uint32_t w_off, in_off, stage_bytes, stg_off, bar_off, vec_off, total;
};
__host__ __device__ inline RowsSmem rows_smem(const RowsShape& s, int nin, int store_warps) {
    RowsSmem l;
    l.w_off = 0;
    l.in_off = (uint32_t)s.KB * BLK_BYTES;
    l.stage_bytes = (uint32_t)nin * s.KB * BLK_BYTES;
    l.stg_off = l.in_off + (uint32_t)s.NST * l.stage_bytes;
    l.bar_off = l.stg_off + (uint32_t)store_warps * s.NSB * RW_STG;
    l.vec_off = l.bar_off + 512;
    l.total = l.vec_off + 512 + 1024;      // bias vector + slack for the manual 1024-byte alignment
    return l;
}

// per-warp staging: wait until the TMA store that last used the chosen buffer has read it
__device__ __forceinline__ uint8_t* rows_stage_acquire(uint8_t* stg, int nsb, uint32_t sb, int lane) {
    if (lane == 0) {
        if (nsb == 2) tc::tma_store_wait_read<1>();
        else tc::tma_store_wait_read<0>();
    }
    __syncwarp();
    return stg + (nsb == 2 ? (sb & 1u) * RW_STG : 0u);
}
__device__ __forceinline__ void rows_stage_store(const CUtensorMap* tm, const uint8_t* buf, int lane, int col0, int row0,
                                                 bool any_row) {
    tc::fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0 && any_row) {
        tc::tma_store_2d(tm, buf, col0, row0);
        tc::tma_store_commit();
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Row-owner fp32 epilogue shared by the second Linear (bias, ReLU, dropout, sign bits) and the dX product of the first
// (plain): out[row, c] for the warp's 32 rows, 32 columns at a time.
// ------------------------------------------------------------------------------------------------------------------
struct EpiF32 {
    int relu_out;
    Dropout drop;
    uint32_t* posmask;      // nullable: [rows, ceil(H/32)]
};

__device__ __forceinline__ void rows_epilogue_f32(const EpiF32& e, const CUtensorMap* tm_out, const float* bias_s, int H,
                                                  uint32_t taddr, uint64_t* acc_empty, uint8_t* stg, int nsb, uint32_t& sb,
                                                  int lane, int q, int64_t r0, int cnt) {
    const int rl = q * 32 + lane;
    const int64_t row = r0 + rl;
    const bool ok = rl < cnt;
    const bool any_row = q * 32 < cnt;
    const int nchunk = H >> 5;
    const bool use_mask = e.drop.enabled && e.drop.mask != nullptr;
    const uint32_t dseed = dropout_seed(e.drop);
    const int W = (H + 31) >> 5;
#pragma unroll 1
    for (int c = 0; c < nchunk; ++c) {
        float v[32];
        tc::tmem_ld_32x32(taddr + c * 32, v);
        tc::tmem_ld_wait();
        if (c == nchunk - 1) {
            tc::tc_fence_before();
            tc::mbar_arrive(acc_empty);
        }
        uint32_t kw = 0xffffffffu;
        if (e.drop.enabled) {
            if (use_mask) {
                kw = 0u;
                if (ok) {
                    const uint4* mp = reinterpret_cast<const uint4*>(e.drop.mask + row * H + c * 32);
                    const uint4 m0 = __ldg(mp), m1 = __ldg(mp + 1);
                    const uint32_t w[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
#pragma unroll
                    for (int j = 0; j < 32; ++j) kw |= (((w[j >> 2] >> ((j & 3) * 8)) & 0xffu) != 0u ? 1u : 0u) << j;
                }
            } else {
                kw = dropout_word(e.drop, (uint32_t)row, (uint32_t)c, dseed);
            }
        }
        uint32_t posw = 0u;
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
            const float4 b = *reinterpret_cast<const float4*>(bias_s + c * 32 + j4 * 4);
            const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int j = j4 * 4 + i;
                float z = v[j] + bb[i];
                if (e.relu_out) z = fmaxf(z, 0.f);
                if (e.drop.enabled) z = ((kw >> j) & 1u) ? z * e.drop.scale : 0.f;
                posw |= (z > 0.f ? 1u : 0u) << j;
                v[j] = z;
            }
        }
        uint8_t* buf = rows_stage_acquire(stg, nsb, sb, lane);
        const uint32_t rowa = tc::smem_u32(buf) + (uint32_t)lane * 128u;
#pragma unroll
        for (int jj = 0; jj < 8; ++jj)
            tc::sts128(rowa + (uint32_t)((jj ^ (lane & 7)) << 4), __float_as_uint(v[4 * jj]), __float_as_uint(v[4 * jj + 1]),
                       __float_as_uint(v[4 * jj + 2]), __float_as_uint(v[4 * jj + 3]));
        rows_stage_store(tm_out, buf, lane, c * 32, (int)(r0 + q * 32), any_row);
        ++sb;
        if (e.posmask && ok) e.posmask[row * W + c] = posw;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Ops
// ------------------------------------------------------------------------------------------------------------------
// Second Linear of the node MLP with BatchNorm + ReLU folded into the operand path:
//   a1 = ReLU(z1 * scale + shift) (in place in the stage; stored as bf16 for dW2 when wanted)
//   h  = Dropout(ReLU(a1 W2^T + b2)) fp32, sign bits of h
struct OpRowsLin2 {
    static constexpr int NIN = 1;
    static constexpr bool XF = true, SWAP = false;
    struct Params {
        const float* scale;     // [H]
        const float* shift;     // [H]
        const float* bias;      // [H] nullable
        int xf_out;             // store the transformed tile (a1)
        EpiF32 epi;
    };
    struct XfState {
        float sc[8], sf[8];
    };
    __device__ static void xf_init(const Params& p, int cc, XfState& s) {
        load8_f32(p.scale, cc * 8, 0, s.sc);
        load8_f32(p.shift, cc * 8, 0, s.sf);
    }
    __device__ static void xf_unit(const XfState& s, uint32_t a0, uint32_t) {
        const uint4 qv = tc::lds128(a0);
        float v[8];
        unpack8(qv, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fmaxf(fmaf(v[i], s.sc[i], s.sf[i]), 0.f);
        uint32_t o[4];
        pack8(v, o);
        tc::sts128(a0, o[0], o[1], o[2], o[3]);
    }
};

// dX product of the first Linear with the BatchNorm backward folded into the operand path:
//   dz1 = cA * g + cB * z1 + cC (in place over g in the stage; stored as bf16 for dW1)
//   dx  = dz1 W1 fp32
struct OpRowsBwd1 {
    static constexpr int NIN = 2;
    static constexpr bool XF = true, SWAP = false;
    struct Params {
        const float* cA;
        const float* cB;
        const float* cC;
        const float* bias;      // unused (nullptr)
        int xf_out;
        EpiF32 epi;
    };
    struct XfState {
        float a[8], b[8], c[8];
    };
    __device__ static void xf_init(const Params& p, int cc, XfState& s) {
        load8_f32(p.cA, cc * 8, 0, s.a);
        load8_f32(p.cB, cc * 8, 0, s.b);
        load8_f32(p.cC, cc * 8, 0, s.c);
    }
    __device__ static void xf_unit(const XfState& s, uint32_t a0, uint32_t a1) {
        const uint4 qg = tc::lds128(a0), qz = tc::lds128(a1);
        float g[8], z[8];
        unpack8(qg, g);
        unpack8(qz, z);
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] = fmaf(s.a[i], g[i], fmaf(s.b[i], z[i], s.c[i]));
        uint32_t o[4];
        pack8(g, o);
        tc::sts128(a0, o[0], o[1], o[2], o[3]);
    }
};

// First Linear of the node MLP: z1 = x W1^T + b1 as bf16, BatchNorm batch statistics sum(z), sum(z^2) per channel from
// the fp32 accumulators of the swapped product.
struct OpRowsLin1 {
    static constexpr int NIN = 1;
    static constexpr bool XF = false, SWAP = true;
    struct Params {
        const float* bias;          // [H] nullable
        float* stat_partials;       // nullable: [gridDim * MAX_GROUPS][2][H]
        int xf_out;                 // unused
        EpiF32 epi;                 // unused
    };
    struct XfState {};
    __device__ static void xf_init(const Params&, int, XfState&) {}
    __device__ static void xf_unit(const XfState&, uint32_t, uint32_t) {}
};

template <class Op>
__global__ void __launch_bounds__(RW_THREADS, 1)
k_rows(const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_in0,
       const __grid_constant__ CUtensorMap tm_in1, const __grid_constant__ CUtensorMap tm_xf,
       const __grid_constant__ CUtensorMap tm_out, const RowsShape sh, const typename Op::Params p) {
    constexpr int NIN = Op::NIN;
    constexpr bool XF = Op::XF, SWAP = Op::SWAP;
    constexpr int NACC = XF ? 3 : (SWAP ? 2 : 4);          // row-owner accumulators = row-owner epilogue groups
    constexpr int STORE_WARPS = NACC * 4;
#ifdef GSATB_HOST_SIM
    uint8_t* smem_raw = simt::dyn_smem();
#else
    extern __shared__ uint8_t smem_raw[];
#endif
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const RowsSmem L = rows_smem(sh, NIN, STORE_WARPS);
    uint8_t* sW = smem + L.w_off;
    uint8_t* sIn = smem + L.in_off;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);
    uint64_t* w_full = bars;                  // [1]
    uint64_t* in_full = bars + 1;             // [NST]   TMA bytes of a stage have landed
    uint64_t* xf_done = bars + 9;             // [NST]   the transform warps have rewritten the stage
    uint64_t* in_free = bars + 17;            // [NST]   the MMAs (and the operand store) have read the stage
    uint64_t* acc_full = bars + 25;           // [4]
    uint64_t* acc_empty = bars + 29;          // [4]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 33);
    float* bias_s = reinterpret_cast<float*>(smem + L.vec_off);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int H = sh.H, KB = sh.KB;

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&tm_w);
        tc::tma_prefetch_desc(&tm_in0);
        tc::tma_prefetch_desc(&tm_out);
        tc::mbar_init(w_full, 1);
        for (int i = 0; i < sh.NST; ++i) {
            tc::mbar_init(&in_full[i], 1);
            tc::mbar_init(&xf_done[i], 128);
            tc::mbar_init(&in_free[i], XF ? 2 : 1);
        }
        for (int i = 0; i < 4; ++i) {
            tc::mbar_init(&acc_full[i], 1);
            tc::mbar_init(&acc_empty[i], SWAP ? 256 : 128);
        }
        tc::fence_barrier_init();
    }
    if (warp == 2) {
        tc::tmem_alloc(tmem_slot, 512);
        tc::tmem_relinquish();
    }
    if (threadIdx.x < 128) bias_s[threadIdx.x] = (p.bias && (int)threadIdx.x < H) ? __ldg(p.bias + threadIdx.x) : 0.f;
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        tc::reg_dec<CTL_REGS>();
        if (warp == 0) {
            // ===================== TMA loads =====================
            if (lane == 0) {
                tc::mbar_arrive_expect_tx(w_full, (uint32_t)KB * BLK_BYTES);
                for (int kb = 0; kb < KB; ++kb) tc::tma_load_2d(sW + kb * BLK_BYTES, &tm_w, w_full, kb * KBLK, 0);
                uint32_t ti = 0;
                for (int tile = blockIdx.x; tile < sh.num_tiles; tile += gridDim.x, ++ti) {
                    const uint32_t s = ti % sh.NST, u = ti / sh.NST;
                    tc::mbar_wait(&in_free[s], (u & 1) ^ 1);
                    tc::mbar_arrive_expect_tx(&in_full[s], L.stage_bytes);
                    uint8_t* st = sIn + (size_t)s * L.stage_bytes;
                    for (int kb = 0; kb < KB; ++kb)
                        tc::tma_load_2d(st + kb * BLK_BYTES, &tm_in0, &in_full[s], kb * KBLK, tile * TILE_ROWS);
                    if (NIN == 2)
                        for (int kb = 0; kb < KB; ++kb)
                            tc::tma_load_2d(st + (KB + kb) * BLK_BYTES, &tm_in1, &in_full[s], kb * KBLK, tile * TILE_ROWS);
                }
            }
        } else if (warp == 1) {
            // ===================== MMA issuer =====================
            if (lane == 0) {
                const uint32_t idesc = tc::make_idesc_bf16(128, H);          // rows x channels
                const uint32_t idesc_t = tc::make_idesc_bf16(128, 128);      // channels (zero-padded to 128) x rows
                long long w_in = 0, w_acc = 0, t_all = clock64(), t0;
                tc::mbar_wait(w_full, 0);
                uint32_t ti = 0;
                for (int tile = blockIdx.x; tile < sh.num_tiles; tile += gridDim.x, ++ti) {
                    const uint32_t s = ti % sh.NST, u = ti / sh.NST;
                    const uint32_t a = ti % NACC, ua = ti / NACC;
                    t0 = clock64();
                    tc::mbar_wait(XF ? &xf_done[s] : &in_full[s], u & 1);
                    w_in += clock64() - t0;
                    t0 = clock64();
                    tc::mbar_wait(&acc_empty[a], (ua & 1) ^ 1);
                    w_acc += clock64() - t0;
                    tc::tc_fence_after();
                    const uint32_t x_base = tc::smem_u32(sIn + (size_t)s * L.stage_bytes);
                    const uint32_t w_base = tc::smem_u32(sW);
                    for (int kb = 0; kb < KB; ++kb) {
                        const uint64_t xd = tc::make_desc_k_sw128(x_base + kb * BLK_BYTES);
                        const uint64_t wd = tc::make_desc_k_sw128(w_base + kb * BLK_BYTES);
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4)
                            tc::mma_bf16_ss(tmem_base + a * 128, xd + (uint64_t)(k4 * 2), wd + (uint64_t)(k4 * 2), idesc,
                                            (kb | k4) != 0);
                    }
                    if (SWAP) {
                        for (int kb = 0; kb < KB; ++kb) {
                            const uint64_t xd = tc::make_desc_k_sw128(x_base + kb * BLK_BYTES);
                            const uint64_t wd = tc::make_desc_k_sw128(w_base + kb * BLK_BYTES);
#pragma unroll
                            for (int k4 = 0; k4 < 4; ++k4)
                                tc::mma_bf16_ss(tmem_base + (2 + a) * 128, wd + (uint64_t)(k4 * 2), xd + (uint64_t)(k4 * 2),
                                                idesc_t, (kb | k4) != 0);
                        }
                    }
                    tc::mma_commit(&in_free[s]);
                    tc::mma_commit(&acc_full[a]);
                }
                if (sh.dbg) {
                    long long* d = sh.dbg + (size_t)blockIdx.x * 16;
                    d[0] = clock64() - t_all;
                    d[1] = w_in;
                    d[2] = w_acc;
                }
            }
        } else if (warp == 2) {
            // ===================== operand store (a1 / dz1 leave as the bf16 tensors the dW GEMMs read) =====================
            if (XF && lane == 0) {
                uint32_t ti = 0;
                for (int tile = blockIdx.x; tile < sh.num_tiles; tile += gridDim.x, ++ti) {
                    const uint32_t s = ti % sh.NST, u = ti / sh.NST;
                    tc::mbar_wait(&xf_done[s], u & 1);
                    if (p.xf_out) {
                        const uint8_t* st = sIn + (size_t)s * L.stage_bytes;
                        for (int kb = 0; kb < KB; ++kb) tc::tma_store_2d(&tm_xf, st + kb * BLK_BYTES, kb * KBLK, tile * TILE_ROWS);
                        tc::tma_store_commit();
                        tc::tma_store_wait_read<0>();
                    }
                    tc::mbar_arrive(&in_free[s]);
                }
                tc::tma_store_wait_all<0>();
            }
        }
    } else {
        tc::reg_inc<EPI4_REGS>();
        const int q = warp & 3;
        if (XF && warp >= 16) {
            // ===================== transform warps: the stage becomes the A operand in place =====================
            const int pt = threadIdx.x - 16 * 32;               // 0..127
            const int cprw = H >> 3, rpp = 128 / cprw;          // 16-byte chunks per row; rows covered per pass
            const int cc = pt % cprw, row0 = pt / cprw;
            typename Op::XfState xs;
            Op::xf_init(p, cc, xs);
            const uint32_t koff = (uint32_t)(cc >> 3) * BLK_BYTES;
            const int kin = (cc & 7) * 8;
            uint32_t ti = 0;
            long long w_in = 0, t_work = 0, t0;
            for (int tile = blockIdx.x; tile < sh.num_tiles; tile += gridDim.x, ++ti) {
                const uint32_t s = ti % sh.NST, u = ti / sh.NST;
                t0 = clock64();
                tc::group_mbar_wait(pt == 0, &in_full[s], u & 1, RW_BAR_XF, 128);
                w_in += clock64() - t0;
                t0 = clock64();
                const uint32_t st = tc::smem_u32(sIn + (size_t)s * L.stage_bytes) + koff;
#pragma unroll 4
                for (int i = 0; i < cprw; ++i) {
                    const uint32_t off = tc::sw128_offset(row0 + i * rpp, kin);
                    Op::xf_unit(xs, st + off, st + (uint32_t)KB * BLK_BYTES + off);
                }
                tc::fence_proxy_async_smem();
                tc::mbar_arrive(&xf_done[s]);
                t_work += clock64() - t0;
            }
            if (sh.dbg && pt == 0) {
                long long* d = sh.dbg + (size_t)blockIdx.x * 16;
                d[6] = w_in;
                d[7] = t_work;
            }
        } else if (SWAP && warp >= 12) {
            // ===================== channel-owner statistics from the swapped accumulators =====================
            const int grp = (warp - 12) >> 2;
            const int ch = q * 32 + lane;
            const bool ch_ok = ch < H;
            const bool warp_ok = q * 32 < H;
            const float b = bias_s[ch & 127];
            float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
            uint32_t ti = 0;
            for (int tile = blockIdx.x; tile < sh.num_tiles; tile += gridDim.x, ++ti) {
                const uint32_t a = ti % NACC, ua = ti / NACC;
                if ((int)a != grp) continue;
                const int64_t r0 = (int64_t)tile * TILE_ROWS;
                const int64_t left = sh.rows - r0;
                const int cnt = left < TILE_ROWS ? (int)left : TILE_ROWS;
                if (lane == 0) tc::mbar_wait(&acc_full[a], ua & 1);
                __syncwarp();
                tc::tc_fence_after();
                const uint32_t taddr = tmem_base + (2 + a) * 128 + ((uint32_t)(q * 32) << 16);
                if (warp_ok) {
#pragma unroll 1
                    for (int c = 0; c < 4; ++c) {
                        float v[32];
                        tc::tmem_ld_32x32(taddr + c * 32, v);
                        tc::tmem_ld_wait();
                        if (cnt == TILE_ROWS) {
#pragma unroll
                            for (int j = 0; j < 32; j += 2) {
                                const float y0 = v[j] + b, y1 = v[j + 1] + b;
                                s1a += y0;
                                s1b += y1;
                                s2a = fmaf(y0, y0, s2a);
                                s2b = fmaf(y1, y1, s2b);
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; j += 2) {
                                const float y0 = c * 32 + j < cnt ? v[j] + b : 0.f, y1 = c * 32 + j + 1 < cnt ? v[j + 1] + b : 0.f;
                                s1a += y0;
                                s1b += y1;
                                s2a = fmaf(y0, y0, s2a);
                                s2b = fmaf(y1, y1, s2b);
                            }
                        }
                    }
                }
                tc::tc_fence_before();
                tc::mbar_arrive(&acc_empty[a]);
            }
            if constexpr (SWAP) {
                if (p.stat_partials && ch_ok) {
                    const size_t part = (size_t)blockIdx.x * MAX_GROUPS + grp;
                    p.stat_partials[(part * 2 + 0) * H + ch] = s1a + s1b;
                    p.stat_partials[(part * 2 + 1) * H + ch] = s2a + s2b;
                }
            }
        } else {
            // ===================== row-owner epilogue groups =====================
            const int ew = warp - 4, grp = ew >> 2;
            uint8_t* stg = smem + L.stg_off + (size_t)ew * sh.NSB * RW_STG;
            uint32_t sb = 0, ti = 0;
            long long w_acc = 0, t_work = 0, t0;
            for (int tile = blockIdx.x; tile < sh.num_tiles; tile += gridDim.x, ++ti) {
                const uint32_t a = ti % NACC, ua = ti / NACC;
                if ((int)a != grp) continue;
                const int64_t r0 = (int64_t)tile * TILE_ROWS;
                const int64_t left = sh.rows - r0;
                const int cnt = left < TILE_ROWS ? (int)left : TILE_ROWS;
                t0 = clock64();
                if (lane == 0) tc::mbar_wait(&acc_full[a], ua & 1);
                __syncwarp();
                w_acc += clock64() - t0;
                tc::tc_fence_after();
                t0 = clock64();
                const uint32_t taddr = tmem_base + a * 128 + ((uint32_t)(q * 32) << 16);
                if constexpr (SWAP) {
                    // z1 = acc + b1 as bf16: two 32-column chunks fill one 128-byte staging row (64 channels)
                    const bool any_row = q * 32 < cnt;
                    const int nchunk = H >> 5;
                    uint8_t* buf = nullptr;
#pragma unroll 1
                    for (int c = 0; c < nchunk; ++c) {
                        float v[32];
                        tc::tmem_ld_32x32(taddr + c * 32, v);
                        tc::tmem_ld_wait();
                        if (c == nchunk - 1) {
                            tc::tc_fence_before();
                            tc::mbar_arrive(&acc_empty[a]);
                        }
                        if ((c & 1) == 0) buf = rows_stage_acquire(stg, sh.NSB, sb, lane);
                        const uint32_t rowa = tc::smem_u32(buf) + (uint32_t)lane * 128u;
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            const float4 b0 = *reinterpret_cast<const float4*>(bias_s + c * 32 + jj * 8);
                            const float4 b1 = *reinterpret_cast<const float4*>(bias_s + c * 32 + jj * 8 + 4);
                            const float* x = v + jj * 8;
                            tc::sts128(rowa + (uint32_t)((((c & 1) * 4 + jj) ^ (lane & 7)) << 4),
                                       tc::pack_bf16(x[0] + b0.x, x[1] + b0.y), tc::pack_bf16(x[2] + b0.z, x[3] + b0.w),
                                       tc::pack_bf16(x[4] + b1.x, x[5] + b1.y), tc::pack_bf16(x[6] + b1.z, x[7] + b1.w));
                        }
                        if ((c & 1) == 1 || c == nchunk - 1) {
                            rows_stage_store(&tm_out, buf, lane, (c >> 1) * 64, (int)(r0 + q * 32), any_row);
                            ++sb;
                        }
                    }
                } else {
                    rows_epilogue_f32(p.epi, &tm_out, bias_s, H, taddr, &acc_empty[a], stg, sh.NSB, sb, lane, q, r0, cnt);
                }
                t_work += clock64() - t0;
            }
            if (lane == 0) tc::tma_store_wait_all<0>();
            if (sh.dbg && ew == 0 && lane == 0) {
                long long* d = sh.dbg + (size_t)blockIdx.x * 16;
                d[4] = w_acc;
                d[5] = t_work;
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) tc::tmem_dealloc(tmem_base, 512);
}

// ---- host side -----------------------------------------------------------------------------------------------------
// [rows, cols] row-major tensor (ld elements per row) -> SWIZZLE_128B boxes of box_cols x box_rows elements
inline int make_rows_tmap(CUtensorMap* tm, CUtensorMapDataType dt, int es, const void* base, int64_t rows, int cols, int64_t ld,
                          int box_cols, int box_rows) {
    PFN_tmapEncodeTiled fn = get_encode_fn();
    if (!fn) return GSATB_ELAUNCH;
    if ((reinterpret_cast<uintptr_t>(base) & 15u) != 0 || ((ld * es) & 15) != 0) return GSATB_EALIGN;
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)ld * es};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? GSATB_OK : GSATB_EINVAL;
}

inline int rows_env(const char* name, int dflt, int lo, int hi) {
    const char* e = getenv(name);
    if (!e) return dflt;
    const int v = atoi(e);
    return v < lo ? lo : (v > hi ? hi : v);
}

// in0 / in1: bf16 [rows, H] inputs; xf_out: bf16 [rows, H] (nullable); out: fp32 or bf16 [rows, H]
template <class Op>
int launch_rows(const void* w_bf16_padded, const void* in0, const void* in1, void* xf_out, void* out, bool out_bf16,
                int64_t rows, int H, typename Op::Params p, cudaStream_t st) {
    if (rows <= 0) return GSATB_OK;
    if (H != 64 && H != 128) return GSATB_ESHAPE;
    if (rows > (int64_t)INT32_MAX - TILE_ROWS) return GSATB_ESHAPE;
    constexpr int STORE_WARPS = (Op::XF ? 3 : (Op::SWAP ? 2 : 4)) * 4;
    RowsShape sh;
    sh.rows = rows;
    sh.num_tiles = (int)((rows + TILE_ROWS - 1) / TILE_ROWS);
    sh.H = H;
    sh.KB = H / 64;
    sh.NSB = rows_env("GSATB_ROWS_NSB", 1, 1, 2);
    sh.dbg = profile_buffer();
    // input stages: as many as fit beside the resident weights and the staging buffers (at most RW_MAX_ST)
    const int fixed = sh.KB * BLK_BYTES + STORE_WARPS * sh.NSB * RW_STG + 512 + 512 + 1024;
    int nst = (227 * 1024 - fixed) / (Op::NIN * sh.KB * BLK_BYTES);
    if (nst > RW_MAX_ST) nst = RW_MAX_ST;
    nst = rows_env("GSATB_ROWS_NST", nst, 1, nst);
    if (nst < 2) return GSATB_ESHAPE;
    sh.NST = nst;
    p.xf_out = xf_out != nullptr;
    CUtensorMap tw, t0, t1, tx, to;
    int rc = make_weight_tmap(&tw, w_bf16_padded, 128, sh.KB * KBLK);
    if (rc != GSATB_OK) return rc;
    rc = make_act_tmap(&t0, in0, rows, H, H);
    if (rc != GSATB_OK) return rc;
    t1 = t0;
    if (Op::NIN == 2) {
        rc = make_act_tmap(&t1, in1, rows, H, H);
        if (rc != GSATB_OK) return rc;
    }
    tx = t0;
    if (xf_out) {
        rc = make_act_tmap(&tx, xf_out, rows, H, H);
        if (rc != GSATB_OK) return rc;
    }
    if (out_bf16) rc = make_rows_tmap(&to, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out, rows, H, H, 64, 32);
    else rc = make_rows_tmap(&to, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, out, rows, H, H, 32, 32);
    if (rc != GSATB_OK) return rc;
    const RowsSmem L = rows_smem(sh, Op::NIN, STORE_WARPS);
    if (L.total > 227 * 1024) return GSATB_ESHAPE;
    static bool attr_set = false;      // per instantiation; set once (not a stream operation)
    if (!attr_set) {
        if (cudaFuncSetAttribute(k_rows<Op>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
            return GSATB_ELAUNCH;
        attr_set = true;
    }
    const int grid = sh.num_tiles < GSATB_NUM_SMS ? sh.num_tiles : GSATB_NUM_SMS;
    k_rows<Op><<<grid, RW_THREADS, L.total, st>>>(tw, t0, t1, tx, to, sh, p);
    if (cudaPeekAtLastError() != cudaSuccess) return GSATB_ELAUNCH;
    return GSATB_OK;
}

__global__ void k_rows_reduce_partials(const float* __restrict__ partials, int parts, int width, double* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= width) return;
    double acc = 0.0;
    for (int q = 0; q < parts; ++q) acc += (double)partials[(size_t)q * width + j];
    out[j] = acc;
}

inline EpiF32 plain_epi() {
    EpiF32 e;
    e.relu_out = 0;
    e.drop = make_dropout(nullptr, 0, 0.f, 0, 1);
    e.posmask = nullptr;
    return e;
}

}  // namespace

extern "C" int gsatb_gin_rows_supported(int K, int H1, int H) { return (K == H1 && H1 == H && (H == 64 || H == 128)) ? 1 : 0; }

extern "C" int gsatb_gin_rows_lin1(const void* x_bf16, const void* w1_bf16_padded, const float* bias, void* z1_bf16,
                                   float* stat_partials, double* stats, int64_t rows, int H, gsatb_stream_t stream) {
    if (rows < 0 || H <= 0) return GSATB_EINVAL;
    if (rows == 0) return GSATB_OK;
    if (!x_bf16 || !w1_bf16_padded || !z1_bf16) return GSATB_EINVAL;
    if (stat_partials && !stats) return GSATB_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    if (stat_partials) cudaMemsetAsync(stat_partials, 0, (size_t)GSATB_NUM_SMS * MAX_GROUPS * 2 * H * sizeof(float), st);
    OpRowsLin1::Params p{bias, stat_partials, 0, plain_epi()};
    int rc = launch_rows<OpRowsLin1>(w1_bf16_padded, x_bf16, nullptr, nullptr, z1_bf16, true, rows, H, p, st);
    if (rc != GSATB_OK) return rc;
    if (stat_partials) {
        k_rows_reduce_partials<<<(2 * H + 127) / 128, 128, 0, st>>>(stat_partials, GSATB_NUM_SMS * MAX_GROUPS, 2 * H, stats);
        GSATB_CHECK_LAUNCH();
    }
    return GSATB_OK;
}

extern "C" int gsatb_gin_rows_lin2(const void* z1_bf16, const float* bn_scale, const float* bn_shift,
                                   const void* w2_bf16_padded, const float* bias, void* a1_bf16, float* h,
                                   uint32_t* posmask_out, const uint8_t* drop_mask, uint64_t drop_seed, float pdrop,
                                   int64_t rows, int H, gsatb_stream_t stream) {
    if (rows < 0 || H <= 0) return GSATB_EINVAL;
    if (rows == 0) return GSATB_OK;
    if (!z1_bf16 || !bn_scale || !bn_shift || !w2_bf16_padded || !h) return GSATB_EINVAL;
    if (!gsatb_aligned16(bn_scale) || !gsatb_aligned16(bn_shift) || (drop_mask && !gsatb_aligned16(drop_mask))) return GSATB_EALIGN;
    EpiF32 e;
    e.relu_out = 1;
    e.drop = make_dropout(drop_mask, drop_seed, pdrop, pdrop > 0.f, 1);
    e.posmask = posmask_out;
    OpRowsLin2::Params p{bn_scale, bn_shift, bias, 0, e};
    return launch_rows<OpRowsLin2>(w2_bf16_padded, z1_bf16, nullptr, a1_bf16, h, false, rows, H, p, (cudaStream_t)stream);
}

extern "C" int gsatb_gin_rows_bwd1(const void* g_bf16, const void* z1_bf16, const float* cA, const float* cB, const float* cC,
                                   const void* w1t_bf16_padded, void* dz1_bf16, float* dx, int64_t rows, int H,
                                   gsatb_stream_t stream) {
    if (rows < 0 || H <= 0) return GSATB_EINVAL;
    if (rows == 0) return GSATB_OK;
    if (!g_bf16 || !z1_bf16 || !cA || !cB || !cC || !w1t_bf16_padded || !dx) return GSATB_EINVAL;
    if (!gsatb_aligned16(cA) || !gsatb_aligned16(cB) || !gsatb_aligned16(cC)) return GSATB_EALIGN;
    OpRowsBwd1::Params p{cA, cB, cC, nullptr, 0, plain_epi()};
    return launch_rows<OpRowsBwd1>(w1t_bf16_padded, g_bf16, z1_bf16, dz1_bf16, dx, false, rows, H, p, (cudaStream_t)stream);
}
