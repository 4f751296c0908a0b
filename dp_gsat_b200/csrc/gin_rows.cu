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
    static constexpr int MAX_ST = 2;       // input stages (measured at 4.9 M x 128: 2 stages 0.89 ms, 3-4 stages 0.92)
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
    static constexpr int MAX_ST = RW_MAX_ST;
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
    static constexpr int MAX_ST = RW_MAX_ST;
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
                    const size_t part = (size_t)blockIdx.x * 2 + grp;          // [gridDim][2 channel-owner groups]
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

// ------------------------------------------------------------------------------------------------------------------
// gin_bwd2 in the row-owner orientation:  d2 = dh * (h > 0) * drop_scale (sign bits from the forward);  da1 = d2 W2;
// g = da1 * (ReLU(BN(z1)) > 0);  BatchNorm-backward statistics sum(g), sum(g * xhat) per channel.
//   warp 0       TMA loads: W2^T once, then the fp32 dh tile of every tile as [128 rows x 32 fp32] SWIZZLE_128B boxes
//   warps 16-19  transform, thread = row: per 64-channel block the row's 256 bytes of dh are read from two boxes, masked,
//                scaled, rounded to bf16 and written IN PLACE over the first of the two boxes, which thereby becomes the
//                [128 x 64] bf16 K-block of the A operand (a thread only ever overwrites bytes it has itself consumed)
//   warp 1       MMA issuer;  warp 2: TMEM allocator + TMA store of the d2 K-blocks (operand of dW2)
//   warps 4-11   two epilogue groups, per [32 rows x 64 channels] block of the warp: row phase (lane = row: accumulator ->
//                bf16 in the staging buffer), then channel phase (lane = channel pair: gate from z1, statistics, in place),
//                then one TMA store
// ------------------------------------------------------------------------------------------------------------------
struct RowsBwd2Params {
    const uint32_t* posmask;    // [rows, H / 32]
    float drop_scale;
    const uint16_t* z1;         // bf16 [rows, H]
    const float* sc;            // BatchNorm folded: a1 = relu(z1 * sc + sf)
    const float* sf;
    const float* mu;
    float* stat_partials;       // [gridDim][2][H]: sum g, sum g * (z1 - mu) over the CTA's tiles
};
constexpr int RW2_NACC = 2;             // accumulators = epilogue groups (warps 4-11; three groups with one staging buffer
                                        // each measured slower: 1.50 vs 1.36 ms at 4.9 M x 128)
constexpr int RW2_ROW_WARPS = 4 * RW2_NACC;
constexpr int RW2_STG = 8192;       // per epilogue warp: two [32 rows x 128 bytes] staging buffers
constexpr int RW2_BAR_RED = 3;      // named barrier of the epilogue warps' final partial-sum exchange

struct Rows2Smem {
    uint32_t in_off, stage_bytes, stg_off, bar_off, total;
};
__host__ __device__ inline Rows2Smem rows2_smem(const RowsShape& s) {
    Rows2Smem l;
    l.in_off = (uint32_t)s.KB * BLK_BYTES;
    l.stage_bytes = (uint32_t)(s.H / 32) * BLK_BYTES;
    l.stg_off = l.in_off + (uint32_t)s.NST * l.stage_bytes;
    l.bar_off = l.stg_off + RW2_ROW_WARPS * RW2_STG;
    l.total = l.bar_off + 512 + 1024;
    return l;
}

__global__ void __launch_bounds__(RW_THREADS, 1)
k_rows_bwd2(const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_dh,
            const __grid_constant__ CUtensorMap tm_d2, const __grid_constant__ CUtensorMap tm_g, const RowsShape sh,
            const RowsBwd2Params p) {
    constexpr int NACC = RW2_NACC;
#ifdef GSATB_HOST_SIM
    uint8_t* smem_raw = simt::dyn_smem();
#else
    extern __shared__ uint8_t smem_raw[];
#endif
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const Rows2Smem L = rows2_smem(sh);
    uint8_t* sW = smem;
    uint8_t* sIn = smem + L.in_off;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);
    uint64_t* w_full = bars;
    uint64_t* in_full = bars + 1;             // [NST <= 6]
    uint64_t* xf_done = bars + 7;
    uint64_t* in_free = bars + 13;
    uint64_t* acc_full = bars + 19;           // [3]
    uint64_t* acc_empty = bars + 22;          // [3]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 25);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int H = sh.H, KB = sh.KB, NB = sh.H >> 5;

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&tm_w);
        tc::tma_prefetch_desc(&tm_dh);
        tc::tma_prefetch_desc(&tm_d2);
        tc::tma_prefetch_desc(&tm_g);
        tc::mbar_init(w_full, 1);
        for (int i = 0; i < sh.NST; ++i) {
            tc::mbar_init(&in_full[i], 1);
            tc::mbar_init(&xf_done[i], 128);
            tc::mbar_init(&in_free[i], 2);
        }
        for (int i = 0; i < NACC; ++i) {
            tc::mbar_init(&acc_full[i], 1);
            tc::mbar_init(&acc_empty[i], 128);
        }
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

    if (warp < 4) {
        tc::reg_dec<CTL_REGS>();
        if (warp == 0) {
            if (lane == 0) {
                tc::mbar_arrive_expect_tx(w_full, (uint32_t)KB * BLK_BYTES);
                for (int kb = 0; kb < KB; ++kb) tc::tma_load_2d(sW + kb * BLK_BYTES, &tm_w, w_full, kb * KBLK, 0);
                uint32_t ti = 0;
                for (int tile = blockIdx.x; tile < sh.num_tiles; tile += gridDim.x, ++ti) {
                    const uint32_t s = ti % sh.NST, u = ti / sh.NST;
                    tc::mbar_wait(&in_free[s], (u & 1) ^ 1);
                    tc::mbar_arrive_expect_tx(&in_full[s], L.stage_bytes);
                    uint8_t* st = sIn + (size_t)s * L.stage_bytes;
                    for (int b = 0; b < NB; ++b) tc::tma_load_2d(st + b * BLK_BYTES, &tm_dh, &in_full[s], b * 32, tile * TILE_ROWS);
                }
            }
        } else if (warp == 1) {
            if (lane == 0) {
                const uint32_t idesc = tc::make_idesc_bf16(128, H);
                long long w_in = 0, w_acc = 0, t_all = clock64(), t0;
                tc::mbar_wait(w_full, 0);
                uint32_t ti = 0;
                for (int tile = blockIdx.x; tile < sh.num_tiles; tile += gridDim.x, ++ti) {
                    const uint32_t s = ti % sh.NST, u = ti / sh.NST;
                    const uint32_t a = ti % NACC, ua = ti / NACC;
                    t0 = clock64();
                    tc::mbar_wait(&xf_done[s], u & 1);
                    w_in += clock64() - t0;
                    t0 = clock64();
                    tc::mbar_wait(&acc_empty[a], (ua & 1) ^ 1);
                    w_acc += clock64() - t0;
                    tc::tc_fence_after();
                    const uint32_t x_base = tc::smem_u32(sIn + (size_t)s * L.stage_bytes);
                    const uint32_t w_base = tc::smem_u32(sW);
                    for (int kb = 0; kb < KB; ++kb) {
                        const uint64_t xd = tc::make_desc_k_sw128(x_base + 2 * kb * BLK_BYTES);     // d2 K-block kb lies over dh box 2 kb
                        const uint64_t wd = tc::make_desc_k_sw128(w_base + kb * BLK_BYTES);
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4)
                            tc::mma_bf16_ss(tmem_base + a * 128, xd + (uint64_t)(k4 * 2), wd + (uint64_t)(k4 * 2), idesc,
                                            (kb | k4) != 0);
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
            if (lane == 0) {
                uint32_t ti = 0;
                for (int tile = blockIdx.x; tile < sh.num_tiles; tile += gridDim.x, ++ti) {
                    const uint32_t s = ti % sh.NST, u = ti / sh.NST;
                    tc::mbar_wait(&xf_done[s], u & 1);
                    const uint8_t* st = sIn + (size_t)s * L.stage_bytes;
                    for (int kb = 0; kb < KB; ++kb) tc::tma_store_2d(&tm_d2, st + 2 * kb * BLK_BYTES, kb * KBLK, tile * TILE_ROWS);
                    tc::tma_store_commit();
                    tc::tma_store_wait_read<0>();
                    tc::mbar_arrive(&in_free[s]);
                }
                tc::tma_store_wait_all<0>();
            }
        }
    } else {
        tc::reg_inc<EPI4_REGS>();
        const int q = warp & 3;
        if (warp >= 16) {
            // ===================== transform: fp32 dh boxes -> bf16 d2 K-blocks, in place =====================
            const int row = threadIdx.x - 16 * 32;               // 0..127: thread = row, one 64-channel block after the other
            const int pt = row;
            const int W = H >> 5;
            const uint32_t r7 = (uint32_t)(row & 7);
            uint32_t ti = 0;
            long long w_in = 0, t_work = 0, t0;
            for (int tile = blockIdx.x; tile < sh.num_tiles; tile += gridDim.x, ++ti) {
                const uint32_t s = ti % sh.NST, u = ti / sh.NST;
                const int64_t grow = (int64_t)tile * TILE_ROWS + row;
                uint32_t mk[4] = {0u, 0u, 0u, 0u};
                if (grow < sh.rows) {
                    const uint32_t* pm = p.posmask + grow * W;
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (i < W) mk[i] = __ldg(pm + i);
                }
                t0 = clock64();
                tc::group_mbar_wait(pt == 0, &in_full[s], u & 1, RW_BAR_XF, 128);
                w_in += clock64() - t0;
                t0 = clock64();
#pragma unroll
                for (int kb = 0; kb < 2; ++kb)
                if (kb < KB) {
                    const uint32_t m0 = mk[2 * kb], m1 = mk[2 * kb + 1];
                    const uint32_t b0 = tc::smem_u32(sIn + (size_t)s * L.stage_bytes) + (uint32_t)(2 * kb) * BLK_BYTES + (uint32_t)row * 128u;
                    uint4 qv[16];
#pragma unroll
                    for (int jx = 0; jx < 8; ++jx) qv[jx] = tc::lds128(b0 + (((uint32_t)jx ^ r7) << 4));
#pragma unroll
                    for (int jx = 0; jx < 8; ++jx) qv[8 + jx] = tc::lds128(b0 + BLK_BYTES + (((uint32_t)jx ^ r7) << 4));
                    // (every load of the row precedes every store: the asm statements are volatile and keep their order)
#pragma unroll
                    for (int jx = 0; jx < 8; ++jx) {          // output chunk jx = channels 8 jx .. 8 jx + 7 of the block
                        const uint32_t mw = (jx < 4 ? m0 : m1) >> ((jx & 3) * 8);
                        const uint4 x0 = qv[2 * jx], x1 = qv[2 * jx + 1];
                        const float f[8] = {__uint_as_float(x0.x), __uint_as_float(x0.y), __uint_as_float(x0.z), __uint_as_float(x0.w),
                                            __uint_as_float(x1.x), __uint_as_float(x1.y), __uint_as_float(x1.z), __uint_as_float(x1.w)};
                        float v[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] = ((mw >> i) & 1u) ? f[i] * p.drop_scale : 0.f;
                        uint32_t o[4];
                        pack8(v, o);
                        tc::sts128(b0 + (((uint32_t)jx ^ r7) << 4), o[0], o[1], o[2], o[3]);
                    }
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
        } else if (warp < 4 + RW2_ROW_WARPS) {       // (warps 12-15 have no role in this kernel)
            // ===================== epilogue groups: row-owner phase, then channel-pair phase, per 64-channel block ==========
            // Row phase (lane = row): accumulator -> bf16 da1 in the warp's staging buffer.  Channel phase (lane = channel
            // pair, the warp's 32 rows in order): z1 comes straight from global memory -- 128 contiguous bytes per row per
            // warp, issued before the accumulator is awaited -- the gate is applied to the staged values in place, sum g
            // and sum g (z1 - mu) accumulate in a fixed order (deterministic, no atomics), and the block leaves by TMA.
            const int ew = warp - 4, grp = ew >> 2;
            const int nblk = H >> 6;
            uint8_t* stg = smem + L.stg_off + (size_t)ew * RW2_STG;
            float s1[2][2] = {{0.f, 0.f}, {0.f, 0.f}}, s2[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
            uint32_t sb = 0, ti = 0;
            long long w_acc = 0, t_work = 0, t0;
            for (int tile = blockIdx.x; tile < sh.num_tiles; tile += gridDim.x, ++ti) {
                const uint32_t a = ti % NACC, ua = ti / NACC;
                if ((int)a != grp) continue;
                const int64_t r0 = (int64_t)tile * TILE_ROWS;
                const int64_t left = sh.rows - r0;
                const int cnt = left < TILE_ROWS ? (int)left : TILE_ROWS;
                const int64_t rbase = r0 + q * 32;
                const bool any_row = q * 32 < cnt;
                uint32_t zz[32];
                // rows past the end of the tensor re-read the last valid row (in bounds; masked in the channel phase)
                const int64_t rfirst = rbase < sh.rows ? rbase : sh.rows - 1;
                const int rmax = (int)((sh.rows - 1 - rfirst) < 31 ? (sh.rows - 1 - rfirst) : 31);
                auto load_z = [&](int b) {
                    const uint32_t* zp = reinterpret_cast<const uint32_t*>(p.z1 + rfirst * H + b * 64) + lane;
#pragma unroll
                    for (int r = 0; r < 32; ++r) {
                        zz[r] = __ldg(zp);
                        if (r < rmax) zp += H >> 1;
                    }
                };
                load_z(0);
                {   // pull this warp's z1 rows of the group's NEXT tile towards L2 (lane = row)
                    const int64_t nrow = rbase + (int64_t)NACC * gridDim.x * TILE_ROWS + lane;
                    if (nrow < sh.rows) {
                        tc::prefetch_l2(p.z1 + nrow * H);
                        if (nblk > 1) tc::prefetch_l2(p.z1 + nrow * H + 64);
                    }
                }
                t0 = clock64();
                if (lane == 0) tc::mbar_wait(&acc_full[a], ua & 1);
                __syncwarp();
                w_acc += clock64() - t0;
                tc::tc_fence_after();
                t0 = clock64();
                const uint32_t taddr = tmem_base + a * 128 + ((uint32_t)(q * 32) << 16);
                // (one copy of the block body: unrolled over the blocks the kernel's hot code outgrew the instruction cache --
                // stall_no_inst was the top stall reason in ncu)
#pragma unroll 1
                for (int b = 0; b < nblk; ++b) {
                    {
                        const int ch = b * 64 + 2 * lane;
                        const float sc0 = __ldg(p.sc + ch), sc1 = __ldg(p.sc + ch + 1), sf0 = __ldg(p.sf + ch), sf1 = __ldg(p.sf + ch + 1);
                        const float mu0 = __ldg(p.mu + ch), mu1 = __ldg(p.mu + ch + 1);
                        float t1a = 0.f, t1b = 0.f, t2a = 0.f, t2b = 0.f;
                        if (lane == 0) tc::tma_store_wait_read<1>();       // the store that last used this buffer has read it
                        __syncwarp();
                        uint8_t* buf = stg + (sb & 1u) * 4096u;
                        const uint32_t rowa = tc::smem_u32(buf) + (uint32_t)lane * 128u;
#pragma unroll
                        for (int cq = 0; cq < 4; ++cq) {           // 16 accumulator columns at a time (z1 of the block is live)
                            float v[16];
                            tc::tmem_ld_32x16(taddr + b * 64 + cq * 16, v);
                            tc::tmem_ld_wait();
                            if (b == nblk - 1 && cq == 3) {
                                tc::tc_fence_before();
                                tc::mbar_arrive(&acc_empty[a]);
                            }
#pragma unroll
                            for (int jj = 0; jj < 2; ++jj)
                                tc::sts128(rowa + (uint32_t)(((cq * 2 + jj) ^ (lane & 7)) << 4), tc::pack_bf16(v[8 * jj], v[8 * jj + 1]),
                                           tc::pack_bf16(v[8 * jj + 2], v[8 * jj + 3]), tc::pack_bf16(v[8 * jj + 4], v[8 * jj + 5]),
                                           tc::pack_bf16(v[8 * jj + 6], v[8 * jj + 7]));
                        }
                        __syncwarp();
                        const uint32_t ba = tc::smem_u32(buf) + (uint32_t)(lane & 3) * 4u;
                        const uint32_t cb = (uint32_t)(lane >> 2);
#pragma unroll
                        for (int r = 0; r < 32; ++r) {
                            const uint32_t ad = ba + (uint32_t)r * 128u + ((cb ^ (uint32_t)(r & 7)) << 4);
                            const uint32_t w = tc::lds32(ad);
                            const float z0 = __uint_as_float(zz[r] << 16), z1v = __uint_as_float(zz[r] & 0xFFFF0000u);
                            const bool ok = q * 32 + r < cnt;
                            // the same fma as the forward's a1 = relu(z1 * sc + sf)
                            const uint32_t w0 = (ok && fmaf(z0, sc0, sf0) > 0.f) ? (w << 16) : 0u;
                            const uint32_t w1 = (ok && fmaf(z1v, sc1, sf1) > 0.f) ? (w & 0xFFFF0000u) : 0u;
                            tc::sts32(ad, (w0 >> 16) | w1);
                            const float g0 = __uint_as_float(w0), g1 = __uint_as_float(w1);
                            t1a += g0;
                            t1b += g1;
                            t2a = fmaf(g0, z0 - mu0, t2a);
                            t2b = fmaf(g1, z1v - mu1, t2b);
                        }
                        if (b == 0) s1[0][0] += t1a, s1[0][1] += t1b, s2[0][0] += t2a, s2[0][1] += t2b;
                        else s1[1][0] += t1a, s1[1][1] += t1b, s2[1][0] += t2a, s2[1][1] += t2b;
                        if (b + 1 < nblk) load_z(b + 1);
                        rows_stage_store(&tm_g, buf, lane, b * 64, (int)rbase, any_row);
                        ++sb;
                    }
                }
                t_work += clock64() - t0;
            }
            if (lane == 0) tc::tma_store_wait_all<0>();
            __syncwarp();
            {   // the CTA's partial sums: the eight warps' values meet in their (now idle) staging buffers and are added in
                // warp order -> one [2][H] row per CTA for the fixed-order reduction kernel
                float* red = reinterpret_cast<float*>(stg);
#pragma unroll
                for (int b = 0; b < 2; ++b)
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int ch = b * 64 + 2 * lane + i;
                        if (ch < H) red[ch] = s1[b][i], red[H + ch] = s2[b][i];
                    }
                tc::named_bar_sync(RW2_BAR_RED, RW2_ROW_WARPS * 32);
                const int t = threadIdx.x - 4 * 32;
                if (p.stat_partials && t < 2 * H) {
                    float acc = 0.f;
#pragma unroll
                    for (int w = 0; w < RW2_ROW_WARPS; ++w) acc += reinterpret_cast<const float*>(smem + L.stg_off + (size_t)w * RW2_STG)[t];
                    p.stat_partials[(size_t)blockIdx.x * 2 * H + t] = acc;
                }
            }
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

// stats[j] = sum over parts of s1; stats[H + j] = rstd[j] * sum over parts of s2   (fixed order, fp64)
__global__ void k_rows_bwd2_reduce(const float* __restrict__ partials, int parts, int H, const float* __restrict__ rstd,
                                   float* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= 2 * H) return;
    double acc = 0.0;
    for (int q = 0; q < parts; ++q) acc += (double)partials[(size_t)q * 2 * H + j];
    out[j] = j < H ? (float)acc : (float)(acc * (double)rstd[j - H]);
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

inline int rows_env(const char* name, int dflt, int lo, int hi);
// persistent grid of the row-owner kernels: one CTA per SM, fewer when there are fewer tiles
inline int rows_grid(int64_t rows) {
    const int64_t tiles = (rows + TILE_ROWS - 1) / TILE_ROWS;
    const int max_grid = rows_env("GSATB_ROWS_GRID", GSATB_NUM_SMS, 1, GSATB_NUM_SMS);      // (tests: many tiles per CTA at small sizes)
    return (int)(tiles < max_grid ? tiles : max_grid);
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
    if (nst > Op::MAX_ST) nst = Op::MAX_ST;
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
    const int grid = rows_grid(rows);
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
    OpRowsLin1::Params p{bias, stat_partials, 0, plain_epi()};
    int rc = launch_rows<OpRowsLin1>(w1_bf16_padded, x_bf16, nullptr, nullptr, z1_bf16, true, rows, H, p, st);
    if (rc != GSATB_OK) return rc;
    if (stat_partials) {       // every CTA of the launch wrote both of its partial rows: no memset, parts = 2 x grid
        k_rows_reduce_partials<<<(2 * H + 63) / 64, 64, 0, st>>>(stat_partials, 2 * rows_grid(rows), 2 * H, stats);
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

extern "C" size_t gsatb_gin_rows_stat_partials_elems(int H) { return (size_t)GSATB_NUM_SMS * 2 * H; }

extern "C" int gsatb_gin_rows_bwd2(const float* dh, const uint32_t* posmask, float drop_scale, const void* w2t_bf16_padded,
                                   const void* z1_bf16, const float* bn_scale, const float* bn_shift, const float* mean,
                                   const float* rstd, void* d2_bf16, void* g_bf16, float* stat_partials, float* stats,
                                   int64_t rows, int H, gsatb_stream_t stream) {
    if (rows < 0 || H <= 0) return GSATB_EINVAL;
    if (rows == 0) return GSATB_OK;
    if (!dh || !posmask || !w2t_bf16_padded || !z1_bf16 || !bn_scale || !bn_shift || !mean || !rstd || !d2_bf16 || !g_bf16 ||
        !stat_partials || !stats)
        return GSATB_EINVAL;
    if (H != 64 && H != 128) return GSATB_ESHAPE;
    if (rows > (int64_t)INT32_MAX - TILE_ROWS) return GSATB_ESHAPE;
    if (!gsatb_aligned16(z1_bf16)) return GSATB_EALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    RowsShape sh;
    sh.rows = rows;
    sh.num_tiles = (int)((rows + TILE_ROWS - 1) / TILE_ROWS);
    sh.H = H;
    sh.KB = H / 64;
    sh.NSB = 1;
    sh.dbg = profile_buffer();
    const int fixed = sh.KB * BLK_BYTES + RW2_ROW_WARPS * RW2_STG + 512 + 1024;
    int nst = (227 * 1024 - fixed) / ((H / 32) * BLK_BYTES);
    if (nst > 6) nst = 6;
    nst = rows_env("GSATB_ROWS_NST", nst, 1, nst);
    if (nst < 2) return GSATB_ESHAPE;
    sh.NST = nst;
    CUtensorMap tw, tdh, td2, tg;
    int rc = make_weight_tmap(&tw, w2t_bf16_padded, 128, sh.KB * KBLK);
    if (rc != GSATB_OK) return rc;
    rc = make_rows_tmap(&tdh, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dh, rows, H, H, 32, TILE_ROWS);
    if (rc != GSATB_OK) return rc;
    rc = make_act_tmap(&td2, d2_bf16, rows, H, H);
    if (rc != GSATB_OK) return rc;
    rc = make_rows_tmap(&tg, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g_bf16, rows, H, H, 64, 32);
    if (rc != GSATB_OK) return rc;
    const Rows2Smem L = rows2_smem(sh);
    if (L.total > 227 * 1024) return GSATB_ESHAPE;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(k_rows_bwd2, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
            return GSATB_ELAUNCH;
        attr_set = true;
    }
    RowsBwd2Params p{posmask, drop_scale, (const uint16_t*)z1_bf16, bn_scale, bn_shift, mean, stat_partials};
    const int grid = rows_grid(rows);
    k_rows_bwd2<<<grid, RW_THREADS, L.total, st>>>(tw, tdh, td2, tg, sh, p);
    if (cudaPeekAtLastError() != cudaSuccess) return GSATB_ELAUNCH;
    k_rows_bwd2_reduce<<<(2 * H + 63) / 64, 64, 0, st>>>(stat_partials, grid, H, rstd, stats);      // every CTA wrote its row
    GSATB_CHECK_LAUNCH();
    return GSATB_OK;
}
