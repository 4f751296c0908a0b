// Persistent, warp-specialised tcgen05 GEMM skeleton shared by every dense layer of the path (node MLPs, extractor
// MLP, their backward dX products).  "Swap-AB" orientation:
//
//      D^T[out channel (TMEM lane), row (TMEM column)]  =  W[out, K] (A operand)  x  X[row, K]^T (B operand)
//
// so that one epilogue thread owns one output CHANNEL and walks the 128 rows (nodes / edges) of the tile in its
// registers: per-channel reductions over rows -- BatchNorm batch statistics, per-graph InstanceNorm statistics,
// bias gradients -- are thread-local, with no shuffles and no atomics.
//
// Two role layouts (Op::TMA_B), 20 warps either way:
//   warp 0      TMA producer: streams W in [128 x 64] bf16 K-blocks (SWIZZLE_128B) through an mbarrier ring
//   warp 1      MMA issuer: one elected thread, tcgen05.mma kind::f16 (bf16 x bf16 -> fp32 in TMEM), M=128 N=128 K=16
//   warp 2      TMEM allocator (NGRP accumulators of 128 columns)
//   TMA_B = false (the B operand needs a fused prologue: fp32 -> bf16, BatchNorm/ReLU/dropout, row gathers):
//     warps 4-11   epilogue, NGRP = 2 warpgroups (one per accumulator slot)
//     warps 12-19  B-operand producers: global -> fused prologue -> bf16 -> swizzled shared memory
//   TMA_B = true (the B operand is a row-major bf16 tensor as it lies in global memory):
//     warp 3       TMA producer of the B operand ([128 rows x 64] boxes straight into the swizzled layout)
//     warps 4-19   epilogue, NGRP = 4 warpgroups / accumulator slots (the epilogues are issue bound: four warps per
//                  SM sub-partition hide each other's TMEM / shared-memory / dependent-issue latencies)
//
// A tile = up to 128 rows; for the extractor the tiles are graph-aligned (whole graphs per tile) so that the
// per-graph InstanceNorm closes inside the tile.  A TMA box of a short tile also brings in the first rows of the
// next tile; they only feed accumulator columns >= cnt, which no epilogue reads.
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "tc.cuh"

namespace tcg {

constexpr int TILE_ROWS = 128;
constexpr int KBLK = 64;                      // bf16 elements per 128-byte swizzle row
constexpr int BLK_BYTES = TILE_ROWS * 128;    // one [128 x 64] bf16 K-block = 16 KiB
constexpr int EPI_GROUPS = 2;                 // epilogue warpgroups of the producer layout (sizes the BatchNorm partials)
constexpr int MAX_GROUPS = 4;
constexpr int THREADS = 640;
constexpr int EPI_WARP0 = 4, PRO_WARP0 = 12, PRO_WARPS = 8;
constexpr int MAX_SEG = 32;                   // graphs per tile the InstanceNorm epilogues support
constexpr int MISC_PER_GROUP = 6144;          // epilogue scratch per group: segment tables + reductions
constexpr int EPI_BAR0 = 8;                   // named barriers EPI_BAR0 + grp: epilogue-group syncs around the staging buffer
constexpr int CTL_REGS = 56, EPI_REGS = 104, PRO_REGS = 104, EPI4_REGS = 104;   // setmaxnreg budgets (see the kernel)
constexpr int ACC_BAR0 = 3;                   // named barriers ACC_BAR0 + grp: accumulator-full wait of a group

struct Tiling {
    int64_t rows;               // total rows
    int num_tiles;
    const int32_t* tile_row;    // [num_tiles + 1] first row of each tile (nullptr: uniform 128-row tiles)
    const int32_t* tile_seg;    // [num_tiles + 1] first graph of each tile (InstanceNorm epilogues), nullable
    const int32_t* seg_ptr;     // [G + 1] row offsets of the graphs, nullable
    long long* dbg;             // optional per-CTA cycle counters [gridDim][16] (gsatb_tc_set_profile_buffer)
};

struct Shape {
    int K, KB;      // reduction size and its 64-blocks
    int OUT, NMB;   // output channels and their 128-blocks
    int NGRP;       // accumulator slots = epilogue groups (2 or 4)
    int SPT;        // accumulator slots a tile occupies in the chunk-outer order (NMB rounded up to a divisor of NGRP)
    int ORDER;      // 0: chunk-outer (all NMB accumulators of a tile live at once; needs SPT <= NGRP)
                    // 1: block-outer, whole-tile B buffer (NCH == 1) re-used by the NMB channel blocks
                    // 2: block-outer with K-chunks: the producers re-stage the tile's chunks for every block
    int NA;         // W ring stages
    int NBUF;       // B buffers (2 when they fit, else 1)
    int KBC;        // K-blocks per B buffer
    int NCH;        // K-chunks per tile = ceil(KB / KBC)
    int STAGE;      // bytes of epilogue staging buffer per epilogue group (Op::STAGE_BYTES)
};

struct SmemLayout {
    uint32_t a_off, b_off, stage_off, bar_off, misc_off, total;
};

__host__ __device__ inline SmemLayout smem_layout(const Shape& s) {
    SmemLayout l;
    l.a_off = 0;
    l.b_off = l.a_off + (uint32_t)s.NA * BLK_BYTES;
    l.stage_off = l.b_off + (uint32_t)s.NBUF * s.KBC * BLK_BYTES;
    l.bar_off = l.stage_off + (uint32_t)s.NGRP * s.STAGE;
    l.misc_off = l.bar_off + 256;
    l.total = l.misc_off + (uint32_t)s.NGRP * MISC_PER_GROUP + 1024;   // + slack for the manual 1024-byte alignment
    return l;
}

// ---- epilogue context + staging helpers -----------------------------------------------------------------------
// An epilogue thread owns one output CHANNEL (TMEM lane) and sees the tile's rows as TMEM columns, while every
// tensor in global memory is row-major [rows, channels].  Touching global memory straight from that layout means
// 2- or 4-byte accesses at a row stride: 64-128 bytes per warp instruction and one L1 transaction each, which left
// these kernels 3-12x off their HBM roofline (ncu, round 1).  Instead each epilogue group (128 threads = 128
// channels) owns a staging buffer in shared memory laid out [row][128 channels]; channel threads read / write it
// at 2-4 bytes (conflict-free: a warp covers 64-128 contiguous bytes) and the group moves it to / from global
// memory with 16-byte cp.async loads and 16-byte coalesced stores.
struct EpiCtx {
    uint32_t taddr;      // TMEM address of this warp's lane quarter of the accumulator
    int ch, ch0, nch;    // this thread's channel, first channel of the block, valid channels in the block (<= 128)
    bool ch_ok;
    int64_t r0;          // first global row of the tile
    int cnt, tile;       // rows in the tile, tile index
    uint8_t* misc;       // MISC_BYTES / EPI_GROUPS bytes of scratch of this group
    uint8_t* stage;      // Shape::STAGE bytes of staging buffer of this group
    int q, lane, grp, gtid;   // warp quarter, lane, epilogue group, thread index inside the group (== ch - ch0)
    uint64_t* acc_empty;
};
__device__ __forceinline__ void epi_sync(const EpiCtx& c) { tc::named_bar_sync(EPI_BAR0 + c.grp, 128); }
// Every Op::epilogue calls this exactly once, as soon as it has read the accumulator for the last time.
__device__ __forceinline__ void epi_release_acc(const EpiCtx& c) {
    tc::tc_fence_before();
    tc::mbar_arrive(c.acc_empty);
}
#ifdef GSATB_HOST_SIM      // tests/simt: the copy completes at once
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) { memcpy(smem_dst, gsrc, 16); }
__device__ __forceinline__ void cp_async_commit() {}
template <int N>
__device__ __forceinline__ void cp_async_wait() {}
#else
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tc::smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
#endif

// Asynchronously copy rows [row_lo, row_lo + nrows) x channels [ch0, ch0 + nch) of the row-major tensor `g`
// (element size ES, leading dimension ld elements; g points at element (row 0 of the TILE, channel 0)) into `buf`
// laid out [row - row_lo][128 channels].  Caller: cp_async_commit / cp_async_wait + epi_sync before reading.
template <int ES>
__device__ __forceinline__ void stage_load_async(const EpiCtx& c, uint8_t* buf, const void* g, int ld, int row_lo,
                                                 int nrows) {
    constexpr int CPR = 128 * ES / 16, EPC = 16 / ES;     // 16-byte chunks per staged row, elements per chunk
    const uint8_t* gb = reinterpret_cast<const uint8_t*>(g) + ((int64_t)row_lo * ld + c.ch0) * ES;
    const bool vec = ((reinterpret_cast<uintptr_t>(gb) | (uintptr_t)((int64_t)ld * ES)) & 15u) == 0 && (c.nch % EPC) == 0;
    if (vec) {
        const int total = nrows * CPR;
#pragma unroll 4
        for (int i = c.gtid; i < total; i += 128) {
            const int row = i / CPR, k = i % CPR;
            if (k * EPC < c.nch) cp_async16(buf + row * (128 * ES) + k * 16, gb + (int64_t)row * ld * ES + k * 16);
        }
    } else {                                              // unaligned shapes: element-wise, synchronous
        for (int i = c.gtid; i < nrows * 128; i += 128) {
            const int row = i >> 7, k = i & 127;
            if (k < c.nch) {
                if (ES == 2) reinterpret_cast<uint16_t*>(buf)[i] = reinterpret_cast<const uint16_t*>(gb)[(int64_t)row * ld + k];
                else reinterpret_cast<uint32_t*>(buf)[i] = reinterpret_cast<const uint32_t*>(gb)[(int64_t)row * ld + k];
            }
        }
    }
}
// Copy `buf` ([row - row_lo][128 channels]) to rows [row_lo, row_lo + nrows) x channels [ch0, ch0 + nch) of `g`
// with 16-byte coalesced stores.  Caller: epi_sync between the channel threads' writes to buf and this call.
template <int ES>
__device__ __forceinline__ void stage_store(const EpiCtx& c, const uint8_t* buf, void* g, int ld, int row_lo, int nrows) {
    constexpr int CPR = 128 * ES / 16, EPC = 16 / ES;
    uint8_t* gb = reinterpret_cast<uint8_t*>(g) + ((int64_t)row_lo * ld + c.ch0) * ES;
    const bool vec = ((reinterpret_cast<uintptr_t>(gb) | (uintptr_t)((int64_t)ld * ES)) & 15u) == 0 && (c.nch % EPC) == 0;
    if (vec) {
        const int total = nrows * CPR;
#pragma unroll 4
        for (int i = c.gtid; i < total; i += 128) {
            const int row = i / CPR, k = i % CPR;
            if (k * EPC < c.nch)
                *reinterpret_cast<uint4*>(gb + (int64_t)row * ld * ES + k * 16) =
                    *reinterpret_cast<const uint4*>(buf + row * (128 * ES) + k * 16);
        }
    } else {
        for (int i = c.gtid; i < nrows * 128; i += 128) {
            const int row = i >> 7, k = i & 127;
            if (k < c.nch) {
                if (ES == 2) reinterpret_cast<uint16_t*>(gb)[(int64_t)row * ld + k] = reinterpret_cast<const uint16_t*>(buf)[i];
                else reinterpret_cast<uint32_t*>(gb)[(int64_t)row * ld + k] = reinterpret_cast<const uint32_t*>(buf)[i];
            }
        }
    }
}

__device__ __forceinline__ void tile_range(const Tiling& t, int tile, int64_t& r0, int& cnt) {
    if (t.tile_row) {
        r0 = t.tile_row[tile];
        cnt = t.tile_row[tile + 1] - (int)r0;
    } else {
        r0 = (int64_t)tile * TILE_ROWS;
        int64_t left = t.rows - r0;
        cnt = left < TILE_ROWS ? (int)left : TILE_ROWS;
    }
}

// Enumerates the (channel block, K-chunk) steps of one tile in the order of Shape::ORDER.  Used identically by the W
// producer, the B producers and the MMA issuer so that the three rings stay in lock step.
//   f(mb, chk, new_b)   new_b: this step starts a new B buffer (first use of the buffer the producers filled next)
template <class F>
__device__ __forceinline__ void for_steps(const Shape& sh, F f) {
    if (sh.ORDER == 0) {
        for (int chk = 0; chk < sh.NCH; ++chk)
            for (int mb = 0; mb < sh.NMB; ++mb) f(mb, chk, mb == 0);
    } else if (sh.ORDER == 1) {
        for (int mb = 0; mb < sh.NMB; ++mb) f(mb, 0, mb == 0);
    } else {
        for (int mb = 0; mb < sh.NMB; ++mb)
            for (int chk = 0; chk < sh.NCH; ++chk) f(mb, chk, true);
    }
}
// B buffers a tile consumes
__device__ __forceinline__ int b_buffers_per_tile(const Shape& sh) {
    return sh.ORDER == 0 ? sh.NCH : (sh.ORDER == 1 ? 1 : sh.NMB * sh.NCH);
}

// The kernel.  `Op` supplies:
//    struct Params                               (copied by value into the kernel)
//    struct EpiState                             (per-thread state living across tiles)
//    static constexpr bool TMA_B                 role layout (see the top of this file)
//    static constexpr int STAGE_BYTES            epilogue staging buffer per epilogue group
//    TMA_B == false:  struct Raw; static void load8(P, grow, k, K, Raw&)      issue the loads of 8 consecutive k of a row
//                     static void transform8(P, Raw, grow, k, K, uint32_t out[4])     fused prologue -> 8 bf16
//    static void epi_init(P, EpiState&, ch, ch_ok, first)
//    static void epi_prefetch(P, Tiling, EpiCtx)   called BEFORE the accumulator is awaited (async loads into stage)
//    static void epilogue(P, Tiling, EpiState&, EpiCtx)   must call epi_release_acc(ctx) exactly once
//    static void epi_finish(P, EpiState&, ch, ch_ok, last, grp)
template <class Op>
__global__ void __launch_bounds__(THREADS, 1)
k_tc_gemm(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_b, const Tiling tl,
          const Shape sh, const typename Op::Params p) {
    constexpr bool TMA_B = Op::TMA_B;
    constexpr int NGRP = TMA_B ? 4 : 2;
#ifdef GSATB_HOST_SIM
    uint8_t* smem_raw = simt::dyn_smem();
#else
    extern __shared__ uint8_t smem_raw[];
#endif
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const SmemLayout L = smem_layout(sh);
    uint8_t* sA = smem + L.a_off;
    uint8_t* sB = smem + L.b_off;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);
    uint64_t* a_full = bars;             // [NA]  (NA <= 8)
    uint64_t* a_empty = bars + 8;        // [NA]
    uint64_t* b_full = bars + 16;        // [2]
    uint64_t* b_empty = bars + 18;       // [2]
    uint64_t* acc_full = bars + 20;      // [4]
    uint64_t* acc_empty = bars + 24;     // [4]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 28);
    uint8_t* misc = smem + L.misc_off;   // epilogue scratch (segment tables, cross-warp reductions)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&tmap_w);
        if (TMA_B) tc::tma_prefetch_desc(&tmap_b);
        for (int i = 0; i < sh.NA; ++i) {
            tc::mbar_init(&a_full[i], 1);
            tc::mbar_init(&a_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            tc::mbar_init(&b_full[i], TMA_B ? 1 : PRO_WARPS * 32);
            tc::mbar_init(&b_empty[i], 1);
        }
        for (int i = 0; i < MAX_GROUPS; ++i) {
            tc::mbar_init(&acc_full[i], 1);
            tc::mbar_init(&acc_empty[i], 128);
        }
        tc::fence_barrier_init();
    }
    if (warp == 2) {
        tc::tmem_alloc(tmem_slot, NGRP * 128);
        tc::tmem_relinquish();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int slots_per_tile = sh.ORDER == 0 ? sh.SPT : sh.NMB;      // accumulator indices a tile advances by

    // Register budget per role (the kernel is launched with 96 registers x 640 threads): the four control warps keep
    // 56 (32 made the single-thread MMA issue loop spill to local memory, which put ~1000 cycles between MMA
    // batches) and hand the rest to the warps that hold tiles of data in registers.
    //   producer layout: 128 x 56 + 256 x 104 (epilogue) + 256 x 104 (producers) = 60416 <= 61440
    //   TMA-B layout:    128 x 56 + 512 x 104 (epilogue)                        = 60416
    // (each setmaxnreg is the first statement of its role's branch: ptxas allocates registers per region)
    if (warp < 4) {
      tc::reg_dec<CTL_REGS>();
      if (warp == 0) {
        // ===================== TMA producer: W K-blocks =====================
        if (lane == 0) {
            uint32_t ca = 0;
            for (int tile = blockIdx.x; tile < tl.num_tiles; tile += gridDim.x) {
                for_steps(sh, [&](int mb, int chk, bool) {
                    const int kb0 = chk * sh.KBC, kb1 = min(sh.KB, kb0 + sh.KBC);
                    for (int kb = kb0; kb < kb1; ++kb, ++ca) {
                        const uint32_t s = ca % sh.NA, use = ca / sh.NA;
                        tc::mbar_wait(&a_empty[s], (use & 1) ^ 1);
                        tc::mbar_arrive_expect_tx(&a_full[s], BLK_BYTES);
                        tc::tma_load_2d(sA + s * BLK_BYTES, &tmap_w, &a_full[s], kb * KBLK, mb * TILE_ROWS);
                    }
                });
            }
        }
      } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = tc::make_idesc_bf16(128, TILE_ROWS);
            uint32_t ca = 0, it = 0, acc0 = 0;      // W ring counter, B buffer counter, accumulator index of the tile's block 0
            uint32_t buf = 0;
            long long w_b = 0, w_acc = 0, w_a = 0, t_all = clock64(), t0;
            for (int tile = blockIdx.x; tile < tl.num_tiles; tile += gridDim.x, acc0 += slots_per_tile) {
                for_steps(sh, [&](int mb, int chk, bool new_b) {
                    if (new_b) {
                        buf = it % sh.NBUF;
                        t0 = clock64();
                        tc::mbar_wait(&b_full[buf], (it / sh.NBUF) & 1);
                        w_b += clock64() - t0;
                        tc::tc_fence_after();
                        ++it;
                    }
                    const uint32_t b_base = tc::smem_u32(sB + (size_t)buf * sh.KBC * BLK_BYTES);
                    const int kb0 = chk * sh.KBC, kb1 = min(sh.KB, kb0 + sh.KBC);
                    const uint32_t acc = acc0 + mb, slot = acc % NGRP, us = acc / NGRP;
                    if (chk == 0) {
                        t0 = clock64();
                        tc::mbar_wait(&acc_empty[slot], (us & 1) ^ 1);
                        w_acc += clock64() - t0;
                        tc::tc_fence_after();
                    }
                    const uint32_t d_tmem = tmem_base + slot * 128;
                    for (int kb = kb0; kb < kb1; ++kb, ++ca) {
                        const uint32_t s = ca % sh.NA, use = ca / sh.NA;
                        t0 = clock64();
                        tc::mbar_wait(&a_full[s], use & 1);
                        w_a += clock64() - t0;
                        tc::tc_fence_after();
                        const uint64_t a_desc = tc::make_desc_k_sw128(tc::smem_u32(sA + s * BLK_BYTES));
                        const uint64_t b_desc = tc::make_desc_k_sw128(b_base + (kb - kb0) * BLK_BYTES);
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4)      // 4 x (K = 16 bf16 = 32 bytes) inside the 128-byte row
                            tc::mma_bf16_ss(d_tmem, a_desc + (uint64_t)(k4 * 2), b_desc + (uint64_t)(k4 * 2), idesc,
                                            (kb | k4) != 0);
                        tc::mma_commit(&a_empty[s]);        // ring slot free once these MMAs have read it
                    }
                    if (chk == sh.NCH - 1) tc::mma_commit(&acc_full[slot]);       // accumulator complete -> epilogue
                    // last use of this B buffer: ORDER 0: last block of the chunk; 1: last block; 2: every step
                    const bool last_use = sh.ORDER == 2 || mb == sh.NMB - 1;
                    if (last_use) tc::mma_commit(&b_empty[buf]);                  // B buffer free -> producers
                });
            }
            if (tl.dbg) {
                long long* d = tl.dbg + (size_t)blockIdx.x * 16;
                d[0] = clock64() - t_all;
                d[1] = w_b;
                d[2] = w_acc;
                d[3] = w_a;
            }
        }
      } else if (TMA_B && warp == 3) {
        // ===================== TMA producer: B K-blocks (row-major bf16 tensor) =====================
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < tl.num_tiles; tile += gridDim.x) {
                int64_t r0;
                int cnt;
                tile_range(tl, tile, r0, cnt);
                for_steps(sh, [&](int, int chk, bool new_b) {
                    if (!new_b) return;
                    const uint32_t buf = it % sh.NBUF, ub = it / sh.NBUF;
                    ++it;
                    tc::mbar_wait(&b_empty[buf], (ub & 1) ^ 1);
                    uint8_t* bt = sB + (size_t)buf * sh.KBC * BLK_BYTES;
                    const int kb0 = chk * sh.KBC, nkb = min(sh.KB, kb0 + sh.KBC) - kb0;
                    tc::mbar_arrive_expect_tx(&b_full[buf], (uint32_t)nkb * BLK_BYTES);
                    for (int kbl = 0; kbl < nkb; ++kbl)
                        tc::tma_load_2d(bt + (size_t)kbl * BLK_BYTES, &tmap_b, &b_full[buf], (kb0 + kbl) * KBLK, (int)r0);
                });
            }
        }
      }
    } else if (warp < EPI_WARP0 + 4 * NGRP) {
        // ===================== epilogue: NGRP warpgroups, one per accumulator slot =====================
        if (TMA_B) tc::reg_inc<EPI4_REGS>();
        else tc::reg_inc<EPI_REGS>();
        const int grp = (warp - EPI_WARP0) >> 2;
        const int q = (warp - EPI_WARP0) & 3;               // == warp % 4: the TMEM lane quarter this warp may access
        typename Op::EpiState st;
        EpiCtx cx;
        cx.misc = misc + grp * MISC_PER_GROUP;
        cx.stage = smem + L.stage_off + (size_t)grp * sh.STAGE;
        cx.q = q;
        cx.lane = lane;
        cx.grp = grp;
        cx.gtid = q * 32 + lane;
        uint32_t acc0 = 0;
        bool first = true;
        long long w_full = 0, t_epi = 0, t0;
        for (int tile = blockIdx.x; tile < tl.num_tiles; tile += gridDim.x, acc0 += slots_per_tile) {
            tile_range(tl, tile, cx.r0, cx.cnt);
            cx.tile = tile;
            for (int mb = 0; mb < sh.NMB; ++mb) {
                const uint32_t acc = acc0 + mb, slot = acc % NGRP, us = acc / NGRP;
                if ((int)slot != grp) continue;
                cx.ch0 = mb * 128;
                cx.ch = cx.ch0 + cx.gtid;
                cx.ch_ok = cx.ch < sh.OUT;
                cx.nch = min(128, sh.OUT - cx.ch0);
                cx.acc_empty = &acc_empty[slot];
                if (first || sh.NMB > 1) Op::epi_init(p, st, cx.ch, cx.ch_ok, first);
                first = false;
                Op::epi_prefetch(p, tl, cx);
                t0 = clock64();
                tc::group_mbar_wait(q == 0 && lane == 0, &acc_full[slot], us & 1, ACC_BAR0 + grp, 128);
                w_full += clock64() - t0;
                tc::tc_fence_after();
                cx.taddr = tmem_base + slot * 128 + ((uint32_t)(q * 32) << 16);
                t0 = clock64();
                Op::epilogue(p, tl, st, cx);
                t_epi += clock64() - t0;
                if (sh.NMB > 1) Op::epi_finish(p, st, cx.ch, cx.ch_ok, false, grp);
            }
        }
        if (sh.NMB == 1) {
            if (first) Op::epi_init(p, st, cx.gtid, cx.gtid < sh.OUT, true);
            Op::epi_finish(p, st, cx.gtid, cx.gtid < sh.OUT, true, grp);
        }
        if (tl.dbg && grp == 0 && q == 0 && lane == 0) {
            long long* d = tl.dbg + (size_t)blockIdx.x * 16;
            d[4] = w_full;
            d[5] = t_epi;
        }
    } else {
        // ===================== B-operand producers =====================
        if constexpr (!TMA_B) {
            tc::reg_inc<PRO_REGS>();
            const int pt = threadIdx.x - PRO_WARP0 * 32;        // 0..255
            const int r_in = (pt & 31) >> 3, c_in = pt & 7, pw = pt >> 5;
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < tl.num_tiles; tile += gridDim.x) {
                int64_t r0;
                int cnt;
                tile_range(tl, tile, r0, cnt);
                for_steps(sh, [&](int, int chk, bool new_b) {
                    if (!new_b) return;
                    const uint32_t buf = it % sh.NBUF, ub = it / sh.NBUF;
                    ++it;
                    long long t0 = clock64();
                    tc::group_mbar_wait(pt == 0, &b_empty[buf], (ub & 1) ^ 1, 2, PRO_WARPS * 32);
                    long long t1 = clock64();
                    uint8_t* bt = sB + (size_t)buf * sh.KBC * BLK_BYTES;
                    const int kb0 = chk * sh.KBC, nkb = min(sh.KB, kb0 + sh.KBC) - kb0;
                    // work unit = (K-block, group of 4 rows): 8 lanes cover the 8 16-byte chunks of one 128-byte row.
                    // PRO_UNROLL units are loaded back to back before any is transformed, to keep HBM requests in
                    // flight.  Loads are unconditional on clamped coordinates (K % 8 == 0, so an 8-chunk is wholly in
                    // or out) and invalid chunks are zeroed by a select: no branches, all loads stay in flight.
                    const int units = nkb * (TILE_ROWS / 4);
                    constexpr int PRO_UNROLL = Op::UNROLL;
                    for (int u0 = pw; u0 < units; u0 += PRO_WARPS * PRO_UNROLL) {
                        typename Op::Raw raw[PRO_UNROLL];
#pragma unroll
                        for (int j = 0; j < PRO_UNROLL; ++j) {
                            const int u = min(u0 + j * PRO_WARPS, units - 1);
                            const int kbl = u / (TILE_ROWS / 4), rg = u % (TILE_ROWS / 4);
                            const int row = rg * 4 + r_in, k = (kb0 + kbl) * KBLK + c_in * 8;
                            const bool ok = row < cnt && k < sh.K;
                            Op::load8(p, r0 + (ok ? row : 0), ok ? k : 0, sh.K, raw[j]);
                        }
#pragma unroll
                        for (int j = 0; j < PRO_UNROLL; ++j) {
                            const int uj = u0 + j * PRO_WARPS;
                            const int u = min(uj, units - 1);
                            const int kbl = u / (TILE_ROWS / 4), rg = u % (TILE_ROWS / 4);
                            const int row = rg * 4 + r_in, k = (kb0 + kbl) * KBLK + c_in * 8;
                            const bool ok = row < cnt && k < sh.K;
                            uint32_t o[4];
                            Op::transform8(p, raw[j], r0 + (ok ? row : 0), ok ? k : 0, sh.K, o);
                            if (uj < units)
                                *reinterpret_cast<uint4*>(bt + (size_t)kbl * BLK_BYTES + tc::sw128_offset(row, c_in * 8)) =
                                    ok ? make_uint4(o[0], o[1], o[2], o[3]) : make_uint4(0u, 0u, 0u, 0u);
                        }
                    }
                    tc::fence_proxy_async_smem();
                    tc::mbar_arrive(&b_full[buf]);
                    if (tl.dbg && pt == 0) {
                        long long* d = tl.dbg + (size_t)blockIdx.x * 16;
                        d[6] += t1 - t0;
                        d[7] += clock64() - t1;
                        d[8] += 1;
                    }
                });
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) tc::tmem_dealloc(tmem_base, NGRP * 128);
}

// ---- host side ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                        CUtensorMapFloatOOBfill);

inline PFN_tmapEncodeTiled get_encode_fn() {
    static PFN_tmapEncodeTiled fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = (PFN_tmapEncodeTiled)ptr;
    }
    return fn;
}

// W padded bf16 [rows_pad (multiple of 128), k_pad (multiple of 64)] row-major -> [128 x 64] SWIZZLE_128B boxes
inline int make_weight_tmap(CUtensorMap* tm, const void* w, int rows_pad, int k_pad) {
    PFN_tmapEncodeTiled fn = get_encode_fn();
    if (!fn) return GSATB_ELAUNCH;
    cuuint64_t gdim[2] = {(cuuint64_t)k_pad, (cuuint64_t)rows_pad};
    cuuint64_t gstride[1] = {(cuuint64_t)k_pad * 2};
    cuuint32_t box[2] = {KBLK, TILE_ROWS};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? GSATB_OK : GSATB_EINVAL;
}

// row-major bf16 activations [rows, K] (leading dimension ld elements) -> [128 rows x 64] SWIZZLE_128B boxes; rows
// and columns outside the tensor read as zero
inline int make_act_tmap(CUtensorMap* tm, const void* x, int64_t rows, int K, int ld) {
    PFN_tmapEncodeTiled fn = get_encode_fn();
    if (!fn) return GSATB_ELAUNCH;
    if ((reinterpret_cast<uintptr_t>(x) & 15u) != 0 || (ld % 8) != 0) return GSATB_EALIGN;
    cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {KBLK, TILE_ROWS};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(x), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? GSATB_OK : GSATB_EINVAL;
}

inline Shape make_shape(int K, int OUT, int stage_bytes, bool tma_b) {
    Shape s;
    s.K = K;
    s.KB = (K + KBLK - 1) / KBLK;
    s.OUT = OUT;
    s.NMB = (OUT + 127) / 128;
    s.NGRP = tma_b ? 4 : 2;
    s.SPT = s.NMB == 3 ? 4 : s.NMB;
    s.STAGE = stage_bytes;
    // 16 KiB blocks left for the W ring and the B buffers
    const int budget = (227 * 1024 - 256 - 1024 - 1024 - s.NGRP * (MISC_PER_GROUP + stage_bytes)) / BLK_BYTES;
    const int min_a = 3;
    s.NBUF = 2;
    s.KBC = s.KB;
    if (s.SPT <= s.NGRP) {                       // chunk-outer: shrink the K-chunk until two buffers fit
        s.ORDER = 0;
        while (s.KBC > 1 && 2 * s.KBC + min_a > budget) s.KBC = (s.KBC + 1) / 2;
        if (2 * s.KBC + min_a > budget) s.NBUF = 1;
    } else if (2 * s.KB + min_a <= budget) {
        s.ORDER = 1;
    } else if (s.KB + min_a <= budget) {
        s.ORDER = 1;
        s.NBUF = 1;
    } else {
        s.ORDER = 2;
        while (s.KBC > 1 && 2 * s.KBC + min_a > budget) s.KBC = (s.KBC + 1) / 2;
    }
    s.NCH = (s.KB + s.KBC - 1) / s.KBC;
    int na = budget - s.NBUF * s.KBC;
    s.NA = na > 8 ? 8 : na;
    return s;
}

inline long long*& profile_buffer() {
    static long long* buf = nullptr;
    return buf;
}

// b_bf16: the row-major bf16 B tensor [rows, K] (ld_b elements per row) for Op::TMA_B ops, ignored otherwise
template <class Op>
int launch(const void* w_bf16_padded, const Tiling& tl_in, int K, int OUT, const typename Op::Params& p,
           cudaStream_t st, const void* b_bf16 = nullptr, int ld_b = 0) {
    if (tl_in.num_tiles <= 0) return GSATB_OK;
    Tiling tl = tl_in;
    tl.dbg = profile_buffer();
    Shape sh = make_shape(K, OUT, Op::STAGE_BYTES, Op::TMA_B);
    if (sh.NA < 2 || sh.NMB > 4) return GSATB_ESHAPE;      // (K limits are the entry points' business: the K loop is generic)
    CUtensorMap tm, tmb;
    int rc = make_weight_tmap(&tm, w_bf16_padded, sh.NMB * 128, sh.KB * KBLK);
    if (rc != GSATB_OK) return rc;
    if (Op::TMA_B) {
        if (!b_bf16) return GSATB_EINVAL;
        rc = make_act_tmap(&tmb, b_bf16, tl.rows, K, ld_b);
        if (rc != GSATB_OK) return rc;
    } else {
        tmb = tm;
    }
    SmemLayout L = smem_layout(sh);
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(k_tc_gemm<Op>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
            return GSATB_ELAUNCH;
        attr_set = true;
    }
    int grid = tl.num_tiles < GSATB_NUM_SMS ? tl.num_tiles : GSATB_NUM_SMS;
    k_tc_gemm<Op><<<grid, THREADS, L.total, st>>>(tm, tmb, tl, sh, p);
    if (cudaPeekAtLastError() != cudaSuccess) return GSATB_ELAUNCH;
    return GSATB_OK;
}

}  // namespace tcg
