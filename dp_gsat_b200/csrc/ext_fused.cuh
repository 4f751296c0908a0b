// Shared pieces of the fused extractor kernels (ext_fused_fwd.cu / ext_fused_bwd.cu): the slot layout of a tile, the
// segment table every role derives from it, register-resident per-graph epilogue bodies, dropout keep bits, the gather
// producer.
//
// Tile = a run of whole graphs (segments).  Inside a tile every graph starts at an 8-aligned SLOT and its tail slots
// up to the next multiple of 8 are padding (zero rows in the operand tile): slot = TMEM column of the swap-AB
// accumulators D^T[channel (lane), slot (column)], so an 8-column tcgen05.ld block never straddles two graphs and the
// per-graph InstanceNorm statistics are plain unmasked thread-local sums (a zero row contributes zero).
// A tile holds <= 128 slots and <= 16 graphs; the MMA N is the slot count rounded up to 16.
//
// Epilogue shape (ncu-driven, round 2): a graph of NBLK <= 8 blocks is handled by a fully static template instance
// (switch on NBLK): all its accumulator values are loaded with NBLK back-to-back tcgen05.ld.x8 into static registers,
// ONE tcgen05.wait, then both sweeps run from registers with no predicates.  Shared memory is touched through 32-bit
// shared addresses (tc::sts128 / lds32): pointers that went through integer arithmetic make nvcc emit generic accesses
// with 64-bit address math.
#pragma once
#include "tc_ops_common.cuh"

namespace extf {

using namespace tcg;

constexpr int TILE_SLOTS = 128;
constexpr int MAX_TSEG = 16;
constexpr int BRICK = 16384;               // one [128 x 64] bf16 SWIZZLE_128B block
constexpr int EXT_THREADS = 512;           // 16 warps: 4 control, 2 epilogue warpgroups, 1 producer warpgroup
constexpr int EXT_CTL_REGS = 56, EXT_PRO_REGS = 96, EXT_EPI_REGS = 176;      // 128 x 56 + 128 x 96 + 256 x 176 = 64512 <= 65536

__host__ __device__ inline int pad8(int n) { return (n + 7) & ~7; }
__host__ __device__ inline int pad16(int n) { return (n + 15) & ~15; }
__host__ __device__ inline int pad128(int n) { return (n + 127) & ~127; }

// Per-warp view of a tile's segments, held in registers: lane s < nseg owns segment s.
struct SegTable {
    int g0, nseg;        // first graph of the tile, graphs in the tile
    int n, row0, slot0;  // (lane s) rows of segment s, its first global row, its first slot
    int total;           // slots used by the tile (multiple of 8)
};
__device__ __forceinline__ SegTable load_seg_table(const int32_t* __restrict__ tile_seg, const int32_t* __restrict__ seg_ptr,
                                                   int tile, int lane) {
    SegTable t;
    t.g0 = __ldg(tile_seg + tile);
    t.nseg = __ldg(tile_seg + tile + 1) - t.g0;
    const bool in = lane < t.nseg;
    t.row0 = in ? __ldg(seg_ptr + t.g0 + lane) : 0;
    t.n = in ? __ldg(seg_ptr + t.g0 + lane + 1) - t.row0 : 0;
    int incl = pad8(t.n);
#pragma unroll
    for (int off = 1; off < MAX_TSEG; off <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += y;
    }
    t.slot0 = incl - pad8(t.n);
    t.total = __shfl_sync(0xffffffffu, incl, MAX_TSEG - 1);
    return t;
}
// slots used by a tile, for a single thread (MMA issuer)
__device__ __forceinline__ int tile_total_slots(const int32_t* __restrict__ tile_seg, const int32_t* __restrict__ seg_ptr,
                                                int tile) {
    const int g0 = __ldg(tile_seg + tile), g1 = __ldg(tile_seg + tile + 1);
    int tot = 0, prev = __ldg(seg_ptr + g0);
    for (int g = g0; g < g1; ++g) {
        const int nx = __ldg(seg_ptr + g + 1);
        tot += pad8(nx - prev);
        prev = nx;
    }
    return tot;
}

// NB back-to-back 8-column loads into static register positions (every register has exactly one, unconditional,
// defining instruction: run-time selected shapes or skipped loads make nvcc keep the array in local memory)
template <int NB>
__device__ __forceinline__ void tmem_ld_blocks(uint32_t taddr, float (&v)[8 * NB]) {
#pragma unroll
    for (int b = 0; b < NB; ++b) tc::tmem_ld_32x8(taddr + 8 * b, &v[8 * b]);
}

// byte offset of the 16-byte group of slots [8 blk, 8 blk + 8) of channel row `ch` (0..127) inside an MN-major
// SWIZZLE_128B tile [128 channels (K)][128 slots (N)]: two 16 KiB slabs of 64 slots, 128-byte rows, chunk ^= row & 7
__device__ __forceinline__ uint32_t mn_tile_offset_blk(int ch, int blk) {
    return (uint32_t)((blk >> 3) * BRICK + ch * 128 + (((blk & 7) ^ (ch & 7)) << 4));
}

// Transposing reduction over the 32 lanes of a warp: in: a[j] = this lane's (channel's) term of slot j; out: lane l
// holds the sum over all 32 lanes of a[l % W].
template <int W>
__device__ __forceinline__ float transpose_reduce(float (&a)[W], int lane) {
#pragma unroll
    for (int off = W / 2; off >= 1; off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float send = upper ? a[i] : a[i + off];
            const float keep = upper ? a[i + off] : a[i];
            a[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    float r = a[0];
#pragma unroll
    for (int off = W; off < 32; off <<= 1) r += __shfl_xor_sync(0xffffffffu, r, off);
    return r;
}

// ---- dropout keep bits ------------------------------------------------------------------------------------------
// Channel-word scheme (dropout_chan_bits32: thread-local, regenerated in backward from the same effective seed), or an
// injected uint8 mask [rows, C] (parity tests).
struct DropCtx {
    const Dropout* d;    // lives in the kernel parameter bank
    uint32_t seed;       // effective seed of this launch
    bool on, use_mask;
    int C;               // row stride of the injected mask
};
__device__ __forceinline__ DropCtx make_drop_ctx(const Dropout& d, uint32_t seed, int C) {
    DropCtx dc;
    dc.d = &d;
    dc.seed = seed;
    dc.on = d.enabled != 0;
    dc.use_mask = dc.on && d.mask != nullptr;
    dc.C = C;
    return dc;
}
// keep bits of rows row0 .. row0 + 31 of channel ch (bit j; rows >= row0 + nvalid read as dropped in mask mode)
__device__ __forceinline__ uint32_t keep_bits32(const DropCtx& dc, int ch, bool ch_ok, int row0, int nvalid) {
    if (!dc.on) return 0xffffffffu;
    if (dc.use_mask) {
        uint32_t m = 0;
        for (int j = 0; j < 32; ++j)
            if (j < nvalid && ch_ok && __ldg(dc.d->mask + (int64_t)(row0 + j) * dc.C + ch) != 0) m |= 1u << j;
        return m;
    }
    return dropout_chan_bits32(*dc.d, (uint32_t)ch, (uint32_t)row0, dc.seed);
}

// ---- epilogue 1: InstanceNorm (centred input) -> ReLU -> Dropout -> bf16 rows of an MN-major tile -------------------
template <int NB>
__device__ __forceinline__ float sumsq_blocks(const float (&v)[8 * NB]) {
    float qa = 0.f, qb = 0.f, qc = 0.f, qd = 0.f;
#pragma unroll
    for (int j = 0; j < 8 * NB; j += 4) {
        qa = fmaf(v[j], v[j], qa);
        qb = fmaf(v[j + 1], v[j + 1], qb);
        qc = fmaf(v[j + 2], v[j + 2], qc);
        qd = fmaf(v[j + 3], v[j + 3], qd);
    }
    return (qa + qb) + (qc + qd);
}
// h = Dropout(ReLU(z~ * rstd)) * scale for 8 slots of one channel (rs = rstd * scale > 0 commutes with the ReLU) -> bf16,
// one 16-byte shared store
__device__ __forceinline__ void emit_h1_blk(const float* v, float rs, uint32_t bits, uint32_t saddr) {
    uint32_t o[4];
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
        float a = fmaxf(v[i], 0.f) * rs, b = fmaxf(v[i + 1], 0.f) * rs;
        a = ((bits >> i) & 1u) ? a : 0.f;
        b = ((bits >> (i + 1)) & 1u) ? b : 0.f;
        o[i >> 1] = tc::pack_bf16(a, b);
    }
    tc::sts128(saddr, o[0], o[1], o[2], o[3]);
}

// ---- epilogue 2 statistics: sums of d = z - K and d^2 over the n valid slots, around a shift K close to the mean --
template <int NB>
__device__ __forceinline__ void shifted_stats(const float (&v)[8 * NB], int nv_last, float Ksh, float& s1, float& s2) {
#pragma unroll
    for (int j = 0; j < 8 * (NB - 1); ++j) {
        const float d = v[j] - Ksh;
        s1 += d;
        s2 = fmaf(d, d, s2);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {              // the graph's last block: nv_last (1..8) valid slots
        const float d = j < nv_last ? v[8 * (NB - 1) + j] - Ksh : 0.f;
        s1 += d;
        s2 = fmaf(d, d, s2);
    }
}

// ---- gather producer (shared by the forward and the backward kernel) ------------------------------------------------
// One warpgroup (128 threads) fills the B-operand tile of GEMM1: row `slot` = bf16(f12[row] - mean_g f12), f12 =
// emb[src] | emb[dst] (node mode: emb[row]), K-major SWIZZLE_128B, KB1 blocks of [128 slots x 64].
//   1. the tile's row -> node indices go to shared memory with one coalesced load per thread (no dependent global loads
//      later);
//   2. the per-graph mean comes from the graph's CONTIGUOUS node rows weighted by out- / in-degree
//      (sum_e emb[src_e] = sum_v outdeg(v) emb[v]: edges never leave their graph), reduced across the row lanes;
//   3. gathers in batches of GATHER_UNROLL independent rows per thread.
struct GatherArgs {
    const float* emb;
    const int32_t* src;          // null: node mode
    const int32_t* dst;
    const int32_t* node_ptr;     // [G + 1]   (edge mode)
    const int32_t* rowptr_src;   // [N + 1]   out-degree = rowptr_src[v + 1] - rowptr_src[v]
    const int32_t* rowptr_dst;   // [N + 1]   in-degree
    int H, Kin, KB1;
};
constexpr int GATHER_UNROLL = 8;
constexpr int GATHER_SCRATCH = 1024 + 4096;      // bytes: slot -> node table [2][128] i32, partial column sums

// Pull the NEXT tile's inputs into L2 while this one is being produced: its graphs' contiguous node rows of emb
// (first touch comes from HBM), its slice of the src / dst index arrays and its node pointers.
__device__ __forceinline__ void prefetch_l2(const void* p) {
#ifndef GSATB_HOST_SIM
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#endif
}
__device__ __forceinline__ void prefetch_next_tile(const GatherArgs& a, const int32_t* __restrict__ tile_seg,
                                                   const int32_t* __restrict__ seg_ptr, int next_tile, int num_tiles, int pt) {
    if (next_tile >= num_tiles) return;
    const int g0 = __ldg(tile_seg + next_tile), g1 = __ldg(tile_seg + next_tile + 1);
    const int r0 = __ldg(seg_ptr + g0), r1 = __ldg(seg_ptr + g1);
    int v0 = r0, v1 = r1;
    if (a.src) {
        v0 = __ldg(a.node_ptr + g0);
        v1 = __ldg(a.node_ptr + g1);
        for (int i = r0 + pt * 32; i < r1; i += 128 * 32) {      // 128-byte lines of the index slices
            prefetch_l2(a.src + i);
            prefetch_l2(a.dst + i);
        }
    }
    const char* base = reinterpret_cast<const char*>(a.emb + (int64_t)v0 * a.H);
    const int64_t bytes = (int64_t)(v1 - v0) * a.H * 4;
    for (int64_t off = (int64_t)pt * 128; off < bytes; off += 128 * 128) prefetch_l2(base + off);
}

// xtile_s / scratch_s: 32-bit shared addresses
// xkb: bytes of one K-block of the tile (max slots per tile * 128)
__device__ __forceinline__ void produce_x_tile(const GatherArgs& a, const SegTable& tb, uint32_t xtile_s, uint32_t xkb,
                                               uint32_t scratch_s, int pt, int lane, int bar_id) {
    const uint32_t sidx_s = scratch_s, part_s = scratch_s + 1024;
    const bool edge = a.src != nullptr;
    const int nck = a.Kin >> 3;
    const int RP = 128 / nck > 0 ? 128 / nck : 1;
    const int ck = pt % nck, rl = pt / nck;
    const bool active = pt < RP * nck;
    const int k0 = ck * 8;
    const bool second = edge && k0 >= a.H;
    const int kk = second ? k0 - a.H : k0;
    const float* __restrict__ embk = a.emb + kk;
    const uint32_t xbuf_s = xtile_s + (uint32_t)(k0 >> 6) * xkb;
    const int kin = k0 & 63;
    // 1. slot -> node ids
    {
        int row = -1;
        for (int s = 0; s < tb.nseg; ++s) {
            const int n = __shfl_sync(0xffffffffu, tb.n, s), sl0 = __shfl_sync(0xffffffffu, tb.slot0, s),
                      r0 = __shfl_sync(0xffffffffu, tb.row0, s);
            if (pt >= sl0 && pt < sl0 + n) row = r0 + pt - sl0;
        }
        tc::sts32(sidx_s + 4 * pt, (uint32_t)(row < 0 ? -1 : (edge ? __ldg(a.src + row) : row)));
        tc::sts32(sidx_s + 512 + 4 * pt, (uint32_t)(row < 0 ? -1 : (edge ? __ldg(a.dst + row) : row)));
    }
    tc::named_bar_sync(bar_id, 128);
    const uint32_t myidx_s = sidx_s + (second ? 512 : 0);
    for (int s = 0; s < tb.nseg; ++s) {
        const int n = __shfl_sync(0xffffffffu, tb.n, s);
        if (n == 0) continue;
        const int npad = pad8(n);
        const int slot0 = __shfl_sync(0xffffffffu, tb.slot0, s), row0 = __shfl_sync(0xffffffffu, tb.row0, s);
        // 2. per-graph mean of the gathered rows
        if (active) {
            float acc[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = 0.f;
            int v0, v1;
            if (edge) {
                v0 = __ldg(a.node_ptr + tb.g0 + s);
                v1 = __ldg(a.node_ptr + tb.g0 + s + 1);
            } else {
                v0 = row0;
                v1 = v0 + n;
            }
            const int32_t* rp = second ? a.rowptr_dst : a.rowptr_src;
#pragma unroll 4
            for (int v = v0 + rl; v < v1; v += RP) {
                const float w = edge ? (float)(__ldg(rp + v + 1) - __ldg(rp + v)) : 1.f;
                float x[8];
                load8_f32(embk + (int64_t)v * a.H, 0, a.H, x);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = fmaf(w, x[i], acc[i]);
            }
            const uint32_t dst = part_s + 4 * (rl * a.Kin + k0);
            tc::sts128(dst, __float_as_uint(acc[0]), __float_as_uint(acc[1]), __float_as_uint(acc[2]), __float_as_uint(acc[3]));
            tc::sts128(dst + 16, __float_as_uint(acc[4]), __float_as_uint(acc[5]), __float_as_uint(acc[6]), __float_as_uint(acc[7]));
        }
        tc::named_bar_sync(bar_id, 128);
        // every row lane sums the RP partials of its own 8 columns (redundantly: no mean buffer to publish)
        float mu[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) mu[i] = 0.f;
        if (active) {
            const float inv_n = 1.f / (float)n;
            for (int j = 0; j < RP; ++j) {
                const uint4 p0 = tc::lds128(part_s + 4 * (j * a.Kin + k0)), p1 = tc::lds128(part_s + 4 * (j * a.Kin + k0) + 16);
                mu[0] += __uint_as_float(p0.x), mu[1] += __uint_as_float(p0.y), mu[2] += __uint_as_float(p0.z), mu[3] += __uint_as_float(p0.w);
                mu[4] += __uint_as_float(p1.x), mu[5] += __uint_as_float(p1.y), mu[6] += __uint_as_float(p1.z), mu[7] += __uint_as_float(p1.w);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) mu[i] *= inv_n;
        }
        tc::named_bar_sync(bar_id, 128);       // the partials are re-used by the next graph
        // 3. gather, centre, round, store (padding slots of the graph: zero rows)
        if (active) {
            for (int r0 = rl; r0 < npad; r0 += RP * GATHER_UNROLL) {
                float x[GATHER_UNROLL][8];
#pragma unroll
                for (int u = 0; u < GATHER_UNROLL; ++u) {
                    const int r = r0 + u * RP;
                    const int node = (int)tc::lds32(myidx_s + 4 * (slot0 + (r < n ? r : 0)));     // clamped: always a valid row
                    load8_f32(embk + (int64_t)node * a.H, 0, a.H, x[u]);
                }
#pragma unroll
                for (int u = 0; u < GATHER_UNROLL; ++u) {
                    const int r = r0 + u * RP;
                    if (r < npad) {
                        uint32_t o[4] = {0u, 0u, 0u, 0u};
                        if (r < n) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) x[u][i] -= mu[i];
                            pack8(x[u], o);
                        }
                        tc::sts128(xbuf_s + tc::sw128_offset(slot0 + r, kin), o[0], o[1], o[2], o[3]);
                    }
                }
            }
        }
    }
    // slots between the last graph and the MMA width, and the K padding up to the 64-block, are zero
    int N = pad16(tb.total);
    if (N < 16) N = 16;
    if (active)
        for (int r = tb.total + rl; r < N; r += RP) tc::sts128(xbuf_s + tc::sw128_offset(r, kin), 0u, 0u, 0u, 0u);
    const int kpad8 = (a.KB1 * 64 - a.Kin) >> 3;          // < 8 chunks of 8
    for (int i = pt; i < kpad8 * N; i += 128) {
        const int r = i / kpad8, c = a.Kin + (i % kpad8) * 8;
        tc::sts128(xtile_s + (uint32_t)(c >> 6) * xkb + tc::sw128_offset(r, c & 63), 0u, 0u, 0u, 0u);
    }
}

}  // namespace extf
