// Shared pieces of the fused extractor kernels (ext_fused_fwd.cu / ext_fused_bwd.cu): the slot layout of a tile, the
// segment table every role derives from it, the accumulator walk in 32/16/8-column pieces, dropout keep bits.
//
// Tile = a run of whole graphs (segments).  Inside a tile every graph starts at an 8-aligned SLOT and its tail slots
// up to the next multiple of 8 are padding (zero rows in the operand tile): slot = TMEM column of the swap-AB
// accumulators D^T[channel (lane), slot (column)], so an 8 / 16 / 32-column tcgen05.ld piece never straddles two
// graphs and the per-graph InstanceNorm statistics are plain unmasked thread-local sums (a zero row contributes zero).
// A tile holds <= 128 slots and <= 16 graphs; the MMA N is the slot count rounded up to 16.
#pragma once
#include "tc_ops_common.cuh"

namespace extf {

using namespace tcg;

constexpr int TILE_SLOTS = 128;
constexpr int MAX_TSEG = 16;
constexpr int BRICK = 16384;               // one [128 x 64] bf16 SWIZZLE_128B block
constexpr int EXT_THREADS = 640;           // 20 warps: 4 control, 2 + 1 epilogue warpgroups, 1 producer warpgroup

__host__ __device__ inline int pad8(int n) { return (n + 7) & ~7; }
__host__ __device__ inline int pad16(int n) { return (n + 15) & ~15; }

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

template <int W>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, float* v);
template <>
__device__ __forceinline__ void tmem_ld_cols<32>(uint32_t taddr, float* v) { tc::tmem_ld_32x32(taddr, v); }
template <>
__device__ __forceinline__ void tmem_ld_cols<16>(uint32_t taddr, float* v) { tc::tmem_ld_32x16(taddr, v); }
template <>
__device__ __forceinline__ void tmem_ld_cols<8>(uint32_t taddr, float* v) { tc::tmem_ld_32x8(taddr, v); }

template <int W>
struct Width {
    static constexpr int value = W;
};
// Walk the npad (multiple of 8) slots of one segment in pieces of 32, 16 and 8 columns: f(Width<W>, off) with `off` the
// slot offset inside the segment.  A piece never crosses a multiple of 32 of `off`, so the keep-bit word of 32
// consecutive rows (dropout_rows32) covers it.
template <class F>
__device__ __forceinline__ void for_pieces(int npad, F f) {
    int off = 0;
#pragma unroll 1
    for (; off + 32 <= npad; off += 32) f(Width<32>{}, off);
    if (off + 16 <= npad) {
        f(Width<16>{}, off);
        off += 16;
    }
    if (off + 8 <= npad) f(Width<8>{}, off);
}

// byte offset of the 16-byte group of slots [slot, slot + 8) (slot % 8 == 0) of channel row `ch` (0..127) inside an
// MN-major SWIZZLE_128B tile [128 channels (K)][128 slots (N)]: two 16 KiB slabs of 64 slots
__device__ __forceinline__ uint32_t mn_tile_offset(int ch, int slot) {
    return (uint32_t)((slot >> 6) * BRICK + ch * 128 + (((((slot & 63) >> 3)) ^ (ch & 7)) << 4));
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

// Keep bits of rows row0 .. row0 + W - 1 (bit j) for this thread's channel.  Word scheme (regenerated in backward), or
// an injected uint8 mask [rows, C] (parity tests).  kw_cache holds the word of the 32-row group the piece lies in.
struct DropCtx {
    Dropout d;
    uint32_t seed;       // effective seed of this launch
    bool on, use_mask;
    int C;               // row stride of the injected mask
};
__device__ __forceinline__ uint32_t keep_word32(const DropCtx& dc, uint32_t row_base32, int ch, int lane) {
    return dropout_rows32(dc.d, row_base32, (uint32_t)ch >> 5, dc.seed, lane);
}
template <int W>
__device__ __forceinline__ uint32_t keep_bits_mask(const DropCtx& dc, int64_t row, int nvalid, int ch, bool ch_ok) {
    uint32_t m = 0;
#pragma unroll
    for (int j = 0; j < W; ++j)
        if (j < nvalid && ch_ok && __ldg(dc.d.mask + (row + j) * dc.C + ch) != 0) m |= 1u << j;
    return m;
}

}  // namespace extf
