// Extractor MLP forward on the tensor cores (reference src/run_gsat.py:909-927 + src/utils/get_model.py:57-68:
// cat(emb[col], emb[row]) -> Linear -> InstanceNorm(batch[col]) -> ReLU -> Dropout -> Linear -> InstanceNorm -> ReLU
// -> Dropout -> Linear(H, 1)).
//
// Launches over graph-aligned tiles (whole graphs, <= 128 rows per tile, so every per-graph InstanceNorm closes
// inside the tile's accumulator):
//   make_f12: f12 = [emb[src] | emb[dst]] gathered and rounded to bf16 [E, 2H] (also the operand of dW1 in backward)
//   ext_fwd1: f12 (TMA-fed B operand, four epilogue groups) -> GEMM1 on tcgen05 -> per-graph InstanceNorm in the
//             epilogue (thread = channel, the tile's rows are its TMEM columns) -> xhat1 stored as bf16 [E, C1]
//             (+ rstd1 [G, C1] for backward)
//   make_h1 : h1 = Dropout(ReLU(xhat1)) as bf16 [E, C1] (elementwise; also the operand of dW2 in backward)
//   ext_fwd2: h1 (TMA-fed B operand, four epilogue groups) -> GEMM2 -> InstanceNorm -> ReLU -> Dropout -> dot with w3
//             (+ b3) reduced across channels with a shuffle transpose -> one logit per row
// The Linear biases in front of an InstanceNorm cancel exactly (the norm subtracts the per-graph mean), so b1 and b2
// are not read; their gradients are exactly zero.
#include "tc_ops_common.cuh"

namespace {

using namespace tcg;

// Walk one accumulator row (one channel; the tile's rows are its TMEM columns) in 32-column chunks, split into
// (chunk, graph) pieces.  Control flow is uniform across the CTA (graph boundaries are per tile, not per channel) and
// the per-element work is a 32-wide statically unrolled, bit-mask predicated body, which keeps the code small: a
// branchy 32x unroll made the first version of this epilogue instruction-cache bound (ncu: stall_no_inst).
//   piece(c, v, mask, s)   columns j of chunk c with bit j of mask set belong to graph s
//   seg_end(s)             graph s is complete
//   chunk_begin(c) / chunk_end(c)
template <class Piece, class SegEnd, class ChunkBegin, class ChunkEnd>
__device__ __forceinline__ void for_pieces(uint32_t taddr, const int* bnd, int nseg, Piece piece, SegEnd seg_end,
                                           ChunkBegin chunk_begin, ChunkEnd chunk_end) {
    const int cnt = bnd[nseg];
    int s = 0;
    while (s < nseg && bnd[s + 1] == bnd[s]) ++s;
#pragma unroll 1
    for (int c = 0; c * 32 < cnt; ++c) {
        float v[32];
        tc::tmem_ld_32x32(taddr + c * 32, v);
        tc::tmem_ld_wait();
        chunk_begin(c);
        const int cbeg = c * 32, cend = min(cbeg + 32, cnt);
#pragma unroll 1
        while (s < nseg && bnd[s] < cend) {
            const int lo = max(bnd[s], cbeg) - cbeg, hi = min(bnd[s + 1], cend) - cbeg;
            const uint32_t m = (hi - lo >= 32) ? 0xffffffffu : (((1u << (hi - lo)) - 1u) << lo);
            piece(c, v, m, s);
            if (bnd[s + 1] <= cend) {
                seg_end(s);
                ++s;
                while (s < nseg && bnd[s + 1] == bnd[s]) ++s;
            } else {
                break;
            }
        }
        chunk_end(c);
    }
}

// Per-graph InstanceNorm of one accumulator row (one channel).  PyG computes the biased variance of the centred values
// in two sweeps; here ONE sweep accumulates the sums of d = x - K and d^2 around a shift K that is close to the mean
// (the previous graph's mean of the same channel; the tile's first element for the first graph), so that
// var = E[d^2] - E[d]^2 does not cancel, and a second sweep emits.  The accumulator is read twice instead of three
// times (the epilogue is issue bound).
// emit(c, j, col, in, xhat) is called for every column; seg_rstd(s, rstd) once per non-empty graph.
template <class Emit, class SegRstd, class ChunkBegin, class ChunkEnd>
__device__ __forceinline__ void instance_norm_rows(uint32_t taddr, const int* bnd, int nseg, float eps, Emit emit,
                                                   SegRstd seg_rstd, ChunkBegin chunk_begin, ChunkEnd chunk_end) {
    float mean[MAX_SEG], rs[MAX_SEG];
    auto nop_c = [](int) {};
    {
        float a1 = 0.f, a2 = 0.f, K = 0.f;
        bool have_k = false;
        for_pieces(taddr, bnd, nseg,
                   [&](int, const float* v, uint32_t m, int) {
                       if (!have_k) {      // first piece of the tile starts at column 0 of chunk 0
                           K = v[0];
                           have_k = true;
                       }
                       float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
                       for (int j = 0; j < 32; j += 2) {
                           const float d0 = ((m >> j) & 1u) ? v[j] - K : 0.f;
                           const float d1 = ((m >> (j + 1)) & 1u) ? v[j + 1] - K : 0.f;
                           s0 += d0;
                           s1 += d1;
                           q0 = fmaf(d0, d0, q0);
                           q1 = fmaf(d1, d1, q1);
                       }
                       a1 += s0 + s1;
                       a2 += q0 + q1;
                   },
                   [&](int s) {
                       const float inv_n = 1.f / (float)(bnd[s + 1] - bnd[s]);
                       const float md = a1 * inv_n;
                       const float var = fmaxf(a2 * inv_n - md * md, 0.f);
                       const float r = 1.f / sqrtf(var + eps);
                       mean[s] = K + md;
                       rs[s] = r;
                       seg_rstd(s, r);
                       K = K + md;      // shift of the next graph
                       a1 = a2 = 0.f;
                   },
                   nop_c, nop_c);
    }
    for_pieces(taddr, bnd, nseg,
               [&](int c, const float* v, uint32_t m, int s) {
                   const float mu = mean[s], r = rs[s];
                   // emit() must be branch-free: compute unconditionally, predicate only stores on `in`
#pragma unroll
                   for (int j = 0; j < 32; ++j) emit(c, j, c * 32 + j, ((m >> j) & 1u) != 0, (v[j] - mu) * r);
               },
               [](int) {}, chunk_begin, chunk_end);
}

// ------------------------------------------------------------------------------------------------------------
struct OpExtFwd1 {
    struct Params {
        uint16_t* xhat;         // bf16 bits [rows, C]
        float* rstd;            // [G, C]
        int C;
        float eps;
    };
    struct EpiState {};
    static constexpr bool TMA_B = true;      // B operand = f12 = [emb[src] | emb[dst]] (or emb rows) as bf16, from make_f12
    // staging: two buffers of 32 rows x 128 channels bf16 (16 KiB per group); the emit pass of the InstanceNorm fills
    // one 32-row chunk, the group stores it as whole rows while the channel threads fill the other buffer.
    static constexpr int STAGE_BYTES = 16384;
    __device__ static void epi_init(const Params&, EpiState&, int, bool, bool) {}
    __device__ static void epi_prefetch(const Params&, const Tiling&, const EpiCtx&) {}
    __device__ static void epilogue(const Params& p, const Tiling& tl, EpiState&, const EpiCtx& cx) {
        const int* bnd;
        int g0;
        const int nseg = load_segments(tl, cx, bnd, g0);
        const int nchunks = (cx.cnt + 31) >> 5;
        if (nchunks == 0) epi_release_acc(cx);
        uint16_t* st16 = reinterpret_cast<uint16_t*>(cx.stage) + cx.gtid;
        instance_norm_rows(
            cx.taddr, bnd, nseg, p.eps,
            [&](int c, int j, int, bool in, float xh) {
                const uint16_t bits = float_to_bf16_bits(xh);
                if (in) st16[((c & 1) * 32 + j) * 128] = bits;
            },
            [&](int s, float r) {
                if (cx.ch_ok) p.rstd[(int64_t)(g0 + s) * p.C + cx.ch] = r;
            },
            [&](int c) {
                if (c == nchunks - 1) epi_release_acc(cx);     // last read of the accumulator has completed
            },
            [&](int c) {
                epi_sync(cx);
                stage_store<2>(cx, cx.stage + (c & 1) * 8192, p.xhat + cx.r0 * p.C, p.C, c * 32, min(32, cx.cnt - c * 32));
            });
    }
    __device__ static void epi_finish(const Params&, EpiState&, int, bool, bool, int) {}
};

// ------------------------------------------------------------------------------------------------------------
struct OpExtFwd2 {
    struct Params {
        const float* w3;         // [H]
        const float* b3;         // [1] nullable
        uint16_t* xhat2;         // bf16 bits [rows, H]
        float* rstd2;            // [G, H]
        Dropout drop2;
        float* logit;            // [rows]
        int H;
        float eps;
    };
    struct EpiState {
        uint32_t par;
    };
    static constexpr bool TMA_B = true;       // B operand = h1 = Dropout(ReLU(xhat1)), bf16 [rows, C1], written by make_h1
    static constexpr int STAGE_BYTES = 16384;      // as OpExtFwd1: two 32-row bf16 chunk buffers for xhat2
    __device__ static void epi_init(const Params&, EpiState& st, int, bool, bool first) {
        if (first) st.par = 0;
    }
    __device__ static void epi_prefetch(const Params&, const Tiling&, const EpiCtx&) {}
    __device__ static void epilogue(const Params& p, const Tiling& tl, EpiState& st, const EpiCtx& cx) {
        const int* bnd;
        int g0;
        const int nseg = load_segments(tl, cx, bnd, g0);
        const int cnt = cx.cnt, ch = cx.ch, q = cx.q, lane = cx.lane;
        const bool ch_ok = cx.ch_ok;
        const int64_t r0 = cx.r0;
        const int nchunks = (cnt + 31) >> 5;
        if (nchunks == 0) epi_release_acc(cx);
        float* red = reinterpret_cast<float*>(cx.misc + 1024) + st.par * 512;   // [4 warps][128 columns]
        for (int c = nchunks; c < 4; ++c) red[q * 128 + c * 32 + lane] = 0.f;   // chunks the norm skips
        const float w3 = ch_ok ? __ldg(p.w3 + ch) : 0.f;
        uint16_t* st16 = reinterpret_cast<uint16_t*>(cx.stage) + cx.gtid;
        const bool use_mask = p.drop2.enabled && p.drop2.mask != nullptr;
        const uint32_t dseed = dropout_seed(p.drop2);
        uint32_t keepw = 0xffffffffu;      // word scheme: keep bits of the chunk's 32 rows for this channel
        float a[32];
        instance_norm_rows(
            cx.taddr, bnd, nseg, p.eps,
            [&](int c, int j, int col, bool in, float xh) {
                const uint16_t bits = float_to_bf16_bits(xh);
                if (in) st16[((c & 1) * 32 + j) * 128] = bits;
                float h = fmaxf(xh, 0.f);
                const bool keep = use_mask ? (in && ch_ok ? __ldg(p.drop2.mask + (r0 + col) * p.H + ch) != 0 : false)
                                           : ((keepw >> j) & 1u) != 0;
                h = (!p.drop2.enabled || keep) ? h * p.drop2.scale : 0.f;
                a[j] = (in && ch_ok) ? h * w3 : a[j];
            },
            [&](int s, float r) {
                if (ch_ok) p.rstd2[(int64_t)(g0 + s) * p.H + ch] = r;
            },
            [&](int c) {
                if (c == nchunks - 1) epi_release_acc(cx);
                if (p.drop2.enabled && !use_mask)
                    keepw = dropout_rows32(p.drop2, (uint32_t)(r0 + c * 32), (uint32_t)ch >> 5, dseed, lane);   // warp-uniform
#pragma unroll
                for (int j = 0; j < 32; ++j) a[j] = 0.f;
            },
            [&](int c) {
                // transpose-reduce: lane l ends with the sum over the warp's 32 channels of column c*32 + l
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) {
                    const bool upper = (lane & off) != 0;
#pragma unroll
                    for (int i = 0; i < off; ++i) {
                        const float send = upper ? a[i] : a[i + off];
                        const float keep = upper ? a[i + off] : a[i];
                        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                    }
                }
                red[q * 128 + c * 32 + lane] = a[0];
                epi_sync(cx);
                stage_store<2>(cx, cx.stage + (c & 1) * 8192, p.xhat2 + r0 * p.H, p.H, c * 32, min(32, cnt - c * 32));
            });
        epi_sync(cx);
        const int col = q * 32 + lane;
        if (col < cnt)
            p.logit[r0 + col] = red[col] + red[128 + col] + red[256 + col] + red[384 + col] + (p.b3 ? __ldg(p.b3) : 0.f);
        st.par ^= 1u;
    }
    __device__ static void epi_finish(const Params&, EpiState&, int, bool, bool, int) {}
};

}  // namespace

// Host-side greedy packing of consecutive graphs into tiles of <= max_rows rows and <= max_seg graphs.
// seg_ptr_host is the HOST copy of edge_ptr / node_ptr ([G+1]).  tile_row / tile_seg need room for G+1 entries.
extern "C" int gsatb_tile_plan_host(const int32_t* seg_ptr_host, int64_t G, int max_rows, int max_seg,
                                    int32_t* tile_row, int32_t* tile_seg, int32_t* num_tiles) {
    if (!seg_ptr_host || !tile_row || !tile_seg || !num_tiles || G < 0) return GSATB_EINVAL;
    if (max_rows > tcg::TILE_ROWS || max_seg > tcg::MAX_SEG) return GSATB_ESHAPE;
    int32_t nt = 0;
    int64_t g = 0;
    while (g < G) {
        const int32_t r0 = seg_ptr_host[g];
        int64_t g1 = g;
        while (g1 < G && (g1 - g) < max_seg && seg_ptr_host[g1 + 1] - r0 <= max_rows) ++g1;
        if (g1 == g) return GSATB_ESHAPE;      // a single graph does not fit one tile
        tile_row[nt] = r0;
        tile_seg[nt] = (int32_t)g;
        ++nt;
        g = g1;
    }
    tile_row[nt] = G > 0 ? seg_ptr_host[G] : 0;
    tile_seg[nt] = (int32_t)G;
    *num_tiles = nt;
    return GSATB_OK;
}

extern "C" int gsatb_tc_ext_fwd1(const void* f12, const void* w1_bf16, const int32_t* tile_row, const int32_t* tile_seg,
                                 const int32_t* seg_ptr, int num_tiles, void* xhat1, float* rstd1, int64_t rows, int K,
                                 int C1, float eps, gsatb_stream_t stream) {
    if (rows < 0 || K <= 0 || C1 <= 0 || num_tiles < 0) return GSATB_EINVAL;
    if (rows == 0 || num_tiles == 0) return GSATB_OK;
    if (!f12 || !w1_bf16 || !tile_row || !tile_seg || !seg_ptr || !xhat1 || !rstd1) return GSATB_EINVAL;
    if (K % 8 != 0 || K > 512) return GSATB_ESHAPE;
    OpExtFwd1::Params p{(uint16_t*)xhat1, rstd1, C1, eps};
    Tiling tl{rows, num_tiles, tile_row, tile_seg, seg_ptr};
    return launch<OpExtFwd1>(w1_bf16, tl, K, C1, p, (cudaStream_t)stream, f12, K);
}

extern "C" int gsatb_tc_ext_fwd2(const void* h1, const void* w2_bf16, const float* w3, const float* b3,
                                 const uint8_t* mask2, uint64_t seed, float pdrop, int training,
                                 const int32_t* tile_row, const int32_t* tile_seg, const int32_t* seg_ptr,
                                 int num_tiles, void* xhat2, float* rstd2, float* logit, int64_t rows, int C1, int H,
                                 float eps, gsatb_stream_t stream) {
    if (rows < 0 || H <= 0 || C1 <= 0 || num_tiles < 0) return GSATB_EINVAL;
    if (rows == 0 || num_tiles == 0) return GSATB_OK;
    if (!h1 || !w2_bf16 || !w3 || !tile_row || !tile_seg || !seg_ptr || !xhat2 || !rstd2 || !logit)
        return GSATB_EINVAL;
    if (C1 % 8 != 0 || C1 > 512 || H > 128) return GSATB_ESHAPE;
    OpExtFwd2::Params p{w3, b3, (uint16_t*)xhat2, rstd2, make_dropout(mask2, seed * 2 + 2, pdrop, training, 1), logit, H, eps};
    Tiling tl{rows, num_tiles, tile_row, tile_seg, seg_ptr};
    return launch<OpExtFwd2>(w2_bf16, tl, C1, H, p, (cudaStream_t)stream, h1, C1);
}
