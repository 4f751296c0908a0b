// Extractor MLP backward (autograd of reference src/utils/get_model.py:57-68 inside src/run_gsat.py:909-927).
//
//   head  (SIMT, thread = channel, CTA = graph): dlogit -> through w3, Dropout2, ReLU2 and InstanceNorm2 -> dz2 (bf16),
//         plus per-graph partials of d w3
//   bwd1  (tcgen05): dh1 = dz2 W2 (A operand = W2^T), epilogue = Dropout1 / ReLU1 masks + InstanceNorm1 backward
//         (per-graph sums are thread-local, as in the forward) -> dz1 (bf16)
//   bwd0  (tcgen05): d f12 = dz1 W1 (A operand = W1^T) -> fp32 [rows, Kin]; the scatter back to the nodes is the
//         deterministic gsatb_gather_concat_bwd
//   h1 / f12 are re-materialised in bf16 by two elementwise kernels for the weight-gradient GEMMs (dW2 = dz2^T h1,
//   dW1 = dz1^T f12), which are plain library GEMMs issued by the Python side.
// Dropout masks are regenerated from the same counter hash (or read from the injected mask tensors).
#include "tc_ops_common.cuh"

namespace {

using namespace tcg;

// ---- head -------------------------------------------------------------------------------------------------------
// CTA = one graph x one 128-channel block, thread = channel.  The graph's [rows][128 channels] block of xhat2 is staged
// in shared memory with 16-byte cp.async loads, both passes read it at 2 bytes (conflict-free), pass 1 overwrites it in
// place with dz2 and the CTA stores it as whole rows (the first version read / wrote global memory at 2 bytes with a
// row stride and was latency bound at 21 % of its HBM roofline).  Graphs have at most 128 rows here (tile plan).
__global__ void __launch_bounds__(128)
k_ext_bwd_head(const float* __restrict__ dlogit, const uint16_t* __restrict__ xhat2, const float* __restrict__ rstd2,
               const float* __restrict__ w3, const int32_t* __restrict__ seg_ptr, Dropout drop2,
               uint16_t* __restrict__ dz2, float* __restrict__ dw3_part, int H) {
    __shared__ __align__(16) uint16_t tile[TILE_ROWS * 128];
    __shared__ float s_dl[TILE_ROWS];
    const int g = blockIdx.x;
    const int ch0 = blockIdx.y * 128, tid = threadIdx.x, lane = tid & 31;
    const int ch = ch0 + tid;
    const bool ch_ok = ch < H;               // no early exit: the dropout words are drawn warp-collectively
    const int chc = ch_ok ? ch : 0;
    const int cw = min(128, H - ch0);        // channels of this block (multiple of 8)
    const int b0 = __ldg(seg_ptr + g), b1 = __ldg(seg_ptr + g + 1);
    const int n = min(b1 - b0, TILE_ROWS);
    if (n <= 0) {
        if (ch_ok) dw3_part[(int64_t)g * H + ch] = 0.f;
        return;
    }
    const int cpr = cw / 8;                  // 16-byte chunks per staged row
    for (int i = tid; i < n * cpr; i += 128) {
        const int row = i / cpr, k = i % cpr;
        cp_async16(tile + row * 128 + k * 8, xhat2 + (int64_t)(b0 + row) * H + ch0 + k * 8);
    }
    cp_async_commit();
    for (int j = tid; j < n; j += 128) s_dl[j] = __ldg(dlogit + b0 + j);
    const float w = __ldg(w3 + chc);
    const uint32_t dseed = dropout_seed(drop2);
    const bool use_mask = drop2.enabled && drop2.mask != nullptr;
    const float rs = __ldg(rstd2 + (int64_t)g * H + chc);
    cp_async_wait<0>();
    __syncthreads();
    float s1 = 0.f, s2 = 0.f, dw = 0.f;
    uint16_t* col = tile + tid;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        const float inv_n = 1.f / (float)n;
        const float m1 = s1 * inv_n, m2 = s2 * inv_n;
        for (int blk = 0; blk < n; blk += 32) {
            // keep bits of rows b0+blk .. +31 for this channel (word scheme, as ext_fwd2 drew them)
            uint32_t keep = 0xffffffffu;
            if (drop2.enabled && !use_mask)
                keep = dropout_rows32(drop2, (uint32_t)(b0 + blk), (uint32_t)ch >> 5, dseed, lane);   // ch >> 5: warp-uniform
            const int nb = min(32, n - blk);
#pragma unroll 4
            for (int j = 0; j < nb; ++j) {
                const int r = blk + j;
                const float x = bf16_bits_to_float(col[r * 128]);
                const bool k = use_mask ? __ldg(drop2.mask + (int64_t)(b0 + r) * H + chc) != 0 : ((keep >> j) & 1u) != 0;
                const float gate = (x > 0.f && k && ch_ok) ? drop2.scale : 0.f;
                const float dl = s_dl[r];
                const float dxh = dl * w * gate;
                if (pass == 0) {
                    s1 += dxh;
                    s2 = fmaf(dxh, x, s2);
                    dw = fmaf(dl, x * gate, dw);
                } else {
                    col[r * 128] = float_to_bf16_bits(rs * (dxh - m1 - x * m2));
                }
            }
        }
    }
    __syncthreads();
    for (int i = tid; i < n * cpr; i += 128) {
        const int row = i / cpr, k = i % cpr;
        *reinterpret_cast<uint4*>(dz2 + (int64_t)(b0 + row) * H + ch0 + k * 8) =
            *reinterpret_cast<const uint4*>(tile + row * 128 + k * 8);
    }
    if (ch_ok) dw3_part[(int64_t)g * H + ch] = dw;
}

// ---- shared pieces walker (same structure as the forward InstanceNorm) ---------------------------------------------
template <class Piece, class SegEnd>
__device__ __forceinline__ void for_pieces_bwd(uint32_t taddr, const int* bnd, int nseg, Piece piece, SegEnd seg_end) {
    const int cnt = bnd[nseg];
    int s = 0;
    while (s < nseg && bnd[s + 1] == bnd[s]) ++s;
#pragma unroll 1
    for (int c = 0; c * 32 < cnt; ++c) {
        float v[32];
        tc::tmem_ld_32x32(taddr + c * 32, v);
        tc::tmem_ld_wait();
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
    }
}

// ---- bwd1: dh1 = dz2 W2, then masks + InstanceNorm1 backward -> dz1 ---------------------------------------------
struct OpExtBwd1 {
    struct Params {
        const uint16_t* xhat1;   // bf16 [rows, C1]
        const float* rstd1;      // [G, C1]
        Dropout drop1;
        uint16_t* dz1;           // bf16 [rows, C1]
        int C1;
    };
    struct EpiState {};
    static constexpr bool TMA_B = true;      // B operand = dz2, bf16 [rows, H] as it lies in memory
    // staging: the tile's [rows][128 channels] block of xhat1 (bf16, 32 KiB) is fetched with cp.async BEFORE the
    // accumulator is awaited; pass 0 forms the gated gradient e = dh1 * keep * scale * (xhat1 > 0) and its per-graph
    // sums (thread-local: thread = channel), remembering the gates as bits; pass 1 overwrites each xhat1 value in
    // place by dz1 = rstd * (e - mean(e) - xhat1 * mean(e * xhat1)), and the group stores the block as whole rows.
    static constexpr int STAGE_BYTES = 32768;
    __device__ static void epi_init(const Params&, EpiState&, int, bool, bool) {}
    __device__ static void epi_prefetch(const Params& p, const Tiling&, const EpiCtx& cx) {
        epi_sync(cx);      // the previous block's copy-out has finished reading the buffer
        stage_load_async<2>(cx, cx.stage, p.xhat1 + cx.r0 * p.C1, p.C1, 0, cx.cnt);
        cp_async_commit();
    }
    __device__ static void epilogue(const Params& p, const Tiling& tl, EpiState&, const EpiCtx& cx) {
        const int* bnd;
        int g0;
        const int nseg = load_segments(tl, cx, bnd, g0);
        const int cnt = bnd[nseg];
        const int ch = cx.ch;
        const bool ch_ok = cx.ch_ok;
        const int chc = ch_ok ? ch : 0;
        const int64_t r0 = cx.r0;
        const bool use_mask = p.drop1.enabled && p.drop1.mask != nullptr;
        const uint8_t* mk = use_mask ? p.drop1.mask + r0 * p.C1 + chc : nullptr;
        const uint32_t dseed = dropout_seed(p.drop1);
        const uint32_t row0 = (uint32_t)r0;
        uint16_t* xs = reinterpret_cast<uint16_t*>(cx.stage) + cx.gtid;      // [row][128]
        float m1[MAX_SEG], m2[MAX_SEG];
        uint32_t gate[4];
        cp_async_wait<0>();
        epi_sync(cx);          // xhat1 block visible to every channel thread
        const int nchunks = (cnt + 31) >> 5;
        if (nchunks == 0) epi_release_acc(cx);
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
            int s = 0;
            while (s < nseg && bnd[s + 1] == bnd[s]) ++s;
            float a1 = 0.f, a2 = 0.f;
#pragma unroll 1
            for (int c = 0; c < nchunks; ++c) {
                float v[32], xv[32];
                tc::tmem_ld_32x32(cx.taddr + c * 32, v);
#pragma unroll
                for (int j = 0; j < 32; ++j) xv[j] = bf16_bits_to_float(xs[(c * 32 + j) * 128]);
                if ((c + 1) * 32 > cnt) {      // rows past the tile end hold stale shared memory (possibly NaN bits)
#pragma unroll
                    for (int j = 0; j < 32; ++j) xv[j] = (c * 32 + j < cnt) ? xv[j] : 0.f;
                }
                uint32_t gt;
                if (pass == 0) {
                    uint32_t keep = 0xffffffffu;
                    if (p.drop1.enabled) {
                        keep = 0u;
                        if (use_mask) {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                keep |= (__ldg(mk + (int64_t)min(c * 32 + j, cnt - 1) * p.C1) != 0 ? 1u : 0u) << j;
                        } else {       // word scheme (as make_h1 drew it): bit j = keep(row c*32 + j, this channel)
                            keep = dropout_rows32(p.drop1, row0 + (uint32_t)(c * 32), (uint32_t)cx.ch >> 5, dseed, cx.lane);
                        }
                    }
                    gt = 0u;
#pragma unroll
                    for (int j = 0; j < 32; ++j) gt |= (xv[j] > 0.f ? 1u : 0u) << j;
                    gt &= keep;
                    // rows past the tile end hold stale shared memory: never gate them in
                    if ((c + 1) * 32 > cnt) gt &= (1u << (cnt - c * 32)) - 1u;
                    if (c == 0) gate[0] = gt;
                    else if (c == 1) gate[1] = gt;
                    else if (c == 2) gate[2] = gt;
                    else gate[3] = gt;
                } else {
                    gt = c == 0 ? gate[0] : (c == 1 ? gate[1] : (c == 2 ? gate[2] : gate[3]));
                }
                tc::tmem_ld_wait();
                if (pass == 1 && c == nchunks - 1) epi_release_acc(cx);
                // e = gate * scale * v is never formed: the sums run over gv = gate ? v : 0 and the dropout scale is
                // folded into the per-graph constants (pass 0: m = scale * sum / n; pass 1: dz1 = A gv + C xhat + B)
                const int cbeg = c * 32, cend = min(cbeg + 32, cnt);
#pragma unroll 1
                while (s < nseg && bnd[s] < cend) {
                    const int lo = max(bnd[s], cbeg) - cbeg, hi = min(bnd[s + 1], cend) - cbeg;
                    const uint32_t m = (hi - lo >= 32) ? 0xffffffffu : (((1u << (hi - lo)) - 1u) << lo);
                    const uint32_t gm = gt & m;
                    if (pass == 0) {
                        float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
#pragma unroll
                        for (int j = 0; j < 32; j += 2) {
                            const float e0 = ((gm >> j) & 1u) ? v[j] : 0.f, e1 = ((gm >> (j + 1)) & 1u) ? v[j + 1] : 0.f;
                            s1a += e0;
                            s1b += e1;
                            s2a = fmaf(e0, xv[j], s2a);
                            s2b = fmaf(e1, xv[j + 1], s2b);
                        }
                        a1 += s1a + s1b;
                        a2 += s2a + s2b;
                    } else {
                        const float rs = ch_ok ? __ldg(p.rstd1 + (int64_t)(g0 + s) * p.C1 + ch) : 0.f;
                        const float cA = rs * p.drop1.scale, cB = -rs * m1[s], cC = -rs * m2[s];
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float gv = ((gm >> j) & 1u) ? v[j] : 0.f;
                            const uint16_t bits = float_to_bf16_bits(fmaf(cA, gv, fmaf(cC, xv[j], cB)));
                            if ((m >> j) & 1u) xs[(cbeg + j) * 128] = bits;
                        }
                    }
                    if (bnd[s + 1] <= cend) {
                        if (pass == 0) {
                            const float sn = p.drop1.scale / (float)(bnd[s + 1] - bnd[s]);
                            m1[s] = a1 * sn;
                            m2[s] = a2 * sn;
                            a1 = a2 = 0.f;
                        }
                        ++s;
                        while (s < nseg && bnd[s + 1] == bnd[s]) ++s;
                    } else {
                        break;
                    }
                }
            }
        }
        epi_sync(cx);          // dz1 block complete in the staging buffer
        stage_store<2>(cx, cx.stage, p.dz1 + r0 * p.C1, p.C1, 0, cnt);
    }
    __device__ static void epi_finish(const Params&, EpiState&, int, bool, bool, int) {}
};

// ---- bwd0: out = x_bf16 W^T (fp32 out), used for d f12 = dz1 W1 ------------------------------------------------------
struct OpLinearBf16In {
    struct Params {
        float* out;          // fp32 [rows, OUT]
        int ldo;
    };
    struct EpiState {};
    static constexpr bool TMA_B = true;      // B operand = x, bf16 [rows, K] as it lies in memory
    static constexpr int STAGE_BYTES = 16384;
    __device__ static void epi_init(const Params&, EpiState&, int, bool, bool) {}
    __device__ static void epi_prefetch(const Params&, const Tiling&, const EpiCtx&) {}
    __device__ static void epilogue(const Params& p, const Tiling&, EpiState&, const EpiCtx& cx) {
        epi_emit_f32<16>(cx, p.out, p.ldo, [](int, float acc) { return acc; });
    }
    __device__ static void epi_finish(const Params&, EpiState&, int, bool, bool, int) {}
};

// ---- re-materialisation for the weight-gradient GEMMs ----------------------------------------------------------------
__global__ void k_ext_make_h1(const uint16_t* __restrict__ xhat1, Dropout drop1, uint16_t* __restrict__ h1,
                              int64_t rows, int C1) {
    const int64_t chunks = rows * (C1 / 8);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < chunks; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / (C1 / 8);
        const int k = (int)(i - row * (C1 / 8)) * 8;
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(xhat1 + row * C1 + k));
        float v[8];
        unpack8(q, v);
        if (drop1.enabled && !drop1.mask) {      // word scheme: one call covers the 32 channels around k
            const uint32_t bits = dropout_word(drop1, (uint32_t)row, (uint32_t)k >> 5, dropout_seed(drop1)) >> (k & 31);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = (v[j] > 0.f && ((bits >> j) & 1u)) ? v[j] * drop1.scale : 0.f;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                v[j] = (v[j] > 0.f && dropout_keep(drop1, row, k + j, C1)) ? v[j] * drop1.scale : 0.f;
        }
        uint32_t o[4];
        pack8(v, o);
        *reinterpret_cast<uint4*>(h1 + row * C1 + k) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

__global__ void k_ext_make_f12(const float* __restrict__ emb, const int32_t* __restrict__ src,
                               const int32_t* __restrict__ dst, uint16_t* __restrict__ f12, int64_t rows, int H) {
    const int K = src ? 2 * H : H;
    const int64_t chunks = rows * (K / 8);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < chunks; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / (K / 8);
        const int k = (int)(i - row * (K / 8)) * 8;
        const int64_t node = src ? (k < H ? __ldg(src + row) : __ldg(dst + row)) : row;
        float v[8];
        load8_f32(emb + node * H, (src && k >= H) ? k - H : k, H, v);
        uint32_t o[4];
        pack8(v, o);
        *reinterpret_cast<uint4*>(f12 + row * K + k) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

}  // namespace

extern "C" int gsatb_tc_ext_bwd_head(const float* dlogit, const void* xhat2, const float* rstd2, const float* w3,
                                     const int32_t* seg_ptr, const uint8_t* mask2, uint64_t seed, float pdrop,
                                     int training, void* dz2, float* dw3_part, int64_t rows, int64_t G, int H,
                                     gsatb_stream_t stream) {
    if (rows < 0 || G < 0 || H <= 0) return GSATB_EINVAL;
    if (rows == 0 || G == 0) return GSATB_OK;
    if (!dlogit || !xhat2 || !rstd2 || !w3 || !seg_ptr || !dz2 || !dw3_part) return GSATB_EINVAL;
    if (H % 8 != 0) return GSATB_ESHAPE;
    if (!gsatb_aligned16(xhat2) || !gsatb_aligned16(dz2)) return GSATB_EALIGN;
    dim3 grid((unsigned)G, (unsigned)((H + 127) / 128));
    k_ext_bwd_head<<<grid, 128, 0, (cudaStream_t)stream>>>(dlogit, (const uint16_t*)xhat2, rstd2, w3, seg_ptr,
                                                          make_dropout(mask2, seed * 2 + 2, pdrop, training, 1),
                                                          (uint16_t*)dz2, dw3_part, H);
    GSATB_CHECK_LAUNCH();
    return GSATB_OK;
}

extern "C" int gsatb_tc_ext_bwd1(const void* dz2, const void* w2t_bf16, const void* xhat1, const float* rstd1,
                                 const uint8_t* mask1, uint64_t seed, float pdrop, int training,
                                 const int32_t* tile_row, const int32_t* tile_seg, const int32_t* seg_ptr,
                                 int num_tiles, void* dz1, int64_t rows, int H, int C1, gsatb_stream_t stream) {
    if (rows < 0 || H <= 0 || C1 <= 0 || num_tiles < 0) return GSATB_EINVAL;
    if (rows == 0 || num_tiles == 0) return GSATB_OK;
    if (!dz2 || !w2t_bf16 || !xhat1 || !rstd1 || !tile_row || !tile_seg || !seg_ptr || !dz1) return GSATB_EINVAL;
    if (H % 8 != 0 || H > 512) return GSATB_ESHAPE;
    OpExtBwd1::Params p{(const uint16_t*)xhat1, rstd1, make_dropout(mask1, seed * 2 + 1, pdrop, training, 1),
                        (uint16_t*)dz1, C1};
    Tiling tl{rows, num_tiles, tile_row, tile_seg, seg_ptr};
    return launch<OpExtBwd1>(w2t_bf16, tl, H, C1, p, (cudaStream_t)stream, dz2, H);
}

extern "C" int gsatb_tc_linear_bf16in_fwd(const void* x_bf16, int ldx, const void* w_bf16, float* out, int ldo,
                                          int64_t rows, int K, int OUT, gsatb_stream_t stream) {
    if (rows < 0 || K <= 0 || OUT <= 0) return GSATB_EINVAL;
    if (rows == 0) return GSATB_OK;
    if (!x_bf16 || !w_bf16 || !out) return GSATB_EINVAL;
    if (K > 512 || K % 8 != 0 || ldx % 8 != 0) return GSATB_ESHAPE;
    OpLinearBf16In::Params p{out, ldo};
    Tiling tl = uniform_tiling(rows);
    return launch<OpLinearBf16In>(w_bf16, tl, K, OUT, p, (cudaStream_t)stream, x_bf16, ldx);
}

extern "C" int gsatb_tc_ext_make_h1(const void* xhat1, const uint8_t* mask1, uint64_t seed, float pdrop, int training,
                                    void* h1, int64_t rows, int C1, gsatb_stream_t stream) {
    if (rows < 0 || C1 <= 0 || C1 % 8 != 0) return GSATB_EINVAL;
    if (rows == 0) return GSATB_OK;
    if (!xhat1 || !h1) return GSATB_EINVAL;
    int64_t blocks = (rows * (C1 / 8) + 255) / 256;
    if (blocks > (int64_t)GSATB_NUM_SMS * 32) blocks = (int64_t)GSATB_NUM_SMS * 32;
    k_ext_make_h1<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        (const uint16_t*)xhat1, make_dropout(mask1, seed * 2 + 1, pdrop, training, 1), (uint16_t*)h1, rows, C1);
    GSATB_CHECK_LAUNCH();
    return GSATB_OK;
}

extern "C" int gsatb_tc_ext_make_f12(const float* emb, const int32_t* src, const int32_t* dst, void* f12, int64_t rows,
                                     int H, gsatb_stream_t stream) {
    if (rows < 0 || H <= 0 || H % 8 != 0) return GSATB_EINVAL;
    if (rows == 0) return GSATB_OK;
    if (!emb || !f12 || ((src == nullptr) != (dst == nullptr))) return GSATB_EINVAL;
    const int K = src ? 2 * H : H;
    int64_t blocks = (rows * (K / 8) + 255) / 256;
    if (blocks > (int64_t)GSATB_NUM_SMS * 32) blocks = (int64_t)GSATB_NUM_SMS * 32;
    k_ext_make_f12<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(emb, src, dst, (uint16_t*)f12, rows, H);
    GSATB_CHECK_LAUNCH();
    return GSATB_OK;
}
