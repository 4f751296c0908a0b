// Dense layers of the path on the 5th-generation tensor cores (tcgen05 / TMEM / TMA), built on tc_gemm.cuh.
//
// Precision contract: operands are rounded to bf16 (8-bit mantissa), products are accumulated in fp32 in TMEM.
// The parity bound for these ops is therefore the documented "bf16 MLP" tolerance (tests/test_gpu_tc.py), not the
// rtol 1e-5 of the fp32 gather / scatter / sampler kernels.
#include "tc_ops_common.cuh"

namespace {

using namespace tcg;

// ------------------------------------------------------------------------------------------------------------
// OpLinear:  out[row, ch] = act( x[row, :] . W[ch, :] + bias[ch] )          (nn.Linear; x fp32, out fp32)
//   optional per-channel column statistics  sum(z), sum(z^2)  over all rows (BatchNorm batch stats), written as
//   per-CTA partials [gridDim][2][OUT] and reduced in a fixed order by k_reduce_partials (deterministic).
// ------------------------------------------------------------------------------------------------------------
struct OpLinear {
    struct Params {
        const float* x;          // fp32 [rows, K] ...
        const uint16_t* x_bf16;  // ... or bf16 [rows, K] (exactly one of the two is given)
        int ldx;
        const float* in_scale;   // [K] optional fused prologue  relu(x * scale + shift)  (BatchNorm + ReLU folded)
        const float* in_shift;
        const float* bias;       // [OUT] nullable
        float* out;
        int ldo;
        int relu_out;
        float* stat_partials;    // nullable
        int OUT;
        Dropout drop;            // applied after the output activation (GIN: dropout(relu(conv)), gin.py:51-52)
    };
    struct EpiState {
        float s1, s2;
    };
    static constexpr int UNROLL = 8;
    struct Raw {
        float v[8];
    };
    __device__ static void load8(const Params& p, int64_t grow, int k, int K, Raw& r) {
        if (p.x_bf16) {
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(p.x_bf16 + grow * p.ldx + k));
            unpack8(q, r.v);
        } else {
            load8_f32(p.x + grow * p.ldx, k, K, r.v);
        }
    }
    __device__ static void transform8(const Params& p, Raw& r, int64_t, int k, int K, uint32_t o[4]) {
        if (p.in_scale) {
            float sc[8], sf[8];
            load8_f32(p.in_scale, k, K, sc);
            load8_f32(p.in_shift, k, K, sf);
#pragma unroll
            for (int i = 0; i < 8; ++i) r.v[i] = fmaxf(fmaf(r.v[i], sc[i], sf[i]), 0.f);
        }
        pack8(r.v, o);
    }
    static constexpr bool TMA_B = false;
    static constexpr int STAGE_BYTES = 32768;
    __device__ static void epi_init(const Params&, EpiState& st, int, bool, bool) { st.s1 = st.s2 = 0.f; }
    __device__ static void epi_prefetch(const Params&, const Tiling&, const EpiCtx&) {}
    // Straight-line code: every column is computed unconditionally (the 32 columns of a chunk are independent
    // instruction streams the scheduler can interleave) and written to the staging buffer; the group then stores the
    // chunk to global memory as whole rows.
    __device__ static void epilogue(const Params& p, const Tiling&, EpiState& st, const EpiCtx& cx) {
        const int ch = cx.ch, cnt = cx.cnt;
        const bool ch_ok = cx.ch_ok;
        const int64_t r0 = cx.r0;
        const float b = (ch_ok && p.bias) ? __ldg(p.bias + ch) : 0.f;
        const bool use_mask = p.drop.enabled && p.drop.mask != nullptr;
        const uint8_t* mk = use_mask ? p.drop.mask + r0 * p.OUT + (ch_ok ? ch : 0) : nullptr;
        const uint32_t dseed = dropout_seed(p.drop);
        uint32_t keepw = 0u;
        float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
        epi_emit_f32<32>(cx, p.out, p.ldo, [&](int col, float acc) {
            const bool ok = col < cnt;
            float z = acc + b;
            const float y = ok ? z : 0.f;
            if (col & 1) {
                s1b += y;
                s2b = fmaf(y, y, s2b);
            } else {
                s1a += y;
                s2a = fmaf(y, y, s2a);
            }
            if (p.relu_out) z = fmaxf(z, 0.f);
            if (p.drop.enabled) {
                // word scheme: the keep bits of 32 rows of this channel are drawn at the first column of each chunk
                if (!use_mask && (col & 31) == 0)
                    keepw = dropout_rows32(p.drop, (uint32_t)(r0 + col), (uint32_t)ch >> 5, dseed, cx.lane);
                const bool k = use_mask ? (ok ? __ldg(mk + (int64_t)col * p.OUT) != 0 : false) : ((keepw >> (col & 31)) & 1u) != 0;
                z = k ? z * p.drop.scale : 0.f;
            }
            return z;
        });
        st.s1 += s1a + s1b;
        st.s2 += s2a + s2b;
    }
    __device__ static void epi_finish(const Params& p, EpiState& st, int ch, bool ch_ok, bool, int grp) {
        if (p.stat_partials && ch_ok) {
            const size_t part = (size_t)blockIdx.x * EPI_GROUPS + grp;
            p.stat_partials[(part * 2 + 0) * p.OUT + ch] = st.s1;
            p.stat_partials[(part * 2 + 1) * p.OUT + ch] = st.s2;
        }
    }
};

// ------------------------------------------------------------------------------------------------------------
// OpLinearBf16:  out[row, ch] = drop(act( x_bf16[row, :] . W[ch, :] + bias[ch] ))    (both Linears of the GIN node MLP)
//   B operand straight from HBM by TMA (x is a bf16 row-major activation tensor: K3's aggregation, or
//   a1 = ReLU(BatchNorm(z1)) from k_bn_relu_bf16), four epilogue groups.  Output bf16 (z1) or fp32 (the layer output
//   h, which K3 gathers in fp32).  The BatchNorm batch statistics sum(z), sum(z^2) are taken from the fp32
//   accumulators BEFORE z is rounded to bf16.
// ------------------------------------------------------------------------------------------------------------
template <bool OUT_BF16>
struct OpLinearBf16 {
    struct Params {
        const float* bias;       // [OUT] nullable
        void* out;               // bf16 or fp32 [rows, OUT]
        int ldo;
        int relu_out;
        float* stat_partials;    // nullable: [gridDim * 4 groups][2][OUT]
        int OUT;
        Dropout drop;
        uint32_t* posmask;       // nullable: [rows, ceil(OUT/32)] bit c%32 of word (row, c/32) = (out[row, c] > 0)
    };
    struct EpiState {
        float s1, s2;
    };
    static constexpr bool TMA_B = true;
    static constexpr int STAGE_BYTES = 16384;
    __device__ static void epi_init(const Params&, EpiState& st, int, bool, bool) { st.s1 = st.s2 = 0.f; }
    __device__ static void epi_prefetch(const Params&, const Tiling&, const EpiCtx&) {}
    __device__ static void epilogue(const Params& p, const Tiling&, EpiState& st, const EpiCtx& cx) {
        const int cnt = cx.cnt;
        const int64_t r0 = cx.r0;
        const float b = (cx.ch_ok && p.bias) ? __ldg(p.bias + cx.ch) : 0.f;
        const bool use_mask = p.drop.enabled && p.drop.mask != nullptr;
        const uint8_t* mk = use_mask ? p.drop.mask + r0 * p.OUT + (cx.ch_ok ? cx.ch : 0) : nullptr;
        const uint32_t dseed = dropout_seed(p.drop);
        uint32_t keepw = 0u, posw = 0u;
        float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
        auto f = [&](int col, float acc) {
            const bool ok = col < cnt;
            float z = acc + b;
            const float y = ok ? z : 0.f;
            if (col & 1) {
                s1b += y;
                s2b = fmaf(y, y, s2b);
            } else {
                s1a += y;
                s2a = fmaf(y, y, s2a);
            }
            if (p.relu_out) z = fmaxf(z, 0.f);
            if (p.drop.enabled) {
                // word scheme: the keep bits of 32 rows of this channel are drawn at the first column of each chunk
                if (!use_mask && (col & 31) == 0)
                    keepw = dropout_rows32(p.drop, (uint32_t)(r0 + col), (uint32_t)cx.ch >> 5, dseed, cx.lane);
                const bool k = use_mask ? (ok ? __ldg(mk + (int64_t)col * p.OUT) != 0 : false) : ((keepw >> (col & 31)) & 1u) != 0;
                z = k ? z * p.drop.scale : 0.f;
            }
            if (p.posmask) {     // sign bits of the layer output for the backward pass (it needs nothing else of h):
                // each thread collects the bits of its channel over the chunk's 32 rows, a warp bit-matrix transpose
                // turns them into one 32-channel word per row
                posw |= ((z > 0.f && cx.ch_ok) ? 1u : 0u) << (col & 31);
                if ((col & 31) == 31) {
                    const uint32_t roww = warp_transpose32(posw, cx.lane);
                    const int row = col - 31 + cx.lane;
                    const int W = (p.OUT + 31) >> 5;       // words per row; warps past the last word column write nothing
                    if (row < cnt && (cx.ch >> 5) < W) p.posmask[(r0 + row) * (int64_t)W + (cx.ch >> 5)] = roww;
                    posw = 0u;
                }
            }
            return z;
        };
        if (OUT_BF16) epi_emit_bf16(cx, reinterpret_cast<uint16_t*>(p.out), p.ldo, f);
        else epi_emit_f32<16>(cx, reinterpret_cast<float*>(p.out), p.ldo, f);
        st.s1 += s1a + s1b;
        st.s2 += s2a + s2b;
    }
    __device__ static void epi_finish(const Params& p, EpiState& st, int ch, bool ch_ok, bool, int grp) {
        if (p.stat_partials && ch_ok) {
            const size_t part = (size_t)blockIdx.x * MAX_GROUPS + grp;
            p.stat_partials[(part * 2 + 0) * p.OUT + ch] = st.s1;
            p.stat_partials[(part * 2 + 1) * p.OUT + ch] = st.s2;
        }
    }
};

// a1 = ReLU(z1 * scale + shift) (BatchNorm folded into two per-channel vectors), bf16 in / bf16 out: the operand of
// the second Linear of the node MLP and of dW2 = d2^T a1 in backward.
__global__ void k_bn_relu_bf16(const uint16_t* __restrict__ z, const float* __restrict__ scale,
                               const float* __restrict__ shift, uint16_t* __restrict__ a, int64_t rows, int C) {
    const int cpr = C / 8;
    const int64_t chunks = rows * cpr;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < chunks; i += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(i % cpr) * 8;
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(z) + i);
        float v[8], sc[8], sf[8];
        unpack8(q, v);
        load8_f32(scale, k, C, sc);
        load8_f32(shift, k, C, sf);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaxf(fmaf(v[j], sc[j], sf[j]), 0.f);
        uint32_t o[4];
        pack8(v, o);
        reinterpret_cast<uint4*>(a)[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// fp32 [OUT, K] (or its transpose) -> zero-padded bf16 [pad128(rows), pad64(cols)]
__global__ void k_prep_weight(const float* __restrict__ w, int OUT, int K, int transpose, __nv_bfloat16* __restrict__ wp,
                              int rows_pad, int cols_pad) {
    const int total = rows_pad * cols_pad;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int r = i / cols_pad, c = i % cols_pad;
        float v = 0.f;
        if (!transpose) {
            if (r < OUT && c < K) v = w[(size_t)r * K + c];
        } else {
            if (r < K && c < OUT) v = w[(size_t)c * K + r];
        }
        wp[i] = __float2bfloat16_rn(v);
    }
}

// out[j] = sum over parts (fixed order, fp64 accumulate) of partials[part][j]
__global__ void k_reduce_partials(const float* __restrict__ partials, int parts, int width, double* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= width) return;
    double acc = 0.0;
    for (int q = 0; q < parts; ++q) acc += (double)partials[(size_t)q * width + j];
    out[j] = acc;
}

}  // namespace

extern "C" int gsatb_tc_prep_weight(const float* w, int OUT, int K, int transpose, void* wp, gsatb_stream_t stream) {
    if (!w || !wp || OUT <= 0 || K <= 0) return GSATB_EINVAL;
    const int rows = transpose ? K : OUT, cols = transpose ? OUT : K;
    const int rows_pad = (rows + 127) / 128 * 128, cols_pad = (cols + 63) / 64 * 64;
    k_prep_weight<<<(rows_pad * cols_pad + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        w, OUT, K, transpose, (__nv_bfloat16*)wp, rows_pad, cols_pad);
    GSATB_CHECK_LAUNCH();
    return GSATB_OK;
}

// Development aid: per-CTA cycle counters of the skeleton's roles ([148][16] int64, zeroed by the caller):
// 0 MMA-thread total, 1 wait B full, 2 wait accumulator free, 3 wait W block, 4 epilogue wait, 5 epilogue work,
// 6 producer wait, 7 producer fill, 8 tiles.  Pass NULL to switch it off.
extern "C" int gsatb_tc_set_profile_buffer(void* buf) {
    tcg::profile_buffer() = (long long*)buf;
    return GSATB_OK;
}

extern "C" size_t gsatb_tc_stat_partials_elems(int OUT) { return (size_t)GSATB_NUM_SMS * tcg::MAX_GROUPS * 2 * OUT; }

extern "C" int gsatb_tc_linear_fwd(const void* x, int x_is_bf16, int ldx, const float* in_scale, const float* in_shift,
                                   const void* w_bf16, const float* bias, float* out, int ldo, int relu_out,
                                   float* stat_partials, double* stats, const uint8_t* drop_mask, uint64_t drop_seed,
                                   float pdrop, int64_t rows, int K, int OUT, gsatb_stream_t stream) {
    if (rows < 0 || K <= 0 || OUT <= 0) return GSATB_EINVAL;
    if (rows == 0) return GSATB_OK;
    if (!x || !w_bf16 || !out) return GSATB_EINVAL;
    if ((in_scale == nullptr) != (in_shift == nullptr)) return GSATB_EINVAL;
    if (stat_partials && (OUT > 128 || !stats)) return GSATB_ESHAPE;
    if (K > 512 || K % 8 != 0 || ldx % (x_is_bf16 ? 8 : 4) != 0) return GSATB_ESHAPE;
    if (!gsatb_aligned16(x) || (in_scale && (!gsatb_aligned16(in_scale) || !gsatb_aligned16(in_shift))))
        return GSATB_EALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    OpLinear::Params p{x_is_bf16 ? nullptr : (const float*)x, x_is_bf16 ? (const uint16_t*)x : nullptr, ldx, in_scale, in_shift, bias, out, ldo, relu_out, stat_partials, OUT,
                       make_dropout(drop_mask, drop_seed, pdrop, pdrop > 0.f, 1)};
    Tiling tl = uniform_tiling(rows);
    if (stat_partials)
        cudaMemsetAsync(stat_partials, 0, gsatb_tc_stat_partials_elems(OUT) * sizeof(float), st);
    int rc = launch<OpLinear>(w_bf16, tl, K, OUT, p, st);
    if (rc != GSATB_OK) return rc;
    if (stat_partials) {
        k_reduce_partials<<<(2 * OUT + 127) / 128, 128, 0, st>>>(stat_partials, GSATB_NUM_SMS * EPI_GROUPS, 2 * OUT, stats);
        GSATB_CHECK_LAUNCH();
    }
    return GSATB_OK;
}

extern "C" int gsatb_tc_linear_bf16_fwd(const void* x_bf16, int ldx, const void* w_bf16, const float* bias, void* out,
                                        int out_is_bf16, int ldo, int relu_out, float* stat_partials, double* stats,
                                        const uint8_t* drop_mask, uint64_t drop_seed, float pdrop, uint32_t* posmask_out,
                                        int64_t rows, int K, int OUT, gsatb_stream_t stream) {
    if (rows < 0 || K <= 0 || OUT <= 0) return GSATB_EINVAL;
    if (rows == 0) return GSATB_OK;
    if (!x_bf16 || !w_bf16 || !out) return GSATB_EINVAL;
    if (stat_partials && (OUT > 128 || !stats)) return GSATB_ESHAPE;
    if (K % 8 != 0 || ldx % 8 != 0) return GSATB_ESHAPE;      // any K: the K loop streams 64-column blocks of both operands
    cudaStream_t st = (cudaStream_t)stream;
    Tiling tl = uniform_tiling(rows);
    if (stat_partials)
        cudaMemsetAsync(stat_partials, 0, gsatb_tc_stat_partials_elems(OUT) * sizeof(float), st);
    const Dropout drop = make_dropout(drop_mask, drop_seed, pdrop, pdrop > 0.f, 1);
    int rc;
    if (out_is_bf16) {
        OpLinearBf16<true>::Params p{bias, out, ldo, relu_out, stat_partials, OUT, drop, posmask_out};
        rc = launch<OpLinearBf16<true>>(w_bf16, tl, K, OUT, p, st, x_bf16, ldx);
    } else {
        OpLinearBf16<false>::Params p{bias, out, ldo, relu_out, stat_partials, OUT, drop, posmask_out};
        rc = launch<OpLinearBf16<false>>(w_bf16, tl, K, OUT, p, st, x_bf16, ldx);
    }
    if (rc != GSATB_OK) return rc;
    if (stat_partials) {
        k_reduce_partials<<<(2 * OUT + 127) / 128, 128, 0, st>>>(stat_partials, GSATB_NUM_SMS * MAX_GROUPS, 2 * OUT, stats);
        GSATB_CHECK_LAUNCH();
    }
    return GSATB_OK;
}

extern "C" int gsatb_bn_relu_bf16(const void* z_bf16, const float* scale, const float* shift, void* a_bf16, int64_t rows,
                                  int C, gsatb_stream_t stream) {
    if (rows < 0 || C <= 0) return GSATB_EINVAL;
    if (rows == 0) return GSATB_OK;
    if (!z_bf16 || !scale || !shift || !a_bf16) return GSATB_EINVAL;
    if (C % 8 != 0) return GSATB_ESHAPE;
    if (!gsatb_aligned16(z_bf16) || !gsatb_aligned16(a_bf16) || !gsatb_aligned16(scale) || !gsatb_aligned16(shift))
        return GSATB_EALIGN;
    int64_t blocks = (rows * (C / 8) + 255) / 256;
    if (blocks > (int64_t)GSATB_NUM_SMS * 32) blocks = (int64_t)GSATB_NUM_SMS * 32;
    k_bn_relu_bf16<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const uint16_t*)z_bf16, scale, shift,
                                                                      (uint16_t*)a_bf16, rows, C);
    GSATB_CHECK_LAUNCH();
    return GSATB_OK;
}
