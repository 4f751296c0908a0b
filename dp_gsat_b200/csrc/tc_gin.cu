// GIN node MLP backward on the tensor cores (autograd of reference src/models/gin.py:55-62 plus the ReLU / Dropout of
// gin.py:50-52):   h = Dropout(ReLU( Linear2( ReLU( BatchNorm1d( Linear1(x) ) ) ) ))
//
//   gin_bwd2: d2 = dh * (h > 0) * 1/(1-p) is formed in the operand load (and written out in bf16 for dW2 / db2);
//             da1 = d2 W2 on tcgen05 (A operand = W2^T); epilogue (thread = channel): reload z1, rebuild
//             a1 = ReLU(BN(z1)) and xhat, g = da1 * (a1 > 0), accumulate the BatchNorm-backward statistics
//             sum(g), sum(g * xhat) per channel thread-locally, store g and a1 in bf16
//   gin_bwd1: dz1 = A*g + B*z1 + C per channel (BatchNorm backward folded into three vectors) is formed in the
//             operand load (and written out in bf16 for dW1 / db1); dx = dz1 W1 on tcgen05 (A operand = W1^T)
// Weight gradients (dW2 = d2^T a1, dW1 = dz1^T x) are plain library GEMMs issued by the Python side.
#include "tc_ops_common.cuh"

namespace {

using namespace tcg;

template <bool MASK>
struct OpGinBwd2 {
    struct Params {
        const float* dh;        // [N, H] upstream gradient of the module output
        const float* h;         // [N, H] saved module output (post ReLU / Dropout)           (MASK == false)
        const uint32_t* posmask;   // [N, ceil(H/32)] sign bits of h written by the forward   (MASK == true)
        float drop_scale;       // 1/(1-p) when dropout was applied, else 1
        const uint16_t* z1;     // bf16 [N, H1] saved Linear1 output
        const float* bn_scale;  // [H1] gamma * rstd          (BatchNorm folded:  a1 = relu(z1 * scale + shift))
        const float* bn_shift;  // [H1] beta - mean * scale
        const float* mean;      // [H1]
        const float* rstd;      // [H1]
        uint16_t* d2;           // bf16 [N, H]   out
        uint16_t* g;            // bf16 [N, H1]  out
        uint16_t* a1;           // bf16 [N, H1]  out, nullable (the forward already keeps a1 for dW2)
        float* stat_partials;   // [gridDim * EPI_GROUPS][2][H1]
        int H, H1;
    };
    struct EpiState {
        float s1, s2;
    };
    // with the sign bits of h the producers fetch 32 + 4 bytes per 8 elements instead of 64, so a whole tile's loads
    // fit in flight at once (UNROLL 8); reading h itself needs two rounds (measured: 2.23 -> 1.66 ms at cfg4)
    static constexpr int UNROLL = MASK ? 8 : 4;
    struct Raw {
        float dh[8];
        float h[MASK ? 1 : 8];
        uint32_t m;
    };
    __device__ static void load8(const Params& p, int64_t grow, int k, int K, Raw& r) {
        load8_f32(p.dh + grow * p.H, k, K, r.dh);
        if (MASK) r.m = __ldg(p.posmask + grow * (int64_t)((p.H + 31) >> 5) + (k >> 5)) >> (k & 31);
        else load8_f32(p.h + grow * p.H, k, K, r.h);
    }
    __device__ static void transform8(const Params& p, Raw& r, int64_t grow, int k, int, uint32_t o[4]) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const bool pos = MASK ? ((r.m >> i) & 1u) != 0 : r.h[MASK ? 0 : i] > 0.f;
            v[i] = pos ? r.dh[i] * p.drop_scale : 0.f;
        }
        pack8(v, o);
        *reinterpret_cast<uint4*>(p.d2 + grow * p.H + k) = make_uint4(o[0], o[1], o[2], o[3]);
    }
    // staging per group: two 32-row chunk buffers of bf16 z1 (cp.async prefetch, 2 x 8 KiB) and two of packed
    // (g bf16 | a1 bf16 << 16) words (2 x 16 KiB) that the group splits into the two bf16 output tensors with
    // 16-byte stores.
    static constexpr bool TMA_B = false;
    static constexpr int STAGE_BYTES = 49152;
    __device__ static void epi_init(const Params&, EpiState& st, int, bool, bool) { st.s1 = st.s2 = 0.f; }
    __device__ static void epi_prefetch(const Params& p, const Tiling&, const EpiCtx& cx) {
        epi_sync(cx);      // the previous block's copy-out has finished reading the buffers
        if (cx.cnt > 0) stage_load_async<2>(cx, cx.stage + 32768, p.z1 + cx.r0 * p.H1, p.H1, 0, min(32, cx.cnt));
        cp_async_commit();
    }
    __device__ static void store_split(const Params& p, const EpiCtx& cx, const uint32_t* buf, int row_lo, int nrows) {
        uint16_t* gg = p.g + (cx.r0 + row_lo) * p.H1 + cx.ch0;
        uint16_t* ga = p.a1 + (cx.r0 + row_lo) * p.H1 + cx.ch0;
        const bool vec = (p.H1 % 8) == 0 && (cx.nch % 8) == 0 &&
                         ((reinterpret_cast<uintptr_t>(gg) | (p.a1 ? reinterpret_cast<uintptr_t>(ga) : 0)) & 15u) == 0;
        if (vec) {
            for (int i = cx.gtid; i < nrows * 16; i += 128) {
                const int row = i >> 4, k = (i & 15) * 8;
                if (k < cx.nch) {
                    const uint4 w0 = *reinterpret_cast<const uint4*>(buf + row * 128 + k);
                    const uint4 w1 = *reinterpret_cast<const uint4*>(buf + row * 128 + k + 4);
                    uint4 og, oa;
                    og.x = __byte_perm(w0.x, w0.y, 0x5410); oa.x = __byte_perm(w0.x, w0.y, 0x7632);
                    og.y = __byte_perm(w0.z, w0.w, 0x5410); oa.y = __byte_perm(w0.z, w0.w, 0x7632);
                    og.z = __byte_perm(w1.x, w1.y, 0x5410); oa.z = __byte_perm(w1.x, w1.y, 0x7632);
                    og.w = __byte_perm(w1.z, w1.w, 0x5410); oa.w = __byte_perm(w1.z, w1.w, 0x7632);
                    *reinterpret_cast<uint4*>(gg + (int64_t)row * p.H1 + k) = og;
                    if (p.a1) *reinterpret_cast<uint4*>(ga + (int64_t)row * p.H1 + k) = oa;
                }
            }
        } else {
            for (int i = cx.gtid; i < nrows * 128; i += 128) {
                const int row = i >> 7, k = i & 127;
                if (k < cx.nch) {
                    gg[(int64_t)row * p.H1 + k] = (uint16_t)(buf[i] & 0xFFFFu);
                    if (p.a1) ga[(int64_t)row * p.H1 + k] = (uint16_t)(buf[i] >> 16);
                }
            }
        }
    }
    __device__ static void epilogue(const Params& p, const Tiling&, EpiState& st, const EpiCtx& cx) {
        const int cnt = cx.cnt;
        const int chc = cx.ch_ok ? cx.ch : 0;
        const float sc = __ldg(p.bn_scale + chc), sf = __ldg(p.bn_shift + chc);
        const float mu = __ldg(p.mean + chc), rs = __ldg(p.rstd + chc);
        float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
        const int nchunks = (cnt + 31) >> 5;
        if (nchunks == 0) {
            cp_async_wait<0>();
            epi_release_acc(cx);
        }
#pragma unroll 1
        for (int c = 0; c < nchunks; ++c) {
            cp_async_wait<0>();
            epi_sync(cx);      // chunk c of z1 is visible; chunk c-1 has been copied out by every thread
            if (c + 1 < nchunks) {
                stage_load_async<2>(cx, cx.stage + 32768 + ((c + 1) & 1) * 8192, p.z1 + cx.r0 * p.H1, p.H1, (c + 1) * 32,
                                    min(32, cnt - (c + 1) * 32));
                cp_async_commit();
            }
            float v[32];
            tc::tmem_ld_32x32(cx.taddr + c * 32, v);
            tc::tmem_ld_wait();
            if (c == nchunks - 1) epi_release_acc(cx);
            const uint16_t* zb = reinterpret_cast<const uint16_t*>(cx.stage + 32768 + (c & 1) * 8192);
            uint32_t* buf = reinterpret_cast<uint32_t*>(cx.stage + (c & 1) * 16384);
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                const bool ok0 = c * 32 + j < cnt && cx.ch_ok, ok1 = c * 32 + j + 1 < cnt && cx.ch_ok;
                const float z0 = ok0 ? bf16_bits_to_float(zb[j * 128 + cx.gtid]) : 0.f;
                const float z1 = ok1 ? bf16_bits_to_float(zb[(j + 1) * 128 + cx.gtid]) : 0.f;
                const float a0 = fmaxf(fmaf(z0, sc, sf), 0.f), a1v = fmaxf(fmaf(z1, sc, sf), 0.f);
                const float g0 = (a0 > 0.f && ok0) ? v[j] : 0.f, g1 = (a1v > 0.f && ok1) ? v[j + 1] : 0.f;
                const float x0 = (z0 - mu) * rs, x1 = (z1 - mu) * rs;
                s1a += g0;
                s1b += g1;
                s2a = fmaf(g0, x0, s2a);
                s2b = fmaf(g1, x1, s2b);
                buf[j * 128 + cx.gtid] = (uint32_t)float_to_bf16_bits(g0) | ((uint32_t)float_to_bf16_bits(a0) << 16);
                buf[(j + 1) * 128 + cx.gtid] = (uint32_t)float_to_bf16_bits(g1) | ((uint32_t)float_to_bf16_bits(a1v) << 16);
            }
            epi_sync(cx);
            store_split(p, cx, buf, c * 32, min(32, cnt - c * 32));
        }
        st.s1 += s1a + s1b;
        st.s2 += s2a + s2b;
    }
    __device__ static void epi_finish(const Params& p, EpiState& st, int ch, bool ch_ok, bool, int grp) {
        if (ch_ok) {
            const size_t part = (size_t)blockIdx.x * EPI_GROUPS + grp;
            p.stat_partials[(part * 2 + 0) * p.H1 + ch] = st.s1;
            p.stat_partials[(part * 2 + 1) * p.H1 + ch] = st.s2;
        }
    }
};

struct OpGinBwd1 {
    struct Params {
        const uint16_t* g;    // bf16 [N, H1]
        const uint16_t* z1;   // bf16 [N, H1]
        const float* cA;      // [H1]  dz1 = cA * g + cB * z1 + cC
        const float* cB;
        const float* cC;
        uint16_t* dz1;        // bf16 [N, H1] out
        float* dx;            // [N, Kin] out
        int H1, Kin;
    };
    struct EpiState {};
    static constexpr int UNROLL = 8;
    struct Raw {
        uint4 g, z;
    };
    __device__ static void load8(const Params& p, int64_t grow, int k, int, Raw& r) {
        r.g = __ldg(reinterpret_cast<const uint4*>(p.g + grow * p.H1 + k));
        r.z = __ldg(reinterpret_cast<const uint4*>(p.z1 + grow * p.H1 + k));
    }
    __device__ static void transform8(const Params& p, Raw& r, int64_t grow, int k, int K, uint32_t o[4]) {
        float g[8], z[8], a[8], b[8], c[8];
        unpack8(r.g, g);
        unpack8(r.z, z);
        load8_f32(p.cA, k, K, a);
        load8_f32(p.cB, k, K, b);
        load8_f32(p.cC, k, K, c);
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] = fmaf(a[i], g[i], fmaf(b[i], z[i], c[i]));
        pack8(g, o);
        *reinterpret_cast<uint4*>(p.dz1 + grow * p.H1 + k) = make_uint4(o[0], o[1], o[2], o[3]);
    }
    static constexpr bool TMA_B = false;
    static constexpr int STAGE_BYTES = 32768;
    __device__ static void epi_init(const Params&, EpiState&, int, bool, bool) {}
    __device__ static void epi_prefetch(const Params&, const Tiling&, const EpiCtx&) {}
    __device__ static void epilogue(const Params& p, const Tiling&, EpiState&, const EpiCtx& cx) {
        epi_emit_f32<32>(cx, p.dx, p.Kin, [](int, float acc) { return acc; });
    }
    __device__ static void epi_finish(const Params&, EpiState&, int, bool, bool, int) {}
};

__global__ void k_reduce_partials_f(const float* __restrict__ partials, int parts, int width, float* __restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= width) return;
    double acc = 0.0;
    for (int q = 0; q < parts; ++q) acc += (double)partials[(size_t)q * width + j];
    out[j] = (float)acc;
}

}  // namespace

extern "C" int gsatb_tc_gin_bwd2(const float* dh, const float* h, const uint32_t* posmask, float drop_scale, const void* w2t_bf16,
                                 const void* z1, const float* bn_scale, const float* bn_shift, const float* mean,
                                 const float* rstd, void* d2, void* g, void* a1, float* stat_partials, float* stats,
                                 int64_t N, int H, int H1, gsatb_stream_t stream) {
    if (N < 0 || H <= 0 || H1 <= 0) return GSATB_EINVAL;
    if (N == 0) return GSATB_OK;
    if (!dh || (!h && !posmask) || !w2t_bf16 || !z1 || !bn_scale || !bn_shift || !mean || !rstd || !d2 || !g ||
        !stat_partials || !stats)
        return GSATB_EINVAL;
    if (H % 8 != 0 || H > 512 || H1 > 128) return GSATB_ESHAPE;
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(stat_partials, 0, (size_t)GSATB_NUM_SMS * EPI_GROUPS * 2 * H1 * sizeof(float), st);
    int rc;
    if (posmask) {
        OpGinBwd2<true>::Params p{dh, h, posmask, drop_scale, (const uint16_t*)z1, bn_scale, bn_shift, mean, rstd,
                                  (uint16_t*)d2, (uint16_t*)g, (uint16_t*)a1, stat_partials, H, H1};
        rc = launch<OpGinBwd2<true>>(w2t_bf16, uniform_tiling(N), H, H1, p, st);
    } else {
        OpGinBwd2<false>::Params p{dh, h, posmask, drop_scale, (const uint16_t*)z1, bn_scale, bn_shift, mean, rstd,
                                   (uint16_t*)d2, (uint16_t*)g, (uint16_t*)a1, stat_partials, H, H1};
        rc = launch<OpGinBwd2<false>>(w2t_bf16, uniform_tiling(N), H, H1, p, st);
    }
    if (rc != GSATB_OK) return rc;
    k_reduce_partials_f<<<(2 * H1 + 127) / 128, 128, 0, st>>>(stat_partials, GSATB_NUM_SMS * EPI_GROUPS, 2 * H1, stats);
    GSATB_CHECK_LAUNCH();
    return GSATB_OK;
}

extern "C" int gsatb_tc_gin_bwd1(const void* g, const void* z1, const float* cA, const float* cB, const float* cC,
                                 const void* w1t_bf16, void* dz1, float* dx, int64_t N, int H1, int Kin,
                                 gsatb_stream_t stream) {
    if (N < 0 || H1 <= 0 || Kin <= 0) return GSATB_EINVAL;
    if (N == 0) return GSATB_OK;
    if (!g || !z1 || !cA || !cB || !cC || !w1t_bf16 || !dz1 || !dx) return GSATB_EINVAL;
    if (H1 % 8 != 0 || H1 > 512) return GSATB_ESHAPE;
    OpGinBwd1::Params p{(const uint16_t*)g, (const uint16_t*)z1, cA, cB, cC, (uint16_t*)dz1, dx, H1, Kin};
    return launch<OpGinBwd1>(w1t_bf16, uniform_tiling(N), H1, Kin, p, (cudaStream_t)stream);
}
