// Attention-aware GINEConv message passing:
//   reference src/models/conv_layers.py:37-66 over PyG GINEConv:  m_e = relu(x[src(e)] + ef[e]) * att[e]
//   out[i] = sum_{e: dst(e) = i} m_e + (1 + eps) * x[i]            (ef = lin(edge_attr), [E, H], computed by the caller)
// backward (CSC rows, by source j):  t_e = att_e * g[dst(e)] * 1[x[j] + ef[e] > 0]
//   dx[j] = sum_e t_e + (1 + eps) g[j] ;  d ef[e] = t_e ;  d att[e] = <relu(x[j] + ef[e]), g[dst(e)]>
// Row-per-sub-warp CSR / CSC walks in edge order (deterministic, no atomics), two edges in flight per row.  This is the
// first, untuned version (the K3 entry-list redesign has not been ported to it); HBM bound: fwd 8NH + 4EH + 12E,
// bwd 12NH + 8EH + 16E.
#include "common.cuh"

namespace {

constexpr int GINE_THREADS = 256;

__device__ __forceinline__ float4 relu_add4(const float4& a, const float4& b) {
    return make_float4(fmaxf(a.x + b.x, 0.f), fmaxf(a.y + b.y, 0.f), fmaxf(a.z + b.z, 0.f), fmaxf(a.w + b.w, 0.f));
}

template <int LPR, int NV, bool HAS_ATT>
__global__ void __launch_bounds__(GINE_THREADS)
k_gine_fwd(const float4* __restrict__ x, const float4* __restrict__ ef, const float* __restrict__ att,
           const int32_t* __restrict__ rowptr, const int32_t* __restrict__ eid, const int32_t* __restrict__ nbr,
           float self_scale, float4* __restrict__ out, int64_t N, int HV) {
    constexpr int RPW = 32 / LPR;
    const int lane = threadIdx.x & 31, sub = lane / LPR, sl = lane % LPR;
    const int64_t warp_global = (blockIdx.x * (int64_t)(GINE_THREADS / 32)) + (threadIdx.x >> 5);
    const int64_t warps_total = (int64_t)gridDim.x * (GINE_THREADS / 32);
    for (int64_t row = warp_global * RPW + sub; row < N; row += warps_total * RPW) {
        const int beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
        float4 acc[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int p = beg; p < end; p += 2) {
            const bool two = p + 1 < end;
            const int n0 = __ldg(nbr + p), e0 = __ldg(eid + p);
            const int n1 = two ? __ldg(nbr + p + 1) : n0, e1 = two ? __ldg(eid + p + 1) : e0;
            const float a0 = HAS_ATT ? __ldg(att + e0) : 1.f, a1 = two ? (HAS_ATT ? __ldg(att + e1) : 1.f) : 0.f;
            float4 x0[NV], f0[NV], x1[NV], f1[NV];
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const int c = sl + v * LPR;
                if (c < HV) {
                    x0[v] = ldg_f4(x + (int64_t)n0 * HV + c);
                    f0[v] = ldg_stream_f4(ef + (int64_t)e0 * HV + c);
                    x1[v] = ldg_f4(x + (int64_t)n1 * HV + c);
                    f1[v] = ldg_stream_f4(ef + (int64_t)e1 * HV + c);
                }
            }
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                if (sl + v * LPR < HV) {
                    fma4(acc[v], a0, relu_add4(x0[v], f0[v]));      // edge order preserved: p, then p + 1
                    if (two) fma4(acc[v], a1, relu_add4(x1[v], f1[v]));
                }
            }
        }
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const int c = sl + v * LPR;
            if (c < HV) {
                fma4(acc[v], self_scale, ldg_stream_f4(x + row * HV + c));
                out[row * HV + c] = acc[v];
            }
        }
    }
}

template <int LPR, int NV, bool HAS_ATT>
__global__ void __launch_bounds__(GINE_THREADS)
k_gine_bwd(const float4* __restrict__ g, const float4* __restrict__ x, const float4* __restrict__ ef,
           const float* __restrict__ att, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ eid,
           const int32_t* __restrict__ nbr, float self_scale, float4* __restrict__ dx, float4* __restrict__ def_,
           float* __restrict__ datt, int64_t N, int HV) {
    constexpr int RPW = 32 / LPR;
    const int lane = threadIdx.x & 31, sub = lane / LPR, sl = lane % LPR;
    const unsigned submask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (sub * LPR));
    const int64_t warp_global = (blockIdx.x * (int64_t)(GINE_THREADS / 32)) + (threadIdx.x >> 5);
    const int64_t warps_total = (int64_t)gridDim.x * (GINE_THREADS / 32);
    for (int64_t row0 = warp_global * RPW; row0 < N; row0 += warps_total * RPW) {
        const int64_t row = row0 + sub;
        const bool live = row < N;
        const int beg = live ? __ldg(rowptr + row) : 0, end = live ? __ldg(rowptr + row + 1) : 0;
        float4 acc[NV], xj[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
            const int c = sl + v * LPR;
            xj[v] = (live && c < HV) ? ldg_stream_f4(x + row * HV + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // the shuffle reduction below is warp-collective inside a sub-warp: every lane of the sub-warp walks the row
        for (int p = beg; p < end; ++p) {
            const int d = __ldg(nbr + p), e = __ldg(eid + p);
            const float a = HAS_ATT ? __ldg(att + e) : 1.f;
            float part = 0.f;
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const int c = sl + v * LPR;
                if (c < HV) {
                    const float4 gg = ldg_f4(g + (int64_t)d * HV + c);
                    const float4 ff = ldg_stream_f4(ef + (int64_t)e * HV + c);
                    const float4 pre = make_float4(xj[v].x + ff.x, xj[v].y + ff.y, xj[v].z + ff.z, xj[v].w + ff.w);
                    float4 t;
                    t.x = pre.x > 0.f ? a * gg.x : 0.f;
                    t.y = pre.y > 0.f ? a * gg.y : 0.f;
                    t.z = pre.z > 0.f ? a * gg.z : 0.f;
                    t.w = pre.w > 0.f ? a * gg.w : 0.f;
                    acc[v].x += t.x; acc[v].y += t.y; acc[v].z += t.z; acc[v].w += t.w;
                    if (def_) def_[(int64_t)e * HV + c] = t;
                    part += fmaxf(pre.x, 0.f) * gg.x + fmaxf(pre.y, 0.f) * gg.y + fmaxf(pre.z, 0.f) * gg.z +
                            fmaxf(pre.w, 0.f) * gg.w;
                }
            }
            if (datt) {
#pragma unroll
                for (int o = LPR / 2; o > 0; o >>= 1) part += __shfl_xor_sync(submask, part, o);
                if (sl == 0) datt[e] = part;
            }
        }
        if (live) {
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const int c = sl + v * LPR;
                if (c < HV) {
                    fma4(acc[v], self_scale, ldg_stream_f4(g + row * HV + c));
                    dx[row * HV + c] = acc[v];
                }
            }
        }
    }
}

inline int gine_lpr(int HV) {
    int l = 1;
    while (l < HV && l < 32) l <<= 1;
    return l;
}
inline unsigned gine_grid(int64_t N, int lpr) {
    int64_t rows_per_block = (GINE_THREADS / 32) * (32 / lpr);
    int64_t blocks = (N + rows_per_block - 1) / rows_per_block;
    const int64_t cap = (int64_t)GSATB_NUM_SMS * 8 * 4;
    return (unsigned)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace

#define GINE_DISPATCH(KERNEL, ...)                                                             \
    do {                                                                                       \
        const int lpr = gine_lpr(HV);                                                          \
        const int nv = (HV + lpr - 1) / lpr;                                                   \
        const unsigned grid = gine_grid(N, lpr);                                               \
        if (nv == 1) {                                                                         \
            switch (lpr) {                                                                     \
                case 1: KERNEL<1, 1, A><<<grid, GINE_THREADS, 0, st>>>(__VA_ARGS__); break;    \
                case 2: KERNEL<2, 1, A><<<grid, GINE_THREADS, 0, st>>>(__VA_ARGS__); break;    \
                case 4: KERNEL<4, 1, A><<<grid, GINE_THREADS, 0, st>>>(__VA_ARGS__); break;    \
                case 8: KERNEL<8, 1, A><<<grid, GINE_THREADS, 0, st>>>(__VA_ARGS__); break;    \
                case 16: KERNEL<16, 1, A><<<grid, GINE_THREADS, 0, st>>>(__VA_ARGS__); break;  \
                default: KERNEL<32, 1, A><<<grid, GINE_THREADS, 0, st>>>(__VA_ARGS__); break;  \
            }                                                                                  \
        } else if (nv == 2) KERNEL<32, 2, A><<<grid, GINE_THREADS, 0, st>>>(__VA_ARGS__);      \
        else if (nv == 3) KERNEL<32, 3, A><<<grid, GINE_THREADS, 0, st>>>(__VA_ARGS__);        \
        else if (nv == 4) KERNEL<32, 4, A><<<grid, GINE_THREADS, 0, st>>>(__VA_ARGS__);        \
        else return GSATB_ESHAPE;                                                              \
    } while (0)

extern "C" int gsatb_gine_aggregate_fwd(const float* x, const float* edge_feat, const float* att,
                                        const int32_t* rowptr_dst, const int32_t* eid_by_dst, const int32_t* src_by_dst,
                                        float eps, float* out, int64_t N, int64_t E, int H, gsatb_stream_t stream) {
    if (N < 0 || E < 0 || H <= 0) return GSATB_EINVAL;
    if (N == 0) return GSATB_OK;
    if (!x || !out || !rowptr_dst || (E > 0 && (!edge_feat || !src_by_dst || !eid_by_dst))) return GSATB_EINVAL;
    if (H % 4 != 0 || H > 512) return GSATB_ESHAPE;
    if (!gsatb_aligned16(x) || !gsatb_aligned16(out) || (edge_feat && !gsatb_aligned16(edge_feat))) return GSATB_EALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    const int HV = H / 4;
    const float ss = 1.f + eps;
    if (att) {
        constexpr bool A = true;
        GINE_DISPATCH(k_gine_fwd, (const float4*)x, (const float4*)edge_feat, att, rowptr_dst, eid_by_dst, src_by_dst, ss,
                      (float4*)out, N, HV);
    } else {
        constexpr bool A = false;
        GINE_DISPATCH(k_gine_fwd, (const float4*)x, (const float4*)edge_feat, att, rowptr_dst, eid_by_dst, src_by_dst, ss,
                      (float4*)out, N, HV);
    }
    GSATB_CHECK_LAUNCH();
    return GSATB_OK;
}

extern "C" int gsatb_gine_aggregate_bwd(const float* gout, const float* x, const float* edge_feat, const float* att,
                                        const int32_t* rowptr_src, const int32_t* eid_by_src, const int32_t* dst_by_src,
                                        float eps, float* dx, float* dedge_feat, float* datt, int64_t N, int64_t E, int H,
                                        gsatb_stream_t stream) {
    if (N < 0 || E < 0 || H <= 0) return GSATB_EINVAL;
    if (N == 0) return GSATB_OK;
    if (!gout || !x || !dx || !rowptr_src || (E > 0 && (!edge_feat || !dst_by_src || !eid_by_src))) return GSATB_EINVAL;
    if (H % 4 != 0 || H > 512) return GSATB_ESHAPE;
    if (!gsatb_aligned16(gout) || !gsatb_aligned16(x) || !gsatb_aligned16(dx) ||
        (edge_feat && !gsatb_aligned16(edge_feat)) || (dedge_feat && !gsatb_aligned16(dedge_feat)))
        return GSATB_EALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    const int HV = H / 4;
    const float ss = 1.f + eps;
    if (att) {
        constexpr bool A = true;
        GINE_DISPATCH(k_gine_bwd, (const float4*)gout, (const float4*)x, (const float4*)edge_feat, att, rowptr_src,
                      eid_by_src, dst_by_src, ss, (float4*)dx, (float4*)dedge_feat, datt, N, HV);
    } else {
        constexpr bool A = false;
        GINE_DISPATCH(k_gine_bwd, (const float4*)gout, (const float4*)x, (const float4*)edge_feat, att, rowptr_src,
                      eid_by_src, dst_by_src, ss, (float4*)dx, (float4*)dedge_feat, datt, N, HV);
    }
    GSATB_CHECK_LAUNCH();
    return GSATB_OK;
}
