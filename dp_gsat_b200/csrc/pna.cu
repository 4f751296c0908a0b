// K4 -- PNA multi-aggregator message passing (reference src/models/conv_layers.py:160-185, aggregators :193-226):
//   m_e = cat(x[dst(e)], x[src(e)], [edge_feat_e]) * att_e ;  per destination node: sum / mean / min / max / var / std
// One pass over the CSR-by-dst row computes every aggregator at once (the reference runs 6-7 separate
// torch_scatter passes over a materialised [E, F] message tensor, min/max with atomics + an arg pass); rows are
// walked in order, so results are deterministic.  The arg-min / arg-max edge ids are kept for backward, with ties
// broken towards the smallest edge id (torch_scatter's CPU rule; its CUDA path is racy).
//
// Backward is two row-parallel kernels that re-derive d m_e from the saved per-node statistics:
//   by destination (CSR): d x_i part, d edge_feat, d att
//   by source (CSC):      d x_j part
// HBM bound: 4NH + 4E*He + 8E + 4N (in) + 4N*A*F (out)  (SURVEY.md §8d).
#include "common.cuh"

namespace {

constexpr int PNA_THREADS = 256;
enum { AGG_SUM = 0, AGG_MEAN = 1, AGG_MIN = 2, AGG_MAX = 3, AGG_VAR = 4, AGG_STD = 5 };

struct AggList {
    int n;
    int code[8];
};

// message component (float4 column c of the F = 2H [+ He] wide message) of edge p in CSR order
__device__ __forceinline__ float4 msg_feat(const float4* __restrict__ x, const float4* __restrict__ ea, int64_t i,
                                           int src, int eid, int c, int HV, int HeV) {
    if (c < HV) return __ldg(x + i * HV + c);
    if (c < 2 * HV) return __ldg(x + (int64_t)src * HV + (c - HV));
    return __ldg(ea + (int64_t)eid * HeV + (c - 2 * HV));
}

template <int NV>
__global__ void __launch_bounds__(PNA_THREADS)
k_pna_fwd(const float4* __restrict__ x, const float4* __restrict__ ea, const float* __restrict__ att,
          const int32_t* __restrict__ rowptr, const int32_t* __restrict__ eid, const int32_t* __restrict__ nbr,
          AggList aggs, float4* __restrict__ out, float4* __restrict__ stat_mean, float4* __restrict__ stat_msq,
          int4* __restrict__ argmin, int4* __restrict__ argmax, int64_t N, int HV, int HeV) {
    const int FV = 2 * HV + HeV;
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = (blockIdx.x * (int64_t)(PNA_THREADS / 32)) + (threadIdx.x >> 5);
    const int64_t warps_total = (int64_t)gridDim.x * (PNA_THREADS / 32);
    for (int64_t i = warp_global; i < N; i += warps_total) {
        const int beg = __ldg(rowptr + i), end = __ldg(rowptr + i + 1);
        float4 s[NV], q[NV], mn[NV], mx[NV];
        int4 amn[NV], amx[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            s[v] = q[v] = make_float4(0.f, 0.f, 0.f, 0.f);
            mn[v] = make_float4(3.4e38f, 3.4e38f, 3.4e38f, 3.4e38f);
            mx[v] = make_float4(-3.4e38f, -3.4e38f, -3.4e38f, -3.4e38f);
            amn[v] = amx[v] = make_int4(-1, -1, -1, -1);
        }
        for (int p = beg; p < end; ++p) {
            const int e = __ldg(eid + p), j = __ldg(nbr + p);
            const float a = att ? __ldg(att + e) : 1.f;
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const int c = lane + v * 32;
                if (c < FV) {
                    float4 m = msg_feat(x, ea, i, j, e, c, HV, HeV);
                    m.x *= a; m.y *= a; m.z *= a; m.w *= a;
                    s[v].x += m.x; s[v].y += m.y; s[v].z += m.z; s[v].w += m.w;
#define UPD(comp)                                                                                      \
    if (m.comp < mn[v].comp || (m.comp == mn[v].comp && e < amn[v].comp)) { mn[v].comp = m.comp; amn[v].comp = e; } \
    if (m.comp > mx[v].comp || (m.comp == mx[v].comp && e < amx[v].comp)) { mx[v].comp = m.comp; amx[v].comp = e; }
                    UPD(x) UPD(y) UPD(z) UPD(w)
#undef UPD
                }
            }
        }
        const int cnt = end - beg;
        const float inv = 1.f / (float)max(cnt, 1);
        // Variance from the CENTRED values in a second walk over the row (its operands are L1/L2 resident).  The
        // reference evaluates E[m^2] - E[m]^2, which cancels in fp32 and makes the relu gate of `std` flip at random
        // for near-constant neighbourhoods; the centred form is exact there and closer to the fp64 result.
        for (int p = beg; p < end; ++p) {
            const int e = __ldg(eid + p), j = __ldg(nbr + p);
            const float a = att ? __ldg(att + e) : 1.f;
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const int c = lane + v * 32;
                if (c < FV) {
                    const float4 f = msg_feat(x, ea, i, j, e, c, HV, HeV);
                    const float dx_ = f.x * a - s[v].x * inv, dy_ = f.y * a - s[v].y * inv;
                    const float dz_ = f.z * a - s[v].z * inv, dw_ = f.w * a - s[v].w * inv;
                    q[v].x = fmaf(dx_, dx_, q[v].x); q[v].y = fmaf(dy_, dy_, q[v].y);
                    q[v].z = fmaf(dz_, dz_, q[v].z); q[v].w = fmaf(dw_, dw_, q[v].w);
                }
            }
        }
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const int c = lane + v * 32;
            if (c >= FV) continue;
            float4 mean = make_float4(s[v].x * inv, s[v].y * inv, s[v].z * inv, s[v].w * inv);
            float4 var = make_float4(q[v].x * inv, q[v].y * inv, q[v].z * inv, q[v].w * inv);
            float4 msq = var;      // saved statistic = centred variance
            if (cnt == 0) {
                mn[v] = mx[v] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            stat_mean[i * FV + c] = mean;
            stat_msq[i * FV + c] = msq;
            argmin[i * FV + c] = amn[v];
            argmax[i * FV + c] = amx[v];
            for (int k = 0; k < aggs.n; ++k) {
                float4 o;
                switch (aggs.code[k]) {
                    case AGG_SUM: o = s[v]; break;
                    case AGG_MEAN: o = mean; break;
                    case AGG_MIN: o = mn[v]; break;
                    case AGG_MAX: o = mx[v]; break;
                    case AGG_VAR: o = var; break;
                    default:
                        o = make_float4(sqrtf(fmaxf(var.x, 0.f) + 1e-5f), sqrtf(fmaxf(var.y, 0.f) + 1e-5f),
                                        sqrtf(fmaxf(var.z, 0.f) + 1e-5f), sqrtf(fmaxf(var.w, 0.f) + 1e-5f));
                        break;
                }
                out[(i * aggs.n + k) * FV + c] = o;
            }
        }
    }
}

// d m_e[c] for edge e into node i, from the upstream gradient of every aggregator
__device__ __forceinline__ float4 dmsg(const float4* __restrict__ gout, const float4* __restrict__ stat_mean,
                                       const float4* __restrict__ stat_msq, const int4* __restrict__ argmin,
                                       const int4* __restrict__ argmax, const AggList& aggs, int64_t i, int c, int FV,
                                       int cnt, int e, const float4& m) {
    const float inv = 1.f / (float)max(cnt, 1);
    const float4 mean = __ldg(stat_mean + i * FV + c), msq = __ldg(stat_msq + i * FV + c);
    float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < aggs.n; ++k) {
        const float4 g = __ldg(gout + (i * aggs.n + k) * FV + c);
        switch (aggs.code[k]) {
            case AGG_SUM: d.x += g.x; d.y += g.y; d.z += g.z; d.w += g.w; break;
            case AGG_MEAN: d.x += g.x * inv; d.y += g.y * inv; d.z += g.z * inv; d.w += g.w * inv; break;
            case AGG_MIN: {
                const int4 a = __ldg(argmin + i * FV + c);
                d.x += a.x == e ? g.x : 0.f; d.y += a.y == e ? g.y : 0.f;
                d.z += a.z == e ? g.z : 0.f; d.w += a.w == e ? g.w : 0.f;
                break;
            }
            case AGG_MAX: {
                const int4 a = __ldg(argmax + i * FV + c);
                d.x += a.x == e ? g.x : 0.f; d.y += a.y == e ? g.y : 0.f;
                d.z += a.z == e ? g.z : 0.f; d.w += a.w == e ? g.w : 0.f;
                break;
            }
            case AGG_VAR:
                d.x += g.x * 2.f * (m.x - mean.x) * inv; d.y += g.y * 2.f * (m.y - mean.y) * inv;
                d.z += g.z * 2.f * (m.z - mean.z) * inv; d.w += g.w * 2.f * (m.w - mean.w) * inv;
                break;
            default: {   // std = sqrt(relu(var) + 1e-5): zero gradient through the relu where var <= 0
                const float vx = msq.x, vy = msq.y, vz = msq.z, vw = msq.w;   // saved centred variance
                d.x += vx > 0.f ? g.x * (m.x - mean.x) * inv / sqrtf(vx + 1e-5f) : 0.f;
                d.y += vy > 0.f ? g.y * (m.y - mean.y) * inv / sqrtf(vy + 1e-5f) : 0.f;
                d.z += vz > 0.f ? g.z * (m.z - mean.z) * inv / sqrtf(vz + 1e-5f) : 0.f;
                d.w += vw > 0.f ? g.w * (m.w - mean.w) * inv / sqrtf(vw + 1e-5f) : 0.f;
                break;
            }
        }
    }
    return d;
}

// by destination: dx_i[f] (f < H), d edge_feat, d att
template <int NV>
__global__ void __launch_bounds__(PNA_THREADS)
k_pna_bwd_dst(const float4* __restrict__ gout, const float4* __restrict__ x, const float4* __restrict__ ea,
              const float* __restrict__ att, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ eid,
              const int32_t* __restrict__ nbr, AggList aggs, const float4* __restrict__ stat_mean,
              const float4* __restrict__ stat_msq, const int4* __restrict__ argmin, const int4* __restrict__ argmax,
              float4* __restrict__ dx, float4* __restrict__ dea, float* __restrict__ datt, int64_t N, int HV, int HeV) {
    const int FV = 2 * HV + HeV;
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = (blockIdx.x * (int64_t)(PNA_THREADS / 32)) + (threadIdx.x >> 5);
    const int64_t warps_total = (int64_t)gridDim.x * (PNA_THREADS / 32);
    for (int64_t i = warp_global; i < N; i += warps_total) {
        const int beg = __ldg(rowptr + i), end = __ldg(rowptr + i + 1), cnt = end - beg;
        float4 acc[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int p = beg; p < end; ++p) {
            const int e = __ldg(eid + p), j = __ldg(nbr + p);
            const float a = att ? __ldg(att + e) : 1.f;
            float dot = 0.f;
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const int c = lane + v * 32;
                if (c < FV) {
                    const float4 f = msg_feat(x, ea, i, j, e, c, HV, HeV);
                    const float4 m = make_float4(f.x * a, f.y * a, f.z * a, f.w * a);
                    const float4 d = dmsg(gout, stat_mean, stat_msq, argmin, argmax, aggs, i, c, FV, cnt, e, m);
                    dot += dot4(d, f);
                    if (c < HV) fma4(acc[v], a, d);
                    else if (c >= 2 * HV && dea)
                        dea[(int64_t)e * HeV + (c - 2 * HV)] = make_float4(d.x * a, d.y * a, d.z * a, d.w * a);
                }
            }
            if (datt) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
                if (lane == 0) datt[e] = dot;
            }
        }
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const int c = lane + v * 32;
            if (c < HV) dx[i * HV + c] = acc[v];      // x_i part; the x_j part is added by k_pna_bwd_src
        }
    }
}

// by source: dx_j[f] += sum over out-edges e = (j -> i) of att_e * d m_e[H + f]
template <int NV>
__global__ void __launch_bounds__(PNA_THREADS)
k_pna_bwd_src(const float4* __restrict__ gout, const float4* __restrict__ x, const float* __restrict__ att,
              const int32_t* __restrict__ rowptr_src, const int32_t* __restrict__ eid_by_src,
              const int32_t* __restrict__ dst_by_src, const int32_t* __restrict__ rowptr_dst, AggList aggs,
              const float4* __restrict__ stat_mean, const float4* __restrict__ stat_msq,
              const int4* __restrict__ argmin, const int4* __restrict__ argmax, float4* __restrict__ dx, int64_t N,
              int HV, int HeV) {
    const int FV = 2 * HV + HeV;
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = (blockIdx.x * (int64_t)(PNA_THREADS / 32)) + (threadIdx.x >> 5);
    const int64_t warps_total = (int64_t)gridDim.x * (PNA_THREADS / 32);
    for (int64_t j = warp_global; j < N; j += warps_total) {
        const int beg = __ldg(rowptr_src + j), end = __ldg(rowptr_src + j + 1);
        float4 acc[NV], xj[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const int c = lane + v * 32;
            acc[v] = c < HV ? dx[j * HV + c] : make_float4(0.f, 0.f, 0.f, 0.f);
            xj[v] = c < HV ? __ldg(x + j * HV + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int p = beg; p < end; ++p) {
            const int e = __ldg(eid_by_src + p);
            const int64_t i = __ldg(dst_by_src + p);
            const float a = att ? __ldg(att + e) : 1.f;
            const int cnt = __ldg(rowptr_dst + i + 1) - __ldg(rowptr_dst + i);
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const int c = lane + v * 32;
                if (c < HV) {
                    const float4 m = make_float4(xj[v].x * a, xj[v].y * a, xj[v].z * a, xj[v].w * a);
                    const float4 d = dmsg(gout, stat_mean, stat_msq, argmin, argmax, aggs, i, HV + c, FV, cnt, e, m);
                    fma4(acc[v], a, d);
                }
            }
        }
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const int c = lane + v * 32;
            if (c < HV) dx[j * HV + c] = acc[v];
        }
    }
}

inline unsigned pna_grid(int64_t N) {
    int64_t b = (N + PNA_THREADS / 32 - 1) / (PNA_THREADS / 32);
    const int64_t cap = (int64_t)GSATB_NUM_SMS * 32;
    return (unsigned)(b > cap ? cap : (b < 1 ? 1 : b));
}

inline bool make_aggs(const int* codes, int n, AggList& a) {
    if (n < 1 || n > 8) return false;
    a.n = n;
    for (int k = 0; k < n; ++k) {
        if (codes[k] < 0 || codes[k] > 5) return false;
        a.code[k] = codes[k];
    }
    return true;
}

}  // namespace

extern "C" int gsatb_pna_aggregate_fwd(const float* x, const float* edge_feat, const float* att,
                                       const int32_t* rowptr_dst, const int32_t* eid_by_dst, const int32_t* src_by_dst,
                                       const int* agg_codes, int n_aggs, float* out, float* stat_mean, float* stat_msq,
                                       int32_t* argmin, int32_t* argmax, int64_t N, int64_t E, int H, int He,
                                       gsatb_stream_t stream) {
    if (N < 0 || E < 0 || H <= 0 || He < 0) return GSATB_EINVAL;
    if (N == 0) return GSATB_OK;
    if (!x || !rowptr_dst || !out || !stat_mean || !stat_msq || !argmin || !argmax || !agg_codes) return GSATB_EINVAL;
    if (E > 0 && (!eid_by_dst || !src_by_dst)) return GSATB_EINVAL;
    if (He > 0 && E > 0 && !edge_feat) return GSATB_EINVAL;      // (an edgeless batch has an empty, null edge_feat)
    if (H % 4 != 0 || He % 4 != 0) return GSATB_ESHAPE;
    AggList aggs;
    if (!make_aggs(agg_codes, n_aggs, aggs)) return GSATB_EINVAL;
    const int FV = (2 * H + He) / 4;
    const int nv = (FV + 31) / 32;
    cudaStream_t st = (cudaStream_t)stream;
#define PNA_FWD(V)                                                                                                 \
    k_pna_fwd<V><<<pna_grid(N), PNA_THREADS, 0, st>>>((const float4*)x, (const float4*)edge_feat, att, rowptr_dst,     \
                                                      eid_by_dst, src_by_dst, aggs, (float4*)out, (float4*)stat_mean, \
                                                      (float4*)stat_msq, (int4*)argmin, (int4*)argmax, N, H / 4, He / 4)
    if (nv == 1) PNA_FWD(1);
    else if (nv == 2) PNA_FWD(2);
    else if (nv == 3) PNA_FWD(3);
    else if (nv <= 6) PNA_FWD(6);
    else return GSATB_ESHAPE;
#undef PNA_FWD
    GSATB_CHECK_LAUNCH();
    return GSATB_OK;
}

extern "C" int gsatb_pna_aggregate_bwd(const float* gout, const float* x, const float* edge_feat, const float* att,
                                       const int32_t* rowptr_dst, const int32_t* eid_by_dst, const int32_t* src_by_dst,
                                       const int32_t* rowptr_src, const int32_t* eid_by_src, const int32_t* dst_by_src,
                                       const int* agg_codes, int n_aggs, const float* stat_mean, const float* stat_msq,
                                       const int32_t* argmin, const int32_t* argmax, float* dx, float* dedge_feat,
                                       float* datt, int64_t N, int64_t E, int H, int He, gsatb_stream_t stream) {
    if (N < 0 || E < 0 || H <= 0 || He < 0) return GSATB_EINVAL;
    if (N == 0) return GSATB_OK;
    if (!gout || !x || !rowptr_dst || !rowptr_src || !stat_mean || !stat_msq || !argmin || !argmax || !dx ||
        !agg_codes)
        return GSATB_EINVAL;
    if (H % 4 != 0 || He % 4 != 0) return GSATB_ESHAPE;
    AggList aggs;
    if (!make_aggs(agg_codes, n_aggs, aggs)) return GSATB_EINVAL;
    const int FV = (2 * H + He) / 4;
    const int nv = (FV + 31) / 32, nvh = (H / 4 + 31) / 32;
    cudaStream_t st = (cudaStream_t)stream;
#define PNA_BD(V)                                                                                                   \
    k_pna_bwd_dst<V><<<pna_grid(N), PNA_THREADS, 0, st>>>(                                                          \
        (const float4*)gout, (const float4*)x, (const float4*)edge_feat, att, rowptr_dst, eid_by_dst, src_by_dst, aggs, \
        (const float4*)stat_mean, (const float4*)stat_msq, (const int4*)argmin, (const int4*)argmax, (float4*)dx,     \
        (float4*)dedge_feat, datt, N, H / 4, He / 4)
    if (nv == 1) PNA_BD(1);
    else if (nv == 2) PNA_BD(2);
    else if (nv == 3) PNA_BD(3);
    else if (nv <= 6) PNA_BD(6);
    else return GSATB_ESHAPE;
#undef PNA_BD
#define PNA_BS(V)                                                                                                    \
    k_pna_bwd_src<V><<<pna_grid(N), PNA_THREADS, 0, st>>>((const float4*)gout, (const float4*)x, att, rowptr_src,      \
                                                          eid_by_src, dst_by_src, rowptr_dst, aggs,                    \
                                                          (const float4*)stat_mean, (const float4*)stat_msq,           \
                                                          (const int4*)argmin, (const int4*)argmax, (float4*)dx, N,    \
                                                          H / 4, He / 4)
    if (nvh == 1) PNA_BS(1);
    else if (nvh == 2) PNA_BS(2);
    else return GSATB_ESHAPE;
#undef PNA_BS
    GSATB_CHECK_LAUNCH();
    return GSATB_OK;
}
