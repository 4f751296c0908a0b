// K3 -- attention-weighted GIN aggregation (gather - scale - segmented sum over CSR), forward and backward,
// and K5 graph readout.
//
// Replaces reference src/models/conv_layers.py:14-34 + PyG MessagePassing.propagate (index_select of x_j,
// x_j * edge_atten, torch_scatter scatter-add with atomics, out += (1+eps) x) and its autograd backward
// (index_add_ atomics).  Rows are walked in CSR order so sums are sequential and run-to-run deterministic.
//
// Mapping: a sub-warp of LPR lanes owns one destination row; each lane carries NV float4 accumulators, so a
// row of H floats is covered by LPR*NV 128-bit columns.  H=64 -> 16 lanes/row, 2 rows per warp; H=128 -> a whole
// warp per row; H=300 -> 75 float4 = 32 lanes x 3.  Index / attention slices of a row are fetched by the whole
// sub-warp in one coalesced load and broadcast with shuffles, then the neighbour rows are gathered with
// independent 128-bit loads (4 in flight per lane) to cover HBM/L2 latency at the low average degree (~2) of the
// GSAT datasets.
//
// HBM bound.  Algorithmic bytes / launch: fwd 8NH + 8E + 4N, bwd 12NH + 16E.
#include "common.cuh"

namespace {

constexpr int AGG_THREADS = 256;

template <int LPR, int NV, bool HAS_ATT>
__global__ void __launch_bounds__(AGG_THREADS)
k_gin_aggregate_fwd(const float4* __restrict__ x, const float* __restrict__ att, const int32_t* __restrict__ rowptr,
                    const int32_t* __restrict__ eid, const int32_t* __restrict__ nbr, float self_scale,
                    float4* __restrict__ out, int64_t N, int HV) {
    constexpr int ROWS_PER_WARP = 32 / LPR;
    const int lane = threadIdx.x & 31;
    const int sub = lane / LPR;            // which row of the warp
    const int sl = lane % LPR;             // lane inside the row
    const unsigned submask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (sub * LPR));
    const int64_t warp_global = (blockIdx.x * (int64_t)(AGG_THREADS / 32)) + (threadIdx.x >> 5);
    const int64_t warps_total = (int64_t)gridDim.x * (AGG_THREADS / 32);

    for (int64_t row0 = warp_global * ROWS_PER_WARP; row0 < N; row0 += warps_total * ROWS_PER_WARP) {
        const int64_t row = row0 + sub;
        const bool live = row < N;
        int beg = 0, end = 0;
        if (live) {
            beg = __ldg(rowptr + row);
            end = __ldg(rowptr + row + 1);
        }
        float4 acc[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);

        for (int p0 = beg; p0 < end; p0 += LPR) {
            // one coalesced fetch of up to LPR (neighbour, attention) pairs for this row
            int my_n = 0;
            float my_a = 1.f;
            if (p0 + sl < end) {
                my_n = __ldg(nbr + p0 + sl);
                if (HAS_ATT) my_a = __ldg(att + __ldg(eid + p0 + sl));
            }
            const int cnt = min(LPR, end - p0);
            for (int q0 = 0; q0 < cnt; q0 += 4) {
                int n[4];
                float a[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    n[u] = __shfl_sync(submask, my_n, sub * LPR + ((q0 + u) % LPR));
                    a[u] = __shfl_sync(submask, my_a, sub * LPR + ((q0 + u) % LPR));
                }
                float4 g[4][NV];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (q0 + u < cnt) {
#pragma unroll
                        for (int v = 0; v < NV; ++v) {
                            const int c = sl + v * LPR;
                            if (c < HV) g[u][v] = ldg_f4(x + (int64_t)n[u] * HV + c);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (q0 + u < cnt) {
#pragma unroll
                        for (int v = 0; v < NV; ++v) {
                            const int c = sl + v * LPR;
                            if (c < HV) {
                                if (HAS_ATT) fma4(acc[v], a[u], g[u][v]);
                                else {
                                    acc[v].x += g[u][v].x;
                                    acc[v].y += g[u][v].y;
                                    acc[v].z += g[u][v].z;
                                    acc[v].w += g[u][v].w;
                                }
                            }
                        }
                    }
                }
            }
        }
        if (live) {
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const int c = sl + v * LPR;
                if (c < HV) {
                    float4 xs = ldg_stream_f4(x + row * HV + c);
                    float4 o;
                    o.x = acc[v].x + self_scale * xs.x;   // reference: out = scatter(...); out += (1+eps) * x
                    o.y = acc[v].y + self_scale * xs.y;
                    o.z = acc[v].z + self_scale * xs.z;
                    o.w = acc[v].w + self_scale * xs.w;
                    out[row * HV + c] = o;
                }
            }
        }
    }
}

// backward over CSC rows (edges grouped by source j): dx[j] = sum att_e g[dst_e] + (1+eps) g[j];
// datt[e] = <x[j], g[dst_e]>
template <int LPR, int NV, bool HAS_ATT, bool WANT_DATT>
__global__ void __launch_bounds__(AGG_THREADS)
k_gin_aggregate_bwd(const float4* __restrict__ g, const float4* __restrict__ x, const float* __restrict__ att,
                    const int32_t* __restrict__ rowptr, const int32_t* __restrict__ eid,
                    const int32_t* __restrict__ nbr, float self_scale, float4* __restrict__ dx,
                    float* __restrict__ datt, int64_t N, int HV) {
    constexpr int ROWS_PER_WARP = 32 / LPR;
    const int lane = threadIdx.x & 31;
    const int sub = lane / LPR;
    const int sl = lane % LPR;
    const unsigned submask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (sub * LPR));
    const int64_t warp_global = (blockIdx.x * (int64_t)(AGG_THREADS / 32)) + (threadIdx.x >> 5);
    const int64_t warps_total = (int64_t)gridDim.x * (AGG_THREADS / 32);

    for (int64_t row0 = warp_global * ROWS_PER_WARP; row0 < N; row0 += warps_total * ROWS_PER_WARP) {
        const int64_t row = row0 + sub;
        const bool live = row < N;
        int beg = 0, end = 0;
        if (live) {
            beg = __ldg(rowptr + row);
            end = __ldg(rowptr + row + 1);
        }
        float4 acc[NV], xr[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
            xr[v] = make_float4(0.f, 0.f, 0.f, 0.f);
            const int c = sl + v * LPR;
            if (WANT_DATT && live && c < HV && end > beg) xr[v] = ldg_stream_f4(x + row * HV + c);
        }
        for (int p0 = beg; p0 < end; p0 += LPR) {
            int my_n = 0, my_e = 0;
            float my_a = 1.f;
            if (p0 + sl < end) {
                my_n = __ldg(nbr + p0 + sl);
                if (HAS_ATT || WANT_DATT) my_e = __ldg(eid + p0 + sl);
                if (HAS_ATT) my_a = __ldg(att + my_e);
            }
            const int cnt = min(LPR, end - p0);
            float my_dot = 0.f;   // lane q keeps the dot product of edge p0+q
            for (int q0 = 0; q0 < cnt; q0 += 4) {
                int n[4];
                float a[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    n[u] = __shfl_sync(submask, my_n, sub * LPR + ((q0 + u) % LPR));
                    a[u] = __shfl_sync(submask, my_a, sub * LPR + ((q0 + u) % LPR));
                }
                float4 gg[4][NV];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (q0 + u < cnt) {
#pragma unroll
                        for (int v = 0; v < NV; ++v) {
                            const int c = sl + v * LPR;
                            if (c < HV) gg[u][v] = ldg_f4(g + (int64_t)n[u] * HV + c);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float part = 0.f;
                    if (q0 + u < cnt) {
#pragma unroll
                        for (int v = 0; v < NV; ++v) {
                            const int c = sl + v * LPR;
                            if (c < HV) {
                                fma4(acc[v], a[u], gg[u][v]);
                                if (WANT_DATT) part += dot4(xr[v], gg[u][v]);
                            }
                        }
                    }
                    if (WANT_DATT) {
#pragma unroll
                        for (int o = LPR / 2; o > 0; o >>= 1) part += __shfl_xor_sync(submask, part, o);
                        if (sl == ((q0 + u) % LPR)) my_dot = part;
                    }
                }
            }
            if (WANT_DATT && p0 + sl < end) datt[my_e] = my_dot;
        }
        if (live) {
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const int c = sl + v * LPR;
                if (c < HV) {
                    float4 gs = ldg_stream_f4(g + row * HV + c);
                    float4 o;
                    o.x = acc[v].x + self_scale * gs.x;
                    o.y = acc[v].y + self_scale * gs.y;
                    o.z = acc[v].z + self_scale * gs.z;
                    o.w = acc[v].w + self_scale * gs.w;
                    dx[row * HV + c] = o;
                }
            }
        }
    }
}

inline int pick_lpr(int HV) {
    int l = 1;
    while (l < HV && l < 32) l <<= 1;
    return l;
}

inline unsigned agg_grid(int64_t N, int lpr) {
    int64_t rows_per_block = (AGG_THREADS / 32) * (32 / lpr);
    int64_t blocks = (N + rows_per_block - 1) / rows_per_block;
    int64_t cap = (int64_t)GSATB_NUM_SMS * 8 * 4;   // 8 resident CTAs/SM, 4 waves max, then grid-stride
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

template <bool HAS_ATT>
int launch_fwd(const float* x, const float* att, const int32_t* rowptr, const int32_t* eid, const int32_t* nbr,
               float self_scale, float* out, int64_t N, int HV, cudaStream_t st) {
    const int lpr = pick_lpr(HV);
    const int nv = (HV + lpr - 1) / lpr;
    const unsigned grid = agg_grid(N, lpr);
#define FWD_CASE(L, V)                                                                                       \
    k_gin_aggregate_fwd<L, V, HAS_ATT><<<grid, AGG_THREADS, 0, st>>>((const float4*)x, att, rowptr, eid, nbr, \
                                                                     self_scale, (float4*)out, N, HV)
    if (nv == 1) {
        switch (lpr) {
            case 1: FWD_CASE(1, 1); break;
            case 2: FWD_CASE(2, 1); break;
            case 4: FWD_CASE(4, 1); break;
            case 8: FWD_CASE(8, 1); break;
            case 16: FWD_CASE(16, 1); break;
            default: FWD_CASE(32, 1); break;
        }
    } else if (nv == 2) FWD_CASE(32, 2);
    else if (nv == 3) FWD_CASE(32, 3);
    else if (nv == 4) FWD_CASE(32, 4);
    else return GSATB_ESHAPE;
#undef FWD_CASE
    return GSATB_OK;
}

template <bool HAS_ATT, bool WANT_DATT>
int launch_bwd(const float* g, const float* x, const float* att, const int32_t* rowptr, const int32_t* eid,
               const int32_t* nbr, float self_scale, float* dx, float* datt, int64_t N, int HV, cudaStream_t st) {
    const int lpr = pick_lpr(HV);
    const int nv = (HV + lpr - 1) / lpr;
    const unsigned grid = agg_grid(N, lpr);
#define BWD_CASE(L, V)                                                                                  \
    k_gin_aggregate_bwd<L, V, HAS_ATT, WANT_DATT><<<grid, AGG_THREADS, 0, st>>>(                         \
        (const float4*)g, (const float4*)x, att, rowptr, eid, nbr, self_scale, (float4*)dx, datt, N, HV)
    if (nv == 1) {
        switch (lpr) {
            case 1: BWD_CASE(1, 1); break;
            case 2: BWD_CASE(2, 1); break;
            case 4: BWD_CASE(4, 1); break;
            case 8: BWD_CASE(8, 1); break;
            case 16: BWD_CASE(16, 1); break;
            default: BWD_CASE(32, 1); break;
        }
    } else if (nv == 2) BWD_CASE(32, 2);
    else if (nv == 3) BWD_CASE(32, 3);
    else if (nv == 4) BWD_CASE(32, 4);
    else return GSATB_ESHAPE;
#undef BWD_CASE
    return GSATB_OK;
}

// ---- K5 readout ------------------------------------------------------------------------------------------
template <int LPR>
__global__ void __launch_bounds__(AGG_THREADS)
k_pool_fwd(const float4* __restrict__ x, const int32_t* __restrict__ node_ptr, float4* __restrict__ out, int64_t G,
           int HV, int mean) {
    constexpr int ROWS_PER_WARP = 32 / LPR;
    const int lane = threadIdx.x & 31, sub = lane / LPR, sl = lane % LPR;
    const int64_t warp_global = (blockIdx.x * (int64_t)(AGG_THREADS / 32)) + (threadIdx.x >> 5);
    const int64_t warps_total = (int64_t)gridDim.x * (AGG_THREADS / 32);
    for (int64_t g0 = warp_global * ROWS_PER_WARP; g0 < G; g0 += warps_total * ROWS_PER_WARP) {
        const int64_t gi = g0 + sub;
        if (gi >= G) continue;
        const int beg = __ldg(node_ptr + gi), end = __ldg(node_ptr + gi + 1);
        const float inv = (mean && end > beg) ? 1.f / (float)(end - beg) : 1.f;
        for (int c = sl; c < HV; c += LPR) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            int i = beg;
            for (; i + 4 <= end; i += 4) {
                float4 a0 = ldg_stream_f4(x + (int64_t)(i + 0) * HV + c);
                float4 a1 = ldg_stream_f4(x + (int64_t)(i + 1) * HV + c);
                float4 a2 = ldg_stream_f4(x + (int64_t)(i + 2) * HV + c);
                float4 a3 = ldg_stream_f4(x + (int64_t)(i + 3) * HV + c);
                acc.x += a0.x; acc.y += a0.y; acc.z += a0.z; acc.w += a0.w;
                acc.x += a1.x; acc.y += a1.y; acc.z += a1.z; acc.w += a1.w;
                acc.x += a2.x; acc.y += a2.y; acc.z += a2.z; acc.w += a2.w;
                acc.x += a3.x; acc.y += a3.y; acc.z += a3.z; acc.w += a3.w;
            }
            for (; i < end; ++i) {
                float4 a0 = ldg_stream_f4(x + (int64_t)i * HV + c);
                acc.x += a0.x; acc.y += a0.y; acc.z += a0.z; acc.w += a0.w;
            }
            if (mean) { acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv; }
            out[gi * HV + c] = acc;
        }
    }
}

__global__ void k_pool_bwd(const float4* __restrict__ gout, const int32_t* __restrict__ node_ptr,
                           const int32_t* __restrict__ node_graph, float4* __restrict__ dx, int64_t N, int HV,
                           int mean) {
    const int64_t total = N * HV;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / HV;
        const int c = (int)(i - row * HV);
        const int g = __ldg(node_graph + row);
        float4 v = __ldg(gout + (int64_t)g * HV + c);
        if (mean) {
            const float inv = 1.f / (float)max(1, __ldg(node_ptr + g + 1) - __ldg(node_ptr + g));
            v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
        }
        dx[i] = v;
    }
}

}  // namespace

extern "C" int gsatb_gin_aggregate_fwd(const float* x, const float* att, const int32_t* rowptr_dst,
                                       const int32_t* eid_by_dst, const int32_t* src_by_dst, float eps, float* out,
                                       int64_t N, int64_t E, int H, gsatb_stream_t stream) {
    if (N < 0 || E < 0 || H <= 0) return GSATB_EINVAL;
    if (N == 0) return GSATB_OK;
    if (!x || !out || !rowptr_dst || (E > 0 && (!src_by_dst || (att && !eid_by_dst)))) return GSATB_EINVAL;
    if (H % 4 != 0 || H > 512) return GSATB_ESHAPE;
    if (!gsatb_aligned16(x) || !gsatb_aligned16(out)) return GSATB_EALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = att ? launch_fwd<true>(x, att, rowptr_dst, eid_by_dst, src_by_dst, 1.f + eps, out, N, H / 4, st)
                 : launch_fwd<false>(x, att, rowptr_dst, eid_by_dst, src_by_dst, 1.f + eps, out, N, H / 4, st);
    if (rc != GSATB_OK) return rc;
    GSATB_CHECK_LAUNCH();
    return GSATB_OK;
}

extern "C" int gsatb_gin_aggregate_bwd(const float* gout, const float* x, const float* att,
                                       const int32_t* rowptr_src, const int32_t* eid_by_src,
                                       const int32_t* dst_by_src, float eps, float* dx, float* datt, int64_t N,
                                       int64_t E, int H, gsatb_stream_t stream) {
    if (N < 0 || E < 0 || H <= 0) return GSATB_EINVAL;
    if (N == 0) return GSATB_OK;
    if (!gout || !dx || !rowptr_src || (E > 0 && !dst_by_src)) return GSATB_EINVAL;
    if (datt && (!x || !eid_by_src)) return GSATB_EINVAL;
    if (att && !eid_by_src) return GSATB_EINVAL;
    if (H % 4 != 0 || H > 512) return GSATB_ESHAPE;
    if (!gsatb_aligned16(gout) || !gsatb_aligned16(dx) || (x && !gsatb_aligned16(x))) return GSATB_EALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    const float ss = 1.f + eps;
    const int HV = H / 4;
    int rc;
    if (att && datt) rc = launch_bwd<true, true>(gout, x, att, rowptr_src, eid_by_src, dst_by_src, ss, dx, datt, N, HV, st);
    else if (att) rc = launch_bwd<true, false>(gout, x, att, rowptr_src, eid_by_src, dst_by_src, ss, dx, datt, N, HV, st);
    else if (datt) rc = launch_bwd<false, true>(gout, x, att, rowptr_src, eid_by_src, dst_by_src, ss, dx, datt, N, HV, st);
    else rc = launch_bwd<false, false>(gout, x, att, rowptr_src, eid_by_src, dst_by_src, ss, dx, datt, N, HV, st);
    if (rc != GSATB_OK) return rc;
    GSATB_CHECK_LAUNCH();
    return GSATB_OK;
}

extern "C" int gsatb_pool_fwd(const float* x, const int32_t* node_ptr, float* out, int64_t N, int64_t G, int H,
                              int mean, gsatb_stream_t stream) {
    if (N < 0 || G < 0 || H <= 0) return GSATB_EINVAL;
    if (G == 0) return GSATB_OK;
    if (!out || !node_ptr || (N > 0 && !x)) return GSATB_EINVAL;
    if (H % 4 != 0) return GSATB_ESHAPE;
    if (!gsatb_aligned16(x) || !gsatb_aligned16(out)) return GSATB_EALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    const int HV = H / 4;
    const int lpr = pick_lpr(HV);
    int64_t rows_per_block = (AGG_THREADS / 32) * (32 / lpr);
    int64_t blocks = (G + rows_per_block - 1) / rows_per_block;
    if (blocks > (int64_t)GSATB_NUM_SMS * 32) blocks = (int64_t)GSATB_NUM_SMS * 32;
    switch (lpr) {
        case 1: k_pool_fwd<1><<<(unsigned)blocks, AGG_THREADS, 0, st>>>((const float4*)x, node_ptr, (float4*)out, G, HV, mean); break;
        case 2: k_pool_fwd<2><<<(unsigned)blocks, AGG_THREADS, 0, st>>>((const float4*)x, node_ptr, (float4*)out, G, HV, mean); break;
        case 4: k_pool_fwd<4><<<(unsigned)blocks, AGG_THREADS, 0, st>>>((const float4*)x, node_ptr, (float4*)out, G, HV, mean); break;
        case 8: k_pool_fwd<8><<<(unsigned)blocks, AGG_THREADS, 0, st>>>((const float4*)x, node_ptr, (float4*)out, G, HV, mean); break;
        case 16: k_pool_fwd<16><<<(unsigned)blocks, AGG_THREADS, 0, st>>>((const float4*)x, node_ptr, (float4*)out, G, HV, mean); break;
        default: k_pool_fwd<32><<<(unsigned)blocks, AGG_THREADS, 0, st>>>((const float4*)x, node_ptr, (float4*)out, G, HV, mean); break;
    }
    GSATB_CHECK_LAUNCH();
    return GSATB_OK;
}

extern "C" int gsatb_pool_bwd(const float* gout, const int32_t* node_ptr, const int32_t* node_graph, float* dx,
                              int64_t N, int64_t G, int H, int mean, gsatb_stream_t stream) {
    if (N < 0 || G < 0 || H <= 0) return GSATB_EINVAL;
    if (N == 0) return GSATB_OK;
    if (!gout || !dx || !node_graph || !node_ptr) return GSATB_EINVAL;
    if (H % 4 != 0) return GSATB_ESHAPE;
    if (!gsatb_aligned16(gout) || !gsatb_aligned16(dx)) return GSATB_EALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t total = N * (H / 4);
    int64_t blocks = (total + 255) / 256;
    if (blocks > (int64_t)GSATB_NUM_SMS * 32) blocks = (int64_t)GSATB_NUM_SMS * 32;
    k_pool_bwd<<<(unsigned)blocks, 256, 0, st>>>((const float4*)gout, node_ptr, node_graph, (float4*)dx, N, H / 4, mean);
    GSATB_CHECK_LAUNCH();
    return GSATB_OK;
}

// ---- edge feature gather for the extractor (reference src/run_gsat.py:912-914: cat(emb[col], emb[row])) ------
namespace {

__global__ void k_gather_concat_fwd(const float4* __restrict__ emb, const int32_t* __restrict__ src,
                                    const int32_t* __restrict__ dst, float4* __restrict__ out, int64_t E, int HV) {
    const int64_t total = E * 2 * HV;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = i / (2 * HV);
        const int c = (int)(i - e * 2 * HV);
        const int node = c < HV ? __ldg(src + e) : __ldg(dst + e);
        out[i] = ldg_f4(emb + (int64_t)node * HV + (c < HV ? c : c - HV));
    }
}

// demb[j] = sum_{e: src(e)=j} g[e, 0:H] + sum_{e: dst(e)=j} g[e, H:2H]   (CSC rows then CSR rows, sequential)
template <int LPR>
__global__ void __launch_bounds__(AGG_THREADS)
k_gather_concat_bwd(const float4* __restrict__ g, const int32_t* __restrict__ rowptr_src,
                    const int32_t* __restrict__ eid_by_src, const int32_t* __restrict__ rowptr_dst,
                    const int32_t* __restrict__ eid_by_dst, float4* __restrict__ demb, int64_t N, int HV) {
    constexpr int ROWS_PER_WARP = 32 / LPR;
    const int lane = threadIdx.x & 31, sub = lane / LPR, sl = lane % LPR;
    const int64_t warp_global = (blockIdx.x * (int64_t)(AGG_THREADS / 32)) + (threadIdx.x >> 5);
    const int64_t warps_total = (int64_t)gridDim.x * (AGG_THREADS / 32);
    for (int64_t row0 = warp_global * ROWS_PER_WARP; row0 < N; row0 += warps_total * ROWS_PER_WARP) {
        const int64_t row = row0 + sub;
        if (row >= N) continue;
        for (int c = sl; c < HV; c += LPR) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int p = __ldg(rowptr_src + row), pe = __ldg(rowptr_src + row + 1); p < pe; ++p) {
                float4 v = ldg_stream_f4(g + (int64_t)__ldg(eid_by_src + p) * 2 * HV + c);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
            for (int p = __ldg(rowptr_dst + row), pe = __ldg(rowptr_dst + row + 1); p < pe; ++p) {
                float4 v = ldg_stream_f4(g + (int64_t)__ldg(eid_by_dst + p) * 2 * HV + HV + c);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
            demb[row * HV + c] = acc;
        }
    }
}

}  // namespace

extern "C" int gsatb_gather_concat_fwd(const float* emb, const int32_t* src, const int32_t* dst, float* out,
                                       int64_t E, int H, gsatb_stream_t stream) {
    if (E < 0 || H <= 0) return GSATB_EINVAL;
    if (E == 0) return GSATB_OK;
    if (!emb || !src || !dst || !out) return GSATB_EINVAL;
    if (H % 4 != 0) return GSATB_ESHAPE;
    if (!gsatb_aligned16(emb) || !gsatb_aligned16(out)) return GSATB_EALIGN;
    int64_t total = E * 2 * (H / 4);
    int64_t blocks = (total + 255) / 256;
    if (blocks > (int64_t)GSATB_NUM_SMS * 32) blocks = (int64_t)GSATB_NUM_SMS * 32;
    k_gather_concat_fwd<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const float4*)emb, src, dst, (float4*)out, E, H / 4);
    GSATB_CHECK_LAUNCH();
    return GSATB_OK;
}

extern "C" int gsatb_gather_concat_bwd(const float* g, const int32_t* rowptr_src, const int32_t* eid_by_src,
                                       const int32_t* rowptr_dst, const int32_t* eid_by_dst, float* demb, int64_t N,
                                       int H, gsatb_stream_t stream) {
    if (N < 0 || H <= 0) return GSATB_EINVAL;
    if (N == 0) return GSATB_OK;
    if (!g || !rowptr_src || !rowptr_dst || !demb) return GSATB_EINVAL;
    if (H % 4 != 0) return GSATB_ESHAPE;
    if (!gsatb_aligned16(g) || !gsatb_aligned16(demb)) return GSATB_EALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    const int HV = H / 4;
    const int lpr = pick_lpr(HV);
    const unsigned grid = agg_grid(N, lpr);
#define GC_CASE(L) k_gather_concat_bwd<L><<<grid, AGG_THREADS, 0, st>>>((const float4*)g, rowptr_src, eid_by_src, rowptr_dst, eid_by_dst, (float4*)demb, N, HV)
    switch (lpr) {
        case 1: GC_CASE(1); break;
        case 2: GC_CASE(2); break;
        case 4: GC_CASE(4); break;
        case 8: GC_CASE(8); break;
        case 16: GC_CASE(16); break;
        default: GC_CASE(32); break;
    }
#undef GC_CASE
    GSATB_CHECK_LAUNCH();
    return GSATB_OK;
}
