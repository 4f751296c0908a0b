// Attention-aware LEConv message passing (second half):
//   reference src/models/conv_layers.py:69-92 over PyG LEConv, used by SPMotifNet (src/models/spmotif_gnn.py:58-63):
//     m_e   = ((a[src(e)] - b[dst(e)]) * w[e]) * att[e]            a = lin1(x), b = lin2(x), w = edge_weight (both
//     out[i] = sum_{e: dst(e) = i} m_e  +  add[i]                  optional), add = lin3(x) (computed by the caller)
//   backward:
//     da[j]   = sum_{e: src(e) = j} (g[dst(e)] * att[e]) * w[e]                     (CSC walk, by source)
//     db[i]   = - sum_{e: dst(e) = i} (g[i] * att[e]) * w[e]                       (CSR walk, by destination)
//     datt[e] = <(a[src] - b[dst]) * w[e], g[dst]> ;  dw[e] = <a[src] - b[dst], g[dst] * att[e]> ;  dadd = g
// Row-per-sub-warp CSR / CSC walks in edge order: deterministic, no atomics, products rounded in the reference's
// order ((diff * w) * att; an absent factor is an exact multiplication by 1).  First, untuned version (same shape as
// csrc/gine.cu); HBM bound: fwd 12NH + 12E (+4NH with add), bwd 20NH + 24E.
// Launches go through GSATB_LAUNCH so that tests/simt can run the very same source on the host SIMT emulator.
#include "common.cuh"

namespace {

constexpr int LE_THREADS = 256;

__device__ __forceinline__ float4 sub4(const float4& a, const float4& b) {
    return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
}
__device__ __forceinline__ float4 scale4(const float4& a, float s) {
    return make_float4(a.x * s, a.y * s, a.z * s, a.w * s);
}
__device__ __forceinline__ void add4(float4& acc, const float4& v) {
    acc.x += v.x;
    acc.y += v.y;
    acc.z += v.z;
    acc.w += v.w;
}
__device__ __forceinline__ float sum_prod4(const float4& a, const float4& b) {
    return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
}

// LPR lanes own one row (LPR = 1..32, power of two); each lane holds NV float4 column groups: c = sl + v * LPR.
template <int LPR, int NV>
__global__ void __launch_bounds__(LE_THREADS)
k_le_fwd(const float4* __restrict__ a, const float4* __restrict__ b, const float* __restrict__ w,
         const float* __restrict__ att, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ eid,
         const int32_t* __restrict__ nbr, const float4* __restrict__ add, float4* __restrict__ out, int64_t N, int HV) {
    constexpr int RPW = 32 / LPR;
    const int lane = threadIdx.x & 31, sub = lane / LPR, sl = lane % LPR;
    const int64_t warp_global = (blockIdx.x * (int64_t)(LE_THREADS / 32)) + (threadIdx.x >> 5);
    const int64_t warps_total = (int64_t)gridDim.x * (LE_THREADS / 32);
    for (int64_t row = warp_global * RPW + sub; row < N; row += warps_total * RPW) {
        const int beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
        float4 acc[NV], bi[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const int c = sl + v * LPR;
            acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
            bi[v] = c < HV ? ldg_stream_f4(b + row * HV + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int p = beg; p < end; ++p) {
            const int n = __ldg(nbr + p), e = __ldg(eid + p);
            const float we = w ? __ldg(w + e) : 1.f, ae = att ? __ldg(att + e) : 1.f;
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const int c = sl + v * LPR;
                if (c < HV) add4(acc[v], scale4(scale4(sub4(ldg_f4(a + (int64_t)n * HV + c), bi[v]), we), ae));
            }
        }
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const int c = sl + v * LPR;
            if (c < HV) {
                if (add) add4(acc[v], ldg_stream_f4(add + row * HV + c));
                out[row * HV + c] = acc[v];
            }
        }
    }
}

// by source row j (CSC): da[j], and per outgoing edge the two dot products
template <int LPR, int NV>
__global__ void __launch_bounds__(LE_THREADS)
k_le_bwd_src(const float4* __restrict__ g, const float4* __restrict__ a, const float4* __restrict__ b,
             const float* __restrict__ w, const float* __restrict__ att, const int32_t* __restrict__ rowptr,
             const int32_t* __restrict__ eid, const int32_t* __restrict__ nbr, float4* __restrict__ da,
             float* __restrict__ dw, float* __restrict__ datt, int64_t N, int HV) {
    constexpr int RPW = 32 / LPR;
    const int lane = threadIdx.x & 31, sub = lane / LPR, sl = lane % LPR;
    const unsigned submask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (sub * LPR));
    const int64_t warp_global = (blockIdx.x * (int64_t)(LE_THREADS / 32)) + (threadIdx.x >> 5);
    const int64_t warps_total = (int64_t)gridDim.x * (LE_THREADS / 32);
    for (int64_t row0 = warp_global * RPW; row0 < N; row0 += warps_total * RPW) {
        const int64_t row = row0 + sub;
        const bool live = row < N;
        const int beg = live ? __ldg(rowptr + row) : 0, end = live ? __ldg(rowptr + row + 1) : 0;
        float4 acc[NV], aj[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const int c = sl + v * LPR;
            acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
            aj[v] = (live && c < HV) ? ldg_stream_f4(a + row * HV + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // the shuffle reductions are collective inside a sub-warp: every lane of the sub-warp walks the whole row
        for (int p = beg; p < end; ++p) {
            const int d = __ldg(nbr + p), e = __ldg(eid + p);
            const float we = w ? __ldg(w + e) : 1.f, ae = att ? __ldg(att + e) : 1.f;
            float part_att = 0.f, part_w = 0.f;
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const int c = sl + v * LPR;
                if (c < HV) {
                    const float4 gg = ldg_f4(g + (int64_t)d * HV + c);
                    const float4 diff = sub4(aj[v], ldg_f4(b + (int64_t)d * HV + c));
                    const float4 ga = scale4(gg, ae);
                    add4(acc[v], scale4(ga, we));
                    part_att += sum_prod4(scale4(diff, we), gg);
                    part_w += sum_prod4(diff, ga);
                }
            }
            if (datt || dw) {
#pragma unroll
                for (int o = LPR / 2; o > 0; o >>= 1) {
                    part_att += __shfl_xor_sync(submask, part_att, o);
                    part_w += __shfl_xor_sync(submask, part_w, o);
                }
                if (sl == 0) {
                    if (datt) datt[e] = part_att;
                    if (dw) dw[e] = part_w;
                }
            }
        }
        if (live) {
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const int c = sl + v * LPR;
                if (c < HV) da[row * HV + c] = acc[v];
            }
        }
    }
}

// by destination row i (CSR): db[i] = - sum_e (g[i] * att[e]) * w[e]
template <int LPR, int NV>
__global__ void __launch_bounds__(LE_THREADS)
k_le_bwd_dst(const float4* __restrict__ g, const float* __restrict__ w, const float* __restrict__ att,
             const int32_t* __restrict__ rowptr, const int32_t* __restrict__ eid, float4* __restrict__ db, int64_t N,
             int HV) {
    constexpr int RPW = 32 / LPR;
    const int lane = threadIdx.x & 31, sub = lane / LPR, sl = lane % LPR;
    const int64_t warp_global = (blockIdx.x * (int64_t)(LE_THREADS / 32)) + (threadIdx.x >> 5);
    const int64_t warps_total = (int64_t)gridDim.x * (LE_THREADS / 32);
    for (int64_t row = warp_global * RPW + sub; row < N; row += warps_total * RPW) {
        const int beg = __ldg(rowptr + row), end = __ldg(rowptr + row + 1);
        float4 acc[NV], gi[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const int c = sl + v * LPR;
            acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
            gi[v] = c < HV ? ldg_stream_f4(g + row * HV + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int p = beg; p < end; ++p) {
            const int e = __ldg(eid + p);
            const float we = w ? __ldg(w + e) : 1.f, ae = att ? __ldg(att + e) : 1.f;
#pragma unroll
            for (int v = 0; v < NV; ++v) add4(acc[v], scale4(scale4(gi[v], ae), we));
        }
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const int c = sl + v * LPR;
            if (c < HV) db[row * HV + c] = make_float4(-acc[v].x, -acc[v].y, -acc[v].z, -acc[v].w);
        }
    }
}

inline int le_lpr(int HV) {
    int l = 1;
    while (l < HV && l < 32) l <<= 1;
    return l;
}
inline unsigned le_grid(int64_t N, int lpr) {
    const int64_t rows_per_block = (LE_THREADS / 32) * (32 / lpr);
    const int64_t blocks = (N + rows_per_block - 1) / rows_per_block;
    const int64_t cap = (int64_t)GSATB_NUM_SMS * 8 * 4;
    return (unsigned)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

}  // namespace

// KERNEL<LPR, NV> for the width HV (float4 columns): one row per LPR lanes while HV <= 32, else NV groups per lane
#define LE_DISPATCH(KERNEL, ...)                                                                     \
    do {                                                                                             \
        const int lpr = le_lpr(HV);                                                                  \
        const int nv = (HV + lpr - 1) / lpr;                                                         \
        const unsigned grid = le_grid(N, lpr);                                                       \
        auto k1 = KERNEL<1, 1>;                                                                      \
        auto k2 = KERNEL<2, 1>;                                                                      \
        auto k4 = KERNEL<4, 1>;                                                                      \
        auto k8 = KERNEL<8, 1>;                                                                      \
        auto k16 = KERNEL<16, 1>;                                                                    \
        auto k32 = KERNEL<32, 1>;                                                                    \
        auto k32x2 = KERNEL<32, 2>;                                                                  \
        auto k32x3 = KERNEL<32, 3>;                                                                  \
        auto k32x4 = KERNEL<32, 4>;                                                                  \
        if (nv == 1) {                                                                               \
            switch (lpr) {                                                                           \
                case 1: GSATB_LAUNCH(k1, grid, LE_THREADS, st, __VA_ARGS__); break;                  \
                case 2: GSATB_LAUNCH(k2, grid, LE_THREADS, st, __VA_ARGS__); break;                  \
                case 4: GSATB_LAUNCH(k4, grid, LE_THREADS, st, __VA_ARGS__); break;                  \
                case 8: GSATB_LAUNCH(k8, grid, LE_THREADS, st, __VA_ARGS__); break;                  \
                case 16: GSATB_LAUNCH(k16, grid, LE_THREADS, st, __VA_ARGS__); break;                \
                default: GSATB_LAUNCH(k32, grid, LE_THREADS, st, __VA_ARGS__); break;                \
            }                                                                                        \
        } else if (nv == 2) GSATB_LAUNCH(k32x2, grid, LE_THREADS, st, __VA_ARGS__);                  \
        else if (nv == 3) GSATB_LAUNCH(k32x3, grid, LE_THREADS, st, __VA_ARGS__);                    \
        else if (nv == 4) GSATB_LAUNCH(k32x4, grid, LE_THREADS, st, __VA_ARGS__);                    \
        else return GSATB_ESHAPE;                                                                    \
    } while (0)

extern "C" int gsatb_le_aggregate_fwd(const float* a, const float* b, const float* edge_weight, const float* att,
                                      const int32_t* rowptr_dst, const int32_t* eid_by_dst, const int32_t* src_by_dst,
                                      const float* add, float* out, int64_t N, int64_t E, int H,
                                      gsatb_stream_t stream) {
    if (N < 0 || E < 0 || H <= 0) return GSATB_EINVAL;
    if (N == 0) return GSATB_OK;
    if (!a || !b || !out || !rowptr_dst || (E > 0 && (!src_by_dst || !eid_by_dst))) return GSATB_EINVAL;
    if (H % 4 != 0 || H > 512) return GSATB_ESHAPE;
    if (!gsatb_aligned16(a) || !gsatb_aligned16(b) || !gsatb_aligned16(out) || (add && !gsatb_aligned16(add)))
        return GSATB_EALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    const int HV = H / 4;
    LE_DISPATCH(k_le_fwd, (const float4*)a, (const float4*)b, edge_weight, att, rowptr_dst, eid_by_dst, src_by_dst,
                (const float4*)add, (float4*)out, N, HV);
    GSATB_CHECK_LAUNCH();
    return GSATB_OK;
}

extern "C" int gsatb_le_aggregate_bwd(const float* gout, const float* a, const float* b, const float* edge_weight,
                                      const float* att, const int32_t* rowptr_src, const int32_t* eid_by_src,
                                      const int32_t* dst_by_src, const int32_t* rowptr_dst, const int32_t* eid_by_dst,
                                      float* da, float* db, float* dedge_weight, float* datt, int64_t N, int64_t E,
                                      int H, gsatb_stream_t stream) {
    if (N < 0 || E < 0 || H <= 0) return GSATB_EINVAL;
    if (N == 0) return GSATB_OK;
    if (!gout || !a || !b || !da || !db || !rowptr_src || !rowptr_dst ||
        (E > 0 && (!dst_by_src || !eid_by_src || !eid_by_dst)))
        return GSATB_EINVAL;
    if (H % 4 != 0 || H > 512) return GSATB_ESHAPE;
    if (!gsatb_aligned16(gout) || !gsatb_aligned16(a) || !gsatb_aligned16(b) || !gsatb_aligned16(da) ||
        !gsatb_aligned16(db))
        return GSATB_EALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    const int HV = H / 4;
    LE_DISPATCH(k_le_bwd_src, (const float4*)gout, (const float4*)a, (const float4*)b, edge_weight, att, rowptr_src,
                eid_by_src, dst_by_src, (float4*)da, dedge_weight, datt, N, HV);
    LE_DISPATCH(k_le_bwd_dst, (const float4*)gout, edge_weight, att, rowptr_dst, eid_by_dst, (float4*)db, N, HV);
    GSATB_CHECK_LAUNCH();
    return GSATB_OK;
}
