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
#include <cuda_bf16.h>
#include <cstdlib>
#include "common.cuh"

namespace {

constexpr int AGG_THREADS = 256;
constexpr int AGG_WORKERS = 8;                        // gather warps of a K3 CTA
constexpr int K3_THREADS = (AGG_WORKERS + 1) * 32;    // + one stager warp
constexpr int AGG_TILE = 64;      // destination rows per CTA tile
constexpr int AGG_EMAX = 1024;    // edges of a tile staged in shared memory (denser tiles take the slow path)

// Both K3 kernels are persistent: a CTA walks tiles of AGG_TILE consecutive CSR rows.
//
// Warp 8 (the "stager") runs ONE TILE AHEAD.  It fetches the next tile's row pointers and its (neighbour, edge id,
// attention) slices -- a chain of three dependent but fully coalesced, batched global loads -- and writes them to
// the other half of a double buffer in shared memory as a FLAT ENTRY LIST: the edges of each row in CSR order,
// followed by one "self" entry (neighbour = the row itself, weight = 1 + eps).  So the GIN self term is just the
// last term of the row's sum (same order and same fused multiply-add as "out = scatter(...); out += (1+eps)*x").
//
// The eight worker warps then walk contiguous slices of that list: U entries are gathered back to back (U x 512 B
// in flight per warp at H = 128) and consumed in order; a row boundary is one compare + one 128-bit store.  There
// are no per-row predicated slots, no dependent index loads and ~40 instructions per row on the critical path (the
// first version chased rowptr -> nbr/eid -> att -> x serially per warp: 46 % / 51 % of the measured HBM peak,
// latency bound; a row-per-sub-warp rewrite with 4 predicated slots per row was issue bound at ~250 instructions
// per row -- ncu, profiles/r1_k3_*.txt).
struct AggTile {
    int rp[AGG_TILE + 1];              // raw row pointers of the tile
    int staged;
    int nbr[AGG_EMAX + AGG_TILE];      // entry list: neighbour row
    int eid[AGG_EMAX + AGG_TILE];      //             edge id (-1 for the self entry)
    int row[AGG_EMAX + AGG_TILE];      //             global row of the entry (backward: x[row] for d att)
    float att[AGG_EMAX + AGG_TILE];    //             weight
};

// executed by ONE warp
template <bool HAS_ATT, bool NEED_EID>
__device__ __forceinline__ void agg_stage(AggTile& t, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ eid,
                                          const int32_t* __restrict__ nbr, const float* __restrict__ att, int64_t t0,
                                          int nr, float self_scale, int lane) {
    constexpr int RP_PER_LANE = (AGG_TILE + 1 + 31) / 32;
    int rp[RP_PER_LANE];
#pragma unroll
    for (int k = 0; k < RP_PER_LANE; ++k) {
        const int i = lane + k * 32;
        rp[k] = i <= nr ? __ldg(rowptr + t0 + i) : 0;
    }
#pragma unroll
    for (int k = 0; k < RP_PER_LANE; ++k) {
        const int i = lane + k * 32;
        if (i <= nr) t.rp[i] = rp[k];
    }
    __syncwarp();
    const int e0 = t.rp[0], ne = t.rp[nr] - e0;
    const bool staged = ne <= AGG_EMAX;
    if (lane == 0) t.staged = staged ? 1 : 0;
    if (!staged) return;
    for (int r = lane; r < nr; r += 32) {             // self entries close each row
        const int slot = t.rp[r + 1] - e0 + r;
        t.nbr[slot] = (int)(t0 + r);
        t.att[slot] = self_scale;
        if (NEED_EID) {
            t.eid[slot] = -1;
            t.row[slot] = (int)(t0 + r);
        }
    }
    constexpr int U = 8;
    for (int base = 0; base < ne; base += 32 * U) {
        int nb[U], ed[U];
        float av[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const int i = base + lane + k * 32;
            const bool ok = i < ne;
            nb[k] = ok ? __ldg(nbr + e0 + i) : 0;
            ed[k] = (ok && (HAS_ATT || NEED_EID)) ? __ldg(eid + e0 + i) : 0;
        }
#pragma unroll
        for (int k = 0; k < U; ++k) av[k] = (HAS_ATT && base + lane + k * 32 < ne) ? __ldg(att + ed[k]) : 1.f;
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const int i = base + lane + k * 32;
            if (i < ne) {
                int lo = 0, hi = nr;                  // row of edge i: largest r with rp[r] <= e0 + i
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (t.rp[mid] <= e0 + i) lo = mid;
                    else hi = mid;
                }
                const int slot = i + lo;
                t.nbr[slot] = nb[k];
                if (NEED_EID) {
                    t.eid[slot] = ed[k];
                    t.row[slot] = (int)(t0 + lo);
                }
                t.att[slot] = av[k];
            }
        }
    }
}

// slow path for one row of a tile that is too dense to stage (reads its slices from global memory)
template <int LPR, int NV, bool HAS_ATT>
__device__ __forceinline__ void agg_row_slow(const float4* __restrict__ x, const float* __restrict__ att,
                                             const int32_t* __restrict__ eid, const int32_t* __restrict__ nbr, int beg,
                                             int end, int sl, int HV, float4 (&acc)[NV]) {
    for (int p = beg; p < end; ++p) {
        const int n = __ldg(nbr + p);
        const float a = HAS_ATT ? __ldg(att + __ldg(eid + p)) : 1.f;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const int c = sl + v * LPR;
            if (c < HV) fma4(acc[v], a, ldg_f4(x + (int64_t)n * HV + c));
        }
    }
}

// one float4 of a row, as fp32 or rounded to bf16 (the tensor-core node MLP consumes bf16 operands: writing them here
// saves a quarter of the kernel's traffic and the separate cast pass)
template <bool OUT_BF16>
__device__ __forceinline__ void store_row4(void* out, int64_t idx, const float4& v) {
    if (OUT_BF16) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
        reinterpret_cast<uint2*>(out)[idx] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
    } else {
        reinterpret_cast<float4*>(out)[idx] = v;
    }
}

template <int LPR, int NV, bool HAS_ATT, bool OUT_BF16>
__global__ void __launch_bounds__(K3_THREADS, 3)
k_gin_aggregate_fwd(const float4* __restrict__ x, const float* __restrict__ att, const int32_t* __restrict__ rowptr,
                    const int32_t* __restrict__ eid, const int32_t* __restrict__ nbr, float self_scale,
                    void* __restrict__ out, int64_t N, int HV) {
    constexpr int RPW = 32 / LPR;                      // sub-warps (rows side by side) per warp
    constexpr int UNITS = AGG_WORKERS * RPW;           // sub-warps per CTA
    constexpr int U = (NV == 1) ? 8 : (NV == 2 ? 4 : 2);   // entries in flight per sub-warp
    __shared__ AggTile tiles[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / LPR, sl = lane % LPR;
    const int64_t ntiles = (N + AGG_TILE - 1) / AGG_TILE;
    if (warp == AGG_WORKERS && (int64_t)blockIdx.x < ntiles)
        agg_stage<HAS_ATT, false>(tiles[0], rowptr, eid, nbr, att, (int64_t)blockIdx.x * AGG_TILE,
                                  (int)min((int64_t)AGG_TILE, N - (int64_t)blockIdx.x * AGG_TILE), self_scale, lane);
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        __syncthreads();       // tile `it` is staged; every worker has left tile it-1 (its buffer may be refilled)
        const int64_t t0 = tile * AGG_TILE;
        const int nr = (int)min((int64_t)AGG_TILE, N - t0);
        if (warp == AGG_WORKERS) {
            const int64_t nt = tile + gridDim.x;
            if (nt < ntiles)
                agg_stage<HAS_ATT, false>(tiles[(it + 1) & 1], rowptr, eid, nbr, att, nt * AGG_TILE,
                                          (int)min((int64_t)AGG_TILE, N - nt * AGG_TILE), self_scale, lane);
            continue;
        }
        const AggTile& t = tiles[it & 1];
        const int e0 = t.rp[0];
        const int per = (nr + UNITS - 1) / UNITS;
        const int ra = min(nr, (warp * RPW + sub) * per), rb = min(nr, ra + per);     // this sub-warp's rows
        if (!t.staged) {
            for (int r = ra; r < rb; ++r) {
                float4 acc[NV];
#pragma unroll
                for (int v = 0; v < NV; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
                agg_row_slow<LPR, NV, HAS_ATT>(x, att, eid, nbr, t.rp[r], t.rp[r + 1], sl, HV, acc);
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    const int c = sl + v * LPR;
                    if (c < HV) {
                        fma4(acc[v], self_scale, ldg_stream_f4(x + (t0 + r) * HV + c));
                        store_row4<OUT_BF16>(out, (t0 + r) * HV + c, acc[v]);
                    }
                }
            }
            continue;
        }
        int r = ra;
        int q = ra < nr ? t.rp[ra] - e0 + ra : 0;                         // first entry of this sub-warp
        const int qend = ra < rb ? t.rp[rb] - e0 + rb : q;               // one past its last entry
        int rend = ra < rb ? t.rp[ra + 1] - e0 + ra + 1 : 0;             // one past the current row's self entry
        float4 acc[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        while (q < qend) {
            float a[U];
            float4 g[U][NV];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int qq = min(q + u, qend - 1);                      // clamped: loads stay unconditional
                a[u] = t.att[qq];
                const float4* src = x + (int64_t)t.nbr[qq] * HV + sl;
#pragma unroll
                for (int v = 0; v < NV; ++v) g[u][v] = (v * LPR + sl < HV) ? ldg_f4(src + v * LPR) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (q + u < qend) {
#pragma unroll
                    for (int v = 0; v < NV; ++v) fma4(acc[v], a[u], g[u][v]);
                    if (q + u + 1 == rend) {                              // the self entry just closed row r
#pragma unroll
                        for (int v = 0; v < NV; ++v) {
                            if (v * LPR + sl < HV) store_row4<OUT_BF16>(out, (t0 + r) * HV + v * LPR + sl, acc[v]);
                            acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                        ++r;
                        rend = t.rp[min(r + 1, nr)] - e0 + r + 1;
                    }
                }
            }
            q += U;
        }
    }
}

// transpose-reduce of U per-lane partial sums over the LPR lanes of a sub-warp: afterwards lane (sl % U) ... every
// lane l < U of the sub-warp holds the full sum of value l (9 shuffles for U = 8 over 32 lanes instead of 40).
template <int LPR, int U>
__device__ __forceinline__ float subwarp_transpose_reduce(float (&v)[U], int sl, unsigned mask) {
    // fold the value index into the lane index while halving the number of live values
    int live = U;
#pragma unroll
    for (int off = LPR / 2; off >= 1; off >>= 1) {
        if (live > 1) {
            const bool upper = (sl & off) != 0;
            const int half = live / 2;
#pragma unroll
            for (int i = 0; i < U / 2; ++i) {
                if (i < half) {
                    const float send = upper ? v[i] : v[i + half];
                    const float keep = upper ? v[i + half] : v[i];
                    v[i] = keep + __shfl_xor_sync(mask, send, off);
                }
            }
            live = half;
        } else {
            v[0] += __shfl_xor_sync(mask, v[0], off);
        }
    }
    return v[0];
}

// backward over CSC rows (edges grouped by source j): dx[j] = sum att_e g[dst_e] + (1+eps) g[j];
// datt[e] = <x[j], g[dst_e]>
template <int LPR, int NV, bool HAS_ATT, bool WANT_DATT>
__global__ void __launch_bounds__(K3_THREADS, 3)
k_gin_aggregate_bwd(const float4* __restrict__ g, const float4* __restrict__ x, const float* __restrict__ att,
                    const int32_t* __restrict__ rowptr, const int32_t* __restrict__ eid,
                    const int32_t* __restrict__ nbr, float self_scale, float4* __restrict__ dx,
                    float* __restrict__ datt, int64_t N, int HV) {
    constexpr int RPW = 32 / LPR;
    constexpr int UNITS = AGG_WORKERS * RPW;
    // with d att every entry also fetches x[its own row] (L1 hits: consecutive entries share the row)
    constexpr int U = WANT_DATT ? ((NV == 1) ? 4 : (NV == 2 ? 2 : 1)) : ((NV == 1) ? 8 : (NV == 2 ? 4 : 2));
    __shared__ AggTile tiles[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / LPR, sl = lane % LPR;
    const unsigned submask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (sub * LPR));
    const int64_t ntiles = (N + AGG_TILE - 1) / AGG_TILE;
    if (warp == AGG_WORKERS && (int64_t)blockIdx.x < ntiles)
        agg_stage<HAS_ATT, WANT_DATT>(tiles[0], rowptr, eid, nbr, att, (int64_t)blockIdx.x * AGG_TILE,
                                      (int)min((int64_t)AGG_TILE, N - (int64_t)blockIdx.x * AGG_TILE), self_scale, lane);
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        __syncthreads();
        const int64_t t0 = tile * AGG_TILE;
        const int nr = (int)min((int64_t)AGG_TILE, N - t0);
        if (warp == AGG_WORKERS) {
            const int64_t nt = tile + gridDim.x;
            if (nt < ntiles)
                agg_stage<HAS_ATT, WANT_DATT>(tiles[(it + 1) & 1], rowptr, eid, nbr, att, nt * AGG_TILE,
                                              (int)min((int64_t)AGG_TILE, N - nt * AGG_TILE), self_scale, lane);
            continue;
        }
        const AggTile& t = tiles[it & 1];
        const int e0 = t.rp[0];
        const int per = (nr + UNITS - 1) / UNITS;
        const int ra = min(nr, (warp * RPW + sub) * per), rb = min(nr, ra + per);
        if (!t.staged) {
            for (int r = ra; r < rb; ++r) {
                float4 acc[NV], xr[NV];
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
                    const int c = sl + v * LPR;
                    xr[v] = (WANT_DATT && c < HV) ? ldg_stream_f4(x + (t0 + r) * HV + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                for (int p = t.rp[r]; p < t.rp[r + 1]; ++p) {
                    const int n = __ldg(nbr + p);
                    const int e = (HAS_ATT || WANT_DATT) ? __ldg(eid + p) : 0;
                    const float a = HAS_ATT ? __ldg(att + e) : 1.f;
                    float part = 0.f;
#pragma unroll
                    for (int v = 0; v < NV; ++v) {
                        const int c = sl + v * LPR;
                        if (c < HV) {
                            const float4 gg = ldg_f4(g + (int64_t)n * HV + c);
                            fma4(acc[v], a, gg);
                            if (WANT_DATT) part += dot4(xr[v], gg);
                        }
                    }
                    if (WANT_DATT) {
#pragma unroll
                        for (int o = LPR / 2; o > 0; o >>= 1) part += __shfl_xor_sync(submask, part, o);
                        if (sl == 0) datt[e] = part;
                    }
                }
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    const int c = sl + v * LPR;
                    if (c < HV) {
                        fma4(acc[v], self_scale, ldg_stream_f4(g + (t0 + r) * HV + c));
                        dx[(t0 + r) * HV + c] = acc[v];
                    }
                }
            }
            continue;
        }
        int r = ra;
        int q = ra < nr ? t.rp[ra] - e0 + ra : 0;
        const int qend = ra < rb ? t.rp[rb] - e0 + rb : q;
        int rend = ra < rb ? t.rp[ra + 1] - e0 + ra + 1 : 0;
        float4 acc[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        while (q < qend) {
            float a[U];
            float4 gg[U][NV], xx[U][NV];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int qq = min(q + u, qend - 1);
                a[u] = t.att[qq];
                const float4* src = g + (int64_t)t.nbr[qq] * HV + sl;
                const float4* xsrc = WANT_DATT ? x + (int64_t)t.row[qq] * HV + sl : nullptr;
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    const bool cok = v * LPR + sl < HV;
                    gg[u][v] = cok ? ldg_f4(src + v * LPR) : make_float4(0.f, 0.f, 0.f, 0.f);
                    if (WANT_DATT) xx[u][v] = cok ? ldg_f4(xsrc + v * LPR) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            float part[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                part[u] = 0.f;
                if (q + u < qend) {
#pragma unroll
                    for (int v = 0; v < NV; ++v) {
                        fma4(acc[v], a[u], gg[u][v]);
                        if (WANT_DATT) part[u] += dot4(xx[u][v], gg[u][v]);
                    }
                    if (q + u + 1 == rend) {
#pragma unroll
                        for (int v = 0; v < NV; ++v) {
                            if (v * LPR + sl < HV) dx[(t0 + r) * HV + v * LPR + sl] = acc[v];
                            acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                        ++r;
                        rend = t.rp[min(r + 1, nr)] - e0 + r + 1;
                    }
                }
            }
            if (WANT_DATT) {
                // lane l < U of the sub-warp ends with the dot product of entry q + l; self entries carry eid -1
                if (LPR >= U) {
                    const float d = subwarp_transpose_reduce<LPR, U>(part, sl, submask);
                    // after the fold, value index = bits of sl selected by the first log2(U) offsets (LPR/2, LPR/4, ...)
                    int idx = 0;
                    {
                        int live = U;
                        for (int off = LPR / 2; off >= 1 && live > 1; off >>= 1) {
                            live >>= 1;
                            if (sl & off) idx += live;
                        }
                    }
                    constexpr int LOWMASK = (LPR / U) - 1;      // lanes whose remaining (low) bits are zero own a value
                    if ((sl & LOWMASK) == 0 && q + idx < qend) {
                        const int e = t.eid[q + idx];
                        if (e >= 0) datt[e] = d;
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        float d = part[u];
#pragma unroll
                        for (int o = LPR / 2; o > 0; o >>= 1) d += __shfl_xor_sync(submask, d, o);
                        if (sl == 0 && q + u < qend) {
                            const int e = t.eid[q + u];
                            if (e >= 0) datt[e] = d;
                        }
                    }
                }
            }
            q += U;
        }
    }
}

// Rows of 16 / 32 float4 (H = 64 / 128) can be walked by half as many lanes carrying two columns each: twice the rows per
// warp instruction, i.e. half the per-entry instruction overhead per byte (H = 64 was issue bound at 16 lanes per row).
// Development switch GSATB_K3_WIDE: 0 = off, 1 = H = 64 only, 2 = H = 64 and 128 (default; measured on the B200 at
// N = 4.9 M: H = 64 fwd 0.717 -> 0.688 ms, H = 128 fwd 1.057 -> 1.022 ms, backward within 2 %).
inline int k3_wide_lanes(int HV) {
    static const int mode = [] {
        const char* e = getenv("GSATB_K3_WIDE");
        return e ? atoi(e) : 2;
    }();
    if (HV == 16 && mode >= 1) return 8;
    if (HV == 32 && mode >= 2) return 16;
    return 0;
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

// K3 kernels are persistent: `per_sm` resident CTAs per SM, each walking tiles with a grid stride
inline unsigned agg_tile_grid(int64_t N, int per_sm) {
    int64_t blocks = (N + AGG_TILE - 1) / AGG_TILE;
    const int64_t cap = (int64_t)GSATB_NUM_SMS * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

template <bool HAS_ATT, bool OUT_BF16>
int launch_fwd(const float* x, const float* att, const int32_t* rowptr, const int32_t* eid, const int32_t* nbr,
               float self_scale, void* out, int64_t N, int HV, cudaStream_t st) {
    int lpr = pick_lpr(HV);
    int nv = (HV + lpr - 1) / lpr;
    const unsigned grid = agg_tile_grid(N, 3);
    const int wide = k3_wide_lanes(HV);      // narrow rows: fewer lanes per row, two float4 columns per lane
    if (wide) lpr = wide, nv = 2;
#define FWD_CASE(L, V)                                                                                       \
    k_gin_aggregate_fwd<L, V, HAS_ATT, OUT_BF16><<<grid, K3_THREADS, 0, st>>>((const float4*)x, att, rowptr, eid, nbr, \
                                                                               self_scale, out, N, HV)
    if (wide == 8) FWD_CASE(8, 2);
    else if (wide == 16) FWD_CASE(16, 2);
    else if (nv == 1) {
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
    int lpr = pick_lpr(HV);
    int nv = (HV + lpr - 1) / lpr;
    const unsigned grid = agg_tile_grid(N, 3);
    const int wide = k3_wide_lanes(HV);
    if (wide) lpr = wide, nv = 2;
#define BWD_CASE(L, V)                                                                                  \
    k_gin_aggregate_bwd<L, V, HAS_ATT, WANT_DATT><<<grid, K3_THREADS, 0, st>>>(                         \
        (const float4*)g, (const float4*)x, att, rowptr, eid, nbr, self_scale, (float4*)dx, datt, N, HV)
    if (wide == 8) BWD_CASE(8, 2);
    else if (wide == 16) BWD_CASE(16, 2);
    else if (nv == 1) {
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
    int rc = att ? launch_fwd<true, false>(x, att, rowptr_dst, eid_by_dst, src_by_dst, 1.f + eps, out, N, H / 4, st)
                 : launch_fwd<false, false>(x, att, rowptr_dst, eid_by_dst, src_by_dst, 1.f + eps, out, N, H / 4, st);
    if (rc != GSATB_OK) return rc;
    GSATB_CHECK_LAUNCH();
    return GSATB_OK;
}

extern "C" int gsatb_gin_aggregate_fwd_bf16(const float* x, const float* att, const int32_t* rowptr_dst,
                                            const int32_t* eid_by_dst, const int32_t* src_by_dst, float eps,
                                            void* out_bf16, int64_t N, int64_t E, int H, gsatb_stream_t stream) {
    if (N < 0 || E < 0 || H <= 0) return GSATB_EINVAL;
    if (N == 0) return GSATB_OK;
    if (!x || !out_bf16 || !rowptr_dst || (E > 0 && (!src_by_dst || (att && !eid_by_dst)))) return GSATB_EINVAL;
    if (H % 4 != 0 || H > 512) return GSATB_ESHAPE;
    if (!gsatb_aligned16(x) || (reinterpret_cast<uintptr_t>(out_bf16) & 7u) != 0) return GSATB_EALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = att ? launch_fwd<true, true>(x, att, rowptr_dst, eid_by_dst, src_by_dst, 1.f + eps, out_bf16, N, H / 4, st)
                 : launch_fwd<false, true>(x, att, rowptr_dst, eid_by_dst, src_by_dst, 1.f + eps, out_bf16, N, H / 4, st);
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

// the same reduction over a bf16 g (d f12 as the fused extractor backward writes it in bf16 mode): half the bytes of the
// step's largest intermediate; fp32 accumulation in the same order.  A lane owns 8 channels (one 16-byte load per edge).
__device__ __forceinline__ void acc_bf16x8(float (&a)[8], const uint4& v) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        a[2 * i] += __uint_as_float(w[i] << 16);
        a[2 * i + 1] += __uint_as_float(w[i] & 0xFFFF0000u);
    }
}
template <int LPR>
__global__ void __launch_bounds__(AGG_THREADS)
k_gather_concat_bwd_bf16(const uint4* __restrict__ g, const int32_t* __restrict__ rowptr_src,
                         const int32_t* __restrict__ eid_by_src, const int32_t* __restrict__ rowptr_dst,
                         const int32_t* __restrict__ eid_by_dst, float4* __restrict__ demb, int64_t N, int HV8) {
    constexpr int ROWS_PER_WARP = 32 / LPR;
    const int lane = threadIdx.x & 31, sub = lane / LPR, sl = lane % LPR;
    const int64_t warp_global = (blockIdx.x * (int64_t)(AGG_THREADS / 32)) + (threadIdx.x >> 5);
    const int64_t warps_total = (int64_t)gridDim.x * (AGG_THREADS / 32);
    for (int64_t row0 = warp_global * ROWS_PER_WARP; row0 < N; row0 += warps_total * ROWS_PER_WARP) {
        const int64_t row = row0 + sub;
        if (row >= N) continue;
        for (int c = sl; c < HV8; c += LPR) {
            float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            for (int p = __ldg(rowptr_src + row), pe = __ldg(rowptr_src + row + 1); p < pe; ++p)
                acc_bf16x8(acc, __ldg(g + (int64_t)__ldg(eid_by_src + p) * 2 * HV8 + c));
            for (int p = __ldg(rowptr_dst + row), pe = __ldg(rowptr_dst + row + 1); p < pe; ++p)
                acc_bf16x8(acc, __ldg(g + (int64_t)__ldg(eid_by_dst + p) * 2 * HV8 + HV8 + c));
            demb[row * (2 * HV8) + 2 * c] = make_float4(acc[0], acc[1], acc[2], acc[3]);
            demb[row * (2 * HV8) + 2 * c + 1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
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

extern "C" int gsatb_gather_concat_bwd_bf16(const void* g_bf16, const int32_t* rowptr_src, const int32_t* eid_by_src,
                                            const int32_t* rowptr_dst, const int32_t* eid_by_dst, float* demb, int64_t N,
                                            int H, gsatb_stream_t stream) {
    if (N < 0 || H <= 0) return GSATB_EINVAL;
    if (N == 0) return GSATB_OK;
    if (!g_bf16 || !rowptr_src || !rowptr_dst || !demb) return GSATB_EINVAL;
    if (H % 8 != 0) return GSATB_ESHAPE;
    if (!gsatb_aligned16(g_bf16) || !gsatb_aligned16(demb)) return GSATB_EALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    const int HV = H / 8;
    const int lpr = pick_lpr(HV);
    const unsigned grid = agg_grid(N, lpr);
#define GC_CASE(L) k_gather_concat_bwd_bf16<L><<<grid, AGG_THREADS, 0, st>>>((const uint4*)g_bf16, rowptr_src, eid_by_src, rowptr_dst, eid_by_dst, (float4*)demb, N, HV)
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
