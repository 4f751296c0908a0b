// K0 -- index builder: CSR-by-dst, CSC-by-src, reverse-edge map, graph segment pointers, structure flags.
//
// Replaces reference src/utils/utils.py:19-25 (reorder_like: sort_edge_index + argsort().argsort() + equality
// check), src/run_gsat.py:242-243 (is_undirected, torch_sparse.transpose) and the per-call COO handling of PyG /
// torch_scatter.  The two canonical orders are the stable ascending (src,dst) and (dst,src) orders -- exactly the
// two rankings reorder_like compares -- built with a from-scratch stable LSD radix sort (8-bit digits) over 32-bit
// node ids: order(dst,src) = stable-sort-by-dst( stable-sort-by-src(identity) ).  All counters are integer, so
// the outputs are bit-exact and run-to-run deterministic.
//
// Traffic: the builder runs once per batch (not per step); algorithmic bytes 16E + 16E + 8(N+1) + 8(G+1)
// (SURVEY.md §8d), the radix passes add 16 B/edge/pass of scratch traffic.
#include "common.cuh"

namespace {

constexpr int RS_THREADS = 256;
constexpr int RS_ROUNDS = 8;
constexpr int RS_TILE = RS_THREADS * RS_ROUNDS;
constexpr int RS_WARPS = RS_THREADS / 32;

__global__ void k_convert_edges(const int64_t* __restrict__ ei, int64_t E, int64_t N, int32_t* __restrict__ src,
                                int32_t* __restrict__ dst, int32_t* __restrict__ ident, int32_t* flags) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= E) return;
    int64_t s = ei[e], d = ei[E + e];
    bool bad = s < 0 || s >= N || d < 0 || d >= N;
    if (bad) {
        atomicAdd(&flags[3], 1);
        s = 0;
        d = 0;
    }
    src[e] = (int32_t)s;
    dst[e] = (int32_t)d;
    ident[e] = (int32_t)e;
}

__global__ void __launch_bounds__(RS_THREADS) k_radix_hist(const int32_t* __restrict__ keys, int64_t E, int shift,
                                                           int numTiles, uint32_t* __restrict__ hist) {
    __shared__ uint32_t s_h[256];
    s_h[threadIdx.x] = 0;
    __syncthreads();
    int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll
    for (int r = 0; r < RS_ROUNDS; ++r) {
        int64_t i = base + r * RS_THREADS + threadIdx.x;
        if (i < E) atomicAdd(&s_h[((uint32_t)keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * numTiles + blockIdx.x] = s_h[threadIdx.x];
}

// block-wide inclusive scan of one value per thread (256 threads)
__device__ __forceinline__ uint32_t block_inclusive_scan_256(uint32_t v, uint32_t* s_warp /* [8] */) {
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    if (lane == 31) s_warp[warp] = v;
    __syncthreads();
    uint32_t add = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w)
        if (w < warp) add += s_warp[w];
    __syncthreads();
    return v + add;
}

// one block per digit: exclusive scan of that digit's per-tile counts, digit total to totals[d]
__global__ void __launch_bounds__(RS_THREADS) k_scan_rows(uint32_t* __restrict__ hist, int numTiles,
                                                          uint32_t* __restrict__ totals) {
    __shared__ uint32_t s_warp[RS_WARPS];
    __shared__ uint32_t s_carry;
    uint32_t* row = hist + (size_t)blockIdx.x * numTiles;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < numTiles; base += RS_THREADS) {
        int i = base + threadIdx.x;
        uint32_t v = i < numTiles ? row[i] : 0u;
        uint32_t incl = block_inclusive_scan_256(v, s_warp);
        uint32_t carry = s_carry;
        if (i < numTiles) row[i] = carry + incl - v;
        __syncthreads();
        if (threadIdx.x == RS_THREADS - 1) s_carry = carry + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) totals[blockIdx.x] = s_carry;
}

__global__ void __launch_bounds__(RS_THREADS) k_radix_scatter(const int32_t* __restrict__ kin,
                                                              const int32_t* __restrict__ vin,
                                                              int32_t* __restrict__ kout, int32_t* __restrict__ vout,
                                                              int64_t E, int shift, int numTiles,
                                                              const uint32_t* __restrict__ hist,
                                                              const uint32_t* __restrict__ totals) {
    __shared__ uint32_t s_base[256];
    __shared__ uint32_t s_wcnt[RS_WARPS][257];
    __shared__ uint32_t s_warp[RS_WARPS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    {
        uint32_t t = totals[tid];
        uint32_t incl = block_inclusive_scan_256(t, s_warp);
        s_base[tid] = incl - t + hist[(size_t)tid * numTiles + tile];
    }
    __syncthreads();
    for (int r = 0; r < RS_ROUNDS; ++r) {
        int64_t i = (int64_t)tile * RS_TILE + r * RS_THREADS + tid;
        bool valid = i < E;
        int32_t k = valid ? kin[i] : 0;
        int32_t v = valid ? vin[i] : 0;
        uint32_t d = valid ? (((uint32_t)k >> shift) & 255u) : 256u;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) s_wcnt[w][tid] = 0;
        __syncthreads();
        unsigned mask = __match_any_sync(0xffffffffu, d);
        int rank = __popc(mask & ((1u << lane) - 1u));
        if (valid && rank == 0) s_wcnt[warp][d] = (uint32_t)__popc(mask);
        __syncthreads();
        {
            uint32_t run = s_base[tid];
#pragma unroll
            for (int w = 0; w < RS_WARPS; ++w) {
                uint32_t c = s_wcnt[w][tid];
                s_wcnt[w][tid] = run;
                run += c;
            }
            s_base[tid] = run;
        }
        __syncthreads();
        if (valid) {
            uint32_t pos = s_wcnt[warp][d] + (uint32_t)rank;
            kout[pos] = k;
            vout[pos] = v;
        }
        __syncthreads();
    }
}

__global__ void k_gather_keys(const int32_t* __restrict__ table, const int32_t* __restrict__ idx,
                              int32_t* __restrict__ out, int64_t E) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < E) out[i] = table[idx[i]];
}

// sorted keys -> segment pointers: ptr[v] = first position whose key >= v ; ptr[n] = E
template <typename KeyT>
__global__ void k_ptr_from_sorted(const KeyT* __restrict__ keys, int64_t E, int64_t n, int32_t* __restrict__ ptr,
                                  int32_t* flags, int flag_slot) {
    int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p > E) return;
    int64_t cur = p < E ? (int64_t)keys[p] : n;
    int64_t prev = p > 0 ? (int64_t)keys[p - 1] : -1;
    if (p < E && (cur < 0 || cur >= n)) {
        atomicAdd(&flags[3], 1);
        return;
    }
    if (prev >= n || prev < -1) return;
    if (cur < prev) {
        if (flag_slot >= 0) atomicAdd(&flags[flag_slot], 1);
        return;
    }
    for (int64_t v = prev + 1; v <= cur; ++v) ptr[v] = (int32_t)p;
}

__global__ void k_finalize_order(const int32_t* __restrict__ eid_sorted, const int32_t* __restrict__ other_end,
                                 int32_t* __restrict__ eid_out, int32_t* __restrict__ other_out, int64_t E) {
    int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= E) return;
    int32_t e = eid_sorted[p];
    eid_out[p] = e;
    other_out[p] = other_end[e];
}

__global__ void k_reverse_map(const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                              const int32_t* __restrict__ eid_by_src, const int32_t* __restrict__ eid_by_dst,
                              int32_t* __restrict__ rev, int64_t E, int32_t* flags) {
    int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= E) return;
    int32_t a = eid_by_src[p], b = eid_by_dst[p];
    int32_t sa = src[a], da = dst[a];
    bool ok = (sa == dst[b]) && (da == src[b]);
    rev[a] = ok ? b : -1;
    if (!ok) atomicAdd(&flags[0], 1);
    if (p > 0) {
        int32_t a0 = eid_by_src[p - 1];
        if (src[a0] == sa && dst[a0] == da) atomicAdd(&flags[1], 1);
    }
}

__global__ void k_node_graph(const int64_t* __restrict__ batch, int64_t N, int32_t* __restrict__ node_graph) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < N) node_graph[i] = (int32_t)batch[i];
}

__global__ void k_edge_graph(const int32_t* __restrict__ src, const int32_t* __restrict__ dst,
                             const int32_t* __restrict__ node_graph, int32_t* __restrict__ edge_graph, int64_t E,
                             int32_t* flags) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= E) return;
    int32_t g = node_graph[src[e]];
    if (g != node_graph[dst[e]]) atomicAdd(&flags[2], 1);
    edge_graph[e] = g;
}

struct Pair {
    int32_t* k;
    int32_t* v;
};

inline int bits_for(int64_t n) {
    int b = 1;
    while (b < 31 && ((int64_t)1 << b) < n) ++b;
    return b;
}

inline unsigned grid1d(int64_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

// stable LSD radix sort of (keys, vals); first pass reads (kin, vin), passes ping-pong between `first` and `second`
// starting by writing `first`.  Returns the pair holding the sorted result.
Pair radix_sort_pairs(const int32_t* kin, const int32_t* vin, Pair first, Pair second, int64_t E, int passes,
                      int numTiles, uint32_t* hist, uint32_t* totals, cudaStream_t st) {
    Pair out = first, alt = second;
    const int32_t* ck = kin;
    const int32_t* cv = vin;
    for (int p = 0; p < passes; ++p) {
        int shift = 8 * p;
        k_radix_hist<<<numTiles, RS_THREADS, 0, st>>>(ck, E, shift, numTiles, hist);
        k_scan_rows<<<256, RS_THREADS, 0, st>>>(hist, numTiles, totals);
        k_radix_scatter<<<numTiles, RS_THREADS, 0, st>>>(ck, cv, out.k, out.v, E, shift, numTiles, hist, totals);
        ck = out.k;
        cv = out.v;
        Pair t = out;
        out = alt;
        alt = t;
    }
    return alt;   // the pair written last
}

}  // namespace

extern "C" size_t gsatb_index_build_workspace(int64_t N, int64_t E, int64_t G) {
    (void)N;
    (void)G;
    int64_t numTiles = (E + RS_TILE - 1) / RS_TILE;
    if (numTiles < 1) numTiles = 1;
    size_t pad = 256;
    size_t bytes = 0;
    bytes += 7 * (((size_t)(E > 0 ? E : 1) * 4 + pad - 1) / pad * pad);   // ident + 3 (key,val) pairs
    bytes += ((size_t)256 * numTiles * 4 + pad - 1) / pad * pad;          // per-tile digit histograms
    bytes += 256 * 4 + pad;                                               // digit totals
    return bytes;
}

extern "C" int gsatb_index_build(const int64_t* edge_index, const int64_t* batch, int64_t N, int64_t E, int64_t G,
                                 int32_t* src, int32_t* dst, int32_t* rev, int32_t* rowptr_dst, int32_t* eid_by_dst,
                                 int32_t* src_by_dst, int32_t* rowptr_src, int32_t* eid_by_src, int32_t* dst_by_src,
                                 int32_t* node_ptr, int32_t* edge_ptr, int32_t* node_graph, int32_t* edge_graph,
                                 int32_t* flags, void* ws, size_t ws_bytes, gsatb_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (N < 0 || E < 0 || G < 0 || N >= (1ll << 31) - 1 || E >= (1ll << 31) - 1) return GSATB_EINVAL;
    if (!rowptr_dst || !rowptr_src || !node_ptr || !edge_ptr || !flags) return GSATB_EINVAL;
    if (E > 0 && (!edge_index || !src || !dst || !rev || !eid_by_dst || !src_by_dst || !eid_by_src || !dst_by_src ||
                  !edge_graph || !ws))
        return GSATB_EINVAL;
    if (N > 0 && (!batch || !node_graph)) return GSATB_EINVAL;
    if (ws_bytes < gsatb_index_build_workspace(N, E, G)) return GSATB_EWS_TOO_SMALL;

    cudaMemsetAsync(flags, 0, 4 * sizeof(int32_t), st);
    const int T = 256;
    if (N > 0) {
        k_node_graph<<<grid1d(N, T), T, 0, st>>>(batch, N, node_graph);
        k_ptr_from_sorted<int64_t><<<grid1d(N + 1, T), T, 0, st>>>(batch, N, G, node_ptr, flags, 2);
    } else {
        cudaMemsetAsync(node_ptr, 0, (size_t)(G + 1) * 4, st);
    }
    if (E == 0) {
        cudaMemsetAsync(rowptr_dst, 0, (size_t)(N + 1) * 4, st);
        cudaMemsetAsync(rowptr_src, 0, (size_t)(N + 1) * 4, st);
        cudaMemsetAsync(edge_ptr, 0, (size_t)(G + 1) * 4, st);
        GSATB_CHECK_LAUNCH();
        return GSATB_OK;
    }
    if (N == 0) return GSATB_EINVAL;

    // carve the workspace
    const size_t pad = 256;
    size_t seg = ((size_t)E * 4 + pad - 1) / pad * pad;
    char* w = (char*)ws;
    int32_t* ident = (int32_t*)w;
    Pair P0{(int32_t*)(w + seg), (int32_t*)(w + 2 * seg)};
    Pair P1{(int32_t*)(w + 3 * seg), (int32_t*)(w + 4 * seg)};
    Pair P2{(int32_t*)(w + 5 * seg), (int32_t*)(w + 6 * seg)};
    int numTiles = (int)((E + RS_TILE - 1) / RS_TILE);
    uint32_t* hist = (uint32_t*)(w + 7 * seg);
    uint32_t* totals = (uint32_t*)((char*)hist + ((size_t)256 * numTiles * 4 + pad - 1) / pad * pad);

    k_convert_edges<<<grid1d(E, T), T, 0, st>>>(edge_index, E, N, src, dst, ident, flags);
    k_edge_graph<<<grid1d(E, T), T, 0, st>>>(src, dst, node_graph, edge_graph, E, flags);
    k_ptr_from_sorted<int32_t><<<grid1d(E + 1, T), T, 0, st>>>(edge_graph, E, G, edge_ptr, flags, 2);

    const int passes = (bits_for(N) + 7) / 8;

    // order (dst, src): stable sort by src, then stable sort by dst
    {
        Pair r1 = radix_sort_pairs(src, ident, P0, P1, E, passes, numTiles, hist, totals, st);
        Pair o1 = (r1.k == P0.k) ? P1 : P0;
        k_gather_keys<<<grid1d(E, T), T, 0, st>>>(dst, r1.v, o1.k, E);
        Pair r2 = radix_sort_pairs(o1.k, r1.v, P2, o1, E, passes, numTiles, hist, totals, st);
        k_ptr_from_sorted<int32_t><<<grid1d(E + 1, T), T, 0, st>>>(r2.k, E, N, rowptr_dst, flags, -1);
        k_finalize_order<<<grid1d(E, T), T, 0, st>>>(r2.v, src, eid_by_dst, src_by_dst, E);
    }
    // order (src, dst): stable sort by dst, then stable sort by src
    {
        Pair r1 = radix_sort_pairs(dst, ident, P0, P1, E, passes, numTiles, hist, totals, st);
        Pair o1 = (r1.k == P0.k) ? P1 : P0;
        k_gather_keys<<<grid1d(E, T), T, 0, st>>>(src, r1.v, o1.k, E);
        Pair r2 = radix_sort_pairs(o1.k, r1.v, P2, o1, E, passes, numTiles, hist, totals, st);
        k_ptr_from_sorted<int32_t><<<grid1d(E + 1, T), T, 0, st>>>(r2.k, E, N, rowptr_src, flags, -1);
        k_finalize_order<<<grid1d(E, T), T, 0, st>>>(r2.v, dst, eid_by_src, dst_by_src, E);
    }
    k_reverse_map<<<grid1d(E, T), T, 0, st>>>(src, dst, eid_by_src, eid_by_dst, rev, E, flags);
    GSATB_CHECK_LAUNCH();
    return GSATB_OK;
}

// ---- stable sort of edge ids by one int32 key (used by the line-graph builder: members of a source group in primal
// edge order) ------------------------------------------------------------------------------------------------------
namespace {
__global__ void k_iota(int32_t* __restrict__ v, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) v[i] = (int32_t)i;
}
}  // namespace

extern "C" size_t gsatb_stable_order_workspace(int64_t E) {
    int64_t numTiles = (E + RS_TILE - 1) / RS_TILE;
    if (numTiles < 1) numTiles = 1;
    const size_t pad = 256;
    size_t bytes = 5 * (((size_t)(E > 0 ? E : 1) * 4 + pad - 1) / pad * pad);      // ident + 2 (key,val) pairs
    bytes += ((size_t)256 * numTiles * 4 + pad - 1) / pad * pad + 256 * 4 + pad;
    return bytes;
}

// order[p] = id of the p-th element in ascending key order, ties in ascending id order (stable LSD radix sort)
extern "C" int gsatb_stable_order(const int32_t* keys, int64_t E, int64_t key_range, int32_t* order, void* ws,
                                  size_t ws_bytes, gsatb_stream_t stream) {
    if (E < 0 || key_range < 0 || E >= (1ll << 31) - 1) return GSATB_EINVAL;
    if (E == 0) return GSATB_OK;
    if (!keys || !order || !ws) return GSATB_EINVAL;
    if (ws_bytes < gsatb_stable_order_workspace(E)) return GSATB_EWS_TOO_SMALL;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t pad = 256;
    const size_t seg = ((size_t)E * 4 + pad - 1) / pad * pad;
    char* w = (char*)ws;
    int32_t* ident = (int32_t*)w;
    Pair P0{(int32_t*)(w + seg), (int32_t*)(w + 2 * seg)};
    Pair P1{(int32_t*)(w + 3 * seg), (int32_t*)(w + 4 * seg)};
    const int numTiles = (int)((E + RS_TILE - 1) / RS_TILE);
    uint32_t* hist = (uint32_t*)(w + 5 * seg);
    uint32_t* totals = (uint32_t*)((char*)hist + ((size_t)256 * numTiles * 4 + pad - 1) / pad * pad);
    k_iota<<<grid1d(E, 256), 256, 0, st>>>(ident, E);
    const int passes = (bits_for(key_range > 1 ? key_range : 2) + 7) / 8;
    Pair r = radix_sort_pairs(keys, ident, P0, P1, E, passes, numTiles, hist, totals, st);
    cudaMemcpyAsync(order, r.v, (size_t)E * 4, cudaMemcpyDeviceToDevice, st);
    GSATB_CHECK_LAUNCH();
    return GSATB_OK;
}
