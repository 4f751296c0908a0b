// Line-graph ("dual") builder of the DP-GSAT fork, SURVEY.md section 8f row 1.
//
// Replaces the O(E d) Python dict loops of reference src/datasets/mutag_dual.py:342-378 (and ba_2motifs_dual.py:35-62):
//   one dual node per directed primal edge (a, b), in primal edge order;
//   primal edges are grouped by their FIRST endpoint a, groups taken in order of first appearance of a in the edge
//   list (dict insertion order), members of a group in primal edge order m_0 < m_1 < ...;
//   for i < j (i outer, j inner) the dual edges (m_i, m_j) then (m_j, m_i) are appended.
// Optional `halve`: the fork's later relabelling (mutag_dual.py:536-548) gives the two directions of a primal edge,
// which are consecutive rows 2k, 2k+1 of the edge list, ONE dual node id -- here id = edge >> 1 (0-based).
//
// Built on K0's CSC row pointers (rowptr_src) and `members` = the edge ids stably sorted by source node (members of a
// group in ascending edge id; gsatb_stable_order -- K0's own eid_by_src orders a group by destination) in two calls,
// because the
// output size E_d = sum_v d(v) (d(v) - 1) is data dependent:
//   count: w[e] = d(d-1) at the first member e of each group, 0 elsewhere; exclusive prefix sum over e (so groups are
//          laid out in order of first appearance) and the total
//   fill : one thread per CSC position writes the pairs of its member against the later members of the group
// Integer work, bit-exact against oracle/gsat_oracle.py::line_graph_dual; HBM bound (16 B per dual edge written).
#include "common.cuh"

namespace {

constexpr int LG_THREADS = 256;
constexpr int LG_ITEMS = 4;                       // elements per thread in the scan kernels
constexpr int LG_TILE = LG_THREADS * LG_ITEMS;

__device__ __forceinline__ long long lg_weight(const int32_t* __restrict__ src, const int32_t* __restrict__ rowptr,
                                               const int32_t* __restrict__ eid_by_src, int64_t e) {
    const int s = __ldg(src + e);
    const int b = __ldg(rowptr + s), c = __ldg(rowptr + s + 1) - b;
    return (__ldg(eid_by_src + b) == (int32_t)e) ? (long long)c * (c - 1) : 0ll;
}

__device__ __forceinline__ long long block_sum(long long v, long long* s_warp) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = v;
    __syncthreads();
    long long t = 0;
    for (int w = 0; w < LG_THREADS / 32; ++w) t += s_warp[w];
    __syncthreads();
    return t;
}

__global__ void __launch_bounds__(LG_THREADS)
k_lg_tile_sums(const int32_t* __restrict__ src, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ eid_by_src,
               int64_t E, long long* __restrict__ tile_sum) {
    __shared__ long long s_warp[LG_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * LG_TILE;
    long long v = 0;
#pragma unroll
    for (int k = 0; k < LG_ITEMS; ++k) {
        const int64_t e = base + k * LG_THREADS + threadIdx.x;
        if (e < E) v += lg_weight(src, rowptr, eid_by_src, e);
    }
    const long long t = block_sum(v, s_warp);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = t;
}

// exclusive scan of the tile sums by ONE block (fixed order), total -> total[0]
__global__ void __launch_bounds__(LG_THREADS)
k_lg_scan_tiles(long long* __restrict__ tile_sum, int64_t ntiles, long long* __restrict__ total) {
    __shared__ long long s_warp[LG_THREADS / 32];
    __shared__ long long s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < ntiles; base += LG_THREADS) {
        const int64_t i = base + threadIdx.x;
        const long long v = i < ntiles ? tile_sum[i] : 0;
        // inclusive scan inside the block
        long long x = v;
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_warp[w] = x;
        __syncthreads();
        long long woff = 0;
        for (int q = 0; q < w; ++q) woff += s_warp[q];
        long long blk = 0;
        for (int q = 0; q < LG_THREADS / 32; ++q) blk += s_warp[q];
        const long long carry = s_carry;
        if (i < ntiles) tile_sum[i] = carry + woff + x - v;       // exclusive
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + blk;
        __syncthreads();
    }
    if (threadIdx.x == 0) total[0] = s_carry;
}

__global__ void __launch_bounds__(LG_THREADS)
k_lg_offsets(const int32_t* __restrict__ src, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ eid_by_src,
             int64_t E, const long long* __restrict__ tile_off, long long* __restrict__ offs) {
    // thread t owns LG_ITEMS CONSECUTIVE elements of the tile: exclusive scan = thread prefix + scan of thread sums
    __shared__ long long s_thread[LG_THREADS];
    const int64_t base = (int64_t)blockIdx.x * LG_TILE + (int64_t)threadIdx.x * LG_ITEMS;
    long long w[LG_ITEMS], tsum = 0;
#pragma unroll
    for (int k = 0; k < LG_ITEMS; ++k) {
        const int64_t e = base + k;
        w[k] = e < E ? lg_weight(src, rowptr, eid_by_src, e) : 0;
        tsum += w[k];
    }
    s_thread[threadIdx.x] = tsum;
    __syncthreads();
    if (threadIdx.x == 0) {                     // 256 values: a serial exclusive scan by one thread is enough
        long long run = tile_off[blockIdx.x];
        for (int t = 0; t < LG_THREADS; ++t) {
            const long long v = s_thread[t];
            s_thread[t] = run;
            run += v;
        }
    }
    __syncthreads();
    long long run = s_thread[threadIdx.x];
#pragma unroll
    for (int k = 0; k < LG_ITEMS; ++k) {
        const int64_t e = base + k;
        if (e < E) offs[e] = run;
        run += w[k];
    }
}

__global__ void __launch_bounds__(LG_THREADS)
k_lg_fill(const int32_t* __restrict__ src, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ eid_by_src,
          const long long* __restrict__ offs, int64_t E, int64_t Ed, int halve, long long* __restrict__ dsrc,
          long long* __restrict__ ddst) {
    const int64_t p = (int64_t)blockIdx.x * LG_THREADS + threadIdx.x;        // CSC position
    if (p >= E) return;
    const int e = __ldg(eid_by_src + p);
    const int s = __ldg(src + e);
    const int b = __ldg(rowptr + s), c = __ldg(rowptr + s + 1) - b;
    const int i = (int)(p - b);
    if (c < 2 || i >= c - 1) return;
    const long long base = offs[__ldg(eid_by_src + b)];
    const long long row = (long long)i * c - (long long)i * (i + 1) / 2;     // pairs (i', *) with i' < i
    const long long me = halve ? (e >> 1) : e;
    for (int j = i + 1; j < c; ++j) {
        const int mj = __ldg(eid_by_src + b + j);
        const long long other = halve ? (mj >> 1) : mj;
        const long long pos = base + 2 * (row + (j - i - 1));
        if (pos + 1 < Ed) {
            dsrc[pos] = me;
            ddst[pos] = other;
            dsrc[pos + 1] = other;
            ddst[pos + 1] = me;
        }
    }
}

__global__ void k_lg_batch(const int32_t* __restrict__ src, const long long* __restrict__ node_graph, int64_t nd, int halve,
                           long long* __restrict__ dual_batch) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nd) return;
    const int64_t e = halve ? 2 * k : k;
    dual_batch[k] = node_graph[__ldg(src + e)];
}

inline int64_t lg_tiles(int64_t E) { return (E + LG_TILE - 1) / LG_TILE; }

}  // namespace

extern "C" size_t gsatb_line_graph_workspace(int64_t N, int64_t E) {
    (void)N;
    return (size_t)(lg_tiles(E) + 1) * sizeof(long long) + 256;
}

extern "C" int gsatb_line_graph_count(const int32_t* src, const int32_t* rowptr_src, const int32_t* eid_by_src, int64_t N,
                                      int64_t E, int64_t* offs, int64_t* total, void* ws, size_t ws_bytes,
                                      gsatb_stream_t stream) {
    if (N < 0 || E < 0 || !total) return GSATB_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    if (E == 0) {
        cudaMemsetAsync(total, 0, sizeof(int64_t), st);
        return GSATB_OK;
    }
    if (!src || !rowptr_src || !eid_by_src || !offs || !ws) return GSATB_EINVAL;
    if (ws_bytes < gsatb_line_graph_workspace(N, E)) return GSATB_EWS_TOO_SMALL;
    const int64_t nt = lg_tiles(E);
    long long* tile_sum = (long long*)ws;
    k_lg_tile_sums<<<(unsigned)nt, LG_THREADS, 0, st>>>(src, rowptr_src, eid_by_src, E, tile_sum);
    k_lg_scan_tiles<<<1, LG_THREADS, 0, st>>>(tile_sum, nt, (long long*)total);
    k_lg_offsets<<<(unsigned)nt, LG_THREADS, 0, st>>>(src, rowptr_src, eid_by_src, E, tile_sum, (long long*)offs);
    GSATB_CHECK_LAUNCH();
    return GSATB_OK;
}

extern "C" int gsatb_line_graph_fill(const int32_t* src, const int32_t* rowptr_src, const int32_t* eid_by_src,
                                     const int64_t* offs, const int64_t* node_graph, int64_t N, int64_t E, int halve,
                                     int64_t* dual_edge_index, int64_t Ed, int64_t* dual_batch, gsatb_stream_t stream) {
    if (N < 0 || E < 0 || Ed < 0) return GSATB_EINVAL;
    if (E == 0) return GSATB_OK;
    if (!src || !rowptr_src || !eid_by_src || !offs || (Ed > 0 && !dual_edge_index)) return GSATB_EINVAL;
    if (halve && (E & 1)) return GSATB_ESHAPE;
    cudaStream_t st = (cudaStream_t)stream;
    if (Ed > 0)
        k_lg_fill<<<(unsigned)((E + LG_THREADS - 1) / LG_THREADS), LG_THREADS, 0, st>>>(
            src, rowptr_src, eid_by_src, (const long long*)offs, E, Ed, halve, (long long*)dual_edge_index,
            (long long*)dual_edge_index + Ed);
    if (dual_batch && node_graph) {
        const int64_t nd = halve ? E / 2 : E;
        k_lg_batch<<<(unsigned)((nd + 255) / 256), 256, 0, st>>>(src, (const long long*)node_graph, nd, halve,
                                                                  (long long*)dual_batch);
    }
    GSATB_CHECK_LAUNCH();
    return GSATB_OK;
}
