// On-device per-batch explanation metrics.
// Replaces the Python loop over graphs of reference src/run_gsat.py:783-791 (GSAT.get_precision_at_k: boolean masks
// over all edges + numpy argsort per graph, after a .cpu() of the attention):
//   precision[g] = (sum of exp_labels over the k edges of graph g with the largest attention) / k
// One CTA per graph over its contiguous edge range (edge_ptr from K0).  Instead of sorting, every edge computes its rank
// = number of edges of the graph that beat it (larger attention, ties to the smaller edge index -- numpy's argsort is
// not stable, so ties are unspecified in the reference; attention averaged over reverse edges ties in pairs that carry
// the same label); an edge counts iff rank < k.  O(n^2) comparisons on values staged in shared memory (n ~ 50 .. 400),
// integer label sums: deterministic.
#include "common.cuh"

namespace {

constexpr int PK_THREADS = 128;
constexpr int PK_CHUNK = 1024;

__global__ void __launch_bounds__(PK_THREADS)
k_precision_at_k(const float* __restrict__ att, const float* __restrict__ labels, const int32_t* __restrict__ edge_ptr,
                 int k, float* __restrict__ out) {
    __shared__ float s_att[PK_CHUNK];
    __shared__ float s_red[PK_THREADS / 32];
    const int g = blockIdx.x;
    const int b0 = __ldg(edge_ptr + g), b1 = __ldg(edge_ptr + g + 1);
    const int n = b1 - b0;
    float local = 0.f;
    for (int base = 0; base < n; base += PK_THREADS) {          // this thread's edge (uniform trip count)
        const int i = base + threadIdx.x;
        const bool live = i < n;
        const float ai = live ? __ldg(att + b0 + i) : 0.f;
        int rank = 0;
        for (int c0 = 0; c0 < n; c0 += PK_CHUNK) {
            const int cn = min(PK_CHUNK, n - c0);
            __syncthreads();
            for (int j = threadIdx.x; j < cn; j += PK_THREADS) s_att[j] = __ldg(att + b0 + c0 + j);
            __syncthreads();
            if (live) {
                for (int j = 0; j < cn; ++j) {
                    const float aj = s_att[j];
                    rank += (aj > ai || (aj == ai && c0 + j < i)) ? 1 : 0;
                }
            }
        }
        if (live && rank < k) local += __ldg(labels + b0 + i);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < PK_THREADS / 32; ++w) t += s_red[w];
        out[g] = t / (float)k;
    }
}

}  // namespace

extern "C" int gsatb_precision_at_k(const float* att, const float* exp_labels, const int32_t* edge_ptr, int64_t G, int k,
                                    float* precision, gsatb_stream_t stream) {
    if (G < 0 || k <= 0) return GSATB_EINVAL;
    if (G == 0) return GSATB_OK;
    if (!att || !exp_labels || !edge_ptr || !precision) return GSATB_EINVAL;
    k_precision_at_k<<<(unsigned)G, PK_THREADS, 0, (cudaStream_t)stream>>>(att, exp_labels, edge_ptr, k, precision);
    GSATB_CHECK_LAUNCH();
    return GSATB_OK;
}
