// K2 -- fused concrete (Gumbel-sigmoid) sampler + undirected reverse-edge average + information (KL) loss,
// K2' node->edge lift, and the reverse gather behind reorder_like.
//
// Replaces reference src/run_gsat.py:866-885 (sampling / concrete_sample), :241-247 (transpose + reorder_like +
// average), :126-132 and example/gsat.py:30-31 (info loss), src/run_gsat.py:870-875 (lift): ~25 elementwise /
// gather / reduction launches and 3 argsorts collapse into one pass over the edges.
//
// HBM bound: 12-20 B/edge forward, 20 B/edge backward.  The partner's attention is recomputed
// from logit[rev] and its own noise instead of being re-read after a grid sync.  Accurate logf/expf are used
// (no ex2/lg2 approximations): the parity bar for this op is rtol 1e-5 against the fp32 oracle.
#include "common.cuh"

namespace {

constexpr int SMP_THREADS = 256;

__device__ __forceinline__ float sigmoidf_acc(float z) { return 1.f / (1.f + expf(-z)); }

__device__ __forceinline__ float draw_u(const float* noise_u, uint64_t seed, uint64_t offset, int64_t e) {
    if (noise_u) return __ldg(noise_u + e);
    uint64_t ctr = offset + (uint64_t)e;
    uint4 r = philox4x32_10(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u),
                            make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    return u01_clamped(r.x);
}

__device__ __forceinline__ float concrete(float logit, float u, float inv_temp, bool training) {
    if (!training) return sigmoidf_acc(logit);
    float noise = logf(u) - logf(1.0f - u);
    return sigmoidf_acc((logit + noise) * inv_temp);
}

// f(a) = a log(a/r + 1e-6) + (1-a) log((1-a)/(1-r+1e-6) + 1e-6)
__device__ __forceinline__ float info_f(float a, float r) {
    const float c = 1.f - r + 1e-6f;
    return a * logf(a / r + 1e-6f) + (1.f - a) * logf((1.f - a) / c + 1e-6f);
}
__device__ __forceinline__ float info_df(float a, float r) {
    const float c = 1.f - r + 1e-6f;
    const float t1 = a / r + 1e-6f, t2 = (1.f - a) / c + 1e-6f;
    return logf(t1) + (a / r) / t1 - logf(t2) - ((1.f - a) / c) / t2;
}

__global__ void __launch_bounds__(SMP_THREADS)
k_sample_fwd(const float* __restrict__ logit, const float* __restrict__ noise_u, const int32_t* __restrict__ rev,
             const float* __restrict__ r_tensor, float r_scalar, float inv_temp, int mode, uint64_t seed,
             uint64_t offset_in, const unsigned long long* __restrict__ step, float* __restrict__ att,
             float* __restrict__ edge_att, float* __restrict__ partial, int64_t E) {
    // the optional device step counter selects a fresh 2^32-wide window of the Philox stream on every graph replay
    const uint64_t offset = offset_in + (step ? ((uint64_t)__ldg(step) << 32) : 0ull);
    const bool training = mode & GSATB_MODE_TRAINING, average = mode & GSATB_MODE_AVERAGE;
    const bool on_edge = mode & GSATB_MODE_INFO_ON_EDGE_ATT, want_info = !(mode & GSATB_MODE_NO_INFO);
    float local = 0.f;
    for (int64_t e = blockIdx.x * (int64_t)SMP_THREADS + threadIdx.x; e < E; e += (int64_t)gridDim.x * SMP_THREADS) {
        const float a = concrete(__ldg(logit + e), training ? draw_u(noise_u, seed, offset, e) : 0.5f, inv_temp, training);
        float ea = a;
        if (average) {
            const int32_t p = __ldg(rev + e);
            if (p >= 0) {
                const float b = (p == e) ? a
                                         : concrete(__ldg(logit + p), training ? draw_u(noise_u, seed, offset, p) : 0.5f,
                                                    inv_temp, training);
                ea = (a + b) / 2.f;
            }
        }
        if (att) att[e] = a;
        if (edge_att) edge_att[e] = ea;
        if (want_info) {
            const float r = r_tensor ? __ldg(r_tensor + e) : r_scalar;
            local += info_f(on_edge ? ea : a, r);
        }
    }
    if (want_info) {
        // fixed-shape block reduction -> one partial per block (deterministic)
        __shared__ float s_red[SMP_THREADS / 32];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = local;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < SMP_THREADS / 32; ++w) t += s_red[w];
            partial[blockIdx.x] = t;
        }
    }
}

__global__ void __launch_bounds__(1024) k_final_mean(const float* __restrict__ partial, int n, float inv_count,
                                                     float* __restrict__ out) {
    __shared__ double s_red[32];
    double local = 0.0;
    for (int i = threadIdx.x; i < n; i += 1024) local += (double)partial[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 32; ++w) t += s_red[w];
        out[0] = (float)(t * (double)inv_count);
    }
}

__global__ void __launch_bounds__(SMP_THREADS)
k_sample_bwd(const float* __restrict__ g_att, const float* __restrict__ g_edge_att, const float* __restrict__ g_info,
             const float* __restrict__ att, const float* __restrict__ edge_att, const int32_t* __restrict__ rev,
             const float* __restrict__ r_tensor, float r_scalar, float inv_temp, int mode,
             float* __restrict__ dlogit, int64_t E) {
    const bool training = mode & GSATB_MODE_TRAINING, average = mode & GSATB_MODE_AVERAGE;
    const bool on_edge = mode & GSATB_MODE_INFO_ON_EDGE_ATT, want_info = !(mode & GSATB_MODE_NO_INFO) && g_info;
    const float gi = want_info ? __ldg(g_info) / (float)E : 0.f;
    for (int64_t e = blockIdx.x * (int64_t)SMP_THREADS + threadIdx.x; e < E; e += (int64_t)gridDim.x * SMP_THREADS) {
        const float a = __ldg(att + e);
        float ga = g_att ? __ldg(g_att + e) : 0.f;
        // gradient arriving on edge_att[e'] for e' in {e, rev[e]} (rev is an involution on matched edges)
        auto g_on_edge = [&](int64_t q) -> float {
            float v = g_edge_att ? __ldg(g_edge_att + q) : 0.f;
            if (want_info && on_edge) {
                const float r = r_tensor ? __ldg(r_tensor + q) : r_scalar;
                v += gi * info_df(__ldg(edge_att + q), r);
            }
            return v;
        };
        if (average) {
            const int32_t p = __ldg(rev + e);
            if (p < 0) ga += g_on_edge(e);
            else if (p == e) ga += g_on_edge(e);
            else ga += 0.5f * (g_on_edge(e) + g_on_edge(p));
        } else {
            ga += g_on_edge(e);
        }
        if (want_info && !on_edge) {
            const float r = r_tensor ? __ldg(r_tensor + e) : r_scalar;
            ga += gi * info_df(a, r);
        }
        const float scale = training ? inv_temp : 1.f;
        dlogit[e] = ga * a * (1.f - a) * scale;
    }
}

__global__ void k_gather_rev(const float* __restrict__ v, const int32_t* __restrict__ rev, float* __restrict__ out,
                             int64_t E, int C) {
    const int64_t total = E * C;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = i / C;
        const int c = (int)(i - e * C);
        const int32_t p = __ldg(rev + e);
        out[i] = p >= 0 ? __ldg(v + (int64_t)p * C + c) : 0.f;
    }
}

__global__ void k_lift_fwd(const float* __restrict__ a, const int32_t* __restrict__ src,
                           const int32_t* __restrict__ dst, float* __restrict__ out, int64_t E) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x)
        out[e] = __ldg(a + __ldg(src + e)) * __ldg(a + __ldg(dst + e));
}

// d a[j] = sum_{e: src(e)=j} g_e a[dst(e)]  +  sum_{e: dst(e)=j} g_e a[src(e)]   (CSC then CSR order, sequential)
__global__ void k_lift_bwd(const float* __restrict__ g, const float* __restrict__ a,
                           const int32_t* __restrict__ rowptr_dst, const int32_t* __restrict__ eid_by_dst,
                           const int32_t* __restrict__ src_by_dst, const int32_t* __restrict__ rowptr_src,
                           const int32_t* __restrict__ eid_by_src, const int32_t* __restrict__ dst_by_src,
                           float* __restrict__ da, int64_t N) {
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < N; j += (int64_t)gridDim.x * blockDim.x) {
        float acc = 0.f;
        for (int p = __ldg(rowptr_src + j), pe = __ldg(rowptr_src + j + 1); p < pe; ++p)
            acc = fmaf(__ldg(g + __ldg(eid_by_src + p)), __ldg(a + __ldg(dst_by_src + p)), acc);
        for (int p = __ldg(rowptr_dst + j), pe = __ldg(rowptr_dst + j + 1); p < pe; ++p)
            acc = fmaf(__ldg(g + __ldg(eid_by_dst + p)), __ldg(a + __ldg(src_by_dst + p)), acc);
        da[j] = acc;
    }
}

__global__ void k_match_orders(const int32_t* __restrict__ order_to, const int32_t* __restrict__ order_from,
                               const int32_t* __restrict__ src_to, const int32_t* __restrict__ dst_to,
                               const int32_t* __restrict__ src_from, const int32_t* __restrict__ dst_from,
                               int32_t* __restrict__ map, int32_t* mismatch, int64_t E) {
    int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= E) return;
    const int32_t t = order_to[p], f = order_from[p];
    map[t] = f;
    if (src_to[t] != src_from[f] || dst_to[t] != dst_from[f]) atomicAdd(mismatch, 1);
}

inline int sample_blocks(int64_t E) {
    int64_t b = (E + SMP_THREADS - 1) / SMP_THREADS;
    int64_t cap = (int64_t)GSATB_NUM_SMS * 8;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace

extern "C" size_t gsatb_sample_workspace(int64_t E) {
    (void)E;
    return (size_t)GSATB_NUM_SMS * 8 * sizeof(float) + 256;
}

extern "C" int gsatb_sample_avg_info_fwd(const float* logit, const float* noise_u, const int32_t* rev,
                                         const float* r_tensor, float r_scalar, float temp, int mode, uint64_t seed,
                                         uint64_t offset, float* att, float* edge_att, float* info_mean, int64_t E,
                                         void* ws, size_t ws_bytes, gsatb_stream_t stream) {
    if (E < 0 || temp <= 0.f) return GSATB_EINVAL;
    const bool want_info = !(mode & GSATB_MODE_NO_INFO);
    if (E == 0) {
        if (want_info && info_mean) cudaMemsetAsync(info_mean, 0, sizeof(float), (cudaStream_t)stream);
        return GSATB_OK;
    }
    if (!logit || (!att && !edge_att)) return GSATB_EINVAL;
    if ((mode & GSATB_MODE_AVERAGE) && !rev) return GSATB_EINVAL;
    if (want_info && (!info_mean || !ws)) return GSATB_EINVAL;
    if (want_info && ws_bytes < gsatb_sample_workspace(E)) return GSATB_EWS_TOO_SMALL;
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = sample_blocks(E);
    k_sample_fwd<<<blocks, SMP_THREADS, 0, st>>>(logit, noise_u, rev, r_tensor, r_scalar, 1.f / temp, mode, seed,
                                                 offset, gsatb_step_counter_ref(), att, edge_att, (float*)ws, E);
    if (want_info) k_final_mean<<<1, 1024, 0, st>>>((const float*)ws, blocks, 1.f / (float)E, info_mean);
    GSATB_CHECK_LAUNCH();
    return GSATB_OK;
}

extern "C" int gsatb_sample_avg_info_bwd(const float* g_att, const float* g_edge_att, const float* g_info,
                                         const float* att, const float* edge_att, const int32_t* rev,
                                         const float* r_tensor, float r_scalar, float temp, int mode, float* dlogit,
                                         int64_t E, gsatb_stream_t stream) {
    if (E < 0 || temp <= 0.f) return GSATB_EINVAL;
    if (E == 0) return GSATB_OK;
    if (!att || !dlogit) return GSATB_EINVAL;
    if ((mode & GSATB_MODE_AVERAGE) && !rev) return GSATB_EINVAL;
    if ((mode & GSATB_MODE_INFO_ON_EDGE_ATT) && !(mode & GSATB_MODE_NO_INFO) && g_info && !edge_att) return GSATB_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    k_sample_bwd<<<sample_blocks(E), SMP_THREADS, 0, st>>>(g_att, g_edge_att, g_info, att, edge_att, rev, r_tensor,
                                                           r_scalar, 1.f / temp, mode, dlogit, E);
    GSATB_CHECK_LAUNCH();
    return GSATB_OK;
}

extern "C" int gsatb_gather_rev(const float* v, const int32_t* rev, float* out, int64_t E, int C,
                                gsatb_stream_t stream) {
    if (E < 0 || C <= 0) return GSATB_EINVAL;
    if (E == 0) return GSATB_OK;
    if (!v || !rev || !out) return GSATB_EINVAL;
    int64_t blocks = (E * C + 255) / 256;
    if (blocks > (int64_t)GSATB_NUM_SMS * 16) blocks = (int64_t)GSATB_NUM_SMS * 16;
    k_gather_rev<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(v, rev, out, E, C);
    GSATB_CHECK_LAUNCH();
    return GSATB_OK;
}

extern "C" int gsatb_lift_fwd(const float* node_att, const int32_t* src, const int32_t* dst, float* edge_att,
                              int64_t E, gsatb_stream_t stream) {
    if (E < 0) return GSATB_EINVAL;
    if (E == 0) return GSATB_OK;
    if (!node_att || !src || !dst || !edge_att) return GSATB_EINVAL;
    int64_t blocks = (E + 255) / 256;
    if (blocks > (int64_t)GSATB_NUM_SMS * 16) blocks = (int64_t)GSATB_NUM_SMS * 16;
    k_lift_fwd<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(node_att, src, dst, edge_att, E);
    GSATB_CHECK_LAUNCH();
    return GSATB_OK;
}

extern "C" int gsatb_lift_bwd(const float* g_edge, const float* node_att, const int32_t* rowptr_dst,
                              const int32_t* eid_by_dst, const int32_t* src_by_dst, const int32_t* rowptr_src,
                              const int32_t* eid_by_src, const int32_t* dst_by_src, float* d_node, int64_t N,
                              gsatb_stream_t stream) {
    if (N < 0) return GSATB_EINVAL;
    if (N == 0) return GSATB_OK;
    if (!g_edge || !node_att || !rowptr_dst || !rowptr_src || !d_node) return GSATB_EINVAL;
    int64_t blocks = (N + 255) / 256;
    if (blocks > (int64_t)GSATB_NUM_SMS * 16) blocks = (int64_t)GSATB_NUM_SMS * 16;
    k_lift_bwd<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(g_edge, node_att, rowptr_dst, eid_by_dst,
                                                                  src_by_dst, rowptr_src, eid_by_src, dst_by_src,
                                                                  d_node, N);
    GSATB_CHECK_LAUNCH();
    return GSATB_OK;
}

extern "C" int gsatb_match_orders(const int32_t* order_to, const int32_t* order_from, const int32_t* src_to,
                                  const int32_t* dst_to, const int32_t* src_from, const int32_t* dst_from,
                                  int32_t* map, int32_t* mismatch, int64_t E, gsatb_stream_t stream) {
    if (E < 0) return GSATB_EINVAL;
    if (!mismatch) return GSATB_EINVAL;
    cudaMemsetAsync(mismatch, 0, sizeof(int32_t), (cudaStream_t)stream);
    if (E == 0) return GSATB_OK;
    if (!order_to || !order_from || !src_to || !dst_to || !src_from || !dst_from || !map) return GSATB_EINVAL;
    k_match_orders<<<(unsigned)((E + 255) / 256), 256, 0, (cudaStream_t)stream>>>(order_to, order_from, src_to, dst_to,
                                                                                src_from, dst_from, map, mismatch, E);
    GSATB_CHECK_LAUNCH();
    return GSATB_OK;
}
