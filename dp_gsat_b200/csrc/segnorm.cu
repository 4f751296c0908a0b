// Per-graph InstanceNorm over contiguous row segments -- the normalisation between the Linears of the extractor
// MLP (reference src/utils/get_model.py:47-68 with torch_geometric InstanceNorm: eps 1e-5, no affine, biased
// variance of the centred values; src/run_gsat.py:912-915 passes batch[col] as the segment id).
//
// Replaces 2 scatter-adds + 2 index_selects + ~6 elementwise launches (and a batch.max() host sync) per norm.
// One CTA column-slab per graph: thread = channel, rows of the segment walked sequentially (coalesced across
// channels), mean -> centred variance -> normalise; the segment is re-read from L1/L2, so HBM traffic stays at
// read-once / write-once: 8*M*C bytes forward, 12*M*C backward.
#include "common.cuh"

namespace {

constexpr int SN_THREADS = 128;

__global__ void __launch_bounds__(SN_THREADS)
k_segnorm_fwd(const float* __restrict__ x, const int32_t* __restrict__ seg_ptr, float* __restrict__ y,
              float* __restrict__ rstd, int C, float eps) {
    const int g = blockIdx.x;
    const int c = blockIdx.y * SN_THREADS + threadIdx.x;
    if (c >= C) return;
    const int beg = __ldg(seg_ptr + g), end = __ldg(seg_ptr + g + 1);
    const int n = end - beg;
    if (n <= 0) {
        rstd[(int64_t)g * C + c] = 0.f;
        return;
    }
    const float inv_n = 1.f / (float)n;
    const float* xp = x + (int64_t)beg * C + c;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int i = 0;
    for (; i + 4 <= n; i += 4) {
        s0 += __ldg(xp + (int64_t)(i + 0) * C);
        s1 += __ldg(xp + (int64_t)(i + 1) * C);
        s2 += __ldg(xp + (int64_t)(i + 2) * C);
        s3 += __ldg(xp + (int64_t)(i + 3) * C);
    }
    for (; i < n; ++i) s0 += __ldg(xp + (int64_t)i * C);
    const float mean = ((s0 + s1) + (s2 + s3)) * inv_n;
    s0 = s1 = s2 = s3 = 0.f;
    i = 0;
    for (; i + 4 <= n; i += 4) {
        float d0 = __ldg(xp + (int64_t)(i + 0) * C) - mean;
        float d1 = __ldg(xp + (int64_t)(i + 1) * C) - mean;
        float d2 = __ldg(xp + (int64_t)(i + 2) * C) - mean;
        float d3 = __ldg(xp + (int64_t)(i + 3) * C) - mean;
        s0 = fmaf(d0, d0, s0);
        s1 = fmaf(d1, d1, s1);
        s2 = fmaf(d2, d2, s2);
        s3 = fmaf(d3, d3, s3);
    }
    for (; i < n; ++i) {
        float d0 = __ldg(xp + (int64_t)i * C) - mean;
        s0 = fmaf(d0, d0, s0);
    }
    const float var = ((s0 + s1) + (s2 + s3)) * inv_n;
    const float rs = 1.f / sqrtf(var + eps);
    rstd[(int64_t)g * C + c] = rs;
    float* yp = y + (int64_t)beg * C + c;
    for (i = 0; i < n; ++i) yp[(int64_t)i * C] = (__ldg(xp + (int64_t)i * C) - mean) * rs;
}

__global__ void __launch_bounds__(SN_THREADS)
k_segnorm_bwd(const float* __restrict__ gy, const float* __restrict__ y, const float* __restrict__ rstd,
              const int32_t* __restrict__ seg_ptr, float* __restrict__ gx, int C) {
    const int g = blockIdx.x;
    const int c = blockIdx.y * SN_THREADS + threadIdx.x;
    if (c >= C) return;
    const int beg = __ldg(seg_ptr + g), end = __ldg(seg_ptr + g + 1);
    const int n = end - beg;
    if (n <= 0) return;
    const float inv_n = 1.f / (float)n;
    const float* gp = gy + (int64_t)beg * C + c;
    const float* yp = y + (int64_t)beg * C + c;
    float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
    int i = 0;
    for (; i + 2 <= n; i += 2) {
        float g0 = __ldg(gp + (int64_t)i * C), g1 = __ldg(gp + (int64_t)(i + 1) * C);
        float y0 = __ldg(yp + (int64_t)i * C), y1 = __ldg(yp + (int64_t)(i + 1) * C);
        a0 += g0;
        a1 += g1;
        b0 = fmaf(g0, y0, b0);
        b1 = fmaf(g1, y1, b1);
    }
    for (; i < n; ++i) {
        float g0 = __ldg(gp + (int64_t)i * C), y0 = __ldg(yp + (int64_t)i * C);
        a0 += g0;
        b0 = fmaf(g0, y0, b0);
    }
    const float m1 = (a0 + a1) * inv_n, m2 = (b0 + b1) * inv_n;
    const float rs = __ldg(rstd + (int64_t)g * C + c);
    float* op = gx + (int64_t)beg * C + c;
    for (i = 0; i < n; ++i) {
        float g0 = __ldg(gp + (int64_t)i * C), y0 = __ldg(yp + (int64_t)i * C);
        op[(int64_t)i * C] = rs * (g0 - m1 - y0 * m2);
    }
}

}  // namespace

extern "C" int gsatb_segnorm_fwd(const float* x, const int32_t* seg_ptr, float* y, float* rstd, int64_t M, int64_t G,
                                 int C, float eps, gsatb_stream_t stream) {
    if (M < 0 || G < 0 || C <= 0) return GSATB_EINVAL;
    if (G == 0 || M == 0) return GSATB_OK;
    if (!x || !seg_ptr || !y || !rstd) return GSATB_EINVAL;
    if (G > 2147483647ll) return GSATB_ESHAPE;
    dim3 grid((unsigned)G, (unsigned)((C + SN_THREADS - 1) / SN_THREADS));
    k_segnorm_fwd<<<grid, SN_THREADS, 0, (cudaStream_t)stream>>>(x, seg_ptr, y, rstd, C, eps);
    GSATB_CHECK_LAUNCH();
    return GSATB_OK;
}

extern "C" int gsatb_segnorm_bwd(const float* gy, const float* y, const float* rstd, const int32_t* seg_ptr,
                                 float* gx, int64_t M, int64_t G, int C, gsatb_stream_t stream) {
    if (M < 0 || G < 0 || C <= 0) return GSATB_EINVAL;
    if (G == 0 || M == 0) return GSATB_OK;
    if (!gy || !y || !rstd || !seg_ptr || !gx) return GSATB_EINVAL;
    dim3 grid((unsigned)G, (unsigned)((C + SN_THREADS - 1) / SN_THREADS));
    k_segnorm_bwd<<<grid, SN_THREADS, 0, (cudaStream_t)stream>>>(gy, y, rstd, seg_ptr, gx, C);
    GSATB_CHECK_LAUNCH();
    return GSATB_OK;
}
