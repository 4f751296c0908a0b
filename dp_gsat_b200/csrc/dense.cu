// Streaming helpers of the dense layers (nn.Linear / BatchNorm1d of the reference: src/models/gin.py:22-25,55-62,42;
// src/models/pna.py:20-50; src/utils/get_model.py:57-68; src/models/conv_layers.py:149,164) that sit around the tcgen05
// GEMM kernels.  All HBM bound, all deterministic (fixed-order reductions, no atomics).
//
//  gsatb_split_bf16      fp32 [rows, C] -> bf16 operand of a tensor-core GEMM.  nseg = 1: plain rounding (bf16 mode).
//                        nseg = 6: the "split-bf16 x3" strict mode -- every fp32 value is written as the exact sum
//                        h + m + l of three bf16 numbers (24 mantissa bits), and a product x * w is evaluated as the six
//                        significant partial products  l h + m m + h l + m h + h m + h h  (the dropped ones are below
//                        2^-24 relative) by laying the parts out as SIX consecutive K-segments of one ordinary bf16 GEMM
//                        with fp32 accumulation:  A side (l,m,h,m,h,h), B side (h,m,l,h,m,h): smallest products first.  The tcgen05 kernels run
//                        unchanged with K' = 6 K and deliver fp32-accurate products (rtol 1e-5 parity mode).
//                        layout 0 concatenates the segments along the feature axis ([rows, nseg*C]: forward and
//                        input-gradient GEMMs, contraction over features); layout 1 stacks them along the row axis
//                        ([nseg*rows, C]: weight-gradient GEMM, contraction over rows).
//  gsatb_colsum          out[c] = sum_r x[r, c]  (bias gradients), fp64 accumulation in a fixed order
//  gsatb_bn_stats        BatchNorm1d training statistics: shifted sums in fp64 -> mean, rstd; running-statistics update
//  gsatb_bn_apply        y = (x - mean) * rstd * gamma + beta  (+ ReLU)
//  gsatb_bn_bwd_stats    sum dy, sum dy * xhat  (the ReLU gate taken from y when fused)
//  gsatb_bn_bwd_apply    dx = gamma * rstd * (g - sum(g)/n - xhat * sum(g * xhat)/n)   (training) | gamma * rstd * g (eval)
#include <cuda_bf16.h>
#include "common.cuh"

namespace {

__device__ __forceinline__ uint16_t bf16_bits(float f) {
    __nv_bfloat16 h = __float2bfloat16_rn(f);
    return *reinterpret_cast<uint16_t*>(&h);
}
__device__ __forceinline__ float bf16_val(uint16_t b) { return __uint_as_float((uint32_t)b << 16); }

// part 0 / 1 / 2 of the three-way bf16 split of v (h = rn(v), m = rn(v - h), l = rn(v - h - m); the subtractions are exact)
__device__ __forceinline__ uint16_t split_part(float v, int part) {
    const uint16_t h = bf16_bits(v);
    if (part == 0) return h;
    const float r1 = v - bf16_val(h);
    const uint16_t m = bf16_bits(r1);
    if (part == 1) return m;
    return bf16_bits(r1 - bf16_val(m));
}

// Fast path (C % 8 == 0, 16-byte aligned rows): one thread = 8 consecutive INPUT elements of a row -- two 128-bit loads,
// the three parts computed once, one 128-bit store per segment.  Every input element is read once however many segments
// are written (the generic path below re-reads and re-splits it per segment and divides by C per element).
__global__ void __launch_bounds__(256)
k_split_bf16_vec(const float* __restrict__ x, int64_t rows, int C, int64_t ldx, int nseg, uint32_t pattern, int layout,
                 uint16_t* __restrict__ out, int64_t ld_out) {
    const int cpr = C >> 3;                                  // 8-element chunks per input row
    const int64_t total = rows * cpr;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cpr;
        const int c0 = (int)(i - r * cpr) << 3;
        const float4 a = *reinterpret_cast<const float4*>(x + r * ldx + c0);
        const float4 b = *reinterpret_cast<const float4*>(x + r * ldx + c0 + 4);
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint16_t part[3][8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint16_t h = bf16_bits(v[k]);
            const float r1 = v[k] - bf16_val(h);
            const uint16_t m = bf16_bits(r1);
            part[0][k] = h;
            part[1][k] = m;
            part[2][k] = bf16_bits(r1 - bf16_val(m));
        }
        for (int s = 0; s < nseg; ++s) {
            const int ps = (int)((pattern >> (2 * s)) & 3u);
            const uint16_t* o = ps == 0 ? part[0] : (ps == 1 ? part[1] : part[2]);
            uint4 q;
            q.x = (uint32_t)o[0] | ((uint32_t)o[1] << 16);
            q.y = (uint32_t)o[2] | ((uint32_t)o[3] << 16);
            q.z = (uint32_t)o[4] | ((uint32_t)o[5] << 16);
            q.w = (uint32_t)o[6] | ((uint32_t)o[7] << 16);
            uint16_t* dst = layout == 0 ? out + r * ld_out + (int64_t)s * C + c0 : out + ((int64_t)s * rows + r) * ld_out + c0;
            *reinterpret_cast<uint4*>(dst) = q;
        }
    }
}
// zero the padding columns [width, ld_out) of every output row (fast path only; width % 8 == 0)
__global__ void __launch_bounds__(256)
k_split_pad_zero(uint16_t* __restrict__ out, int64_t out_rows, int width, int64_t ld_out) {
    const int cpr = (int)((ld_out - width) >> 3);
    const int64_t total = out_rows * cpr;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cpr;
        const int c0 = width + ((int)(i - r * cpr) << 3);
        *reinterpret_cast<uint4*>(out + r * ld_out + c0) = make_uint4(0u, 0u, 0u, 0u);
    }
}

// generic path: one thread = 8 consecutive output elements of one output row (one 16-byte store)
__global__ void __launch_bounds__(256)
k_split_bf16(const float* __restrict__ x, int64_t rows, int C, int64_t ldx, int nseg, uint32_t pattern, int layout,
             uint16_t* __restrict__ out, int64_t ld_out) {
    const int64_t out_rows = layout == 0 ? rows : rows * nseg;
    const int64_t chunks_per_row = ld_out / 8;
    const int64_t total = out_rows * chunks_per_row;
    const int width = layout == 0 ? nseg * C : C;          // valid columns of an output row
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t orow = i / chunks_per_row;
        const int j0 = (int)(i % chunks_per_row) * 8;
        int64_t r = orow;
        int seg_row = 0;
        if (layout == 1) {
            seg_row = (int)(orow / rows);
            r = orow % rows;
        }
        uint16_t o[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int j = j0 + k;
            uint16_t v = 0;
            if (j < width) {
                const int s = layout == 0 ? j / C : seg_row;
                const int c = layout == 0 ? j - s * C : j;
                v = split_part(x[r * ldx + c], (int)((pattern >> (2 * s)) & 3u));
            }
            o[k] = v;
        }
        uint4 q;
        q.x = (uint32_t)o[0] | ((uint32_t)o[1] << 16);
        q.y = (uint32_t)o[2] | ((uint32_t)o[3] << 16);
        q.z = (uint32_t)o[4] | ((uint32_t)o[5] << 16);
        q.w = (uint32_t)o[6] | ((uint32_t)o[7] << 16);
        *reinterpret_cast<uint4*>(out + orow * ld_out + j0) = q;
    }
}

// ---- column reductions ------------------------------------------------------------------------------------------
// Block b owns rows [b * rpb, (b+1) * rpb); thread (ty, tx): column tx + 32 * blockIdx.y ... walks rows ty, ty + 8, ...
// Per-block partial [NV][C] in fp64, folded over the 8 row lanes in a fixed order; a second kernel adds the blocks'
// partials in block order.  MODE selects the summands:
//   0: x                                  (NV = 1)   bias gradient
//   1: x - shift, (x - shift)^2           (NV = 2)   BatchNorm statistics, shift = row 0 (kills the cancellation)
//   2: g, g * xhat                        (NV = 2)   BatchNorm backward sums, g = dy (* [y > 0] when the ReLU is fused)
constexpr int CR_TX = 32, CR_TY = 8;

template <int MODE>
__global__ void __launch_bounds__(CR_TX * CR_TY)
k_col_partials(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ y,
               const float* __restrict__ mean, const float* __restrict__ rstd, int64_t rows, int C, int64_t ld,
               int64_t rpb, double* __restrict__ part) {
    constexpr int NV = MODE == 0 ? 1 : 2;
    __shared__ double red[CR_TY][NV][CR_TX];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int c = blockIdx.y * CR_TX + tx;
    const bool ok = c < C;
    const int64_t r0 = (int64_t)blockIdx.x * rpb, r1 = min(rows, r0 + rpb);
    double s0 = 0.0, s1 = 0.0;
    float sh = 0.f, mu = 0.f, rs = 0.f;
    if (ok) {
        if (MODE == 1) sh = a[c];
        if (MODE == 2) {
            mu = mean[c];
            rs = rstd[c];
        }
    }
    if (ok) {
        // fp32 partial sums over short runs of 16 rows, promoted to fp64: error of a run ~1e-7 relative, no growth with rows
        for (int64_t rb = r0 + ty * 16; rb < r1; rb += CR_TY * 16) {
            float p0 = 0.f, p1 = 0.f;
            const int64_t re = min(r1, rb + 16);
            for (int64_t r = rb; r < re; ++r) {
                const float v = a[r * ld + c];
                if (MODE == 0) {
                    p0 += v;
                } else if (MODE == 1) {
                    const float d = v - sh;
                    p0 += d;
                    p1 = fmaf(d, d, p1);
                } else {
                    float g = v;
                    if (y && !(y[r * ld + c] > 0.f)) g = 0.f;
                    const float xh = (b[r * ld + c] - mu) * rs;
                    p0 += g;
                    p1 = fmaf(g, xh, p1);
                }
            }
            s0 += (double)p0;
            s1 += (double)p1;
        }
    }
    red[ty][0][tx] = s0;
    if (NV == 2) red[ty][NV - 1][tx] = s1;
    __syncthreads();
    if (ty == 0 && ok) {
        double t0 = 0.0, t1 = 0.0;
        for (int k = 0; k < CR_TY; ++k) {
            t0 += red[k][0][tx];
            if (NV == 2) t1 += red[k][NV - 1][tx];
        }
        part[((size_t)blockIdx.x * NV + 0) * C + c] = t0;
        if (NV == 2) part[((size_t)blockIdx.x * NV + 1) * C + c] = t1;
    }
}

// MODE 0: out_f[c] = sum.   MODE 2: out_f[c] = sum g (d beta), out_f[C + c] = sum g xhat (d gamma); sums64 gets both.
// MODE 1: mean / rstd of the batch, optional running-statistics update (momentum, unbiased variance); sums64 (nullable)
//         receives the raw [sum d, sum d^2] and is used by the data-parallel statistics exchange.
template <int MODE>
__global__ void k_col_finish(const double* __restrict__ part, int nblocks, int C, int64_t rows, const float* __restrict__ x0,
                             float eps, float momentum, float* __restrict__ running_mean, float* __restrict__ running_var,
                             float* __restrict__ out0, float* __restrict__ out1, double* __restrict__ sums64) {
    constexpr int NV = MODE == 0 ? 1 : 2;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double t0 = 0.0, t1 = 0.0;
    for (int b = 0; b < nblocks; ++b) {
        t0 += part[((size_t)b * NV + 0) * C + c];
        if (NV == 2) t1 += part[((size_t)b * NV + 1) * C + c];
    }
    if (sums64) {
        sums64[c] = t0;
        if (NV == 2) sums64[C + c] = t1;
    }
    if (MODE == 0) {
        out0[c] = (float)t0;
    } else if (MODE == 2) {
        out0[c] = (float)t0;
        out1[c] = (float)t1;
    } else {
        const double n = (double)rows;
        const double md = t0 / n;
        double var = t1 / n - md * md;                       // biased, on shifted data: no catastrophic cancellation
        if (var < 0.0) var = 0.0;
        const double mean = (double)x0[c] + md;
        out0[c] = (float)mean;
        out1[c] = (float)(1.0 / sqrt(var + (double)eps));
        if (running_mean) {
            const double unb = rows > 1 ? var * n / (n - 1.0) : var;
            running_mean[c] = (float)((1.0 - (double)momentum) * (double)running_mean[c] + (double)momentum * mean);
            running_var[c] = (float)((1.0 - (double)momentum) * (double)running_var[c] + (double)momentum * unb);
        }
    }
}

inline int col_blocks(int64_t rows) {
    int64_t b = (rows + 1023) / 1024;
    if (b > 4 * GSATB_NUM_SMS) b = 4 * GSATB_NUM_SMS;
    return b < 1 ? 1 : (int)b;
}

template <int MODE>
int run_col(const float* a, const float* b, const float* y, const float* mean, const float* rstd, int64_t rows, int C,
            int64_t ld, float eps, float momentum, float* rm, float* rv, float* out0, float* out1, double* sums64,
            void* ws, size_t ws_bytes, cudaStream_t st) {
    constexpr int NV = MODE == 0 ? 1 : 2;
    const int nb = col_blocks(rows);
    if (ws_bytes < (size_t)nb * NV * C * sizeof(double)) return GSATB_EWS_TOO_SMALL;
    const int64_t rpb = (rows + nb - 1) / nb;
    dim3 grid(nb, (C + CR_TX - 1) / CR_TX), block(CR_TX, CR_TY);
    k_col_partials<MODE><<<grid, block, 0, st>>>(a, b, y, mean, rstd, rows, C, ld, rpb, (double*)ws);
    GSATB_CHECK_LAUNCH();
    k_col_finish<MODE><<<(C + 127) / 128, 128, 0, st>>>((const double*)ws, nb, C, rows, a, eps, momentum, rm, rv, out0, out1,
                                                         sums64);
    GSATB_CHECK_LAUNCH();
    return GSATB_OK;
}

// y = (x - mean) * rstd * gamma + beta (+ ReLU): 4 channels per thread when C % 4 == 0
__global__ void __launch_bounds__(256)
k_bn_apply(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ rstd,
           const float* __restrict__ gamma, const float* __restrict__ beta, int relu, float* __restrict__ y, int64_t rows,
           int C) {
    const int64_t total = rows * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const float sc = rstd[c] * (gamma ? gamma[c] : 1.f);
        float v = fmaf(x[i] - mean[c], sc, beta ? beta[c] : 0.f);
        if (relu) v = fmaxf(v, 0.f);
        y[i] = v;
    }
}

__global__ void __launch_bounds__(256)
k_bn_bwd_apply(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ y,
               const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
               const float* __restrict__ sum_g, const float* __restrict__ sum_gx, float inv_n, int training,
               float* __restrict__ dx, int64_t rows, int C) {
    const int64_t total = rows * C;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        float g = dy[i];
        if (y && !(y[i] > 0.f)) g = 0.f;
        const float rs = rstd[c];
        const float coef = rs * (gamma ? gamma[c] : 1.f);
        float v = g;
        if (training) v = g - sum_g[c] * inv_n - (x[i] - mean[c]) * rs * (sum_gx[c] * inv_n);
        dx[i] = coef * v;
    }
}

inline unsigned stream_grid(int64_t total) {
    int64_t b = (total + 255) / 256;
    if (b > (int64_t)GSATB_NUM_SMS * 16) b = (int64_t)GSATB_NUM_SMS * 16;
    return (unsigned)(b < 1 ? 1 : b);
}

}  // namespace

extern "C" int gsatb_split_bf16(const float* x, int64_t rows, int C, int64_t ldx, int nseg, int pattern, int layout,
                                void* out_bf16, int64_t ld_out, gsatb_stream_t stream) {
    if (rows < 0 || C <= 0 || nseg < 1 || nseg > 8 || (layout != 0 && layout != 1) || ldx < C) return GSATB_EINVAL;
    if (rows == 0) return GSATB_OK;
    if (!x || !out_bf16) return GSATB_EINVAL;
    const int64_t width = layout == 0 ? (int64_t)nseg * C : C;
    if (ld_out % 8 != 0 || ld_out < width) return GSATB_ESHAPE;
    if (!gsatb_aligned16(out_bf16)) return GSATB_EALIGN;
    for (int s = 0; s < nseg; ++s)
        if ((((uint32_t)pattern >> (2 * s)) & 3u) > 2u) return GSATB_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    if (C % 8 == 0 && ldx % 4 == 0 && gsatb_aligned16(x)) {
        const int64_t out_rows = layout == 0 ? rows : rows * nseg;
        k_split_bf16_vec<<<stream_grid(rows * (C / 8)), 256, 0, st>>>(x, rows, C, ldx, nseg, (uint32_t)pattern, layout,
                                                                    (uint16_t*)out_bf16, ld_out);
        GSATB_CHECK_LAUNCH();
        if (ld_out > width) {
            k_split_pad_zero<<<stream_grid(out_rows * ((ld_out - width) / 8)), 256, 0, st>>>((uint16_t*)out_bf16, out_rows,
                                                                                            (int)width, ld_out);
            GSATB_CHECK_LAUNCH();
        }
        return GSATB_OK;
    }
    const int64_t total = (layout == 0 ? rows : rows * nseg) * (ld_out / 8);
    k_split_bf16<<<stream_grid(total), 256, 0, st>>>(x, rows, C, ldx, nseg, (uint32_t)pattern, layout,
                                                     (uint16_t*)out_bf16, ld_out);
    GSATB_CHECK_LAUNCH();
    return GSATB_OK;
}

extern "C" size_t gsatb_col_workspace(int64_t rows, int C) {
    if (rows < 0 || C <= 0) return 0;
    return (size_t)col_blocks(rows) * 2 * C * sizeof(double);
}

extern "C" int gsatb_colsum(const float* x, int64_t rows, int C, int64_t ld, float* out, void* workspace, size_t ws_bytes,
                            gsatb_stream_t stream) {
    if (rows < 0 || C <= 0 || ld < C) return GSATB_EINVAL;
    if (!out) return GSATB_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    if (rows == 0) return cudaMemsetAsync(out, 0, (size_t)C * 4, st) == cudaSuccess ? GSATB_OK : GSATB_ELAUNCH;
    if (!x || !workspace) return GSATB_EINVAL;
    return run_col<0>(x, nullptr, nullptr, nullptr, nullptr, rows, C, ld, 0.f, 0.f, nullptr, nullptr, out, nullptr, nullptr,
                      workspace, ws_bytes, st);
}

extern "C" int gsatb_bn_stats(const float* x, int64_t rows, int C, float eps, float momentum, float* running_mean,
                              float* running_var, float* mean, float* rstd, double* sums64, void* workspace,
                              size_t ws_bytes, gsatb_stream_t stream) {
    if (rows <= 0 || C <= 0) return GSATB_EINVAL;
    if (!x || !mean || !rstd || !workspace) return GSATB_EINVAL;
    if ((running_mean == nullptr) != (running_var == nullptr)) return GSATB_EINVAL;
    return run_col<1>(x, nullptr, nullptr, nullptr, nullptr, rows, C, C, eps, momentum, running_mean, running_var, mean,
                      rstd, sums64, workspace, ws_bytes, (cudaStream_t)stream);
}

extern "C" int gsatb_bn_apply(const float* x, const float* mean, const float* rstd, const float* gamma, const float* beta,
                              int relu, float* y, int64_t rows, int C, gsatb_stream_t stream) {
    if (rows < 0 || C <= 0) return GSATB_EINVAL;
    if (rows == 0) return GSATB_OK;
    if (!x || !mean || !rstd || !y) return GSATB_EINVAL;
    k_bn_apply<<<stream_grid(rows * C), 256, 0, (cudaStream_t)stream>>>(x, mean, rstd, gamma, beta, relu, y, rows, C);
    GSATB_CHECK_LAUNCH();
    return GSATB_OK;
}

extern "C" int gsatb_bn_bwd_stats(const float* dy, const float* x, const float* y_relu, const float* mean, const float* rstd,
                                  int64_t rows, int C, float* dbeta, float* dgamma, double* sums64, void* workspace,
                                  size_t ws_bytes, gsatb_stream_t stream) {
    if (rows <= 0 || C <= 0) return GSATB_EINVAL;
    if (!dy || !x || !mean || !rstd || !dbeta || !dgamma || !workspace) return GSATB_EINVAL;
    return run_col<2>(dy, x, y_relu, mean, rstd, rows, C, C, 0.f, 0.f, nullptr, nullptr, dbeta, dgamma, sums64, workspace,
                      ws_bytes, (cudaStream_t)stream);
}

extern "C" int gsatb_bn_bwd_apply(const float* dy, const float* x, const float* y_relu, const float* mean, const float* rstd,
                                  const float* gamma, const float* sum_g, const float* sum_gx, float inv_n, int training,
                                  float* dx, int64_t rows, int C, gsatb_stream_t stream) {
    if (rows < 0 || C <= 0) return GSATB_EINVAL;
    if (rows == 0) return GSATB_OK;
    if (!dy || !x || !mean || !rstd || !dx) return GSATB_EINVAL;
    if (training && (!sum_g || !sum_gx)) return GSATB_EINVAL;
    k_bn_bwd_apply<<<stream_grid(rows * C), 256, 0, (cudaStream_t)stream>>>(dy, x, y_relu, mean, rstd, gamma, sum_g, sum_gx,
                                                                           inv_n, training, dx, rows, C);
    GSATB_CHECK_LAUNCH();
    return GSATB_OK;
}

// ---- BatchNorm folding for the tensor-core GIN layer (tc.py): one launch instead of ~15 elementwise library launches ----
namespace {

__global__ void k_bn_fold_fwd(const double* __restrict__ stats, double n_host, const double* __restrict__ n_dev,
                              const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float momentum,
                              float* __restrict__ running_mean, float* __restrict__ running_var,
                              long long* __restrict__ nbt, int training, float* __restrict__ mean, float* __restrict__ rstd,
                              float* __restrict__ scale, float* __restrict__ shift, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0 && training && nbt) *nbt += 1;
    if (c >= C) return;
    float mu, var;
    if (training) {
        const double n = n_dev ? *n_dev : n_host;
        const double m = stats[c] / n;
        double v = stats[C + c] / n - m * m;                  // biased variance, from the fp32 accumulators' fp64 sums
        if (v < 0.0) v = 0.0;
        mu = (float)m;
        var = (float)v;
        if (running_mean) {
            const double nm1 = n - 1.0 > 1.0 ? n - 1.0 : 1.0;
            const float unb = (float)(n / nm1);
            running_mean[c] = running_mean[c] * (1.f - momentum) + momentum * mu;
            running_var[c] = running_var[c] * (1.f - momentum) + momentum * (var * unb);
        }
    } else {
        mu = running_mean[c];
        var = running_var[c];
    }
    const float rs = 1.f / sqrtf(var + eps);
    const float sc = (gamma ? gamma[c] : 1.f) * rs;
    mean[c] = mu;
    rstd[c] = rs;
    scale[c] = sc;
    shift[c] = (beta ? beta[c] : 0.f) - mu * sc;
}

// dz1 = cA * g + cB * z1 + cC  with  coef = gamma * rstd:  training: cA = coef, cB = -coef rstd mean(g xhat),
// cC = -coef mean(g) - cB mean;   eval (running statistics are constants): cA = coef, cB = cC = 0
__global__ void k_bn_fold_bwd(const float* __restrict__ sum_g, const float* __restrict__ sum_gx, double n_host,
                              const double* __restrict__ n_dev, const float* __restrict__ gamma,
                              const float* __restrict__ mean, const float* __restrict__ rstd, int training,
                              float* __restrict__ cA, float* __restrict__ cB, float* __restrict__ cC, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float coef = (gamma ? gamma[c] : 1.f) * rstd[c];
    float b = 0.f, cc = 0.f;
    if (training) {
        const float inv_n = (float)(1.0 / (n_dev ? *n_dev : n_host));
        b = -coef * rstd[c] * (sum_gx[c] * inv_n);
        cc = -coef * (sum_g[c] * inv_n) - b * mean[c];
    }
    cA[c] = coef;
    cB[c] = b;
    cC[c] = cc;
}

}  // namespace

extern "C" int gsatb_bn_fold_fwd(const double* stats, double n_host, const double* n_dev, const float* gamma, const float* beta,
                                 float eps, float momentum, float* running_mean, float* running_var,
                                 int64_t* num_batches_tracked, int training, float* mean, float* rstd, float* scale,
                                 float* shift, int C, gsatb_stream_t stream) {
    if (C <= 0 || !mean || !rstd || !scale || !shift) return GSATB_EINVAL;
    if (training && (!stats || (!n_dev && !(n_host > 0.0)))) return GSATB_EINVAL;
    if (!training && (!running_mean || !running_var)) return GSATB_EINVAL;
    if ((running_mean == nullptr) != (running_var == nullptr)) return GSATB_EINVAL;
    k_bn_fold_fwd<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(stats, n_host, n_dev, gamma, beta, eps, momentum,
                                                                     running_mean, running_var,
                                                                     (long long*)num_batches_tracked, training, mean, rstd,
                                                                     scale, shift, C);
    GSATB_CHECK_LAUNCH();
    return GSATB_OK;
}

extern "C" int gsatb_bn_fold_bwd(const float* sum_g, const float* sum_gx, double n_host, const double* n_dev,
                                 const float* gamma, const float* mean, const float* rstd, int training, float* cA, float* cB,
                                 float* cC, int C, gsatb_stream_t stream) {
    if (C <= 0 || !rstd || !cA || !cB || !cC) return GSATB_EINVAL;
    if (training && (!sum_g || !sum_gx || !mean || (!n_dev && !(n_host > 0.0)))) return GSATB_EINVAL;
    k_bn_fold_bwd<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sum_g, sum_gx, n_host, n_dev, gamma, mean, rstd, training,
                                                                     cA, cB, cC, C);
    GSATB_CHECK_LAUNCH();
    return GSATB_OK;
}
