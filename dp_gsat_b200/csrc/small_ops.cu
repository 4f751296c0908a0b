// Weight / bias gradient of a Linear layer with a SMALL input width F (the node encoder Linear(x_dim, H) of
// reference src/models/gin.py:22-25 / pna.py:20-25: x_dim = 10 on BA-2Motifs, 14 / 31 on Mutag):
//      dW[h, f] = sum_n g[n, h] * x[n, f]        db[h] = sum_n g[n, h]
// The library route (autograd of addmm) runs an fp32 "large-K" sgemm over K = N rows at a few percent of the HBM
// roofline (3 ms at N = 4.9 M, H = 128); this kernel streams g once (4NH bytes) with 128-bit loads, keeps the
// [4 channels x (F+1)] partial products of a thread in registers and reduces the per-CTA partials in a fixed order
// (deterministic).  HBM bound: 4NH + 4NF bytes.
#include "common.cuh"

namespace {

constexpr int SDW_THREADS = 128;
constexpr int SDW_CHUNK = 64;      // rows of x staged per step
constexpr int SDW_FMAX = 16;       // F + 1 (bias column) <= SDW_FMAX

__global__ void __launch_bounds__(SDW_THREADS)
k_small_dw(const float4* __restrict__ g, const float* __restrict__ x, float* __restrict__ part, int64_t N, int HV, int F,
           int64_t rows_per_cta) {
    __shared__ float xs[SDW_CHUNK][SDW_FMAX];
    __shared__ float red[4][128][SDW_FMAX + 1];
    const int tid = threadIdx.x, cg = tid & 31, rsub = tid >> 5;
    const int c4 = blockIdx.y * 32 + cg;
    const bool c_ok = c4 < HV;
    const int FE = F + 1;
    float acc[4][SDW_FMAX];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int f = 0; f < SDW_FMAX; ++f) acc[k][f] = 0.f;
    const int64_t r_beg = blockIdx.x * rows_per_cta, r_end = min(N, r_beg + rows_per_cta);
    for (int64_t r0 = r_beg; r0 < r_end; r0 += SDW_CHUNK) {
        const int nr = (int)min((int64_t)SDW_CHUNK, r_end - r0);
        __syncthreads();
        for (int i = tid; i < SDW_CHUNK * SDW_FMAX; i += SDW_THREADS) {
            const int r = i / SDW_FMAX, f = i % SDW_FMAX;
            float v = 0.f;
            if (r < nr) v = f < F ? __ldg(x + (r0 + r) * F + f) : (f == F ? 1.f : 0.f);
            xs[r][f] = v;
        }
        __syncthreads();
#pragma unroll 4
        for (int r = rsub; r < SDW_CHUNK; r += 4) {
            float4 gv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < nr && c_ok) gv = ldg_stream_f4(g + (r0 + r) * HV + c4);
#pragma unroll
            for (int f = 0; f < SDW_FMAX; ++f) {
                const float xv = xs[r][f];
                acc[0][f] = fmaf(gv.x, xv, acc[0][f]);
                acc[1][f] = fmaf(gv.y, xv, acc[1][f]);
                acc[2][f] = fmaf(gv.z, xv, acc[2][f]);
                acc[3][f] = fmaf(gv.w, xv, acc[3][f]);
            }
        }
    }
    // fold the 4 row sub-groups (fixed order), then one partial [128 channels][F+1] per CTA
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int f = 0; f < SDW_FMAX; ++f) red[rsub][cg * 4 + k][f] = acc[k][f];
    __syncthreads();
    const int H = HV * 4;
    const int ch = blockIdx.y * 128 + tid;
    if (ch < H) {
        float* out = part + ((size_t)blockIdx.x * H + ch) * FE;
        for (int f = 0; f < FE; ++f) out[f] = (red[0][tid][f] + red[1][tid][f]) + (red[2][tid][f] + red[3][tid][f]);
    }
}

// dW[h, f] / db[h] = sum over CTAs (fixed order, fp64 accumulate) of the partials
__global__ void k_small_dw_reduce(const float* __restrict__ part, int parts, int H, int F, float* __restrict__ dW,
                                  float* __restrict__ db) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int FE = F + 1;
    if (i >= H * FE) return;
    double a = 0.0;
    for (int q = 0; q < parts; ++q) a += (double)part[(size_t)q * H * FE + i];
    const int h = i / FE, f = i % FE;
    if (f < F) dW[h * F + f] = (float)a;
    else if (db) db[h] = (float)a;
}

inline int sdw_parts(int64_t N) {
    int64_t p = (N + 1023) / 1024;
    const int64_t cap = (int64_t)GSATB_NUM_SMS * 8;
    return (int)(p < 1 ? 1 : (p > cap ? cap : p));
}

}  // namespace

extern "C" size_t gsatb_linear_small_dw_workspace(int64_t N, int H, int F) {
    return (size_t)sdw_parts(N) * (size_t)H * (size_t)(F + 1) * sizeof(float) + 256;
}

extern "C" int gsatb_linear_small_dw(const float* g, const float* x, float* dW, float* db, int64_t N, int H, int F,
                                     void* ws, size_t ws_bytes, gsatb_stream_t stream) {
    if (N < 0 || H <= 0 || F <= 0) return GSATB_EINVAL;
    if (!dW || !ws || (N > 0 && (!g || !x))) return GSATB_EINVAL;
    if (H % 4 != 0 || F + 1 > SDW_FMAX) return GSATB_ESHAPE;
    if (!gsatb_aligned16(g)) return GSATB_EALIGN;
    if (ws_bytes < gsatb_linear_small_dw_workspace(N, H, F)) return GSATB_EWS_TOO_SMALL;
    cudaStream_t st = (cudaStream_t)stream;
    const int parts = sdw_parts(N);
    const int64_t rows_per_cta = ((N + parts - 1) / parts + SDW_CHUNK - 1) / SDW_CHUNK * SDW_CHUNK;
    dim3 grid((unsigned)parts, (unsigned)((H + 127) / 128));
    k_small_dw<<<grid, SDW_THREADS, 0, st>>>((const float4*)g, x, (float*)ws, N, H / 4, F, rows_per_cta > 0 ? rows_per_cta : SDW_CHUNK);
    k_small_dw_reduce<<<(H * (F + 1) + 127) / 128, 128, 0, st>>>((const float*)ws, parts, H, F, dW, db);
    GSATB_CHECK_LAUNCH();
    return GSATB_OK;
}
