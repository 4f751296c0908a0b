// Self-test kernels for the host SIMT emulator (tests/simt/simt.h): known-answer checks of the primitives the kernel
// sources rely on -- __syncthreads with shared memory, full- and sub-warp shuffles, ballot, partial last warp,
// multi-dimensional grids.  TEST INFRASTRUCTURE ONLY.
#include "simt.h"

namespace {

// block-wide sum through shared memory + tree reduction (needs real barrier semantics between the steps)
__global__ void k_block_sum(const int* in, int* out, int n) {
    __shared__ int sh[256];
    const int tid = threadIdx.x;
    int v = 0;
    for (int i = blockIdx.x * blockDim.x + tid; i < n; i += gridDim.x * blockDim.x) v += in[i];
    sh[tid] = v;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (tid < s) sh[tid] += sh[tid + s];
        __syncthreads();
    }
    if (tid == 0) out[blockIdx.x] = sh[0];
}

// every sub-warp of `lpr` lanes reduces its own values with xor shuffles under its own mask; sub-warps run different
// trip counts (as the row walks of gine.cu / leconv.cu do)
__global__ void k_subwarp(const float* in, float* out, int lpr) {
    const int lane = threadIdx.x & 31, sub = lane / lpr;
    const unsigned submask = lpr == 32 ? 0xffffffffu : (((1u << lpr) - 1u) << (sub * lpr));
    const int group = (blockIdx.x * blockDim.x + threadIdx.x) / lpr;
    float total = 0.f;
    for (int it = 0; it <= group % 3; ++it) {                 // divergent trip counts between sub-warps
        float v = in[(blockIdx.x * blockDim.x + threadIdx.x)] * (float)(it + 1);
        for (int o = lpr / 2; o > 0; o >>= 1) v += __shfl_xor_sync(submask, v, o);
        total += v;
    }
    if (lane % lpr == 0) out[group] = total;
}

__global__ void k_ballot_partial(unsigned* out) {
    const unsigned b = __ballot_sync(0xffffffffu, (threadIdx.x % 3) == 0);      // 40 threads: last warp has 8 lanes
    const int up = __shfl_up_sync(0xffffffffu, (int)threadIdx.x, 1);
    const int dn = __shfl_down_sync(0xffffffffu, (int)threadIdx.x, 2);
    const int bc = __shfl_sync(0xffffffffu, (int)threadIdx.x, 5);
    out[threadIdx.x * 4 + 0] = b;
    out[threadIdx.x * 4 + 1] = (unsigned)up;
    out[threadIdx.x * 4 + 2] = (unsigned)dn;
    out[threadIdx.x * 4 + 3] = (unsigned)bc;
}

__global__ void k_grid3(int* out) {
    const int b = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    const int t = threadIdx.x + blockDim.x * threadIdx.y;
    atomicAdd(out + b, t);
}

}  // namespace

extern "C" int simt_selftest_block_sum(const int* in, int* out, int n, int blocks, int threads) {
    simt::launch(dim3(blocks), dim3(threads), [&]() { k_block_sum(in, out, n); });
    return 0;
}
extern "C" int simt_selftest_subwarp(const float* in, float* out, int lpr, int blocks, int threads) {
    simt::launch(dim3(blocks), dim3(threads), [&]() { k_subwarp(in, out, lpr); });
    return 0;
}
extern "C" int simt_selftest_ballot(unsigned* out) {
    simt::launch(dim3(1), dim3(40), [&]() { k_ballot_partial(out); });
    return 0;
}
extern "C" int simt_selftest_grid3(int* out) {
    simt::launch(dim3(2, 3, 2), dim3(4, 2), [&]() { k_grid3(out); });
    return 0;
}
