// Weight-gradient products on the tensor cores:  D[m, n] = sum_r A[r, m] * B[r, n]   (dW = dz^T h, K = rows), plus the
// bias gradient  db[m] = sum_r A[r, m]  from the same pass.  Replaces the autograd weight-gradient GEMMs of
// src/utils/get_model.py:57-68 (extractor MLP) and src/models/gin.py:55-62 (GIN node MLP) at src/run_gsat.py:634,
// which round 1 ran as library GEMMs.
//
// Both operands are bf16 activations as the other kernels of the step leave them in HBM; each may be
//   row-major     [rows, C]  (ld = elements per row)      -> TMA boxes of 64 channels x 64 rows = an MN-major
//                                                            SWIZZLE_128B operand tile (K = rows runs across smem rows)
//   channel-major [C, rows]  (ld = elements per channel)  -> TMA boxes of 64 rows x 128 channels = the usual K-major tile
//   tile-major    [rows / 128][ld channels][128 rows]       -> the same boxes out of one contiguous block per 128 rows (the
//                                                            slot space of the fused extractor kernels, layout code 2)
// so no transposed copy of an activation is ever made.
//
// Split-K over the rows: grid = (splits, slabs); a slab is one 128-channel block of A times up to 256 channels of B
// (accumulator [128 lanes x n_chunk columns] in TMEM, + 16 columns for the bias gradient: A times a tile of ones).
// Warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, warps 4-7 = epilogue (accumulator -> fp32
// partial in the workspace).  A second tiny kernel adds the partials in split order: deterministic.
#include "tc_ops_common.cuh"

namespace {

using namespace tcg;

constexpr int DW_THREADS = 256;
constexpr int DW_KSTEP = 64;                 // rows per pipeline stage
constexpr int DW_A_BYTES = 128 * 128;        // [128 channels x 64 rows] bf16
constexpr int DW_ONES_BYTES = 16 * 128;      // 16 "channels" of ones x 64 rows (any layout of ones is ones)
constexpr int DW_BIAS_COLS = 16;

struct DwParams {
    int64_t rows;
    int M, N;            // channels of A (output rows of D) and of B (output columns of D)
    int a_cm, b_cm;      // 1: channel-major operand, 2: tile-major (channel-major inside blocks of 128 rows)
    int lda, ldb;        // tile-major operands: channel rows per 128-row block
    int n_chunk;         // B channels per slab: multiple of 64, <= 256
    int n_chunks;        // slabs per A block
    int slabs;           // A blocks x n_chunks
    int mpc;             // A blocks per CTA: 2 when no bias gradient is wanted and two accumulators fit in TMEM -- every B
                         // tile then feeds two MMAs per fetch (the 4-block dW1 of the extractor pulled 38 GB through L2
                         // per launch, 25 GB with pairs: 3.8-4.4 -> 3.1 ms at cfg4)
    int splits;
    int stages;
    float* ws;           // [splits][slabs][128][n_chunk + 16]
};

__global__ void __launch_bounds__(DW_THREADS, 1)
k_tc_dw(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const DwParams p) {
#ifdef GSATB_HOST_SIM
    uint8_t* smem_raw = simt::dyn_smem();
#else
    extern __shared__ uint8_t smem_raw[];
#endif
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int b_bytes = p.n_chunk * 128;
    const int stage_bytes = p.mpc * DW_A_BYTES + b_bytes;
    uint8_t* ones = smem + (size_t)p.stages * stage_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ones + DW_ONES_BYTES);
    uint64_t* full = bars;              // [stages <= 8]
    uint64_t* empty = bars + 8;         // [stages]
    uint64_t* acc_full = bars + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int split = blockIdx.x;
    // CTA (split, y): mpc == 1: slab y = (A block y / n_chunks, B chunk y % n_chunks); mpc == 2 (n_chunks == 1): A blocks 2y, 2y + 1
    const int mb = p.mpc == 2 ? 2 * (int)blockIdx.y : (int)blockIdx.y / p.n_chunks;
    const int nc = p.mpc == 2 ? 0 : (int)blockIdx.y % p.n_chunks;
    const int64_t total_steps = (p.rows + DW_KSTEP - 1) / DW_KSTEP;
    // K-steps are dealt round-robin over the splits (split s takes steps s, s + splits, ...): at any moment the CTAs of a
    // slab stream ADJACENT 64-row slices of the operands, so a channel-major operand ([C, rows]: one 128-byte piece per
    // channel row and step) is read as long contiguous runs per DRAM page instead of 128-byte pieces ~rows/splits apart.
    const int64_t s0 = split, s1 = total_steps, sstep = p.splits;

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&tmap_a);
        tc::tma_prefetch_desc(&tmap_b);
        for (int i = 0; i < p.stages; ++i) {
            tc::mbar_init(&full[i], 1);
            tc::mbar_init(&empty[i], 1);
        }
        tc::mbar_init(acc_full, 1);
        tc::fence_barrier_init();
    }
    if (warp == 2) {
        tc::tmem_alloc(tmem_slot, 512);
        tc::tmem_relinquish();
    }
    if (warp == 3) {      // the tile of ones behind the bias gradient
        uint32_t* o = reinterpret_cast<uint32_t*>(ones);
        for (int i = lane; i < DW_ONES_BYTES / 4; i += 32) o[i] = 0x3F803F80u;      // bf16 1.0 | 1.0
        tc::fence_proxy_async_smem();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int64_t st = s0; st < s1; st += sstep, ++it) {
                const uint32_t s = it % p.stages, use = it / p.stages;
                tc::mbar_wait(&empty[s], (use & 1) ^ 1);
                tc::mbar_arrive_expect_tx(&full[s], (uint32_t)stage_bytes);
                uint8_t* sA0 = smem + (size_t)s * stage_bytes;
                uint8_t* sB = sA0 + p.mpc * DW_A_BYTES;
                const int r = (int)(st * DW_KSTEP);
                for (int ja = 0; ja < p.mpc; ++ja) {
                    uint8_t* sA = sA0 + ja * DW_A_BYTES;
                    const int mbj = mb + ja;
                    if (p.a_cm == 2) {
                        tc::tma_load_2d(sA, &tmap_a, &full[s], r & 127, (r >> 7) * p.lda + mbj * 128);
                    } else if (p.a_cm) {
                        tc::tma_load_2d(sA, &tmap_a, &full[s], r, mbj * 128);                   // box {64 rows, 128 ch}
                    } else {
                        tc::tma_load_2d(sA, &tmap_a, &full[s], mbj * 128, r);                   // box {64 ch, 64 rows}
                        tc::tma_load_2d(sA + 8192, &tmap_a, &full[s], mbj * 128 + 64, r);
                    }
                }
                if (p.b_cm == 2) {
                    for (int j = 0; j * 128 < p.n_chunk; ++j)
                        tc::tma_load_2d(sB + j * 16384, &tmap_b, &full[s], r & 127, (r >> 7) * p.ldb + nc * p.n_chunk + j * 128);
                } else if (p.b_cm) {
                    for (int j = 0; j * 128 < p.n_chunk; ++j)
                        tc::tma_load_2d(sB + j * 16384, &tmap_b, &full[s], r, nc * p.n_chunk + j * 128);
                } else {
                    for (int j = 0; j * 64 < p.n_chunk; ++j)
                        tc::tma_load_2d(sB + j * 8192, &tmap_b, &full[s], nc * p.n_chunk + j * 64, r);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = tc::make_idesc_bf16(128, p.n_chunk, p.a_cm ? 0 : 1, p.b_cm ? 0 : 1);
            const uint32_t idesc_b = tc::make_idesc_bf16(128, DW_BIAS_COLS, p.a_cm ? 0 : 1, 0);
            const uint64_t ones_desc = tc::make_desc_k_sw128(tc::smem_u32(ones));
            uint32_t it = 0;
            for (int64_t st = s0; st < s1; st += sstep, ++it) {
                const uint32_t s = it % p.stages, use = it / p.stages;
                tc::mbar_wait(&full[s], use & 1);
                tc::tc_fence_after();
                const uint32_t a_addr = tc::smem_u32(smem + (size_t)s * stage_bytes);
                const uint32_t b_addr = a_addr + p.mpc * DW_A_BYTES;
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {          // 4 x (K = 16 rows)
                    const uint64_t a_desc = p.a_cm ? tc::make_desc_k_sw128(a_addr) + (uint64_t)(k4 * 2)
                                                   : tc::make_desc_mn_sw128(a_addr + k4 * 2048, 8192);
                    const uint64_t b_desc = p.b_cm ? tc::make_desc_k_sw128(b_addr) + (uint64_t)(k4 * 2)
                                                   : tc::make_desc_mn_sw128(b_addr + k4 * 2048, 8192);
                    const uint32_t accum = (it | (uint32_t)k4) != 0;
                    tc::mma_bf16_ss(tmem_base, a_desc, b_desc, idesc, accum);
                    if (p.mpc == 2) {                     // second A block of the pair on the same B tile (no bias columns)
                        const uint64_t a1_desc = p.a_cm ? tc::make_desc_k_sw128(a_addr + DW_A_BYTES) + (uint64_t)(k4 * 2)
                                                        : tc::make_desc_mn_sw128(a_addr + DW_A_BYTES + k4 * 2048, 8192);
                        tc::mma_bf16_ss(tmem_base + p.n_chunk, a1_desc, b_desc, idesc, accum);
                    } else {
                        tc::mma_bf16_ss(tmem_base + p.n_chunk, a_desc, ones_desc + (uint64_t)(k4 * 2), idesc_b, accum);
                    }
                }
                tc::mma_commit(&empty[s]);
            }
            tc::mma_commit(acc_full);
        }
    } else if (warp >= 4) {
        const int q = warp - 4;
        const int TW = p.n_chunk + DW_BIAS_COLS;
        const bool have = s1 > s0;
        if (have) {
            tc::mbar_wait(acc_full, 0);
            tc::tc_fence_after();
        }
        for (int ja = 0; ja < p.mpc; ++ja) {
            const int slab = (mb + ja) * p.n_chunks + nc;
            float* out = p.ws + ((size_t)(split * p.slabs + slab) * 128 + q * 32 + lane) * TW;
            // mpc == 2: accumulator ja sits at column ja * n_chunk and there are no bias columns (written as zeros)
            const uint32_t taddr = tmem_base + (uint32_t)(ja * p.n_chunk) + ((uint32_t)(q * 32) << 16);
            for (int c = 0; c < TW; c += 16) {
                float v[16];
                if (have && (p.mpc == 1 || c < p.n_chunk)) {
                    tc::tmem_ld_32x16(taddr + c, v);
                    tc::tmem_ld_wait();
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = 0.f;
                }
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                    *reinterpret_cast<float4*>(out + c + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 2) tc::tmem_dealloc(tmem_base, 512);
}

// out[m, n] = sum over splits (in order) of the slab partials; db[m] from the bias column of the nc = 0 slabs.
__global__ void k_dw_reduce(const float* __restrict__ ws, int splits, int slabs, int n_chunk, int n_chunks, int M, int N,
                            float* __restrict__ out, int ldo, float* __restrict__ db, int accumulate) {
    const int TW = n_chunk + DW_BIAS_COLS;
    const int64_t total = (int64_t)M * (N + 1);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int m = (int)(i / (N + 1)), n = (int)(i % (N + 1));
        const bool bias = n == N;
        if (bias && !db) continue;
        const int mb = m >> 7, nc = bias ? 0 : n / n_chunk, col = bias ? n_chunk : n % n_chunk;
        const int slab = mb * n_chunks + nc;
        float acc = 0.f;
        for (int s = 0; s < splits; ++s) acc += ws[((size_t)(s * slabs + slab) * 128 + (m & 127)) * TW + col];
        if (bias) db[m] = accumulate ? db[m] + acc : acc;
        else out[(int64_t)m * ldo + n] = accumulate ? out[(int64_t)m * ldo + n] + acc : acc;
    }
}

struct DwPlan {
    int n_chunk, n_chunks, m_blocks, slabs, splits, stages, mpc;
    size_t ws_bytes, smem;
};

inline DwPlan dw_plan(int64_t rows, int M, int N, int b_cm, bool want_bias = true) {
    DwPlan d;
    const int unit = b_cm ? 128 : 64;          // B channels per TMA box
    const int npad = (N + unit - 1) / unit * unit;
    d.n_chunks = (npad + 255) / 256;
    d.n_chunk = ((npad / unit + d.n_chunks - 1) / d.n_chunks) * unit;
    d.m_blocks = (M + 127) / 128;
    d.slabs = d.m_blocks * d.n_chunks;
    // pairs of A blocks per CTA when nothing stands in the way: no bias columns wanted, one B chunk, two accumulators in TMEM
    d.mpc = (!want_bias && d.n_chunks == 1 && d.m_blocks % 2 == 0 && 2 * d.n_chunk <= 512) ? 2 : 1;
    const int64_t steps = (rows + DW_KSTEP - 1) / DW_KSTEP;
    int splits = GSATB_NUM_SMS / (d.slabs / d.mpc);
    if (splits < 1) splits = 1;
    if ((int64_t)splits > steps) splits = steps > 0 ? (int)steps : 1;
    d.splits = splits;
    const int stage_bytes = d.mpc * DW_A_BYTES + d.n_chunk * 128;
    int stages = (227 * 1024 - 1024 - DW_ONES_BYTES - 256) / stage_bytes;
    d.stages = stages > 8 ? 8 : stages;
    d.smem = (size_t)d.stages * stage_bytes + DW_ONES_BYTES + 256 + 1024;
    d.ws_bytes = (size_t)d.splits * d.slabs * 128 * (d.n_chunk + DW_BIAS_COLS) * sizeof(float);
    return d;
}

// operand tensor map: channel-major [C, rows] -> boxes {64 rows, 128 ch}; row-major [rows, C] -> boxes {64 ch, 64 rows}
inline int make_dw_tmap(CUtensorMap* tm, const void* x, int64_t rows, int C, int64_t ld, int cm) {
    PFN_tmapEncodeTiled fn = get_encode_fn();
    if (!fn) return GSATB_ELAUNCH;
    if ((reinterpret_cast<uintptr_t>(x) & 15u) != 0 || (ld % 8) != 0) return GSATB_EALIGN;
    cuuint64_t gdim[2], gstride[1];
    cuuint32_t box[2], estr[2] = {1, 1};
    if (cm == 2) {      // [rows / 128][ld][128]: rows of the map = block * ld + channel, 128 columns
        if (rows % 128 != 0 || ld % 128 != 0 || ld < C) return GSATB_ESHAPE;
        gdim[0] = 128, gdim[1] = (cuuint64_t)(rows / 128) * (cuuint64_t)ld;
        box[0] = 64, box[1] = 128;
        ld = 128;
    } else if (cm) {
        gdim[0] = (cuuint64_t)rows, gdim[1] = (cuuint64_t)C;
        box[0] = 64, box[1] = 128;
    } else {
        gdim[0] = (cuuint64_t)C, gdim[1] = (cuuint64_t)rows;
        box[0] = 64, box[1] = 64;
    }
    gstride[0] = (cuuint64_t)ld * 2;
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(x), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? GSATB_OK : GSATB_EINVAL;
}

}  // namespace

extern "C" size_t gsatb_tc_dw_workspace(int64_t rows, int M, int N) {
    if (rows < 0 || M <= 0 || N <= 0) return 0;
    size_t w = 0;
    for (int b_cm = 0; b_cm < 2; ++b_cm)
        for (int bias = 0; bias < 2; ++bias) {
            const size_t x = dw_plan(rows, M, N, b_cm, bias != 0).ws_bytes;
            w = x > w ? x : w;
        }
    return w;
}

extern "C" int gsatb_tc_dw(const void* a_bf16, int a_channel_major, int64_t lda, const void* b_bf16, int b_channel_major,
                           int64_t ldb, int64_t rows, int M, int N, float* dW, int ldo, float* db, int accumulate,
                           void* workspace, size_t ws_bytes, gsatb_stream_t stream) {
    if (rows < 0 || M <= 0 || N <= 0 || ldo < N) return GSATB_EINVAL;
    if (!dW) return GSATB_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    if (rows == 0) {
        if (!accumulate) {
            for (int m = 0; m < M; ++m)
                if (cudaMemsetAsync(dW + (size_t)m * ldo, 0, (size_t)N * 4, st) != cudaSuccess) return GSATB_ELAUNCH;
            if (db && cudaMemsetAsync(db, 0, (size_t)M * 4, st) != cudaSuccess) return GSATB_ELAUNCH;
        }
        return GSATB_OK;
    }
    if (!a_bf16 || !b_bf16 || !workspace) return GSATB_EINVAL;
    if (a_channel_major < 0 || a_channel_major > 2 || b_channel_major < 0 || b_channel_major > 2) return GSATB_EINVAL;
    const DwPlan d = dw_plan(rows, M, N, b_channel_major ? 1 : 0, db != nullptr);
    if (ws_bytes < d.ws_bytes) return GSATB_EWS_TOO_SMALL;
    if (d.stages < 2) return GSATB_ESHAPE;
    CUtensorMap ta, tb;
    int rc = make_dw_tmap(&ta, a_bf16, rows, M, lda, a_channel_major);
    if (rc != GSATB_OK) return rc;
    rc = make_dw_tmap(&tb, b_bf16, rows, N, ldb, b_channel_major);
    if (rc != GSATB_OK) return rc;
    DwParams p{rows, M, N, a_channel_major, b_channel_major, (int)lda, (int)ldb, d.n_chunk, d.n_chunks, d.slabs, d.mpc, d.splits,
               d.stages, (float*)workspace};
    if (cudaFuncSetAttribute(k_tc_dw, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
        return GSATB_ELAUNCH;
    k_tc_dw<<<dim3(d.splits, d.slabs / d.mpc), DW_THREADS, d.smem, st>>>(ta, tb, p);
    GSATB_CHECK_LAUNCH();
    const int64_t total = (int64_t)M * (N + 1);
    int grid = (int)((total + 255) / 256);
    if (grid > 4 * GSATB_NUM_SMS) grid = 4 * GSATB_NUM_SMS;
    k_dw_reduce<<<grid, 256, 0, st>>>((const float*)workspace, d.splits, d.slabs, d.n_chunk, d.n_chunks, M, N, dW, ldo, db,
                                      accumulate);
    GSATB_CHECK_LAUNCH();
    return GSATB_OK;
}
