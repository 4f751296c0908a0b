// Host implementations of the PTX wrappers of dp_gsat_b200/csrc/tc.cuh for the SIMT-emulator build (tests/simt).
// TEST INFRASTRUCTURE ONLY.  Functional models, not timing models:
//   * an mbarrier is its 64-bit shared-memory word: phase bit, pending arrivals, expected arrivals, pending tx bytes;
//   * TMA box loads, tcgen05.mma and cp.async complete at once (then signal their mbarrier), so every wait in the
//     kernels is eventually satisfied or reported as a dead-lock by the scheduler;
//   * TMEM is a per-block float[128 lanes][512 columns]; tcgen05.mma reads its operands from shared memory through the
//     same descriptors the hardware would decode (start address, stride byte offset, SWIZZLE_128B address XOR) and
//     accumulates in fp32 in k order (the hardware's order is unspecified: results agree to rounding, not bitwise);
//   * named barriers (bar.sync id, n) are real rendezvous points of n threads.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <cuda.h>

namespace tc {

inline uint32_t smem_u32(const void* p) {
    const uint8_t* base = simt::dyn_smem();
    const uint8_t* q = static_cast<const uint8_t*>(p);
    if (q < base || q >= base + simt::kDynSmemBytes) simt::fail("smem_u32 of a pointer outside dynamic shared memory");
    return (uint32_t)(q - base);
}

// ---- mbarrier: [63] phase | [62:48] expected | [47:32] pending | [31:0] tx bytes (signed) ------------------------
struct MbarView {
    uint64_t* w;
    uint32_t phase() const { return (uint32_t)(*w >> 63); }
    uint32_t expected() const { return (uint32_t)((*w >> 48) & 0x7fff); }
    int32_t pending() const { return (int32_t)((*w >> 32) & 0xffff); }
    int32_t tx() const { return (int32_t)(*w & 0xffffffffu); }
    void set(uint32_t ph, uint32_t ex, int32_t pe, int32_t t) {
        *w = ((uint64_t)ph << 63) | ((uint64_t)ex << 48) | ((uint64_t)(uint32_t)(pe & 0xffff) << 32) | (uint32_t)t;
    }
    void settle() {                       // phase completes when every arrival is in and every expected byte landed
        if (pending() == 0 && tx() == 0) set(phase() ^ 1u, expected(), (int32_t)expected(), 0);
        ++simt::S().progress;
    }
};
inline void mbar_init(uint64_t* bar, uint32_t count) { MbarView{bar}.set(0, count, (int32_t)count, 0); }
inline void fence_barrier_init() {}
inline void mbar_arrive(uint64_t* bar) {
    MbarView m{bar};
    if (m.pending() <= 0) simt::fail("mbarrier.arrive beyond the expected arrival count");
    m.set(m.phase(), m.expected(), m.pending() - 1, m.tx());
    m.settle();
}
inline void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    MbarView m{bar};
    if (m.pending() <= 0) simt::fail("mbarrier.arrive.expect_tx beyond the expected arrival count");
    m.set(m.phase(), m.expected(), m.pending() - 1, m.tx() + (int32_t)bytes);
    m.settle();
}
inline void mbar_complete_tx(uint64_t* bar, uint32_t bytes) {
    MbarView m{bar};
    m.set(m.phase(), m.expected(), m.pending(), m.tx() - (int32_t)bytes);
    m.settle();
}
// the phase with parity `parity` has completed  <=>  the barrier is now in the other phase
inline bool mbar_try_wait(uint64_t* bar, uint32_t parity) { return MbarView{bar}.phase() != (parity & 1u); }
inline bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t) {
    if (mbar_try_wait(bar, parity)) return true;
    simt::yield();
    return mbar_try_wait(bar, parity);
}
inline void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) simt::yield();
}
inline void named_bar_sync(int id, int nthreads) { simt::named_barrier(id, nthreads); }
inline void group_mbar_wait(bool leader, uint64_t* bar, uint32_t parity, int bar_id, int nthreads) {
    if (leader) mbar_wait(bar, parity);
    named_bar_sync(bar_id, nthreads);
}
inline void fence_proxy_async_smem() {}
template <int N>
inline void reg_dec() {}
template <int N>
inline void reg_inc() {}

// ---- explicit shared-space accesses on 32-bit shared addresses (= offsets into the block's dynamic shared memory) ----
inline void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    if (saddr & 15u) simt::fail("st.shared.v4 not 16-byte aligned");
    uint32_t* q = reinterpret_cast<uint32_t*>(simt::dyn_smem() + saddr);
    q[0] = a, q[1] = b, q[2] = c, q[3] = d;
}
inline void sts32(uint32_t saddr, uint32_t a) { *reinterpret_cast<uint32_t*>(simt::dyn_smem() + saddr) = a; }
inline void sts_f32(uint32_t saddr, float a) { *reinterpret_cast<float*>(simt::dyn_smem() + saddr) = a; }
inline uint32_t lds32(uint32_t saddr) { return *reinterpret_cast<const uint32_t*>(simt::dyn_smem() + saddr); }
inline float lds_f32(uint32_t saddr) { return *reinterpret_cast<const float*>(simt::dyn_smem() + saddr); }
inline uint4 lds128(uint32_t saddr) {
    if (saddr & 15u) simt::fail("ld.shared.v4 not 16-byte aligned");
    return *reinterpret_cast<const uint4*>(simt::dyn_smem() + saddr);
}
inline uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) {
    return (uint32_t)(((((uint64_t)hi) << 32) | lo) >> (sh & 31u));
}

// ---- TMA -------------------------------------------------------------------------------------------------------
inline void tma_prefetch_desc(const void*) {}
inline void prefetch_l2(const void*) {}
inline uint32_t swizzle128(uint32_t byte_addr) { return byte_addr ^ (((byte_addr >> 7) & 7u) << 4); }
inline void tma_load_2d(void* smem_dst, const void* desc, uint64_t* bar, int crd0, int crd1) {
    const CUtensorMap& tm = *static_cast<const CUtensorMap*>(desc);
    uint8_t* base = simt::dyn_smem();
    const uint32_t dst = smem_u32(smem_dst);
    if (tm.swizzle == CU_TENSOR_MAP_SWIZZLE_128B && (dst & 1023u)) simt::fail("SWIZZLE_128B box not 1024-byte aligned");
    const uint32_t row_bytes = tm.box[0] * tm.elem_bytes;
    for (uint32_t r = 0; r < tm.box[1]; ++r)
        for (uint32_t c = 0; c < tm.box[0]; ++c) {
            const int64_t gr = (int64_t)crd1 + r, gc = (int64_t)crd0 + c;
            const uint32_t es = tm.elem_bytes;
            uint8_t v[4] = {0, 0, 0, 0};                               // out-of-bounds elements read as zero
            if (gr >= 0 && gc >= 0 && (uint64_t)gr < tm.dim[1] && (uint64_t)gc < tm.dim[0])
                memcpy(v, tm.base + (uint64_t)gr * tm.row_stride + (uint64_t)gc * es, es);
            uint32_t off = dst + r * row_bytes + c * es;
            if (tm.swizzle == CU_TENSOR_MAP_SWIZZLE_128B) off = swizzle128(off);
            memcpy(base + off, v, es);
        }
    mbar_complete_tx(bar, tm.box[1] * row_bytes);
}

// ---- tcgen05 ---------------------------------------------------------------------------------------------------
inline void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {          // called by one full warp
    if (simt::lane_id() == 0) *smem_result = simt::tmem_alloc_cols(ncols);
}
inline void tmem_relinquish() {}
inline void tmem_dealloc(uint32_t, uint32_t) {}
inline void tc_fence_before() {}
inline void tc_fence_after() {}

inline float bf16_at(const uint8_t* base, uint32_t off) {
    const uint32_t u = (uint32_t)(*reinterpret_cast<const uint16_t*>(base + off)) << 16;
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}
// Byte address of element (i = M/N index, k) of an operand tile described by a shared-memory descriptor.
//   K-major  SWIZZLE_128B: 128-byte rows of 64 K elements, row i at (i / 8) * SBO + (i % 8) * 128
//   MN-major SWIZZLE_128B: for every k a 128-byte row of 64 consecutive M/N elements; 8 k-rows per 1024-byte atom, the
//                          next 8 k at + SBO, the next 64 M/N elements at + LBO (canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO))
//                          in 16-byte units)
inline uint32_t operand_addr(uint64_t desc, bool mn_major, int i, int k) {
    const uint32_t a0 = (uint32_t)(desc & 0x3fff) << 4, lbo = (uint32_t)((desc >> 16) & 0x3fff) << 4,
                   sbo = (uint32_t)((desc >> 32) & 0x3fff) << 4;
    if (!mn_major) return swizzle128(a0 + (uint32_t)(i >> 3) * sbo + (uint32_t)(i & 7) * 128 + (uint32_t)k * 2);
    return swizzle128(a0 + (uint32_t)(i >> 6) * lbo + (uint32_t)(k >> 3) * sbo + (uint32_t)(k & 7) * 128 + (uint32_t)(i & 63) * 2);
}
// D[tmem lane m][column n] (+)= sum_k A[m][k] * B[n][k]; M, N, operand majors from the instruction descriptor,
// K = 16 (kind::f16)
inline void mma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    const int M = (int)((idesc >> 24) & 0x1f) << 4, N = (int)((idesc >> 17) & 0x3f) << 3;
    const bool a_mn = (idesc >> 15) & 1u, b_mn = (idesc >> 16) & 1u;
    if (M != 128) simt::fail("only the M = 128 accumulator layout (row m -> TMEM lane m) is modelled");
    if (N < 16 || N > 256 || (N & 15)) simt::fail("tcgen05.mma: N must be a multiple of 16 in [16, 256] for M = 128");
    if (((desc_a >> 61) & 7) != 2 || ((desc_b >> 61) & 7) != 2) simt::fail("only SWIZZLE_128B operands are modelled");
    const uint8_t* base = simt::dyn_smem();
    float* tm = simt::tmem();
    const uint32_t col0 = tmem_d & 0xffff;
    if ((tmem_d >> 16) != 0 || col0 + N > simt::kTmemCols) simt::fail("MMA accumulator outside the allocated TMEM");
    // decode both operand tiles once (the address arithmetic of the descriptors is the expensive part), then accumulate
    static thread_local float a[128][16], b[256][16];
    for (int m = 0; m < M; ++m)
        for (int k = 0; k < 16; ++k) a[m][k] = bf16_at(base, operand_addr(desc_a, a_mn, m, k));
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < 16; ++k) b[n][k] = bf16_at(base, operand_addr(desc_b, b_mn, n, k));
    for (int m = 0; m < M; ++m) {
        float* row = tm + (size_t)m * simt::kTmemCols + col0;
        for (int n = 0; n < N; ++n) {
            float acc = accumulate ? row[n] : 0.f;
            for (int k = 0; k < 16; ++k) acc += a[m][k] * b[n][k];
            row[n] = acc;
        }
    }
}
inline void mma_commit(uint64_t* bar) { mbar_arrive(bar); }         // the emulated MMAs have already completed
inline void tmem_ld_32x32(uint32_t taddr, float* v) {
    const uint32_t lane0 = taddr >> 16, col = taddr & 0xffff;
    if (lane0 != 32u * ((uint32_t)(simt::S().cur >> 5) & 3u)) simt::fail("tcgen05.ld: a warp may only read its own TMEM lane quarter");
    const float* row = simt::tmem() + (size_t)(lane0 + simt::lane_id()) * simt::kTmemCols + col;
    for (int j = 0; j < 32; ++j) v[j] = row[j];
}
inline void tmem_ld_cols(uint32_t taddr, float* v, int ncols) {
    const uint32_t lane0 = taddr >> 16, col = taddr & 0xffff;
    if (lane0 != 32u * ((uint32_t)(simt::S().cur >> 5) & 3u)) simt::fail("tcgen05.ld: a warp may only read its own TMEM lane quarter");
    if (col + ncols > (uint32_t)simt::kTmemCols) simt::fail("tcgen05.ld beyond the TMEM columns");
    const float* row = simt::tmem() + (size_t)(lane0 + simt::lane_id()) * simt::kTmemCols + col;
    for (int j = 0; j < ncols; ++j) v[j] = row[j];
}
inline void tmem_ld_32x16(uint32_t taddr, float* v) { tmem_ld_cols(taddr, v, 16); }
inline void tmem_ld_32x8(uint32_t taddr, float* v) { tmem_ld_cols(taddr, v, 8); }
inline void tmem_ld_wait() {}

// TMA store: shared-memory box (same layout rules as the load) -> global tensor; out-of-bounds elements are dropped
inline void tma_store_2d(const void* desc, const void* smem_src, int crd0, int crd1) {
    const CUtensorMap& tm = *static_cast<const CUtensorMap*>(desc);
    const uint8_t* base = simt::dyn_smem();
    const uint32_t src = smem_u32(smem_src);
    if (tm.swizzle == CU_TENSOR_MAP_SWIZZLE_128B && (src & 1023u)) simt::fail("SWIZZLE_128B box not 1024-byte aligned");
    const uint32_t row_bytes = tm.box[0] * tm.elem_bytes;
    for (uint32_t r = 0; r < tm.box[1]; ++r)
        for (uint32_t c = 0; c < tm.box[0]; ++c) {
            const int64_t gr = (int64_t)crd1 + r, gc = (int64_t)crd0 + c;
            if (gr < 0 || gc < 0 || (uint64_t)gr >= tm.dim[1] || (uint64_t)gc >= tm.dim[0]) continue;
            const uint32_t es = tm.elem_bytes;
            uint32_t off = src + r * row_bytes + c * es;
            if (tm.swizzle == CU_TENSOR_MAP_SWIZZLE_128B) off = swizzle128(off);
            memcpy(const_cast<uint8_t*>(tm.base) + (uint64_t)gr * tm.row_stride + (uint64_t)gc * es, base + off, es);
        }
}
inline void tma_store_commit() {}
template <int N>
inline void tma_store_wait_read() {}
template <int N>
inline void tma_store_wait_all() {}

}  // namespace tc
