// Host-side SIMT emulator: TEST INFRASTRUCTURE ONLY (tests/simt/, `-m "not gpu"` suite).
//
// Lets the kernel sources of dp_gsat_b200/csrc/ be compiled by plain g++ (-DGSATB_HOST_SIM; build_sim.py rewrites the
// <<<...>>> launches in a generated copy) and run on the CPU with CUDA's execution semantics:
// every thread of a block is a cooperative fiber (ucontext); blocks run one after the other; __syncthreads() and the
// *_sync warp primitives are real rendezvous points (a lane named in a mask that never arrives is reported as a
// dead-lock instead of hanging, a lane that calls with a mask it is not part of aborts).  Global / shared memory are
// plain host memory, so out-of-bounds indexing shows up under the usual host tools.  This checks the INDEXING,
// MASK and REDUCTION LOGIC of a kernel without a GPU; it says nothing about races between warps (the schedule is
// deterministic) or about performance.  The product library never includes this file.
#pragma once
#include <setjmp.h>
#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <unordered_map>
#include <vector>

// ---------------------------------------------------------------------------------------------------------------
// vector types / qualifiers
// ---------------------------------------------------------------------------------------------------------------
struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
struct uint3 { unsigned x, y, z; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }

#define __global__ static
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __grid_constant__
#define __align__(n) __attribute__((aligned(n)))

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
enum cudaDeviceAttr { cudaDevAttrComputeCapabilityMajor = 75 };
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = 10; return cudaSuccess; }   // "sm_100"
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) {      // launches run to completion,
    std::memset(p, v, n);                                                                // so stream order is call order
    return cudaSuccess;
}
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) {
    std::memmove(d, s, n);
    return cudaSuccess;
}

namespace simt {

// One fiber per thread SLOT of a block, created once and reused by every block and launch (its body loops: run the
// kernel for the (block, thread) the scheduler has set, report done, wait to be resumed for the next one).
struct Fiber {
    ucontext_t ctx;      // first entry only (makecontext); afterwards fibers switch with _setjmp / _longjmp, which --
    jmp_buf jb;          // unlike swapcontext -- do not make a sigprocmask system call per switch
    char* stack = nullptr;
    bool started = false;
    bool done = false;
};
struct WarpBarrier {
    uint32_t arrived = 0;
    uint64_t gen = 0;
};
struct Warp {
    uint64_t slot[32];
    uint32_t exists = 0;                                      // lanes that exist in this (possibly partial) warp
    std::unordered_map<uint32_t, WarpBarrier> bars;           // one rendezvous per distinct mask
};

struct NamedBarrier {
    int arrived = 0;
    uint64_t gen = 0;
};
constexpr size_t kDynSmemBytes = 256 * 1024;        // >= the 227 KiB a CTA can opt into
constexpr uint32_t kTmemCols = 512;

struct State {
    ucontext_t sched;
    jmp_buf sched_jb;
    uint8_t* dyn_smem = nullptr;  // 1024-aligned dynamic shared memory of the running block
    float* tmem = nullptr;        // [128 lanes][kTmemCols] tensor memory of the running block
    uint32_t tmem_next = 0;       // bump allocator of TMEM columns
    NamedBarrier named[16];
    std::vector<Fiber*> fibers;   // pool, grows to the largest block seen; addresses and stacks never move
    std::vector<Warp> warps;
    const std::function<void()>* body = nullptr;
    int cur = -1;
    int live = 0;                 // fibers not finished
    int block_arrived = 0;        // fibers waiting in __syncthreads
    uint64_t block_gen = 0;
    uint64_t progress = 0;        // bumped on every arrival / release / exit: no change over a full pass = dead-lock
    dim3 block_dim, grid_dim;
};
inline State& S() {
    static State s;
    return s;
}
constexpr size_t kStackBytes = 256 * 1024;

[[noreturn]] inline void fail(const char* msg) {
    std::fprintf(stderr, "[simt] %s (block thread %d)\n", msg, S().cur);
    std::abort();
}
inline void yield() {
    State& s = S();
    if (_setjmp(s.fibers[s.cur]->jb) == 0) _longjmp(s.sched_jb, 1);
}
inline void trampoline() {
    State& s = S();
    for (;;) {                               // s.cur / threadIdx / blockIdx / s.body were set by the scheduler
        (*s.body)();
        s.fibers[s.cur]->done = true;
        --s.live;
        ++s.progress;
        yield();                             // resumed when this slot gets its next (block, thread)
    }
}
// scheduler side: run fiber t until it yields or finishes
inline void resume(int t) {
    State& s = S();
    Fiber& f = *s.fibers[t];
    if (_setjmp(s.sched_jb) == 0) {
        if (!f.started) {
            f.started = true;
            setcontext(&f.ctx);
        } else {
            _longjmp(f.jb, 1);
        }
    }
}
inline int lane_id() { return S().cur & 31; }
inline Warp& warp() { return S().warps[S().cur >> 5]; }

inline void warp_barrier(unsigned mask) {
    State& s = S();
    const int lane = lane_id();
    Warp& w = warp();
    if (!((mask >> lane) & 1u)) fail("a lane called a *_sync primitive with a mask that does not name it");
    const uint32_t eff = mask & w.exists;
    WarpBarrier& b = w.bars[mask];
    ++s.progress;
    b.arrived |= 1u << lane;
    if ((b.arrived & eff) == eff) {
        b.arrived = 0;
        ++b.gen;
        return;
    }
    const uint64_t gen = b.gen;
    while (b.gen == gen) yield();
}
template <class T>
inline uint64_t to_bits(T v) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    uint64_t u = 0;
    std::memcpy(&u, &v, sizeof(T));
    return u;
}
template <class T>
inline T from_bits(uint64_t u) {
    T v;
    std::memcpy(&v, &u, sizeof(T));
    return v;
}
template <class T>
inline T shfl_from(unsigned mask, T v, int src_lane) {
    Warp& w = warp();
    const int lane = lane_id();
    w.slot[lane] = to_bits(v);
    warp_barrier(mask);
    T r = v;
    if (src_lane >= 0 && src_lane < 32) {
        if (!((mask >> src_lane) & 1u)) fail("shuffle reads a lane outside the mask (undefined in CUDA)");
        r = from_bits<T>(w.slot[src_lane]);
    }
    warp_barrier(mask);
    return r;
}

inline void launch(dim3 grid, dim3 block, const std::function<void()>& body, size_t dyn_smem_bytes = 0);

inline uint8_t* dyn_smem() { return S().dyn_smem; }
inline float* tmem() { return S().tmem; }
inline uint32_t tmem_alloc_cols(uint32_t ncols) {
    State& s = S();
    if (ncols < 32 || (ncols & (ncols - 1)) || s.tmem_next + ncols > kTmemCols) fail("tcgen05.alloc: bad column count");
    const uint32_t base = s.tmem_next;
    s.tmem_next += ncols;
    return base;                                      // lane 0, column `base`
}
// bar.sync id, nthreads
inline void named_barrier(int id, int nthreads) {
    State& s = S();
    if (id < 0 || id >= 16) fail("named barrier id out of range");
    NamedBarrier& b = s.named[id];
    ++s.progress;
    if (++b.arrived == nthreads) {
        b.arrived = 0;
        ++b.gen;
        return;
    }
    const uint64_t gen = b.gen;
    while (b.gen == gen) yield();
}

}  // namespace simt

// built-in variables (set by the scheduler before a fiber is resumed)
inline uint3 threadIdx, blockIdx;
inline dim3 blockDim, gridDim;

inline void simt::launch(dim3 grid, dim3 block, const std::function<void()>& body, size_t dyn_smem_bytes) {
    State& s = S();
    if (dyn_smem_bytes > kDynSmemBytes) fail("dynamic shared memory request too large");
    if (!s.dyn_smem) {
        s.dyn_smem = static_cast<uint8_t*>(std::aligned_alloc(1024, kDynSmemBytes));
        s.tmem = static_cast<float*>(std::malloc(sizeof(float) * 128 * kTmemCols));
    }
    const int nthreads = (int)(block.x * block.y * block.z);
    if (nthreads <= 0 || nthreads > 1024) fail("block size out of range");
    if (grid.x == 0 || grid.y == 0 || grid.z == 0 || grid.y > 65535 || grid.z > 65535) fail("grid size out of range");
    s.body = &body;
    s.block_dim = block;
    s.grid_dim = grid;
    while ((int)s.fibers.size() < nthreads) {                 // new thread slots: one context + stack each, made once
        Fiber* f = new Fiber();
        f->stack = static_cast<char*>(std::malloc(kStackBytes));
        getcontext(&f->ctx);
        f->ctx.uc_stack.ss_sp = f->stack;
        f->ctx.uc_stack.ss_size = kStackBytes;
        f->ctx.uc_link = nullptr;                             // the trampoline never returns
        makecontext(&f->ctx, (void (*)())simt::trampoline, 0);
        s.fibers.push_back(f);
    }
    blockDim = block;
    gridDim = grid;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                blockIdx = uint3{bx, by, bz};
                s.tmem_next = 0;
                for (auto& nb : s.named) nb = NamedBarrier();
                s.warps.assign((nthreads + 31) / 32, Warp());
                for (int t = 0; t < nthreads; ++t) {
                    s.warps[t >> 5].exists |= 1u << (t & 31);
                    s.fibers[t]->done = false;
                }
                s.live = nthreads;
                s.block_arrived = 0;
                while (s.live > 0) {
                    const uint64_t before = s.progress;
                    for (int t = 0; t < nthreads; ++t) {
                        if (s.fibers[t]->done) continue;
                        s.cur = t;
                        threadIdx = uint3{(unsigned)t % block.x, ((unsigned)t / block.x) % block.y,
                                          (unsigned)t / (block.x * block.y)};
                        resume(t);
                        if (s.block_arrived > 0 && s.block_arrived == s.live) {       // __syncthreads released
                            s.block_arrived = 0;
                            ++s.block_gen;
                            ++s.progress;
                        }
                    }
                    if (s.live > 0 && s.progress == before)
                        fail("dead-lock: threads wait at a barrier / *_sync primitive that the others never reach");
                }
            }
    s.cur = -1;
}

// ---------------------------------------------------------------------------------------------------------------
// intrinsics
// ---------------------------------------------------------------------------------------------------------------
inline void __syncthreads() {
    simt::State& s = simt::S();
    const uint64_t gen = s.block_gen;
    ++s.block_arrived;
    ++s.progress;
    while (s.block_gen == gen) simt::yield();
}
inline void __syncwarp(unsigned mask = 0xffffffffu) { simt::warp_barrier(mask); }
template <class T>
inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    const int lane = simt::lane_id();
    return simt::shfl_from(mask, v, (lane & ~(width - 1)) | (src & (width - 1)));
}
template <class T>
inline T __shfl_xor_sync(unsigned mask, T v, int lane_mask, int width = 32) {
    const int lane = simt::lane_id();
    const int src = lane ^ lane_mask;
    return simt::shfl_from(mask, v, (src & ~(width - 1)) == (lane & ~(width - 1)) ? src : lane);
}
template <class T>
inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    const int lane = simt::lane_id();
    const int src = lane + (int)delta;
    return simt::shfl_from(mask, v, (src & ~(width - 1)) == (lane & ~(width - 1)) ? src : lane);
}
template <class T>
inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    const int lane = simt::lane_id();
    const int src = lane - (int)delta;
    return simt::shfl_from(mask, v, (src >= 0 && (src & ~(width - 1)) == (lane & ~(width - 1))) ? src : lane);
}
inline unsigned __ballot_sync(unsigned mask, int pred) {
    simt::Warp& w = simt::warp();
    w.slot[simt::lane_id()] = pred ? 1u : 0u;
    simt::warp_barrier(mask);
    unsigned r = 0;
    for (int l = 0; l < 32; ++l)
        if (((mask & w.exists) >> l) & 1u) r |= (unsigned)(w.slot[l] & 1u) << l;
    simt::warp_barrier(mask);
    return r;
}
template <class T>
inline unsigned __match_any_sync(unsigned mask, T v) {
    simt::Warp& w = simt::warp();
    const int lane = simt::lane_id();
    w.slot[lane] = simt::to_bits(v);
    simt::warp_barrier(mask);
    unsigned r = 0;
    for (int l = 0; l < 32; ++l)
        if ((((mask & w.exists) >> l) & 1u) && w.slot[l] == w.slot[lane]) r |= 1u << l;
    simt::warp_barrier(mask);
    return r;
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, !pred) == 0; }

template <class T>
inline T __ldg(const T* p) { return *p; }
template <class T>
inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <class T>
inline T atomicMax(T* p, T v) { T o = *p; *p = std::max(o, v); return o; }
template <class T>
inline T atomicMin(T* p, T v) { T o = *p; *p = std::min(o, v); return o; }
template <class T>
inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
inline long long clock64() { return 0; }
inline float __uint_as_float(unsigned u) { return simt::from_bits<float>(simt::to_bits(u)); }
inline unsigned __float_as_uint(float f) { return simt::from_bits<unsigned>(simt::to_bits(f)); }
inline unsigned __byte_perm(unsigned x, unsigned y, unsigned sel) {       // PTX prmt, default mode
    const uint64_t both = ((uint64_t)y << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; ++i) {
        const unsigned c = (sel >> (4 * i)) & 0xf;
        unsigned byte = (unsigned)((both >> (8 * (c & 7))) & 0xff);
        if (c & 8) byte = (byte & 0x80) ? 0xff : 0x00;                     // sign replication
        r |= byte << (8 * i);
    }
    return r;
}
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
inline int __float_as_int(float f) { return simt::from_bits<int>(simt::to_bits(f)); }
inline float __int_as_float(int i) { return simt::from_bits<float>(simt::to_bits(i)); }
using std::max;
using std::min;
inline int64_t min(int64_t a, int b) { return a < b ? a : (int64_t)b; }
inline int64_t min(int a, int64_t b) { return a < b ? (int64_t)a : b; }
inline int64_t max(int64_t a, int b) { return a > b ? a : (int64_t)b; }
inline int64_t max(int a, int64_t b) { return a > b ? (int64_t)a : b; }
