// Host stand-in for the driver-API pieces csrc/tc_gemm.cuh uses (<cuda.h>) in the SIMT-emulator build (tests/simt):
// a CUtensorMap that simply records the tiled-encode arguments, and the enums of cuTensorMapEncodeTiled.
#pragma once
#include <cstdint>
#include "simt.h"

typedef uint32_t cuuint32_t;
typedef uint64_t cuuint64_t;
typedef int CUresult;
enum { CUDA_SUCCESS = 0 };
enum CUtensorMapDataType { CU_TENSOR_MAP_DATA_TYPE_FLOAT32 = 7, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 = 9 };
enum CUtensorMapInterleave { CU_TENSOR_MAP_INTERLEAVE_NONE = 0 };
enum CUtensorMapSwizzle { CU_TENSOR_MAP_SWIZZLE_NONE = 0, CU_TENSOR_MAP_SWIZZLE_128B = 3 };
enum CUtensorMapL2promotion { CU_TENSOR_MAP_L2_PROMOTION_L2_256B = 3 };
enum CUtensorMapFloatOOBfill { CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE = 0 };

struct alignas(64) CUtensorMap {
    const uint8_t* base;
    uint64_t dim[2];        // elements: [0] innermost (K), [1] rows
    uint64_t row_stride;    // bytes between rows
    uint32_t box[2];        // elements: [0] innermost, [1] rows
    uint32_t elem_bytes;
    uint32_t swizzle;
    uint8_t pad[128 - 56];
};
static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap is a 128-byte opaque object");

static inline CUresult simt_tensor_map_encode_tiled(CUtensorMap* tm, CUtensorMapDataType dt, cuuint32_t rank, void* base,
                                                    const cuuint64_t* gdim, const cuuint64_t* gstride,
                                                    const cuuint32_t* box, const cuuint32_t* estr, CUtensorMapInterleave,
                                                    CUtensorMapSwizzle sw, CUtensorMapL2promotion,
                                                    CUtensorMapFloatOOBfill) {
    if (rank != 2 || (dt != CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 && dt != CU_TENSOR_MAP_DATA_TYPE_FLOAT32) || estr[0] != 1 || estr[1] != 1) return 1;
    const uint32_t es = dt == CU_TENSOR_MAP_DATA_TYPE_FLOAT32 ? 4u : 2u;
    if ((reinterpret_cast<uintptr_t>(base) & 15u) || (gstride[0] & 15u)) return 1;      // the driver's alignment rules
    if (box[0] * es > 128 && sw == CU_TENSOR_MAP_SWIZZLE_128B) return 1;                   // inner box <= swizzle span
    tm->base = static_cast<const uint8_t*>(base);
    tm->dim[0] = gdim[0];
    tm->dim[1] = gdim[1];
    tm->row_stride = gstride[0];
    tm->box[0] = box[0];
    tm->box[1] = box[1];
    tm->elem_bytes = es;
    tm->swizzle = (uint32_t)sw;
    return CUDA_SUCCESS;
}

// cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", ...)
enum cudaDriverEntryPointQueryResult { cudaDriverEntryPointSuccess = 0 };
enum { cudaEnableDefault = 0 };
static inline cudaError_t cudaGetDriverEntryPoint(const char*, void** fn, int, cudaDriverEntryPointQueryResult* q) {
    *fn = reinterpret_cast<void*>(&simt_tensor_map_encode_tiled);
    *q = cudaDriverEntryPointSuccess;
    return cudaSuccess;
}
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <class F>
static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
