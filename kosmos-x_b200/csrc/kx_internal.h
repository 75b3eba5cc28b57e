// Internal helpers shared by the translation units of libkosmosx_sm100.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/kosmosx_b200.h"

namespace kx {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int device_sm_count();            // <= 0 when there is no usable device (error text set)

typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
struct DriverApi {
    encode_tiled_fn encode_tiled = nullptr;
};
const DriverApi& driver_api();    // resolved through cudaGetDriverEntryPoint (no link-time libcuda dependency)

bool make_tmap_bf16_2d(CUtensorMap* tm, const void* ptr, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                       uint32_t box_inner, uint32_t box_outer);

bool make_tmap_f32_2d(CUtensorMap* tm, const void* ptr, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                      uint32_t box_inner, uint32_t box_outer);     // fp32, SWIZZLE_128B (box_inner <= 32)

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s launch failed: %s", what, cudaGetErrorString(e));
        return KX_ERR_LAUNCH;
    }
    count_launch();
    return KX_OK;
}

}  // namespace kx
