// rt_display.h -- device display chain (rt_display.cu): tone map + percentile contrast stretch + uint8 pack.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace b200rt {

size_t display_scratch_bytes(int64_t n_floats);
cudaError_t launch_display_u8(const float* d_accum, uint8_t* d_rgb8, int64_t n_floats, float exposure, void* d_scratch,
                              size_t scratch_bytes, cudaStream_t stream, int* n_launches);

}  // namespace b200rt
