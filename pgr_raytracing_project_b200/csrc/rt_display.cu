// rt_display.cu -- the reference's display chain on the device:
//   interaction.py:1435-1439  _tone_map        x*e / (1 + x*e), clip [0,1]
//   interaction.py:1441-1449  _enhance_display  percentile(2) / percentile(98) contrast stretch over ALL values
//   gui.py:73                 ImageDisplay      clip * 255 -> uint8 (truncation)
// so that W*H*3 BYTES cross PCIe per displayed frame instead of the float frame plus ~6 full-frame numpy passes
// and two np.percentile calls on the host.  Bit-exact against numpy (>= 2.0 float32 semantics): the percentiles
// are exact order statistics (radix sort of the tone-mapped values) combined with numpy's own _lerp formula in
// float32; every elementwise step is the float32 operation numpy performs.
#include <cub/cub.cuh>

#include "rt_display.h"

namespace b200rt {

namespace {

__global__ void k_tonemap_f32(const float* __restrict__ accum, float* __restrict__ out, int64_t n, float exposure) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        float x = __fmul_rn(accum[k], exposure);
        x = __fdiv_rn(x, __fadd_rn(1.0f, x));
        out[k] = fminf(fmaxf(x, 0.0f), 1.0f);
    }
}

// numpy percentile, method "linear", float32 input: virtual index (n-1)*q in float64, gamma cast to float32,
// _lerp: a + (b-a)*t, and b - (b-a)*(1-t) where t >= 0.5
__device__ __forceinline__ float np_percentile(const float* __restrict__ sorted, int64_t n, double q) {
    const double vi = (double)(n - 1) * q;
    int64_t lo = (int64_t)floor(vi);
    if (lo < 0) lo = 0;
    if (lo > n - 1) lo = n - 1;
    const int64_t hi = lo + 1 < n ? lo + 1 : n - 1;
    const float t = (float)(vi - (double)lo);
    const float a = sorted[lo], b = sorted[hi];
    const float d = __fsub_rn(b, a);
    if (t >= 0.5f) return __fsub_rn(b, __fmul_rn(d, __fsub_rn(1.0f, t)));
    return __fadd_rn(a, __fmul_rn(d, t));
}

__global__ void k_stretch_u8(const float* __restrict__ tone, const float* __restrict__ sorted, uint8_t* __restrict__ out, int64_t n) {
    const float mn = np_percentile(sorted, n, 2.0 / 100.0), mx = np_percentile(sorted, n, 98.0 / 100.0);
    const bool stretch = mx > mn;
    const float range = __fsub_rn(mx, mn);
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        float v = tone[k];
        if (stretch) v = fminf(fmaxf(__fdiv_rn(__fsub_rn(v, mn), range), 0.0f), 1.0f);
        out[k] = (uint8_t)__fmul_rn(fminf(fmaxf(v, 0.0f), 1.0f), 255.0f);
    }
}

inline int grid_of(int64_t n) { int64_t g = (n + 255) / 256; return (int)(g > 148 * 16 ? 148 * 16 : (g < 1 ? 1 : g)); }

}  // namespace

size_t display_scratch_bytes(int64_t n) {
    size_t sort_bytes = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, sort_bytes, (const float*)nullptr, (float*)nullptr, (int)n);
    return 2 * (size_t)n * sizeof(float) + ((sort_bytes + 255) & ~(size_t)255) + 256;
}

cudaError_t launch_display_u8(const float* d_accum, uint8_t* d_rgb8, int64_t n, float exposure, void* d_scratch, size_t scratch_bytes,
                              cudaStream_t stream, int* n_launches) {
    if (n == 0) return cudaSuccess;
    float* tone = static_cast<float*>(d_scratch);
    float* sorted = tone + n;
    char* cub_tmp = reinterpret_cast<char*>(sorted + n);
    cub_tmp = reinterpret_cast<char*>(((uintptr_t)cub_tmp + 255) & ~(uintptr_t)255);
    size_t cub_bytes = scratch_bytes - (size_t)(cub_tmp - static_cast<char*>(d_scratch));
    k_tonemap_f32<<<grid_of(n), 256, 0, stream>>>(d_accum, tone, n, exposure);
    cudaError_t e = cub::DeviceRadixSort::SortKeys(cub_tmp, cub_bytes, tone, sorted, (int)n, 0, 32, stream);
    if (e != cudaSuccess) return e;
    k_stretch_u8<<<grid_of(n), 256, 0, stream>>>(tone, sorted, d_rgb8, n);
    *n_launches += 3;
    return cudaGetLastError();
}

}  // namespace b200rt
