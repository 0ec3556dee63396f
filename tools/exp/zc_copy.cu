// Experiment: can SM-issued stores into mapped pinned host memory carry a finished frame at PCIe rate?
//   a) DMA reference (cudaMemcpyAsync D2H)           b) flat coalesced 16-byte copy, G CTAs x 256 threads
//   c) tile pattern: one warp copies a 32x32-pixel tile (32 rows x 384 B at a 23040-B pitch), W warps in flight
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o zc_copy zc_copy.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void k_flat(const float4* __restrict__ src, float4* __restrict__ dst, size_t n) {
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) dst[k] = src[k];
}
// tiles_x x tiles_y tiles of 32 rows x 24 float4; warp w takes tiles w, w + n_warps, ...
__global__ void k_tiles(const float4* __restrict__ src, float4* __restrict__ dst, int tiles_x, int tiles_y, int pitch4) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    for (int t = warp; t < tiles_x * tiles_y; t += n_warps) {
        const int ty = t / tiles_x, tx = t - ty * tiles_x;
        const size_t base = (size_t)ty * 32 * pitch4 + (size_t)tx * 24;
        if (lane < 24) {
#pragma unroll 8
            for (int r = 0; r < 32; ++r) dst[base + (size_t)r * pitch4 + lane] = src[base + (size_t)r * pitch4 + lane];
        }
    }
}
int main() {
    const int W = 1920, H = 1080;
    const size_t bytes = (size_t)W * H * 12, n4 = bytes / 16;
    float4 *d, *h, *hd;
    CK(cudaMalloc(&d, bytes));
    CK(cudaHostAlloc(&h, bytes, cudaHostAllocMapped));
    CK(cudaHostGetDevicePointer(&hd, h, 0));
    CK(cudaMemset(d, 1, bytes));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    auto report = [&](const char* name, int p, float ms) { printf("%-28s %5d  %.3f ms  %.1f GB/s\n", name, p, ms, bytes / ms / 1e6); };
    float ms;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(a); CK(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost)); cudaEventRecord(b); cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms, a, b); if (rep == 2) report("DMA cudaMemcpyAsync", 0, ms);
    }
    for (int g : {8, 16, 37, 74, 148, 296, 592}) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(a); k_flat<<<g, 256>>>(d, hd, n4); cudaEventRecord(b); CK(cudaEventSynchronize(b));
            cudaEventElapsedTime(&ms, a, b); if (rep == 2) report("flat 16B copy, CTAs", g, ms);
        }
    }
    for (int g : {8, 16, 37, 74, 148, 296}) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(a); k_tiles<<<g, 256>>>(d, hd, W / 32, (H + 31) / 32 - 1, W * 12 / 16); cudaEventRecord(b); CK(cudaEventSynchronize(b));
            cudaEventElapsedTime(&ms, a, b); if (rep == 2) report("tile copy (33 of 34 rows), CTAs", g, ms);
        }
    }
    return 0;
}
