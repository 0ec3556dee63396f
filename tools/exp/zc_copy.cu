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
// d) bulk (TMA) stores: each CTA stages CHUNK bytes in shared memory with ordinary loads, then ONE thread issues
//    cp.async.bulk.global.shared::cta for the whole chunk (UBLKCP): does the async-proxy path packetise better on PCIe?
template <int CHUNK>
__global__ void k_bulk(const float4* __restrict__ src, float4* __restrict__ dst, size_t n4) {
    extern __shared__ __align__(128) unsigned char smem[];
    float4* buf = reinterpret_cast<float4*>(smem);
    const size_t per = CHUNK / 16, n_chunks = n4 / per;
    for (size_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        for (int k = threadIdx.x; k < (int)per; k += blockDim.x) buf[k] = src[c * per + k];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned s_addr = (unsigned)__cvta_generic_to_shared(buf);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(dst + c * per), "r"(s_addr), "r"(CHUNK) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncthreads();
    }
}
// e) the same with 384-byte rows at a frame pitch (the tile pattern): 32 bulk stores per tile
__global__ void k_bulk_tiles(const float4* __restrict__ src, float4* __restrict__ dst, int tiles_x, int tiles_y, int pitch4) {
    __shared__ __align__(128) float4 buf[32 * 24];
    for (int t = blockIdx.x; t < tiles_x * tiles_y; t += gridDim.x) {
        const int ty = t / tiles_x, tx = t - ty * tiles_x;
        const size_t base = (size_t)ty * 32 * pitch4 + (size_t)tx * 24;
        for (int k = threadIdx.x; k < 32 * 24; k += blockDim.x) buf[k] = src[base + (size_t)(k / 24) * pitch4 + (k % 24)];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x < 32) {
            unsigned s_addr = (unsigned)__cvta_generic_to_shared(buf + threadIdx.x * 24);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 384;" :: "l"(dst + base + (size_t)threadIdx.x * pitch4), "r"(s_addr) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncthreads();
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
    {   // f) DMA engine and SM stores AT THE SAME TIME, each moving a share of the frame: does the link take more than either alone?
        cudaStream_t s1, s2; cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
        cudaEvent_t e1, e2; cudaEventCreate(&e1); cudaEventCreate(&e2);
        for (int pct : {30, 50, 70}) {
            const size_t n_dma = (n4 * pct / 100) & ~(size_t)63;
            for (int rep = 0; rep < 3; ++rep) {
                cudaDeviceSynchronize();
                cudaEventRecord(a, s1);
                cudaStreamWaitEvent(s2, a, 0);
                CK(cudaMemcpyAsync(h, d, n_dma * 16, cudaMemcpyDeviceToHost, s1));
                k_flat<<<37, 256, 0, s2>>>(d + n_dma, hd + n_dma, n4 - n_dma);
                cudaEventRecord(e2, s2); cudaStreamWaitEvent(s1, e2, 0);
                cudaEventRecord(b, s1); CK(cudaEventSynchronize(b));
                cudaEventElapsedTime(&ms, a, b); if (rep == 2) report("DMA (pct of bytes) + SM stores", pct, ms);
            }
        }
    }
    for (int g : {8, 37, 148, 296}) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(a); k_bulk<16384><<<g, 256, 16384>>>(d, hd, n4); cudaEventRecord(b); CK(cudaEventSynchronize(b));
            cudaEventElapsedTime(&ms, a, b); if (rep == 2) report("bulk store 16 KB chunks, CTAs", g, ms);
        }
    }
    for (int g : {8, 37, 148, 296}) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(a); k_bulk_tiles<<<g, 256>>>(d, hd, W / 32, (H + 31) / 32 - 1, W * 12 / 16); cudaEventRecord(b); CK(cudaEventSynchronize(b));
            cudaEventElapsedTime(&ms, a, b); if (rep == 2) report("bulk store tiles (33/34 rows), CTAs", g, ms);
        }
    }
    return 0;
}
