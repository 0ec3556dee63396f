// Micro-benchmark: scalar FFMA against packed FFMA2 (fma.rn.f32x2, sm_100) issue throughput.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu && ./ffma2
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_scalar(float* out, int iters, float a, float b) {
    float x[8];
    for (int k = 0; k < 8; ++k) x[k] = threadIdx.x + k;
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = __fmaf_rn(x[k], a, b);
    float s = 0; for (int k = 0; k < 8; ++k) s += x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_packed(float* out, int iters, float a, float b) {
    float2 x[4];
    for (int k = 0; k < 4; ++k) x[k] = make_float2(threadIdx.x + 2 * k, threadIdx.x + 2 * k + 1);
    const float2 aa = make_float2(a, a), bb = make_float2(b, b);
    for (int i = 0; i < iters; ++i)
#pragma unroll
        for (int k = 0; k < 4; ++k) x[k] = __ffma2_rn(x[k], aa, bb);
    float s = 0; for (int k = 0; k < 4; ++k) s += x[k].x + x[k].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 200000;
    for (int rep = 0; rep < 2; ++rep) {
        float ms;
        cudaEventRecord(e0); k_scalar<<<148 * 8, 256>>>(d, iters, 1.0001f, 0.5f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double flops = 2.0 * 8 * iters * 148.0 * 8 * 256;
        printf("scalar FFMA : %.3f ms  %.1f TFLOP/s\n", ms, flops / ms / 1e9);
        cudaEventRecord(e0); k_packed<<<148 * 8, 256>>>(d, iters, 1.0001f, 0.5f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("packed FFMA2: %.3f ms  %.1f TFLOP/s\n", ms, flops / ms / 1e9);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
