// Experiment: device->pinned-host copy of a 1920x1080x3 float frame as 1 contiguous copy, as 12 row bands,
// and as 8x6 rectangular regions (cudaMemcpy2DAsync).   nvcc -O2 -o copy2d copy2d.cu
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
int main() {
    const int W = 1920, H = 1080; const size_t pitch = (size_t)W * 12, bytes = pitch * H;
    float *d, *h; cudaMalloc(&d, bytes); cudaHostAlloc(&h, bytes, cudaHostAllocDefault);
    cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    auto run = [&](const char* name, int by, int bx) {
        for (int rep = 0; rep < 3; ++rep) {
            auto t0 = std::chrono::steady_clock::now();
            for (int it = 0; it < 20; ++it) {
                for (int j = 0; j < by; ++j) for (int i = 0; i < bx; ++i) {
                    int y0 = H * j / by, y1 = H * (j + 1) / by, x0 = W * i / bx, x1 = W * (i + 1) / bx;
                    size_t off = (size_t)y0 * pitch + (size_t)x0 * 12;
                    if (bx == 1) cudaMemcpyAsync((char*)h + off, (char*)d + off, (size_t)(y1 - y0) * pitch, cudaMemcpyDeviceToHost, s);
                    else cudaMemcpy2DAsync((char*)h + off, pitch, (char*)d + off, pitch, (size_t)(x1 - x0) * 12, y1 - y0, cudaMemcpyDeviceToHost, s);
                }
                cudaStreamSynchronize(s);
            }
            double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / 20;
            if (rep == 2) printf("%-28s %7.1f us/frame  %.1f GB/s\n", name, us, bytes / us / 1e3);
        }
    };
    run("1 contiguous", 1, 1); run("12 row bands", 12, 1); run("34 row bands", 34, 1); run("8x6 regions (2D)", 8, 6); run("8x2 regions (2D)", 8, 2); run("1x6 columns (2D)", 1, 6);
    return 0;
}
