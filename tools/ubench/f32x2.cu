// Issue-rate microbenchmark: FFMA vs FFMA2 (packed fp32x2) on sm_100a.  nvcc -arch=sm_100a -o f32x2 f32x2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k_ffma(float* out, int iters) {
    float a[8], b = 1.0001f, c = 0.5f;
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], b, c);
    float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float* out, int iters) {
    uint64_t a[8], b, c;
    asm("mov.b64 %0, {%1,%2};" : "=l"(b) : "f"(1.0001f), "f"(1.0002f));
    asm("mov.b64 %0, {%1,%2};" : "=l"(c) : "f"(0.5f), "f"(0.25f));
    for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1,%2};" : "=l"(a[i]) : "f"((float)threadIdx.x + i), "f"((float)i));
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(b), "l"(c));
    float s = 0;
    for (int i = 0; i < 8; ++i) { float x, y; asm("mov.b64 {%0,%1}, %2;" : "=f"(x), "=f"(y) : "l"(a[i])); s += x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int which = 0; which < 2; ++which) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (which == 0) k_ffma<<<148 * 4, 512>>>(d, iters); else k_ffma2<<<148 * 4, 512>>>(d, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double inst = 148.0 * 4 * 512 * (double)iters * 8;
        printf("%s: %.3f ms, %.1f G thread-instr/s, %.1f TFLOP/s\n", which ? "FFMA2" : "FFMA", ms, inst / ms / 1e6,
               inst * (which ? 4 : 2) / ms / 1e9);
    }
    return 0;
}
