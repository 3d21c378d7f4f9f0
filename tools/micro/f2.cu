#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pk(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(unsigned long long v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) { unsigned long long d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__global__ void k_fma2(float* out, int iters) {
    unsigned long long a[8];
    for (int k = 0; k < 8; ++k) a[k] = pk(threadIdx.x * 1e-3f + k, threadIdx.x * 2e-3f + k);
    const unsigned long long m = pk(0.999f, 0.998f), b = pk(1e-3f, 2e-3f);
    for (int i = 0; i < iters; ++i) {
        #pragma unroll
        for (int r = 0; r < 16; ++r)
            #pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = fma2(a[k], m, b);
    }
    float s = 0.f;
    for (int k = 0; k < 8; ++k) { float x, y; upk(a[k], x, y); s += x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_fma1(float* out, int iters) {
    float a[8];
    for (int k = 0; k < 8; ++k) a[k] = threadIdx.x * 1e-3f + k;
    const float m = 0.999f, b = 1e-3f;
    for (int i = 0; i < iters; ++i) {
        #pragma unroll
        for (int r = 0; r < 16; ++r)
            #pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = fmaf(a[k], m, b);
    }
    float s = 0.f; for (int k = 0; k < 8; ++k) s += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// 3 distinct varying operands
__global__ void k_fma3(float* out, int iters) {
    float a[8], x[8], y[8];
    for (int k = 0; k < 8; ++k) { a[k] = threadIdx.x * 1e-3f + k; x[k] = 0.999f + k * 1e-4f; y[k] = 1e-3f * k; }
    for (int i = 0; i < iters; ++i) {
        #pragma unroll
        for (int r = 0; r < 16; ++r)
            #pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = fmaf(a[k], x[(k + r) & 7], y[(k + 3 * r) & 7]);
    }
    float s = 0.f; for (int k = 0; k < 8; ++k) s += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    int dev = 0; cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 4000;
    float* d; cudaMalloc(&d, blocks * threads * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int which = 0; which < 3; ++which) for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        if (which == 0) k_fma1<<<blocks, threads>>>(d, iters); else if (which == 1) k_fma2<<<blocks, threads>>>(d, iters); else k_fma3<<<blocks, threads>>>(d, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * 8 * 16 * (double)iters * blocks * threads * (which == 1 ? 2 : 1);
        if (rep == 2) printf("%s: %.3f ms %.2f TFLOP/s\n", which == 0 ? "FFMA (2 const operands)" : which == 1 ? "FFMA2 f32x2" : "FFMA 3 varying operands", ms, fl / ms * 1e-9);
    }
    return 0;
}
