// Micro-benchmark: issue rate of packed FFMA2 vs scalar FFMA on sm_100a (decides whether the FFT butterflies
// should use add/mul/fma.f32x2).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared f32x2.cu -o f32x2 && ./f32x2
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

template <int MODE>
__global__ void k(float* out, int iters, float s) {
    float a[16];
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
    float m = s, c = s * 0.5f;
    if (MODE == 0) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], m, c);
        }
    } else {
        unsigned long long* p = reinterpret_cast<unsigned long long*>(a);
        unsigned long long mm, cc;
        float2 m2 = make_float2(m, m), c2 = make_float2(c, c);
        mm = *reinterpret_cast<unsigned long long*>(&m2);
        cc = *reinterpret_cast<unsigned long long*>(&c2);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], mm, cc);
        }
    }
    float r = 0;
    for (int i = 0; i < 16; ++i) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

int main() {
    float* out;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 20000;
    for (int threads : {128, 256, 512, 1024}) {
        for (int mode = 0; mode < 2; ++mode) {
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                if (mode == 0) k<0><<<148, threads>>>(out, iters, 0.999f);
                else k<1><<<148, threads>>>(out, iters, 0.999f);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
            }
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            double fma_per_clk_sm = double(iters) * 16 * threads / (ms * 1e-3 * 1.965e9);
            printf("threads/SM %4d  %s  %.3f ms  -> %.1f FMA lanes/clk/SM (at 1965 MHz)\n", threads, mode ? "FFMA2" : "FFMA ", ms,
                   fma_per_clk_sm);
        }
    }
    return 0;
}
