// Micro-benchmark: can the warp scheduler issue another instruction in the second cycle a packed FFMA2 occupies the FP32 pipe?
// Per loop iteration a thread runs 8 independent FFMA2 (or 16 FFMA) plus K "filler" instructions of another pipe
// (LOP3 on the ALU pipe, LDS on the LSU, MOV).  If fillers are free up to 8 per iteration, FFMA2 leaves its second cycle to
// them and the extractor's floor is max(issue slots, FP32 pipe cycles); if time grows from the first filler on, packed
// instructions hold the issue port for both cycles and the floor is issue slots + packed instructions.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared coissue.cu -o coissue && ./coissue
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ float fma1(float a, float b, float c) {
    float r;
    asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ unsigned lop(unsigned a, unsigned b) {
    unsigned r;
    asm volatile("{ .reg .b32 t; shl.b32 t, %1, 5; xor.b32 %0, t, %2; }" : "=r"(r) : "r"(a), "r"(b));   // SHF + LOP3 (ptxas folds plain xor chains)
    return r;
}
__device__ __forceinline__ unsigned lds(unsigned addr) {
    unsigned r;
    asm volatile("ld.volatile.shared.b32 %0, [%1];" : "=r"(r) : "r"(addr) : "memory");
    return r;
}

// PACKED: 8 FFMA2 per iteration, else 16 FFMA; FILL: 0 = LOP3, 1 = LDS; K fillers per iteration
template <bool PACKED, int FILL, int K>
__global__ void k(float* out, int iters, float s, unsigned seed) {
    __shared__ unsigned sm[1024];
    sm[threadIdx.x & 1023] = seed;
    __syncthreads();
    float a[16];
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
    unsigned x[8];
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
    const unsigned saddr = static_cast<unsigned>(__cvta_generic_to_shared(sm)) + 4 * (threadIdx.x & 1023);
    const float m = s, c = s * 0.5f;
    unsigned long long* p = reinterpret_cast<unsigned long long*>(a);
    float2 m2 = make_float2(m, m), c2 = make_float2(c, c);
    const unsigned long long mm = *reinterpret_cast<unsigned long long*>(&m2), cc = *reinterpret_cast<unsigned long long*>(&c2);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (PACKED) p[i] = fma2(p[i], mm, cc);
            else { a[2 * i] = fma1(a[2 * i], m, c); a[2 * i + 1] = fma1(a[2 * i + 1], m, c); }
#pragma unroll
            for (int j = 0; j < (K + 7 - i) / 8; ++j) {
                if (FILL == 0) x[i] = lop(x[i], seed);
                else x[i] += lds(saddr + 4 * ((i + 8 * j + it) & 31));
            }
        }
    }
    float r = 0;
    for (int i = 0; i < 16; ++i) r += a[i];
    unsigned xr = 0;
    for (int i = 0; i < 8; ++i) xr ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r + float(xr);
}

template <bool PACKED, int FILL, int K>
static void run(float* out, int threads, const char* name) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 20000;
    float ms = 0;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        k<PACKED, FILL, K><<<148, threads>>>(out, iters, 0.999f, 12345u);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    }
    // cycles per iteration per scheduler (4 per SM): warps per scheduler = threads / 128
    const double cyc = ms * 1e-3 * 1.965e9 / iters / (threads / 128.0);
    const int fill_slots = FILL == 0 ? 2 * K : K;      // a LOP3 filler is SHF + LOP3
    printf("%-6s %s filler instructions %2d: %.3f ms  %.2f cycles / (warp iteration)  [FP32 pipe: 16, issue slots: %d]\n", name,
           FILL == 0 ? "ALU " : "LDS ", fill_slots, ms, cyc, (PACKED ? 8 : 16) + fill_slots);
}

int main() {
    float* out;
    cudaMalloc(&out, 148 * 1024 * 4);
    const int threads = 512;
#define ROW(P, F, NAME) run<P, F, 0>(out, threads, NAME); run<P, F, 2>(out, threads, NAME); run<P, F, 4>(out, threads, NAME); \
                        run<P, F, 6>(out, threads, NAME); run<P, F, 8>(out, threads, NAME); run<P, F, 12>(out, threads, NAME);
    ROW(true, 0, "FFMA2") ROW(false, 0, "FFMA") ROW(true, 1, "FFMA2") ROW(false, 1, "FFMA")
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return 0;
}
