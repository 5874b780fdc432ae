// Measured evidence for BASELINE.json north_star (2): is a tensor-core contraction a better way than the register radix FFT to do
// a 32-point DFT stage of the extractor's 1024-point transform?  (SURVEY 7.2-3 argued it on paper; this measures it.)
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --expt-relaxed-constexpr -cudart shared -I../../seld_b200/csrc tc_dft32.cu -o tc_dft32
//   ./tc_dft32 [reg|tc_floor|tc_full|check]
//
// One "stage" = what one warp does in stage 2 of a frame: 32 columns x a 32-point complex DFT, data in shared memory.
//   reg       the kernel's own code path: stage2_load_fft (32 LDS.64 + packed-FP32 radix-2 FFT in registers) + store, 16 warps per SM
//   tc_floor  the same contraction as a real GEMM on the 5th-generation tensor cores: D[64 x 32] = A[64 x 64] B[64 x 32] with
//             A = [[Wr, -Wi], [Wi, Wr]] (the DFT matrix), B = [Er; Ei].  The 1e-4 dB log-mel tolerance needs ~22 mantissa bits, so
//             every product is a 3xTF32 split (A_hi B_hi + A_hi B_lo + A_lo B_hi): 3 x 8 tcgen05.mma.kind::tf32 (M64 N32 K8) per
//             stage.  tc_floor issues ONLY the MMAs, back to back, four accumulators in flight: the tensor pipe's own floor.
//   tc_full   + what a kernel would have to do around them: write B as a hi / lo split in the UMMA shared-memory layout
//             (2 x 8 KB per stage instead of the 8 KB exchange buffer) and read D back from tensor memory (tcgen05.ld).
//   check     numerical error of both against a float64 DFT.
// Result (B200, profiles/r2_microbench_tc_dft32.txt): the register FFT wins by a wide margin; the tensor pipe alone needs more
// cycles per stage than the whole CUDA-core stage.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

#include "extract_core.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

using namespace seld;

// ------------------------------------------------------------------------------------------------ register FFT (the kernel's path)
__global__ void __launch_bounds__(512) reg_kernel(const float2* in, float2* out, int iters) {
    extern __shared__ __align__(16) unsigned char sm[];
    using G = Geo<32>;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float2* E = reinterpret_cast<float2*>(sm) + warp * G::E_ELEMS;
    for (int i = lane; i < 32 * 32; i += 32) E[(i >> 5) * G::EP + (i & 31)] = in[(blockIdx.x * 16 + warp) * 1024 + i];
    __syncwarp();
    for (int it = 0; it < iters; ++it) {
        float2 u[32];
        stage2_load_fft<32>(E, u, lane);
        __syncwarp();
#pragma unroll
        for (int p = 0; p < 32; ++p) E[bitrev(p, 5) * G::EP + lane] = make_float2(u[p].x * 0.03125f, u[p].y * 0.03125f);   // (scaled: stays bounded)
        __syncwarp();
    }
    for (int i = lane; i < 32 * 32; i += 32) out[(blockIdx.x * 16 + warp) * 1024 + i] = E[(i >> 5) * G::EP + (i & 31)];
}

// ------------------------------------------------------------------------------------------------ tensor-core DFT
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ unsigned long long umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (unsigned long long)((saddr >> 4) & 0x3FFF) | ((unsigned long long)((lbo >> 4) & 0x3FFF) << 16) |
           ((unsigned long long)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_tf32_ss(uint32_t d, unsigned long long da, unsigned long long db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
// kind::tf32 instruction descriptor: F32 accumulate (bit 4), A / B format TF32 = 2 (bits 7-9, 10-12), K-major, N >> 3 at 17, M >> 4 at 24
constexpr uint32_t kIdescTf32 = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((64u >> 4) << 24);
constexpr int A_LBO = 128, A_SBO = 16 * 128;        // A [64 rows][64 K] fp32: [8 row groups][16 K units][8 rows][16 B]
constexpr int B_LBO = 128, B_SBO = 16 * 128;        // B [32 rows][64 K]
constexpr int A_BYTES = 64 * 64 * 4, B_BYTES = 32 * 64 * 4;

struct TcArgs { const float* a_hi; const float* a_lo; const float2* in; float* out; int iters; int full; };

// 128 threads.  Warp w = TMEM quadrant w.  D (M = 64): row m -> lane (m % 16) + 32 (m / 16), 32 columns per accumulator slot.
__global__ void __launch_bounds__(128) tc_kernel(TcArgs g) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ __align__(8) unsigned long long s_bar[4];
    __shared__ uint32_t s_tmem;
    unsigned char* sAh = sm;
    unsigned char* sAl = sm + A_BYTES;
    unsigned char* sB = sm + 2 * A_BYTES;           // 4 slots x (hi, lo)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid < 4) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar[tid]))); asm volatile("fence.mbarrier_init.release.cluster;"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    // A (hi, lo) in the UMMA layout: element (m, k) at (m >> 3) * SBO + (k >> 2) * LBO + (m & 7) * 16 + (k & 3) * 4
    for (int e = tid; e < 64 * 64; e += 128) {
        const int m = e >> 6, k = e & 63;
        const int off = (m >> 3) * A_SBO + (k >> 2) * A_LBO + (m & 7) * 16 + (k & 3) * 4;
        *reinterpret_cast<float*>(sAh + off) = g.a_hi[e];
        *reinterpret_cast<float*>(sAl + off) = g.a_lo[e];
    }
    // B slots: column n = k2 (row of the K-major operand), K = (n1 re | 32 + n1 im); input E[n1][k2]
    auto write_b = [&](int slot, const float2* E) {
        for (int e = tid; e < 32 * 32; e += 128) {
            const int n1 = e >> 5, k2 = e & 31;
            const float2 v = E[e];
#pragma unroll
            for (int part = 0; part < 2; ++part) {
                const float x = part ? v.y : v.x;
                const int k = n1 + 32 * part;
                const float hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);      // the 10-bit-mantissa part the tensor core keeps
                const int off = (k2 >> 3) * B_SBO + (k >> 2) * B_LBO + (k2 & 7) * 16 + (k & 3) * 4;
                *reinterpret_cast<float*>(sB + (2 * slot) * B_BYTES + off) = hi;
                *reinterpret_cast<float*>(sB + (2 * slot + 1) * B_BYTES + off) = x - hi;
            }
        }
    };
    const float2* myin = g.in + (size_t)blockIdx.x * 1024;
    for (int s = 0; s < 4; ++s) write_b(s, myin);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = s_tmem;
    auto issue = [&](int slot) {            // 24 MMAs: 3 split products x 8 K steps of 8
        const uint32_t ah = smem_u32(sAh), al = smem_u32(sAl), bh = smem_u32(sB + (2 * slot) * B_BYTES), bl = bh + B_BYTES;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            mma_tf32_ss(tmem + 32 * slot, umma_desc(ah + j * 2 * A_LBO, A_LBO, A_SBO), umma_desc(bh + j * 2 * B_LBO, B_LBO, B_SBO), kIdescTf32, j > 0);
            mma_tf32_ss(tmem + 32 * slot, umma_desc(ah + j * 2 * A_LBO, A_LBO, A_SBO), umma_desc(bl + j * 2 * B_LBO, B_LBO, B_SBO), kIdescTf32, 1u);
            mma_tf32_ss(tmem + 32 * slot, umma_desc(al + j * 2 * A_LBO, A_LBO, A_SBO), umma_desc(bh + j * 2 * B_LBO, B_LBO, B_SBO), kIdescTf32, 1u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar[slot])) : "memory");
    };
    uint32_t par[4] = {0, 0, 0, 0};
    float keep = 0.f;
    if (!g.full) {
        // tensor-pipe floor: nothing but the MMAs, four accumulators in flight
        if (warp == 0 && elect_one()) {
            for (int it = 0; it < g.iters; ++it) {
                const int slot = it & 3;
                if (it >= 4) { mbar_wait(smem_u32(&s_bar[slot]), par[slot]); par[slot] ^= 1; }
                issue(slot);
            }
            for (int s = 0; s < 4 && s < g.iters; ++s) { mbar_wait(smem_u32(&s_bar[s]), par[s]); par[s] ^= 1; }
        }
    } else {
        // every stage: operand split written to shared memory, MMAs, accumulator read back (lanes 0..15 of each quadrant, 32 columns)
        for (int it = 0; it < g.iters; ++it) {
            const int slot = it & 3;
            write_b(slot, myin);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;");
            __syncthreads();
            if (warp == 0 && elect_one()) { asm volatile("tcgen05.fence::after_thread_sync;"); issue(slot); }
            mbar_wait(smem_u32(&s_bar[slot]), par[slot]);
            par[slot] ^= 1;
            asm volatile("tcgen05.fence::after_thread_sync;");
            float v[32];
            tmem_ld16(tmem + ((32u * warp) << 16) + 32 * slot, v);
            tmem_ld16(tmem + ((32u * warp) << 16) + 32 * slot + 16, v + 16);
#pragma unroll
            for (int i = 0; i < 32; ++i) keep += v[i];
            asm volatile("tcgen05.fence::before_thread_sync;");
        }
    }
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    // dump accumulator slot 0 (or the last one used): D row m = 16 q + (lane & 15) for lanes < 16
    {
        const int slot = g.full ? ((g.iters - 1) & 3) : 0;
        float v[32];
        tmem_ld16(tmem + ((32u * warp) << 16) + 32 * slot, v);
        tmem_ld16(tmem + ((32u * warp) << 16) + 32 * slot + 16, v + 16);
        if (lane < 16)
            for (int n = 0; n < 32; ++n) g.out[((size_t)blockIdx.x * 64 + 16 * warp + lane) * 32 + n] = v[n] + 0.f * keep;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
}

int main(int argc, char** argv) {
    const char* what = argc > 1 ? argv[1] : "check";
    const int n_sm = 148;
    std::vector<float2> in((size_t)n_sm * 16 * 1024);
    unsigned s = 777u;
    for (auto& v : in) {
        s = s * 1664525u + 1013904223u; v.x = float(int(s >> 9) & 0xffff) / 32768.0f - 1.0f;
        s = s * 1664525u + 1013904223u; v.y = float(int(s >> 9) & 0xffff) / 32768.0f - 1.0f;
    }
    // A = [[Wr, -Wi], [Wi, Wr]], W[k1][n1] = exp(-2 pi i k1 n1 / 32); hi / lo split for the 3xTF32 product
    std::vector<float> a_hi(64 * 64), a_lo(64 * 64);
    for (int m = 0; m < 64; ++m)
        for (int k = 0; k < 64; ++k) {
            const int k1 = m & 31, n1 = k & 31;
            const double ang = -2.0 * M_PI * double((k1 * n1) % 32) / 32.0;
            const double wr = cos(ang), wi = sin(ang);
            const double val = (m < 32) ? (k < 32 ? wr : -wi) : (k < 32 ? wi : wr);
            const float f = float(val);
            uint32_t bits; memcpy(&bits, &f, 4); bits &= 0xffffe000u;
            float hi; memcpy(&hi, &bits, 4);
            a_hi[m * 64 + k] = hi;
            a_lo[m * 64 + k] = float(val - double(hi));
        }
    float2 *d_in, *d_out2; float *d_ah, *d_al, *d_out;
    CK(cudaMalloc(&d_in, in.size() * 8)); CK(cudaMalloc(&d_out2, in.size() * 8)); CK(cudaMalloc(&d_out, (size_t)n_sm * 64 * 32 * 4));
    CK(cudaMalloc(&d_ah, 64 * 64 * 4)); CK(cudaMalloc(&d_al, 64 * 64 * 4));
    CK(cudaMemcpy(d_in, in.data(), in.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_ah, a_hi.data(), 64 * 64 * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_al, a_lo.data(), 64 * 64 * 4, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reg_smem = 16 * Geo<32>::E_ELEMS * 8, tc_smem = 2 * A_BYTES + 8 * B_BYTES + 1024;
    CK(cudaFuncSetAttribute(reg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, reg_smem));
    CK(cudaFuncSetAttribute(tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem));
    float ms = 0.f;
    if (!strcmp(what, "reg")) {
        const int iters = 2000;
        reg_kernel<<<n_sm, 512, reg_smem>>>(d_in, d_out2, 10);
        cudaEventRecord(e0); reg_kernel<<<n_sm, 512, reg_smem>>>(d_in, d_out2, iters); cudaEventRecord(e1);
        CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms, e0, e1);
        const double stages_per_sm = 16.0 * iters;
        printf("reg      : %.3f ms for %d stages per warp, 16 warps/SM -> %.1f ns per stage per SM (%.0f SM-cycles at 1.965 GHz)\n", ms, iters,
               ms * 1e6 / stages_per_sm, ms * 1e6 / stages_per_sm * 1.965);
    } else if (!strcmp(what, "tc_floor") || !strcmp(what, "tc_full")) {
        const int full = !strcmp(what, "tc_full"), iters = 4000;
        TcArgs g{d_ah, d_al, d_in, d_out, 8, full};
        tc_kernel<<<n_sm, 128, tc_smem>>>(g);
        g.iters = iters;
        cudaEventRecord(e0); tc_kernel<<<n_sm, 128, tc_smem>>>(g); cudaEventRecord(e1);
        CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms, e0, e1);
        printf("%-9s: %.3f ms for %d stages per SM (24 tcgen05.mma.kind::tf32 M64 N32 K8 each) -> %.1f ns per stage per SM (%.0f SM-cycles)\n", what, ms, iters,
               ms * 1e6 / iters, ms * 1e6 / iters * 1.965);
    } else {
        // numerics: one stage both ways against a float64 DFT of block 0 (the tensor-core kernel transforms its block's first 32x32)
        TcArgs g{d_ah, d_al, d_in, d_out, 1, 1};
        tc_kernel<<<n_sm, 128, tc_smem>>>(g);
        reg_kernel<<<n_sm, 512, reg_smem>>>(d_in, d_out2, 1);
        CK(cudaDeviceSynchronize());
        std::vector<float> tc(64 * 32); std::vector<float2> rg(1024);
        CK(cudaMemcpy(tc.data(), d_out, 64 * 32 * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(rg.data(), d_out2, 1024 * 8, cudaMemcpyDeviceToHost));
        double e_tc = 0, e_rg = 0, ref_max = 0;
        for (int k1 = 0; k1 < 32; ++k1)
            for (int k2 = 0; k2 < 32; ++k2) {
                double yr = 0, yi = 0;
                for (int n1 = 0; n1 < 32; ++n1) {
                    const double ang = -2.0 * M_PI * double((k1 * n1) % 32) / 32.0, xr = in[n1 * 32 + k2].x, xi = in[n1 * 32 + k2].y;
                    yr += xr * cos(ang) - xi * sin(ang);
                    yi += xr * sin(ang) + xi * cos(ang);
                }
                ref_max = fmax(ref_max, fmax(fabs(yr), fabs(yi)));
                e_tc = fmax(e_tc, fmax(fabs(tc[k1 * 32 + k2] - yr), fabs(tc[(32 + k1) * 32 + k2] - yi)));
                e_rg = fmax(e_rg, fmax(fabs(rg[k1 * 32 + k2].x * 32.0 - yr), fabs(rg[k1 * 32 + k2].y * 32.0 - yi)));
            }
        printf("check    : max |error| vs float64 DFT, relative to the largest output (%.2f): 3xTF32 tensor core %.2e, packed-FP32 register FFT %.2e\n",
               ref_max, e_tc / ref_max, e_rg / ref_max);
    }
    return 0;
}
