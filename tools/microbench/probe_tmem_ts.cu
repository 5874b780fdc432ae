// Probe: where does tcgen05.mma put / expect its operands for the shapes the fused GCC-PHAT path wants?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared probe_tmem_ts.cu -o probe_tmem_ts
//   ./probe_tmem_ts <test>         (one test per process: a faulting variant must not poison the others)
//
// Questions (undocumented in the guides of this image; answered by dumping tensor memory):
//   test 0  SS  M=64  N=8   A, B from shared memory (no-swizzle K-major, B with a padded 144-byte K pitch): D row -> TMEM lane?
//   test 1  TS  M=64  N=8   A in TMEM lanes (m % 16) + 32 (m / 16): accepted? element k -> column k/2, half k%2?
//   test 2  TS  M=64  N=8   A and D both in the UPPER half-subpartitions (lane field 16): a second, interleaved atom?
//   test 3  TS  M=128 N=16  A row m in TMEM lane m
//   test 4  SS  M=64  N=8   D in the upper half-subpartitions
//   test 5  TS  M=64  N=8   A lower, D upper (is datapath equivalence of A and D required?)
//   test 6  SS  M=64  N=16  B rows 8..15 in a second 8-row group SBO bytes away (two frame teams in one MMA)
//   test 7  timing: cycles per TS MMA (M=64, K=16) for N = 8, 16, 24, 48, issued back to back by one thread
//   test 8  timing: the same with the MMAs dealt round-robin over 1, 2, 4, 8 independent accumulators (is test 7 a
//           dependent-accumulate latency or an issue / pipe cost?), and with 4 threads of 4 warps issuing concurrently
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

constexpr int KK = 64;                  // contraction length of the layout tests (4 MMAs of K = 16)
constexpr int A_COL = 128, D_COL = 64;  // TMEM columns of the A operand / the accumulator
constexpr int LBO_A = 128, SBO_A = 1024;
constexpr int LBO_B = 144, SBO_B = 8 * 144;

struct Args {
    const __half* A;     // [128][KK] row-major
    const __half* B;     // [16][KK]
    float* out;          // [128 lanes][16 columns] dump of the accumulator region
    long long* cycles;   // test 7
    int test;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ unsigned long long umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    unsigned long long d = 0;
    d |= (unsigned long long)((saddr >> 4) & 0x3FFF);
    d |= (unsigned long long)((lbo >> 4) & 0x3FFF) << 16;
    d |= (unsigned long long)((sbo >> 4) & 0x3FFF) << 32;
    d |= 1ull << 46;
    return d;
}
__device__ __forceinline__ void mma_ss(uint32_t d, unsigned long long da, unsigned long long db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc));
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, unsigned long long db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a), "l"(db), "r"(idesc), "r"(acc));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
                    "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
                 "tcgen05.wait::ld.sync.aligned;"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}

__device__ __forceinline__ bool elect_one() {      // one lane of a converged warp; lets ptxas issue UTCHMMA without a per-thread waterfall loop
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) { return (1u << 4) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24); }

__global__ void __launch_bounds__(128) probe(Args g) {
    __shared__ __align__(1024) unsigned char sA[16 * SBO_A];          // 128 rows x 64 halfs, no swizzle
    __shared__ __align__(1024) unsigned char sB[4 * SBO_B + 8192];    // 16+ rows x 64 halfs, 144-byte K pitch (+ slack for the timing test)
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = s_tmem;
    const uint32_t tq = tmem + ((uint32_t)(32 * warp) << 16);          // this warp's lane quadrant

    // operands -> shared memory
    for (int e = tid; e < 128 * (KK / 8); e += 128) {
        const int r = e / (KK / 8), j = e % (KK / 8);
        *reinterpret_cast<uint4*>(sA + (r >> 3) * SBO_A + j * LBO_A + (r & 7) * 16) = *reinterpret_cast<const uint4*>(g.A + r * KK + j * 8);
    }
    for (int e = tid; e < 16 * (KK / 8); e += 128) {
        const int r = e / (KK / 8), j = e % (KK / 8);
        *reinterpret_cast<uint4*>(sB + (r >> 3) * SBO_B + j * LBO_B + (r & 7) * 16) = *reinterpret_cast<const uint4*>(g.B + r * KK + j * 8);
    }
    // tensor memory: zero everything, then the A operand of the TS tests
    {
        uint32_t z[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) z[i] = 0u;
        for (int c = 0; c < 512; c += 16) tmem_st16(tq + c, z);
        int row = -1;
        if (g.test == 1 || g.test == 5 || g.test == 7) row = (lane < 16) ? 16 * warp + lane : -1;
        if (g.test == 8) row = 16 * warp + (lane & 15);
        if (g.test == 2) row = (lane >= 16) ? 16 * warp + lane - 16 : -1;
        if (g.test == 3) row = 32 * warp + lane;
        uint32_t w[KK / 2];
#pragma unroll
        for (int c = 0; c < KK / 2; ++c) w[c] = (row >= 0) ? *reinterpret_cast<const uint32_t*>(g.A + row * KK + 2 * c) : 0u;
        if (g.test == 1 || g.test == 2 || g.test == 3 || g.test == 5 || g.test == 7 || g.test == 8) {
            tmem_st16(tq + A_COL, w);
            tmem_st16(tq + A_COL + 16, w + 16);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");

    if (g.test == 8) {
        const uint32_t b0 = smem_u32(sB);
        const uint32_t id = idesc_f16(64, 8);
        if (warp == 0 && elect_one()) {
            const int nacc[4] = {1, 2, 4, 8};
            uint32_t parity = 0;
            for (int t = 0; t < 5; ++t) {                 // t == 4: 8 accumulators, alternating lower / upper atoms
                const int na = nacc[t < 4 ? t : 3];
                const long long c0 = clock64();
                for (int rep = 0; rep < 1024; ++rep) {
                    const int acc_i = rep % na;
                    const uint32_t up = (t == 4 && (rep & 1)) ? (16u << 16) : 0u;
                    mma_ts(tmem + up + D_COL + 8 * acc_i, tmem + up + A_COL + 8 * (rep & 3), umma_desc(b0 + (rep & 3) * 2 * LBO_B, LBO_B, SBO_B), id,
                           rep >= na ? 1u : 0u);
                }
                const long long c1 = clock64();
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar)));
                mbar_wait(smem_u32(&s_bar), parity);
                parity ^= 1;
                const long long c2 = clock64();
                g.cycles[2 * t] = c1 - c0;
                g.cycles[2 * t + 1] = c2 - c0;
            }
        }
        __syncthreads();
        // four issuing threads (lane 0 of each warp), each with its own accumulator, 256 MMAs each
        __shared__ __align__(8) unsigned long long s_bar4[4];
        if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar4[warp])));
        asm volatile("fence.mbarrier_init.release.cluster;");
        __syncthreads();
        const long long c0 = clock64();
        if (elect_one()) {
            for (int rep = 0; rep < 256; ++rep)
                mma_ts(tmem + D_COL + 8 * warp, tmem + A_COL + 8 * (rep & 3), umma_desc(b0 + (rep & 3) * 2 * LBO_B, LBO_B, SBO_B), id, rep ? 1u : 0u);
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar4[warp])));
            mbar_wait(smem_u32(&s_bar4[warp]), 0);
        }
        __syncthreads();
        if (tid == 0) g.cycles[10] = clock64() - c0;
        g.cycles[11] = 0;
    }
    if (warp == 0 && g.test != 8 && elect_one()) {
        const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
        if (g.test == 7) {
            const int ns[4] = {8, 16, 24, 48};
            uint32_t parity = 0;
            for (int t = 0; t < 4; ++t) {
                const uint32_t id = idesc_f16(64, ns[t]);
                const long long c0 = clock64();
                for (int rep = 0; rep < 256; ++rep) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        mma_ts(tmem + D_COL, tmem + A_COL + 8 * j, umma_desc(b0 + j * 2 * LBO_B, LBO_B, SBO_B), id, (rep | j) ? 1u : 0u);
                }
                const long long c1 = clock64();
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar)));
                mbar_wait(smem_u32(&s_bar), parity);
                parity ^= 1;
                const long long c2 = clock64();
                g.cycles[2 * t] = c1 - c0;          // issue time of 1024 MMAs
                g.cycles[2 * t + 1] = c2 - c0;      // until the last one has completed
            }
        } else {
            for (int j = 0; j < KK / 16; ++j) {
                const unsigned long long da = umma_desc(a0 + j * 2 * LBO_A, LBO_A, SBO_A);
                const unsigned long long db = umma_desc(b0 + j * 2 * LBO_B, LBO_B, SBO_B);
                const uint32_t acc = j > 0;
                const uint32_t up = 16u << 16;
                switch (g.test) {
                    case 0: mma_ss(tmem + D_COL, da, db, idesc_f16(64, 8), acc); break;
                    case 1: mma_ts(tmem + D_COL, tmem + A_COL + 8 * j, db, idesc_f16(64, 8), acc); break;
                    case 2: mma_ts(tmem + up + D_COL, tmem + up + A_COL + 8 * j, db, idesc_f16(64, 8), acc); break;
                    case 3: mma_ts(tmem + D_COL, tmem + A_COL + 8 * j, db, idesc_f16(128, 16), acc); break;
                    case 4: mma_ss(tmem + up + D_COL, da, db, idesc_f16(64, 8), acc); break;
                    case 5: mma_ts(tmem + up + D_COL, tmem + A_COL + 8 * j, db, idesc_f16(64, 8), acc); break;
                    case 6: mma_ss(tmem + D_COL, da, db, idesc_f16(64, 16), acc); break;
                    default: break;
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar)));
        }
    }
    if (g.test < 7) mbar_wait(smem_u32(&s_bar), 0);
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    uint32_t v[16];
    tmem_ld16(tq + D_COL, v);
#pragma unroll
    for (int i = 0; i < 16; ++i) g.out[(32 * warp + lane) * 16 + i] = __uint_as_float(v[i]);
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int main(int argc, char** argv) {
    const int test = argc > 1 ? atoi(argv[1]) : 0;
    std::vector<__half> A(128 * KK), B(16 * KK);
    std::vector<float> Af(128 * KK), Bf(16 * KK);
    unsigned s = 12345u;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return float(int((s >> 20) & 127) - 64) / 64.0f; };   // exact in fp16
    for (int i = 0; i < 128 * KK; ++i) { Af[i] = rnd(); A[i] = __float2half(Af[i]); }
    for (int i = 0; i < 16 * KK; ++i) { Bf[i] = rnd(); B[i] = __float2half(Bf[i]); }
    std::vector<float> ref(128 * 16);
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 16; ++n) {
            float acc = 0.f;
            for (int k = 0; k < KK; ++k) acc += Af[m * KK + k] * Bf[n * KK + k];
            ref[m * 16 + n] = acc;
        }
    Args g{};
    __half *dA, *dB;
    CK(cudaMalloc(&dA, A.size() * 2));
    CK(cudaMalloc(&dB, B.size() * 2));
    CK(cudaMalloc(&g.out, 128 * 16 * 4));
    CK(cudaMalloc(&g.cycles, 16 * 8));
    CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(g.out, 0, 128 * 16 * 4));
    g.A = dA; g.B = dB; g.test = test;
    probe<<<1, 128>>>(g);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    if (test == 7) {
        long long c[8];
        CK(cudaMemcpy(c, g.cycles, sizeof c, cudaMemcpyDeviceToHost));
        const int ns[4] = {8, 16, 24, 48};
        for (int t = 0; t < 4; ++t)
            printf("test 7: TS M=64 N=%2d K=16: issue %.2f cycles/MMA, complete %.2f cycles/MMA (1024 MMAs back to back)\n", ns[t],
                   c[2 * t] / 1024.0, c[2 * t + 1] / 1024.0);
        return 0;
    }
    if (test == 8) {
        long long c[16];
        CK(cudaMemcpy(c, g.cycles, sizeof c, cudaMemcpyDeviceToHost));
        const char* nm[5] = {"1 accumulator", "2 accumulators", "4 accumulators", "8 accumulators", "8 accumulators, lower/upper atoms alternating"};
        for (int t = 0; t < 5; ++t)
            printf("test 8: TS M=64 N=8 K=16, %s: issue %.2f, complete %.2f cycles/MMA (1024 MMAs, one thread)\n", nm[t], c[2 * t] / 1024.0, c[2 * t + 1] / 1024.0);
        printf("test 8: 4 threads x 256 MMAs, own accumulators: %.2f cycles/MMA overall\n", c[10] / 1024.0);
        return 0;
    }
    std::vector<float> out(128 * 16);
    CK(cudaMemcpy(out.data(), g.out, out.size() * 4, cudaMemcpyDeviceToHost));
    const int ncol = (test == 3 || test == 6) ? 16 : 8;
    printf("test %d: TMEM lane -> matching row of A.B^T (%d columns compared; '.' = all zero, '?' = non-zero but no row matches)\n", test, ncol);
    int matched = 0;
    for (int l = 0; l < 128; ++l) {
        bool zero = true;
        for (int n = 0; n < 16; ++n) zero = zero && out[l * 16 + n] == 0.f;
        int hit = -1;
        for (int m = 0; m < 128 && hit < 0; ++m) {
            bool ok = true;
            for (int n = 0; n < ncol; ++n) ok = ok && fabsf(out[l * 16 + n] - ref[m * 16 + n]) < 1e-3f;
            if (ok) hit = m;
        }
        if (hit >= 0) { printf(" %d:%d", l, hit); ++matched; }
        else if (!zero) printf(" %d:?", l);
        if (l % 32 == 31) printf("\n");
    }
    printf("test %d: %d lanes hold a row of the product\n", test, matched);
    // first non-matching lane, for diagnosis
    for (int l = 0; l < 128; ++l) {
        bool zero = true;
        for (int n = 0; n < 16; ++n) zero = zero && out[l * 16 + n] == 0.f;
        if (zero) continue;
        printf("lane %d:", l);
        for (int n = 0; n < 16; ++n) printf(" %.4f", out[l * 16 + n]);
        printf("\n   ref row 0:");
        for (int n = 0; n < 8; ++n) printf(" %.4f", ref[n]);
        printf("\n");
        break;
    }
    return 0;
}
