// Stand-alone GCC-PHAT lag projection on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a: seld_gcc_gemm.
//
// reference feature_extractor.py:209-211 computes, per microphone pair and frame,
//     cc = irfft(exp(1j * angle(R)))[lags -32..31]
// Only 64 of the 1024 irfft outputs are kept, so the pruned inverse transform is the dense contraction
//     cc[row, lag] = sum_K  A[row, K] * B[K, lag],      K = 1024
// with one row per (frame, pair), A = the unit phasor P[k] = R/|R| laid out as
//     A[row, 0] = Re P[0]   A[row, 1] = Re P[512]   A[row, 2k] = Re P[k]   A[row, 2k+1] = Im P[k]   (1 <= k <= 511)
// and B = the matching rows of the inverse DFT basis (seld_b200/tables.py: gcc_basis).  FP16 operands with FP32
// accumulation keep the result within ~5e-5 of the float32 irfft (tolerance 1e-3).
//
// The extractor no longer calls this kernel: it runs the same contraction per frame INSIDE extract_kernel, with the basis
// resident in tensor memory and the phasor rows never leaving shared memory (extract_core.cuh, "fused tensor-core GCC").
// This file keeps the dense [rows, 1024] x [1024, 64] form behind the C ABI for callers that hold phasor rows in HBM, and
// as the numerical pin of the fp16 basis (tests/test_gpu_gcc_gemm.py).
//
// Kernel: persistent CTAs of 192 threads, M = 128 rows per tile, K walked in 16 chunks of 64.  Warp 4 = producer: one
// cp.async.bulk per operand per chunk (both operands arrive as ready-made K-major SWIZZLE_128B images) into a 3-stage
// mbarrier ring; warp 5 = one ELECTED thread issuing 4 x tcgen05.mma (M128 N64 K16, kind::f16) per chunk and
// tcgen05.commit to free the stage; warps 0-3 = epilogue (tcgen05.ld, warp w <-> TMEM lanes 32w..32w+31); the accumulator
// is double-buffered in TMEM so tile i+1's MMAs overlap tile i's epilogue.
#include <cuda_fp16.h>

#include "plan.h"

namespace seld {

constexpr int GM = 128;          // rows per tile
constexpr int GN = 64;           // lags
constexpr int GK = 1024;         // contraction length
constexpr int GKC = 64;          // K elements per chunk (128 bytes per row)
constexpr int NCHUNK = GK / GKC;
constexpr int GSTAGES = 3;
constexpr int GA_BYTES = GM * GKC * 2;       // 16 KB
constexpr int GB_BYTES = GN * GKC * 2;       // 8 KB
constexpr int GSTAGE_BYTES = GA_BYTES + GB_BYTES;
// Operand images.  Both operands arrive in global memory ALREADY in the shared-memory image of the canonical K-major
// SWIZZLE_128B UMMA layout, one contiguous block per (tile, chunk): row r of a chunk is 128 contiguous bytes at r * 128,
// its 16-byte unit u stored at unit position u ^ (r % 8) (Swizzle<3,4,3>); 8-row atoms are 1024 bytes apart (SBO).  One
// cp.async.bulk per operand per chunk lands it -- no per-thread copies, no proxy fence -- and the extractor can write
// every row segment with one coalesced 128-byte store.
constexpr int GSBO = 1024;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address, LBO field 1 (unused
// for a swizzled K-major operand whose K extent fits one 128-byte atom), SBO = 1024 bytes between 8-row atoms,
// version 1, layout type 2 = SWIZZLE_128B.  A K = 16 step advances the start address by 32 bytes inside the atom.
__device__ __forceinline__ unsigned long long umma_desc_sw128(uint32_t saddr) {
    unsigned long long d = 0;
    d |= (unsigned long long)((saddr >> 4) & 0x3FFF);
    d |= 1ull << 16;
    d |= (unsigned long long)((GSBO >> 4) & 0x3FFF) << 32;
    d |= 1ull << 46;
    d |= 2ull << 61;
    return d;
}

// kind::f16 instruction descriptor: F32 accumulate, F16 x F16, both K-major, N = 64, M = 128 (cute::UMMA::InstrDescriptor)
constexpr uint32_t kIdesc = (1u << 4) | ((GN >> 3) << 17) | ((GM >> 4) << 24);

struct GccGemmArgs {
    const __half* A;          // operand image [tile][16 chunks][16 KB]
    const __half* Bt;         // operand image [16 chunks][8 KB]
    long long n_tiles;
    float scale;              // epilogue factor (the basis is stored x512)
    float* dense_out;         // [n_tiles * 128][64]
    long long dense_rows;     // valid rows
};

__device__ __forceinline__ bool elect_one_lane() {      // one lane of a converged warp: lets ptxas issue UTCHMMA without a per-thread loop
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// Warp roles (192 threads): warps 0-3 epilogue (TMEM lane quarter w), warp 4 bulk-copy producer, warp 5 MMA issuer.
__global__ void __launch_bounds__(192) gcc_gemm_kernel(GccGemmArgs g) {
    extern __shared__ __align__(1024) unsigned char gsm[];
    __shared__ __align__(8) unsigned long long s_full[GSTAGES], s_empty[GSTAGES], s_acc_full[2], s_acc_empty[2];
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t base = (smem_u32(gsm) + 1023u) & ~1023u;      // SWIZZLE_128B atoms must be 1024-byte aligned

    if (tid == 0) {
        for (int s = 0; s < GSTAGES; ++s) { mbar_init(smem_u32(&s_full[s]), 1); mbar_init(smem_u32(&s_empty[s]), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&s_acc_full[s]), 1); mbar_init(smem_u32(&s_acc_empty[s]), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = s_tmem;

    if (warp == 4) {
        // ---------------- producer: two bulk copies per chunk, GSTAGES deep, running ahead across tiles
        if (lane == 0) {
            uint32_t q = 0;
            for (long long tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
                const unsigned char* a_src = reinterpret_cast<const unsigned char*>(g.A) + tile * (long long)NCHUNK * GA_BYTES;
                for (int c = 0; c < NCHUNK; ++c, ++q) {
                    const uint32_t s = q % GSTAGES;
                    if (q >= GSTAGES) mbar_wait(smem_u32(&s_empty[s]), ((q / GSTAGES) - 1) & 1);
                    const uint32_t bar = smem_u32(&s_full[s]), dst = base + s * GSTAGE_BYTES;
                    mbar_expect_tx(bar, GSTAGE_BYTES);
                    bulk_g2s(dst, a_src + (long long)c * GA_BYTES, GA_BYTES, bar);
                    bulk_g2s(dst + GA_BYTES, reinterpret_cast<const unsigned char*>(g.Bt) + (long long)c * GB_BYTES, GB_BYTES, bar);
                }
            }
        }
    } else if (warp == 5) {
        // ---------------- MMA issuer: 4 x (M128 N64 K16) per chunk into TMEM accumulator (tile parity)
        // (an `if (lane == 0)` here makes ptxas wrap every UTCHMMA in a per-thread ELECT loop: ~60 cycles per MMA, measured
        //  with tools/microbench/probe_tmem_ts.cu; elect.sync issues them back to back)
        if (elect_one_lane()) {
            uint32_t q = 0, lt = 0;
            for (long long tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ++lt) {
                const uint32_t as = lt & 1;
                if (lt >= 2) mbar_wait(smem_u32(&s_acc_empty[as]), ((lt >> 1) - 1) & 1);     // epilogue drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;");
                const uint32_t d_tmem = tmem + as * GN;
                for (int c = 0; c < NCHUNK; ++c, ++q) {
                    const uint32_t s = q % GSTAGES;
                    mbar_wait(smem_u32(&s_full[s]), (q / GSTAGES) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    const uint32_t sa = base + s * GSTAGE_BYTES, sbb = sa + GA_BYTES;
#pragma unroll
                    for (int j = 0; j < GKC / 16; ++j) {
                        const unsigned long long da = umma_desc_sw128(sa + j * 32);
                        const unsigned long long db = umma_desc_sw128(sbb + j * 32);
                        const uint32_t accum = (c > 0 || j > 0) ? 1u : 0u;
                        asm volatile(
                            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                            ::"r"(d_tmem), "l"(da), "l"(db), "r"(kIdesc), "r"(accum));
                    }
                    // frees the stage for the producer once these MMAs have read it
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_empty[s])));
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_acc_full[as])));
            }
        }
    } else {
        // ---------------- epilogue warps 0-3: thread (warp w, lane l) owns row 32w + l = TMEM lane 32w + l
        uint32_t lt = 0;
        for (long long tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ++lt) {
            const uint32_t as = lt & 1;
            mbar_wait(smem_u32(&s_acc_full[as]), (lt >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;");
            uint32_t v[64];
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + as * GN;
#define SELD_TMEM_LD32(OFF)                                                                                             \
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                       \
                         "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                      \
                         "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"       \
                         : "=r"(v[OFF + 0]), "=r"(v[OFF + 1]), "=r"(v[OFF + 2]), "=r"(v[OFF + 3]), "=r"(v[OFF + 4]),      \
                           "=r"(v[OFF + 5]), "=r"(v[OFF + 6]), "=r"(v[OFF + 7]), "=r"(v[OFF + 8]), "=r"(v[OFF + 9]),      \
                           "=r"(v[OFF + 10]), "=r"(v[OFF + 11]), "=r"(v[OFF + 12]), "=r"(v[OFF + 13]), "=r"(v[OFF + 14]), \
                           "=r"(v[OFF + 15]), "=r"(v[OFF + 16]), "=r"(v[OFF + 17]), "=r"(v[OFF + 18]), "=r"(v[OFF + 19]), \
                           "=r"(v[OFF + 20]), "=r"(v[OFF + 21]), "=r"(v[OFF + 22]), "=r"(v[OFF + 23]), "=r"(v[OFF + 24]), \
                           "=r"(v[OFF + 25]), "=r"(v[OFF + 26]), "=r"(v[OFF + 27]), "=r"(v[OFF + 28]), "=r"(v[OFF + 29]), \
                           "=r"(v[OFF + 30]), "=r"(v[OFF + 31])                                                           \
                         : "r"(taddr + OFF))
            SELD_TMEM_LD32(0);
            SELD_TMEM_LD32(32);
#undef SELD_TMEM_LD32
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&s_acc_empty[as]));      // the MMA warp may reuse this accumulator

            const long long row = tile * GM + tid;
            if (row < g.dense_rows) {
                float4* d4 = reinterpret_cast<float4*>(g.dense_out + row * GN);
#pragma unroll
                for (int n = 0; n < GN; n += 4)
                    d4[n / 4] = make_float4(__uint_as_float(v[n]) * g.scale, __uint_as_float(v[n + 1]) * g.scale,
                                            __uint_as_float(v[n + 2]) * g.scale, __uint_as_float(v[n + 3]) * g.scale);
            }
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
}

int launch_gcc_gemm(const GccGemmArgs& g, int num_sms, cudaStream_t st) {
    if (g.n_tiles <= 0) return SELD_OK;
    const int smem = GSTAGES * GSTAGE_BYTES + 1024;      // 73 KB -> up to 3 CTAs per SM (128 TMEM columns each)
    static unsigned long long configured = 0;             // bit d: attribute set on device d
    if (first_use_on_device(&configured))
        SELD_CUDA_TRY(cudaFuncSetAttribute(gcc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    long long grid = (long long)num_sms * 2;
    if (grid > g.n_tiles) grid = g.n_tiles;
    gcc_gemm_kernel<<<(int)grid, 192, smem, st>>>(g);
    SELD_CUDA_TRY(cudaGetLastError());
    seld::note_launch();
    return SELD_OK;
}

}  // namespace seld

using namespace seld;

// Stand-alone entry point (what tests/test_gpu_gcc_gemm.py drives): out[rows][64] = scale * A * B with both operands
// given as UMMA operand images (see seld_b200.tables.gcc_operand_image): a_img [ceil(rows/128)][16][16 KB],
// bt_img [16][8 KB].
extern "C" int seld_gcc_gemm(const void* a_img_dev, const void* bt_img_dev, int64_t rows, float scale, float* out_dev,
                             void* stream) {
    if (!a_img_dev || !bt_img_dev || !out_dev || rows < 0) { set_error("bad argument"); return SELD_EINVAL; }
    if ((reinterpret_cast<uintptr_t>(a_img_dev) | reinterpret_cast<uintptr_t>(bt_img_dev) | reinterpret_cast<uintptr_t>(out_dev)) % 16) {
        set_error("operands must be 16-byte aligned");
        return SELD_EINVAL;
    }
    static unsigned long long checked = 0;      // one device check per device (the query is slow and takes driver locks)
    if (first_use_on_device(&checked)) {
        const int rc = seld_device_check(-1);
        if (rc != SELD_OK) return rc;
    }
    const int sms = device_sm_count();
    GccGemmArgs g{};
    g.A = static_cast<const __half*>(a_img_dev);
    g.Bt = static_cast<const __half*>(bt_img_dev);
    g.n_tiles = (rows + GM - 1) / GM;
    g.scale = scale;
    g.dense_out = out_dev;
    g.dense_rows = rows;
    return launch_gcc_gemm(g, sms, static_cast<cudaStream_t>(stream));
}
