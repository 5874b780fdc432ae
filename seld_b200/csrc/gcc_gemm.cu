// GCC-PHAT lag projection on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
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
// Kernel: one CTA = 128 threads owns M = 128 rows at a time.  K is walked in 16 chunks of 64; each chunk of A (128 x 64
// halfs = 16 KB, read from the phasor scratch rows the extractor wrote) and of B^T (64 x 64 halfs = 8 KB, L2 resident)
// is brought into shared memory with coalesced 16-byte cp.async in a (padded) no-swizzle K-major core-matrix layout, three
// stages deep; one elected thread issues 4 x tcgen05.mma (M128 N64 K16, kind::f16) per chunk accumulating in 64 TMEM
// columns, and tcgen05.commit on the stage's mbarrier tells the loaders when the stage may be overwritten.  The epilogue
// reads the accumulator with tcgen05.ld (warp w owns TMEM lanes 32w..32w+31 = its 32 rows), scales, and scatters the 64
// lags of each row into the GCC channel of the feature tensor ([t][mel = lag index][4 + pair]).
#include <cuda_fp16.h>

#include <atomic>

#include "plan.h"

namespace seld {

constexpr int GM = 128;          // rows per tile
constexpr int GN = 64;           // lags
constexpr int GK = 1024;         // contraction length
constexpr int GKC = 64;          // K elements per chunk (128 bytes per row)
constexpr int GSTAGES = 4;          // chunks in flight: GSTAGES - 1 ahead of the MMA
// Shared-memory operand layout (K-major, no swizzle): element (row r, k) of a stage lives at
//     (r / 8) * GSBO + (k / 8) * GLBO + (r % 8) * 16 + (k % 8) * 2
// i.e. 8-row x 16-byte core matrices, K-adjacent core matrices GLBO = 144 bytes apart (128 + 16 bytes of padding) and
// 8-row groups GSBO = 8 * 144 bytes apart.  The padding makes the coalesced global->shared copy (8 consecutive lanes =
// the 8 k-groups of one row) hit 8 different 16-byte bank groups instead of one.
constexpr int GLBO = 144;
constexpr int GSBO = 8 * GLBO;               // 1152
constexpr int GA_BYTES = (GM / 8) * GSBO;    // 18 KB
constexpr int GB_BYTES = (GN / 8) * GSBO;    // 9 KB
constexpr int GSTAGE_BYTES = GA_BYTES + GB_BYTES;

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
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src));
}

// K-major, no-swizzle ("interleave") shared-memory matrix descriptor: 8-row x 16-byte core matrices; LBO = byte distance
// between the two core matrices of one K = 16 step, SBO = byte distance between 8-row groups (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ unsigned long long umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    unsigned long long d = 0;
    d |= (unsigned long long)((saddr >> 4) & 0x3FFF);
    d |= (unsigned long long)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (unsigned long long)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= 1ull << 46;                        // descriptor version 1 (Blackwell); layout_type 0 = SWIZZLE_NONE
    return d;
}

// kind::f16 instruction descriptor: F32 accumulate, F16 x F16, both K-major, N = 64, M = 128 (cute::UMMA::InstrDescriptor)
constexpr uint32_t kIdesc = (1u << 4) | ((GN >> 3) << 17) | ((GM >> 4) << 24);

struct GccGemmArgs {          // (also declared in extract.cu, which drives the kernel in scatter mode)
    const __half* A;          // [rows][1024]
    const __half* Bt;         // [64][1024]   (lag-major: K contiguous)
    long long rows;
    float scale;              // epilogue factor (the basis is stored x512)
    float* dense_out;         // debug: [rows][64]; nullptr in production
    float* feat;              // production: feature tensor [clip][t_out][64][n_ch]
    int frames_per_clip;      // rows are ((clip * frames_per_clip + t) * 6 + pair)
    int t_out, n_ch;
    int blocked;              // A layout: 0 = row-major [rows][1024]; 1 = tile-blocked [tile][chunk][128 rows][64] (fused path)
};

__global__ void __launch_bounds__(128) gcc_gemm_kernel(GccGemmArgs g) {
    extern __shared__ __align__(1024) unsigned char gsm[];
    __shared__ __align__(8) unsigned long long s_bar[GSTAGES];
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t base = (smem_u32(gsm) + 127u) & ~127u;        // stage blocks start on a core-matrix boundary

    if (tid == 0) {
        for (int s = 0; s < GSTAGES; ++s) mbar_init(smem_u32(&s_bar[s]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(64));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = s_tmem;

    uint32_t commits[GSTAGES] = {};        // tcgen05.commit count per stage barrier (uniform across threads)
    const long long n_tiles = (g.rows + GM - 1) / GM;

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        // Coalesced copies: 8 consecutive threads fetch the 8 16-byte k-groups (one 128-byte line) of one row; a pass of
        // the 128 threads covers 16 rows, so A takes 8 passes and B^T 4.  Rows past the end are clamped (their results
        // are computed and discarded).
        const int gk = tid & 7, r0 = tid >> 3;
        const long long tile_row0 = tile * GM;

        auto load_chunk = [&](int c) {
            const int s = c % GSTAGES;
            const uint32_t sb = base + s * GSTAGE_BYTES;
#pragma unroll
            for (int i = 0; i < GM / 16; ++i) {
                const int r = r0 + 16 * i;
                const __half* src;
                if (g.blocked) {
                    // the extractor wrote this (tile, chunk) as one contiguous 16 KB block: full DRAM pages, and rows past
                    // the end exist as (uninitialised, discarded) padding
                    src = g.A + ((tile * (GK / GKC) + c) * GM + r) * GKC + gk * 8;
                } else {
                    long long grow = tile_row0 + r;
                    if (grow >= g.rows) grow = g.rows - 1;
                    src = g.A + grow * GK + c * GKC + gk * 8;
                }
                cp_async16(sb + (r >> 3) * GSBO + gk * GLBO + (r & 7) * 16, src);
            }
#pragma unroll
            for (int i = 0; i < GN / 16; ++i) {
                const int r = r0 + 16 * i;
                cp_async16(sb + GA_BYTES + (r >> 3) * GSBO + gk * GLBO + (r & 7) * 16, g.Bt + (long long)r * GK + c * GKC + gk * 8);
            }
        };

        constexpr int NCHUNK = GK / GKC;
        constexpr int AHEAD = GSTAGES - 1;
        // prologue: the previous tile's MMAs have all retired (its last commit was awaited), so every stage is free
#pragma unroll
        for (int c = 0; c < AHEAD; ++c) {
            load_chunk(c);
            asm volatile("cp.async.commit_group;");
        }
        for (int c = 0; c < NCHUNK; ++c) {
            if (c + AHEAD < NCHUNK) {
                const int s2 = (c + AHEAD) % GSTAGES;
                // the stage is free once the MMAs of its previous user (chunk c-1, or the previous tile) have completed
                if (commits[s2] > 0) mbar_wait(smem_u32(&s_bar[s2]), (commits[s2] - 1) & 1);
                load_chunk(c + AHEAD);
            }
            asm volatile("cp.async.commit_group;");
            asm volatile("cp.async.wait_group %0;" ::"n"(AHEAD) : "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the MMA
            __syncthreads();
            const int s = c % GSTAGES;
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;");
                const uint32_t sa = base + s * GSTAGE_BYTES, sbb = sa + GA_BYTES;
#pragma unroll
                for (int j = 0; j < GKC / 16; ++j) {
                    const unsigned long long da = umma_desc(sa + j * 2 * GLBO, GLBO, GSBO);
                    const unsigned long long db = umma_desc(sbb + j * 2 * GLBO, GLBO, GSBO);
                    const uint32_t accum = (c > 0 || j > 0) ? 1u : 0u;
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                        ::"r"(tmem), "l"(da), "l"(db), "r"(kIdesc), "r"(accum));
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar[s])));
            }
            commits[s] += 1;
        }
        // all MMAs of the tile are complete when the last chunk's commit has arrived (MMAs retire in order)
        {
            const int s = (NCHUNK - 1) % GSTAGES;
            mbar_wait(smem_u32(&s_bar[s]), (commits[s] - 1) & 1);
        }
        asm volatile("tcgen05.fence::after_thread_sync;");

        // ---- epilogue: thread (warp w, lane l) owns row 32w + l = TMEM lane 32w + l, 64 columns
        uint32_t v[64];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
#define SELD_TMEM_LD32(OFF)                                                                                             \
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                           \
                     "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                          \
                     "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"           \
                     : "=r"(v[OFF + 0]), "=r"(v[OFF + 1]), "=r"(v[OFF + 2]), "=r"(v[OFF + 3]), "=r"(v[OFF + 4]),          \
                       "=r"(v[OFF + 5]), "=r"(v[OFF + 6]), "=r"(v[OFF + 7]), "=r"(v[OFF + 8]), "=r"(v[OFF + 9]),          \
                       "=r"(v[OFF + 10]), "=r"(v[OFF + 11]), "=r"(v[OFF + 12]), "=r"(v[OFF + 13]), "=r"(v[OFF + 14]),     \
                       "=r"(v[OFF + 15]), "=r"(v[OFF + 16]), "=r"(v[OFF + 17]), "=r"(v[OFF + 18]), "=r"(v[OFF + 19]),     \
                       "=r"(v[OFF + 20]), "=r"(v[OFF + 21]), "=r"(v[OFF + 22]), "=r"(v[OFF + 23]), "=r"(v[OFF + 24]),     \
                       "=r"(v[OFF + 25]), "=r"(v[OFF + 26]), "=r"(v[OFF + 27]), "=r"(v[OFF + 28]), "=r"(v[OFF + 29]),     \
                       "=r"(v[OFF + 30]), "=r"(v[OFF + 31])                                                               \
                     : "r"(taddr + OFF))
        SELD_TMEM_LD32(0);
        SELD_TMEM_LD32(32);
#undef SELD_TMEM_LD32
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");

        const long long row = tile * GM + tid;
        if (row < g.rows) {
            if (g.dense_out != nullptr) {
                float4* d4 = reinterpret_cast<float4*>(g.dense_out + row * GN);
#pragma unroll
                for (int n = 0; n < GN; n += 4)
                    d4[n / 4] = make_float4(__uint_as_float(v[n]) * g.scale, __uint_as_float(v[n + 1]) * g.scale,
                                            __uint_as_float(v[n + 2]) * g.scale, __uint_as_float(v[n + 3]) * g.scale);
            } else {
                const long long frame = row / 6;
                const int pair = int(row - frame * 6);
                const long long clip = frame / g.frames_per_clip;
                const int t = int(frame - clip * g.frames_per_clip);
                float* dst = g.feat + ((clip * g.t_out + t) * GN) * g.n_ch + 4 + pair;
#pragma unroll
                for (int n = 0; n < GN; ++n) dst[n * g.n_ch] = __uint_as_float(v[n]) * g.scale;
            }
        }
        // the accumulator may be overwritten by the next tile only after every warp has read it
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;");
    }

    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64));
}

int launch_gcc_gemm(const GccGemmArgs& g, int num_sms, cudaStream_t st) {
    if (g.rows <= 0) return SELD_OK;
    const int smem = GSTAGES * GSTAGE_BYTES + 128;  // 108 KB -> 2 CTAs per SM (64 TMEM columns each)
    static std::atomic<int> configured{0};
    if (!configured.load(std::memory_order_acquire)) {
        SELD_CUDA_TRY(cudaFuncSetAttribute(gcc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured.store(1, std::memory_order_release);
    }
    const long long n_tiles = (g.rows + GM - 1) / GM;
    long long grid = (long long)num_sms * 2;
    if (grid > n_tiles) grid = n_tiles;
    gcc_gemm_kernel<<<(int)grid, 128, smem, st>>>(g);
    SELD_CUDA_TRY(cudaGetLastError());
    return SELD_OK;
}

}  // namespace seld

using namespace seld;

// Stand-alone entry point (also what tests/test_gpu_gcc_gemm.py drives): out[rows][64] = scale * A[rows][1024] * Bt^T.
extern "C" int seld_gcc_gemm(const void* a_dev, const void* bt_dev, int64_t rows, float scale, float* out_dev, void* stream) {
    if (!a_dev || !bt_dev || !out_dev || rows < 0) { set_error("bad argument"); return SELD_EINVAL; }
    if ((reinterpret_cast<uintptr_t>(a_dev) | reinterpret_cast<uintptr_t>(bt_dev) | reinterpret_cast<uintptr_t>(out_dev)) % 16) {
        set_error("operands must be 16-byte aligned");
        return SELD_EINVAL;
    }
    static int checked = 0, sms = 148;          // one device query per process (it is slow and takes driver locks)
    if (!checked) {
        const int rc = seld_device_check(-1);
        if (rc != SELD_OK) return rc;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        checked = 1;
    }
    GccGemmArgs g{};
    g.A = static_cast<const __half*>(a_dev);
    g.Bt = static_cast<const __half*>(bt_dev);
    g.rows = rows;
    g.scale = scale;
    g.dense_out = out_dev;
    g.feat = nullptr;
    g.frames_per_clip = 1;
    g.t_out = 1;
    g.n_ch = 10;
    g.blocked = 0;
    return launch_gcc_gemm(g, sms, static_cast<cudaStream_t>(stream));
}
