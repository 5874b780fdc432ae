// Fused STFT -> log-mel + intensity-vector / GCC-PHAT extractor, sm_100a.
//
// Replaces reference feature_extractor.py:53-88 (extract_features) and the feature half of :117-149
// (pad / truncate), batched over clips.  A team of two warps owns one STFT frame of all four channels; see
// extract_core.cuh for the per-lane algorithm.  Persistent CTAs (one per SM, limited by shared memory)
// walk super-chunks of consecutive frames so that the 2.13x overlap between neighbouring frames is served
// by L1/L2 and HBM sees every sample once.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <string>
#include <vector>

#include <cuda_fp16.h>

#include "extract_core.cuh"
#include "mel_pieces.h"
#include "plan.h"

namespace seld {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
int cuda_fail(cudaError_t e, const char* what) {
    set_error(std::string(what) + ": " + cudaGetErrorString(e));
    return SELD_ECUDA;
}

static std::atomic<long long> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int device_sm_count() {
    static std::atomic<int> cache[64];
    int dev = 0;
    cudaGetDevice(&dev);
    std::atomic<int>& c = cache[dev & 63];
    int v = c.load(std::memory_order_acquire);
    if (v == 0) {
        v = 148;
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        c.store(v, std::memory_order_release);
    }
    return v;
}

bool first_use_on_device(unsigned long long* slot_bits) {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    return !(__atomic_fetch_or(slot_bits, bit, __ATOMIC_ACQ_REL) & bit);
}

struct ExtractArgs {
    const float* wav;
    int layout;
    int n_clips;
    long long n_samples;
    int t_raw, t_out, t_tot;
    int t_lo, t_hi;           // frames [t_lo, t_hi) lie wholly inside the clip (no reflection); the rest are edge frames
    float* out;
    unsigned int* clip_max_key;
    const float* window;
    const float2* tw_t;
    const float2* tw_lin;
    const float2* w01;
    const unsigned long long* endmask;
    const int* slot0;
    const int* slot1;
    const int* ov;
    const int* pb;
    int hop, n_mels;
    int x_bytes;              // per-team piece / GCC exchange buffer size (SmemPlan::x_bytes)
    int seg_major;            // piece records in the segment-major layout: gather_lanes (else the compact layout + gather_phase)
    int x_zero_f2;            // float2 elements of the piece buffer that must read as zero where no piece is stored
    int gcc_tc;               // MIC, n_fft 1024, 64 lags: fused tensor-core lag projection (extract_core.cuh)
    const void* gcc_basis;    // fp16 [64 lags][1024] basis of that projection (x512), row-major
    int tf_variant;           // FOA plan created with SELD_MODE_FOA_TF (magnitude mel, 20 log10, zero-padded tail)
    int lanes;                // FOA, n_fft 1024: the bank has a flush-free lane form (mel_pieces.h)
    const float* w4;          // [64 * 9][4] lane-form weights
    const int* lane_beg;      // [64]
    const int* gtab;          // [64][kLaneGatherMax]
    int gather_n0, gather_n1;
    int frames_per_clip;      // frames this launch handles per clip (interior or edge count)
    int origin;               // frame t starts at sample t*hop - n_fft/2 + origin (0: centred STFT; n_fft/2: uncentred chunks)
    int fpw;                  // consecutive frames per team per super-chunk
    int fsc;                  // frames per super-chunk (teams * fpw)
    long long n_super;        // super-chunks of fsc frames
};

static int frames_per_team(int dflt) {   // consecutive frames a team handles per super-chunk (SELD_FPW overrides, for experiments)
    const char* e = getenv("SELD_FPW");
    const int n = e ? atoi(e) : 0;
    // measured (600 planar FOA clips, 16 warps, shared taps kept in tensor memory): 1 -> 9.72 ms, 2 -> 9.27, 4 -> 9.13, 8 -> 9.10;
    // interleaved 9.82 / 9.52 / 9.22 / 9.15.  Kernels without the tap reuse are flat from 2 upwards.
    return n > 0 ? n : dflt;
}

__host__ __device__ constexpr int align16(int x) { return (x + 15) & ~15; }
#ifndef SELD_EARLY_TAIL
#define SELD_EARLY_TAIL 0
#endif
// FOA experiment (kept off): request the next frame's 15 new taps right after the bin phase, so that their latency hides
// behind the gather and the row store instead of being waited for at the top of the frame (16 % of the stall samples).
// Measured, 600 clips: 10.37 ms with it, 9.25 ms without -- 30 more live registers at the 128-register budget of four
// warps per scheduler cost more (112 B of spills, a tighter gather) than the exposed load.
constexpr bool kEarlyTail = SELD_EARLY_TAIL;
#ifndef SELD_BULK_TAIL
#define SELD_BULK_TAIL 1
#endif
// FOA, planar input: the 15 new taps of the team's NEXT frame (two channels x 1 920 contiguous bytes per warp) are staged through
// shared memory by two cp.async.bulk copies issued right after the bin phase -- into the warp's own exchange buffer, which is dead
// from there until the next frame's stage-1 store -- and read back with 30 conflict-free LDS at the top of the frame: no
// registers are held across the gather (the early register request above lost to exactly that) and the HBM / L2 latency is off
// the critical path.  north_star (1)'s "staged through TMA / shared memory", measured: see DESIGN.md 4.1.
constexpr bool kBulkTail = SELD_BULK_TAIL;
#ifndef SELD_BULK_MIC
#define SELD_BULK_MIC 1
#endif
// Fused MIC kernel, planar input: it has neither registers for a prefetch nor tensor-memory columns for the shared taps, so the
// frame's global loads sat exposed at the top of the frame (12 % of the stall samples).  The WHOLE next frame of the warp's channel
// pair (2 x 4 096 contiguous bytes) is staged by two cp.async.bulk copies into the warp's exchange buffer -- a part of the operand
// tile, free as soon as the team pair's MMA chain has completed -- issued right after that wait, so the copies run under the GCC
// epilogue, the pair barrier and the row store, and the frame starts with 64 conflict-free LDS instead of 64 LDG.
constexpr bool kBulkMic = SELD_BULK_MIC;
#ifndef SELD_GCC_DEAD
#define SELD_GCC_DEAD 1
#endif
constexpr bool kGccDead = SELD_GCC_DEAD;               // (experiments: cost of the dead-channel detection)
constexpr int kGccTileBytes = GT_BYTES;          // fused GCC: phasor tile (the exchange buffers alias it) ...
constexpr int kGccNyqBytes = GT_NYQ_BYTES + 16;  // ... + the Nyquist column + the dead-channel flags of the two warps

// Shared-memory plan of one kernel variant.  CTA-shared tables first, then one region per frame team (two warps): an
// exchange buffer per warp (its spectrum overwrites it in place), the piece / GCC exchange buffer, the staged output row.
// n_fft = 1024 keeps the per-lane constant tables in tensor memory, so only the small index tables stay here.
// TC: variant bits -- bit 0: MIC: fused tensor-core GCC / FOA: the TensorFlow variant; bit 1 (FOA, n_fft 1024): flush-free lane form of the mel bank
template <int R, int MODE, int TC>
struct SmemPlan {
    using G = Geo<R>;
    static constexpr bool TM = (R == 32);                               // per-lane constants in tensor memory
    static constexpr bool tc = (TC & 1) && MODE == MODE_MIC && R == 32;       // fused tensor-core GCC: the spectrum buffer doubles as the MMA's B tile
    static constexpr bool lanes = (TC & 2) && MODE == MODE_FOA && R == 32;    // flush-free lane form: two more index tables
    static constexpr bool NEED_TW = !TM || (MODE == MODE_MIC && !tc);   // stage-1 twiddles (also the fast CUDA-core GCC stage 2)
    static constexpr bool NEED_W01 = !TM;
    static constexpr bool NEED_WIN = !TM;
    static constexpr bool NEED_TWLIN = (MODE == MODE_MIC && !tc);
    __host__ __device__ static constexpr int table_bytes(int n_mels) {
        return (NEED_TW ? align16(G::N * 8) : 0) + (NEED_W01 ? align16(G::TL * G::BPT * 8) : 0) + align16(G::TL * 8) + 2 * align16(G::TL * 4) +
               align16((n_mels + 2) * 4) + align16(64 * 4) + 16 + (NEED_WIN ? align16(G::N * 4) : 0) + (NEED_TWLIN ? align16(G::N * 8) : 0) +
               (tc ? 64 : 0) +                                          // one mbarrier per frame team
               (((kBulkTail && MODE == MODE_FOA && R == 32) || (kBulkMic && tc)) ? 128 : 0) +  // one mbarrier per warp (bulk-staged taps)
               (lanes ? align16(G::TL * 4) + align16(64 * kLaneGatherMax * 4) : 0);
    }
    __host__ __device__ static constexpr int x_bytes(int n_slots) {
        const int pieces = align16(n_slots * PieceGeo<MODE>::PSTRIDE * 8);
        const int exchange = (MODE == MODE_MIC && !tc) ? align16(G::E_ELEMS * 8) : 0;
        if (lanes) return align16((kLaneZeroRec + 1) * kLaneRecWords * 4);      // 64 lane records + the zero record
        return pieces > exchange ? pieces : exchange;
    }
    __host__ __device__ static constexpr int spec_bytes() {            // the two exchange / spectrum buffers of a team
        return tc ? kGccTileBytes + kGccNyqBytes : 2 * align16(G::E_ELEMS * 8);
    }
    __host__ __device__ static constexpr int team_bytes(int n_mels, int xb) {
        return spec_bytes() + xb + align16(n_mels * (MODE == MODE_FOA ? 7 : 10) * 4);
    }
};

// Warps per CTA.  n_fft = 1024: up to 16 warps = 8 teams at 128 registers per thread (four warps per scheduler; measured
// 2 -> 4 -> 8 warps with whole frames per warp: 35 -> 17.7 -> 11.5 ms, then teams of two: 12 warps 9.9 ms).
// (Round 1's MIC kernel kept 12 warps at 168 registers: 18.0 ms at 16 against 17.0.  The fused-GCC kernel of round 2 is latency /
//  issue bound with the phasor copy-out gone, fits 128 registers with 76 B of spills once the whole-frame register prefetch is
//  dropped, and gains from the fourth warp per scheduler: 15.2 ms at 16 warps against 16.8 at 12, SELD_MIC_WARPS.)
template <int R, int MODE>
#ifndef SELD_MIC_WARPS
#define SELD_MIC_WARPS 16
#endif
__host__ __device__ constexpr int max_warps() { return R <= 16 ? 16 : (R == 32 ? (MODE == MODE_FOA ? 16 : SELD_MIC_WARPS) : 4); }

__device__ __forceinline__ void team_bar(int id) { asm volatile("bar.sync %0, 64;" :: "r"(id) : "memory"); }
__device__ __forceinline__ void pair_bar(int id) { asm volatile("bar.sync %0, 128;" :: "r"(id) : "memory"); }
__device__ __forceinline__ unsigned smem_addr(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }

// EDGE = false: interior frames only, loads specialised on LAYOUT (the hot kernel).
// EDGE = true : the few frames per clip that need reflection, plus the zero padding rows (generic strided loads).
// TC (MIC, n_fft 1024, 64 lags): the GCC lag projection runs on the tensor cores inside this kernel (extract_core.cuh,
// "fused tensor-core GCC").  Two neighbouring teams form a PAIR there: their four warps cover the four tensor-memory
// subpartitions, which is what reading an M = 64 accumulator back takes, so the pair reads both teams' accumulators.
template <int R, int MODE, int LAYOUT, bool EDGE, int TC>
__global__ void __launch_bounds__(max_warps<R, MODE>() * 32, 1) extract_kernel(ExtractArgs a) {
    using G = Geo<R>;
    constexpr int C = (MODE == MODE_FOA) ? 7 : 10;
    constexpr int TL = G::TL;
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int h = warp & 1;                   // which channel pair this warp transforms
    const int team = warp >> 1;
    const int u = h * 32 + lane;              // lane within the team
    const int bar_id = 1 + team;              // named barrier of the team (0 is __syncthreads)
    // per-lane constant tables (window, twiddles, mel weights) live in tensor memory for n_fft = 1024: tcgen05.ld keeps
    // ~18 % of the shared-memory wavefronts off the LSU data pipe this kernel is bound by (extract_core.cuh)
    using SP = SmemPlan<R, MODE, TC>;
    constexpr bool TM = SP::TM;
    constexpr bool FUSED = SP::tc;
    // TC on a FOA plan selects the TensorFlow variant of the features (reference data_loader.py:310-349,
    // get_preprocessed_x_tf): mel bank on |X|, 20 log10 without a floor, frames from sample 0 with a zero-padded tail
    constexpr bool TFV = (TC & 1) && MODE == MODE_FOA;
    constexpr bool LANES = SP::lanes;
    constexpr bool PACKED_TW = FUSED;      // fused MIC: packed stage-1 twiddle products + 64-bit exchange stores (extract_core.cuh: cmul_rt_packed)
    // ---- CTA-shared tables
    unsigned char* p = smem;
    float2* s_tw_t = nullptr;
    if constexpr (SP::NEED_TW) { s_tw_t = reinterpret_cast<float2*>(p);  p += align16(G::N * 8); }
    float2* s_w01 = nullptr;
    if constexpr (SP::NEED_W01) { s_w01 = reinterpret_cast<float2*>(p);  p += align16(TL * G::BPT * 8); }
    unsigned long long* s_endmask = reinterpret_cast<unsigned long long*>(p);  p += align16(TL * 8);
    int* s_slot0 = reinterpret_cast<int*>(p);  p += align16(TL * 4);
    int* s_slot1 = reinterpret_cast<int*>(p);  p += align16(TL * 4);
    int* s_pb = reinterpret_cast<int*>(p);  p += align16((a.n_mels + 2) * 4);
    int* s_ov = reinterpret_cast<int*>(p);  p += align16(64 * 4);
    unsigned& s_tmem_base = *reinterpret_cast<unsigned*>(p);  p += 16;      // TMEM allocation of this CTA
    float* s_win = nullptr;
    if constexpr (SP::NEED_WIN) { s_win = reinterpret_cast<float*>(p);  p += align16(G::N * 4); }
    float2* s_tw_lin = nullptr;
    if constexpr (SP::NEED_TWLIN) { s_tw_lin = reinterpret_cast<float2*>(p);  p += align16(G::N * 8); }
    unsigned long long* s_mbar = nullptr;                                    // fused GCC: "the MMAs of team t have completed"
    if constexpr (FUSED) { s_mbar = reinterpret_cast<unsigned long long*>(p);  p += 64; }
    int* s_lane_beg = nullptr;
    int* s_gtab = nullptr;
    if constexpr (LANES) {
        s_lane_beg = reinterpret_cast<int*>(p);  p += align16(TL * 4);
        s_gtab = reinterpret_cast<int*>(p);  p += align16(64 * kLaneGatherMax * 4);
        for (int i = threadIdx.x; i < TL; i += blockDim.x) s_lane_beg[i] = a.lane_beg[i];
        for (int i = threadIdx.x; i < 64 * kLaneGatherMax; i += blockDim.x) s_gtab[i] = a.gtab[i];
    }
    constexpr bool BULK = kBulkTail && MODE == MODE_FOA && R == 32 && !EDGE && LAYOUT == LAYOUT_PLANAR_CL;
    unsigned long long* s_ldbar = nullptr;                                   // bulk-staged tail taps: one mbarrier per warp
    constexpr bool BULKM = kBulkMic && FUSED && !EDGE && LAYOUT == LAYOUT_PLANAR_CL;      // fused MIC: the whole next frame staged
    if constexpr ((kBulkTail && MODE == MODE_FOA && R == 32) || (kBulkMic && FUSED)) { s_ldbar = reinterpret_cast<unsigned long long*>(p);  p += 128; }
    const float wscale = ((EDGE ? a.layout : LAYOUT) == LAYOUT_PCM16_LC) ? (1.0f / 32768.0f) : 1.0f;    // exact: folds the PCM decode
    for (int i = threadIdx.x; i < G::N; i += blockDim.x) {
        if constexpr (SP::NEED_WIN) s_win[i] = a.window[i] * wscale;
        if constexpr (SP::NEED_TW) s_tw_t[i] = a.tw_t[i];
        if constexpr (SP::NEED_TWLIN) s_tw_lin[i] = a.tw_lin[i];
    }
    if constexpr (SP::NEED_W01) { for (int i = threadIdx.x; i < TL * G::BPT; i += blockDim.x) s_w01[i] = a.w01[i]; }
    for (int i = threadIdx.x; i < a.n_mels + 2; i += blockDim.x) s_pb[i] = a.pb[i];
    if (threadIdx.x < TL) {
        s_endmask[threadIdx.x] = a.endmask[threadIdx.x];
        s_slot0[threadIdx.x] = a.slot0[threadIdx.x];
        s_slot1[threadIdx.x] = a.slot1[threadIdx.x];
        s_ov[threadIdx.x] = a.ov[threadIdx.x];
    }
    if constexpr (BULK || BULKM) {
        if (threadIdx.x < 16) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_addr(&s_ldbar[threadIdx.x])));
            asm volatile("fence.mbarrier_init.release.cluster;");
        }
    }
    if constexpr (FUSED) {
        if (threadIdx.x < 8) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 2;" :: "r"(smem_addr(&s_mbar[threadIdx.x])));      // one commit per warp of the team
            asm volatile("fence.mbarrier_init.release.cluster;");
        }
    }
    unsigned taddr = 0;
    if constexpr (TM) {
        static_assert(!TM || 2 * G::BPT <= 32, "mel weights of a lane must fit 32 TMEM columns");
        if (warp == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                         :: "r"(smem_addr(&s_tmem_base)), "r"(TMEM_COLS));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;");
        taddr = s_tmem_base + ((32u * (warp & 3)) << 16);
        if (warp < 4) {                           // one writer per TMEM lane quadrant
            float r[16];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
#pragma unroll
                for (int i = 0; i < 16; ++i) r[i] = a.window[lane + 32 * (16 * c + i)] * wscale;
                tmem_st16(taddr + TMEM_COL_WIN + 16 * c, r);
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) {
#pragma unroll
                for (int i = 0; i < 8; ++i) { const float2 t = a.tw_t[(8 * g + i) * 32 + lane]; r[2 * i] = t.x; r[2 * i + 1] = t.y; }
                tmem_st16(taddr + TMEM_COL_TW + 16 * g, r);
            }
            if constexpr (LANES) {                // 4 weights per bin step: columns 96 .. 131
                const float* w4 = a.w4 + size_t(u) * (4 * G::BPT);
#pragma unroll
                for (int c = 0; c < 2; ++c) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) r[i] = w4[16 * c + i];
                    tmem_st16(taddr + TMEM_COL_W01 + 16 * c, r);
                }
                tmem_st2(taddr + TMEM_COL_W01 + 32, w4[32], w4[33]);
                tmem_st2(taddr + TMEM_COL_W01 + 34, w4[34], w4[35]);
            } else {
                const float* w01 = reinterpret_cast<const float*>(a.w01) + size_t(u) * (2 * G::BPT);     // u: quadrant parity == warp parity
#pragma unroll
                for (int c = 0; c < 2; ++c) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) r[i] = (16 * c + i < 2 * G::BPT) ? w01[16 * c + i] : 0.f;
                    tmem_st16(taddr + TMEM_COL_W01 + 16 * c, r);
                }
            }
            if constexpr (FUSED) {
                // the lag-projection basis, A operand of every MMA of this kernel: lane l of quadrant q holds lag row
                // 16 q + (l & 15), K half (l >> 4), two fp16 per column -> 256 columns
                const uint4* src = reinterpret_cast<const uint4*>(a.gcc_basis) + (size_t(16 * warp + (lane & 15)) * 1024 + size_t(lane >> 4) * 512) / 8;
#pragma unroll 1
                for (int g = 0; g < 16; ++g) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint4 q = __ldg(src + 4 * g + i);
                        r[4 * i] = __uint_as_float(q.x); r[4 * i + 1] = __uint_as_float(q.y);
                        r[4 * i + 2] = __uint_as_float(q.z); r[4 * i + 3] = __uint_as_float(q.w);
                    }
                    tmem_st16(taddr + TMEM_COL_BASIS + 16 * g, r);
                }
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;");
    }
    // ---- per-team regions
    const int team_stride = SP::team_bytes(a.n_mels, a.x_bytes);
    unsigned char* tp = p + size_t(team) * team_stride;
    unsigned char* tile = tp;                                                        // fused GCC: spectrum / phasor tile + Nyquist column
    unsigned char* nyq = tp + kGccTileBytes;
    float2* S0 = reinterpret_cast<float2*>(tp);                                      // warp 0: exchange, then spectrum of pair 0
    float2* S1 = reinterpret_cast<float2*>(tp + align16(G::E_ELEMS * 8));            // warp 1: exchange, then spectrum of pair 1
    tp += SP::spec_bytes();
    float2* X = reinterpret_cast<float2*>(tp);  tp += a.x_bytes;                    // mel pieces, then GCC exchange
    float* acc = reinterpret_cast<float*>(tp);
    float2* E = h ? S1 : S0;
    const int row_elems = a.n_mels * C;
    for (int i = u; i < a.x_zero_f2; i += TL) X[i] = make_float2(0.f, 0.f);    // segment-major slots without a piece stay zero
    if constexpr (FUSED) {
        static_assert(!FUSED || 2 * align16(G::E_ELEMS * 8) <= kGccTileBytes, "the exchange buffers alias the phasor tile");
        for (int i = u; i < kGccNyqBytes / 4; i += TL) reinterpret_cast<float*>(nyq)[i] = 0.f;
    }
    __syncthreads();
    if constexpr (TM) asm volatile("tcgen05.fence::after_thread_sync;");
    const float* wlane = SP::NEED_WIN ? s_win + lane : nullptr;      // window taps of this lane when they are not in TMEM

    const Tables tb{nullptr, s_tw_t, s_tw_lin, s_w01, s_endmask, s_slot0, s_slot1, s_pb, s_ov, nullptr, s_lane_beg, s_gtab, a.gather_n0, a.gather_n1};
    const long long total_frames = (long long)a.n_clips * a.frames_per_clip;

    float run_max = -INFINITY;
    int run_clip = -1;
    // fused GCC: partner team of the pair, its staged row, phase parities of the two teams' MMA barriers
    const int pteam = team ^ 1;
    float* acc_part = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(acc) + (pteam - team) * team_stride);
    unsigned par_pair = 0;
    const unsigned tmem_base = taddr - ((32u * (warp & 3)) << 16);
    volatile unsigned* dead_flags = reinterpret_cast<volatile unsigned*>(nyq + GT_NYQ_BYTES);     // [2]: bits (channel 2h, 2h + 1) of warp h

    // forward FFT stage 2 of this warp; the fused kernel lays the spectrum out as the MMA's B tile (both warps' rows
    // interleave there, so every lane must hold its column before anyone stores)
    auto stage2 = [&]() {
        if constexpr (FUSED) {
            float2 uu[32];
            stage2_load_fft<R, PACKED_TW ? R + 1 : G::EP>(E, uu, lane);
            team_bar(bar_id);
            stage2_store_tile(uu, tile, nyq, lane, h);
        } else {
            stage2_forward<R>(E, E, lane);
        }
    };
    // fused GCC: is a channel of this warp's pair exactly zero over the whole windowed frame?  (the reference's phase
    // transform of a dead microphone is a sign pattern of the live partner, extract_core.cuh: dead_pair)
    auto note_dead = [&](const float2* v) {
        if constexpr (FUSED && kGccDead) {
            // quick reject on four of the 32 taps (all-zero there is necessary for a dead channel); the full scan only runs when
            // one of the two channels passes it -- digital silence -- so a live frame pays 12 instructions instead of 40
            unsigned ox = __float_as_uint(v[5].x) | __float_as_uint(v[12].x) | __float_as_uint(v[19].x) | __float_as_uint(v[26].x);
            unsigned oy = __float_as_uint(v[5].y) | __float_as_uint(v[12].y) | __float_as_uint(v[19].y) | __float_as_uint(v[26].y);
            unsigned dx = __all_sync(0xffffffffu, (ox << 1) == 0), dy = __all_sync(0xffffffffu, (oy << 1) == 0);
            if (dx | dy) {                               // (warp-uniform)
#pragma unroll
                for (int n2 = 0; n2 < R; ++n2) { ox |= __float_as_uint(v[n2].x); oy |= __float_as_uint(v[n2].y); }
                dx = __all_sync(0xffffffffu, (ox << 1) == 0);
                dy = __all_sync(0xffffffffu, (oy << 1) == 0);
            }
            if (lane == 0) dead_flags[h] = dx | (dy << 1);
        }
    };

    // running clip maximum of this warp; a NaN log-mel value makes it NaN (reference: db.max(), feature_extractor.py:65-71):
    // every maximum on the way is NaN-propagating, the key of a NaN maximum sorts above +inf
    auto max_key = [&](float m) -> unsigned { return (m != m) ? kNanKey : float_to_key(m); };
    auto note_max = [&](int clip, float mx) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = max_nan(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (clip != run_clip) {
            if (run_clip >= 0 && lane == 0) atomicMax(&a.clip_max_key[run_clip], max_key(run_max));
            run_clip = clip;
            run_max = -INFINITY;
        }
        run_max = max_nan(run_max, mx);
    };

    // everything after the team's two packed FFTs: mel pieces, gather, GCC, row store, running clip maximum.
    // mine: this team has a frame; part (fused GCC only): the partner team has one whose accumulator this team helps read.
    auto finish_frame = [&](bool mine, bool part, int clip, int t, float* row, auto&& early) {
        float mx = -INFINITY;
        if constexpr (FUSED) {
            const int pair = team >> 1;
            if (mine) {
                team_bar(bar_id);                                        // both spectra are in the tile
                const unsigned dead = kGccDead ? (dead_flags[0] | (dead_flags[1] << 2)) : 0u;
                if (dead) bin_phase_gcc_fused<true>(tile, nyq, tb, X, u, taddr + TMEM_COL_W01, dead);
                else bin_phase_gcc_fused<false>(tile, nyq, tb, X, u, taddr + TMEM_COL_W01, 0u);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the phasor rows -> visible to the tensor core
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            }
            pair_bar(9 + pair);                                          // both teams' rows are in place (or absent)
            // one MMA chain per pair: the even team issues if it has a frame, else the odd one
            if (mine && (!(team & 1) || !part)) {
                if (elect_one())
                    gcc_issue_mma(smem_addr(tile) - unsigned(team & 1) * unsigned(team_stride), unsigned(team_stride), tmem_base, gcc_dcol(pair),
                                  smem_addr(&s_mbar[pair]), h);
                __syncwarp();
            }
            if (mine) mx = gather_lanes<MODE>(X, tb, acc, a.n_mels, u);  // log-mel while the MMAs run
            mbar_wait_parity(smem_addr(&s_mbar[pair]), par_pair);
            par_pair ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (mine) early();                                           // (the tile is free: the next frame's samples are staged into it)
            {
                float* mine_row = mine ? acc : nullptr;
                float* part_row = part ? acc_part : nullptr;
                gcc_epilogue(taddr, gcc_dcol(pair), warp & 3, lane, (team & 1) ? part_row : mine_row, (team & 1) ? mine_row : part_row);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            pair_bar(9 + pair);                                          // both rows are complete; the accumulator is free
            if (!mine) return;
        } else {
            team_bar(bar_id);                                            // both spectra are in place
            if constexpr (LANES) bin_phase_lanes<R, TM, TFV>(S0, S1, tb, X, 1e-8f, u, taddr + TMEM_COL_W01);
            else bin_phase<R, MODE, TM, TFV>(S0, S1, tb, X, 1e-8f, u, taddr + TMEM_COL_W01);
            team_bar(bar_id);
            early();                                                     // (FOA: the next frame's new samples are requested here)
            if constexpr (LANES) mx = gather_records<TFV>(X, tb, acc, a.n_mels, u);
            else mx = a.seg_major ? gather_lanes<MODE, TFV>(X, tb, acc, a.n_mels, u) : gather_phase<MODE, TFV>(X, tb, acc, a.n_mels, u);
            if constexpr (MODE == MODE_MIC) {
                team_bar(bar_id);                                        // pieces consumed: X is the GCC exchange buffer now
                if (h == 0) {
                    gcc_stage1<R, 0>(S0, S1, X, lane);
                    __syncwarp();
                    gcc_stage2<R, 0>(X, tb, acc, a.n_mels, lane);
                    __syncwarp();
                    gcc_stage1<R, 1>(S0, S1, X, lane);
                    __syncwarp();
                    gcc_stage2<R, 1>(X, tb, acc, a.n_mels, lane);
                    __syncwarp();
                    gcc_stage1<R, 2>(S0, S1, X, lane);
                    __syncwarp();
                    gcc_stage2<R, 2>(X, tb, acc, a.n_mels, lane);
                    __syncwarp();
                    for (int i = lane; i < a.x_zero_f2; i += 32) X[i] = make_float2(0.f, 0.f);   // X goes back to being the piece buffer
                }
            }
            team_bar(bar_id);                                            // the row is complete
        }
        // (an asynchronous bulk store of the row -- fence.proxy.async + cp.async.bulk shared -> global by one lane -- measured
        //  no faster than these 128-bit copies: 9.99 vs 9.93 ms)
        if (row != nullptr) store_row(acc, row_elems, row, u, TL);
        note_max(clip, mx);
    };

    if constexpr (EDGE) {
        // edge frame j of a clip -> frame index t; -1 for the zero padding rows past the last frame
        auto edge_t = [&](long long g) -> int {
            const int j = int(g % a.frames_per_clip);
            return (j < a.t_lo) ? j : a.t_hi + (j - a.t_lo);
        };
        for (long long sc = blockIdx.x; sc < a.n_super; sc += gridDim.x) {
            for (int i = 0; i < a.fpw; ++i) {
                const long long g = sc * a.fsc + team * a.fpw + i;
                const long long gp = sc * a.fsc + pteam * a.fpw + i;
                const bool in_range = g < total_frames;
                const bool part = FUSED && gp < total_frames && edge_t(gp) < a.t_raw;
                if (!in_range && !part) continue;
                int clip = 0, t = 0;
                float* row = nullptr;
                bool mine = false;
                if (in_range) {
                    clip = int(g / a.frames_per_clip);
                    t = edge_t(g);
                    row = (t < a.t_out) ? a.out + ((long long)clip * a.t_out + t) * row_elems : nullptr;
                    if (t >= a.t_raw) {                   // zero padding rows (reference :142-145)
                        for (int e = u; e < row_elems; e += TL) row[e] = 0.f;
                    } else {
                        mine = true;
                    }
                }
                if (mine) {
                    ClipSrc src;
                    src.base = a.wav + (long long)clip * 4 * a.n_samples;
                    src.n_samples = a.n_samples;
                    if (a.layout == LAYOUT_PLANAR_CL) { src.chan_stride = a.n_samples; src.samp_stride = 1; }
                    else { src.chan_stride = 1; src.samp_stride = 4; }
                    const long long start = (long long)t * a.hop - G::N / 2 + a.origin;
                    {
                        float wreg[R];
                        if constexpr (TM) {
                            tmem_ld16(taddr + TMEM_COL_WIN, wreg);
                            tmem_ld16(taddr + TMEM_COL_WIN + 16, wreg + 16);
                        } else {
#pragma unroll
                            for (int n2 = 0; n2 < R; ++n2) wreg[n2] = wlane[32 * n2];
                        }
                        float2 v[R];
                        if (a.layout == LAYOUT_PCM16_LC)
                            stage1_load_reflect_pcm16<R>(reinterpret_cast<const short*>(a.wav) + (long long)clip * 4 * a.n_samples,
                                                         a.n_samples, h, start, wreg, v, lane);
                        else
                            stage1_load_reflect<R>(src, 2 * h, 2 * h + 1, start, wreg, v, lane, TFV);
                        note_dead(v);
                        if constexpr (TM) stage1_fft_store_tm<R, PACKED_TW>(v, taddr + TMEM_COL_TW, E, lane);
                        else stage1_fft_store<R>(v, tb, E, lane);
                    }
                    __syncwarp();
                    stage2();
                }
                if (mine || part) finish_frame(mine, part, clip, t, row, [] {});
            }
        }
    } else {
        // Interior frames, software-pipelined: the raw samples of the team's NEXT frame (this warp's channel pair) are
        // requested before the current frame's FFT starts, so HBM/L2 latency hides behind the arithmetic (requesting them
        // later -- after the bin phase -- measured 6 % slower).
        // Frame indices are 32-bit here (launch_one refuses launches of 2^31 frames or more) and (clip, t) of the NEXT frame is
        // tracked incrementally: one division per run of fpw frames instead of two 64-bit divisions per frame (those and the
        // 64-bit index products were ~150 of the ~1800 instructions a warp spends on a frame).
        const int sc_step = gridDim.x;                    // super-chunk sc -> CTA sc mod grid
        int sc = blockIdx.x;
        const int sc_end = int(a.n_super);
        const int n_frames = int(total_frames);
        const int fpc = a.frames_per_clip;
        int fi = 0;
        int run_base = sc * a.fsc + team * a.fpw;         // frame index of the team's frame fi = 0 of super-chunk sc
        auto frame_at = [&](int s, int base, int i) -> int {        // -1 past the end
            return (s < sc_end && base + i < n_frames) ? base + i : -1;
        };
        struct Pos { int clip, t; };
        auto locate = [&](int g) -> Pos {                 // (one 32-bit division: only at the first frame of a run)
            const int c = int(unsigned(g) / unsigned(fpc));
            return Pos{c, a.t_lo + (g - c * fpc)};
        };
        float2 raw[R];
        long long start = 0;
        int clip = 0, t = 0;
        auto request_at = [&](Pos p) {              // issue the loads of the frame at (clip, t) (this warp's pair) into raw[]
            clip = p.clip;
            t = p.t;
            start = (long long)t * a.hop - G::N / 2 + a.origin;
            if constexpr (LAYOUT == LAYOUT_PCM16_LC) {
                stage1_load_raw_pcm16_pair<R>(reinterpret_cast<const short*>(a.wav) + (long long)clip * 4 * a.n_samples, h, start, raw, lane);
            } else {
                ClipSrc src;
                src.base = a.wav + (long long)clip * 4 * a.n_samples;
                src.n_samples = a.n_samples;
                if constexpr (LAYOUT == LAYOUT_PLANAR_CL) { src.chan_stride = a.n_samples; src.samp_stride = 1; }
                else { src.chan_stride = 1; src.samp_stride = 4; }
                stage1_load_raw<R, LAYOUT>(src, 2 * h, 2 * h + 1, start, raw, lane);
            }
        };
        auto request = [&](int g) { request_at(locate(g)); };
        // FOA at n_fft = 1024 runs four warps per scheduler at 128 registers: the 2R prefetch registers do not fit, the loads
        // are issued at the top of the frame instead and the other warps cover their latency.  Measured alternatives: requesting
        // them after the bin phase spills 376 B (12.3 ms instead of 9.2); staging them through tensor memory four taps per
        // bin step (tcgen05.st, then four tcgen05.ld at the next frame) couples the load latency into the team barriers (10.5 ms)
        constexpr bool PREFETCH = !(R == 32 && (MODE == MODE_FOA || SELD_MIC_WARPS > 12));      // (experiment: MIC at 16 warps has no registers for it either)
        const int part_off = (pteam - team) * a.fpw;      // the partner team's frames sit fpw further (or back)
        int g = frame_at(sc, run_base, fi);
        int gpart = FUSED ? frame_at(sc, run_base + part_off, fi) : -1;
        Pos next_pos = g >= 0 ? locate(g) : Pos{0, 0};    // (clip, t) of frame g, carried one iteration ahead
        if (PREFETCH && g >= 0) request(g);
        // Consecutive frames share R - 15 taps per lane (hop 480 = 15 * 32 samples: tap n2 of frame t + 1 is tap n2 + 15 of
        // frame t).  The 16-warp kernel parks those 17 taps in its own tensor-memory columns and fetches only the 15 new
        // ones from global memory for the following frame: half the global-load wavefronts, half the L2 requests.
        constexpr int SH = 15;
        constexpr bool KEEP = !PREFETCH && TM && R == 32 && MODE == MODE_FOA;      // (the fused MIC kernel's tensor memory is full: tables + basis + accumulators)
        const bool keep_ok = KEEP && a.hop == SH * 32;
        const unsigned tkeep = taddr + TMEM_COL_KEEP + (LANES ? 8 : 0) + 64 * (warp >> 2);      // (the lane form's weights end at column 132)
        int prev_clip = -1, prev_t = -2;
        // taps [n_lo, n_hi) of the frame starting at sample `fs` of clip `c` (this warp's channel pair) -> r[]
        auto load_taps = [&](int c, long long fs, int n_lo, int n_hi, float2* r) {
            if constexpr (LAYOUT == LAYOUT_PCM16_LC) {
                const float* p = reinterpret_cast<const float*>(reinterpret_cast<const short*>(a.wav) + ((long long)c * a.n_samples + fs + lane) * 4 + 2 * h);
#pragma unroll
                for (int n2 = 0; n2 < R; ++n2) if (n2 >= n_lo && n2 < n_hi) r[n2].x = p[64 * n2];
            } else if constexpr (LAYOUT == LAYOUT_INTERLEAVED_LC) {
                const float2* p = reinterpret_cast<const float2*>(a.wav + ((long long)c * a.n_samples + fs + lane) * 4 + 2 * h);
#pragma unroll
                for (int n2 = 0; n2 < R; ++n2) if (n2 >= n_lo && n2 < n_hi) r[n2] = p[64 * n2];
            } else {
                const float* pa = a.wav + ((long long)c * 4 + 2 * h) * a.n_samples + fs + lane;
                const float* pb = pa + a.n_samples;
#pragma unroll
                for (int n2 = 0; n2 < R; ++n2) if (n2 >= n_lo && n2 < n_hi) r[n2] = make_float2(pa[32 * n2], pb[32 * n2]);
            }
        };
        // (kEarlyTail: the SH new taps of the team's NEXT frame requested right after the bin phase)
        bool have_tail = false;                       // every frame but a warp's first gets its tail early
        unsigned ld_parity = 0;
        const bool bulk_ok = (BULK || BULKM) && a.hop % 4 == 0 && a.n_samples % 4 == 0 && a.origin % 4 == 0 &&
                             (reinterpret_cast<unsigned long long>(a.wav) & 15ull) == 0;      // 16-byte aligned bulk copies
        bool staged = false;                          // fused MIC: this frame's samples were staged by the previous iteration
        // (prefetching just the 15 new taps of the next frame in 30 registers was tried on top of this: 104 B of spills and
        //  9.58 ms instead of 9.06)
        // (Fused GCC: deferring a frame's epilogue into the next iteration -- accumulators and staged rows double-buffered, the
        //  MMAs running under the next frame's FFTs -- was built and measured: 18.5 ms per 600 clips against 17.3 for waiting
        //  right after the gather; the tensor pipe is 9 % busy, the wait is not what this kernel is short of.)
#pragma unroll 1
        while (g >= 0 || gpart >= 0) {
            const bool mine = g >= 0;
            const Pos pos = next_pos;
            if (++fi == a.fpw) { fi = 0; sc += sc_step; run_base += sc_step * a.fsc; }
            const int g_next = frame_at(sc, run_base, fi);
            const int gpart_next = FUSED ? frame_at(sc, run_base + part_off, fi) : -1;
            if constexpr (!PREFETCH) {
                if (g_next >= 0) {
                    if (fi == 0 || !mine) next_pos = locate(g_next);
                    else if (pos.t + 1 == a.t_hi) next_pos = Pos{pos.clip + 1, a.t_lo};
                    else next_pos = Pos{pos.clip, pos.t + 1};
                }
            }
            if constexpr (FUSED) {
                if (!mine) {                              // only the partner has a frame: help read its accumulator
                    finish_frame(false, true, 0, 0, nullptr, [] {});
                    g = g_next;
                    gpart = gpart_next;
                    continue;
                }
            }
            if constexpr (!PREFETCH) {
                if constexpr (KEEP) {
                    clip = pos.clip;
                    t = pos.t;
                    start = (long long)t * a.hop - G::N / 2 + a.origin;
                    if (keep_ok && clip == prev_clip && t == prev_t + 1) {
                        float* rf = reinterpret_cast<float*>(raw);
                        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                        tmem_ld16(tkeep, rf);
                        tmem_ld16(tkeep + 16, rf + 16);
                        tmem_ld2(tkeep + 32, rf[32], rf[33]);                     // taps 0 .. 16
                    } else {
                        load_taps(clip, start, 0, R - SH, raw);
                    }
                    if (BULK && bulk_ok && have_tail) {       // staged by the bulk copies issued after the previous bin phase
                        mbar_wait_parity(smem_addr(&s_ldbar[warp]), ld_parity);
                        ld_parity ^= 1u;
                        const float* ea = reinterpret_cast<const float*>(E) + lane;
#pragma unroll
                        for (int n2 = R - SH; n2 < R; ++n2) raw[n2] = make_float2(ea[32 * (n2 - (R - SH))], ea[32 * SH + 32 * (n2 - (R - SH))]);
                        __syncwarp();                         // every lane has its taps before anyone's stage-1 store lands in E
                    } else if (!(kEarlyTail && have_tail)) {
                        load_taps(clip, start, R - SH, R, raw);   // (kEarlyTail: requested before the previous gather)
                    }
                    have_tail = true;
                    if (keep_ok) {                                                // taps 15 .. 31 are taps 0 .. 16 of the next frame
                        const float* rf = reinterpret_cast<const float*>(raw);
                        tmem_st16(tkeep, rf + 2 * SH);
                        tmem_st16(tkeep + 16, rf + 2 * SH + 16);
                        tmem_st2(tkeep + 32, rf[2 * SH + 32], rf[2 * SH + 33]);
                    }
                    prev_clip = clip;
                    prev_t = t;
                } else if constexpr (BULKM) {
                    if (staged) {
                        clip = pos.clip;
                        t = pos.t;
                        mbar_wait_parity(smem_addr(&s_ldbar[warp]), ld_parity);
                        ld_parity ^= 1u;
                        const float* ea = reinterpret_cast<const float*>(E) + lane;
#pragma unroll
                        for (int n2 = 0; n2 < R; ++n2) raw[n2] = make_float2(ea[32 * n2], ea[32 * R + 32 * n2]);
                        __syncwarp();                         // every lane has its taps before anyone's stage-1 store lands in E
                        staged = false;
                    } else {
                        request_at(pos);
                    }
                } else {
                    request_at(pos);
                }
            }
            float2 v[R];
            if constexpr (TM) {
                float w[R];
                tmem_ld16(taddr + TMEM_COL_WIN, w);
                tmem_ld16(taddr + TMEM_COL_WIN + 16, w + 16);
                if constexpr (LAYOUT == LAYOUT_PCM16_LC) apply_window_pcm16<R>(raw, 0, w, v);
                else apply_window<R>(raw, w, v);
            } else {
                if constexpr (LAYOUT == LAYOUT_PCM16_LC) apply_window_pcm16<R, 32>(raw, 0, wlane, v);
                else apply_window<R, 32>(raw, wlane, v);
            }
            const int clip_now = clip, t_now = t;
            float* row = (t < a.t_out) ? a.out + ((long long)clip * a.t_out + t) * row_elems : nullptr;
            if (PREFETCH && g_next >= 0) request(g_next);
            {
                note_dead(v);
                if constexpr (TM) stage1_fft_store_tm<R, PACKED_TW>(v, taddr + TMEM_COL_TW, E, lane);
                else stage1_fft_store<R>(v, tb, E, lane);
                __syncwarp();
                stage2();
                finish_frame(true, gpart >= 0, clip_now, t_now, row, [&] {
                    if constexpr (BULK) {
                        if (bulk_ok && g_next >= 0 && lane == 0) {
                            const int c2 = next_pos.clip, t2 = next_pos.t;
                            const long long s2 = (long long)t2 * a.hop - G::N / 2 + a.origin + 32 * (R - SH);
                            const float* pa = a.wav + ((long long)c2 * 4 + 2 * h) * a.n_samples + s2;
                            const unsigned bar = smem_addr(&s_ldbar[warp]), dst = smem_addr(E);
                            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // E was read through the generic proxy
                            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(2 * 32 * SH * 4) : "memory");
                            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                         :: "r"(dst), "l"(pa), "r"(32 * SH * 4), "r"(bar) : "memory");
                            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                         :: "r"(dst + 32 * SH * 4), "l"(pa + a.n_samples), "r"(32 * SH * 4), "r"(bar) : "memory");
                        }
                    } else if constexpr (BULKM) {
                        if (bulk_ok && g_next >= 0) {
                            if (lane == 0) {
                                const int c2 = next_pos.clip, t2 = next_pos.t;
                                const long long s2 = (long long)t2 * a.hop - G::N / 2 + a.origin;
                                const float* pa = a.wav + ((long long)c2 * 4 + 2 * h) * a.n_samples + s2;
                                const unsigned bar = smem_addr(&s_ldbar[warp]), dst = smem_addr(E);
                                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // E was written through the generic proxy
                                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(2 * G::N * 4) : "memory");
                                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                             :: "r"(dst), "l"(pa), "r"(G::N * 4), "r"(bar) : "memory");
                                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                             :: "r"(dst + G::N * 4), "l"(pa + a.n_samples), "r"(G::N * 4), "r"(bar) : "memory");
                            }
                            staged = true;
                        }
                    } else if constexpr (KEEP && kEarlyTail) {
                        if (g_next >= 0) {
                            const int c2 = next_pos.clip, t2 = next_pos.t;
                            load_taps(c2, (long long)t2 * a.hop - G::N / 2 + a.origin, R - SH, R, raw);
                        }
                    }
                });
            }
            g = g_next;
            gpart = gpart_next;
        }
    }
    if (run_clip >= 0 && lane == 0) atomicMax(&a.clip_max_key[run_clip], max_key(run_max));
    if constexpr (TM) {
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(s_tmem_base), "r"(TMEM_COLS));
    }
}

__global__ void clip_max_decode_kernel(const unsigned int* keys, int n, float* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = key_to_float(keys[i]);
}

template <int R, int MODE, int LAYOUT, bool EDGE, int TC>
static int launch_one(const seld_plan* plan, ExtractArgs a, cudaStream_t stream) {
    a.frames_per_clip = EDGE ? a.t_lo + (a.t_tot - a.t_hi) : a.t_hi - a.t_lo;
    if (a.frames_per_clip <= 0) return SELD_OK;
    // shared-memory geometry of this kernel variant: as many frame teams as fit, at most max_warps / 2
    using SP = SmemPlan<R, MODE, TC>;
    a.x_bytes = SP::x_bytes(plan->n_slots);
    const int tb = SP::table_bytes(plan->n_mels);
    const int wb = SP::team_bytes(plan->n_mels, a.x_bytes);
    int teams = (plan->max_smem_optin - tb) / wb;
    if (teams > max_warps<R, MODE>() / 2) teams = max_warps<R, MODE>() / 2;
    if (const char* e = getenv("SELD_WARPS")) {           // experiments only: fewer warps per CTA
        const int n = atoi(e) / 2;
        if (n >= 1 && n < teams) teams = n;
    }
    if (SP::tc) teams &= ~1;                              // fused GCC: teams work in pairs (extract_kernel)
    if (teams < 1) { set_error("n_mels too large for the shared-memory budget"); return SELD_EUNSUPPORTED; }
    a.fpw = frames_per_team((R == 32 && MODE == MODE_FOA && !EDGE) ? 8 : 2);
    a.fsc = teams * a.fpw;
    const long long per_super = a.fsc;
    a.n_super = ((long long)a.n_clips * a.frames_per_clip + per_super - 1) / per_super;
    if ((long long)a.n_clips * a.frames_per_clip + (long long)(plan->grid + 1) * per_super >= (1ll << 31)) {
        set_error("seld_extract: 2^31 or more frames in one launch (split the batch)");      // the kernel indexes frames in 32 bits
        return SELD_EINVAL;
    }
    long long grid = a.n_super < plan->grid ? a.n_super : plan->grid;
    static std::atomic<unsigned long long> configured{0};        // bit d: attribute set on device d (per instantiation)
    const unsigned long long bit = 1ull << (plan->device & 63);
    if (!(configured.load(std::memory_order_acquire) & bit)) {
        SELD_CUDA_TRY(cudaFuncSetAttribute(extract_kernel<R, MODE, LAYOUT, EDGE, TC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           plan->max_smem_optin));
        configured.fetch_or(bit, std::memory_order_release);
    }
    extract_kernel<R, MODE, LAYOUT, EDGE, TC><<<(int)grid, teams * 64, tb + teams * wb, stream>>>(a);
    SELD_CUDA_TRY(cudaGetLastError());
    seld::note_launch();
    return SELD_OK;
}

template <int R, int MODE, int TC>
static int launch_kernels(const seld_plan* plan, const ExtractArgs& a, cudaStream_t stream) {
    int rc;
    if (a.layout == LAYOUT_PLANAR_CL) rc = launch_one<R, MODE, LAYOUT_PLANAR_CL, false, TC>(plan, a, stream);
    else if (a.layout == LAYOUT_INTERLEAVED_LC) rc = launch_one<R, MODE, LAYOUT_INTERLEAVED_LC, false, TC>(plan, a, stream);
    else if constexpr (MODE == MODE_FOA && (TC & 1)) { set_error("the TF-variant extractor takes float32 input"); return SELD_EUNSUPPORTED; }
    else rc = launch_one<R, MODE, LAYOUT_PCM16_LC, false, TC>(plan, a, stream);
    if (rc != SELD_OK) return rc;
    return launch_one<R, MODE, LAYOUT_PLANAR_CL, true, TC>(plan, a, stream);  // edge frames: layout taken from a.layout
}

template <int R, int MODE>
static int launch_mode(const seld_plan* plan, const ExtractArgs& a, cudaStream_t stream) {
    constexpr bool can_tc = (MODE == MODE_MIC && R == 32);
    if (can_tc && a.gcc_tc) return launch_kernels<R, MODE, can_tc ? 1 : 0>(plan, a, stream);
    if constexpr (MODE == MODE_FOA && R == 32) {
        if (a.lanes) return a.tf_variant ? launch_kernels<R, MODE, 3>(plan, a, stream) : launch_kernels<R, MODE, 2>(plan, a, stream);
        if (a.tf_variant) return launch_kernels<R, MODE, 1>(plan, a, stream);
    }
    return launch_kernels<R, MODE, 0>(plan, a, stream);
}

template <int R>
static int launch_extract(const seld_plan* plan, const ExtractArgs& a, cudaStream_t stream) {
    return plan->mode == SELD_MODE_FOA ? launch_mode<R, MODE_FOA>(plan, a, stream) : launch_mode<R, MODE_MIC>(plan, a, stream);
}

}  // namespace seld

using namespace seld;

extern "C" {

const char* seld_last_error(void) { return g_last_error.c_str(); }
int seld_version(void) { return 2; }
int64_t seld_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int seld_device_check(int device) {
    int dev = device;
    if (dev < 0) {
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) { set_error(std::string("no CUDA device: ") + cudaGetErrorString(e)); return SELD_ENODEVICE; }
    }
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) { set_error(std::string("no CUDA device: ") + cudaGetErrorString(e)); return SELD_ENODEVICE; }
    if (prop.major != 10) {
        char buf[128];
        snprintf(buf, sizeof buf, "seld_b200 needs an sm_100 (B200) device; device %d is sm_%d%d", dev, prop.major,
                 prop.minor);
        set_error(buf);
        return SELD_ENODEVICE;
    }
    return SELD_OK;
}

int seld_plan_create(int sample_rate, int n_fft, int win_length, int hop_length, int n_mels, int n_chan, int mode,
                     const float* window_host, const float* mel_fb_host, seld_plan_t* plan_out) {
    if (!plan_out || !window_host || !mel_fb_host) { set_error("null argument"); return SELD_EINVAL; }
    *plan_out = nullptr;
    if (n_fft != 256 && n_fft != 512 && n_fft != 1024 && n_fft != 2048) {
        set_error("n_fft must be one of 256, 512, 1024, 2048");
        return SELD_EUNSUPPORTED;
    }
    if (n_chan != 4) { set_error("the fused extractor needs exactly 4 channels"); return SELD_EUNSUPPORTED; }
    const bool tf_variant = (mode == SELD_MODE_FOA_TF);
    if (tf_variant) {
        if (n_fft != 1024) { set_error("SELD_MODE_FOA_TF is built for n_fft = 1024 (reference data_loader.py:311-312)"); return SELD_EUNSUPPORTED; }
        mode = SELD_MODE_FOA;
    }
    if (mode != SELD_MODE_FOA && mode != SELD_MODE_MIC) { set_error("invalid mode"); return SELD_EINVAL; }
    if (win_length <= 0 || win_length > n_fft || hop_length <= 0 || n_mels <= 0 || sample_rate <= 0) {
        set_error("invalid STFT geometry");
        return SELD_EINVAL;
    }
    if (mode == SELD_MODE_MIC && ((n_mels & 1) || n_mels > n_fft)) {
        set_error("MIC mode needs an even n_mels <= n_fft (the reference concatenates n_mels GCC lags)");
        return SELD_EINVAL;
    }
    int rc = seld_device_check(-1);
    if (rc != SELD_OK) return rc;

    seld_plan* plan = new seld_plan();
    memset(plan, 0, sizeof(*plan));
    plan->sample_rate = sample_rate;
    plan->n_fft = n_fft;
    plan->win_length = win_length;
    plan->hop = hop_length;
    plan->n_mels = n_mels;
    plan->n_chan = n_chan;
    plan->mode = mode;
    plan->tf_variant = tf_variant ? 1 : 0;
    plan->n_bins = n_fft / 2 + 1;
    plan->n_out_ch = (mode == SELD_MODE_FOA) ? 7 : 10;
    cudaGetDevice(&plan->device);
    cudaDeviceGetAttribute(&plan->num_sms, cudaDevAttrMultiProcessorCount, plan->device);
    cudaDeviceGetAttribute(&plan->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, plan->device);

    // piece form of the mel bank (<= 2 adjacent non-zeros per row, increasing centres)
    MelPieces mp;
    const std::string perr = build_mel_pieces(mel_fb_host, plan->n_bins, n_mels, mp);
    if (!perr.empty()) {
        delete plan;
        set_error(perr);
        return SELD_EUNSUPPORTED;
    }
    plan->n_pieces = mp.n_pieces;
    plan->max_pieces_per_seg = mp.max_pieces_per_seg;
    plan->n_slots = mp.n_slots;
    plan->seg_major = mp.seg_major ? 1 : 0;
    std::vector<float> tw(2 * (size_t)n_fft);
    for (int j = 0; j < n_fft; ++j) {
        const double ang = -2.0 * 3.14159265358979323846264338327950288 * double(j) / double(n_fft);
        tw[2 * j] = float(cos(ang));
        tw[2 * j + 1] = float(sin(ang));
    }
    cudaError_t e = cudaSuccess;
    auto up = [&](void** dst, const void* src, size_t bytes) {
        if (e != cudaSuccess) return;
        e = cudaMalloc(dst, bytes);
        if (e == cudaSuccess) e = cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
    };
    up((void**)&plan->window, window_host, sizeof(float) * n_fft);
    up((void**)&plan->twiddle, tw.data(), sizeof(float) * 2 * n_fft);
    std::vector<float> twt(2 * (size_t)n_fft);          // tw_t[k2*32 + lane] = W^(lane*k2): lane-contiguous, conflict-free
    for (int k2 = 0; k2 < n_fft / 32; ++k2)
        for (int l = 0; l < 32; ++l) {
            const int j = (l * k2) % n_fft;
            twt[2 * (k2 * 32 + l)] = tw[2 * j];
            twt[2 * (k2 * 32 + l) + 1] = tw[2 * j + 1];
        }
    up((void**)&plan->tw_t, twt.data(), sizeof(float) * 2 * n_fft);
    up((void**)&plan->w01, mp.w01.data(), sizeof(float) * mp.w01.size());
    up((void**)&plan->endmask, mp.endmask.data(), sizeof(unsigned long long) * mp.endmask.size());
    up((void**)&plan->slot0, mp.slot0.data(), sizeof(int) * mp.slot0.size());
    up((void**)&plan->slot1, mp.slot1.data(), sizeof(int) * mp.slot1.size());
    up((void**)&plan->ov, mp.ov.data(), sizeof(int) * mp.ov.size());
    up((void**)&plan->pb, mp.pb.data(), sizeof(int) * mp.pb.size());
    plan->lanes_ok = mp.lanes_ok ? 1 : 0;
    plan->gather_n[0] = mp.gather_n[0];
    plan->gather_n[1] = mp.gather_n[1];
    up((void**)&plan->w4, mp.w4.data(), sizeof(float) * mp.w4.size());
    up((void**)&plan->lane_beg, mp.lane_beg.data(), sizeof(int) * mp.lane_beg.size());
    up((void**)&plan->gtab, mp.gtab.data(), sizeof(int) * mp.gtab.size());
    if (e != cudaSuccess) {
        seld_plan_destroy(plan);
        return cuda_fail(e, "plan table upload");
    }
    if (e == cudaSuccess && mode == SELD_MODE_MIC && n_fft == 1024 && n_mels == 64) {
        // basis of the tensor-core lag projection (extract_core.cuh): [64 lags][K = 1024] fp16, x512, row-major.
        // K = 0: Re P[0], K = 1: Re P[512] (rides in the unused Im P[0] slot), K = 2k / 2k + 1: Re / Im P[k].
        std::vector<__half> bt((size_t)64 * 1024);
        const double two_pi = 2.0 * 3.14159265358979323846264338327950288;
        for (int j = 0; j < 64; ++j) {
            const int lag = j - 32;
            for (int K = 0; K < 1024; ++K) {
                double val;
                const int k = K >> 1;
                if (K == 0) val = 0.5;                                     // 512 * (1/1024): DC
                else if (K == 1) val = (lag & 1) ? -0.5 : 0.5;             // Nyquist: cos(pi lag) / 1024 * 512
                else {
                    const double ang = two_pi * double(k) * double(lag) / 1024.0;
                    val = (K & 1) ? -sin(ang) : cos(ang);                  // 512 * (2/1024) * {cos, -sin}
                }
                bt[(size_t)j * 1024 + K] = __float2half(float(val));
            }
        }
        up((void**)&plan->gcc_bt, bt.data(), sizeof(__half) * bt.size());
    }
    if (e != cudaSuccess) {
        seld_plan_destroy(plan);
        return cuda_fail(e, "gcc basis upload");
    }
    // every kernel variant needs room for at least one frame team (checked again, per variant, at launch)
    {
        const int n_ch = (mode == SELD_MODE_FOA) ? 7 : 10;
        const long long need = 3ll * (n_fft + 64) * 8 + (long long)plan->n_slots * 56 + (long long)n_mels * n_ch * 4 + 5ll * n_fft * 4 + 8192;
        if (need > plan->max_smem_optin) {
            seld_plan_destroy(plan);
            set_error("n_mels too large for the shared-memory budget");
            return SELD_EUNSUPPORTED;
        }
    }
    plan->grid = plan->num_sms;
    plan->stats_blocks = plan->num_sms * 4;
    *plan_out = plan;
    return SELD_OK;
}

int seld_plan_destroy(seld_plan_t plan) {
    if (!plan) return SELD_OK;
    cudaFree(plan->window);
    cudaFree(plan->twiddle);
    cudaFree(plan->tw_t);
    cudaFree(plan->w01);
    cudaFree(plan->endmask);
    cudaFree(plan->slot0);
    cudaFree(plan->slot1);
    cudaFree(plan->ov);
    cudaFree(plan->pb);
    cudaFree(plan->w4);
    cudaFree(plan->lane_beg);
    cudaFree(plan->gtab);
    cudaFree(plan->gcc_bt);
    delete plan;
    return SELD_OK;
}

int seld_plan_out_channels(seld_plan_t plan) { return plan ? plan->n_out_ch : SELD_EINVAL; }

int64_t seld_plan_num_frames(seld_plan_t plan, int64_t n_samples) {
    if (!plan || n_samples < 0) return SELD_EINVAL;
    return 1 + n_samples / plan->hop;
}

static int extract_common(seld_plan_t plan, const void* wav_void, int layout, int n_clips, int64_t n_samples, int t_out,
                          float* feat_raw_dev, uint32_t* clip_max_key_dev, void* workspace_dev, int64_t workspace_bytes,
                          void* stream, bool centered = true, bool pad_end = false) {
    const float* wav_dev = static_cast<const float*>(wav_void);
    if (!plan || !wav_dev || !feat_raw_dev || !clip_max_key_dev) { set_error("null argument"); return SELD_EINVAL; }
    if (layout != LAYOUT_PLANAR_CL && layout != LAYOUT_INTERLEAVED_LC && layout != LAYOUT_PCM16_LC) { set_error("invalid layout"); return SELD_EINVAL; }
    if (reinterpret_cast<uintptr_t>(wav_void) % 16 != 0) { set_error("wav_dev must be 16-byte aligned"); return SELD_EINVAL; }
    if (n_clips < 0 || t_out < 0) { set_error("negative size"); return SELD_EINVAL; }
    if (centered && n_samples <= plan->n_fft / 2) {
        set_error("reflect padding needs n_fft/2 < number of samples (torch.stft raises here too)");
        return SELD_EINVAL;
    }
    if (!centered && !pad_end && n_samples < plan->n_fft) { set_error("an uncentred chunk needs at least n_fft samples"); return SELD_EINVAL; }
    if (pad_end != (plan->tf_variant != 0)) { set_error("seld_extract_tf and SELD_MODE_FOA_TF plans go together"); return SELD_EINVAL; }
    if (pad_end && n_samples < 1) { set_error("empty clip"); return SELD_EINVAL; }
    if (n_clips == 0) return SELD_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ExtractArgs a;
    a.wav = wav_dev;
    a.layout = layout;
    a.n_clips = n_clips;
    a.n_samples = n_samples;
    a.t_raw = centered ? int(1 + n_samples / plan->hop) : int(1 + (n_samples - plan->n_fft) / plan->hop);
    if (pad_end) a.t_raw = int((n_samples + plan->hop - 1) / plan->hop);      // tf.signal.stft(pad_end=True): ceil(L / hop) frames
    a.tf_variant = plan->tf_variant;
    a.origin = centered ? 0 : plan->n_fft / 2;
    a.t_out = t_out;
    a.t_tot = a.t_raw > t_out ? a.t_raw : t_out;
    a.out = feat_raw_dev;
    a.clip_max_key = clip_max_key_dev;
    a.window = plan->window;
    a.tw_t = reinterpret_cast<const float2*>(plan->tw_t);
    a.tw_lin = reinterpret_cast<const float2*>(plan->twiddle);
    a.w01 = reinterpret_cast<const float2*>(plan->w01);
    a.endmask = plan->endmask;
    a.slot0 = plan->slot0;
    a.slot1 = plan->slot1;
    a.ov = plan->ov;
    a.pb = plan->pb;
    // MIC at the production geometry: the lag projection runs on the tensor cores inside the extractor (no scratch).
    // workspace_bytes < 0 asks for the CUDA-core inverse transforms instead (tests compare the two).
    (void)workspace_dev;
    a.gcc_tc = (plan->gcc_bt != nullptr && plan->seg_major && workspace_bytes >= 0) ? 1 : 0;
    a.gcc_basis = plan->gcc_bt;
    a.seg_major = plan->seg_major;
    a.x_zero_f2 = plan->seg_major ? plan->n_slots * (plan->mode == SELD_MODE_FOA ? PieceGeo<MODE_FOA>::PSTRIDE : PieceGeo<MODE_MIC>::PSTRIDE) : 0;
    a.lanes = (plan->lanes_ok && plan->mode == SELD_MODE_FOA && plan->n_fft == 1024 && !getenv("SELD_NO_LANES")) ? 1 : 0;
    a.w4 = plan->w4;
    a.lane_beg = plan->lane_beg;
    a.gtab = plan->gtab;
    a.gather_n0 = plan->gather_n[0];
    a.gather_n1 = plan->gather_n[1];
    if (a.lanes) a.x_zero_f2 = (kLaneZeroRec + 1) * (kLaneRecWords / 2);      // (only the never-written zero record matters)
    // frames [t_lo, t_hi) need no reflection: t*hop - n_fft/2 >= 0 and t*hop + n_fft/2 <= n_samples
    {
        const long long half = plan->n_fft / 2;
        long long lo = centered ? (half + plan->hop - 1) / plan->hop : 0;
        long long hi = centered ? ((n_samples >= half) ? (n_samples - half) / plan->hop + 1 : 0) : a.t_raw;
        if (pad_end) hi = (n_samples >= plan->n_fft) ? (n_samples - plan->n_fft) / plan->hop + 1 : 0;    // frames wholly inside the clip
        if (hi > a.t_raw) hi = a.t_raw;
        if (lo > hi) lo = hi;
        a.t_lo = int(lo);
        a.t_hi = int(hi);
    }
    a.hop = plan->hop;
    a.n_mels = plan->n_mels;
    SELD_CUDA_TRY(cudaMemsetAsync(clip_max_key_dev, 0, sizeof(uint32_t) * n_clips, st));
    switch (plan->n_fft) {
        case 256: return launch_extract<8>(plan, a, st);
        case 512: return launch_extract<16>(plan, a, st);
        case 1024: return launch_extract<32>(plan, a, st);
        default: return launch_extract<64>(plan, a, st);
    }
}

int64_t seld_extract_workspace_bytes(seld_plan_t plan, int n_clips, int64_t n_samples, int t_out) {
    (void)plan; (void)n_clips; (void)n_samples; (void)t_out;
    return 0;           // the fused tensor-core GCC path needs no scratch (kept for ABI stability)
}

int seld_extract(seld_plan_t plan, const float* wav_dev, int layout, int n_clips, int64_t n_samples, int t_out,
                 float* feat_raw_dev, uint32_t* clip_max_key_dev, void* workspace_dev, int64_t workspace_bytes, void* stream) {
    if (layout != SELD_LAYOUT_PLANAR_CL && layout != SELD_LAYOUT_INTERLEAVED_LC) { set_error("invalid layout"); return SELD_EINVAL; }
    return extract_common(plan, wav_dev, layout, n_clips, n_samples, t_out, feat_raw_dev, clip_max_key_dev, workspace_dev,
                          workspace_bytes, stream);
}

int seld_extract_pcm16(seld_plan_t plan, const int16_t* pcm_dev, int n_clips, int64_t n_samples, int t_out,
                       float* feat_raw_dev, uint32_t* clip_max_key_dev, void* workspace_dev, int64_t workspace_bytes,
                       void* stream) {
    return extract_common(plan, pcm_dev, LAYOUT_PCM16_LC, n_clips, n_samples, t_out, feat_raw_dev, clip_max_key_dev,
                          workspace_dev, workspace_bytes, stream);
}

int seld_extract_chunks(seld_plan_t plan, const float* wav_dev, int layout, int n_chunks, int64_t n_samples, int t_out,
                        float* feat_raw_dev, uint32_t* chunk_max_key_dev, void* workspace_dev, int64_t workspace_bytes,
                        void* stream) {
    if (layout != SELD_LAYOUT_PLANAR_CL && layout != SELD_LAYOUT_INTERLEAVED_LC) { set_error("invalid layout"); return SELD_EINVAL; }
    return extract_common(plan, wav_dev, layout, n_chunks, n_samples, t_out, feat_raw_dev, chunk_max_key_dev, workspace_dev,
                          workspace_bytes, stream, /*centered=*/false);
}

int seld_extract_tf(seld_plan_t plan, const float* wav_dev, int layout, int n_clips, int64_t n_samples, int t_out,
                    float* feat_raw_dev, uint32_t* clip_max_key_dev, void* stream) {
    if (layout != SELD_LAYOUT_PLANAR_CL && layout != SELD_LAYOUT_INTERLEAVED_LC) { set_error("invalid layout"); return SELD_EINVAL; }
    return extract_common(plan, wav_dev, layout, n_clips, n_samples, t_out, feat_raw_dev, clip_max_key_dev, nullptr, 0, stream,
                          /*centered=*/false, /*pad_end=*/true);
}

int seld_clip_max_decode(const uint32_t* clip_max_key_dev, int n_clips, float* clip_max_dev, void* stream) {
    if (!clip_max_key_dev || !clip_max_dev || n_clips < 0) { set_error("bad argument"); return SELD_EINVAL; }
    if (n_clips == 0) return SELD_OK;
    clip_max_decode_kernel<<<(n_clips + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(clip_max_key_dev,
                                                                                                  n_clips, clip_max_dev);
    SELD_CUDA_TRY(cudaGetLastError());
    seld::note_launch();
    return SELD_OK;
}

}  // extern "C"
