// Fused training-batch augmentation, sm_100a: ONE pass over x[B][T][F][C] that applies, per sample and with draws made
// ON THE DEVICE (Philox keyed by the global sample index -- no host random numbers, no table upload):
//
//   level jitter      random_ups_and_downs, reference trainv2.py:120-124: one N(0, stddev^2) scalar added to channels [:4]
//   spatial           foa_intensity_vec_aug (transforms.py:78-114, C = 7) or acs_aug (:155-199, C = 17): a channel
//                     gather + sign per sample, applied consistently to the label coordinates y[B][T_y][4][n_classes]
//   time / frequency  mask (transforms.py:6-43) per `period`-frame chunk, the same bands as seld_mask draws
//
// out = keep(t, f) * sign[c] * (in[src[c]] + (src[c] < 4 ? offset : 0)); masked positions are value * 0 like the reference's
// `specs * mask` (so -0.0 / NaN survive).  The reference applies these as separate tf.data map stages (jitter, masks per
// sample; spatial per batch); every one of them is a per-(t, f) channel map, so they commute into this single read + write of
// the batch (275 MB for 256 x [300, 64, 7]).  The Philox streams are the ones transforms.py's host functions use
// (STREAM_IV_AUG 0x100, STREAM_ACS_AUG 0x101, STREAM_LEVEL_JITTER 0x102, masks: chunk index), so the fused launch equals
// the separate calls bit for bit.
#include <math.h>

#include "philox.cuh"
#include "plan.h"
#include "seld_common.cuh"

namespace seld {

constexpr uint32_t kStreamIvAug = 0x100u, kStreamAcsAug = 0x101u, kStreamLevelJitter = 0x102u;
constexpr int kMaxAugChan = 32;

// reference transforms.py:143-152 (arXiv:2101.02919, table 1): [[mic channel], [foa channel]] for the 8 swaps
__constant__ signed char c_channel_list[8][2][4] = {
    {{1, 3, 0, 2}, {0, -3, -2, 1}}, {{3, 1, 2, 0}, {0, -3, 2, -1}}, {{0, 1, 2, 3}, {0, 1, 2, 3}}, {{1, 0, 3, 2}, {0, -1, -2, 3}},
    {{2, 0, 3, 1}, {0, 3, -2, -1}}, {{0, 2, 1, 3}, {0, 3, 2, 1}}, {{3, 2, 1, 0}, {0, -1, 2, -3}}, {{2, 3, 0, 1}, {0, 1, -2, -3}}};
// reference transforms.py:122-139 (mic_gcc_perm): index of the GCC pair of two microphones
__constant__ signed char c_pair_decode[4][4] = {{0, 0, 1, 2}, {0, 0, 3, 4}, {1, 3, 0, 5}, {2, 4, 5, 0}};

struct SampleAug {            // per-sample tables, built by one thread per block
    int x_src[kMaxAugChan];   // source channel of output channel c
    float x_sgn[kMaxAugChan];
    int y_src[4];
    float y_sgn[4];
    float offset;             // level jitter on channels [:4]
    int draw_word;            // packed draws for draws_out: flips | swap << 3 (IV) or the swap index (ACS)
};

__device__ void build_sample_aug(SampleAug& s, int spatial, int C, float level_stddev, unsigned long long seed, unsigned long long gs) {
    const uint32_t k0 = uint32_t(seed), k1 = uint32_t(seed >> 32), s0 = uint32_t(gs), s1 = uint32_t(gs >> 32);
    for (int c = 0; c < C; ++c) { s.x_src[c] = c; s.x_sgn[c] = 1.f; }
    for (int j = 0; j < 4; ++j) { s.y_src[j] = j; s.y_sgn[j] = 1.f; }
    s.offset = 0.f;
    s.draw_word = 0;
    int perm[3] = {0, 1, 2}, feat_perm[3] = {0, 1, 2};
    float sgn_out[3] = {1.f, 1.f, 1.f};                       // sign of output intensity / coordinate axis j
    if (spatial == 1) {                                       // foa_intensity_vec_aug
        uint32_t w[4];
        philox4x32_10(s0, s1, kStreamIvAug, 0u, k0, k1, w);
        const int flip[3] = {int(w[0] & 1u), int(w[1] & 1u), int(w[2] & 1u)};      // tf.random.uniform([B, 3], 0, 2)
        const int swap = int(w[3] & 1u);                                          // tf.random.uniform([B, 1], maxval=2)
        perm[0] = 2 * swap; perm[1] = 1; perm[2] = 2 - 2 * swap;                   // :100-101
        int check = 0;
        for (int j = 0; j < 3; ++j) check += perm[j] != j;
        for (int j = 0; j < 3; ++j) feat_perm[j] = (perm[j] + check) % 3;          // :103-104
        for (int j = 0; j < 3; ++j) sgn_out[j] = flip[feat_perm[j]] ? -1.f : 1.f;  // flips happen BEFORE the gather (:96-97)
        s.draw_word = flip[0] | (flip[1] << 1) | (flip[2] << 2) | (swap << 3);
    } else if (spatial == 2) {                                // acs_aug
        const int idx = int(philox4x32_10_first(s0, s1, kStreamAcsAug, 0u, k0, k1) % 8u);
        int mic[4];
        for (int j = 0; j < 4; ++j) mic[j] = c_channel_list[idx][0][j];
        for (int j = 0; j < 3; ++j) {
            const int f = c_channel_list[idx][1][1 + j];
            sgn_out[j] = f < 0 ? -1.f : 1.f;                                       // sign applied AFTER the gather (:182)
            perm[j] = (f < 0 ? -f : f) - 1;                                        // :176
        }
        int check = 0;
        for (int j = 0; j < 3; ++j) check += perm[j] != j;
        for (int j = 0; j < 3; ++j) feat_perm[j] = (perm[j] + check) % 3;          // :179
        for (int j = 0; j < 4; ++j) s.x_src[7 + j] = 7 + mic[j];                   // :190
        int q = 0;
        for (int i = 0; i < 4; ++i)
            for (int j = i + 1; j < 4; ++j) s.x_src[11 + q++] = 11 + c_pair_decode[mic[i]][mic[j]];      // :188-189
        s.draw_word = idx;
    }
    if (spatial != 0) {
        for (int j = 0; j < 3; ++j) {
            s.x_src[1 + j] = 1 + perm[j];                     // FOA channels follow `perm`
            s.x_src[4 + j] = 4 + feat_perm[j];                // intensity vectors follow `feat_perm`, with signs
            s.x_sgn[4 + j] = sgn_out[j];
            s.y_src[1 + j] = 1 + feat_perm[j];                // label x, y, z
            s.y_sgn[1 + j] = sgn_out[j];
        }
    }
    if (level_stddev > 0.f) {                                 // Box-Muller in float64 on words 0, 1 of (sample, STREAM_LEVEL_JITTER)
        uint32_t w[4];
        philox4x32_10(s0, s1, kStreamLevelJitter, 0u, k0, k1, w);
        const double u1 = (double(w[0]) + 0.5) / 4294967296.0, u2 = (double(w[1]) + 0.5) / 4294967296.0;
        s.offset = float(double(level_stddev) * sqrt(-2.0 * log(u1)) * cos(2.0 * 3.14159265358979323846 * u2));
    }
}

struct AugArgs {
    const float* x_in;
    float* x_out;
    unsigned T, F, C;
    int f_shift;              // log2(F) when F is a power of two, else -1
    int spatial;
    float level_stddev;
    int period, n_chunks;
    int time_max, time_n, freq_max, freq_n;
    unsigned long long seed, sample_offset;
    int* draws_out;           // [B][2]: packed spatial draw, level offset bits
};

// grid = (work blocks, B)
template <int CT>            // compile-time channel count (0: runtime a.C)
__global__ void __launch_bounds__(256) augment_kernel(AugArgs a) {
    extern __shared__ unsigned char sm[];
    __shared__ SampleAug s_aug;
    const unsigned C = CT ? CT : a.C;
    unsigned char* keep_t = sm;                                         // [T]
    unsigned char* keep_f = sm + ((a.T + 15u) & ~15u);                  // [n_chunks][F]
    const unsigned long long b = blockIdx.y;
    const unsigned long long gs = a.sample_offset + b;
    for (unsigned i = threadIdx.x; i < a.T; i += blockDim.x) keep_t[i] = 1;
    for (unsigned i = threadIdx.x; i < a.n_chunks * a.F; i += blockDim.x) keep_f[i] = 1;
    if (threadIdx.x == 0) {
        build_sample_aug(s_aug, a.spatial, int(C), a.level_stddev, a.seed, gs);
        if (a.draws_out != nullptr && blockIdx.x == 0) {
            a.draws_out[2 * b] = s_aug.draw_word;
            a.draws_out[2 * b + 1] = __float_as_int(s_aug.offset);
        }
    }
    __syncthreads();
    // bands: the draws of mask_kernel (mask.cu) -- counter (sample, chunk, axis << 24 | mask << 1 | draw)
    const int n_masks = a.time_n + a.freq_n;
    for (int m = threadIdx.x; m < a.n_chunks * n_masks; m += blockDim.x) {
        const int chunk = m / n_masks, mm = m - chunk * n_masks;
        const bool is_time = mm < a.time_n;
        const int mi = is_time ? mm : mm - a.time_n;
        const int total = is_time ? a.period : int(a.F);
        int mx = is_time ? a.time_max : a.freq_max;
        if (mx <= 0) mx = total;
        const uint32_t k0 = uint32_t(a.seed), k1 = uint32_t(a.seed >> 32);
        const uint32_t c3 = (uint32_t(is_time ? 0 : 1) << 24) | (uint32_t(mi) << 1);
        const uint32_t u_size = philox4x32_10_first(uint32_t(gs), uint32_t(gs >> 32), uint32_t(chunk), c3, k0, k1);
        const uint32_t u_off = philox4x32_10_first(uint32_t(gs), uint32_t(gs >> 32), uint32_t(chunk), c3 | 1u, k0, k1);
        const int size = int(u_size % uint32_t(mx));
        const int off = int(u_off % uint32_t(total - size));
        unsigned char* dst = is_time ? keep_t + chunk * a.period + off : keep_f + chunk * a.F + off;
        for (int i = 0; i < size; ++i) dst[i] = 0;            // (bands of one chunk may overlap: plain stores of 0 commute)
    }
    __syncthreads();

    const unsigned n_pos = a.T * a.F;
    const float* src = a.x_in + b * (unsigned long long)n_pos * C;
    float* dst = a.x_out + b * (unsigned long long)n_pos * C;
    const float off = s_aug.offset;
    if constexpr (CT > 0 && CT <= 10) {
        // one thread = one (t, f) position = CT contiguous floats (a warp covers 32 * CT contiguous floats; CT loads in flight)
        for (unsigned o = blockIdx.x * blockDim.x + threadIdx.x; o < n_pos; o += gridDim.x * blockDim.x) {
            const unsigned t = o / a.F, f = o - t * a.F;
            const bool keep = keep_t[t] && keep_f[(t / a.period) * a.F + f];
            const float* p = src + (unsigned long long)o * CT;
            float* q = dst + (unsigned long long)o * CT;
            float r[CT];
#pragma unroll
            for (int c = 0; c < CT; ++c) {                    // the sources of a position share one or two cache lines
                const int sc = s_aug.x_src[c];
                r[c] = s_aug.x_sgn[c] * (p[sc] + (sc < 4 ? off : 0.f));
            }
#pragma unroll
            for (int c = 0; c < CT; ++c) q[c] = keep ? r[c] : r[c] * 0.0f;
        }
    } else {
        // wide rows (acs_aug: 17 channels): one thread = one element, consecutive lanes on consecutive floats, four independent
        // loads in flight per thread
        const unsigned n_el = n_pos * C, stride = gridDim.x * blockDim.x;
        constexpr int UN = 4;
        for (unsigned e0 = blockIdx.x * blockDim.x + threadIdx.x; e0 < n_el; e0 += UN * stride) {
            float r[UN];
            bool keep[UN];
#pragma unroll
            for (int k = 0; k < UN; ++k) {
                const unsigned e = e0 + k * stride;
                if (e < n_el) {
                    const unsigned o = e / C, c = e - o * C;
                    const unsigned t = a.f_shift >= 0 ? (o >> a.f_shift) : o / a.F, f = o - t * a.F;
                    const int sc = s_aug.x_src[c];
                    r[k] = s_aug.x_sgn[c] * (src[(unsigned long long)o * C + sc] + (sc < 4 ? off : 0.f));
                    keep[k] = keep_t[t] && keep_f[(t / a.period) * a.F + f];
                }
            }
#pragma unroll
            for (int k = 0; k < UN; ++k) {
                const unsigned e = e0 + k * stride;
                if (e < n_el) dst[e] = keep[k] ? r[k] : r[k] * 0.0f;
            }
        }
    }
}

// labels y[B][T_y][4][n_cls]: out[b, t, j, k] = sign[j] * in[b, t, src[j], k], same per-sample draws
__global__ void __launch_bounds__(256) augment_labels_kernel(const float* __restrict__ y_in, float* __restrict__ y_out, unsigned per_sample,
                                                             unsigned n_cls, int spatial, unsigned long long seed, unsigned long long sample_offset) {
    __shared__ SampleAug s_aug;
    const unsigned long long b = blockIdx.y;
    if (threadIdx.x == 0) build_sample_aug(s_aug, spatial, spatial == 2 ? 17 : 7, 0.f, seed, sample_offset + b);
    __syncthreads();
    const float* src = y_in + b * per_sample;
    float* dst = y_out + b * per_sample;
    const unsigned row = 4 * n_cls;
    for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < per_sample; e += gridDim.x * blockDim.x) {
        const unsigned o = e / row, r = e - o * row;
        const unsigned j = r / n_cls, k = r - j * n_cls;
        dst[e] = s_aug.y_sgn[j] * src[o * row + unsigned(s_aug.y_src[j]) * n_cls + k];
    }
}

}  // namespace seld

using namespace seld;

extern "C" int seld_augment_batch(const float* x_in_dev, float* x_out_dev, int64_t n_samples, int64_t t, int64_t f, int n_chan,
                                  const float* y_in_dev, float* y_out_dev, int64_t t_y, int n_classes, int spatial, float level_stddev,
                                  int period, int time_max, int time_n, int freq_max, int freq_n, uint64_t seed, uint64_t sample_offset,
                                  int32_t* draws_out_dev, void* stream) {
    if (!x_in_dev || !x_out_dev || n_samples < 0 || t < 0 || f < 1 || n_chan < 1) { set_error("bad argument"); return SELD_EINVAL; }
    if (x_in_dev == x_out_dev && spatial != 0) { set_error("the spatial augmentations are out of place"); return SELD_EINVAL; }
    if (n_chan > kMaxAugChan) { set_error("at most 32 channels"); return SELD_EUNSUPPORTED; }
    if (spatial < 0 || spatial > 2) { set_error("spatial must be 0 (none), 1 (foa_intensity_vec_aug) or 2 (acs_aug)"); return SELD_EINVAL; }
    if ((spatial == 1 && n_chan != 7) || (spatial == 2 && n_chan != 17)) {
        set_error("foa_intensity_vec_aug needs 7 channels, acs_aug 17");
        return SELD_EINVAL;
    }
    if (level_stddev > 0.f && n_chan < 4) { set_error("the level jitter needs the 4 log-mel channels"); return SELD_EINVAL; }
    if ((y_in_dev == nullptr) != (y_out_dev == nullptr)) { set_error("y_in and y_out go together"); return SELD_EINVAL; }
    if (y_in_dev != nullptr && (t_y < 0 || n_classes < 1)) { set_error("bad label shape"); return SELD_EINVAL; }
    if (time_n < 0 || freq_n < 0 || time_n + freq_n > 256) { set_error("at most 256 masks per chunk"); return SELD_EUNSUPPORTED; }
    if (period <= 0) period = (int)t;
    if (n_samples == 0 || t == 0) return SELD_OK;
    if (time_n + freq_n > 0) {
        if (t % period != 0) { set_error("(spec time length / period)' rest must be 0"); return SELD_EINVAL; }
        if (time_n > 0 && time_max > period) { set_error("time max_mask_size exceeds the period"); return SELD_EINVAL; }
        if (freq_n > 0 && freq_max > f) { set_error("freq max_mask_size exceeds the axis length"); return SELD_EINVAL; }
    }
    if (t * f >= (1ll << 31) || n_samples > 65535) { set_error("sample too large or more than 65535 samples"); return SELD_EUNSUPPORTED; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    AugArgs a;
    a.x_in = x_in_dev; a.x_out = x_out_dev;
    a.T = (unsigned)t; a.F = (unsigned)f; a.C = (unsigned)n_chan;
    a.f_shift = -1;
    for (int sh = 0; sh < 31; ++sh) if ((1ll << sh) == f) a.f_shift = sh;
    a.spatial = spatial; a.level_stddev = level_stddev;
    if (time_n + freq_n == 0) period = (int)t;               // no masks: one chunk
    a.period = period;
    a.n_chunks = int(t / period);
    a.time_max = time_max; a.time_n = time_n; a.freq_max = freq_max; a.freq_n = freq_n;
    a.seed = seed; a.sample_offset = sample_offset;
    a.draws_out = draws_out_dev;
    const size_t smem = ((a.T + 15u) & ~15u) + (size_t)a.n_chunks * a.F;
    if (smem > 160 * 1024) { set_error("augment: time axis too long for shared memory"); return SELD_EUNSUPPORTED; }
    if (t * f * n_chan >= (1ll << 31)) { set_error("sample too large (2^31 elements)"); return SELD_EUNSUPPORTED; }
    const long long n_items = (n_chan == 7 || n_chan == 10) ? t * f : (t * f * n_chan + 3) / 4;     // positions | 4-element groups
    long long bx = (n_items + 255) / 256;
    const long long cap = ((long long)device_sm_count() * 16 + n_samples - 1) / n_samples;
    if (bx > cap) bx = cap < 1 ? 1 : cap;
    dim3 grid((unsigned)bx, (unsigned)n_samples);
#define SELD_AUG_LAUNCH(CT)                                                                                                        \
    do {                                                                                                                           \
        static unsigned long long configured = 0;                                                                                  \
        if (smem > 48 * 1024 && first_use_on_device(&configured))                                                                  \
            SELD_CUDA_TRY(cudaFuncSetAttribute(augment_kernel<CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));      \
        augment_kernel<CT><<<grid, 256, smem, st>>>(a);                                                                            \
    } while (0)
    if (n_chan == 7) SELD_AUG_LAUNCH(7);
    else if (n_chan == 10) SELD_AUG_LAUNCH(10);
    else if (n_chan == 17) SELD_AUG_LAUNCH(17);
    else SELD_AUG_LAUNCH(0);
#undef SELD_AUG_LAUNCH
    SELD_CUDA_TRY(cudaGetLastError());
    seld::note_launch();
    if (y_in_dev != nullptr && t_y > 0) {
        const long long per_sample = t_y * 4 * n_classes;
        if (per_sample >= (1ll << 31)) { set_error("label sample too large"); return SELD_EUNSUPPORTED; }
        long long by = (per_sample + 255) / 256;
        if (by > 8) by = 8;
        augment_labels_kernel<<<dim3((unsigned)by, (unsigned)n_samples), 256, 0, st>>>(y_in_dev, y_out_dev, (unsigned)per_sample,
                                                                                      (unsigned)n_classes, spatial, seed, sample_offset);
        SELD_CUDA_TRY(cudaGetLastError());
        seld::note_launch();
    }
    return SELD_OK;
}
