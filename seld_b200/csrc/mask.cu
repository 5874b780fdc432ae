// Fused time / frequency spectrogram masking, sm_100a.
//
// Replaces reference transforms.py:6-43 (mask: period-wise) and :46-75 (simple_mask: whole axis), both axes
// in one in-place pass over x[n][t][mid][f][c].  Only the masked bands are touched (x <- x * 0, so -0.0 and
// NaN survive exactly as in the reference's `specs * mask`); everything else is neither read nor written,
// which is what keeps the HBM traffic at the masked fraction instead of a full read + write.
//
// Draws per (sample, chunk) and mask: size in [0, max), then offset in [0, total - size)
// (transforms.py:21-22 / :60-61).  PHILOX_COUNTER mode is the product's own reproducible stream (the
// reference has none in graph mode); TF_EAGER_COMPAT reproduces TensorFlow-2's eager stream from host-made
// op seeds and is what the known answers of reference transforms_test.py:8-30 are checked with.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "philox.cuh"
#include "plan.h"
#include "seld_common.cuh"

namespace seld {

struct MaskArgs {
    void* x;
    long long n_samples, t, mid, f, c;
    int period;          // rows per chunk (== t when the whole time axis is one chunk)
    int n_chunks;
    int time_max, time_n, freq_max, freq_n;
    unsigned long long seed, sample_offset;
    int rng_mode;
    const long long* op_seed2;
    int* draws_out;
    int splits;
};

template <typename T> __device__ __forceinline__ T times_zero(T v);
template <> __device__ __forceinline__ float times_zero(float v) { return v * 0.0f; }
template <> __device__ __forceinline__ double times_zero(double v) { return v * 0.0; }
template <> __device__ __forceinline__ __half times_zero(__half v) { return __float2half(__half2float(v) * 0.0f); }
template <> __device__ __forceinline__ __nv_bfloat16 times_zero(__nv_bfloat16 v) { return __float2bfloat16(__bfloat162float(v) * 0.0f); }
template <> __device__ __forceinline__ int times_zero(int) { return 0; }
template <> __device__ __forceinline__ long long times_zero(long long) { return 0; }
template <> __device__ __forceinline__ short times_zero(short) { return 0; }
template <> __device__ __forceinline__ unsigned char times_zero(unsigned char) { return 0; }

// grid = (n_samples * n_chunks, splits); shared: keep_t[period] bytes + bands
template <typename T>
__global__ void __launch_bounds__(256) mask_kernel(MaskArgs a) {
    extern __shared__ unsigned char sm[];
    unsigned char* keep_t = sm;
    int* bands = reinterpret_cast<int*>(sm + ((a.period + 15) & ~15));   // [(time_n + freq_n)][2] = offset, size
    const long long sc = blockIdx.x;
    const long long sample = sc / a.n_chunks;
    const int chunk = int(sc % a.n_chunks);
    const int n_masks = a.time_n + a.freq_n;

    for (int i = threadIdx.x; i < a.period; i += blockDim.x) keep_t[i] = 1;
    if (threadIdx.x < n_masks) {
        const int m = threadIdx.x;
        const bool is_time = m < a.time_n;
        const int mi = is_time ? m : m - a.time_n;
        const int total = is_time ? a.period : int(a.f);
        int mx = is_time ? a.time_max : a.freq_max;
        if (mx <= 0) mx = total;
        uint32_t u_size, u_off;
        if (a.rng_mode == SELD_RNG_PHILOX_COUNTER) {
            const unsigned long long gs = a.sample_offset + (unsigned long long)sample;
            const uint32_t k0 = uint32_t(a.seed), k1 = uint32_t(a.seed >> 32);
            const uint32_t c3 = (uint32_t(is_time ? 0 : 1) << 24) | (uint32_t(mi) << 1);
            u_size = philox4x32_10_first(uint32_t(gs), uint32_t(gs >> 32), uint32_t(chunk), c3, k0, k1);
            u_off = philox4x32_10_first(uint32_t(gs), uint32_t(gs >> 32), uint32_t(chunk), c3 | 1u, k0, k1);
        } else {
            long long base = ((sample * a.n_chunks + chunk) * n_masks + m) * 2;
            if (a.rng_mode == SELD_RNG_TF_EAGER_TWO_PASS)
                base = sample * (long long)a.n_chunks * n_masks * 2 +
                       (is_time ? ((long long)chunk * a.time_n + mi) * 2 : ((long long)a.n_chunks * a.time_n + (long long)chunk * a.freq_n + mi) * 2);
            const unsigned long long s2a = (unsigned long long)a.op_seed2[base], s2b = (unsigned long long)a.op_seed2[base + 1];
            const uint32_t k0 = uint32_t(a.seed), k1 = uint32_t(a.seed >> 32);
            u_size = philox4x32_10_first(0u, 0u, uint32_t(s2a), uint32_t(s2a >> 32), k0, k1);
            u_off = philox4x32_10_first(0u, 0u, uint32_t(s2b), uint32_t(s2b >> 32), k0, k1);
        }
        const int size = int(u_size % uint32_t(mx));
        const int off = int(u_off % uint32_t(total - size));
        bands[2 * m] = off;
        bands[2 * m + 1] = size;
        if (a.draws_out != nullptr && blockIdx.y == 0) {
            a.draws_out[(sc * n_masks + m) * 2] = off;
            a.draws_out[(sc * n_masks + m) * 2 + 1] = size;
        }
    }
    __syncthreads();
    for (int m = 0; m < a.time_n; ++m) {
        const int off = bands[2 * m], size = bands[2 * m + 1];
        for (int i = threadIdx.x; i < size; i += blockDim.x) keep_t[off + i] = 0;
    }
    __syncthreads();

    T* x = reinterpret_cast<T*>(a.x) + (sample * a.t + (long long)chunk * a.period) * a.mid * a.f * a.c;
    const long long fc = a.f * a.c;
    const long long row = a.mid * fc;
    // one warp per row of the chunk; rows are dealt over (blockIdx.y, warp)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (long long tt = (long long)blockIdx.y * nwarps + warp; tt < a.period; tt += (long long)a.splits * nwarps) {
        T* xr = x + tt * row;
        if (!keep_t[tt]) {
            for (long long e = lane; e < row; e += 32) xr[e] = times_zero<T>(xr[e]);
        } else {
            for (int m = a.time_n; m < n_masks; ++m) {
                const long long off = (long long)bands[2 * m] * a.c, len = (long long)bands[2 * m + 1] * a.c;
                for (long long mi = 0; mi < a.mid; ++mi) {
                    T* pb = xr + mi * fc + off;
                    for (long long e = lane; e < len; e += 32) pb[e] = times_zero<T>(pb[e]);
                }
            }
        }
    }
}

template <typename T>
static int launch_mask(const MaskArgs& a, int num_sms, cudaStream_t st) {
    const size_t smem = ((a.period + 15) & ~15) + sizeof(int) * 2 * (a.time_n + a.freq_n);
    if (smem > 200 * 1024) { set_error("mask: period too large for shared memory"); return SELD_EUNSUPPORTED; }
    SELD_CUDA_TRY(cudaFuncSetAttribute(mask_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)(a.n_samples * a.n_chunks), (unsigned)a.splits);
    mask_kernel<T><<<grid, 256, smem, st>>>(a);
    SELD_CUDA_TRY(cudaGetLastError());
    seld::note_launch();
    return SELD_OK;
}

}  // namespace seld

using namespace seld;

extern "C" int seld_mask(void* x_dev, int dtype, int64_t n_samples, int64_t t, int64_t mid, int64_t f, int64_t c,
                         int period, int time_max, int time_n, int freq_max, int freq_n, uint64_t seed,
                         uint64_t sample_offset, int rng_mode, const int64_t* op_seed2_dev, int32_t* draws_out_dev,
                         void* stream) {
    if (!x_dev) { set_error("null argument"); return SELD_EINVAL; }
    if (n_samples < 0 || t < 0 || mid < 1 || f < 1 || c < 1 || time_n < 0 || freq_n < 0) { set_error("bad shape"); return SELD_EINVAL; }
    if (time_n + freq_n > 256) { set_error("at most 256 masks per chunk"); return SELD_EUNSUPPORTED; }
    if (period <= 0) period = (int)t;
    if (n_samples == 0 || t == 0 || time_n + freq_n == 0) return SELD_OK;
    if (t % period != 0) { set_error("(spec time length / period)' rest must be 0"); return SELD_EINVAL; }
    if (time_n > 0 && time_max > period) { set_error("time max_mask_size exceeds the period"); return SELD_EINVAL; }
    if (freq_n > 0 && freq_max > f) { set_error("freq max_mask_size exceeds the axis length"); return SELD_EINVAL; }
    if (rng_mode != SELD_RNG_TF_EAGER_COMPAT && rng_mode != SELD_RNG_TF_EAGER_TWO_PASS && rng_mode != SELD_RNG_PHILOX_COUNTER) {
        set_error("bad rng_mode");
        return SELD_EINVAL;
    }
    if (rng_mode != SELD_RNG_PHILOX_COUNTER && op_seed2_dev == nullptr) { set_error("the TF-eager modes need op_seed2"); return SELD_EINVAL; }
    const int num_sms = device_sm_count();
    MaskArgs a;
    a.x = x_dev;
    a.n_samples = n_samples; a.t = t; a.mid = mid; a.f = f; a.c = c;
    a.period = period;
    a.n_chunks = int(t / period);
    a.time_max = time_max; a.time_n = time_n; a.freq_max = freq_max; a.freq_n = freq_n;
    a.seed = seed; a.sample_offset = sample_offset;
    a.rng_mode = rng_mode;
    a.op_seed2 = reinterpret_cast<const long long*>(op_seed2_dev);
    a.draws_out = draws_out_dev;
    const long long chunks = n_samples * a.n_chunks;
    if (chunks > 0x7fffffffLL) { set_error("too many chunks"); return SELD_EUNSUPPORTED; }
    long long splits = (8LL * num_sms + chunks - 1) / chunks;
    if (splits > (period + 7) / 8) splits = (period + 7) / 8;
    if (splits < 1) splits = 1;
    if (splits > 65535) splits = 65535;
    a.splits = (int)splits;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (dtype) {
        case SELD_DTYPE_F32: return launch_mask<float>(a, num_sms, st);
        case SELD_DTYPE_F64: return launch_mask<double>(a, num_sms, st);
        case SELD_DTYPE_F16: return launch_mask<__half>(a, num_sms, st);
        case SELD_DTYPE_BF16: return launch_mask<__nv_bfloat16>(a, num_sms, st);
        case SELD_DTYPE_I32: return launch_mask<int>(a, num_sms, st);
        case SELD_DTYPE_I64: return launch_mask<long long>(a, num_sms, st);
        case SELD_DTYPE_I16: return launch_mask<short>(a, num_sms, st);
        case SELD_DTYPE_U8: return launch_mask<unsigned char>(a, num_sms, st);
        default: set_error("unsupported dtype"); return SELD_EUNSUPPORTED;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Per-sample channel gather + sign flip: the data movement of the reference's batch-level spatial augmentations
// (foa_intensity_vec_aug transforms.py:78-114, acs_aug :155-199), which only permute / negate channels of the features
// [B, T, F, C] and of the label coordinates [B, T, 4, n_classes].  With x viewed as [n][outer][C][inner]:
//     out[b, o, c, j] = (table[b, c] < 0 ? -1 : 1) * in[b, o, table[b, c] & 0xff, j]
// One thread per output element (coalesced stores; the C sources of a position share one or two cache lines), the
// sample on blockIdx.y so the per-sample table sits in shared memory and all index arithmetic is 32-bit.
namespace seld {
__global__ void __launch_bounds__(256) channel_remap_kernel(const float* __restrict__ in, float* __restrict__ out, unsigned per_sample,
                                                            unsigned C, unsigned inner, const int* __restrict__ table) {
    __shared__ int s_tab[32];
    const long long b = blockIdx.y;
    if (threadIdx.x < C) s_tab[threadIdx.x] = table[b * C + threadIdx.x];
    __syncthreads();
    const float* src = in + b * per_sample;
    float* dst = out + b * per_sample;
    const unsigned ci = C * inner;
    for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < per_sample; e += gridDim.x * blockDim.x) {
        const unsigned o = e / ci, r = e - o * ci;
        const unsigned c = r / inner, j = r - c * inner;
        const int t = s_tab[c];
        const float v = src[o * ci + unsigned(t & 0xff) * inner + j];
        dst[e] = (t < 0 ? -1.0f : 1.0f) * v;
    }
}
}  // namespace seld

extern "C" int seld_channel_remap(const float* in_dev, float* out_dev, int64_t n_samples, int64_t outer, int n_chan, int64_t inner,
                                  const int32_t* table_dev, void* stream) {
    if (!in_dev || !out_dev || !table_dev || n_samples < 0 || outer < 0 || inner < 1 || n_chan < 1) { set_error("bad argument"); return SELD_EINVAL; }
    if (in_dev == out_dev) { set_error("seld_channel_remap is out of place"); return SELD_EINVAL; }
    if (n_chan > 32) { set_error("at most 32 channels"); return SELD_EUNSUPPORTED; }
    const long long per_sample = outer * n_chan * inner;
    if (per_sample >= (1ll << 31) || n_samples > 65535) { set_error("sample too large (2^31 elements) or more than 65535 samples"); return SELD_EUNSUPPORTED; }
    if (per_sample == 0 || n_samples == 0) return SELD_OK;
    long long bx = (per_sample + 255) / 256;
    const long long cap = ((long long)device_sm_count() * 16 + n_samples - 1) / n_samples;
    if (bx > cap) bx = cap < 1 ? 1 : cap;
    dim3 grid((unsigned)bx, (unsigned)n_samples);
    channel_remap_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(in_dev, out_dev, (unsigned)per_sample, (unsigned)n_chan,
                                                                            (unsigned)inner, table_dev);
    SELD_CUDA_TRY(cudaGetLastError());
    seld::note_launch();
    return SELD_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Per-sample level jitter: out[b, p, c] = in[b, p, c] + (c < n_first ? offset[b] : 0) -- the data movement of the
// reference's `random_ups_and_downs` (trainv2.py:120-124: one N(0, 0.2^2) scalar added to the 4 log-mel channels of a
// sample, ahead of the masks).  Out of place it doubles as the copy the masking transform needs anyway; in == out works.
namespace seld {
__global__ void __launch_bounds__(256) channel_offset_kernel(const float* __restrict__ in, float* __restrict__ out, unsigned per_sample,
                                                             unsigned C, unsigned n_first, const float* __restrict__ offset) {
    const long long b = blockIdx.y;
    const float off = offset[b];
    const float* src = in + b * per_sample;
    float* dst = out + b * per_sample;
    for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < per_sample; e += gridDim.x * blockDim.x) {
        const unsigned c = e % C;
        dst[e] = src[e] + (c < n_first ? off : 0.f);
    }
}
}  // namespace seld

extern "C" int seld_channel_offset(const float* in_dev, float* out_dev, int64_t n_samples, int64_t positions, int n_chan, int n_first,
                                   const float* offset_dev, void* stream) {
    if (!in_dev || !out_dev || !offset_dev || n_samples < 0 || positions < 0 || n_chan < 1 || n_first < 0 || n_first > n_chan) {
        set_error("bad argument");
        return SELD_EINVAL;
    }
    const long long per_sample = positions * n_chan;
    if (per_sample >= (1ll << 31) || n_samples > 65535) { set_error("sample too large (2^31 elements) or more than 65535 samples"); return SELD_EUNSUPPORTED; }
    if (per_sample == 0 || n_samples == 0) return SELD_OK;
    long long bx = (per_sample + 255) / 256;
    const long long cap = ((long long)device_sm_count() * 16 + n_samples - 1) / n_samples;
    if (bx > cap) bx = cap < 1 ? 1 : cap;
    dim3 grid((unsigned)bx, (unsigned)n_samples);
    channel_offset_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(in_dev, out_dev, (unsigned)per_sample, (unsigned)n_chan,
                                                                             (unsigned)n_first, offset_dev);
    SELD_CUDA_TRY(cudaGetLastError());
    seld::note_launch();
    return SELD_OK;
}
