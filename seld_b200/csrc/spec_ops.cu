// Stand-alone stages kept for API parity with the reference's public helpers (the hot path is the fused
// extractor in extract.cu):
//   seld_complex_spec   reference feature_extractor.py:153-173
//   seld_foa_iv         reference feature_extractor.py:176-193
//   seld_gcc            reference feature_extractor.py:196-214
#include <math.h>

#include "extract_core.cuh"
#include "plan.h"

namespace seld {

__host__ __device__ constexpr int sp_align16(int x) { return (x + 15) & ~15; }

// One warp per (channel pair, frame): packed FFT of two real channels, split, write rows [chan][t][F].
template <int R>
__global__ void __launch_bounds__(128) complex_spec_kernel(const float* __restrict__ wav, int n_chan, long long n_samples,
                                                           int hop, int t_raw, float scale, const float* __restrict__ window,
                                                           const float2* __restrict__ tw_t, float2* __restrict__ spec) {
    using G = Geo<R>;
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    float2* s_tw_t = reinterpret_cast<float2*>(smem);
    unsigned char* wp = smem + sp_align16(G::N * 8) + size_t(warp) * (sp_align16(G::E_ELEMS * 8) + sp_align16((G::N + 1) * 8));
    float2* E = reinterpret_cast<float2*>(wp);
    float2* S = reinterpret_cast<float2*>(wp + sp_align16(G::E_ELEMS * 8));
    for (int i = threadIdx.x; i < G::N; i += blockDim.x) s_tw_t[i] = tw_t[i];
    float wreg[R];
#pragma unroll
    for (int n2 = 0; n2 < R; ++n2) wreg[n2] = window[lane + 32 * n2];
    __syncthreads();
    const Tables tb{nullptr, s_tw_t, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    const int n_pairs = (n_chan + 1) / 2;
    const long long items = (long long)n_pairs * t_raw;
    ClipSrc src{wav, n_samples, 1, n_samples};
    for (long long it = (long long)blockIdx.x * nwarps + warp; it < items; it += (long long)gridDim.x * nwarps) {
        const int pair = int(it / t_raw), t = int(it % t_raw);
        const int ca = 2 * pair, cb = (2 * pair + 1 < n_chan) ? 2 * pair + 1 : ca;
        stage1_forward<R, LAYOUT_PLANAR_CL>(src, ca, cb, (long long)t * hop - G::N / 2, wreg, tb, E, lane);
        __syncwarp();
        stage2_forward<R>(E, S, lane);
        __syncwarp();
        float2* oa = spec + ((long long)ca * t_raw + t) * G::F;
        float2* ob = spec + ((long long)cb * t_raw + t) * G::F;
        for (int k = lane; k < G::F; k += 32) {
            const float2 z = S[k], zn = S[(G::N - k) & (G::N - 1)];     // Z = FFT(a + i b)
            const float h = 0.5f * scale;
            oa[k] = make_float2(h * (z.x + zn.x), h * (z.y - zn.y));
            if (cb != ca) ob[k] = make_float2(h * (z.y + zn.y), h * (zn.x - z.x));
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(256) foa_iv_kernel(const float2* __restrict__ spec, long long n, float eps,
                                                      float* __restrict__ iv) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float2 w = spec[i], y = spec[n + i], z = spec[2 * n + i], x = spec[3 * n + i];
        const float ix = w.x * x.x + w.y * x.y, iy = w.x * y.x + w.y * y.y, iz = w.x * z.x + w.y * z.y;
        const float nrm = fmaxf(sqrtf(ix * ix + iy * iy + iz * iz), eps);
        iv[i] = ix / nrm;
        iv[n + i] = iy / nrm;
        iv[2 * n + i] = iz / nrm;
    }
}

// One block per frame: phasors of one pair in shared memory, threads over lags (direct pruned inverse DFT).
__global__ void __launch_bounds__(256) gcc_kernel(const float2* __restrict__ spec, int n_chan, long long n_frames, int n_bins,
                                                  int n_lags, int first_lag, float* __restrict__ gcc) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int n = 2 * (n_bins - 1);
    float2* tw = reinterpret_cast<float2*>(smem);            // exp(+2 pi i j / n)
    float2* ph = tw + n;                                     // [n_bins]
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        float s, c;
        sincospif(2.0f * float(j) / float(n), &s, &c);
        tw[j] = make_float2(c, s);
    }
    const long long t = blockIdx.x;
    int pair = 0;
    for (int m = 0; m < n_chan; ++m) {
        for (int q = m + 1; q < n_chan; ++q, ++pair) {
            __syncthreads();
            const float2* xm = spec + ((long long)m * n_frames + t) * n_bins;
            const float2* xq = spec + ((long long)q * n_frames + t) * n_bins;
            for (int k = threadIdx.x; k < n_bins; k += blockDim.x) {
                const float2 a = xm[k], b = xq[k];
                const float2 r = make_float2(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x);   // conj(a) * b
                float2 u = unit_phasor(r);
                if (u.x == 0.f && u.y == 0.f) u = make_float2(1.f, 0.f);                       // angle(0) = 0
                ph[k] = u;
            }
            __syncthreads();
            for (int j = threadIdx.x; j < n_lags; j += blockDim.x) {
                const int lag = first_lag + j;
                const int lm = ((lag % n) + n) % n;
                float acc = ph[0].x + ((lm & 1) ? -ph[n_bins - 1].x : ph[n_bins - 1].x);
                float s = 0.f;
                for (int k = 1; k < n_bins - 1; ++k) {
                    const float2 w = tw[(k * lm) % n];
                    s += ph[k].x * w.x - ph[k].y * w.y;
                }
                gcc[((long long)pair * n_lags + j) * n_frames + t] = (acc + 2.0f * s) / float(n);
            }
        }
    }
}

template <int R>
static int launch_spec(const seld_plan* plan, const float* wav, int n_chan, long long n_samples, float scale, float* spec,
                       cudaStream_t st) {
    using G = Geo<R>;
    const int warps = (R == 64) ? 2 : 4;
    const int smem = sp_align16(G::N * 8) + warps * (sp_align16(G::E_ELEMS * 8) + sp_align16((G::N + 1) * 8));
    SELD_CUDA_TRY(cudaFuncSetAttribute(complex_spec_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int t_raw = int(1 + n_samples / plan->hop);
    const long long items = (long long)((n_chan + 1) / 2) * t_raw;
    long long blocks = (items + warps - 1) / warps;
    if (blocks > plan->num_sms * 4) blocks = plan->num_sms * 4;
    complex_spec_kernel<R><<<(int)blocks, warps * 32, smem, st>>>(wav, n_chan, n_samples, plan->hop, t_raw, scale, plan->window,
                                                                  reinterpret_cast<const float2*>(plan->tw_t),
                                                                  reinterpret_cast<float2*>(spec));
    SELD_CUDA_TRY(cudaGetLastError());
    seld::note_launch();
    return SELD_OK;
}

}  // namespace seld

using namespace seld;

extern "C" {

int seld_complex_spec(seld_plan_t plan, const float* wav_dev, int n_chan, int64_t n_samples, float scale, float* spec_dev,
                      void* stream) {
    if (!plan || !wav_dev || !spec_dev) { set_error("null argument"); return SELD_EINVAL; }
    if (n_chan < 1) { set_error("need at least one channel"); return SELD_EINVAL; }
    if (n_samples <= plan->n_fft / 2) {
        set_error("reflect padding needs n_fft/2 < number of samples (torch.stft raises here too)");
        return SELD_EINVAL;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (plan->n_fft) {
        case 256: return launch_spec<8>(plan, wav_dev, n_chan, n_samples, scale, spec_dev, st);
        case 512: return launch_spec<16>(plan, wav_dev, n_chan, n_samples, scale, spec_dev, st);
        case 1024: return launch_spec<32>(plan, wav_dev, n_chan, n_samples, scale, spec_dev, st);
        default: return launch_spec<64>(plan, wav_dev, n_chan, n_samples, scale, spec_dev, st);
    }
}

int seld_foa_iv(const float* spec_dev, int64_t n, float eps, float* iv_dev, void* stream) {
    if (!spec_dev || !iv_dev || n < 0) { set_error("bad argument"); return SELD_EINVAL; }
    if (n == 0) return SELD_OK;
    long long blocks = (n + 255) / 256;
    if (blocks > device_sm_count() * 16) blocks = device_sm_count() * 16;
    foa_iv_kernel<<<(int)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const float2*>(spec_dev), n, eps,
                                                                              iv_dev);
    SELD_CUDA_TRY(cudaGetLastError());
    seld::note_launch();
    return SELD_OK;
}

int seld_gcc(const float* spec_dev, int n_chan, int64_t n_frames, int n_bins, int n_lags, int first_lag, float* gcc_dev,
             void* stream) {
    if (!spec_dev || !gcc_dev) { set_error("null argument"); return SELD_EINVAL; }
    if (n_chan < 2 || n_bins < 2 || n_lags < 1 || n_frames < 0) { set_error("bad shape"); return SELD_EINVAL; }
    if (n_frames == 0) return SELD_OK;
    const int n = 2 * (n_bins - 1);
    const size_t smem = sizeof(float2) * (size_t(n) + n_bins);
    if (smem > 200 * 1024) { set_error("too many bins"); return SELD_EUNSUPPORTED; }
    SELD_CUDA_TRY(cudaFuncSetAttribute(gcc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gcc_kernel<<<(unsigned)n_frames, 256, smem, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const float2*>(spec_dev),
                                                                                     n_chan, n_frames, n_bins, n_lags,
                                                                                     first_lag, gcc_dev);
    SELD_CUDA_TRY(cudaGetLastError());
    seld::note_launch();
    return SELD_OK;
}

}  // extern "C"
