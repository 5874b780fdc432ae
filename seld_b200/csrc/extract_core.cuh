// Per-warp phases of the fused SELD feature extractor (one warp = one STFT frame of 4 channels).
//
// Reference behaviour replaced (file:line relative to the reference repo):
//   complex_spec            feature_extractor.py:153-173  (centred STFT, reflect pad, periodic Hann
//                                                          zero-padded centred to n_fft)
//   |X|^2 -> mel -> dB      feature_extractor.py:63-71    (top_db clamp is applied by finalize/stats)
//   foa_intensity_vectors   feature_extractor.py:176-193  (+ mel projection, :76)
//   gcc_features            feature_extractor.py:196-214  (phase transform, irfft, 64 centre lags)
//   cat + transpose         feature_extractor.py:84-87    ([t, mel, chan] rows)
//
// Algorithm (N = n_fft = 32*R, lane = threadIdx & 31):
//   stage 1  lane holds x[lane + 32*n2], n2 < R, of TWO real channels packed as re/im; R-point
//            DIF FFT in registers; twiddle W_N^(lane*k2); transposed through the warp's exchange buffer.
//   stage 2  lane k2 runs a 32-point FFT down its column -> Z[R*k1 + k2] written to the spectrum buffer.
//   bins     lane owns a contiguous run of bins k, splits Z[k], Z[N-k] into the two real channels'
//            spectra, forms power / intensity vectors (or per-channel unit phasors), and reduces them
//            into the mel accumulators (each bin feeds at most two adjacent filters).
//   gcc      three packed Hermitian inverse transforms (two pairs each) pruned to the n_mels centre lags.
//
// Every function is per-lane and only communicates through the shared-memory pointers it is given, so
// the same code runs lane-by-lane on the CPU in tests/emu (phase boundary == __syncwarp()).
#pragma once

#include "seld_common.cuh"

namespace seld {

enum { MODE_FOA = 0, MODE_MIC = 1 };
enum { LAYOUT_PLANAR_CL = 0, LAYOUT_INTERLEAVED_LC = 1 };

struct Tables {             // CTA-shared constant tables (shared memory on the device)
    const float* window;    // [N]   periodic Hann(win_length) zero-padded centred to N
    const float2* twiddle;  // [N]   exp(-2 pi i j / N)
    const int* seg;         // [F]   first mel filter fed by bin k (-1: none)
    const float* w0;        // [F]   weight into filter seg[k]
    const float* w1;        // [F]   weight into filter seg[k] + 1
};

struct ClipSrc {            // one clip's samples: channel c, sample i -> base[c * chan_stride + i * samp_stride]
    const float* base;
    long long chan_stride;
    long long samp_stride;
    long long n_samples;
};

template <int R>
struct Geo {
    static constexpr int N = 32 * R;
    static constexpr int F = N / 2 + 1;
    static constexpr int EP = R + 1;                 // padded row length of the exchange buffer
    static constexpr int E_ELEMS = 32 * EP;          // float2 elements
    static constexpr int COLS = (R + 31) / 32;       // stage-2 columns per lane
    static constexpr int BPT = (F + 31) / 32;        // bins per lane in the bin phase
    static constexpr int LOG2R = ilog2(R);
};

#if defined(__CUDA_ARCH__)
#define SELD_SMEM_ADD(ptr, v) atomicAdd((ptr), (v))
#else
#define SELD_SMEM_ADD(ptr, v) (*(ptr) += (v))
#endif

// ---------------------------------------------------------------- stage 1: load, window, R-point FFT, twiddle
template <int R>
SELD_HD void stage1_forward(const ClipSrc& src, int ch_a, int ch_b, long long frame_start, const Tables& tb,
                            float2* E, int lane) {
    using G = Geo<R>;
    float2 v[R];
    const float* xa = src.base + ch_a * src.chan_stride;
    const float* xb = src.base + ch_b * src.chan_stride;
    const long long L = src.n_samples;
#pragma unroll
    for (int n2 = 0; n2 < R; ++n2) {
        const int n = lane + 32 * n2;
        const float w = tb.window[n];
        long long i = frame_start + n;
        if (i < 0) i = -i;                       // reflect, no edge repeat (torch.stft center=True)
        if (i >= L) i = 2 * (L - 1) - i;
        float a = 0.f, b = 0.f;
        if (w != 0.f) {
            a = xa[i * src.samp_stride];
            b = xb[i * src.samp_stride];
        }
        v[n2] = make_float2(w * a, w * b);
    }
    fft_dif<R>(v);
#pragma unroll
    for (int p = 0; p < R; ++p) {
        const int k2 = bitrev(p, G::LOG2R);
        E[lane * G::EP + k2] = cmul(v[p], tb.twiddle[lane * k2]);
    }
}

// ---------------------------------------------------------------- stage 2: 32-point FFT per column
template <int R>
SELD_HD void stage2_forward(const float2* E, float2* S, int lane) {
    using G = Geo<R>;
#pragma unroll
    for (int c = 0; c < G::COLS; ++c) {
        const int k2 = lane + 32 * c;
        if (k2 < R) {
            float2 u[32];
#pragma unroll
            for (int n1 = 0; n1 < 32; ++n1) u[n1] = E[n1 * G::EP + k2];
            fft_dif<32>(u);
#pragma unroll
            for (int p = 0; p < 32; ++p) S[R * bitrev(p, 5) + k2] = u[p];
        }
    }
}

// split packed spectrum Z = FFT(a + i b) at bin k (zn = Z[N-k]) into A[k], B[k]
SELD_HD void unpack2(float2 z, float2 zn, float2& A, float2& B) {
    A = make_float2(0.5f * (z.x + zn.x), 0.5f * (z.y - zn.y));
    B = make_float2(0.5f * (z.y + zn.y), 0.5f * (zn.x - z.x));
}

SELD_HD float2 unit_phasor(float2 a) {   // a / |a|, (0,0) for a == 0; scaled so |a|^2 neither under- nor overflows
    float m = fmaxf(fabsf(a.x), fabsf(a.y));
    if (!(m > 0.f)) return make_float2(0.f, 0.f);
    float sc = (m < 1e-16f) ? 1.8446744e19f : ((m > 1e16f) ? 5.4210109e-20f : 1.f);
    float x = a.x * sc, y = a.y * sc;
#if defined(__CUDA_ARCH__)
    float r = rsqrtf(x * x + y * y);
#else
    float r = 1.0f / sqrtf(x * x + y * y);
#endif
    return make_float2(x * r, y * r);
}

// ---------------------------------------------------------------- bin phase
// NV = 7 (FOA: 4 powers + 3 normalised intensity components) or 4 (MIC: powers; unit phasors are
// written back in place of the packed spectra for the GCC phase).
template <int R, int MODE>
SELD_HD void bin_phase(float2* S0, float2* S1, const Tables& tb, float* acc, int n_mels, int n_out_ch,
                       float eps, int lane) {
    using G = Geo<R>;
    constexpr int N = G::N;
    constexpr int NV = (MODE == MODE_FOA) ? 7 : 4;
    const int kbeg = lane * G::BPT;
    const int kend = (kbeg + G::BPT < G::F) ? kbeg + G::BPT : G::F;
    int cur = -1;
    float a0[NV], a1[NV];
#pragma unroll
    for (int c = 0; c < NV; ++c) { a0[c] = 0.f; a1[c] = 0.f; }

    for (int k = kbeg; k < kend; ++k) {
        const int kn = (N - k) & (N - 1);
        float2 ch[4];
        unpack2(S0[k], S0[kn], ch[0], ch[1]);
        unpack2(S1[k], S1[kn], ch[2], ch[3]);
        float val[NV];
#pragma unroll
        for (int c = 0; c < 4; ++c) val[c] = ch[c].x * ch[c].x + ch[c].y * ch[c].y;
        if constexpr (MODE == MODE_FOA) {
            // W = ch0, Y = ch1, Z = ch2, X = ch3; I = Re(conj(W) * {X, Y, Z})
            float ix = ch[0].x * ch[3].x + ch[0].y * ch[3].y;
            float iy = ch[0].x * ch[1].x + ch[0].y * ch[1].y;
            float iz = ch[0].x * ch[2].x + ch[0].y * ch[2].y;
            float nrm = fmaxf(sqrtf(ix * ix + iy * iy + iz * iz), eps);
            val[4] = ix / nrm;
            val[5] = iy / nrm;
            val[6] = iz / nrm;
        } else {
            float2 u0 = unit_phasor(ch[0]), u1 = unit_phasor(ch[1]);
            float2 u2 = unit_phasor(ch[2]), u3 = unit_phasor(ch[3]);
            if (k == 0 || k == N / 2) {          // real bins: both channels of a pair share one slot
                S0[k] = make_float2(u0.x, u1.x);
                S1[k] = make_float2(u2.x, u3.x);
            } else {
                S0[k] = u0; S0[kn] = u1;
                S1[k] = u2; S1[kn] = u3;
            }
        }
        const int s = tb.seg[k];
        if (s < 0) continue;
        if (s != cur) {
            if (cur >= 0) {
#pragma unroll
                for (int c = 0; c < NV; ++c) {
                    SELD_SMEM_ADD(&acc[cur * n_out_ch + c], a0[c]);
                    if (cur + 1 < n_mels) SELD_SMEM_ADD(&acc[(cur + 1) * n_out_ch + c], a1[c]);
                }
            }
            cur = s;
#pragma unroll
            for (int c = 0; c < NV; ++c) { a0[c] = 0.f; a1[c] = 0.f; }
        }
        const float w0 = tb.w0[k], w1 = tb.w1[k];
#pragma unroll
        for (int c = 0; c < NV; ++c) { a0[c] += w0 * val[c]; a1[c] += w1 * val[c]; }
    }
    if (cur >= 0) {
#pragma unroll
        for (int c = 0; c < NV; ++c) {
            SELD_SMEM_ADD(&acc[cur * n_out_ch + c], a0[c]);
            if (cur + 1 < n_mels) SELD_SMEM_ADD(&acc[(cur + 1) * n_out_ch + c], a1[c]);
        }
    }
}

// ---------------------------------------------------------------- GCC-PHAT, packed inverse transform q
// Pair order (reference feature_extractor.py:207-208): 0:(0,1) 1:(0,2) 2:(0,3) 3:(1,2) 4:(1,3) 5:(2,3).
// Transform q carries pairs 2q (real part of the result) and 2q+1 (imaginary part).
SELD_HD float2 pair_phasor(float2 um, float2 un) {   // exp(i angle(conj(Xm) Xn)); angle(0) = 0 -> 1
    const bool zm = (um.x == 0.f && um.y == 0.f), zn = (un.x == 0.f && un.y == 0.f);
    if (zm || zn) return make_float2(1.f, 0.f);
    return make_float2(um.x * un.x + um.y * un.y, um.x * un.y - um.y * un.x);
}

template <int R, int Q>
SELD_HD void gcc_stage1(const float2* S0, const float2* S1, float2* E, int lane) {
    using G = Geo<R>;
    constexpr int N = G::N;
    constexpr int PM[6] = {0, 0, 0, 1, 1, 2};
    constexpr int PN[6] = {1, 2, 3, 2, 3, 3};
    float2 v[R];
#pragma unroll
    for (int n2 = 0; n2 < R; ++n2) {
        const int k = lane + 32 * n2;
        const bool upper = k > N / 2;
        const int kk = upper ? N - k : k;
        const int kn = (N - kk) & (N - 1);
        float2 u[4];
        const float2 s0 = S0[kk], s0n = S0[kn], s1 = S1[kk], s1n = S1[kn];
        if (kk == 0 || kk == N / 2) {
            u[0] = make_float2(s0.x, 0.f); u[1] = make_float2(s0.y, 0.f);
            u[2] = make_float2(s1.x, 0.f); u[3] = make_float2(s1.y, 0.f);
        } else {
            u[0] = s0; u[1] = s0n; u[2] = s1; u[3] = s1n;
        }
        float2 pa = pair_phasor(u[PM[2 * Q]], u[PN[2 * Q]]);
        float2 pb = pair_phasor(u[PM[2 * Q + 1]], u[PN[2 * Q + 1]]);
        if (upper) { pa.y = -pa.y; pb.y = -pb.y; }       // Hermitian extension P[N-k] = conj(P[k])
        // conj(Pa + i Pb): the forward FFT of the conjugate is the conjugate of the inverse FFT
        v[n2] = make_float2(pa.x - pb.y, -(pa.y + pb.x));
    }
    fft_dif<R>(v);
#pragma unroll
    for (int p = 0; p < R; ++p) E[lane * G::EP + bitrev(p, G::LOG2R)] = v[p];
}

template <int R, int Q>
SELD_HD void gcc_stage2(const float2* E, const Tables& tb, float* acc, int n_mels, int n_out_ch, int lane) {
    using G = Geo<R>;
    constexpr int N = G::N;
    const float inv_n = 1.0f / float(N);
#pragma unroll
    for (int c = 0; c < G::COLS; ++c) {
        const int r = lane + 32 * c;
        if (r < R) {
            float2 u[32];
#pragma unroll
            for (int b = 0; b < 32; ++b) u[b] = E[b * G::EP + r];
            // output j <-> lag j - n_mels/2 <-> n = lag mod N; this column owns n == r (mod R)
            const int half = n_mels / 2;
            int j = (r + half) % R;
            for (; j < n_mels; j += R) {
                const int n = (j - half) & (N - 1);
                float sx = 0.f, sy = 0.f;
#pragma unroll
                for (int b = 0; b < 32; ++b) {
                    const float2 t = tb.twiddle[(b * n) & (N - 1)];
                    sx += u[b].x * t.x - u[b].y * t.y;
                    sy += u[b].x * t.y + u[b].y * t.x;
                }
                acc[j * n_out_ch + 4 + 2 * Q] = sx * inv_n;          // Re conj(sum)
                acc[j * n_out_ch + 4 + 2 * Q + 1] = -sy * inv_n;     // Im conj(sum)
            }
        }
    }
}

// ---------------------------------------------------------------- finish one frame row
// acc[m * C + c]: channels < 4 are mel power -> 10 log10(max(., amin)); the rest pass through.
// Returns this lane's maximum dB value (-inf if it owns no log-mel element).
SELD_HD float finish_row(float* acc, int n_mels, int n_out_ch, float* out_row /* nullable */, int lane) {
    float mx = -INFINITY;
    const int n = n_mels * n_out_ch;
    for (int e = lane; e < n; e += 32) {
        const int c = e % n_out_ch;
        float v = acc[e];
        if (c < 4) {
            v = 10.0f * log10f(fmaxf(v, 1e-10f));
            mx = fmaxf(mx, v);
        }
        if (out_row) out_row[e] = v;
        acc[e] = 0.f;
    }
    return mx;
}

}  // namespace seld
