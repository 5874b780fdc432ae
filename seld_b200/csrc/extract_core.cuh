// Per-lane phases of the fused SELD feature extractor.  A TEAM of two warps (TL = 64 lanes) owns one STFT frame of 4
// channels: each warp transforms one packed channel pair, then all 64 lanes share the bin phase and the mel gather.
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
//            DIF FFT in registers (packed f32x2 butterflies: the kernel is issue-bound and FADD2/FFMA2
//            do two FP32 lanes per issue slot); twiddle W_N^(lane*k2) from a lane-contiguous table;
//            transposed through the warp's exchange buffer with 128-bit stores.
//   stage 2  lane k2 runs a 32-point FFT down its column -> Z[R*k1 + k2], written back IN PLACE over the exchange
//            buffer (linear index) once every lane holds its column in registers.
//   bins     team lane u (0..63) owns a contiguous run of bins k, splits Z[k], Z[N-k] into the two real channels'
//            spectra, forms power / intensity vectors (or per-channel unit phasors), and reduces them
//            into the mel accumulators (each bin feeds at most two adjacent filters).
//   gcc      three packed Hermitian inverse transforms (two pairs each) pruned to the n_mels centre lags.
//
// Every function is per-lane and only communicates through the shared-memory pointers it is given, so
// the same code runs lane-by-lane on the CPU in tests/emu (phase boundary == __syncwarp() or the team barrier).
#pragma once

#include "seld_common.cuh"
#include "mel_pieces.h"
#if defined(__CUDACC__)
#include <cuda_fp16.h>
#include <type_traits>
#endif

namespace seld {

enum { MODE_FOA = 0, MODE_MIC = 1 };
enum { LAYOUT_PLANAR_CL = 0, LAYOUT_INTERLEAVED_LC = 1, LAYOUT_PCM16_LC = 2 /* int16 [sample][4], WAV frame order */ };

struct Tables {             // CTA-shared constant tables (shared memory on the device)
    const float* window;    // [N]      periodic Hann(win_length) zero-padded centred to N
    const float2* tw_t;     // [R][32]  tw_t[k2*32 + lane] = exp(-2 pi i lane*k2 / N)   (lane-contiguous)
    const float2* tw_lin;   // [N]      exp(-2 pi i j / N)  (generic GCC lags only; may be null otherwise)
    // sparse mel bank in "piece" form (built by seld_plan_create): team lane u owns bins [u*BPT, (u+1)*BPT); a piece is
    // a maximal run of one lane's bins that feed the same pair of adjacent filters (seg, seg+1)
    const float2* w01;          // [TL*BPT] 0.25 * (weight into filter seg, weight into filter seg+1); 0 past bin F-1
    const unsigned long long* endmask;  // [TL] bit i set: bin u*BPT + i is the last bin of a piece
    const int* slot0;           // [TL]     record slot of lane u's first piece
    const int* slot1;           // [TL]     slot of its second piece (later pieces follow at +1); layouts: mel_pieces.h
    const int* pb;              // [n_mels + 2] compact layout: pieces with seg == m are [pb[m+1], pb[m+2])
    const int* ov;              // [64] segment-major layout: overflow slots of the 3rd / 4th piece of a segment
    // flush-free lane form of the bank (mel_pieces.h; null / 0 where it does not apply)
    const float* w4;            // [TL*BPT][4] per-bin weights into the lane's four filters (device: tensor memory instead)
    const int* lane_beg;        // [TL] first bin of lane u
    const int* gtab;            // [64][kLaneGatherMax] record words of filter m
    int gather_n0, gather_n1;   // table entries the filters of team warp 0 / 1 need
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
    static constexpr int EP = R + 2;                 // padded row of the exchange buffer: 128-bit stores conflict-free
    static constexpr int E_ELEMS = 32 * EP;          // float2 elements
    static constexpr int COLS = (R + 31) / 32;       // stage-2 columns per lane
    static constexpr int TL = 64;                    // lanes of a frame team (two warps)
    static constexpr int BPT = (F + TL - 1) / TL;    // bins per team lane in the bin phase
    static constexpr int LOG2R = ilog2(R);
};

#if defined(__CUDA_ARCH__)
#define SELD_SMEM_ADD(ptr, v) atomicAdd((ptr), (v))
#else
#define SELD_SMEM_ADD(ptr, v) (*(ptr) += (v))
#endif

// ---------------------------------------------------------------- packed FP32x2 arithmetic (FADD2 / FMUL2 / FFMA2)
SELD_HD float2 padd(float2 a, float2 b) {
#if defined(__CUDA_ARCH__)
    float2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(*reinterpret_cast<unsigned long long*>(&r))
        : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
    return r;
#else
    return make_float2(a.x + b.x, a.y + b.y);
#endif
}
SELD_HD float2 psub(float2 a, float2 b) {
#if defined(__CUDA_ARCH__)
    float2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(*reinterpret_cast<unsigned long long*>(&r))
        : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
    return r;
#else
    return make_float2(a.x - b.x, a.y - b.y);
#endif
}
SELD_HD float2 pmul(float2 a, float2 b) {
#if defined(__CUDA_ARCH__)
    float2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(*reinterpret_cast<unsigned long long*>(&r))
        : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
    return r;
#else
    return make_float2(a.x * b.x, a.y * b.y);
#endif
}
SELD_HD float2 pfma(float2 a, float2 b, float2 c) {
#if defined(__CUDA_ARCH__)
    float2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(*reinterpret_cast<unsigned long long*>(&r))
        : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
          "l"(*reinterpret_cast<unsigned long long*>(&c)));
    return r;
#else
    return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}
// d * (c + i s) in two packed instructions: d*(c,c) then swap(d)*(-s,s) + .
SELD_HD float2 pcmul(float2 d, float c, float s) {
    return pfma(make_float2(d.y, d.x), make_float2(-s, s), pmul(d, make_float2(c, c)));
}
// the same product for a RUN-TIME twiddle as four scalar instructions: the packed form needs the swapped (d.y, d.x) built with
// two moves, so it saves no issue slot (same FP32 pipe cycles, same roundings; measured 8.445 -> 8.405 ms per 600 FOA clips)
SELD_HD float2 cmul_rt(float2 d, float c, float s) {
    return make_float2(fmaf(d.y, -s, d.x * c), fmaf(d.x, s, d.y * c));
}
// ... and packed again with c and s as BROADCAST scalar operands and the swap / sign on d's operand modifiers (no pair to build):
// two issue slots instead of four, same FP32 pipe cycles, same roundings.  The result is an aligned register PAIR, so it wants
// 64-bit stores (a 128-bit store of two results costs four moves into an aligned quad).  Measured per 600 clips, packed + 64-bit
// stores at an odd row pitch against scalar + 128-bit stores: MIC 12.90 against 13.21 ms, FOA 8.21 against 8.02 -- each kernel
// keeps the form that is faster for it (stage1_store_tm's PACKED flag).
SELD_HD float2 cmul_rt_packed(float2 d, float c, float s) {
    return pfma(make_float2(-d.y, d.x), make_float2(s, s), pmul(d, make_float2(c, c)));
}

template <int J, int N>
SELD_HD float2 pmul_tw(float2 d) {   // d * W_N^J, compile-time twiddle
    if constexpr (J == 0) {
        return d;
    } else if constexpr (4 * J == N) {
        return make_float2(d.y, -d.x);
    } else {
        return pcmul(d, Tw<J, N>::re, Tw<J, N>::im);
    }
}

template <int N, int J>
struct PButterflies {
    static SELD_HD void run(float2* v) {
        const float2 a = v[J], b = v[J + N / 2];
        v[J] = padd(a, b);
        v[J + N / 2] = pmul_tw<J, N>(psub(a, b));
        if constexpr (J + 1 < N / 2) PButterflies<N, J + 1>::run(v);
    }
};

// Forward DFT of v[0..N) in registers, result bit-reversed: v[p] = X[bitrev(p)].
template <int N>
SELD_HD void pfft_dif(float2* v) {
    if constexpr (N >= 2) {
        PButterflies<N, 0>::run(v);
        pfft_dif<N / 2>(v);
        pfft_dif<N / 2>(v + N / 2);
    }
}

// ---------------------------------------------------------------- stage 1: load, window, R-point FFT, twiddle
// wreg[n2] = window[lane + 32*n2] is held in registers by the caller for the whole kernel.
// Interior frames (whole frame inside the clip, no reflection): one 64-bit load per tap when the two channels
// of the pair are adjacent in memory (interleaved layout, ch_b == ch_a + 1, ch_a even), else two 32-bit loads.
// (Streaming L1::no_allocate loads were measured and rejected: planar 11.8 -> 12.15 ms, interleaved 12.9 -> 15.1 ms per
// 600 clips -- neighbouring frames re-read the sectors the first one brought into L1.)
template <int R, int LAYOUT>
SELD_HD void stage1_load_raw(const ClipSrc& src, int ch_a, int ch_b, long long frame_start, float2* raw, int lane) {
    if constexpr (LAYOUT == LAYOUT_INTERLEAVED_LC) {
        const float2* p = reinterpret_cast<const float2*>(src.base + (frame_start + lane) * 4 + ch_a);
#pragma unroll
        for (int n2 = 0; n2 < R; ++n2) raw[n2] = p[64 * n2];
    } else {
        const float* pa = src.base + ch_a * src.chan_stride + frame_start + lane;       // planar: samp_stride == 1
        const float* pb = src.base + ch_b * src.chan_stride + frame_start + lane;
#pragma unroll
        for (int n2 = 0; n2 < R; ++n2) raw[n2] = make_float2(pa[32 * n2], pb[32 * n2]);
    }
}

// ---- 16-bit PCM input (the payload of a WAV file, 4 interleaved channels = 8 bytes per sample).  One 32-bit load per
// tap brings the two channels of this warp's packed FFT input.  int16 -> float without I2F: put the 16 bits
// (offset-binary) into the mantissa of 2^23 and subtract 2^23 + 2^15; the 1/32768 of torchaudio's decoder is folded
// into the window taps (an exact power-of-two scaling), so the result equals window * (s / 32768) bit for bit.
// One channel pair only (32-bit load per tap): the team's warp `pair` reads its half of every 8-byte sample.
template <int R>
SELD_HD void stage1_load_raw_pcm16_pair(const short* base, int pair, long long frame_start, float2* raw, int lane) {
    const float* p = reinterpret_cast<const float*>(base + (frame_start + lane) * 4 + 2 * pair);
#pragma unroll
    for (int n2 = 0; n2 < R; ++n2) raw[n2].x = p[64 * n2];
}

SELD_HD float2 pcm16_pair_to_float(float packed) {       // two int16 in one 32-bit word -> (lo, hi) as floats
#if defined(__CUDA_ARCH__)
    const unsigned w = __float_as_uint(packed);
    const unsigned lo = __byte_perm(w, 0x4B000000u, 0x7610) ^ 0x8000u;
    const unsigned hi = __byte_perm(w, 0x4B000000u, 0x7632) ^ 0x8000u;
    return padd(make_float2(__uint_as_float(lo), __uint_as_float(hi)), make_float2(-8421376.0f, -8421376.0f));
#else
    union { float f; short s[2]; } c; c.f = packed;
    return make_float2(float(c.s[0]), float(c.s[1]));
#endif
}

// WS: stride of the window taps (1: a per-lane register array; 32: the shared table offset by the lane)
template <int R, int WS = 1>
SELD_HD void apply_window_pcm16(const float2* raw, int pair, const float* wreg16, float2* v) {
#pragma unroll
    for (int n2 = 0; n2 < R; ++n2)
        v[n2] = pmul(pcm16_pair_to_float(pair ? raw[n2].y : raw[n2].x), make_float2(wreg16[n2 * WS], wreg16[n2 * WS]));
}

// edge frames of PCM16 input (reflection): plain conversions, same value as the interior path
template <int R>
SELD_HD void stage1_load_reflect_pcm16(const short* base, long long n_samples, int pair, long long frame_start,
                                       const float* wreg16, float2* v, int lane) {
#pragma unroll
    for (int n2 = 0; n2 < R; ++n2) {
        long long i = frame_start + lane + 32 * n2;
        if (i < 0) i = -i;
        if (i >= n_samples) i = 2 * (n_samples - 1) - i;
        float a = 0.f, b = 0.f;
        if (wreg16[n2] != 0.f) {
            a = float(base[i * 4 + 2 * pair]);
            b = float(base[i * 4 + 2 * pair + 1]);
        }
        v[n2] = make_float2(wreg16[n2] * a, wreg16[n2] * b);
    }
}

template <int R, int WS = 1>
SELD_HD void apply_window(const float2* raw, const float* wreg, float2* v) {
#pragma unroll
    for (int n2 = 0; n2 < R; ++n2) v[n2] = pmul(raw[n2], make_float2(wreg[n2 * WS], wreg[n2 * WS]));
}

template <int R, int LAYOUT>
SELD_HD void stage1_load_interior(const ClipSrc& src, int ch_a, int ch_b, long long frame_start, const float* wreg,
                                  float2* v, int lane) {
    float2 raw[R];
    stage1_load_raw<R, LAYOUT>(src, ch_a, ch_b, frame_start, raw, lane);
    apply_window<R>(raw, wreg, v);
}

// Edge frames: reflect without edge repeat (torch.stft center=True, pad_mode='reflect').
// zero_tail: samples past the end read as zero instead (tf.signal.stft(pad_end=True), reference data_loader.py:320).
template <int R>
SELD_HD void stage1_load_reflect(const ClipSrc& src, int ch_a, int ch_b, long long frame_start, const float* wreg,
                                 float2* v, int lane, bool zero_tail = false) {
    const float* xa = src.base + ch_a * src.chan_stride;
    const float* xb = src.base + ch_b * src.chan_stride;
    const long long L = src.n_samples;
#pragma unroll
    for (int n2 = 0; n2 < R; ++n2) {
        long long i = frame_start + lane + 32 * n2;
        if (i < 0) i = -i;
        const bool past = i >= L;
        if (past) i = zero_tail ? 0 : 2 * (L - 1) - i;
        float a = 0.f, b = 0.f;
        if (wreg[n2] != 0.f && !(past && zero_tail)) {
            a = xa[i * src.samp_stride];
            b = xb[i * src.samp_stride];
        }
        v[n2] = make_float2(wreg[n2] * a, wreg[n2] * b);
    }
}

template <int R>
SELD_HD void stage1_fft_store(float2* v, const Tables& tb, float2* E, int lane) {
    using G = Geo<R>;
    pfft_dif<R>(v);
    // twiddle W_N^(lane*k2), then two neighbouring k2 per 128-bit store
    float4* E4 = reinterpret_cast<float4*>(E + lane * G::EP);
#pragma unroll
    for (int j = 0; j < R / 2; ++j) {
        const int p0 = bitrev(2 * j, G::LOG2R), p1 = bitrev(2 * j + 1, G::LOG2R);
        const float2 t0 = tb.tw_t[(2 * j) * 32 + lane], t1 = tb.tw_t[(2 * j + 1) * 32 + lane];
        const float2 a = cmul_rt(v[p0], t0.x, t0.y), b = cmul_rt(v[p1], t1.x, t1.y);
        float4 q; q.x = a.x; q.y = a.y; q.z = b.x; q.w = b.y;
        E4[j] = q;
    }
}

// ---------------------------------------------------------------- per-lane constant tables in tensor memory
// The window taps, the stage-1 twiddles and the mel weights are PER-LANE constants: lane l only ever reads its own
// column of each table.  That is exactly the access tensor memory offers a warp (warp w reads TMEM lanes 32 (w % 4) ..
// +31, thread i <-> lane i), and tcgen05.ld does not go through the LSU / shared-memory data pipe this kernel is bound
// by.  Column map of a lane (n_fft = 1024): [0, 32) window taps w[l + 32 n2]; [32, 96) twiddles (re, im) of W^(l k2),
// k2 < 32; [96, 114) mel weights (w0, w1) of team lane u's 9 bins.  Quadrants 0/2 hold the tables of a team's first
// warp, 1/3 those of its second (warp parity == quadrant parity).
#if defined(__CUDACC__)
constexpr int TMEM_COL_WIN = 0, TMEM_COL_TW = 32, TMEM_COL_W01 = 96, TMEM_COL_KEEP = 128, TMEM_COLS = 512;
// [128, 384): per warp (four warps share a lane quadrant, 64 columns each) the samples two consecutive frames have in common

__device__ __forceinline__ void tmem_ld2(unsigned taddr, float& a, float& b) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];\n\ttcgen05.wait::ld.sync.aligned;"
                 : "=f"(a), "=f"(b) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld4(unsigned taddr, float& a, float& b, float& c, float& d) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n\ttcgen05.wait::ld.sync.aligned;"
                 : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(unsigned taddr, float* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
                 "tcgen05.wait::ld.sync.aligned;"
                 : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]), "=f"(r[8]),
                   "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st2(unsigned taddr, float a, float b) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" :: "r"(taddr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void tmem_st16(unsigned taddr, const float* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 :: "r"(taddr), "f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]), "f"(r[4]), "f"(r[5]), "f"(r[6]), "f"(r[7]), "f"(r[8]),
                    "f"(r[9]), "f"(r[10]), "f"(r[11]), "f"(r[12]), "f"(r[13]), "f"(r[14]), "f"(r[15]) : "memory");
}

// stage1_fft_store with the twiddles read from tensor memory (same values, same arithmetic as the shared-table version)
// PACKED: packed twiddle products and 64-bit stores at row pitch R + 1 float2 (odd: the 16 lanes of a half-warp hit 16 different
// bank pairs); else scalar products and 128-bit stores at pitch R + 2.  stage2_load_fft takes the same pitch.
template <int R, bool PACKED = false>
__device__ __forceinline__ void stage1_store_tm(const float2* v, unsigned taddr_tw, float2* E, int lane) {   // v: pfft_dif output
    using G = Geo<R>;
    static_assert(R % 8 == 0, "twiddles are fetched 8 at a time");
    constexpr int EPITCH = PACKED ? R + 1 : G::EP;
    float4* E4 = reinterpret_cast<float4*>(E + lane * G::EP);
#pragma unroll
    for (int g = 0; g < R / 8; ++g) {
        float t[16];                                   // (re, im) of k2 = 8g .. 8g + 7
        tmem_ld16(taddr_tw + 16 * g, t);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int j = 4 * g + jj;
            const int p0 = bitrev(2 * j, G::LOG2R), p1 = bitrev(2 * j + 1, G::LOG2R);
            if constexpr (PACKED) {
                E[lane * EPITCH + 2 * j] = cmul_rt_packed(v[p0], t[4 * jj], t[4 * jj + 1]);
                E[lane * EPITCH + 2 * j + 1] = cmul_rt_packed(v[p1], t[4 * jj + 2], t[4 * jj + 3]);
            } else {
                const float2 a = cmul_rt(v[p0], t[4 * jj], t[4 * jj + 1]), b = cmul_rt(v[p1], t[4 * jj + 2], t[4 * jj + 3]);
                float4 q; q.x = a.x; q.y = a.y; q.z = b.x; q.w = b.y;
                E4[j] = q;
            }
        }
    }
}
template <int R, bool PACKED = false>
__device__ __forceinline__ void stage1_fft_store_tm(float2* v, unsigned taddr_tw, float2* E, int lane) {
    pfft_dif<R>(v);
    stage1_store_tm<R, PACKED>(v, taddr_tw, E, lane);
}
#endif

template <int R, int LAYOUT>
SELD_HD void stage1_forward(const ClipSrc& src, int ch_a, int ch_b, long long frame_start, const float* wreg,
                            const Tables& tb, float2* E, int lane) {
    float2 v[R];
    if (frame_start >= 0 && frame_start + 32 * R <= src.n_samples)      // warp-uniform
        stage1_load_interior<R, LAYOUT>(src, ch_a, ch_b, frame_start, wreg, v, lane);
    else
        stage1_load_reflect<R>(src, ch_a, ch_b, frame_start, wreg, v, lane);
    stage1_fft_store<R>(v, tb, E, lane);
}

// ---------------------------------------------------------------- stage 2: 32-point FFT per column
// Two per-lane phases so the spectrum can overwrite the exchange buffer (S == E is allowed): every lane first pulls
// its column(s) into registers and transforms them; after a __syncwarp() the results go back at linear index k.
template <int R, int EPITCH = Geo<R>::EP>
SELD_HD void stage2_load_fft(const float2* E, float2* u, int lane) {      // u[COLS * 32]
    using G = Geo<R>;
#pragma unroll
    for (int c = 0; c < G::COLS; ++c) {
        const int k2 = lane + 32 * c;
        if (k2 < R) {
#pragma unroll
            for (int n1 = 0; n1 < 32; ++n1) u[32 * c + n1] = E[n1 * EPITCH + k2];
            pfft_dif<32>(u + 32 * c);
        }
    }
}
template <int R>
SELD_HD void stage2_store(const float2* u, float2* S, int lane) {
    using G = Geo<R>;
#pragma unroll
    for (int c = 0; c < G::COLS; ++c) {
        const int k2 = lane + 32 * c;
        if (k2 < R) {
#pragma unroll
            for (int p = 0; p < 32; ++p) S[R * bitrev(p, 5) + k2] = u[32 * c + p];
            if (k2 == 0) S[Geo<R>::N] = u[32 * c];      // S[N] = S[0] (inside the padded buffer): the bin phase reads S[N - k] unmasked
        }
    }
}
#if defined(__CUDACC__)
template <int R>
__device__ __forceinline__ void stage2_forward(const float2* E, float2* S, int lane) {
    float2 u[Geo<R>::COLS * 32];
    stage2_load_fft<R>(E, u, lane);
    __syncwarp();
    stage2_store<R>(u, S, lane);
}
#endif

SELD_HD float rsqrt_ftz(float s) {      // one MUFU.RSQ; subnormal inputs flush to 0 -> +inf (callers clamp)
#if defined(__CUDA_ARCH__)
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(s));
    return r;
#else
    return (s < 1.17549435e-38f) ? INFINITY : 1.0f / sqrtf(s);
#endif
}

SELD_HD float sqrt_ftz(float s) {       // one MUFU.SQRT
#if defined(__CUDA_ARCH__)
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(s));
    return r;
#else
    return sqrtf(s);
#endif
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
// written back in place of the packed spectra for the GCC phase).  With Z = FFT(a + i b):
// 2 A[k] = Z[k] + conj(Z[N-k]), 2 B[k] = -i (Z[k] - conj(Z[N-k])); the factor 2 is carried: powers come out
// 4x (the mel weights are pre-scaled by the exact constant 1/4) and the intensity vector, scale-free apart from
// eps, is produced 4x as well.  Each team lane walks its BPT contiguous bins in straight-line, branch-free code,
// accumulating (into filter seg, into filter seg+1) as one packed FFMA2 per channel; at the end of every piece the
// pair sums are stored (predicated) as one record P[piece][c] = (sum w0 val_c, sum w1 val_c) and the accumulators
// are cleared.  Records are PSTRIDE = 7 | 5 float2 apart: 14 | 10 words, so 16 neighbouring pieces hit 16 different
// bank pairs.  No atomics, fixed order => bit-reproducible.
template <int MODE>
struct PieceGeo {
    static constexpr int NV = (MODE == MODE_FOA) ? 7 : 4;
    static constexpr int PSTRIDE = (MODE == MODE_FOA) ? 7 : 5;      // float2 per piece record
};


// if (flag) rec[c] = acc[c] for c < NV (predicated stores, no branch); then acc[c] *= (flag ? 0 : 1)
template <int NV>
SELD_HD void piece_flush(float2* acc, float2* rec, unsigned flag) {
#if defined(__CUDA_ARCH__)
    const unsigned addr = static_cast<unsigned>(__cvta_generic_to_shared(rec));
    const unsigned long long* a = reinterpret_cast<const unsigned long long*>(acc);
    if constexpr (NV == 7) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %7, 0;\n\t"
            "@p st.shared.b64 [%8], %0;\n\t@p st.shared.b64 [%8+8], %1;\n\t@p st.shared.b64 [%8+16], %2;\n\t"
            "@p st.shared.b64 [%8+24], %3;\n\t@p st.shared.b64 [%8+32], %4;\n\t@p st.shared.b64 [%8+40], %5;\n\t"
            "@p st.shared.b64 [%8+48], %6;\n\t}"
            :: "l"(a[0]), "l"(a[1]), "l"(a[2]), "l"(a[3]), "l"(a[4]), "l"(a[5]), "l"(a[6]), "r"(flag), "r"(addr));
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %4, 0;\n\t"
            "@p st.shared.b64 [%5], %0;\n\t@p st.shared.b64 [%5+8], %1;\n\t@p st.shared.b64 [%5+16], %2;\n\t"
            "@p st.shared.b64 [%5+24], %3;\n\t}"
            :: "l"(a[0]), "l"(a[1]), "l"(a[2]), "l"(a[3]), "r"(flag), "r"(addr));
    }
    const float keep = __uint_as_float((flag ^ 1u) * 0x3f800000u);      // 1.0f or 0.0f
#pragma unroll
    for (int c = 0; c < NV; ++c) acc[c] = pmul(acc[c], make_float2(keep, keep));
#else
    if (flag) {
        for (int c = 0; c < NV; ++c) { rec[c] = acc[c]; acc[c] = make_float2(0.f, 0.f); }
    }
#endif
}

#if defined(__CUDACC__)
// piece_flush for NV = 4 with the record cursor as two shared-memory addresses: if (mask & bit) { store acc to [rec]; acc = 0;
// rec = nxt; nxt += one record } -- the stores, the clear factor and both cursor updates hang off ONE predicate.
__device__ __forceinline__ void piece_flush_cursor(float2* acc, unsigned& rec, unsigned& nxt, unsigned mask, unsigned bit) {
    const unsigned long long* a = reinterpret_cast<const unsigned long long*>(acc);
    float keep;
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b32 t;\n\tand.b32 t, %7, %8;\n\tsetp.ne.u32 p, t, 0;\n\t"
        "@p st.shared.b64 [%1], %3;\n\t@p st.shared.b64 [%1+8], %4;\n\t@p st.shared.b64 [%1+16], %5;\n\t@p st.shared.b64 [%1+24], %6;\n\t"
        "selp.f32 %0, 0f00000000, 0f3F800000, p;\n\tselp.u32 %1, %2, %1, p;\n\t@p add.u32 %2, %2, 40;\n\t}"
        : "=f"(keep), "+r"(rec), "+r"(nxt) : "l"(a[0]), "l"(a[1]), "l"(a[2]), "l"(a[3]), "r"(mask), "r"(bit) : "memory");
    static_assert(PieceGeo<MODE_MIC>::PSTRIDE * 8 == 40, "record pitch of the MIC pieces");
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[c] = pmul(acc[c], make_float2(keep, keep));
}
#endif

SELD_HD float2 pair_phasor(float2 um, float2 un) {   // exp(i angle(conj(Xm) Xn)); angle(0) = 0 -> 1
    const bool zm = (um.x == 0.f && um.y == 0.f), zn = (un.x == 0.f && un.y == 0.f);
    if (zm || zn) return make_float2(1.f, 0.f);
    return make_float2(um.x * un.x + um.y * un.y, um.x * un.y - um.y * un.x);
}

// conj(um) * un for unit phasors um, un (packed arithmetic); (1, 0) when either channel is zero: exp(i angle(0)) = 1
SELD_HD float2 unit_pair(float2 um, float2 un, bool zero) {
    // both packed instructions take one scalar of um as a broadcast operand (no register pair to build): the swap of un and
    // the sign of the second half ride on the operand modifiers of FMUL2 / FFMA2 (12 fewer MOVs per bin than
    // (mx, mx) * un + (my, -my) * swap(un); 14.11 -> 13.75 ms per 600 MIC clips)
    const float2 t = pmul(make_float2(um.y, um.y), make_float2(un.y, un.x));       // (my ny, my nx)
    const float2 r = pfma(make_float2(um.x, um.x), un, make_float2(t.x, -t.y));    // (mx nx + my ny, mx ny - my nx)
    return zero ? make_float2(1.f, 0.f) : r;
}

SELD_HD float pack_half2(float lo, float hi) {          // two floats -> one 32-bit word of two fp16 (round to nearest)
#if defined(__CUDA_ARCH__)
    const __half2 h = __floats2half2_rn(lo, hi);
    return __uint_as_float(*reinterpret_cast<const unsigned*>(&h));
#else
    (void)lo; (void)hi;
    return 0.f;                                          // the tensor-core GCC path is device-only
#endif
}

template <int R, int MODE, bool W_TMEM = false, bool MAG = false>
SELD_HD void bin_phase(float2* S0, float2* S1, const Tables& tb, float2* P, float eps, int u,      // u: team lane, 0..TL-1
                       unsigned taddr_w01 = 0) {                                                    // W_TMEM: mel weights from tensor memory
    using G = Geo<R>;
    constexpr int N = G::N;
    constexpr int NV = PieceGeo<MODE>::NV;
    constexpr int PSTRIDE = PieceGeo<MODE>::PSTRIDE;
    const int kbeg = u * G::BPT;
    const unsigned long long endmask = tb.endmask[u];
    int slot = tb.slot0[u], next = tb.slot1[u];
    float2 acc2[NV];
#pragma unroll
    for (int c = 0; c < NV; ++c) acc2[c] = make_float2(0.f, 0.f);
    const float inv_eps = 1.0f / eps;

    // fully unrolled: measured 12% faster than a x2-rolled loop despite the larger instruction footprint
#pragma unroll
    for (int i = 0; i < G::BPT; ++i) {
        // bins past F-1 carry zero weights; they read their own (in-range, finite) slots of the full spectrum rather than a
        // clamped index, which keeps the lane -> bank map of the last lanes conflict-free (TL * BPT <= N)
        static_assert(G::TL * G::BPT <= N, "bin ownership must stay inside the spectrum buffer");
        const bool valid = kbeg + i < G::F;
        const int k = kbeg + i;
        const int kn = (N - k) & (N - 1);
        const float2 z0 = S0[k], z0n = S0[kn], z1 = S1[k], z1n = S1[kn];
        float2 ch[4];                                  // twice the channel spectra
        ch[0] = make_float2(z0.x + z0n.x, z0.y - z0n.y);
        ch[1] = make_float2(z0.y + z0n.y, z0n.x - z0.x);
        ch[2] = make_float2(z1.x + z1n.x, z1.y - z1n.y);
        ch[3] = make_float2(z1.y + z1n.y, z1n.x - z1.x);
        float val[NV];
#pragma unroll
        for (int c = 0; c < 4; ++c) val[c] = fmaf(ch[c].x, ch[c].x, ch[c].y * ch[c].y);     // 4 |X_c|^2
        if constexpr (MAG) {
            // TF variant (reference data_loader.py:320-322): the mel bank is applied to |X|, not |X|^2; 4 |X| = sqrt(4 * 4 |X|^2)
            // so that the 1/4 folded into the weights still applies
#pragma unroll
            for (int c = 0; c < 4; ++c) val[c] = sqrt_ftz(4.0f * val[c]);
        }
        if constexpr (MODE == MODE_FOA) {
            // W = ch0, Y = ch1, Z = ch2, X = ch3; I = Re(conj(W) * {X, Y, Z})  (here 4 I)
            const float ix = fmaf(ch[0].x, ch[3].x, ch[0].y * ch[3].y);
            const float iy = fmaf(ch[0].x, ch[1].x, ch[0].y * ch[1].y);
            const float iz = fmaf(ch[0].x, ch[2].x, ch[0].y * ch[2].y);
            // 4 / max(|4 I|, 4 eps) = min(4 rsqrt(|4 I|^2), 1/eps); rsqrt(0) = inf -> 1/eps, as maximum(norm, eps)
            const float inv4 = fminf(4.0f * rsqrt_ftz(fmaf(ix, ix, fmaf(iy, iy, iz * iz))), inv_eps);
            val[4] = ix * inv4;
            val[5] = iy * inv4;
            val[6] = iz * inv4;
        } else {
            if (valid) {
                const float2 u0 = unit_phasor(ch[0]), u1 = unit_phasor(ch[1]);
                const float2 u2 = unit_phasor(ch[2]), u3 = unit_phasor(ch[3]);
                if (k == 0 || k == N / 2) {          // real bins: both channels of a pair share one slot
                    S0[k] = make_float2(u0.x, u1.x);
                    S1[k] = make_float2(u2.x, u3.x);
                } else {
                    S0[k] = u0; S0[kn] = u1;
                    S1[k] = u2; S1[kn] = u3;
                }
            }
        }
        float2 w;
#if defined(__CUDA_ARCH__)
        if constexpr (W_TMEM) tmem_ld2(taddr_w01 + 2 * i, w.x, w.y);
        else w = tb.w01[kbeg + i];
#else
        (void)taddr_w01;
        w = tb.w01[kbeg + i];
#endif
#pragma unroll
        for (int c = 0; c < NV; ++c) acc2[c] = pfma(make_float2(val[c], val[c]), w, acc2[c]);
        const unsigned flag = static_cast<unsigned>(endmask >> i) & 1u;
        piece_flush<NV>(acc2, P + slot * PSTRIDE, flag);
        slot = flag ? next : slot;
        next += int(flag);
    }
}

// ---------------------------------------------------------------- bin phase, flush-free lane form (FOA)
// Same per-bin arithmetic as bin_phase<MODE_FOA>; the mel projection differs: lane u's run of bins touches at most four
// consecutive filters (mel_pieces.h), which it accumulates as two packed pairs A = (F0, F1), B = (F2, F3) per channel with the
// per-bin weights (a0, a1, b0, b1) -- two FFMA2 per channel and bin, NO per-bin flush (the piece form spends 7 FMUL2 + 7
// predicated 64-bit stores + slot bookkeeping on every bin although 113 of 576 slots end a piece) -- and stores one record of
// 14 float2 at the end.  Fixed order, no atomics => bit-reproducible.
template <int R, bool W_TMEM = false, bool MAG = false>
SELD_HD void bin_phase_lanes(const float2* S0, const float2* S1, const Tables& tb, float2* P, float eps, int u, unsigned taddr_w4 = 0) {
    using G = Geo<R>;
    constexpr int N = G::N;
    const int kbeg = tb.lane_beg[u];
    float2 A[7], B[7];
#pragma unroll
    for (int c = 0; c < 7; ++c) { A[c] = make_float2(0.f, 0.f); B[c] = make_float2(0.f, 0.f); }
    const float inv_eps = 1.0f / eps;
#pragma unroll
    for (int i = 0; i < G::BPT; ++i) {
        const int k = kbeg + i;                        // (bins past the lane's run, or past F - 1, carry zero weights)
        const int kn = N - k;                          // S[N] holds S[0] (stage2_store): no wrap, compile-time offsets from N - kbeg
        const float2 z0 = S0[k], z0n = S0[kn], z1 = S1[k], z1n = S1[kn];
        // twice the channel spectra, the odd channels with (re, im) swapped: one packed FFMA2 each instead of two scalar adds
        // (same roundings; measured 8.73 -> 8.45 ms per 600 clips)
        float2 ch[4];
        ch[0] = pfma(z0n, make_float2(1.f, -1.f), z0);         // (z0.x + z0n.x, z0.y - z0n.y)
        ch[1] = pfma(z0, make_float2(-1.f, 1.f), z0n);         // (z0n.x - z0.x, z0.y + z0n.y) = (im, re)
        ch[2] = pfma(z1n, make_float2(1.f, -1.f), z1);
        ch[3] = pfma(z1, make_float2(-1.f, 1.f), z1n);
        float val[7];
#pragma unroll
        for (int c = 0; c < 4; ++c)                            // 4 |X_c|^2
            val[c] = (c & 1) ? fmaf(ch[c].y, ch[c].y, ch[c].x * ch[c].x) : fmaf(ch[c].x, ch[c].x, ch[c].y * ch[c].y);
        if constexpr (MAG) {
#pragma unroll
            for (int c = 0; c < 4; ++c) val[c] = sqrt_ftz(4.0f * val[c]);                   // TF variant: 4 |X_c|
        }
        const float ix = fmaf(ch[0].x, ch[3].y, ch[0].y * ch[3].x);
        const float iy = fmaf(ch[0].x, ch[1].y, ch[0].y * ch[1].x);
        const float iz = fmaf(ch[0].x, ch[2].x, ch[0].y * ch[2].y);
        const float inv4 = fminf(4.0f * rsqrt_ftz(fmaf(ix, ix, fmaf(iy, iy, iz * iz))), inv_eps);
        val[4] = ix * inv4;
        val[5] = iy * inv4;
        val[6] = iz * inv4;
        float2 wa, wb;
#if defined(__CUDA_ARCH__)
        if constexpr (W_TMEM) {
            tmem_ld4(taddr_w4 + 4 * i, wa.x, wa.y, wb.x, wb.y);
        } else {
            const float* w = tb.w4 + (size_t(u) * G::BPT + i) * 4;
            wa = make_float2(w[0], w[1]); wb = make_float2(w[2], w[3]);
        }
#else
        (void)taddr_w4;
        const float* w = tb.w4 + (size_t(u) * G::BPT + i) * 4;
        wa = make_float2(w[0], w[1]); wb = make_float2(w[2], w[3]);
#endif
#pragma unroll
        for (int c = 0; c < 7; ++c) {
            A[c] = pfma(make_float2(val[c], val[c]), wa, A[c]);
            B[c] = pfma(make_float2(val[c], val[c]), wb, B[c]);
        }
    }
    float2* rec = P + u * (kLaneRecWords / 2);
#pragma unroll
    for (int c = 0; c < 7; ++c) { rec[2 * c] = A[c]; rec[2 * c + 1] = B[c]; }
}

SELD_HD float fast_db(float x) {            // 10 log10(x), x > 0
#if defined(__CUDA_ARCH__)
    return 3.0102999566398120f * __log2f(x);     // MUFU.LG2: <= 2 ulp of log2 => <= 2.3e-5 dB at -100 dB, ~6e-6 dB typical
#else
    return 3.0102999566398120f * log2f(x);
#endif
}

// maximum that PROPAGATES NaN (torch.max / torch.clamp semantics; fmaxf would drop it): one FMNMX.NAN on the device
SELD_HD float max_nan(float a, float b) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
#else
    return (a != a || b != b) ? NAN : fmaxf(a, b);
#endif
}
// 10 log10(max(power, 1e-10)); a NaN power stays NaN (reference :65-71 through amplitude_to_DB's clamp)
SELD_HD float power_to_db(float pw) { return fast_db(max_nan(pw, 1e-10f)); }
// TF variant: tfio dbscale = 10 log10(x^2) with NO floor (reference data_loader.py:323): log(0) = -inf survives until the
// top_db clamp against the clip maximum
SELD_HD float magnitude_to_db(float m) { return 2.0f * fast_db(m); }

// ---------------------------------------------------------------- gather: pieces -> mel rows
// Team lane u owns filters m = u, u + TL, ...: mel[m][c] = sum_{pieces of seg m} P.x + sum_{pieces of seg m-1} P.y, in
// piece order.  Log-mel channels get 10 log10(max(., 1e-10)) here; returns the lane's maximum dB.
template <int MODE, bool MAG = false>
SELD_HD float gather_phase(const float2* P, const Tables& tb, float* acc, int n_mels, int u) {
    constexpr int NV = PieceGeo<MODE>::NV;
    constexpr int PSTRIDE = PieceGeo<MODE>::PSTRIDE;
    constexpr int C = (MODE == MODE_FOA) ? 7 : 10;
    float mx = -INFINITY;
    for (int m = u; m < n_mels; m += 64) {
        const int p0 = tb.pb[m], p1 = tb.pb[m + 1], p2 = tb.pb[m + 2];
        float sum[NV];
#pragma unroll
        for (int c = 0; c < NV; ++c) sum[c] = 0.f;
#pragma unroll 1
        for (int p = p0; p < p1; ++p) {                // falling slopes of the segment below
#pragma unroll
            for (int c = 0; c < NV; ++c) sum[c] += P[p * PSTRIDE + c].y;
        }
#pragma unroll 1
        for (int p = p1; p < p2; ++p) {                // rising slopes of this filter's own segment
#pragma unroll
            for (int c = 0; c < NV; ++c) sum[c] += P[p * PSTRIDE + c].x;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float v = MAG ? magnitude_to_db(sum[c]) : power_to_db(sum[c]);
            mx = max_nan(mx, v);
            acc[m * C + c] = v;
        }
#pragma unroll
        for (int c = 4; c < NV; ++c) acc[m * C + c] = sum[c];
    }
    return mx;
}

// Fast gather over the segment-major record layout (mel_pieces.h): team lane u totals slots u, u + 65 and the two overflow
// slots ov[u] names (the zero slot for all but the widest filters) with packed adds -- consecutive lanes read consecutive
// records (no bank conflicts) and absent pieces are slots that were zeroed once and never written.  mel[u] = A[u].x +
// A[u-1].y takes one shuffle per channel from the lane below; lane 0 of the team's second warp re-totals segment 31
// itself (same additions in the same order, so the CPU emulation, which always re-totals, is bit-identical).
template <int MODE>
SELD_HD void seg_total(const float2* P, const Tables& tb, int seg, float2* A) {
    constexpr int NV = PieceGeo<MODE>::NV;
    constexpr int PSTRIDE = PieceGeo<MODE>::PSTRIDE;
    const float2* rec = P + seg * PSTRIDE;
    const int ov = tb.ov[seg];
    const float2* r2 = P + (ov & 0xffff) * PSTRIDE;
    const float2* r3 = P + (ov >> 16) * PSTRIDE;
#pragma unroll
    for (int c = 0; c < NV; ++c) A[c] = padd(padd(padd(rec[c], rec[kSegMajorPitch * PSTRIDE + c]), r2[c]), r3[c]);
}

template <int MODE, bool MAG = false>
SELD_HD float gather_lanes(const float2* P, const Tables& tb, float* acc, int n_mels, int u) {
    constexpr int NV = PieceGeo<MODE>::NV;
    constexpr int C = (MODE == MODE_FOA) ? 7 : 10;
    float2 A[NV];
    seg_total<MODE>(P, tb, u, A);
    float below[NV];
#if defined(__CUDA_ARCH__)
#pragma unroll
    for (int c = 0; c < NV; ++c) below[c] = __shfl_up_sync(0xffffffffu, A[c].y, 1);
    if (u == 0) {
#pragma unroll
        for (int c = 0; c < NV; ++c) below[c] = 0.f;
    }
    if (u == 32) {                                         // only the team's second warp takes this path
        float2 B[NV];
        seg_total<MODE>(P, tb, 31, B);
#pragma unroll
        for (int c = 0; c < NV; ++c) below[c] = B[c].y;
    }
#else
    for (int c = 0; c < NV; ++c) below[c] = 0.f;
    if (u > 0) {
        float2 B[NV];
        seg_total<MODE>(P, tb, u - 1, B);
        for (int c = 0; c < NV; ++c) below[c] = B[c].y;
    }
#endif
    float mx = -INFINITY;
    if (u < n_mels) {
#pragma unroll
        for (int c = 0; c < NV; ++c) {
            float v = A[c].x + below[c];
            if (c < 4) {
                v = MAG ? magnitude_to_db(v) : power_to_db(v);
                mx = max_nan(mx, v);
            }
            acc[u * C + c] = v;
        }
    }
    return mx;
}

// Gather of the lane form: filter m = team lane u adds the (at most gather_n) records gtab[m] names, in table order (absent
// entries point at the never-written zero record).
template <bool MAG = false>
SELD_HD float gather_records(const float2* P, const Tables& tb, float* acc, int n_mels, int u) {
    constexpr int C = 7;
    const float* W = reinterpret_cast<const float*>(P);
    const int* gt = tb.gtab + u * kLaneGatherMax;
    const int n = (u < 32) ? tb.gather_n0 : tb.gather_n1;            // warp-uniform
    float sum[C];
#pragma unroll
    for (int c = 0; c < C; ++c) sum[c] = 0.f;
    // (Two table entries per step with the entries preloaded by two 128-bit loads -- 14 independent record loads in flight -- was
    //  measured: 8.09 ms per 600 clips against 8.04 for this loop; the other warps already cover the chain's latency.)
    for (int t = 0; t < n; ++t) {
        const float* r = W + gt[t];
#pragma unroll
        for (int c = 0; c < C; ++c) sum[c] += r[4 * c];
    }
    float mx = -INFINITY;
    if (u < n_mels) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float v = sum[c];
            if (c < 4) {
                v = MAG ? magnitude_to_db(v) : power_to_db(v);
                mx = max_nan(mx, v);
            }
            acc[u * C + c] = v;
        }
    }
    return mx;
}

// ---------------------------------------------------------------- fused tensor-core GCC (MIC, n_fft 1024, 64 lags)
// reference feature_extractor.py:209-211 keeps 64 of the 1024 irfft outputs, so per (frame, pair)
//     cc[lag] = sum_K basis[lag][K] * P[K],   K = 1024 = (Re P[0], Re P[512], Re P[1], Im P[1], ..., Im P[511])
// is a [64 x 1024] x [1024 x 6] contraction per frame.  It runs on the tensor cores INSIDE the extractor:
//   * A operand = the basis (64 lags x 1024, fp16), resident in TENSOR MEMORY for the whole kernel (tcgen05.mma with the A
//     operand in TMEM): M = 64 occupies lanes 0..15 of each 32-lane subpartition, so two "atoms" interleave -- lanes 0..15
//     hold K in [0, 512), lanes 16..31 hold K in [512, 1024) -- and the 128 KB basis takes 256 columns (probed on B200:
//     tools/microbench/probe_tmem_ts.cu; A and D of one MMA must sit on the same datapath lanes).
//   * B operand = the six pair-phasor rows of the frames of a team PAIR (N = 16: two 8-row groups with two unused rows each,
//     one per team), written by the bin phase IN PLACE over
//     the team's spectrum in the no-swizzle K-major core-matrix layout: bin k of row r is the 32-bit word (re, im fp16) at
//         (k >> 2) * 144 + r * 16 + (k & 3) * 4
//     (16-byte K units of 8 rows, padded from 128 to 144 bytes: bank = (k + 4 r) mod 32, so both the stage-2 stores --
//     lane = k mod 32 -- and the bin phase -- lane u owns bins 9u .. 9u + 8 -- are conflict-free).  Before the bin phase the
//     same words hold the packed spectra: rows 0..3 = Re/Im Z0[k], Re/Im Z0[N-k], rows 4..7 the same of Z1.
//   * D = 16 fp32 columns per team pair in TMEM (lower atom + upper atom, summed with one shuffle in the epilogue).
// 64 MMAs (M64 N16 K16) per pair of frames; nothing of this touches HBM.
#if defined(__CUDACC__)
constexpr int GT_UNIT = 144;                    // byte pitch of a 16-byte K unit (8 rows)
constexpr int GT_CHUNK = 8 * GT_UNIT;           // 32 bins
constexpr int GT_BYTES = 128 * GT_UNIT;         // 512 bins: 18 432 bytes (the two exchange buffers alias its first 17 408)
constexpr int GT_NYQ_BYTES = 128;               // Nyquist "column": rows 0, 1 = Z0[512], rows 4, 5 = Z1[512]
constexpr int TMEM_COL_BASIS = 128, TMEM_COL_D = 384;         // D of team pair p: columns TMEM_COL_D + 16 p .. + 15 (even team's 8, odd team's 8)
__device__ __forceinline__ int gcc_dcol(int pair) { return TMEM_COL_D + 16 * pair; }

__device__ __forceinline__ bool elect_one() {   // one lane of a converged warp (lets ptxas issue UTCHMMA without a per-thread loop)
    unsigned pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// stage-2 results of warp h (u[p] = Z_h[32 bitrev(p) + lane]) -> rows 4h .. 4h + 3 of the tile; Z_h[512] -> the Nyquist column
__device__ __forceinline__ void stage2_store_tile(const float2* u, unsigned char* tile, unsigned char* nyq, int lane, int h) {
    const int pm = (32 - lane) & 31;                       // k = 32 k1 + lane > 512 mirrors to bin N - k: position 32 - lane
    unsigned char* bd = tile + (lane >> 2) * GT_UNIT + (lane & 3) * 4 + (4 * h) * 16;
    unsigned char* bm = tile + (pm >> 2) * GT_UNIT + (pm & 3) * 4 + (4 * h + 2) * 16 + (lane == 0 ? GT_CHUNK : 0);
    unsigned char* b16 = (lane == 0) ? nyq + (4 * h) * 16 : bm + 15 * GT_CHUNK;
#pragma unroll
    for (int p = 0; p < 32; ++p) {
        const int k1 = bitrev(p, 5);
        unsigned char* d = (k1 < 16) ? bd + k1 * GT_CHUNK : (k1 == 16 ? b16 : bm + (31 - k1) * GT_CHUNK);
        if (k1 == 0 && lane == 0) d = nyq + (4 * h) * 16 + 8;      // Z_h[0] -> third word of the Nyquist column's units (bin_phase_gcc_fused)
        *reinterpret_cast<float*>(d) = u[p].x;
        *reinterpret_cast<float*>(d + 16) = u[p].y;
    }
}

// six pair phasors conj(u_m) u_n of the per-channel unit phasors u_c = X_c / |X_c| (one MUFU.RSQ per channel on the power the
// log-mel block needs anyway); a pair with a zero channel is 1, as exp(i angle(0)) = 1.  Pair order: reference :207-208.
// A zero (underflowing) bin in a live frame is rare: the fix-up selects sit behind a warp vote.
#ifndef SELD_GCC_VOTE_ZERO
#define SELD_GCC_VOTE_ZERO 1
#endif
template <bool VOTE>      // VOTE: called by the whole (converged) warp
__device__ __forceinline__ void pair_phasors(const float2* ch, const float* val, float2* p) {
    float2 uc[4];
    bool zc[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        zc[c] = !(val[c] >= 1.17549435e-38f);
        const float inv = rsqrt_ftz(val[c]);
        uc[c] = pmul(ch[c], make_float2(inv, inv));
    }
    constexpr int PM[6] = {0, 0, 0, 1, 1, 2}, PN[6] = {1, 2, 3, 2, 3, 3};
    if constexpr (VOTE && SELD_GCC_VOTE_ZERO) {
#pragma unroll
        for (int q = 0; q < 6; ++q) p[q] = unit_pair(uc[PM[q]], uc[PN[q]], false);
        if (__any_sync(0xffffffffu, zc[0] || zc[1] || zc[2] || zc[3])) {
#pragma unroll
            for (int q = 0; q < 6; ++q) if (zc[PM[q]] || zc[PN[q]]) p[q] = make_float2(1.f, 0.f);
        }
    } else {
#pragma unroll
        for (int q = 0; q < 6; ++q) p[q] = unit_pair(uc[PM[q]], uc[PN[q]], zc[PM[q]] || zc[PN[q]]);
    }
}

// ... and without any fix-up: a pair with a zero (underflowing) channel comes out non-finite -- rsqrt.ftz gives +inf -- and the
// caller patches the packed words after its bin loop, behind ONE vote on the smallest power it has seen (bin_phase_gcc_fused).
__device__ __forceinline__ void pair_phasors_raw(const float2* ch, const float* val, float2* p) {
    float2 uc[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float inv = rsqrt_ftz(val[c]);
        uc[c] = pmul(ch[c], make_float2(inv, inv));
    }
    constexpr int PM[6] = {0, 0, 0, 1, 1, 2}, PN[6] = {1, 2, 3, 2, 3, 3};
#pragma unroll
    for (int q = 0; q < 6; ++q) p[q] = unit_pair(uc[PM[q]], uc[PN[q]], false);
}

// The reference on a DEAD channel (an exactly zero spectrum): R = conj(X_m) X_n is a signed zero, torch.angle gives pi where
// its real part is -0 -- which is where the live partner has Re < 0 and Im < 0 (sign bits; measured on the reference's torch
// build for either operand order) -- and exp(1j * pi) = (-1, -8.74e-8) in complex64.  Two dead channels give +0 -> 1.
__device__ __forceinline__ float2 dead_pair(float2 xm, float2 xn, bool dm, bool dn, float2 live_pair) {
    if (!(dm || dn)) return live_pair;
    if (dm && dn) return make_float2(1.f, 0.f);
    const float2 x = dm ? xn : xm;
    const bool neg = (__float_as_uint(x.x) & __float_as_uint(x.y)) >> 31;
    return neg ? make_float2(-1.f, -8.742278e-8f) : make_float2(1.f, 0.f);
}

// Bin phase of the fused MIC kernel: powers -> mel pieces as in bin_phase<MODE_MIC>, pair phasors written in place as the
// B operand.  DEAD (rare, rolled loop): some channel of this frame is exactly zero (dead bits: bit c = channel c).
// (Measured, 600 MIC clips: the channel split as four packed FFMA2 instead of eight scalar adds 13.87 ms against 13.75 -- the eight
//  32-bit loads land in unpaired registers.  Instruction cuts pay in this kernel only when they cost no registers: it runs at the
//  128-register limit with 64 B of spills.)
template <bool DEAD>
__device__ __forceinline__ void bin_phase_gcc_fused(unsigned char* tile, const unsigned char* nyq, const Tables& tb, float2* P, int u,
                                                    unsigned taddr_w01, unsigned dead) {
    constexpr int N = 1024, BPT = 9, NV = 4, PSTRIDE = PieceGeo<MODE_MIC>::PSTRIDE;
    const int kbeg = u * BPT;
    const unsigned long long endmask = tb.endmask[u];
    int slot = tb.slot0[u], next = tb.slot1[u];
    float2 acc2[NV];
#pragma unroll
    for (int c = 0; c < NV; ++c) acc2[c] = make_float2(0.f, 0.f);
    // Integer work hoisted out of the (unrolled) bin loop: bin k = 9u + i sits at tile + 36 k - 32 (k & 3) with (k & 3) = (u + i) & 3,
    // i.e. at colbase[i & 3] + 36 i -- four per-lane bases, the rest immediate offsets; the record cursor is kept as two
    // shared-memory addresses advanced under the flush predicate (piece_flush_cursor).  ~15 -> ~6 integer instructions per bin.
    unsigned char* colbase[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) colbase[j] = tile + 324 * u - 32 * ((u + j) & 3);
    const bool past = u >= 57;                           // bins 513 .. 575 do not exist: zero weights, no stores, Nyquist column reads
    const unsigned emask = static_cast<unsigned>(endmask);
    unsigned rec_addr = static_cast<unsigned>(__cvta_generic_to_shared(P + slot * PSTRIDE));
    unsigned nxt_addr = static_cast<unsigned>(__cvta_generic_to_shared(P + next * PSTRIDE));

    auto spectra = [&](const unsigned char* col, bool real_bin, float2* ch, float* val) {
        float f[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) f[r] = *reinterpret_cast<const float*>(col + 16 * r);
        if (real_bin) { f[2] = f[0]; f[3] = f[1]; f[6] = f[4]; f[7] = f[5]; }      // Z[N - k] is Z[k] itself for k = 0, N/2
        ch[0] = make_float2(f[0] + f[2], f[1] - f[3]);                             // twice the channel spectra
        ch[1] = make_float2(f[1] + f[3], f[2] - f[0]);
        ch[2] = make_float2(f[4] + f[6], f[5] - f[7]);
        ch[3] = make_float2(f[5] + f[7], f[6] - f[4]);
        if constexpr (DEAD) {
#pragma unroll
            for (int c = 0; c < 4; ++c) if ((dead >> c) & 1u) ch[c] = make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) val[c] = fmaf(ch[c].x, ch[c].x, ch[c].y * ch[c].y);
    };
    unsigned vmin = 0x7f800000u;                         // smallest power seen, as its bit pattern (powers are >= +0, or NaN: not small)
    auto phasors = [&](const float2* ch, const float* val, float2* p, auto vote) {
        if constexpr (!DEAD) {
            // live frames: no per-bin zero test.  A bin with a zero channel is rare; its words are patched after the loop
            pair_phasors_raw(ch, val, p);
            vmin = min(min(vmin, min(__float_as_uint(val[0]), __float_as_uint(val[1]))), min(__float_as_uint(val[2]), __float_as_uint(val[3])));
            return;
        }
        pair_phasors<decltype(vote)::value>(ch, val, p);
        if constexpr (DEAD) {
            constexpr int PM[6] = {0, 0, 0, 1, 1, 2}, PN[6] = {1, 2, 3, 2, 3, 3};
#pragma unroll
            for (int q = 0; q < 6; ++q) p[q] = dead_pair(ch[PM[q]], ch[PN[q]], (dead >> PM[q]) & 1u, (dead >> PN[q]) & 1u, p[q]);
        }
    };
    auto step = [&](int i, bool first, bool last) {
        const int k = kbeg + i;
        // bins past N/2 (lanes 57 .. 63) carry zero weights and never close a piece: they read the Nyquist column
        // (bin 512 of lane 56 is the Nyquist column by address: it follows the tile)
        // The DC and the Nyquist bin are real and share word 0 of each row (low / high half).  Lane 0 takes DC in its first step --
        // its spectrum sits in the third words of the Nyquist column's units, so that bin 0's words are write-only here -- and lane
        // 63, which has no bins, takes the Nyquist bin in ITS first step: both store 16-bit halves.  (Lane 0 used to compute both
        // in a divergent block of ~60 instructions that the team's other warp waited for at the barrier.)
        const bool dc = first && k == 0, ny = first && u == 63;
        const unsigned char* col;
        if constexpr (DEAD) col = (k >= N / 2) ? nyq : tile + 36 * k - 32 * (k & 3);
        else col = past ? nyq : colbase[i & 3] + 36 * i;
        if (dc) col = nyq + 8;
        float2 ch[4];
        float val[NV];
        spectra(col, dc || ny || (last && k >= N / 2), ch, val);
        float2 p[6];
        phasors(ch, val, p, std::true_type{});           // every lane (the zero fix-up votes); lanes past N/2 compute and drop
        float w[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) w[q] = pack_half2(p[q].x, p[q].y);
        if constexpr (DEAD) {
            if (k < N / 2 && !dc) {
                float* dst = reinterpret_cast<float*>(tile + 36 * k - 32 * (k & 3));
#pragma unroll
                for (int q = 0; q < 6; ++q) dst[4 * q] = w[q];
            }
        } else {
            // every lane stores (no branch around the six stores): the lanes without a bin -- 57 .. 63, and lane 56 at bin 512 --
            // dump into the unused second word of the Nyquist column's units, lane 0's DC step over the words it has just read
            // (12.75 -> 12.59 ms per 600 clips; the same idea cost 92 B of spills and time before the integer work was hoisted)
            const int dump = (last ? (u >= 56) : past) ? 4 : 0;
            float* dst = reinterpret_cast<float*>(const_cast<unsigned char*>(col) + dump);
#pragma unroll
            for (int q = 0; q < 6; ++q) dst[4 * q] = w[q];
        }
        if (dc || ny) {
            unsigned short* d16 = reinterpret_cast<unsigned short*>(tile + (ny ? 2 : 0));
#pragma unroll
            for (int q = 0; q < 6; ++q) d16[8 * q] = static_cast<unsigned short>(__float_as_uint(w[q]) & 0xffffu);     // fp16(Re)
        }
        float2 wt;
        tmem_ld2(taddr_w01 + 2 * i, wt.x, wt.y);         // (four bins' weights per tcgen05.ld.x8 measured slower: 12.77 against 12.59 ms)
#pragma unroll
        for (int c = 0; c < NV; ++c) acc2[c] = pfma(make_float2(val[c], val[c]), wt, acc2[c]);
        if constexpr (DEAD) {
            const unsigned flag = static_cast<unsigned>(endmask >> i) & 1u;
            piece_flush<NV>(acc2, P + slot * PSTRIDE, flag);
            slot = flag ? next : slot;
            next += int(flag);
        } else {
            piece_flush_cursor(acc2, rec_addr, nxt_addr, emask, 1u << i);
        }
    };
    if constexpr (DEAD) {
#pragma unroll 1
        for (int i = 0; i < BPT; ++i) step(i, i == 0, i == BPT - 1);
    } else {
#pragma unroll
        for (int i = 0; i < BPT; ++i) step(i, i == 0, i == BPT - 1);
        // rare: some lane of the warp met a power below FLT_MIN (the test the per-bin fix-up made: a pair with such a channel is
        // exp(i angle(0)) = 1).  Its words hold inf / NaN halves; every lane rescans its own words.
        if (__any_sync(0xffffffffu, vmin < 0x00800000u)) {
            auto bad = [](unsigned h) { return (h & 0x7c00u) == 0x7c00u; };
#pragma unroll 1
            for (int i = 0; i < BPT; ++i) {
                const int k = kbeg + i;
                if (k >= N / 2) break;
                if (k == 0) continue;                    // (bin 0's words are patched half by half below)
                unsigned* w = reinterpret_cast<unsigned*>(tile + 36 * k - 32 * (k & 3));
#pragma unroll 1
                for (int q = 0; q < 6; ++q) if (bad(w[4 * q]) || bad(w[4 * q] >> 16)) w[4 * q] = 0x00003c00u;      // (1, 0)
            }
            if (u == 0 || u == 63) {                     // DC (low halves) / Nyquist (high halves): real, 1 where a channel is zero
                unsigned short* d16 = reinterpret_cast<unsigned short*>(tile + (u ? 2 : 0));
#pragma unroll 1
                for (int q = 0; q < 6; ++q) if (bad(d16[8 * q])) d16[8 * q] = 0x3c00u;
            }
        }
    }
}

// kind::f16 MMA with the A operand in tensor memory: D[tmem] (+)= A[tmem] * B[smem descriptor]
__device__ __forceinline__ void mma_f16_ts(unsigned d_tmem, unsigned a_tmem, unsigned long long b_desc, unsigned idesc, unsigned accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

// Issued once per team PAIR and frame slot, after both teams' phasor rows are in place (and fenced to the async proxy): the two
// tiles are the two 8-row groups of ONE N = 16 operand (SBO = byte distance between the teams' tiles), so a single chain of
// MMAs (M64 N16 K16, ~11 cycles each -- measured, tools/microbench/probe_tmem_ts.cu -- against 2 x 8 for two N = 8 chains) serves
// both frames: half the UTCHMMA instructions through the MIO queue.  An elected thread of each warp of the issuing team takes
// atom h -- K half h of the basis, datapath lanes 16 h .. 16 h + 15 -- 32 MMAs, and tcgen05.commit arrives on the pair's
// mbarrier (count 2) when they have read the tiles and written D.  A team without a frame contributes stale rows: its eight
// accumulator columns are simply not read.
__device__ __forceinline__ void gcc_issue_mma(unsigned tile0_saddr, unsigned tile_stride, unsigned tmem_base, int dcol, unsigned mbar_saddr, int half) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    constexpr unsigned idesc = (1u << 4) | ((16u >> 3) << 17) | ((64u >> 4) << 24);       // F32 accumulate, F16 x F16, K-major, N = 16, M = 64
    // no-swizzle K-major descriptor (cute::UMMA::SmemDescriptor): start address, LBO = 144 (K-adjacent core matrices),
    // SBO = distance between the two 8-row groups, version 1
    const unsigned long long desc = (unsigned long long)((tile0_saddr >> 4) & 0x3FFF) | ((unsigned long long)(GT_UNIT >> 4) << 16) |
                                    ((unsigned long long)((tile_stride >> 4) & 0x3FFF) << 32) | (1ull << 46);
    const unsigned up = unsigned(half) * (16u << 16);
    const unsigned d = tmem_base + dcol + up, a = tmem_base + TMEM_COL_BASIS + up;
    const unsigned long long desc_h = desc + (unsigned long long)(unsigned(half) * ((2 * GT_UNIT * 32) >> 4));
#pragma unroll
    for (int s = 0; s < 32; ++s) mma_f16_ts(d, a + 8 * s, desc_h + (unsigned long long)((2 * GT_UNIT * s) >> 4), idesc, s > 0);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(mbar_saddr) : "memory");
}

__device__ __forceinline__ void mbar_wait_parity(unsigned bar_saddr, unsigned parity) {
    unsigned done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar_saddr), "r"(parity) : "memory");
    } while (!done);
}

// Epilogue of a pair's accumulator (16 columns: the even team's 8, the odd team's 8), executed by every warp of the pair (four
// warps = the four TMEM subpartitions): warp quadrant q holds lags 16 q .. 16 q + 15 -- K < 512 partial sums in lanes 0..15,
// K >= 512 in lanes 16..31.  The six GCC channels of a lag are 24 contiguous bytes of a team's staged feature row.
__device__ __forceinline__ void gcc_epilogue(unsigned taddr_quadrant, int dcol, int q, int lane, float* acc_even, float* acc_odd) {
    float v[16];
    tmem_ld16(taddr_quadrant + dcol, v);
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        float* acc_row = t ? acc_odd : acc_even;
        if (acc_row == nullptr) continue;                                  // (warp-uniform: that team has no frame in this slot)
        float w[6];
#pragma unroll
        for (int n = 0; n < 6; ++n) w[n] = (v[8 * t + n] + __shfl_xor_sync(0xffffffffu, v[8 * t + n], 16)) * (1.0f / 512.0f);     // the basis is stored x512
        if (lane < 16) {
            float2* d = reinterpret_cast<float2*>(acc_row + (16 * q + lane) * 10 + 4);
            d[0] = make_float2(w[0], w[1]);
            d[1] = make_float2(w[2], w[3]);
            d[2] = make_float2(w[4], w[5]);
        }
    }
}
#endif

// ---------------------------------------------------------------- GCC-PHAT, packed inverse transform q
// Pair order (reference feature_extractor.py:207-208): 0:(0,1) 1:(0,2) 2:(0,3) 3:(1,2) 4:(1,3) 5:(2,3).
// Transform q carries pairs 2q (real part of the result) and 2q+1 (imaginary part).

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
    pfft_dif<R>(v);
#pragma unroll
    for (int p = 0; p < R; ++p) E[lane * G::EP + bitrev(p, G::LOG2R)] = v[p];
}

// W_32^-b as compile-time constants for the fast GCC stage 2
template <int B>
SELD_HD float2 mul_conj_w32(float2 t) { return pcmul(t, Tw<B, 32>::re, -Tw<B, 32>::im); }

template <int B>
struct GccAccum {    // out0 += u[b] W^(b r);  out1 += u[b] W^(b r) W_32^-b
    static SELD_HD void run(const float2* u, const float2* twcol, float2& o0, float2& o1) {
        const float2 w = twcol[B * 32];
        const float2 t = pcmul(u[B], w.x, w.y);
        o0 = padd(o0, t);
        o1 = padd(o1, mul_conj_w32<B>(t));
        if constexpr (B + 1 < 32) GccAccum<B + 1>::run(u, twcol, o0, o1);
    }
};

template <int R, int Q>
SELD_HD void gcc_stage2(const float2* E, const Tables& tb, float* acc, int n_mels, int lane) {
    using G = Geo<R>;
    constexpr int N = G::N;
    constexpr int C = 10;
    const float inv_n = 1.0f / float(N);
    if (R == 32 && n_mels == 64) {   // (dead code for R != 32 is removed: the condition is a constant there)
        // lags 0..31 are n = r, lags -32..-1 are n = N - 32 + r: W^(b n) = W^(b r) resp. W^(b r) W_32^-b, and
        // W^(b r) = tw_t[b][r] is lane-contiguous
        float2 u[32];
#pragma unroll
        for (int b = 0; b < 32; ++b) u[b] = E[b * G::EP + lane];
        float2 o0 = make_float2(0.f, 0.f), o1 = make_float2(0.f, 0.f);
        GccAccum<0>::run(u, tb.tw_t + lane, o0, o1);
        acc[(32 + lane) * C + 4 + 2 * Q] = o0.x * inv_n;
        acc[(32 + lane) * C + 4 + 2 * Q + 1] = -o0.y * inv_n;
        acc[lane * C + 4 + 2 * Q] = o1.x * inv_n;
        acc[lane * C + 4 + 2 * Q + 1] = -o1.y * inv_n;
        return;
    }
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
                    const float2 t = tb.tw_lin[(b * n) & (N - 1)];
                    sx += u[b].x * t.x - u[b].y * t.y;
                    sy += u[b].x * t.y + u[b].y * t.x;
                }
                acc[j * C + 4 + 2 * Q] = sx * inv_n;          // Re conj(sum)
                acc[j * C + 4 + 2 * Q + 1] = -sy * inv_n;     // Im conj(sum)
            }
        }
    }
}

// ---------------------------------------------------------------- finish one frame row
// Coalesced copy of the staged row acc[m * C + c] to global memory.
SELD_HD void store_row(const float* acc, int row_elems, float* out_row, int u, int n_lanes) {
    if ((row_elems & 3) == 0 && (reinterpret_cast<uintptr_t>(out_row) & 15) == 0) {
        const float4* s4 = reinterpret_cast<const float4*>(acc);
        float4* d4 = reinterpret_cast<float4*>(out_row);
        for (int e = u; e < row_elems / 4; e += n_lanes) d4[e] = s4[e];
    } else {
        for (int e = u; e < row_elems; e += n_lanes) out_row[e] = acc[e];
    }
}

}  // namespace seld
