// Shared building blocks of the seld_b200 kernels (sm_100a).
//
// Everything here is __host__ __device__ so that tests/emu/ can run the exact
// per-lane code of the extractor on the CPU (32 lanes executed in sequence per
// phase, a phase boundary being a __syncwarp() on the device).
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define SELD_HD __host__ __device__ __forceinline__
#else
#include <math.h>
#define SELD_HD inline
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
static inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
#endif

namespace seld {

// ---------------------------------------------------------------- compile-time twiddles
// sin/cos of 2*pi*k/n evaluated in double by octant reduction + Taylor series; exact to
// ~1e-17, then rounded once to float.  Used only as template constants (immediates in SASS).
constexpr double kPi = 3.14159265358979323846264338327950288;

constexpr double taylor_sin(double x) {  // |x| <= pi/4
    double x2 = x * x, term = x, sum = x;
    for (int i = 1; i < 12; ++i) { term *= -x2 / double((2 * i) * (2 * i + 1)); sum += term; }
    return sum;
}
constexpr double taylor_cos(double x) {  // |x| <= pi/4
    double x2 = x * x, term = 1.0, sum = 1.0;
    for (int i = 1; i < 12; ++i) { term *= -x2 / double((2 * i - 1) * (2 * i)); sum += term; }
    return sum;
}
// cos(2 pi k / n), sin(2 pi k / n) for 0 <= k < n, n a power of two >= 8
constexpr double cos2pi(int k, int n) {
    k = ((k % n) + n) % n;
    if (k == 0) return 1.0;
    if (4 * k == n) return 0.0;
    if (2 * k == n) return -1.0;
    if (4 * k == 3 * n) return 0.0;
    if (2 * k > n) return cos2pi(n - k, n);              // cos(2pi - a) = cos a
    if (4 * k > n) return -cos2pi(n / 2 - k, n);         // cos(pi - a) = -cos a
    if (8 * k > n) return taylor_sin(2.0 * kPi * double(n / 4 - k) / double(n));   // cos a = sin(pi/2 - a)
    return taylor_cos(2.0 * kPi * double(k) / double(n));
}
constexpr double sin2pi(int k, int n) { return cos2pi(k - n / 4, n); }   // sin a = cos(a - pi/2)

template <int K, int N>
struct Tw {   // forward twiddle W_N^K = exp(-2 pi i K / N)
    static constexpr float re = float(cos2pi(K, N));
    static constexpr float im = float(-sin2pi(K, N));
};

constexpr int bitrev(int v, int bits) {
    int r = 0;
    for (int i = 0; i < bits; ++i) { r = (r << 1) | ((v >> i) & 1); }
    return r;
}
constexpr int ilog2(int n) { return n <= 1 ? 0 : 1 + ilog2(n / 2); }

// ---------------------------------------------------------------- complex helpers
SELD_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
SELD_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
SELD_HD float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
SELD_HD float2 cconj(float2 a) { return make_float2(a.x, -a.y); }

// d * W_N^J with the trivial cases folded at compile time
template <int J, int N>
SELD_HD float2 mul_tw(float2 d) {
    if constexpr (J == 0) {
        return d;
    } else if constexpr (4 * J == N) {            // -i
        return make_float2(d.y, -d.x);
    } else if constexpr (8 * J == N) {            // (1 - i)/sqrt2
        constexpr float h = 0.70710678118654752440f;
        return make_float2((d.x + d.y) * h, (d.y - d.x) * h);
    } else if constexpr (8 * J == 3 * N) {        // (-1 - i)/sqrt2
        constexpr float h = 0.70710678118654752440f;
        return make_float2((d.y - d.x) * h, -(d.x + d.y) * h);
    } else {
        return make_float2(d.x * Tw<J, N>::re - d.y * Tw<J, N>::im, d.x * Tw<J, N>::im + d.y * Tw<J, N>::re);
    }
}

// ---------------------------------------------------------------- in-register radix-2 DIF FFT
// Forward DFT of v[0..N) in place; result in BIT-REVERSED order: v[p] = X[bitrev(p)].
// All indices are compile-time after unrolling, so v stays in registers.
template <int N, int J>
struct Butterflies {
    static SELD_HD void run(float2* v) {
        float2 a = v[J], b = v[J + N / 2];
        v[J] = cadd(a, b);
        v[J + N / 2] = mul_tw<J, N>(csub(a, b));
        if constexpr (J + 1 < N / 2) Butterflies<N, J + 1>::run(v);
    }
};

template <int N>
SELD_HD void fft_dif(float2* v) {
    if constexpr (N >= 2) {
        Butterflies<N, 0>::run(v);
        fft_dif<N / 2>(v);
        fft_dif<N / 2>(v + N / 2);
    }
}

// ---------------------------------------------------------------- ordered float <-> uint key (atomicMax)
SELD_HD uint32_t float_to_key(float f) {
#if defined(__CUDA_ARCH__)
    uint32_t b = __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; uint32_t b = c.u;
#endif
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
constexpr uint32_t kNanKey = 0xFFC00000u;      // key of +NaN: above every finite value and +inf (a NaN clip maximum poisons the clip)
SELD_HD float key_to_float(uint32_t k) {
    uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
#if defined(__CUDA_ARCH__)
    return __uint_as_float(b);
#else
    union { float f; uint32_t u; } c; c.u = b; return c.f;
#endif
}

}  // namespace seld
