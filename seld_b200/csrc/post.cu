// top_db clamp, dataset statistics and normalisation, sm_100a (HBM-bound streaming kernels).
//
//   seld_finalize      torchaudio amplitude_to_DB top_db clamp (reference feature_extractor.py:65-71) fused
//                      with apply_normalizer's (x - mean) / max(std, eps) (reference :226-234)
//   seld_stats         calculate_statistics (reference :218-223): per-(mel, chan) sum / sum of squares in
//                      float64, fixed reduction order, clamp applied on the fly
//   seld_stats_finish  mean / population std from the (all-reduced) accumulators
#include <math.h>

#include "plan.h"
#include "seld_common.cuh"

namespace seld {

// ------------------------------------------------------------------ row walkers
// Both streaming kernels use the same geometry: a block is (G4 float4 column groups) x (RP row phases); a thread owns 4
// fixed (mel, chan) columns -- so its clamp mask, mean and 1/std live in registers -- and walks the block's contiguous row
// slab with UNROLL independent 128-bit loads in flight.  (clip, t) advance incrementally: no division per element.
constexpr int kUnroll = 4;
#ifndef SELD_STATS_UNROLL
#define SELD_STATS_UNROLL 4
#endif
#ifndef SELD_STATS_BLOCKS_PER_SM
#define SELD_STATS_BLOCKS_PER_SM 2      // measured (600 clips, C = 7 | 10): 2 -> 0.611 | 0.865 ms, 4 -> 0.637 | 0.873, 8 -> 0.666 | 0.899; 8 loads in flight: 0.85 | 1.12
#endif
constexpr int kStatsUnroll = SELD_STATS_UNROLL;

struct RowCursor {
    long long clip;
    int t;
    __device__ __forceinline__ void init(long long row, int t_out) { clip = row / t_out; t = int(row - clip * t_out); }
    __device__ __forceinline__ void advance(int rows, int t_out) {
        t += rows;
        while (t >= t_out) { t -= t_out; ++clip; }
    }
};

__device__ __forceinline__ float floor_of(const unsigned int* __restrict__ keys, const RowCursor& c, int t_valid, float top_db) {
    return (keys != nullptr && c.t < t_valid) ? key_to_float(__ldg(keys + c.clip)) - top_db : -INFINITY;
}

// torch.max(db, floor): NaN in either operand gives NaN (fmaxf alone would drop it) -- a clip with a NaN sample has a NaN
// maximum, hence a NaN floor, hence NaN log-mel everywhere, as in the reference (feature_extractor.py:65-71)
__device__ __forceinline__ float clamp_db(float x, float fl) {
    const float r = fmaxf(x, fl);
    return (x != x || fl != fl) ? NAN : r;
}

// ------------------------------------------------------------------ finalize (vector path: row_len % 4 == 0)
__global__ void __launch_bounds__(512) finalize_rows_kernel(const float* __restrict__ in, const unsigned int* __restrict__ keys,
                                                            long long n_rows, int t_out, int t_valid, int row_len, int n_ch,
                                                            float top_db, const float* __restrict__ mean,
                                                            const float* __restrict__ stdv, float eps, int g4, int rp,
                                                            float* __restrict__ out) {
    const int g = threadIdx.x % g4, ph = threadIdx.x / g4;
    const long long rows_per_block = (n_rows + gridDim.x - 1) / gridDim.x;
    const long long r0 = (long long)blockIdx.x * rows_per_block;
    long long r1 = r0 + rows_per_block;
    if (r1 > n_rows) r1 = n_rows;
    bool logmel[4];
    float m[4], inv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        logmel[j] = ((4 * g + j) % n_ch) < 4;
        m[j] = mean ? mean[4 * g + j] : 0.f;
        inv[j] = mean ? 1.0f / fmaxf(stdv[4 * g + j], eps) : 1.f;
    }
    RowCursor cur;
    cur.init(r0 + ph, t_out);
    const float4* src = reinterpret_cast<const float4*>(in);
    float4* dst = reinterpret_cast<float4*>(out);
    const long long stride4 = (long long)rp * g4;             // float4 elements between a thread's consecutive rows
    long long r = r0 + ph;
    for (; r + (long long)(kUnroll - 1) * rp < r1; r += (long long)kUnroll * rp) {
        float4 v[kUnroll];
        float fl[kUnroll];
        const long long base = r * g4 + g;
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) v[u] = __ldcs(src + base + u * stride4);
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) { fl[u] = floor_of(keys, cur, t_valid, top_db); cur.advance(rp, t_out); }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            float x[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (logmel[j]) x[j] = clamp_db(x[j], fl[u]);
                x[j] = (x[j] - m[j]) * inv[j];
            }
            __stcs(dst + base + u * stride4, make_float4(x[0], x[1], x[2], x[3]));
        }
    }
    for (; r < r1; r += rp) {
        const float fl = floor_of(keys, cur, t_valid, top_db);
        cur.advance(rp, t_out);
        const float4 q = __ldcs(src + r * g4 + g);
        float x[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (logmel[j]) x[j] = clamp_db(x[j], fl);
            x[j] = (x[j] - m[j]) * inv[j];
        }
        __stcs(dst + r * g4 + g, make_float4(x[0], x[1], x[2], x[3]));
    }
}

// generic fallback (row_len % 4 != 0 or unaligned): one element per thread
__global__ void __launch_bounds__(256) finalize_scalar_kernel(const float* __restrict__ in, const unsigned int* __restrict__ keys,
                                                              long long n, int t_out, int t_valid, int row_len, int n_ch,
                                                              float top_db, const float* __restrict__ mean,
                                                              const float* __restrict__ stdv, float eps, float* __restrict__ out) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
        const long long row = e / row_len;
        const int p = int(e - row * row_len);
        float v = in[e];
        if ((p % n_ch) < 4 && keys != nullptr && int(row % t_out) < t_valid) v = clamp_db(v, key_to_float(keys[row / t_out]) - top_db);
        if (mean != nullptr) v = (v - mean[p]) * (1.0f / fmaxf(stdv[p], eps));
        out[e] = v;
    }
}

// ------------------------------------------------------------------ statistics
// Partial sums stay in float64 registers, are folded over the row phases through shared memory in a fixed order, and
// land in partials[block][2 * row_len]; stats_fold_kernel adds the blocks in block order => run-to-run deterministic.
__global__ void __launch_bounds__(512) stats_partial_kernel(const float* __restrict__ x, const unsigned int* __restrict__ keys,
                                                            long long n_rows, int t_out, int t_valid, int row_len, int n_ch,
                                                            float top_db, int g4, int rp, double* __restrict__ partials) {
    extern __shared__ double red[];   // [rp][2 * row_len]
    const int g = threadIdx.x % g4;
    const int ph = threadIdx.x / g4;
    double s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
    const long long rows_per_block = (n_rows + gridDim.x - 1) / gridDim.x;
    const long long r0 = (long long)blockIdx.x * rows_per_block;
    long long r1 = r0 + rows_per_block;
    if (r1 > n_rows) r1 = n_rows;
    bool logmel[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) logmel[j] = ((4 * g + j) % n_ch) < 4;
    RowCursor cur;
    cur.init(r0 + ph, t_out);
    const float4* src = reinterpret_cast<const float4*>(x);
    const long long stride4 = (long long)rp * g4;
    long long r = r0 + ph;
    auto add = [&](const float4& v4, float fl) {
        float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (logmel[j]) v[j] = clamp_db(v[j], fl);
            const double d = double(v[j]);
            s[j] += d;
            q[j] = fma(d, d, q[j]);
        }
    };
    for (; r + (long long)(kStatsUnroll - 1) * rp < r1; r += (long long)kStatsUnroll * rp) {
        float4 v[kStatsUnroll];
        const long long base = r * g4 + g;
#pragma unroll
        for (int u = 0; u < kStatsUnroll; ++u) v[u] = __ldcs(src + base + u * stride4);
#pragma unroll
        for (int u = 0; u < kStatsUnroll; ++u) {
            const float fl = floor_of(keys, cur, t_valid, top_db);
            cur.advance(rp, t_out);
            add(v[u], fl);
        }
    }
    for (; r < r1; r += rp) {
        const float fl = floor_of(keys, cur, t_valid, top_db);
        cur.advance(rp, t_out);
        add(__ldcs(src + r * g4 + g), fl);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        red[(size_t)ph * 2 * row_len + 4 * g + j] = s[j];
        red[(size_t)ph * 2 * row_len + row_len + 4 * g + j] = q[j];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * row_len; i += blockDim.x) {
        double a = 0.0;
        for (int k = 0; k < rp; ++k) a += red[(size_t)k * 2 * row_len + i];
        partials[(size_t)blockIdx.x * 2 * row_len + i] = a;
    }
}

__global__ void stats_fold_kernel(const double* __restrict__ partials, int n_blocks, int n2, double n_rows,
                                  double* __restrict__ acc) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n2) {
        double a = 0.0;
        for (int b = 0; b < n_blocks; ++b) a += partials[(size_t)b * n2 + i];
        acc[i] += a;
    }
    if (i == 0) acc[n2] += n_rows;
}

// ------------------------------------------------------------------ one-shot all-reduce over NVLink peer memory
// The path's only collective: the per-bin {sum, sum of squares, count} vector, <= 1 281 doubles.  NCCL spends ~70 us
// on it inside an 8-GPU step of 1.5 ms (profiles/r2_bench_n8_a.json); this kernel does it in one launch over buffers every rank
// can address (torch symmetric memory / cudaIpc -- the caller hands in the peer base pointers):
//   publish   acc -> own exchange slot (parity of a device-side epoch counter), system-scope fence
//   signal    flag[my rank] := epoch in EVERY rank's buffer (st.release.sys over NVLink)
//   wait      until every rank's flag in the own buffer has reached the epoch (ld.acquire.sys)
//   reduce    sum the W peer slots in RANK ORDER (L1-bypassing loads) -> acc: every rank adds the same numbers in the same order,
//             so the result is bit-identical on all ranks and run to run
// Two slots suffice: a rank can only overwrite slot e & 1 at epoch e + 2, and to get there it has passed barrier e + 1, which every
// peer signals only after it finished reading epoch e.  The epoch lives in device memory, so a captured CUDA graph replays it.
// Buffer of a rank: [0, 256) flags (uint32 per rank), [256, 264) epoch counter, [512, ...) two slots of n doubles.
constexpr int kPeerHeaderBytes = 512;
constexpr int kPeerMaxWorld = 64;

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(1024) stats_peer_allreduce_kernel(const unsigned long long* __restrict__ peer_base, int rank, int world, int n,
                                                                    double* __restrict__ acc) {
    __shared__ unsigned s_epoch;
    __shared__ unsigned long long s_base[kPeerMaxWorld];
    if (threadIdx.x < world) s_base[threadIdx.x] = peer_base[threadIdx.x];
    __syncthreads();
    unsigned char* mine = reinterpret_cast<unsigned char*>(s_base[rank]);
    if (threadIdx.x == 0) {
        unsigned* ep = reinterpret_cast<unsigned*>(mine + 256);
        s_epoch = *ep + 1u;
        *ep = s_epoch;
    }
    __syncthreads();
    const unsigned epoch = s_epoch;
    const size_t slot_off = kPeerHeaderBytes + size_t(epoch & 1u) * n * sizeof(double);
    double* my_slot = reinterpret_cast<double*>(mine + slot_off);
    for (int i = threadIdx.x; i < n; i += blockDim.x) my_slot[i] = acc[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < world) {
        st_release_sys(reinterpret_cast<unsigned*>(reinterpret_cast<unsigned char*>(s_base[threadIdx.x])) + rank, epoch);
        const unsigned* flag = reinterpret_cast<const unsigned*>(mine) + threadIdx.x;
        while (int(ld_acquire_sys(flag) - epoch) < 0) { }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double a = 0.0;
        for (int r = 0; r < world; ++r) a += ld_relaxed_sys_f64(reinterpret_cast<const double*>(reinterpret_cast<const unsigned char*>(s_base[r]) + slot_off) + i);
        acc[i] = a;
    }
}

// generic (row_len % 4 != 0) fallback: one thread per column, block-strided rows, same fixed-order fold
__global__ void __launch_bounds__(256) stats_partial_scalar_kernel(const float* __restrict__ x, const unsigned int* __restrict__ keys,
                                                                   long long n_rows, int t_out, int t_valid, int row_len,
                                                                   int n_ch, float top_db, double* __restrict__ partials) {
    const long long rows_per_block = (n_rows + gridDim.x - 1) / gridDim.x;
    const long long r0 = (long long)blockIdx.x * rows_per_block;
    long long r1 = r0 + rows_per_block;
    if (r1 > n_rows) r1 = n_rows;
    for (int p = threadIdx.x; p < row_len; p += blockDim.x) {
        const bool logmel = (p % n_ch) < 4;
        double s = 0, q = 0;
        for (long long r = r0; r < r1; ++r) {
            float v = x[r * row_len + p];
            if (logmel && keys != nullptr && int(r % t_out) < t_valid) v = clamp_db(v, key_to_float(keys[r / t_out]) - top_db);
            s += double(v);
            q = fma(double(v), double(v), q);
        }
        partials[(size_t)blockIdx.x * 2 * row_len + p] = s;
        partials[(size_t)blockIdx.x * 2 * row_len + row_len + p] = q;
    }
}

__global__ void stats_finish_kernel(const double* __restrict__ acc, int n, float* __restrict__ mean, float* __restrict__ stdv) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const double cnt = acc[2 * n];
        const double m = acc[i] / cnt;
        double var = acc[n + i] / cnt - m * m;
        if (var < 0.0) var = 0.0;
        mean[i] = float(m);
        stdv[i] = float(sqrt(var));
    }
}

}  // namespace seld

using namespace seld;

extern "C" {

static int sm_count() { return device_sm_count(); }
static int stats_block_count() { return sm_count() * SELD_STATS_BLOCKS_PER_SM; }

int seld_finalize(int n_mels, int n_ch, const float* feat_in_dev, const uint32_t* clip_max_key_dev, int n_clips, int t_out,
                  int t_valid, float top_db, const float* mean_dev, const float* std_dev, float eps, float* feat_out_dev,
                  void* stream) {
    if (!feat_in_dev || !feat_out_dev || n_mels < 1 || n_ch < 1) { set_error("bad argument"); return SELD_EINVAL; }
    if ((mean_dev == nullptr) != (std_dev == nullptr)) { set_error("mean and std must be given together"); return SELD_EINVAL; }
    if (n_clips < 0 || t_out < 0) { set_error("negative size"); return SELD_EINVAL; }
    const int row_len = n_mels * n_ch;
    const long long n = (long long)n_clips * t_out * row_len;
    if (n == 0) return SELD_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int g4 = row_len / 4;
    const bool vec = (row_len % 4 == 0) && g4 <= 512 &&
                     ((reinterpret_cast<uintptr_t>(feat_in_dev) | reinterpret_cast<uintptr_t>(feat_out_dev)) % 16 == 0);
    if (vec) {
        const long long n_rows = (long long)n_clips * t_out;
        const int rp = 512 / g4;
        long long blocks = (long long)sm_count() * 4;
        if (blocks > n_rows) blocks = n_rows;
        finalize_rows_kernel<<<(int)blocks, g4 * rp, 0, st>>>(feat_in_dev, clip_max_key_dev, n_rows, t_out, t_valid, row_len, n_ch,
                                                              top_db, mean_dev, std_dev, eps, g4, rp, feat_out_dev);
    } else {
        long long blocks = (n + 255) / 256;
        const long long cap = (long long)sm_count() * 16;
        if (blocks > cap) blocks = cap;
        finalize_scalar_kernel<<<(int)blocks, 256, 0, st>>>(feat_in_dev, clip_max_key_dev, n, t_out, t_valid, row_len, n_ch, top_db,
                                                            mean_dev, std_dev, eps, feat_out_dev);
    }
    SELD_CUDA_TRY(cudaGetLastError());
    seld::note_launch();
    return SELD_OK;
}

int64_t seld_stats_workspace_doubles(int n_mels, int n_ch) {
    if (n_mels < 1 || n_ch < 1) return SELD_EINVAL;
    return (int64_t)stats_block_count() * 2 * n_mels * n_ch;
}

int seld_stats(int n_mels, int n_ch, const float* feat_dev, const uint32_t* clip_max_key_dev, int n_clips, int t_out,
               int t_valid, float top_db, double* workspace_dev, double* acc_dev, void* stream) {
    if (!feat_dev || !workspace_dev || !acc_dev || n_mels < 1 || n_ch < 1) { set_error("bad argument"); return SELD_EINVAL; }
    if (n_clips < 0 || t_out < 0) { set_error("negative size"); return SELD_EINVAL; }
    const int row_len = n_mels * n_ch;
    const long long n_rows = (long long)n_clips * t_out;
    if (n_rows == 0) return SELD_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int blocks = stats_block_count();
    if (blocks > n_rows) blocks = (int)n_rows;
    const int g4 = row_len / 4;
    if (row_len % 4 == 0 && g4 <= 512 && reinterpret_cast<uintptr_t>(feat_dev) % 16 == 0) {
        const int rp = 512 / g4;
        const size_t smem = sizeof(double) * rp * 2 * row_len;
        SELD_CUDA_TRY(cudaFuncSetAttribute(stats_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        stats_partial_kernel<<<blocks, g4 * rp, smem, st>>>(feat_dev, clip_max_key_dev, n_rows, t_out, t_valid, row_len,
                                                            n_ch, top_db, g4, rp, workspace_dev);
    } else {
        stats_partial_scalar_kernel<<<blocks, 256, 0, st>>>(feat_dev, clip_max_key_dev, n_rows, t_out, t_valid, row_len,
                                                            n_ch, top_db, workspace_dev);
    }
    SELD_CUDA_TRY(cudaGetLastError());
    seld::note_launch();
    const int n2 = 2 * row_len;
    stats_fold_kernel<<<(n2 + 127) / 128, 128, 0, st>>>(workspace_dev, blocks, n2, double(n_rows), acc_dev);
    SELD_CUDA_TRY(cudaGetLastError());
    seld::note_launch();
    return SELD_OK;
}

int64_t seld_stats_peer_buffer_bytes(int n_values) {
    if (n_values < 1) return SELD_EINVAL;
    return kPeerHeaderBytes + 2ll * n_values * (int64_t)sizeof(double);
}

int seld_stats_peer_allreduce(const uint64_t* peer_base_dev, int rank, int world, int n_values, double* acc_dev, void* stream) {
    if (!peer_base_dev || !acc_dev || n_values < 1 || world < 1 || world > kPeerMaxWorld || rank < 0 || rank >= world) {
        set_error("bad argument");
        return SELD_EINVAL;
    }
    stats_peer_allreduce_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const unsigned long long*>(peer_base_dev), rank,
                                                                                  world, n_values, acc_dev);
    SELD_CUDA_TRY(cudaGetLastError());
    seld::note_launch();
    return SELD_OK;
}

int seld_stats_finish(int n_mels, int n_ch, const double* acc_dev, float* mean_dev, float* std_dev, void* stream) {
    if (!acc_dev || !mean_dev || !std_dev || n_mels < 1 || n_ch < 1) { set_error("bad argument"); return SELD_EINVAL; }
    const int n = n_mels * n_ch;
    stats_finish_kernel<<<(n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(acc_dev, n, mean_dev, std_dev);
    SELD_CUDA_TRY(cudaGetLastError());
    seld::note_launch();
    return SELD_OK;
}

}  // extern "C"
