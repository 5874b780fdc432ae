/*
 * seld_b200 -- C ABI of the B200-native (sm_100a) SELD feature-extraction hot path.
 *
 * The reference (IRIS-AUDIO/SELD) has no FFI layer: its boundary is the Python module surface of
 * feature_extractor.py / transforms.py.  seld_b200/feature_extractor.py and seld_b200/transforms.py
 * keep that surface and call the entry points below through ctypes.  Each entry point cites the
 * reference code it replaces (file:line relative to the reference repository).
 *
 * Conventions
 *   - every pointer named *_dev is DEVICE memory owned by the caller; the library never frees it;
 *   - every function returns 0 on success, a negative SELD_E* code on failure; seld_last_error()
 *     returns the message of the calling thread's last failure;
 *   - all work is enqueued on the caller's CUDA stream (`stream` is a cudaStream_t passed as void*);
 *     nothing synchronises the device except seld_plan_create;
 *   - plans are immutable after creation => entry points are re-entrant across threads and streams;
 *   - there is NO CPU fallback: without an sm_100 device seld_plan_create fails with SELD_ENODEVICE.
 */
#ifndef SELD_B200_H
#define SELD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SELD_OK 0
#define SELD_EINVAL (-1)     /* bad argument */
#define SELD_ENODEVICE (-2)  /* no sm_100 CUDA device */
#define SELD_ECUDA (-3)      /* CUDA runtime error (message in seld_last_error) */
#define SELD_EUNSUPPORTED (-4)

#define SELD_MODE_FOA 0 /* 4 log-mel + 3 intensity-vector channels  (reference feature_extractor.py:74-77) */
#define SELD_MODE_MIC 1 /* 4 log-mel + 6 GCC-PHAT channels           (reference feature_extractor.py:78-80) */
#define SELD_MODE_FOA_TF 2 /* the TensorFlow variant of the FOA features (reference data_loader.py:310-349, get_preprocessed_x_tf):
                              mel bank applied to |X| (not |X|^2), dB = 20 log10(mel) with no floor before the top_db clamp (tfio dbscale),
                              frames start at sample t * hop with a zero-padded tail (tf.signal.stft(pad_end=True)); the caller passes
                              tf.signal.hann_window and tf.signal.linear_to_mel_weight_matrix as window / mel table; n_fft = 1024 */

#define SELD_LAYOUT_PLANAR_CL 0      /* wav[clip][chan][sample]  (torchaudio.load layout, feature_extractor.py:43) */
#define SELD_LAYOUT_INTERLEAVED_LC 1 /* wav[clip][sample][chan]  (one 128-bit load = one time step of 4 channels) */

#define SELD_RNG_PHILOX_COUNTER 0   /* Philox4x32-10, counter = (sample, axis slot, chunk, 2*mask+draw) */
#define SELD_RNG_TF_EAGER_COMPAT 1  /* TensorFlow-2 eager op-seed stream (reproduces transforms_test.py:8-30); op_seed2 index
                                      ((sample * n_chunks + chunk) * (time_n + freq_n) + mask) * 2 + draw */
#define SELD_RNG_TF_EAGER_TWO_PASS 2 /* the same stream in the order of TWO reference calls per sample -- mask(axis=-3) then
                                      mask(axis=-2), train.py:157-160: all time draws of all chunks, then all frequency draws:
                                      time  sample * S + (chunk * time_n + mask) * 2 + draw
                                      freq  sample * S + n_chunks * time_n * 2 + (chunk * freq_n + mask) * 2 + draw,
                                      S = n_chunks * (time_n + freq_n) * 2 */

#define SELD_DTYPE_F32 0
#define SELD_DTYPE_F64 1
#define SELD_DTYPE_F16 2
#define SELD_DTYPE_BF16 3
#define SELD_DTYPE_I32 4
#define SELD_DTYPE_I64 5
#define SELD_DTYPE_I16 6
#define SELD_DTYPE_U8 7

typedef struct seld_plan* seld_plan_t;

#if defined(__GNUC__)
#define SELD_API __attribute__((visibility("default")))
#else
#define SELD_API
#endif

SELD_API const char* seld_last_error(void);
SELD_API int seld_version(void);
/* Kernels this library has launched in this process so far (all streams; what bench.py reports as gpu_launches). */
SELD_API int64_t seld_launch_count(void);

/* 0 if device `device` (or the current one when < 0) is compute capability 10.x, else SELD_ENODEVICE. */
SELD_API int seld_device_check(int device);

/*
 * Build an extraction plan: uploads the analysis window, FFT twiddles and the sparse form of the mel bank.
 * Replaces the per-file constants of reference feature_extractor.py:59-60 (MelScale) and :167 (Hann window).
 *   window_host  [n_fft]                 float32 window already zero-padded/centred to n_fft
 *   mel_fb_host  [(n_fft/2+1) * n_mels]  dense float32 filterbank (row = STFT bin); each row may hold at
 *                                        most two non-zeros, in adjacent filters (true of triangular banks)
 *   n_fft in {256, 512, 1024, 2048}; n_chan must be 4; n_mels even for SELD_MODE_MIC.
 */
SELD_API int seld_plan_create(int sample_rate, int n_fft, int win_length, int hop_length, int n_mels, int n_chan, int mode,
                     const float* window_host, const float* mel_fb_host, seld_plan_t* plan_out);
SELD_API int seld_plan_destroy(seld_plan_t plan);

/* Output channels per (frame, mel): 7 (FOA) or 10 (MIC). */
SELD_API int seld_plan_out_channels(seld_plan_t plan);
/* Number of STFT frames of an n_samples-long clip: 1 + n_samples / hop  (torch.stft, center=True). */
SELD_API int64_t seld_plan_num_frames(seld_plan_t plan, int64_t n_samples);

/*
 * Fused extractor: reference feature_extractor.py:53-88 (extract_features) + the feature half of :117-149
 * (pad / truncate to t_out frames), batched over clips.
 *   wav_dev            [n_clips][4][n_samples] (PLANAR_CL) or [n_clips][n_samples][4] (INTERLEAVED_LC) float32
 *   feat_raw_dev       [n_clips][t_out][n_mels][C] float32; log-mel channels are written WITHOUT the top_db
 *                      clamp (it needs the clip-global maximum); rows >= 1 + n_samples/hop are zero
 *   clip_max_key_dev   [n_clips] uint32 order-preserving keys of the per-clip maximum dB over ALL frames
 *                      (including frames >= t_out, as the reference takes the max before truncating);
 *                      reset by this call; decode with seld_clip_max_decode.  A NaN log-mel value makes the key NaN's
 *                      (the reference's db.max() is NaN then, and with it the whole clip's log-mel block).
 *   workspace_dev      unused (kept for ABI stability); workspace_bytes < 0 selects the CUDA-core GCC path for MIC
 *                      plans that would otherwise use the fused tensor-core lag projection (n_fft 1024, 64 lags)
 */
SELD_API int seld_extract(seld_plan_t plan, const float* wav_dev, int layout, int n_clips, int64_t n_samples, int t_out,
                 float* feat_raw_dev, uint32_t* clip_max_key_dev, void* workspace_dev, int64_t workspace_bytes, void* stream);

/*
 * Always 0 since the tensor-core GCC lag projection (reference feature_extractor.py:209-211) runs inside the extractor:
 * basis resident in tensor memory, pair-phasor rows in shared memory, tcgen05.mma per frame -- no scratch in HBM.
 * Kept so that round-1 callers still link.
 */
SELD_API int64_t seld_extract_workspace_bytes(seld_plan_t plan, int n_clips, int64_t n_samples, int t_out);

/*
 * Same as seld_extract for 16-bit PCM input, pcm_dev[n_clips][n_samples][4] int16 -- the frame order of a WAV `data`
 * chunk -- decoded like torchaudio.load (reference feature_extractor.py:43): sample / 32768.  Halves the input bytes
 * (H2D and HBM) against float32; bit-identical to decoding on the host and calling seld_extract.
 */
SELD_API int seld_extract_pcm16(seld_plan_t plan, const int16_t* pcm_dev, int n_clips, int64_t n_samples, int t_out,
                                float* feat_raw_dev, uint32_t* clip_max_key_dev, void* workspace_dev, int64_t workspace_bytes,
                                void* stream);

/*
 * On-the-fly form for training batches (BASELINE.json config 5): every "clip" is a chunk cut from a
 * longer recording WITH its real context, chunk[c] = samples [s0 - n_fft/2, s0 + (t - 1) * hop + n_fft/2), and frame t
 * starts at chunk sample t * hop (no centring, no reflection), so the rows equal rows s0/hop .. of the full-clip
 * extraction bit for bit.  n_samples >= n_fft; frames = 1 + (n_samples - n_fft) / hop.  chunk_max_key_dev receives the
 * chunk maxima; for the reference's clip-global top_db pass the cached clip maxima to seld_finalize instead.
 */
SELD_API int seld_extract_chunks(seld_plan_t plan, const float* wav_dev, int layout, int n_chunks, int64_t n_samples, int t_out,
                                 float* feat_raw_dev, uint32_t* chunk_max_key_dev, void* workspace_dev, int64_t workspace_bytes,
                                 void* stream);

/*
 * Fused extractor for SELD_MODE_FOA_TF plans (reference data_loader.py:310-349, consumed by train.py:210-261 get_tdm_dataset):
 * ceil(n_samples / hop) frames per clip, frame t = samples [t * hop, t * hop + n_fft) with zeros past the end; rows >= that
 * count are zero; the log-mel block is un-clamped (pass the keys to seld_finalize, top_db = 80), -inf where the mel sum is 0.
 */
SELD_API int seld_extract_tf(seld_plan_t plan, const float* wav_dev, int layout, int n_clips, int64_t n_samples, int t_out,
                             float* feat_raw_dev, uint32_t* clip_max_key_dev, void* stream);

SELD_API int seld_clip_max_decode(const uint32_t* clip_max_key_dev, int n_clips, float* clip_max_dev, void* stream);

/*
 * top_db clamp (torchaudio amplitude_to_DB top_db=80, reference feature_extractor.py:65-71) fused with the
 * per-file normaliser of reference feature_extractor.py:226-234:
 *     y = max(x, clip_max - top_db)   on the 4 log-mel channels of rows < t_valid
 *     y = (y - mean) / max(std, eps)  when mean_dev/std_dev are non-NULL ([n_mels][C] float32)
 * clip_max_key_dev may be NULL (no clamp: input already clamped).  feat_out_dev may alias feat_in_dev.
 */
SELD_API int seld_finalize(int n_mels, int n_ch, const float* feat_in_dev, const uint32_t* clip_max_key_dev, int n_clips, int t_out,
                  int t_valid, float top_db, const float* mean_dev, const float* std_dev, float eps,
                  float* feat_out_dev, void* stream);

/*
 * Per-(mel, chan) partial statistics of reference feature_extractor.py:218-223 (calculate_statistics), with the
 * top_db clamp applied on the fly, accumulated in float64 in a fixed order (run-to-run deterministic):
 *     acc_dev[0 .. n)      += sum x        n = n_mels * C
 *     acc_dev[n .. 2n)     += sum x^2
 *     acc_dev[2n]          += number of rows (n_clips * t_out)
 * acc_dev is ADDED to (zero it first); the caller all-reduces acc_dev across ranks (NCCL sum) and then calls
 * seld_stats_finish.  workspace_dev: at least seld_stats_workspace_doubles(n_mels, n_ch) doubles.
 * n_ch = C (7 | 10 | any layout whose first 4 channels per mel are the log-mel ones).
 */
SELD_API int64_t seld_stats_workspace_doubles(int n_mels, int n_ch);
SELD_API int seld_stats(int n_mels, int n_ch, const float* feat_dev, const uint32_t* clip_max_key_dev, int n_clips, int t_out,
               int t_valid, float top_db, double* workspace_dev, double* acc_dev, void* stream);
/* mean = sum/n_rows, std = sqrt(max(sumsq/n_rows - mean^2, 0)) (population, ddof 0) -> float32 [n_mels][C]. */
/*
 * The statistics all-reduce as ONE kernel over NVLink peer memory instead of a library collective (<= 1 281 doubles,
 * latency bound).  Every rank owns an exchange buffer of seld_stats_peer_buffer_bytes(n_values) bytes, zero-initialised once, that
 * ALL ranks can address (CUDA IPC / torch symmetric memory); peer_base_dev[world] holds the base address of every rank's buffer as
 * seen from this rank.  acc_dev[n_values] (n_values = 2 * n_mels * C + 1): in = this rank's sums, out = the sums over all ranks, added
 * in rank order on every rank (bit-identical everywhere).  All ranks must enqueue the call the same number of times (it contains
 * a cross-GPU barrier); the epoch counter lives in the buffer, so the launch can sit in a replayed CUDA graph.
 */
SELD_API int64_t seld_stats_peer_buffer_bytes(int n_values);
SELD_API int seld_stats_peer_allreduce(const uint64_t* peer_base_dev, int rank, int world, int n_values, double* acc_dev, void* stream);

SELD_API int seld_stats_finish(int n_mels, int n_ch, const double* acc_dev, float* mean_dev, float* std_dev, void* stream);

/*
 * Fused time / frequency spectrogram masking, in place: reference transforms.py:6-43 (mask, period-wise) and
 * :46-75 (simple_mask, one chunk).  x_dev is viewed as x[n_samples][t][mid][f][c] elements of `dtype`:
 *   time masks   zero bands of rows inside each chunk of `period` rows of t   (total = period)
 *   freq masks   zero bands of the f axis, drawn independently per chunk       (total = f)
 * (any axis of any-rank input folds onto this view: axis 0 -> time; axis a > 0 -> mid = prod(shape[1:a]),
 * f = shape[a], c = prod(shape[a+1:]);  simple_mask -> period <= 0, meaning one chunk spanning all of t.)
 * Masked elements are MULTIPLIED by zero (float: -0.0 / NaN survive, as in the reference's x * mask); only
 * masked elements are read or written.  Per chunk and mask: size in [0, max), offset in [0, total - size),
 * time masks first, then freq masks (list order of reference train.py:157-160).  *_max <= 0 means "None"
 * (= total); *_n == 0 disables that axis.  Fails with SELD_EINVAL when t % period != 0 (transforms.py:38-39).
 *   rng_mode PHILOX_COUNTER   Philox4x32-10, key = seed, counter = (lo32(s), hi32(s), chunk,
 *                             axis << 24 | mask << 1 | draw) with s = sample_offset + sample, axis 0 = time,
 *                             1 = freq, draw 0 = size, 1 = offset
 *   rng_mode TF_EAGER_COMPAT  TensorFlow-2 eager stream: `seed` = kernel seed (graph seed mod 2^31-1),
 *                             op_seed2_dev[((sample * n_chunks + chunk) * n_masks + mask) * 2 + draw] = the op's
 *                             seed2, generated on the host in draw order
 *   draws_out_dev (nullable)  [n_samples][n_chunks][time_n + freq_n][2] int32 (offset, size)
 */
SELD_API int seld_mask(void* x_dev, int dtype, int64_t n_samples, int64_t t, int64_t mid, int64_t f, int64_t c, int period,
              int time_max, int time_n, int freq_max, int freq_n, uint64_t seed, uint64_t sample_offset, int rng_mode,
              const int64_t* op_seed2_dev, int32_t* draws_out_dev, void* stream);

/*
 * Per-sample channel gather + sign flip, out of place: the data movement of the reference's batch-level spatial
 * augmentations (foa_intensity_vec_aug transforms.py:78-114, acs_aug :155-199).  in/out are float32
 * [n_samples][outer][n_chan][inner]:  out[b, o, c, j] = (table[b, c] < 0 ? -1 : 1) * in[b, o, table[b, c] & 0xff, j].
 * Features [B, T, F, C]: outer = T*F, inner = 1.  Label coordinates [B, T, 4, n_classes]: outer = T, n_chan = 4,
 * inner = n_classes.  table_dev int32 [n_samples][n_chan] (source channel in the low byte, bit 31 = negate);
 * n_chan <= 32, n_samples <= 65535, fewer than 2^31 elements per sample, in_dev != out_dev.
 */
SELD_API int seld_channel_remap(const float* in_dev, float* out_dev, int64_t n_samples, int64_t outer, int n_chan, int64_t inner,
                                const int32_t* table_dev, void* stream);

/*
 * Per-sample level jitter (reference trainv2.py:120-124, `random_ups_and_downs`: one N(0, 0.2^2) scalar added to the four
 * log-mel channels of a sample): out[b, p, c] = in[b, p, c] + (c < n_first ? offset[b] : 0) for float32
 * [n_samples][positions][n_chan]; offset_dev float32 [n_samples].  in_dev == out_dev is allowed.
 */
SELD_API int seld_channel_offset(const float* in_dev, float* out_dev, int64_t n_samples, int64_t positions, int n_chan, int n_first,
                                 const float* offset_dev, void* stream);

/*
 * Fused training-batch augmentation (one read + one write of the batch; draws made on the device, no host random numbers):
 *   level jitter   reference trainv2.py:120-124 random_ups_and_downs: N(0, level_stddev^2) per sample on channels [:4] (0 = off)
 *   spatial        0 none | 1 foa_intensity_vec_aug (transforms.py:78-114, n_chan 7) | 2 acs_aug (:155-199, n_chan 17),
 *                  applied consistently to the label coordinates y[n][t_y][4][n_classes] (y_in_dev / y_out_dev may be NULL)
 *   masks          reference transforms.py:6-43 per `period`-frame chunk: time_n bands of < time_max frames, freq_n bands of
 *                  < freq_max bins -- the same bands seld_mask draws for (seed, sample_offset) in PHILOX_COUNTER mode
 * x_out[b,t,f,c] = keep(t,f) * sign_b[c] * (x_in[b,t,f,src_b[c]] + (src_b[c] < 4 ? offset_b : 0)).  Out of place when
 * spatial != 0.  draws_out_dev (nullable) [n][2] int32: packed spatial draw (IV: flip bits 0..2, swap bit 3; ACS: swap
 * index), level offset as float bits.  Streams: (seed, sample_offset + b, 0x100 | 0x101 | 0x102) as in transforms.py.
 */
SELD_API int seld_augment_batch(const float* x_in_dev, float* x_out_dev, int64_t n_samples, int64_t t, int64_t f, int n_chan,
                                const float* y_in_dev, float* y_out_dev, int64_t t_y, int n_classes, int spatial, float level_stddev,
                                int period, int time_max, int time_n, int freq_max, int freq_n, uint64_t seed, uint64_t sample_offset,
                                int32_t* draws_out_dev, void* stream);

/*
 * Stand-alone stages (API parity with the reference's public helpers; the hot path is seld_extract).
 *   seld_complex_spec     reference feature_extractor.py:153-173; spec_dev [n_chan][T][F] complex64 (frame-major;
 *                         the Python wrapper returns the [C, F, T] transposed view)
 *   seld_foa_iv           reference feature_extractor.py:176-193; spec [4][n] complex64 -> iv [3][n]
 *   seld_gcc              reference feature_extractor.py:196-214; spec [n_chan][T][F] complex64 (frame-major)
 *                         -> gcc [pairs][n_lags][T] float32, row j = lag first_lag + j of the length-2(F-1) irfft
 */
SELD_API int seld_complex_spec(seld_plan_t plan, const float* wav_dev, int n_chan, int64_t n_samples, float scale,
                      float* spec_dev, void* stream);
SELD_API int seld_foa_iv(const float* spec_dev, int64_t n, float eps, float* iv_dev, void* stream);
SELD_API int seld_gcc(const float* spec_dev, int n_chan, int64_t n_frames, int n_bins, int n_lags, int first_lag, float* gcc_dev,
             void* stream);

/*
 * GCC-PHAT lag projection on the tensor cores (tcgen05.mma, FP16 operands, FP32 accumulation in TMEM): the pruned
 * inverse transform of reference feature_extractor.py:210-211 as a dense contraction,
 *     out[rows][64] = scale * A[rows][1024] * B[1024][64]
 * A: unit phasors, row = (frame, pair), K order (Re P[0], Re P[512], Re P[1], Im P[1], ..., Re P[511], Im P[511]);
 * B: the matching inverse-DFT basis (seld_b200.tables.gcc_basis).  Both operands are passed as UMMA operand IMAGES
 * (seld_b200.tables.gcc_operand_image): a_img_dev [ceil(rows/128)][16 chunks][16 KB], bt_img_dev [16 chunks][8 KB],
 * __half, 16-byte aligned -- the layout the kernel bulk-copies straight into shared memory.  The fused MIC extractor
 * writes that image itself and drives the same kernel with an epilogue that assembles complete feature rows; this
 * entry point exists for tests and for callers that hold their own phasors.
 */
SELD_API int seld_gcc_gemm(const void* a_img_dev, const void* bt_img_dev, int64_t rows, float scale, float* out_dev,
                           void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SELD_B200_H */
