"""TEST INFRASTRUCTURE -- CPU restatement of the reference's TensorFlow variant of the on-the-fly extractor
(reference data_loader.py:237-265 foa_intensity_vectors_tf / gcc_features_tf, :310-349 get_preprocessed_x_tf, :188-234 TDM_aug).

PARITY UNPINNED: TensorFlow and tensorflow_io are pip dependencies of the reference (requirements.txt) that are absent
from this image and from /root/reference, and the reference has no test for these functions.  What follows restates the
PUBLISHED algorithms those calls bottom out in:

  tf.signal.hann_window(n, periodic=True)        0.5 - 0.5 cos(2 pi k / n) for even n, evaluated in float32
  tf.signal.stft(x, frame_length, frame_step, fft_length, pad_end=True)
                                                 ceil(L / step) frames starting at t * step, zero-padded past the end,
                                                 windowed, rfft(fft_length)
  tf.signal.linear_to_mel_weight_matrix          HTK mel = 1127 ln(1 + f / 700); triangles LINEAR IN MEL between
                                                 linspace(mel(lo), mel(hi), n_mels + 2); DC row zero; float32
  tfio.experimental.audio.dbscale(x, top_db)     10 log10(x^2) (no floor), then max(., global max - top_db)

Only tests/ and bench.py's CPU legs may import this module.
"""
import numpy as np

_MEL_BREAK, _MEL_Q = 700.0, 1127.0


def hann_window_tf(n, dtype=np.float32):
    even = 1 - n % 2
    denom = dtype(n + even - 1)
    count = np.arange(n).astype(dtype)
    arg = dtype(2 * np.pi) * count / denom
    return (dtype(0.5) - dtype(0.5) * np.cos(arg)).astype(dtype)


def linear_to_mel_weight_matrix_tf(n_mels=64, n_bins=513, sample_rate=24000, lower_hz=0.0, upper_hz=None, dtype=np.float32):
    """[n_bins, n_mels] in `dtype` arithmetic (float32 = TF's default dtype; float64 = the exact table)."""
    upper_hz = sample_rate // 2 if upper_hz is None else upper_hz
    f = dtype
    mel = lambda hz: f(_MEL_Q) * np.log(f(1.0) + hz / f(_MEL_BREAK)).astype(dtype)      # noqa: E731
    nyq = f(sample_rate) / f(2.0)
    lin = np.linspace(f(0.0), nyq, n_bins, dtype=dtype)[1:]
    bins_mel = mel(lin)[:, None]
    edges = np.linspace(mel(np.asarray(f(lower_hz))), mel(np.asarray(f(upper_hz))), n_mels + 2, dtype=dtype)
    lower, center, upper = edges[None, :-2], edges[None, 1:-1], edges[None, 2:]
    lower_slopes = (bins_mel - lower) / (center - lower)
    upper_slopes = (upper - bins_mel) / (upper - center)
    w = np.maximum(f(0.0), np.minimum(lower_slopes, upper_slopes)).astype(dtype)
    return np.concatenate([np.zeros((1, n_mels), dtype=dtype), w], 0)


def stft_tf(wav, frame_length=1024, frame_step=480, fft_length=1024, dtype=np.float64):
    """wav [C, L] -> complex [C, T, fft_length // 2 + 1], T = ceil(L / frame_step) (pad_end=True)."""
    wav = np.asarray(wav)
    n_ch, n = wav.shape
    t = -(-n // frame_step)
    padded = np.zeros((n_ch, (t - 1) * frame_step + frame_length), dtype=dtype)
    padded[:, :n] = wav
    idx = np.arange(t)[:, None] * frame_step + np.arange(frame_length)[None, :]
    frames = padded[:, idx] * hann_window_tf(frame_length).astype(dtype)[None, None, :]
    return np.fft.rfft(frames, n=fft_length, axis=-1)


def foa_intensity_vectors_tf(spec, eps=1e-8):
    """reference data_loader.py:237-251; spec [4, T, F] complex -> [3, T, F]."""
    c0 = np.conj(spec[0])
    iv = np.stack([(c0 * spec[3]).real, (c0 * spec[1]).real, (c0 * spec[2]).real], 0)
    norm = np.maximum(np.sqrt((iv ** 2).sum(0)), eps)
    return iv / norm


def dbscale_tfio(x, top_db=80.0):
    with np.errstate(divide='ignore'):
        log_spec = 10.0 * (np.log(np.square(x)) / np.log(10.0))
    return np.maximum(log_spec, log_spec.max() - top_db)


def get_preprocessed_x_tf(wav, sr, mode='foa', n_mels=64, multiplier=5, max_label_length=600, win_length=1024, hop_length=480,
                          n_fft=1024, mel_dtype=np.float32):
    """reference data_loader.py:310-349 for mode='foa' -> float64 [max_len, n_mels, 7].  (mode='mic' cannot run in the
    reference: gcc_features_tf slices the TIME axis -- cc[-n_mels//2:] on a [T, n_fft] tensor -- and the concat with the
    [4, T, n_mels] log-mel block then fails on the shape.)"""
    if mode != 'foa':
        raise ValueError('invalid mode')
    mel_mat = linear_to_mel_weight_matrix_tf(n_mels, n_fft // 2 + 1, sr, 0.0, sr // 2, dtype=mel_dtype).astype(np.float64)
    spec = stft_tf(wav, win_length, hop_length, n_fft)
    mel_spec = dbscale_tfio(np.abs(spec) @ mel_mat, top_db=80.0)
    foa = foa_intensity_vectors_tf(spec) @ mel_mat
    feat = np.concatenate([mel_spec, foa], 0).transpose(1, 2, 0)
    max_len = max_label_length * multiplier
    if feat.shape[0] < max_len:
        feat = np.pad(feat, ((0, max_len - feat.shape[0]), (0, 0), (0, 0)))
    return feat[:max_len]


def gcc_features_tf(spec, n_mels):
    """reference data_loader.py:254-265, literally: irfft over the last (frequency) axis of [chan, T, F], then the slice on
    axis 0 of the [T, n_fft] result -- i.e. the last n_mels/2 and the first (n_mels+1)/2 FRAMES -> [pairs, n_mels, n_fft]."""
    n_chan = spec.shape[0]
    out = []
    for m in range(n_chan):
        for n in range(m + 1, n_chan):
            r = np.conj(spec[m]) * spec[n]
            cc = np.fft.irfft(np.exp(1j * np.angle(r)), axis=-1)
            out.append(np.concatenate([cc[-(n_mels // 2):], cc[:(n_mels + 1) // 2]], 0))
    return np.stack(out, 0)


def tdm_aug(x, y, tdm_x, tdm_y, draws, sr=24000, label_resolution=0.1, max_overlap_per_frame=2):
    """reference data_loader.py:188-234 with the random draws made explicit: draws[i] = list of (cls, sample_time, offset,
    td_offset) in label frames.  x: list of [4, L] arrays, y: list of [T_y, 4 * n_classes]; both are modified in place
    (as in the reference) and returned."""
    n_cls = y[0].shape[-1] // 4
    spf = int(sr * label_resolution)
    for i in range(len(x)):
        for cls, st, off, tdo in draws[i]:
            frame_y = y[i][off:off + st]
            nondup = 1 - frame_y[..., cls]
            valid = (frame_y[..., :n_cls].sum(-1) < max_overlap_per_frame).astype(frame_y.dtype) * nondup
            if valid.sum() == 0:
                continue
            y[i][off:off + st] += tdm_y[cls][tdo:tdo + st] * valid[:, None]
            x[i][:, off * spf:(off + st) * spf] += tdm_x[cls][:, tdo * spf:(tdo + st) * spf] * np.repeat(valid, spf)[None, :].astype(x[i].dtype)
    return x, y
