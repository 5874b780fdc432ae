"""Host-side constant tables of an extraction plan (window, twiddles, sparse mel bank).

These are the constants the reference rebuilds for every file (Hann window:
reference feature_extractor.py:167; MelScale: :59-60).  Here they are built once per
plan on the host and uploaded by ``seld_plan_create``.
"""
import numpy as np
import torch

from .melscale import melscale_fbanks_htk, sparsify

SUPPORTED_N_FFT = (256, 512, 1024, 2048)


def resolve_stft(n_fft=512, win_length=None, hop_length=None):
    """Default rule of reference feature_extractor.py:160-163."""
    if win_length is None:
        win_length = n_fft
    if hop_length is None:
        hop_length = win_length // 2
    return int(n_fft), int(win_length), int(hop_length)


def padded_window(n_fft: int, win_length: int) -> np.ndarray:
    """float32 periodic Hann(win_length) placed in the centre of an n_fft frame (torch.stft rule)."""
    if not 0 < win_length <= n_fft:
        raise ValueError('win_length must be in (0, n_fft]')
    w = np.zeros(n_fft, dtype=np.float32)
    left = (n_fft - win_length) // 2
    w[left:left + win_length] = torch.hann_window(win_length, dtype=torch.float32).numpy()
    return w


def twiddles(n_fft: int) -> np.ndarray:
    """[n_fft, 2] float32 = exp(-2 pi i j / n_fft), evaluated in float64."""
    j = np.arange(n_fft, dtype=np.float64)
    ang = -2.0 * np.pi * j / n_fft
    return np.stack([np.cos(ang), np.sin(ang)], axis=1).astype(np.float32)


def mel_tables(n_fft: int, sample_rate: int, n_mels: int):
    fb = melscale_fbanks_htk(n_fft // 2 + 1, int(sample_rate), int(n_mels))
    seg, w0, w1 = sparsify(fb.numpy())
    return fb, seg, w0, w1


GCC_BASIS_SCALE = 512.0      # the fp16 basis is stored x512 (entries in [-1, 1]); the GEMM epilogue multiplies by 1/512


def gcc_basis(n_fft: int = 1024, n_lags: int = 64) -> np.ndarray:
    """float16 [n_lags, n_fft] basis B^T of the pruned inverse real FFT behind GCC-PHAT (reference
    feature_extractor.py:210-211): output j is lag j - n_lags/2 of irfft(P, n_fft), written as a contraction over
    K = (Re P[0], Re P[n/2], Re P[1], Im P[1], ..., Re P[n/2-1], Im P[n/2-1]); scaled by GCC_BASIS_SCALE."""
    half = n_fft // 2
    lags = np.arange(n_lags, dtype=np.float64) - n_lags // 2
    bt = np.zeros((n_lags, n_fft), dtype=np.float64)
    bt[:, 0] = 1.0 / n_fft                                        # DC term
    bt[:, 1] = np.where(lags.astype(np.int64) % 2 == 0, 1.0, -1.0) / n_fft      # Nyquist term: cos(pi * lag)
    k = np.arange(1, half, dtype=np.float64)
    ang = 2.0 * np.pi * np.outer(lags, k) / n_fft
    bt[:, 2::2] = 2.0 * np.cos(ang) / n_fft
    bt[:, 3::2] = -2.0 * np.sin(ang) / n_fft
    return (bt * GCC_BASIS_SCALE).astype(np.float16)


def gcc_operand_image(mat: np.ndarray) -> np.ndarray:
    """Row-major float16 [rows, 1024] -> the UMMA operand image seld_gcc_gemm consumes (K-major SWIZZLE_128B): rows are
    padded to a multiple of 128 for the A operand (rows == 64: the B^T operand, one 64-row block per chunk); per block
    of rows, 16 chunks of 64 K-elements; inside a chunk row r is 128 contiguous bytes whose 16-byte unit u is stored at
    unit position u ^ (r % 8).  (The fused extractor writes this layout directly.)"""
    mat = np.ascontiguousarray(mat, dtype=np.float16)
    rows, k = mat.shape
    assert k == 1024
    block = 64 if rows == 64 else 128
    pad = (-rows) % block
    if pad:
        mat = np.concatenate([mat, np.zeros((pad, 1024), np.float16)], 0)
    n_blocks = mat.shape[0] // block
    x = mat.reshape(n_blocks, block, 16, 8, 8)                  # [block][row][chunk][unit][e]
    out = np.empty_like(x)
    r = np.arange(block)
    for u in range(8):
        out[:, r, :, u ^ (r % 8), :] = x[:, r, :, u, :]
    out = out.transpose(0, 2, 1, 3, 4)                          # [block][chunk][row][unit][e]
    return np.ascontiguousarray(out if rows != 64 else out[0])


# --------------------------------------------------------------------------- TensorFlow variant (reference data_loader.py:310-349)
def tf_hann_window(n: int) -> np.ndarray:
    """tf.signal.hann_window(n, periodic=True) in float32 arithmetic: 0.5 - 0.5 cos(2 pi k / (n + even - 1))."""
    even = 1 - n % 2
    count = np.arange(n).astype(np.float32)
    arg = np.float32(2 * np.pi) * count / np.float32(n + even - 1)
    return (np.float32(0.5) - np.float32(0.5) * np.cos(arg)).astype(np.float32)


def tf_mel_weight_matrix(n_mels: int, n_bins: int, sample_rate: int, lower_hz: float = 0.0, upper_hz=None) -> np.ndarray:
    """tf.signal.linear_to_mel_weight_matrix(n_mels, n_bins, sample_rate, lower_hz, upper_hz) in float32 arithmetic:
    HTK mel = 1127 ln(1 + f / 700); triangles linear in MEL between linspace(mel(lo), mel(hi), n_mels + 2); the DC row
    is zero.  Every row has at most two adjacent non-zeros, which is what the extractor's piece form needs."""
    f32 = np.float32
    upper_hz = sample_rate // 2 if upper_hz is None else upper_hz
    mel = lambda hz: (f32(1127.0) * np.log(f32(1.0) + hz / f32(700.0))).astype(f32)      # noqa: E731
    lin = np.linspace(f32(0.0), f32(sample_rate) / f32(2.0), n_bins, dtype=f32)[1:]
    bins_mel = mel(lin)[:, None]
    edges = np.linspace(mel(np.asarray(f32(lower_hz))), mel(np.asarray(f32(upper_hz))), n_mels + 2, dtype=f32)
    lower, center, upper = edges[None, :-2], edges[None, 1:-1], edges[None, 2:]
    w = np.maximum(f32(0.0), np.minimum((bins_mel - lower) / (center - lower), (upper - bins_mel) / (upper - center)))
    return np.concatenate([np.zeros((1, n_mels), dtype=f32), w.astype(f32)], 0)
