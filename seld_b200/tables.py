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
