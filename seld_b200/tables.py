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
