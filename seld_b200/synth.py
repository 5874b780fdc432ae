"""Synthetic 4-channel clips of the DCASE dev-set shape.

``float32[4, L]`` = 0.1 N(0,1) + three sinusoids (440 / 1750 / 6000 Hz, amp 0.05)
with per-channel integer delays in 0..8 samples; the first 10 s scaled by 1e-6
(drives log-mel more than 80 dB below the clip maximum, so the ``top_db`` clamp
of reference feature_extractor.py:65-71 is active); the following 1 s is exact
zeros (exercises ``amin``, the IV ``eps`` and ``angle(0) == 0``).
"""
import math

import torch

TONES_HZ = (440.0, 1750.0, 6000.0)


def clip_delays(seed: int):
    return [(seed * 7 + 3 * c) % 9 for c in range(4)]


def make_clip(seed: int, n_samples: int = 1_440_000, sample_rate: int = 24000,
              device='cpu', quiet_frac: float = 1.0 / 6.0, zero_frac: float = 1.0 / 60.0) -> torch.Tensor:
    """One clip; the generator is seeded on ``device`` (CPU clips are the parity inputs)."""
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    wav = 0.1 * torch.randn(4, n_samples, generator=g, device=dev, dtype=torch.float32)
    t = torch.arange(n_samples, device=dev, dtype=torch.float64)
    for c, d in enumerate(clip_delays(seed)):
        for f in TONES_HZ:
            wav[c] += (0.05 * torch.sin(2.0 * math.pi * f * (t - d) / sample_rate)).float()
    n_quiet = int(n_samples * quiet_frac)
    n_zero = int(n_samples * zero_frac)
    wav[:, :n_quiet] *= 1e-6
    wav[:, n_quiet:n_quiet + n_zero] = 0.0
    return wav.clamp_(-0.999, 0.999)


def make_clips(seeds, n_samples: int = 1_440_000, sample_rate: int = 24000, device='cpu') -> torch.Tensor:
    """``[len(seeds), 4, n_samples]`` float32."""
    seeds = list(seeds)
    out = torch.empty(len(seeds), 4, n_samples, dtype=torch.float32, device=device)
    for i, s in enumerate(seeds):
        out[i] = make_clip(s, n_samples, sample_rate, device)
    return out
