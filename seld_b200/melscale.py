"""HTK triangular mel filterbank, float32, bit-identical to torchaudio's table.

The reference builds ``torchaudio.transforms.MelScale(n_mels, sample_rate)`` for
every file (reference feature_extractor.py:59-60) and applies it to the power
spectrogram and to the normalised intensity vectors (:64, :76).  The log-mel
tolerance (1e-4 dB) needs the *same float32 table* (SURVEY.md section 7.2-2:
recomputing it in float64 alone costs 9.2e-5 dB), so the arithmetic below
follows torchaudio's published ``melscale_fbanks(..., norm=None,
mel_scale='htk')`` step by step in float32 torch ops.  It is a constant built
once per plan on the host; tests check it bit-for-bit against torchaudio.

The table is also converted to the sparse form the kernels use: every STFT bin
lies between two adjacent filter centres, so a row of the dense table has at
most two non-zeros, in adjacent filters ``seg`` and ``seg + 1``.
"""
import math

import numpy as np
import torch


def melscale_fbanks_htk(n_freqs: int, sample_rate: int, n_mels: int,
                        f_min: float = 0.0, f_max: float = None) -> torch.Tensor:
    """Dense ``[n_freqs, n_mels]`` float32 filterbank (torchaudio MelScale defaults)."""
    if f_max is None:
        f_max = float(sample_rate // 2)
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + (f_min / 700.0))
    m_max = 2595.0 * math.log10(1.0 + (f_max / 700.0))
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down, up))


def sparsify(fb) -> tuple:
    """Dense ``[F, M]`` table -> ``(seg int32[F], w0 float32[F], w1 float32[F])``.

    Row ``k`` contributes ``w0[k]`` to filter ``seg[k]`` and ``w1[k]`` to filter
    ``seg[k] + 1``; ``seg[k] == -1`` marks an all-zero row.  Raises if a row
    has more than two non-zeros or two non-adjacent ones (cannot happen for a
    triangular bank with increasing centres; checked because the kernels rely on it).
    """
    fb = np.asarray(fb, dtype=np.float32)
    n_freqs, n_mels = fb.shape
    seg = np.full(n_freqs, -1, dtype=np.int32)
    w0 = np.zeros(n_freqs, dtype=np.float32)
    w1 = np.zeros(n_freqs, dtype=np.float32)
    for k in range(n_freqs):
        nz = np.flatnonzero(fb[k])
        if nz.size == 0:
            continue
        if nz.size > 2 or (nz.size == 2 and nz[1] != nz[0] + 1):
            raise ValueError(f'mel filterbank row {k} is not a pair of adjacent taps: {nz}')
        seg[k] = nz[0]
        w0[k] = fb[k, nz[0]]
        if nz.size == 2:
            w1[k] = fb[k, nz[1]]
    return seg, w0, w1
