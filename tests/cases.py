"""Shared definitions of the golden cases (must match oracle/make_golden.py)."""
import hashlib
import os

import numpy as np
import torch

from seld_b200.synth import make_clip

PROD = dict(win_length=960, hop_length=480, n_fft=1024)

CASES = {
    'prod':     (1000, 36000, 24000, 64, PROD),
    'prodB':    (2000, 36000, 24000, 64, PROD),
    'zeros':    (None, 32000, 16000, 64, {}),
    'default':  (7,    8000,  16000, 64, {}),
    'ragged':   (11,   30007, 24000, 64, PROD),
    'nfft256':  (13,   5000,  8000,  32, dict(n_fft=256)),
    'nfft2048': (17,   20000, 48000, 64, dict(n_fft=2048, win_length=1200, hop_length=600)),
    'loud':     (19,   12000, 24000, 64, PROD),
}

# stated tolerances (BASELINE.json north_star): max abs error
TOL_LOGMEL_DB = 1e-4
TOL_IV = 1e-3
TOL_GCC = 1e-3


def case_input(name):
    seed, n, sr, n_mels, kw = CASES[name]
    if seed is None:
        wav = torch.zeros(4, n)
    else:
        wav = make_clip(seed, n, sr)
        if name == 'loud':
            wav = wav * 300.0
    return wav, sr, n_mels, dict(kw)


def load_golden(name):
    here = os.path.dirname(os.path.abspath(__file__))
    return np.load(os.path.join(here, 'golden', f'extract_{name}.npz'))


def input_matches_golden(wav, golden) -> bool:
    sha = hashlib.sha256(np.ascontiguousarray(wav.numpy()).tobytes()).hexdigest()
    return sha == str(golden['input_sha256'])


def check_features(got, want, mode, what=''):
    """Assert the stated tolerances channel block by channel block; returns the measured errors."""
    assert got.shape == want.shape, (what, got.shape, want.shape)
    assert got.dtype == np.float32
    e_mel = float(np.abs(got[..., :4].astype(np.float64) - want[..., :4]).max())
    e_rest = float(np.abs(got[..., 4:].astype(np.float64) - want[..., 4:]).max())
    assert e_mel <= TOL_LOGMEL_DB, f'{what} log-mel max abs err {e_mel:.3e} dB > {TOL_LOGMEL_DB}'
    tol = TOL_IV if mode == 'foa' else TOL_GCC
    assert e_rest <= tol, f'{what} {"IV" if mode == "foa" else "GCC"} max abs err {e_rest:.3e} > {tol}'
    return e_mel, e_rest
