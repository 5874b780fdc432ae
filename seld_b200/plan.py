"""Extraction plans: one per (sample_rate, STFT geometry, n_mels, mode, device), cached.

A plan owns the device copies of the constants that the reference rebuilds for every file
(reference feature_extractor.py:59-60 MelScale, :167 Hann window).
"""
import ctypes
import math
import threading

import numpy as np
import torch

from . import _lib, tables

_MODES = {'foa': _lib.MODE_FOA, 'mic': _lib.MODE_MIC}
VARIANTS = ('torch', 'tf')
_cache = {}
_cache_lock = threading.Lock()


class ExtractPlan:
    def __init__(self, sample_rate, n_fft, win_length, hop_length, n_mels, mode, normalized=False, variant='torch'):
        if mode not in _MODES:
            raise ValueError('invalid mode')                       # reference feature_extractor.py:81-82
        if variant not in VARIANTS:
            raise ValueError('variant must be "torch" or "tf"')
        if variant == 'tf' and (mode != 'foa' or n_fft != 1024 or win_length != n_fft or normalized):
            # reference data_loader.py:310-312 defaults; its mode='mic' branch cannot run (gcc_features_tf slices the time axis
            # and the concat with the log-mel block fails on the shape)
            raise ValueError('the TF variant is built for mode="foa", win_length = n_fft = 1024')
        if n_fft not in tables.SUPPORTED_N_FFT:
            raise ValueError(f'n_fft must be one of {tables.SUPPORTED_N_FFT} (got {n_fft})')
        _lib.require_device()
        self.sample_rate, self.n_fft, self.win_length, self.hop_length = int(sample_rate), n_fft, win_length, hop_length
        self.n_mels, self.mode, self.normalized = int(n_mels), mode, bool(normalized)
        self.n_bins = n_fft // 2 + 1
        self.n_out_ch = 7 if mode == 'foa' else 10
        self.variant = variant
        window = tables.padded_window(n_fft, win_length) if variant == 'torch' else tables.tf_hann_window(win_length)
        # spectrogram(normalized=True) divides the spectrum by sqrt(sum w^2); folding it into the window is the same map
        self.spec_scale = 1.0 / math.sqrt(float(torch.hann_window(win_length).pow(2.).sum())) if normalized else 1.0
        if normalized:
            window = (window * np.float32(self.spec_scale)).astype(np.float32)
        if variant == 'torch':
            fb = tables.melscale_fbanks_htk(self.n_bins, self.sample_rate, self.n_mels).numpy()
        else:
            fb = tables.tf_mel_weight_matrix(self.n_mels, self.n_bins, self.sample_rate, 0.0, self.sample_rate // 2)
        self._window = np.ascontiguousarray(window, dtype=np.float32)
        self._fb = np.ascontiguousarray(fb, dtype=np.float32)
        handle = ctypes.c_void_p()
        lib = _lib.load()
        mode_code = _lib.MODE_FOA_TF if variant == 'tf' else _MODES[mode]
        _lib.check(lib.seld_plan_create(self.sample_rate, n_fft, win_length, hop_length, self.n_mels, 4, mode_code,
                                        self._window.ctypes.data_as(ctypes.c_void_p),
                                        self._fb.ctypes.data_as(ctypes.c_void_p), ctypes.byref(handle)))
        self.handle = handle
        self.device = torch.cuda.current_device()

    def num_frames(self, n_samples: int) -> int:
        if self.variant == 'tf':                                   # tf.signal.stft(pad_end=True)
            return -(-int(n_samples) // self.hop_length)
        return 1 + int(n_samples) // self.hop_length

    def __del__(self):
        try:
            if getattr(self, 'handle', None):
                _lib.load().seld_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def get_plan(sample_rate, mode='foa', n_mels=64, pad=0, n_fft=512, win_length=None, hop_length=None,
             normalized=False, variant='torch') -> ExtractPlan:
    """Cached plan for the keyword set of reference complex_spec (feature_extractor.py:153-158); `pad` is handled
    by the caller (zero-padding the waveform)."""
    n_fft, win_length, hop_length = tables.resolve_stft(n_fft, win_length, hop_length)
    _lib.require_device()           # cached per device after the first call
    key = (int(sample_rate), n_fft, win_length, hop_length, int(n_mels), mode, bool(normalized), variant, torch.cuda.current_device())
    with _cache_lock:
        plan = _cache.get(key)
        if plan is None:
            plan = ExtractPlan(sample_rate, n_fft, win_length, hop_length, n_mels, mode, normalized, variant)
            _cache[key] = plan
    return plan
