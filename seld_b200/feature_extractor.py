"""Drop-in for the reference's ``feature_extractor`` module, computed on a B200.

Same names, positional order, defaults, error conventions and array layouts as the reference
(feature_extractor.py), so ``from feature_extractor import *`` in the reference's data_loader.py:3 keeps working
when this module is put in front of it on ``sys.path``.  The arithmetic runs in the
sm_100a kernels behind libseld_b200.so; there is no CPU path -- without the library or a B200 every extraction
call raises.
"""
import os
from glob import glob

import numpy as np
import torch

from . import _lib, pipeline
from .data_utils import create_folder, degree_to_radian, radian_to_degree
from .plan import get_plan
from .wavio import load_wav

__all__ = ['extract_seldnet_data', 'extract_features', 'extract_labels', 'preprocess_features_labels', 'complex_spec',
           'foa_intensity_vectors', 'gcc_features', 'calculate_statistics', 'apply_normalizer', 'cartesian_to_polar',
           'polar_to_cartesian', 'get_device', 'create_folder', 'degree_to_radian', 'radian_to_degree']


def get_device():
    """reference utils.py:62-64 -- here a CUDA device is mandatory."""
    _lib.require_device()
    return torch.device('cuda', torch.cuda.current_device())


def _stem(path):
    return os.path.splitext(os.path.basename(path))[0]


# --------------------------------------------------------------------------- dataset driver
def extract_seldnet_data(feature_path: str, feature_output_path: str, label_path: str, label_output_path: str,
                         mode='foa', **kwargs):
    """reference feature_extractor.py:15-50: every ``*.wav`` / ``*.csv`` pair -> ``<name>.npy`` features
    ([3000, 64, C] float32) and labels ([600, 4*n_classes])."""
    if feature_output_path == label_output_path:
        raise ValueError('output folders for features and labels must differ')
    wavs = sorted(glob(os.path.join(feature_path, '*.wav')))
    csvs = sorted(glob(os.path.join(label_path, '*.csv')))
    if len(wavs) != len(csvs):
        raise ValueError('# of features and labels are not matched')
    create_folder(feature_output_path)
    create_folder(label_output_path)
    for wav_path, csv_path in zip(wavs, csvs):
        name = _stem(wav_path)
        if name != _stem(csv_path):
            raise ValueError('feature, label must share the same name')
        wav, rate = load_wav(wav_path)
        feats = extract_features(wav, rate, mode=mode, **kwargs)
        feats, labels = preprocess_features_labels(feats, extract_labels(csv_path))
        np.save(os.path.join(feature_output_path, name + '.npy'), feats)
        np.save(os.path.join(label_output_path, name + '.npy'), labels)


# --------------------------------------------------------------------------- features
def extract_features(wav: torch.Tensor, sample_rate, mode='foa', n_mels=64, **kwargs) -> np.ndarray:
    """reference feature_extractor.py:53-88.  ``wav`` [4, L] float -> float32 ``[1 + L//hop, n_mels, 7 | 10]``
    (log-mel of the 4 channels, then 3 mel-projected intensity vectors or 6 GCC-PHAT pairs)."""
    if mode not in ('foa', 'mic'):
        raise ValueError('invalid mode')
    device = get_device()
    wav = torch.as_tensor(wav)
    if wav.dim() != 2:
        raise ValueError('wav must be [channels, samples]')
    x = wav.to(device=device, dtype=torch.float32).contiguous().unsqueeze(0)
    feat, key = pipeline.extract_batch(x, sample_rate, mode=mode, n_mels=n_mels, **kwargs)
    pipeline.finalize_(feat, key)                 # top_db = 80 against the clip-global maximum (:65-71)
    return feat[0].cpu().numpy()


def complex_spec(wav: torch.Tensor, pad=0, n_fft=512, win_length=None, hop_length=None, normalized=False) -> torch.Tensor:
    """reference feature_extractor.py:153-173.  [C, L] -> complex64 CUDA tensor [C, n_fft//2 + 1, 1 + L//hop]."""
    device = get_device()
    x = torch.as_tensor(wav).to(device=device, dtype=torch.float32)
    lead = x.shape[:-1]
    x = x.reshape(-1, x.shape[-1])
    if pad > 0:
        x = torch.nn.functional.pad(x, (pad, pad))
    x = x.contiguous()
    n_chan, n_samples = x.shape
    # the STFT kernels only need window / twiddles; n_mels and mode are irrelevant here
    plan = get_plan(8000, mode='foa', n_mels=8, n_fft=n_fft, win_length=win_length, hop_length=hop_length,
                    normalized=False)
    t_raw = plan.num_frames(n_samples)
    spec = torch.empty(n_chan, t_raw, plan.n_bins, dtype=torch.complex64, device=device)
    scale = 1.0
    if normalized:
        scale = 1.0 / float(torch.hann_window(plan.win_length).pow(2.).sum().sqrt())
    _lib.check(_lib.load().seld_complex_spec(plan.handle, _lib.ptr(x), n_chan, n_samples, scale, _lib.ptr(spec),
                                             _lib.current_stream_ptr()))
    spec = spec.transpose(-1, -2)                                  # [C, F, T] view of the frame-major buffer
    return spec.reshape(lead + spec.shape[-2:]) if len(lead) != 1 else spec


def _as_complex_cuda(complex_specs):
    x = torch.as_tensor(complex_specs)
    if not torch.is_complex(x):
        x = torch.view_as_complex(x.contiguous())
    return x.to(device=get_device(), dtype=torch.complex64)


def foa_intensity_vectors(complex_specs: torch.Tensor, eps=1e-8) -> torch.Tensor:
    """reference feature_extractor.py:176-193.  [>=4, F, T] complex -> [3, F, T] float32 (x, y, z order)."""
    x = _as_complex_cuda(complex_specs)
    if x.dim() != 3 or x.size(0) < 4:
        raise ValueError('complex_specs must be [chan >= 4, freq, time]')
    x = x[:4].contiguous()
    n = x.shape[1] * x.shape[2]
    out = torch.empty((3,) + tuple(x.shape[1:]), dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().seld_foa_iv(_lib.ptr(x), n, float(eps), _lib.ptr(out), _lib.current_stream_ptr()))
    return out


def gcc_features(complex_specs: torch.Tensor, n_mels: int) -> torch.Tensor:
    """reference feature_extractor.py:196-214.  [C, F, T] complex -> [C(C-1)/2, lags, T] float32; the lags are
    ``cc[-n_mels//2:]`` followed by ``cc[:(n_mels+1)//2]`` of the length-2(F-1) inverse real FFT."""
    x = _as_complex_cuda(complex_specs)
    if x.dim() != 3 or x.size(0) < 2:
        raise ValueError('complex_specs must be [chan >= 2, freq, time]')
    n_chan, n_bins, n_frames = x.shape
    xt = x.transpose(-1, -2).contiguous()                          # frame-major [C, T, F]
    first_lag = -n_mels // 2                                       # python: (-n_mels) // 2, as in the reference slice
    n_lags = -first_lag + (n_mels + 1) // 2
    if -first_lag > 2 * (n_bins - 1) or (n_mels + 1) // 2 > 2 * (n_bins - 1):
        raise ValueError('n_mels exceeds the inverse transform length')
    out = torch.empty(n_chan * (n_chan - 1) // 2, n_lags, n_frames, dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().seld_gcc(_lib.ptr(xt), n_chan, n_frames, n_bins, n_lags, first_lag, _lib.ptr(out),
                                    _lib.current_stream_ptr()))
    return out


# --------------------------------------------------------------------------- labels / padding (host, numpy)
def extract_labels(path: str, n_classes=14, max_frames=None):
    """reference feature_extractor.py:91-114.  DCASE csv rows ``frame, class, track, azimuth, elevation`` ->
    float32 [n_frames, 4*n_classes] = (activity, x, y, z) blocks."""
    rows = []
    with open(path, 'r') as fh:
        for line in fh:
            if not line.strip():
                continue
            frame, cls, _track, azi, ele = (int(v) for v in line.split(','))
            rows.append((frame, cls, azi, ele))
    rows = np.asarray(rows)
    xyz = polar_to_cartesian(rows[:, 2:])
    n_frames = int(rows[:, 0].max()) + 1
    if max_frames is not None:
        n_frames = max(max_frames, n_frames)
    out = np.zeros((n_frames, 4, n_classes), dtype='float32')
    for (frame, cls), pos in zip(rows[:, :2], xyz):
        out[int(frame), 0, int(cls)] = 1.0
        out[int(frame), 1:, int(cls)] = pos
    return out.reshape(n_frames, 4 * n_classes)


def preprocess_features_labels(features: np.ndarray, labels: np.ndarray, max_label_length=600, multiplier=5):
    """reference feature_extractor.py:117-149: zero-pad or cut labels to ``max_label_length`` rows and features to
    ``max_label_length * multiplier`` frames."""
    def fit(a, n):
        if a.shape[0] < n:
            return np.pad(a, [(0, n - a.shape[0])] + [(0, 0)] * (a.ndim - 1), 'constant')
        return a[:n]
    return fit(features, max_label_length * multiplier), fit(labels, max_label_length)


# --------------------------------------------------------------------------- dataset statistics
def _npy_files(folder):
    return sorted(glob(os.path.join(folder, '*.npy')))


def calculate_statistics(feature_path: str):
    """reference feature_extractor.py:218-223: mean / population std over the time axis of ALL files, keepdims
    -> two float32 arrays [1, n_mels, C].  Accumulated in float64 on the GPU (the reference's float32 numpy sum
    is itself off by ~1e-3 at dev-set size)."""
    device = get_device()
    acc = None
    shape = None
    for f in _npy_files(feature_path):
        x = torch.from_numpy(np.ascontiguousarray(np.load(f), dtype=np.float32)).to(device)
        if shape is None:
            shape = tuple(x.shape[1:])
            n_mels, n_ch = (shape if len(shape) == 2 else (int(np.prod(shape)), 1))
        elif tuple(x.shape[1:]) != shape:
            raise ValueError('all feature files must share their trailing dimensions')
        acc = pipeline.partial_statistics(x.reshape(1, x.shape[0], n_mels, n_ch), None, None, acc)
    if acc is None:
        raise ValueError('need at least one array to concatenate')
    mean, std = pipeline.finish_statistics(acc, n_mels, n_ch)
    return mean.cpu().numpy().reshape((1,) + shape), std.cpu().numpy().reshape((1,) + shape)


def apply_normalizer(feature_path, new_feature_path, mean, std, eps=1e-8):
    """reference feature_extractor.py:226-234: ``(x - mean) / max(std, eps)`` per file into ``new_feature_path``."""
    device = get_device()
    create_folder(new_feature_path)
    mean_d = torch.as_tensor(np.asarray(mean, dtype=np.float32)).to(device)
    std_d = torch.as_tensor(np.asarray(std, dtype=np.float32)).to(device)
    for f in _npy_files(feature_path):
        x = torch.from_numpy(np.ascontiguousarray(np.load(f), dtype=np.float32)).to(device)
        shape = tuple(x.shape)
        n_mels, n_ch = (shape[1:] if len(shape) == 3 else (int(np.prod(shape[1:])), 1))
        y = pipeline.finalize_(x.reshape(1, shape[0], n_mels, n_ch), None, None, mean_d, std_d, eps)
        np.save(os.path.join(new_feature_path, os.path.basename(f)), y.reshape(shape).cpu().numpy())


# --------------------------------------------------------------------------- unit conversion (host, numpy)
def cartesian_to_polar(coordinates):
    """reference feature_extractor.py:238-253: [..., (x, y, z)] -> [..., (azimuth deg, elevation deg, r)]."""
    c = np.asarray(coordinates)
    if c.shape[-1] != 3:
        raise ValueError('only 3D cartesian coordinates are allowed')
    x, y, z = c[..., 0], c[..., 1], c[..., 2]
    horiz = np.sqrt(x ** 2 + y ** 2)
    return np.stack([radian_to_degree(np.arctan2(y, x)), radian_to_degree(np.arctan2(z, horiz)),
                     np.sqrt(x ** 2 + y ** 2 + z ** 2)], axis=-1)


def polar_to_cartesian(coordinates):
    """reference feature_extractor.py:256-271: [..., (azimuth deg, elevation deg[, r])] -> [..., (x, y, z)]."""
    c = np.asarray(coordinates)
    azi, ele = degree_to_radian(c[..., 0]), degree_to_radian(c[..., 1])
    r = c[..., 2] if c.shape[-1] == 3 else 1
    return np.stack([r * np.cos(azi) * np.cos(ele), r * np.sin(azi) * np.cos(ele), r * np.sin(ele)], axis=-1)


if __name__ == '__main__':   # reference feature_extractor.py:274-307
    import argparse
    arg = argparse.ArgumentParser()
    arg.add_argument('--mode', default='foa', type=str, choices=['foa', 'mic'])
    arg.add_argument('--gpus', default='0', type=str)
    arg.add_argument('--root', default='/root/datasets/DCASE2020', type=str)
    config = arg.parse_args()
    os.environ['CUDA_VISIBLE_DEVICES'] = config.gpus
    mode = config.mode
    feature_out, label_out, norm_out = f'{mode}_dev', f'{mode}_dev_label', f'{mode}_dev_norm'
    extract_seldnet_data(os.path.join(config.root, f'{mode}_dev'), feature_out,
                         os.path.join(config.root, 'metadata_dev'), label_out,
                         mode=mode, win_length=960, hop_length=480, n_fft=1024)
    mean, std = calculate_statistics(feature_out)
    np.save('mean.npy', mean)
    np.save('std.npy', std)
    apply_normalizer(feature_out, norm_out, mean, std)
