"""Windowing / batching glue of the reference's training input pipeline, on resident device tensors.

Reference surface (data_loader.py):
  data_loader(dataset, preprocessing, sample_transforms, batch_transforms, deterministic, loop_time, batch_size)  :13-56
  seldnet_data_to_dataloader(features, labels, train, label_window_size, drop_remainder, shuffle_size,
                             batch_size, loop_time, **kwargs)                                                    :132-168
and the sliding-window evaluation framing of trainv2.py:158-192 (`ensemble_outputs`).

The reference builds a tf.data graph over host numpy arrays; here the normalised features stay in HBM as one
[sum T, F, C] tensor, a batch of consecutive windows is a zero-copy view of it, the sample transforms run as ONE fused
masking launch per batch (`transforms.sample_masks`) and the batch transforms as one remap launch
(`transforms.foa_intensity_vec_aug` ...).  Shapes, ordering (repeat -> batch -> batch-level shuffle) and the
drop-remainder rule follow the reference; random streams are this package's own (seeded, reproducible).
"""
import numpy as np
import torch

__all__ = ['data_loader', 'seldnet_data_to_dataloader', 'get_preprocessed_x', 'get_preprocessed_x_tf', 'foa_intensity_vectors_tf',
           'gcc_features_tf', 'TDM_aug', 'get_TDMset', 'normalize_over_clips', 'frame_windows', 'overlap_and_add_mean', 'ensemble_outputs']


def _as_tensor(a, device):
    t = a if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a))
    return t.to(device) if device is not None else t


def _apply(ops, x, y, per_sample):
    """Apply a list of (x, y) -> (x, y) callables.  per_sample: the reference maps these over single samples; ops
    carrying ``batched = True`` (transforms.sample_masks) take the whole batch in one launch instead."""
    if ops is None:
        return x, y
    if not isinstance(ops, (list, tuple)):
        ops = [ops]
    for op in ops:
        if not per_sample or getattr(op, 'batched', False):
            x, y = op(x, y)
        else:
            outs = [op(x[i], y[i]) for i in range(x.shape[0])]
            x = torch.stack([torch.as_tensor(o[0]) for o in outs])
            y = torch.stack([torch.as_tensor(o[1]) for o in outs])
    return x, y


class _BatchIterable:
    """repeat(loop_time) -> [sample transforms] -> batch(batch_size, drop_remainder=False) -> [batch transforms]
    (reference data_loader.py:49-54), then an optional tf.data-style buffer shuffle of the BATCHES (:161-164)."""

    def __init__(self, xs, ys, n_samples, batch_size, loop_time, sample_transforms, batch_transforms, shuffle_size, seed):
        self.xs, self.ys, self.n = xs, ys, int(n_samples)
        self.batch_size = int(batch_size)
        self.loop_time = None if loop_time is None else int(loop_time)      # None: repeat for ever (dataset.repeat(None), :49)
        self.sample_transforms, self.batch_transforms = sample_transforms, batch_transforms
        self.shuffle_size = shuffle_size
        self.rng = np.random.default_rng(seed)

    def __len__(self):
        if self.loop_time is None:
            raise TypeError('an endlessly repeating loader (loop_time=None) has no length')
        return -(-self.n * self.loop_time // self.batch_size)

    def batch_order(self):
        """Batch indices in the order they are yielded (buffer shuffle: fill `shuffle_size` slots, emit a random slot,
        refill it with the next batch -- tf.data.Dataset.shuffle semantics)."""
        n_batches = len(self)
        if not self.shuffle_size or self.shuffle_size <= 1:
            return list(range(n_batches))
        buf, out, nxt = [], [], 0
        while nxt < n_batches and len(buf) < self.shuffle_size:
            buf.append(nxt)
            nxt += 1
        while buf:
            i = int(self.rng.integers(len(buf)))
            out.append(buf[i])
            if nxt < n_batches:
                buf[i] = nxt
                nxt += 1
            else:
                buf.pop(i)
        return out

    def _slice(self, lo, hi):
        """Samples [lo, hi) of the repeated stream: a contiguous view where the range stays inside one pass, two views
        joined where it wraps once, a modular gather where a batch spans several passes (batch_size > n)."""
        lo_m, hi_m = lo % self.n, (hi - 1) % self.n + 1
        wraps = (hi - 1) // self.n - lo // self.n
        if wraps == 0:
            return self.xs[lo_m:hi_m], self.ys[lo_m:hi_m]
        if wraps == 1:
            return (torch.cat([self.xs[lo_m:], self.xs[:hi_m]]), torch.cat([self.ys[lo_m:], self.ys[:hi_m]]))
        idx = torch.arange(lo, hi, device=self.xs.device) % self.n
        return self.xs.index_select(0, idx), self.ys.index_select(0, idx.to(self.ys.device))

    def _batch(self, lo, hi):
        x, y = self._slice(lo, hi)
        x, y = _apply(self.sample_transforms, x, y, per_sample=True)
        return _apply(self.batch_transforms, x, y, per_sample=False)

    def __iter__(self):
        if self.loop_time is None:                      # endless: batches in stream order, shuffled buffer-wise on the fly
            if self.n == 0:
                return
            buf, b = [], 0
            size = self.shuffle_size if self.shuffle_size and self.shuffle_size > 1 else 1
            while True:
                while len(buf) < size:
                    buf.append(b)
                    b += 1
                i = int(self.rng.integers(len(buf))) if size > 1 else 0
                nxt = buf.pop(i)
                yield self._batch(nxt * self.batch_size, (nxt + 1) * self.batch_size)
        total = self.n * self.loop_time
        for b in self.batch_order():
            yield self._batch(b * self.batch_size, min((b + 1) * self.batch_size, total))


def data_loader(dataset, preprocessing=None, sample_transforms=None, batch_transforms=None, deterministic=False,
                loop_time=None, batch_size=32, shuffle_size=None, seed=0, device=None):
    """reference data_loader.py:13-56.  ``dataset`` = (xs, ys) with a common leading sample axis (tensors or arrays).
    ``preprocessing`` ops run once up front on the whole set (the reference caches their output); ``deterministic`` is
    accepted for signature parity -- the order here is always deterministic given ``seed``."""
    xs, ys = dataset
    xs, ys = _as_tensor(xs, device), _as_tensor(ys, device)
    xs, ys = _apply(preprocessing, xs, ys, per_sample=True)
    return _BatchIterable(xs, ys, xs.shape[0], batch_size, loop_time, sample_transforms, batch_transforms, shuffle_size, seed)


def seldnet_data_to_dataloader(features, labels, train=True, label_window_size=60, drop_remainder=True, shuffle_size=None,
                               batch_size=32, loop_time=1, device=None, seed=0, **kwargs):
    """reference data_loader.py:132-168.  features: list of [T_i, F, C] (or one [n, T, F, C] tensor), labels: list of
    [T_i / resolution, 4 * n_classes] -> batches x [B, label_window_size * resolution, F, C], y [B, label_window_size,
    4 * n_classes].  Windows are cut from the time-concatenated stream, non-overlapping, remainder dropped; evaluation
    (train=False) yields one clip per batch, unshuffled, a single pass."""
    if isinstance(features, (list, tuple)):
        total_length = labels[0].shape[0]
        feats = torch.cat([_as_tensor(f, device) for f in features], dim=0)
        labs = torch.cat([_as_tensor(l, device) for l in labels], dim=0)
    else:
        feats, labs = _as_tensor(features, device), _as_tensor(labels, device)
        total_length = labs.shape[1] if labs.dim() == 3 else labs.shape[0]
        if feats.dim() == 4:
            feats = feats.reshape(-1, *feats.shape[2:])
        if labs.dim() == 3:
            labs = labs.reshape(-1, labs.shape[-1])
    if feats.shape[0] % labs.shape[0]:
        raise ValueError('feature frames must be a multiple of label frames')
    resolution = feats.shape[0] // labs.shape[0]
    lws = int(label_window_size)
    n_samples = labs.shape[0] // lws
    if not drop_remainder and labs.shape[0] % lws:
        raise ValueError('a ragged last window cannot be batched (the reference fails there too): use drop_remainder=True')
    xs = feats[:n_samples * lws * resolution].reshape(n_samples, lws * resolution, *feats.shape[1:])     # views, no copy
    ys = labs[:n_samples * lws].reshape(n_samples, lws, labs.shape[-1])
    if not train:
        batch_size = total_length // lws
    shuffle = None
    if train:
        shuffle = n_samples // batch_size if shuffle_size is None else shuffle_size
    return _BatchIterable(xs, ys, n_samples, batch_size, loop_time if train else 1, kwargs.get('sample_transforms'),
                          kwargs.get('batch_transforms'), shuffle, seed)


def get_preprocessed_x(wav, sample_rate, mode='foa', n_mels=64, multiplier=5, max_label_length=600, **kwargs):
    """reference data_loader.py:268-308 (the on-the-fly extractor behind ``get_tdm_dataset``): features of one clip
    ``wav [4, L]`` (or a batch ``[n, 4, L]``), top_db-clamped, zero-padded / truncated to ``max_label_length * multiplier``
    frames -> CUDA float32 ``[max_len, n_mels, C]`` (``[n, max_len, n_mels, C]`` for a batch).  One fused launch + the clamp;
    the reference's numpy-or-tensor return type becomes a device tensor."""
    from . import _lib, pipeline
    from .plan import get_plan
    _lib.require_device()
    w = torch.as_tensor(wav)
    single = w.dim() == 2
    dev = w.device if w.is_cuda else torch.device('cuda', torch.cuda.current_device())
    w = (w.unsqueeze(0) if single else w).to(device=dev, dtype=torch.float32).contiguous()
    max_len = int(max_label_length) * int(multiplier)
    feat, key = pipeline.extract_batch(w, sample_rate, mode=mode, n_mels=n_mels, t_out=max_len, **kwargs)
    with torch.cuda.device(w.device):
        plan = get_plan(sample_rate, mode=mode, n_mels=n_mels, **{k: v for k, v in kwargs.items() if k != 'pad'})
    pipeline.finalize_(feat, key, plan.num_frames(w.shape[-1] + 2 * int(kwargs.get('pad', 0))))
    return feat[0] if single else feat


# --------------------------------------------------------------------------- TensorFlow-variant on-the-fly extractor
def get_preprocessed_x_tf(wav, sr, mode='foa', n_mels=64, multiplier=5, max_label_length=600, win_length=1024,
                          hop_length=480, n_fft=1024):
    """reference data_loader.py:310-349 (behind train.py:210-261 get_tdm_dataset): one clip ``wav [4, L]`` (or a batch
    ``[n, 4, L]``) -> CUDA float32 ``[max_label_length * multiplier, n_mels, 7]``: tf.signal.stft(pad_end=True) frames, mel
    bank (tf.signal.linear_to_mel_weight_matrix) on |X|, tfio dbscale(top_db=80), intensity vectors through the same bank,
    zero-padded / truncated.  One fused launch + the clamp.  mode='mic' raises like the reference does: its
    gcc_features_tf slices the time axis and the concat with the log-mel block fails on the shape."""
    from . import _lib, pipeline
    if mode != 'foa':
        raise ValueError('invalid mode')
    _lib.require_device()
    w = torch.as_tensor(wav)
    single = w.dim() == 2
    dev = w.device if w.is_cuda else torch.device('cuda', torch.cuda.current_device())
    w = (w.unsqueeze(0) if single else w).to(device=dev, dtype=torch.float32).contiguous()
    max_len = int(max_label_length) * int(multiplier)
    feat, key = pipeline.extract_batch_tf(w, sr, n_mels=n_mels, t_out=max_len, win_length=win_length, hop_length=hop_length, n_fft=n_fft)
    pipeline.finalize_(feat, key, -(-w.shape[-1] // hop_length))
    return feat[0] if single else feat


def foa_intensity_vectors_tf(spectrogram, eps=1e-8):
    """reference data_loader.py:237-251 for a complex ``[4, time, freq]`` tensor (stand-alone helper; the fused extractor
    computes the same per bin)."""
    s = torch.as_tensor(spectrogram)
    c0 = torch.conj(s[0])
    iv = torch.stack([(c0 * s[3]).real, (c0 * s[1]).real, (c0 * s[2]).real], 0)
    norm = torch.clamp_min(torch.sqrt((iv ** 2).sum(0)), eps)
    return iv / norm


def gcc_features_tf(complex_specs, n_mels):
    """reference data_loader.py:254-265, literally: irfft over the LAST axis of ``[chan, time, freq]``, then
    ``concat(cc[-n_mels//2:], cc[:(n_mels+1)//2], axis=0)`` -- which slices FRAMES, not lags -> ``[pairs, n_mels, n_fft]``
    (the reason the reference's mode='mic' TF path cannot be concatenated with the log-mel block)."""
    s = torch.as_tensor(complex_specs)
    out = []
    for m in range(s.shape[0]):
        for n in range(m + 1, s.shape[0]):
            r = torch.conj(s[m]) * s[n]
            cc = torch.fft.irfft(torch.exp(1j * torch.angle(r)), dim=-1)
            out.append(torch.cat([cc[-(n_mels // 2):], cc[:(n_mels + 1) // 2]], 0))
    return torch.stack(out, 0)


def normalize_over_clips(x):
    """train.py:232: ``(x - reduce_mean(x, 0)) / reduce_std(x, 0)`` -- statistics over the CLIP axis only, one per
    (frame, mel, channel) -- for the stacked ``[n_clips, T, F, C]`` output of get_preprocessed_x_tf."""
    x = torch.as_tensor(x)
    return (x - x.mean(0)) / x.std(0, unbiased=False)


STREAM_TDM = 0x110


def TDM_aug(x, y, tdm_x, tdm_y, sr=24000, label_resolution=0.1, max_overlap_num=5, max_overlap_per_frame=2, min_overlap_sec=1,
            max_overlap_sec=5, seed=0, return_draws=False):
    """reference data_loader.py:188-234, time-domain mixing of single-class recordings into the clips before extraction.
    x: list of ``[4, L]`` waveforms, y: list of ``[T_y, 4 * n_classes]`` labels, tdm_x / tdm_y: per class ``[4, frames]`` /
    ``[time, 4 * n_classes]``.  Per clip, ``max_overlap_num`` classes are drawn with probability ~ 1 / (class length), and for
    each a duration in [min, max) label frames, a position in the clip and a position in the class recording; label frames
    that already hold ``max_overlap_per_frame`` events or the same class are skipped.  The tensors stay where they are (the
    mixing runs on the device when they are CUDA tensors); lists are modified in place and returned, as in the reference.
    Draws: this package's counter-based Philox stream (seed, clip index, STREAM_TDM, event) -- the reference's are
    TensorFlow's unseeded stateful ops."""
    from . import philox
    n_cls = y[0].shape[-1] // 4
    lo, hi = int(min_overlap_sec / label_resolution), int(max_overlap_sec / label_resolution)
    spf = int(sr * label_resolution)
    lens = np.array([int(k.shape[0]) for k in tdm_y], dtype=np.float64)
    cdf = np.cumsum((1.0 / lens) / (1.0 / lens).sum())
    draws = []
    for i in range(len(x)):
        frames_y = int(y[i].shape[0])
        mine = []
        for j in range(max_overlap_num):
            w = philox.sample_words(int(seed), i, 1, STREAM_TDM, draw=j)[0].astype(np.uint64)
            cls = int(np.searchsorted(cdf, (float(w[0]) + 0.5) / 4294967296.0))
            cls = min(cls, len(tdm_y) - 1)
            st = lo + int(w[1] % np.uint64(hi - lo))
            off = int(w[2] % np.uint64(frames_y - st))
            tdo = int(w[3] % np.uint64(int(tdm_y[cls].shape[0]) - st))
            mine.append((cls, st, off, tdo))
            frame_y = y[i][off:off + st]
            nondup = 1 - frame_y[..., cls]
            valid = (frame_y[..., :n_cls].sum(-1) < max_overlap_per_frame).to(frame_y.dtype) * nondup
            if float(valid.sum()) == 0:
                continue
            y[i][off:off + st] += tdm_y[cls][tdo:tdo + st].to(y[i].device) * valid[:, None]
            gain = valid.to(x[i].dtype).repeat_interleave(spf)[None, :].to(x[i].device)
            x[i][:, off * spf:(off + st) * spf] += tdm_x[cls][:, tdo * spf:(tdo + st) * spf].to(x[i].device) * gain
        draws.append(mine)
    return (x, y, draws) if return_draws else (x, y)


def get_TDMset(TDM_PATH):
    """reference data_loader.py:170-185: the per-class recordings / labels written by the reference's TDM preparation
    (``foa_dev_tdm/tdm_noise_<c>.joblib``, ``tdm_label_<c>.joblib``) -> lists of float32 tensors."""
    import os
    from glob import glob
    import joblib
    tdm_path = os.path.join(TDM_PATH, 'foa_dev_tdm')
    class_num = len(glob(tdm_path + '/*label_*.joblib'))
    tdm_x = [torch.as_tensor(np.asarray(joblib.load(os.path.join(tdm_path, f'tdm_noise_{c}.joblib')), dtype=np.float32)) for c in range(class_num)]
    tdm_y = [torch.as_tensor(np.asarray(joblib.load(os.path.join(tdm_path, f'tdm_label_{c}.joblib')))) for c in range(class_num)]
    return tdm_x, tdm_y


# --------------------------------------------------------------------------- sliding-window evaluation (trainv2.py:158-192)
def frame_windows(x, win_size=300, step_size=5):
    """tf.signal.frame(x, win_size, step_size, axis=0) as a zero-copy strided view [n_win, win_size, ...]."""
    if not x.is_contiguous():
        x = x.contiguous()
    n_win = (x.shape[0] - win_size) // step_size + 1
    if n_win <= 0:
        return x.new_empty((0, win_size) + tuple(x.shape[1:]))
    inner = x.stride(0)
    return x.as_strided((n_win, win_size) + tuple(x.shape[1:]), (step_size * inner, inner) + tuple(x.stride()[1:]))


def overlap_and_add_mean(frames):
    """frames [n_win, L, K] of window outputs hopping by ONE output step -> [n_win + L - 1, K]: overlap-add divided by
    the overlap count (tf.signal.overlap_and_add(..., 1) / total_counts, trainv2.py:173-179)."""
    n_win, length, k = frames.shape
    out = frames.new_zeros(n_win + length - 1, k)
    cnt = frames.new_zeros(n_win + length - 1, 1)
    idx = (torch.arange(n_win, device=frames.device)[:, None] + torch.arange(length, device=frames.device)[None, :]).reshape(-1)
    out.index_add_(0, idx, frames.reshape(-1, k))
    cnt.index_add_(0, idx, frames.new_ones(idx.numel(), 1))
    return out / cnt


def ensemble_outputs(model, xs, win_size=300, step_size=5, batch_size=256):
    """trainv2.py:158-192: run ``model`` (a callable: windows [b, win, F, C] -> (sed [b, win/5, n], doa [b, win/5, m])) over
    every ``step_size``-hop window of each clip and average the overlapping per-label-frame outputs."""
    outs = []
    for x in xs:
        windows = frame_windows(_as_tensor(x, None), win_size, step_size)
        sed, doa = [], []
        for i in range(0, windows.shape[0], batch_size):
            s, d = model(windows[i:i + batch_size])
            sed.append(s)
            doa.append(d)
        outs.append((overlap_and_add_mean(torch.cat(sed)), overlap_and_add_mean(torch.cat(doa))))
    return outs
