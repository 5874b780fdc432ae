"""Drop-in for the masking half of the reference's ``transforms`` module, computed on a B200.

``mask`` and ``simple_mask`` keep the reference's names, argument order, defaults and error behaviour
(transforms.py:6-43, :46-75); they accept numpy arrays or torch tensors and return a NEW array of the same kind,
shape and dtype.  ``mask_batch_`` is the fused, in-place form for whole training batches (time and frequency
masks of every sample in one kernel).

Random streams
  * ``set_seed(s)`` mirrors ``tf.random.set_seed(s)`` in eager mode: the next draws follow TensorFlow-2's stream
    bit for bit (this is how the known answers of reference transforms_test.py:8-30 are reproduced).
  * without it (the reference's own training runs never seed, and tf.data draws in a nondeterministic order --
    there is nothing to reproduce) a process-wide counter-based Philox stream is used: draws are a pure function of
    (seed, sample index, axis, chunk, mask, draw), so any batch can be re-generated.
"""
import os
import random
import threading

import numpy as np
import torch

from . import _lib

__all__ = ['mask', 'simple_mask', 'mask_batch_', 'sample_masks', 'augment_batch', 'batch_augment', 'random_ups_and_downs', 'set_seed', 'set_counter_seed', 'foa_intensity_vec_aug', 'acs_aug',
           'mic_gcc_perm', 'channel_list', 'split_total_labels_to_sed_doa']

_MAXINT32 = 2 ** 31 - 1
_state = threading.local()
_global_lock = threading.Lock()
_counter = {'seed': int.from_bytes(os.urandom(8), 'little'), 'next_sample': 0}


# --------------------------------------------------------------------------- random streams (host side)
class _TFEagerStream:
    """TensorFlow-2 eager seeding: graph seed g; every un-seeded op takes ``rng.randint(0, 2**31-1)`` from a
    ``random.Random(g)``; kernel seeds are (g mod M, op mod M) with M = 2**31-1, (0, 0) -> (0, M)."""

    def __init__(self, seed):
        self.graph_seed = int(seed)
        self._rng = random.Random(self.graph_seed)

    def kernel_seed(self):
        return self.graph_seed % _MAXINT32

    def next_seed2(self):
        op = self._rng.randint(0, _MAXINT32) % _MAXINT32
        if self.kernel_seed() == 0 and op == 0:
            op = _MAXINT32
        return op


def set_seed(seed):
    """Like ``tf.random.set_seed``: subsequent ``mask`` / ``simple_mask`` calls on this thread follow TF's eager
    stream.  ``set_seed(None)`` returns to the counter-based stream."""
    _state.tf = None if seed is None else _TFEagerStream(seed)


def set_counter_seed(seed, next_sample=0):
    """Seed the process-wide counter-based stream (used when no TF-compatible seed is set)."""
    with _global_lock:
        _counter['seed'] = int(seed) & (2 ** 64 - 1)
        _counter['next_sample'] = int(next_sample)


def _take_samples(n):
    with _global_lock:
        first = _counter['next_sample']
        _counter['next_sample'] += n
        return _counter['seed'], first


# --------------------------------------------------------------------------- kernel launch
def _launch(x, n_samples, t, mid, f, c, period, time_max, time_n, freq_max, freq_n, seed, sample_offset, rng_mode,
            op_seed2=None, want_draws=False):
    code = _lib.DTYPE_CODES.get(str(x.dtype).replace('torch.', ''))
    if code is None:
        raise TypeError(f'unsupported dtype {x.dtype}')
    n_chunks = 1 if period <= 0 else t // period
    draws = None
    if want_draws:
        draws = torch.empty(n_samples, n_chunks, time_n + freq_n, 2, dtype=torch.int32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().seld_mask(_lib.ptr(x), code, n_samples, t, mid, f, c, int(period),
                                         int(time_max or 0), int(time_n), int(freq_max or 0), int(freq_n),
                                         int(seed), int(sample_offset), rng_mode, _lib.ptr(op_seed2), _lib.ptr(draws),
                                         _lib.current_stream_ptr()))
    return draws


def _to_cuda_copy(specs):
    """-> (fresh contiguous CUDA tensor, restore(tensor) -> same kind as the input)."""
    _lib.require_device()
    dev = torch.device('cuda', torch.cuda.current_device())
    if isinstance(specs, torch.Tensor):
        src_dev = specs.device
        x = specs.to(dev).contiguous()
        if x.data_ptr() == specs.data_ptr():
            x = x.clone()
        return x, (lambda y: y if src_dev.type == 'cuda' else y.to(src_dev))
    arr = np.ascontiguousarray(specs)
    return torch.from_numpy(arr.copy()).to(dev), (lambda y: y.cpu().numpy())


def _single_axis(x, axis, max_mask_size, n_mask, period):
    """Mask one axis of one sample through the reference-signature entry points."""
    nd = x.dim()
    ax = axis % nd if -nd <= axis < nd else None
    if ax is None:
        raise ValueError('axis out of range')
    shape = list(x.shape)
    n_mask = int(n_mask)
    if period is not None and ax == 0:          # time masks inside each period-long chunk
        t, mid, f, c = shape[0], 1, 1, int(np.prod(shape[1:], dtype=np.int64))
        tm, tn, fm, fn, total = max_mask_size, n_mask, 0, 0, period
    elif period is not None:                    # another axis, drawn independently per chunk of the time axis
        t, mid, f, c = shape[0], int(np.prod(shape[1:ax], dtype=np.int64)), shape[ax], int(np.prod(shape[ax + 1:], dtype=np.int64))
        tm, tn, fm, fn, total = 0, 0, max_mask_size, n_mask, shape[ax]
    else:                                       # simple_mask: one chunk, everything before `axis` folds into mid
        t, mid, f, c = 1, int(np.prod(shape[:ax], dtype=np.int64)), shape[ax], int(np.prod(shape[ax + 1:], dtype=np.int64))
        tm, tn, fm, fn, total = 0, 0, max_mask_size, n_mask, shape[ax]
    if max_mask_size is not None and not 0 < int(max_mask_size) <= total:
        raise ValueError('max_mask_size must be in (0, axis length]')
    if x.numel() == 0 or n_mask == 0:
        return
    per = period if period is not None else 0
    n_chunks = 1 if per == 0 else t // per
    tf = getattr(_state, 'tf', None)
    if tf is not None:
        seeds = [tf.next_seed2() for _ in range(n_chunks * n_mask * 2)]      # draw order: chunk, mask, (size, offset)
        op_seed2 = torch.tensor(seeds, dtype=torch.int64, device=x.device)
        _launch(x, 1, t, mid, f, c, per, tm, tn, fm, fn, tf.kernel_seed(), 0, _lib.RNG_TF_EAGER_COMPAT, op_seed2)
    else:
        seed, first = _take_samples(1)
        _launch(x, 1, t, mid, f, c, per, tm, tn, fm, fn, seed, first, _lib.RNG_PHILOX_COUNTER)


def mask(specs, axis, max_mask_size=None, period=100, n_mask=1):
    """reference transforms.py:6-43: cut the time axis (axis 0) into ``period``-long chunks and, in each chunk
    independently, zero ``n_mask`` random bands of ``axis``; size in [0, max_mask_size), offset in [0, total-size)."""
    shape = tuple(specs.shape)
    if shape[0] % period != 0:
        raise ValueError("(spec time length / period)' rest must be 0")
    x, restore = _to_cuda_copy(specs)
    _single_axis(x, axis, max_mask_size, n_mask, int(period))
    return restore(x)


def simple_mask(specs, axis, max_mask_size=None, n_mask=1):
    """reference transforms.py:46-75: ``n_mask`` random zero bands along ``axis`` of the whole array."""
    x, restore = _to_cuda_copy(specs)
    _single_axis(x, axis, max_mask_size, n_mask, None)
    return restore(x)


def mask_batch_(x, time_mask=(24, 1), freq_mask=(16, 1), period=100, seed=None, sample_offset=None, return_draws=False):
    """Fused in-place masking of a CUDA batch ``x[B, T, F, C]``: per sample and per ``period``-frame chunk,
    ``time_mask = (max_size, n)`` bands of frames then ``freq_mask = (max_size, n)`` bands of mel bins -- the
    ``sample_transforms`` of reference train.py:157-160 (24x1, 16x1) / trainv2.py:136-137 (6x10, 8x6) in one pass.
    Draws depend only on (seed, sample_offset + b, axis, chunk, mask, draw).
    After ``set_seed(s)`` (and with ``seed=None``) the bands follow TensorFlow-2's eager stream instead, in the order the
    reference's two calls per sample consume it -- ``mask(x, axis=-3, ...)`` then ``mask(x, axis=-2, ...)``, sample after
    sample: all time draws of all chunks, then all frequency draws -- still as ONE launch."""
    if not (isinstance(x, torch.Tensor) and x.is_cuda and x.is_contiguous() and x.dim() == 4):
        raise ValueError('x must be a contiguous CUDA tensor [B, T, F, C]')
    b, t, f, c = x.shape
    if t % period != 0:
        raise ValueError("(spec time length / period)' rest must be 0")
    tm, tn = time_mask if time_mask else (0, 0)
    fm, fn = freq_mask if freq_mask else (0, 0)
    tf = getattr(_state, 'tf', None)
    if seed is None and tf is not None and b > 0 and (tn + fn) > 0:
        n_chunks = t // period
        seeds = [tf.next_seed2() for _ in range(b * n_chunks * (int(tn) + int(fn)) * 2)]
        op_seed2 = torch.tensor(seeds, dtype=torch.int64, device=x.device)
        return _launch(x, b, t, 1, f, c, int(period), tm, int(tn), fm, int(fn), tf.kernel_seed(), 0, _lib.RNG_TF_EAGER_TWO_PASS,
                       op_seed2, return_draws)
    if seed is None:
        seed, first = _take_samples(b)
        sample_offset = first if sample_offset is None else sample_offset
    elif sample_offset is None:
        sample_offset = 0
    return _launch(x, b, t, 1, f, c, int(period), tm, int(tn), fm, int(fn), int(seed) & (2 ** 64 - 1), int(sample_offset),
                   _lib.RNG_PHILOX_COUNTER, None, return_draws)


STREAM_LEVEL_JITTER = 0x102


def random_ups_and_downs(x, y, stddev=0.2, seed=None, sample_offset=None, return_draws=False):
    """reference trainv2.py:120-124: add ONE N(0, 0.2^2) scalar to channels [:4] (the log-mel block).  The reference maps
    it over single samples [T, F, C]; here x may also be a batch [B, T, F, C] (one independent scalar per sample).
    The scalar is drawn on the device: Box-Muller in float64 on Philox words 0, 1 of (seed, sample, STREAM_LEVEL_JITTER)."""
    xt = torch.as_tensor(x)
    single = xt.dim() == 3
    _lib.require_device()
    xb = (xt.unsqueeze(0) if single else xt).cuda()
    out, _, draws = augment_batch(xb, None, level_jitter=stddev, seed=seed, sample_offset=sample_offset, return_draws=True)
    out = out[0] if single else out
    return (out, y, draws[:, 1].cpu().numpy().view(np.float32).copy()) if return_draws else (out, y)


def sample_masks(time_mask=(24, 1), freq_mask=(16, 1), period=100, seed=None, level_jitter=None):
    """The reference's per-sample transforms (train.py:157-160: ``mask(x, axis=-3, max_mask_size=24, n_mask=1)`` then
    ``mask(x, axis=-2, max_mask_size=16)``; trainv2.py:134-138 puts ``random_ups_and_downs`` in front: ``level_jitter=0.2``)
    as ONE batched transform for ``data_loader.seldnet_data_to_dataloader``: ``(x [B, T, F, C], y) -> (new x, y)``,
    independent draws per sample; jitter, copy and both mask axes are ONE launch (seld_augment_batch)."""
    state = {'next_sample': 0}          # with an explicit seed: this transform's own running sample index, so that every
    lock = threading.Lock()             # batch (and every epoch) draws fresh bands, reproducibly

    def op(x, y):
        first = None
        if seed is not None:
            with lock:
                first = state['next_sample']
                state['next_sample'] += int(x.shape[0])
        out, _ = augment_batch(x, None, None, level_jitter, time_mask, freq_mask, period, seed, first)      # one launch
        return out, y
    op.batched = True
    return op


# --------------------------------------------------------------------------- fused augmentation launch (device draws)
_SPATIAL = {None: 0, 'none': 0, 'foa': 1, 'foa_iv': 1, 'acs': 2}


def augment_batch(x, y=None, spatial=None, level_jitter=None, time_mask=None, freq_mask=None, period=100, seed=None,
                  sample_offset=None, return_draws=False):
    """ONE launch over a CUDA batch ``x [B, T, F, C]`` (+ one tiny launch over the labels ``y [B, T_y, 4 * n_classes]``) that
    applies, with every draw made on the device: the level jitter of trainv2.py:120-124 (``level_jitter`` = stddev), the
    spatial augmentation ``'foa'`` (foa_intensity_vec_aug, transforms.py:78-114) or ``'acs'`` (acs_aug, :155-199), and the
    time / frequency masks of transforms.py:6-43 (``(max_size, n)`` per ``period``-frame chunk).  Returns new tensors
    ``(x, y)`` -- bit-identical to foa_intensity_vec_aug / acs_aug, random_ups_and_downs and mask_batch_ applied one after
    the other with the same ``seed`` and ``sample_offset`` (the Philox streams are shared)."""
    _lib.require_device()
    if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dim() == 4):
        raise ValueError('x must be a CUDA tensor [B, T, F, C]')
    xs = x.to(torch.float32).contiguous()
    b, t, f, c = xs.shape
    code = _SPATIAL.get(spatial)
    if code is None:
        raise ValueError('spatial must be None, "foa" or "acs"')
    ys = yo = None
    n_cls = t_y = 0
    if y is not None and code != 0:
        ys = torch.as_tensor(y).to(device=xs.device, dtype=torch.float32).contiguous()
        if ys.dim() != 3 or ys.shape[0] != b or ys.shape[-1] % 4:
            raise ValueError('y must be [B, T_y, 4 * n_classes]')
        t_y, n_cls = ys.shape[1], ys.shape[2] // 4
        yo = torch.empty_like(ys)
    tm, tn = time_mask if time_mask else (0, 0)
    fm, fn = freq_mask if freq_mask else (0, 0)
    if (tn or fn) and t % period != 0:
        raise ValueError("(spec time length / period)' rest must be 0")
    seed, first = _draw_seed(b, seed, sample_offset)
    out = torch.empty_like(xs)
    draws = torch.empty(b, 2, dtype=torch.int32, device=xs.device) if return_draws else None
    with torch.cuda.device(xs.device):
        _lib.check(_lib.load().seld_augment_batch(_lib.ptr(xs), _lib.ptr(out), b, t, f, c, _lib.ptr(ys), _lib.ptr(yo), t_y, n_cls, code,
                                                  float(level_jitter or 0.0), int(period), int(tm or 0), int(tn), int(fm or 0), int(fn),
                                                  int(seed), int(first), _lib.ptr(draws), _lib.current_stream_ptr()))
    y_out = yo if yo is not None else y
    return (out, y_out, draws) if return_draws else (out, y_out)


def batch_augment(spatial=None, level_jitter=None, time_mask=(24, 1), freq_mask=(16, 1), period=100, seed=None):
    """The whole augmentation chain of train.py:157-165 / trainv2.py:134-138 as ONE batched transform for
    ``data_loader.seldnet_data_to_dataloader(sample_transforms=[...])``: ``(x [B, T, F, C], y) -> (new x, new y)``."""
    state = {'next_sample': 0}
    lock = threading.Lock()

    def op(x, y):
        first = None
        if seed is not None:
            with lock:
                first = state['next_sample']
                state['next_sample'] += int(x.shape[0])
        return augment_batch(x, y, spatial, level_jitter, time_mask, freq_mask, period, seed, first)
    op.batched = True
    return op


# --------------------------------------------------------------------------- batch-level spatial augmentations (f1)
STREAM_IV_AUG, STREAM_ACS_AUG = 0x100, 0x101        # Philox stream ids (masking uses the chunk index < 2**24 there)

# reference transforms.py:143-152 (arXiv:2101.02919, table 1): [[mic channel], [foa channel]] for the 8 swaps
channel_list = [
    [[1, 3, 0, 2], [0, -3, -2, 1]],
    [[3, 1, 2, 0], [0, -3, 2, -1]],
    [[0, 1, 2, 3], [0, 1, 2, 3]],
    [[1, 0, 3, 2], [0, -1, -2, 3]],
    [[2, 0, 3, 1], [0, 3, -2, -1]],
    [[0, 2, 1, 3], [0, 3, 2, 1]],
    [[3, 2, 1, 0], [0, -1, 2, -3]],
    [[2, 3, 0, 1], [0, 1, -2, -3]],
]
_GCC_PAIRS = [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]


def split_total_labels_to_sed_doa(x, y):
    """reference transforms.py:117-119."""
    n_classes = y.shape[-1] // 4
    return x, (y[..., :n_classes], y[..., n_classes:])


def mic_gcc_perm(mic_perm):
    """reference transforms.py:122-139: microphone permutation [B, 4] -> permutation of the 6 GCC pair channels
    (pair order (0,1),(0,2),(0,3),(1,2),(1,3),(2,3)); exact table pinned by transforms_test.py:64-73."""
    mp = np.asarray(mic_perm.cpu() if isinstance(mic_perm, torch.Tensor) else mic_perm, dtype=np.int64)
    decode = np.array([[0, 0, 1, 2], [0, 0, 3, 4], [1, 3, 0, 5], [2, 4, 5, 0]], dtype=np.int64)
    out = np.empty((mp.shape[0], 6), dtype=np.int32)
    for j, (a, b) in enumerate(_GCC_PAIRS):
        out[:, j] = decode[mp[:, a], mp[:, b]]
    return torch.from_numpy(out) if isinstance(mic_perm, torch.Tensor) else out


def _draw_seed(n, seed, sample_offset):
    if seed is None:
        seed, first = _take_samples(n)
        return seed, first if sample_offset is None else sample_offset
    return int(seed) & (2 ** 64 - 1), 0 if sample_offset is None else int(sample_offset)


def _aug_inputs(x, y):
    _lib.require_device()
    dev = torch.device('cuda', torch.cuda.current_device())
    xs = torch.as_tensor(x)
    ys = torch.as_tensor(y)
    xs = xs if xs.is_cuda else xs.to(dev)
    return xs.to(torch.float32).contiguous(), ys.to(device=xs.device, dtype=torch.float32).contiguous()


def foa_intensity_vec_aug(x, y, seed=None, sample_offset=None, return_draws=False):
    """reference transforms.py:78-114 for x [B, T, F, 7], y [B, T, 4*n_classes]: per sample, random sign flips of the
    three intensity / coordinate axes and a random x<->z swap ([0,1,2] or the reference's [2,1,0] channel order),
    applied consistently to the FOA channels 1..3, the intensity vectors 4..6 and the label coordinates.
    Draws (on the device): Philox words 0..2 & 1 = flips, word 3 & 1 = swap, per global sample index."""
    xs, ys = _aug_inputs(x, y)
    if xs.dim() != 4 or xs.shape[-1] != 7 or ys.shape[-1] % 4:
        raise ValueError('x must be [B, T, F, 7] and y [B, T, 4*n_classes]')
    xo, yo, draws = augment_batch(xs, ys, spatial='foa', seed=seed, sample_offset=sample_offset, return_draws=True)
    if not return_draws:
        return xo, yo
    w = draws[:, 0].cpu().numpy().astype(np.int64)
    return xo, yo, {'flip': np.stack([w & 1, (w >> 1) & 1, (w >> 2) & 1], 1), 'swap': (w >> 3) & 1}


def acs_aug(x, y, seed=None, sample_offset=None, return_draws=False):
    """reference transforms.py:155-199, audio channel swapping for x [B, T, F, 17] (4 FOA log-mel, 3 IV, 4 MIC log-mel,
    6 GCC) and y [B, T, 4*n_classes]: one of the 8 rotations / reflections of `channel_list` per sample, applied to
    the FOA channels, intensity vectors (with signs), microphone channels, GCC pair channels and label coordinates.
    Draw (on the device): Philox word 0 % 8 per global sample index."""
    xs, ys = _aug_inputs(x, y)
    if xs.dim() != 4 or xs.shape[-1] != 17 or ys.shape[-1] % 4:
        raise ValueError('x must be [B, T, F, 17] and y [B, T, 4*n_classes]')
    xo, yo, draws = augment_batch(xs, ys, spatial='acs', seed=seed, sample_offset=sample_offset, return_draws=True)
    return (xo, yo, {'idx': draws[:, 0].cpu().numpy().astype(np.int64)}) if return_draws else (xo, yo)
