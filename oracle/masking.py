"""CPU restatement of the reference's spectrogram masking -- TEST INFRASTRUCTURE ONLY.

Follows reference transforms.py:6-43 (``mask``: period-wise) and :46-75
(``simple_mask``: whole axis).  Per mask: size in [0, max_mask_size), then offset in
[0, total - size); band [offset, offset+size) is multiplied by 0 (so -0.0 / NaN survive);
time axis is framed into ``period``-long chunks, each chunk drawn independently, chunks
in order.  Parity status: ``simple_mask`` PINNED by the two known answers of reference
transforms_test.py:8-30 (with oracle/tf_random.TFEagerRandom); ``mask`` has no reference
test -- pinned only through the shared band logic.
"""
import numpy as np


def draw_bands(draw, total: int, max_mask_size, n_mask: int):
    """``draw(maxval)`` -> int.  Returns [(offset, size)] * n_mask in draw order."""
    if max_mask_size is None:
        max_mask_size = total
    bands = []
    for _ in range(n_mask):
        size = draw(max_mask_size)          # transforms.py:21 / :60
        offset = draw(total - size)         # transforms.py:22 / :61
        bands.append((offset, size))
    return bands


def apply_bands(x: np.ndarray, axis: int, bands) -> np.ndarray:
    total = x.shape[axis]
    keep = np.ones(total, dtype=x.dtype)
    for off, size in bands:
        keep[off:off + size] = 0
    shape = [1] * x.ndim
    shape[axis] = total
    return x * keep.reshape(shape)


def simple_mask_ref(x: np.ndarray, axis: int, draw, max_mask_size=None, n_mask=1):
    """reference transforms.py:46-75.  Returns (masked, bands)."""
    x = np.asarray(x)
    bands = draw_bands(draw, x.shape[axis], max_mask_size, n_mask)
    return apply_bands(x, axis, bands), bands


def mask_ref(x: np.ndarray, axis: int, draw_for_chunk, max_mask_size=None, period=100, n_mask=1):
    """reference transforms.py:6-43.  ``draw_for_chunk(chunk)`` -> ``draw(maxval)`` callable.
    Returns (masked, [[(offset, size)] per chunk])."""
    x = np.asarray(x)
    if x.shape[0] % period != 0:
        raise ValueError("(spec time length / period)' rest must be 0")
    ax = axis % x.ndim
    total = period if ax == 0 else x.shape[ax]
    out = np.empty_like(x)
    all_bands = []
    for c in range(x.shape[0] // period):
        bands = draw_bands(draw_for_chunk(c), total, max_mask_size, n_mask)
        out[c * period:(c + 1) * period] = apply_bands(x[c * period:(c + 1) * period], ax, bands)
        all_bands.append(bands)
    return out, all_bands
