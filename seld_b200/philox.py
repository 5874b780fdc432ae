"""Host-side Philox4x32-10 (numpy, vectorised) for the per-sample draws of the batch augmentations.

Same generator and the same (seed -> key, counter) convention as the masking kernel (csrc/mask.cu): counter =
(lo32(sample), hi32(sample), stream_id, draw).  A few integers per sample, so it runs on the host and the resulting
permutation / sign tables are uploaded with the launch."""
import numpy as np

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint64(0x9E3779B9), np.uint64(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Arrays (or scalars) of uint32 values held in uint64 -> the 4 output words."""
    c0, c1, c2, c3 = (np.asarray(v, dtype=np.uint64) & _MASK for v in (c0, c1, c2, c3))
    k0, k1 = np.uint64(k0) & _MASK, np.uint64(k1) & _MASK
    for _ in range(10):
        p0, p1 = _M0 * c0, _M1 * c2
        c0, c1, c2, c3 = ((p1 >> np.uint64(32)) ^ c1 ^ k0) & _MASK, p1 & _MASK, ((p0 >> np.uint64(32)) ^ c3 ^ k1) & _MASK, p0 & _MASK
        k0, k1 = (k0 + _W0) & _MASK, (k1 + _W1) & _MASK
    return c0, c1, c2, c3


def sample_words(seed: int, first_sample: int, n_samples: int, stream_id: int, draw: int = 0):
    """[n_samples, 4] uint32: the Philox block of (sample, stream_id, draw) for every sample."""
    s = np.arange(first_sample, first_sample + n_samples, dtype=np.uint64)
    out = philox4x32_10(s & _MASK, s >> np.uint64(32), np.uint64(stream_id), np.uint64(draw), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    return np.stack(out, axis=1).astype(np.uint32)
