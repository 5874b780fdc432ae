"""Masking oracle: TF-2 eager seeding + Philox reproduce the reference's known answers; Philox known-answer vectors."""
import numpy as np
import pytest

from oracle.masking import mask_ref, simple_mask_ref
from oracle.tf_random import CounterRandom, TFEagerRandom, philox4x32_10

ORG = np.arange(25).reshape(5, 5)


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10
    assert philox4x32_10((0, 0, 0, 0), (0, 0)) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert philox4x32_10((0xffffffff,) * 4, (0xffffffff,) * 2) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert philox4x32_10((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)


def test_reference_simple_mask_known_answers():
    """reference transforms_test.py:8-30."""
    r = TFEagerRandom(100)
    out, bands = simple_mask_ref(ORG, 0, r.uniform_int, None, 1)
    want = ORG.copy()
    want[:3] = 0
    assert bands == [(0, 3)] and np.array_equal(out, want)
    r = TFEagerRandom(2020)
    out, bands = simple_mask_ref(ORG, 1, r.uniform_int, 3, 2)
    want = ORG.copy()
    want[:, [0, 2]] = 0
    assert bands == [(2, 1), (0, 1)] and np.array_equal(out, want)


def test_mask_ref_semantics():
    x = np.random.default_rng(0).standard_normal((300, 64, 7)).astype(np.float32)
    cr = CounterRandom(123)
    out, bands = mask_ref(x, -3, lambda ch: cr.drawer(5, 0, ch), 24, 100, 2)
    assert len(bands) == 3
    for ch, bs in enumerate(bands):
        keep = np.ones(100, bool)
        for off, size in bs:
            assert 0 <= size < 24 and 0 <= off < 100 - size       # the last index is never masked
            keep[off:off + size] = False
        blk = out[ch * 100:(ch + 1) * 100]
        assert np.all(blk[~keep] == 0) and np.array_equal(blk[keep], x[ch * 100:(ch + 1) * 100][keep])
    with pytest.raises(ValueError):
        mask_ref(x[:250], 0, lambda ch: cr.drawer(0, 0, ch))
    # masked value is x * 0: -0.0 for negative inputs, NaN stays NaN
    y = np.array([[-1.0, 2.0], [np.nan, 3.0]], dtype=np.float32)
    out, _ = simple_mask_ref(y, 0, lambda m: [2, 0].pop(0) if m == 2 else 0, None, 1)
