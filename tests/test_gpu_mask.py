"""Masking kernel: bit-exact band positions vs the CPU restatement, the reference's known answers in TF-eager mode."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ORG = np.array([[0, 1, 2, 3, 4], [5, 6, 7, 8, 9], [10, 11, 12, 13, 14], [15, 16, 17, 18, 19], [20, 21, 22, 23, 24]])


def test_reference_known_answers_tf_eager_mode():
    """reference transforms_test.py:8-30, bit for bit (int64 input, as in the reference's test)."""
    from seld_b200 import transforms as T
    T.set_seed(100)
    target = ORG.copy(); target[:3] = 0
    out = T.simple_mask(ORG, axis=0, max_mask_size=None, n_mask=1)
    assert isinstance(out, np.ndarray) and out.dtype == ORG.dtype and np.array_equal(out, target)
    T.set_seed(2020)
    target = ORG.copy(); target[:, [0, 2]] = 0
    assert np.array_equal(T.simple_mask(ORG, axis=1, max_mask_size=3, n_mask=2), target)
    T.set_seed(None)


@pytest.mark.parametrize('seed', [1, 77, 2021])
def test_tf_eager_mask_matches_oracle(seed):
    from oracle.masking import mask_ref, simple_mask_ref
    from oracle.tf_random import TFEagerRandom
    from seld_b200 import transforms as T
    x = np.random.default_rng(seed).standard_normal((300, 64, 7)).astype(np.float32)
    for axis, mx, n in ((-3, 24, 1), (-2, 16, 1), (0, 6, 10), (1, 8, 6), (2, None, 1)):
        r = TFEagerRandom(seed)
        want, _ = mask_ref(x, axis, lambda ch: r.uniform_int, mx, 100, n)
        T.set_seed(seed)
        got = T.mask(x, axis, max_mask_size=mx, n_mask=n)
        assert np.array_equal(got, want), (axis, mx, n)
    r = TFEagerRandom(seed)
    want, _ = simple_mask_ref(x, 1, r.uniform_int, 10, 3)
    T.set_seed(seed)
    assert np.array_equal(T.simple_mask(x, 1, 10, 3), want)
    T.set_seed(None)


@pytest.mark.parametrize('params', [((24, 1), (16, 1)), ((6, 10), (8, 6)), ((24, 6), (8, 1)), (None, (16, 2)), ((24, 1), None)])
def test_fused_batch_mask_bit_exact_vs_oracle(params):
    """Config 5(i): B x [300, 64, 7]; positions from the counter-based Philox stream == CPU restatement."""
    from oracle.masking import apply_bands, draw_bands
    from oracle.tf_random import CounterRandom
    from seld_b200 import transforms as T
    tmask, fmask = params
    b, seed, off = 8, 0xDEADBEEF12345, 1000
    x = torch.randn(b, 300, 64, 7, generator=torch.Generator().manual_seed(5))
    y = x.cuda().clone()
    draws = T.mask_batch_(y, tmask, fmask, period=100, seed=seed, sample_offset=off, return_draws=True)
    got, draws = y.cpu().numpy(), draws.cpu().numpy()
    cr = CounterRandom(seed)
    want = x.numpy().copy()
    for s in range(b):
        for ch in range(3):
            blk = want[s, ch * 100:(ch + 1) * 100]
            bands = []
            if tmask:
                tb = draw_bands(cr.drawer(off + s, 0, ch), 100, tmask[0], tmask[1])
                blk = apply_bands(blk, 0, tb); bands += tb
            if fmask:
                fb = draw_bands(cr.drawer(off + s, 1, ch), 64, fmask[0], fmask[1])
                blk = apply_bands(blk, 1, fb); bands += fb
            want[s, ch * 100:(ch + 1) * 100] = blk
            assert [tuple(d) for d in draws[s, ch]] == bands          # (offset, size) per mask, time first
    assert np.array_equal(got, want)
    # same seed / offset => same masks; different offset => different
    z = x.cuda().clone(); T.mask_batch_(z, tmask, fmask, seed=seed, sample_offset=off)
    assert torch.equal(z, y)


def test_multiply_by_zero_semantics_and_dtypes():
    from seld_b200 import transforms as T
    x = torch.full((100, 4, 2), -3.0)
    x[5, 1, 0] = float('nan')
    T.set_seed(1)
    y = T.simple_mask(x, 0, None, 3)
    assert torch.is_tensor(y) and y.shape == x.shape and not y.is_cuda
    zeroed = (y == 0) & ~torch.isnan(y)
    assert zeroed.any() and torch.all(torch.signbit(y[zeroed]))      # x * 0 = -0.0 for negative x, like the reference
    assert torch.equal(torch.isnan(y), torch.isnan(x))                # NaN * 0 = NaN
    for dt in (torch.float64, torch.float16, torch.bfloat16, torch.int32, torch.int16, torch.uint8, torch.int64):
        a = (torch.arange(1, 601).reshape(100, 3, 2) % 100 + 1).to(dt)
        T.set_seed(9)
        ref = T.simple_mask(a.to(torch.float32), 0, 20, 2)
        T.set_seed(9)
        got = T.simple_mask(a, 0, 20, 2)
        assert got.dtype == dt and torch.equal(got.to(torch.float32), ref)
    T.set_seed(None)
    with pytest.raises(ValueError):
        T.mask(torch.zeros(250, 4, 2), 0)
    with pytest.raises(ValueError):
        T.mask(torch.zeros(300, 4, 2), 1, max_mask_size=9)           # > axis length: tf.random.uniform would raise too


def test_unseeded_stream_is_counter_based_and_advances():
    from seld_b200 import transforms as T
    T.set_seed(None)
    T.set_counter_seed(42, 0)
    x = torch.randn(300, 64, 7).cuda()
    a = T.mask(x, 0, 24); b = T.mask(x, 0, 24)
    assert not torch.equal(a, b)                                      # successive samples draw different masks
    T.set_counter_seed(42, 0)
    assert torch.equal(T.mask(x, 0, 24), a) and torch.equal(T.mask(x, 0, 24), b)
    assert a.is_cuda and torch.equal(x, x)                            # input untouched, result stays on the GPU


@pytest.mark.parametrize('params', [((24, 1), (16, 1)), ((6, 10), (8, 6))])
def test_fused_batch_mask_follows_the_tf_eager_stream_of_two_reference_calls(params):
    """After set_seed(s), ONE fused launch over a batch == the reference's eager sequence per sample: mask(x, axis=-3) -- all
    chunks' time draws -- then mask(x, axis=-2) -- all chunks' frequency draws (train.py:157-160), sample after sample."""
    from oracle.masking import mask_ref
    from oracle.tf_random import TFEagerRandom
    from seld_b200 import transforms as T
    (tm, tn), (fm, fn) = params
    x = np.random.default_rng(3).standard_normal((5, 300, 64, 7)).astype(np.float32)
    r = TFEagerRandom(404)
    want = np.empty_like(x)
    for s in range(x.shape[0]):
        a, _ = mask_ref(x[s], -3, lambda ch: r.uniform_int, tm, 100, tn)
        want[s], _ = mask_ref(a, -2, lambda ch: r.uniform_int, fm, 100, fn)
    T.set_seed(404)
    y = torch.from_numpy(x).cuda()
    T.mask_batch_(y, (tm, tn), (fm, fn))
    T.set_seed(None)
    assert np.array_equal(y.cpu().numpy(), want)
