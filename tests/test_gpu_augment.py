"""CUDA spatial augmentations (seld_channel_remap through transforms.foa_intensity_vec_aug / acs_aug) against the
oracle restatement on the same draws: bit-exact (sign flips and gathers only)."""
import numpy as np
import pytest
import torch

from oracle import augment as A
from seld_b200 import _lib, transforms as T

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('shape_x,shape_y', [((1, 10, 32, 7), (1, 2, 12)), ((5, 300, 64, 7), (5, 60, 56)), ((256, 30, 64, 7), (256, 6, 48))])
def test_iv_aug_matches_oracle(shape_x, shape_y):
    g = torch.Generator().manual_seed(2022)
    x = torch.rand(shape_x, generator=g) - 0.5
    y = torch.rand(shape_y, generator=g) - 0.5
    nx, ny, d = T.foa_intensity_vec_aug(x.cuda(), y.cuda(), seed=99, sample_offset=1000, return_draws=True)
    ox, oy = A.foa_intensity_vec_aug_ref(x.numpy(), y.numpy(), d['flip'], d['swap'])
    assert np.array_equal(nx.cpu().numpy(), ox) and np.array_equal(ny.cpu().numpy(), oy)
    # the reference's own property (transforms_test.py:46-52)
    xf = (x.numpy()[..., -3:] != ox[..., -3:]).astype(np.float32).mean(axis=(1, 2))
    s = shape_y[:-1] + (4, -1)
    yf = (y.numpy().reshape(s)[..., -3:, :] != oy.reshape(s)[..., -3:, :]).astype(np.float32).mean(axis=(1, 3))
    assert np.array_equal(xf, yf)
    if shape_x[0] >= 256:                              # all 16 (flip, swap) combinations occur
        assert len({(tuple(f), int(s_)) for f, s_ in zip(d['flip'], d['swap'])}) == 16


def test_iv_aug_inputs_untouched_and_reproducible():
    x = torch.rand(4, 20, 64, 7, device='cuda')
    y = torch.rand(4, 4, 56, device='cuda')
    x0, y0 = x.clone(), y.clone()
    a = T.foa_intensity_vec_aug(x, y, seed=5)
    b = T.foa_intensity_vec_aug(x, y, seed=5)
    c = T.foa_intensity_vec_aug(x, y, seed=5, sample_offset=4)
    assert torch.equal(x, x0) and torch.equal(y, y0)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    assert not torch.equal(a[0], c[0])


@pytest.mark.parametrize('shape_x,shape_y', [((2, 10, 32, 17), (2, 2, 12)), ((64, 100, 64, 17), (64, 20, 56))])
def test_acs_aug_matches_oracle(shape_x, shape_y):
    g = torch.Generator().manual_seed(2022)
    x = torch.rand(shape_x, generator=g) - 0.5
    y = torch.rand(shape_y, generator=g) - 0.5
    nx, ny, d = T.acs_aug(x.cuda(), y.cuda(), seed=3, return_draws=True)
    ox, oy = A.acs_aug_ref(x.numpy(), y.numpy(), d['idx'])
    assert np.array_equal(nx.cpu().numpy(), ox) and np.array_equal(ny.cpu().numpy(), oy)
    if shape_x[0] >= 64:
        assert set(d['idx'].tolist()) == set(range(8))


def test_channel_remap_errors():
    lib = _lib.load()
    x = torch.zeros(2, 3, 33, device='cuda')
    o = torch.zeros_like(x)
    p = torch.zeros(2, 33, dtype=torch.int32, device='cuda')
    assert lib.seld_channel_remap(_lib.ptr(x), _lib.ptr(o), 2, 3, 33, 1, _lib.ptr(p), None) == -4
    assert lib.seld_channel_remap(_lib.ptr(x), _lib.ptr(x), 2, 3, 4, 1, _lib.ptr(p), None) == -1      # out of place only
    assert lib.seld_channel_remap(None, _lib.ptr(o), 2, 3, 4, 1, _lib.ptr(p), None) == -1
    assert lib.seld_channel_remap(_lib.ptr(x), _lib.ptr(o), 0, 3, 4, 1, _lib.ptr(p), None) == 0
    with pytest.raises(ValueError):
        T.foa_intensity_vec_aug(torch.zeros(1, 2, 3, 6), torch.zeros(1, 2, 12))


def test_random_ups_and_downs_adds_one_scalar_to_the_logmel_channels():
    g = torch.Generator().manual_seed(4)
    x = torch.rand(6, 300, 64, 7, generator=g) - 0.5
    nx, y, offs = T.random_ups_and_downs(x.cuda(), 'labels', seed=21, sample_offset=3, return_draws=True)
    assert y == 'labels' and offs.shape == (6,) and len(set(offs.tolist())) == 6
    want = x.numpy().copy()
    want[..., :4] += offs[:, None, None, None]                                    # reference trainv2.py:120-124
    assert np.array_equal(nx.cpu().numpy(), want)
    one, _ = T.random_ups_and_downs(x[2].cuda(), None, seed=21, sample_offset=5)     # single sample [T, F, C], same draw as sample 2
    assert np.array_equal(one.cpu().numpy(), want[2])


def test_sample_masks_with_level_jitter():
    x = (torch.rand(8, 300, 64, 7) + 1.0).cuda()
    T.set_counter_seed(77)
    op = T.sample_masks(time_mask=(6, 10), freq_mask=(8, 6), level_jitter=0.2)         # trainv2.py:134-138
    out, y = op(x, None)
    assert out.data_ptr() != x.data_ptr() and float(x.min()) >= 1.0
    masked = out == 0
    assert bool(masked.any()) and float(masked.float().mean()) < 0.6
    d = (out - x)[~masked]
    assert float(d.abs().max()) < 1.5
    for b in range(8):                                                             # per sample: one offset on channels 0..3, none on 4..6
        keep = ~masked[b]
        lo = (out[b] - x[b])[..., :4][keep[..., :4]]
        hi = (out[b] - x[b])[..., 4:][keep[..., 4:]]
        assert float(hi.abs().max()) == 0.0
        assert float(lo.max() - lo.min()) < 1e-6


def test_seeded_sample_masks_draw_fresh_bands_every_batch():
    """With an explicit seed the transform keeps its own running sample index: consecutive batches (and epochs) get
    different bands and jitter, and a second transform with the same seed reproduces the sequence."""
    x = (torch.rand(8, 300, 64, 7) + 1.0).cuda()
    runs = []
    for _ in range(2):
        op = T.sample_masks(time_mask=(24, 1), freq_mask=(16, 1), seed=1, level_jitter=0.2)
        runs.append([op(x, None)[0] for _ in range(3)])
    a, b, c = runs[0]
    assert not torch.equal(a == 0, b == 0) and not torch.equal(b == 0, c == 0)
    assert not torch.equal(a[..., :4][(a != 0)[..., :4]][:100], b[..., :4][(b != 0)[..., :4]][:100])
    for first, second in zip(*runs):
        assert torch.equal(first, second)
    # batch k of the transform == one call on samples [8k, 8k + 8) of the stream
    want = x.clone()
    T.mask_batch_(want, (24, 1), (16, 1), seed=1, sample_offset=8)
    assert torch.equal((want == 0), (b == 0))


@pytest.mark.parametrize('spatial,c', [('foa', 7), ('acs', 17), (None, 7), (None, 10)])
def test_fused_augment_batch_equals_the_separate_calls(spatial, c):
    """seld_augment_batch draws on the device from the same Philox streams as the host-side functions: one launch ==
    level jitter -> spatial augmentation -> time / frequency masks, bit for bit (and so == the oracle restatements)."""
    g = torch.Generator().manual_seed(7)
    x = (torch.rand(12, 300, 64, c, generator=g) - 0.5).cuda()
    y = (torch.rand(12, 60, 56, generator=g) - 0.5).cuda()
    seed, off = 1234, 500
    got_x, got_y, draws = T.augment_batch(x, y, spatial=spatial, level_jitter=0.2, time_mask=(6, 10), freq_mask=(8, 6), seed=seed,
                                          sample_offset=off, return_draws=True)
    want_x, _, offs = T.random_ups_and_downs(x, None, stddev=0.2, seed=seed, sample_offset=off, return_draws=True)
    want_y = y
    if spatial == 'foa':
        want_x, want_y, d = T.foa_intensity_vec_aug(want_x, y, seed=seed, sample_offset=off, return_draws=True)
        packed = d['flip'][:, 0] | (d['flip'][:, 1] << 1) | (d['flip'][:, 2] << 2) | (d['swap'] << 3)
        assert np.array_equal(draws[:, 0].cpu().numpy(), packed)
    elif spatial == 'acs':
        want_x, want_y, d = T.acs_aug(want_x, y, seed=seed, sample_offset=off, return_draws=True)
        assert np.array_equal(draws[:, 0].cpu().numpy(), d['idx'])
    want_x = want_x.clone()
    T.mask_batch_(want_x, (6, 10), (8, 6), seed=seed, sample_offset=off)
    assert np.array_equal(draws[:, 1].cpu().numpy().view(np.float32), offs)       # Box-Muller: device float64 == numpy float64
    assert torch.equal(got_x, want_x)
    assert torch.equal(got_y, want_y)
    assert bool((got_x == 0).any()) and x.data_ptr() != got_x.data_ptr()


def test_fused_augment_masks_only_and_transform_wrapper():
    x = (torch.rand(8, 300, 64, 7) + 1.0).cuda()
    a, _ = T.augment_batch(x, None, time_mask=(24, 1), freq_mask=(16, 1), seed=3, sample_offset=16)
    b = x.clone()
    T.mask_batch_(b, (24, 1), (16, 1), seed=3, sample_offset=16)
    assert torch.equal(a, b)
    c, _ = T.augment_batch(x, None, seed=3)                                        # nothing switched on: a copy
    assert torch.equal(c, x) and c.data_ptr() != x.data_ptr()
    op = T.batch_augment(spatial='foa', time_mask=(24, 1), freq_mask=(16, 1), seed=9)
    y = torch.rand(8, 60, 56).cuda()
    x1, y1 = op(x, y)
    x2, y2 = op(x, y)
    assert not torch.equal(x1 == 0, x2 == 0)                                       # the running sample index advances
    want, wy = T.augment_batch(x, y, spatial='foa', time_mask=(24, 1), freq_mask=(16, 1), seed=9, sample_offset=8)
    assert torch.equal(x2, want) and torch.equal(y2, wy)
    with pytest.raises(ValueError):
        T.augment_batch(x, y, spatial='acs')                                       # acs needs 17 channels


def test_device_level_jitter_matches_its_restatement():
    """Box-Muller in float64 on the device vs numpy: the same float32 up to the last bit of a float64 transcendental."""
    x = torch.zeros(512, 1, 1, 7, device='cuda')
    out, _, offs = T.random_ups_and_downs(x, None, stddev=0.2, seed=77, sample_offset=1000, return_draws=True)
    want = A.level_offsets_ref(77, 1000, 512, 0.2)
    assert offs.dtype == np.float32 and np.abs(offs - want).max() <= 1.2e-7 and (offs == want).mean() > 0.9
    assert np.array_equal(out[:, 0, 0, 0].cpu().numpy(), offs) and float(out[..., 4:].abs().max()) == 0.0
