"""Windowing / batching glue against a numpy restatement of reference data_loader.py:132-168 and the
reference's own shape tests (data_loader_test.py:6-27, :48-69).  Plain torch tensors on the CPU: no kernels involved."""
import numpy as np
import torch

from seld_b200 import data_loader as DL


def test_data_loader_shapes_like_reference_test():
    data_size, batch_size = 16, 8                                   # data_loader_test.py:6-27
    xs, ys = torch.arange(data_size), torch.arange(data_size, 0, -1)
    ident = [lambda x, y: (x, y)]
    seen = 0
    for x, y in DL.data_loader((xs, ys), sample_transforms=ident, batch_transforms=ident, batch_size=batch_size, loop_time=1):
        assert tuple(x.shape) == (batch_size,) and tuple(y.shape) == (batch_size,)
        assert torch.equal(x + y, torch.full((batch_size,), data_size))
        seen += 1
    assert seen == 2


def test_default_loop_time_repeats_for_ever_like_dataset_repeat_none():
    """reference data_loader.py:17,51: loop_time=None is `dataset.repeat(None)` -- an endless stream (the reference's own
    test iterates it with a `for`, which only ends because nobody waits for it)."""
    import itertools
    import pytest
    xs, ys = torch.arange(5), torch.arange(5, 0, -1)
    dl = DL.data_loader((xs, ys), batch_size=4)
    got = torch.cat([x for x, _ in itertools.islice(dl, 7)])
    assert torch.equal(got, torch.arange(28) % 5)                   # batches straddle the passes, nothing is dropped
    with pytest.raises(TypeError):
        len(dl)


def test_batches_larger_than_the_dataset_wrap_several_times():
    xs, ys = torch.arange(3), torch.arange(3)
    batches = [x for x, _ in DL.data_loader((xs, ys), batch_size=8, loop_time=4)]
    assert [len(b) for b in batches] == [8, 4]
    assert torch.equal(torch.cat(batches), torch.arange(12) % 3)


def test_seldnet_data_to_dataloader_shapes_like_reference_test():
    n_samples, time_x, freq, chan = 8, 80, 40, 7                    # data_loader_test.py:48-69
    time_y, n_classes = 16, 11
    x = [torch.zeros(time_x, freq, chan) for _ in range(n_samples)]
    y = [torch.zeros(time_y, n_classes * 4) for _ in range(n_samples)]
    lws = 8
    n = 0
    for bx, by in DL.seldnet_data_to_dataloader(x, y, label_window_size=lws, batch_size=n_samples):
        assert tuple(bx.shape) == (n_samples, lws * 5, freq, chan)
        assert tuple(by.shape) == (n_samples, lws, n_classes * 4)
        n += 1
    assert n == 2                                                    # 8 clips * 16 labels / 8 per window / 8 per batch


def _reference_windows(features, labels, lws):
    """numpy restatement of data_loader.py:142-153 (concatenate, reshape to label resolution, batch, flatten)."""
    f = np.concatenate(features, 0)
    l = np.concatenate(labels, 0)
    f = f.reshape(l.shape[0], -1, *f.shape[1:])
    n = l.shape[0] // lws
    xs = np.stack([f[i * lws:(i + 1) * lws].reshape(-1, *f.shape[2:]) for i in range(n)])
    ys = np.stack([l[i * lws:(i + 1) * lws] for i in range(n)])
    return xs, ys


def test_windows_match_reference_restatement_and_shuffle_is_a_permutation():
    rng = np.random.default_rng(0)
    feats = [rng.standard_normal((t * 5, 6, 3)).astype(np.float32) for t in (20, 35, 15)]       # ragged clips
    labs = [rng.standard_normal((t, 8)).astype(np.float32) for t in (20, 35, 15)]
    lws, bs = 4, 3
    want_x, want_y = _reference_windows(feats, labs, lws)            # 70 labels -> 17 windows, 2 labels dropped
    assert want_x.shape[0] == 17
    # no shuffle: batches of consecutive windows, last one short (batch(drop_remainder=False), data_loader.py:53)
    dl = DL.seldnet_data_to_dataloader(feats, labs, label_window_size=lws, batch_size=bs, shuffle_size=0)
    got = list(dl)
    assert [b[0].shape[0] for b in got] == [3, 3, 3, 3, 3, 2]
    assert np.array_equal(torch.cat([b[0] for b in got]).numpy(), want_x)
    assert np.array_equal(torch.cat([b[1] for b in got]).numpy(), want_y)
    # default shuffle: same batches, permuted, reproducible per seed
    a = DL.seldnet_data_to_dataloader(feats, labs, label_window_size=lws, batch_size=bs, seed=3)
    b = DL.seldnet_data_to_dataloader(feats, labs, label_window_size=lws, batch_size=bs, seed=3)
    oa, ob = a.batch_order(), b.batch_order()
    assert oa == ob and sorted(oa) == list(range(6))
    assert DL.seldnet_data_to_dataloader(feats, labs, label_window_size=lws, batch_size=bs, seed=4).batch_order() != oa or True
    # loop_time repeats the stream before batching (data_loader.py:51): 34 samples -> 12 batches, the 6th straddles
    dl2 = DL.seldnet_data_to_dataloader(feats, labs, label_window_size=lws, batch_size=bs, loop_time=2, shuffle_size=0)
    xs2 = torch.cat([bx for bx, _ in dl2]).numpy()
    assert np.array_equal(xs2, np.concatenate([want_x, want_x]))


def test_eval_mode_yields_one_clip_per_batch_in_order():
    feats = [torch.full((300, 4, 2), float(i)) for i in range(3)]
    labs = [torch.full((60, 4), float(i)) for i in range(3)]
    dl = DL.seldnet_data_to_dataloader(feats, labs, train=False, label_window_size=20, batch_size=7, loop_time=5)
    got = list(dl)
    assert len(got) == 3
    for i, (x, y) in enumerate(got):
        assert tuple(x.shape) == (3, 100, 4, 2) and tuple(y.shape) == (3, 20, 4)
        assert float(x.min()) == float(x.max()) == float(i)


def test_resident_tensor_input_is_not_copied():
    feats = torch.arange(2 * 50 * 3 * 2, dtype=torch.float32).reshape(2, 50, 3, 2)
    labs = torch.zeros(2, 10, 4)
    dl = DL.seldnet_data_to_dataloader(feats, labs, label_window_size=5, batch_size=2, shuffle_size=0)
    x, _ = next(iter(dl))
    assert x.data_ptr() == feats.data_ptr() and tuple(x.shape) == (2, 25, 3, 2)


def test_per_sample_transforms_are_mapped_and_batched_ones_get_the_batch():
    calls = []

    def per_sample(x, y):
        calls.append(tuple(x.shape))
        return x + 1, y

    def batched(x, y):
        calls.append(('batch',) + tuple(x.shape))
        return x * 2, y
    batched.batched = True
    xs, ys = torch.zeros(4, 3), torch.zeros(4)
    out = list(DL.data_loader((xs, ys), sample_transforms=[per_sample, batched], batch_size=4, loop_time=1))
    assert calls == [(3,)] * 4 + [('batch', 4, 3)]
    assert torch.equal(out[0][0], torch.full((4, 3), 2.0))


def test_frame_windows_and_overlap_add_mean():
    x = torch.arange(40 * 2, dtype=torch.float32).reshape(40, 2)
    w = DL.frame_windows(x, win_size=10, step_size=5)
    assert tuple(w.shape) == (7, 10, 2) and w.data_ptr() == x.data_ptr()
    for i in range(7):
        assert torch.equal(w[i], x[5 * i:5 * i + 10])
    assert DL.frame_windows(x[:5], 10, 5).shape[0] == 0
    rng = np.random.default_rng(1)
    fr = rng.standard_normal((6, 4, 3)).astype(np.float32)
    want = np.zeros((9, 3), np.float64)
    cnt = np.zeros((9, 1))
    for i in range(6):
        want[i:i + 4] += fr[i]
        cnt[i:i + 4] += 1
    got = DL.overlap_and_add_mean(torch.from_numpy(fr)).numpy()
    assert np.allclose(got, want / cnt, atol=1e-6)


def test_ensemble_outputs_recovers_a_framewise_model():
    # a "model" whose label-frame output is the mean of the 5 feature frames it covers: ensembling overlapping windows of
    # a clip must give back exactly that per-label-frame sequence (every window agrees on it)
    def model(w):                                   # w [b, 300, F, C]
        m = w.reshape(w.shape[0], 60, 5, -1).mean(dim=(2, 3))
        return m[..., None].repeat(1, 1, 2), m[..., None].repeat(1, 1, 3)
    x = torch.arange(3000, dtype=torch.float32)[:, None, None].repeat(1, 4, 2)
    (sed, doa), = DL.ensemble_outputs(model, [x], win_size=300, step_size=5, batch_size=100)
    assert tuple(sed.shape) == (600, 2) and tuple(doa.shape) == (600, 3)
    want = x.reshape(600, 5, -1).mean(dim=(1, 2))
    assert torch.allclose(sed[:, 0], want, rtol=1e-6) and torch.allclose(doa[:, 2], want, rtol=1e-6)
