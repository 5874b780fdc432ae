"""The training input pipeline of reference train.py:157-165 (masks as sample transforms, FOA flip/rotate + label split as
batch transforms) running on a resident device tensor through seld_b200.data_loader."""
import numpy as np
import pytest
import torch

from oracle import augment as A
from seld_b200 import data_loader as DL, transforms as T

pytestmark = pytest.mark.gpu


def test_training_pipeline_on_resident_features():
    g = torch.Generator().manual_seed(0)
    feats = (torch.rand(4, 3000, 64, 7, generator=g) + 0.5).cuda()          # strictly positive: zeros are masks
    labs = torch.rand(4, 600, 56, generator=g).cuda()
    keep = feats.clone()
    T.set_counter_seed(123)
    dl = DL.seldnet_data_to_dataloader(
        feats, labs, label_window_size=60, batch_size=16, seed=1,
        sample_transforms=[T.sample_masks(time_mask=(24, 1), freq_mask=(16, 1))],
        batch_transforms=[lambda x, y: T.foa_intensity_vec_aug(x, y, seed=5), T.split_total_labels_to_sed_doa])
    batches = list(dl)
    assert len(batches) == 3 and sorted(dl.batch_order()) == [0, 1, 2]
    assert torch.equal(feats, keep)                                          # the resident tensor is never modified
    sizes = [b[0].shape[0] for b in batches]
    assert sorted(sizes) == [8, 16, 16]
    for x, (sed, doa) in batches:
        assert tuple(x.shape[1:]) == (300, 64, 7) and tuple(sed.shape[1:]) == (60, 14) and tuple(doa.shape[1:]) == (60, 42)
        zero_rows = (x == 0).all(dim=3).all(dim=2)                           # [B, 300] fully masked frames
        assert int(zero_rows.sum(dim=1).max()) <= 3 * 23                      # < 24 frames in each of the 3 periods
        zero_bins = (x == 0).all(dim=3)                                      # [B, 300, 64]
        assert bool(zero_bins.any())
        assert float(x.abs().max()) <= 1.5


def test_batch_transform_equals_oracle_on_loader_batches():
    g = torch.Generator().manual_seed(2)
    feats = (torch.rand(2, 600, 64, 7, generator=g) - 0.5).cuda()
    labs = (torch.rand(2, 120, 48, generator=g) - 0.5).cuda()
    got = {}

    def aug(x, y):
        nx, ny, d = T.foa_intensity_vec_aug(x, y, seed=9, return_draws=True)
        got['in'] = (x.cpu().numpy(), y.cpu().numpy())
        got['d'] = d
        return nx, ny
    dl = DL.seldnet_data_to_dataloader(feats, labs, label_window_size=60, batch_size=4, shuffle_size=0, batch_transforms=[aug])
    (x, y), = list(dl)
    ox, oy = A.foa_intensity_vec_aug_ref(got['in'][0], got['in'][1], got['d']['flip'], got['d']['swap'])
    assert np.array_equal(x.cpu().numpy(), ox) and np.array_equal(y.cpu().numpy(), oy)


def test_sliding_window_framing_on_device():
    x = torch.arange(3000 * 64 * 7, dtype=torch.float32, device='cuda').reshape(3000, 64, 7)
    w = DL.frame_windows(x, 300, 5)
    assert tuple(w.shape) == (541, 300, 64, 7) and w.data_ptr() == x.data_ptr()
    assert torch.equal(w[17], x[85:385])


@pytest.mark.parametrize('mode', ['foa', 'mic'])
def test_get_preprocessed_x_matches_oracle(mode):
    """reference data_loader.py:268-308 = extract_features + pad / truncate to max_label_length * multiplier frames."""
    from oracle import extractor as O
    from seld_b200.synth import make_clips
    from cases import PROD, check_features
    wav = make_clips(range(40, 42), 24000 * 4)                                  # 4 s -> 201 frames
    for max_label in (30, 50):                                                   # truncate to 150, pad to 250
        got = DL.get_preprocessed_x(wav[0], 24000, mode=mode, max_label_length=max_label, multiplier=5, **PROD)
        ref = O.extract_features_port(wav[0], 24000, mode=mode, **PROD)
        want = O.preprocess_features_port(ref, max_label_length=max_label, multiplier=5)
        assert tuple(got.shape) == want.shape == (max_label * 5, 64, 7 if mode == 'foa' else 10)
        check_features(got.cpu().numpy(), want, mode, f'{mode} {max_label}')
    both = DL.get_preprocessed_x(wav, 24000, mode=mode, max_label_length=30, **PROD)
    assert tuple(both.shape)[:2] == (2, 150)
    assert torch.equal(both[0], DL.get_preprocessed_x(wav[0], 24000, mode=mode, max_label_length=30, **PROD))
