"""The TensorFlow-variant FOA extractor (reference data_loader.py:310-349) on the GPU against the numpy restatement
(oracle/tf_variant.py -- PARITY UNPINNED: TensorFlow / tensorflow_io are absent; the restatement follows their published
algorithms).  Tolerances: log-mel 2e-4 dB above the clamp floor (20 log10 of a magnitude sum doubles the sensitivity of the
power path's 10 log10), IV 1e-3, frame count / padding / clamp exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _check(got, want, what):
    assert got.shape == want.shape, (what, got.shape, want.shape)
    e_mel = np.abs(got[..., :4].astype(np.float64) - want[..., :4])
    e_mel = e_mel[np.isfinite(want[..., :4])]
    assert e_mel.max() <= 2e-4, (what, e_mel.max())
    assert np.array_equal(np.isneginf(got[..., :4]), np.isneginf(want[..., :4])), what
    assert np.abs(got[..., 4:] - want[..., 4:]).max() <= 1e-3, (what, np.abs(got[..., 4:] - want[..., 4:]).max())


@pytest.mark.parametrize('n_samples', [480 * 40, 480 * 40 + 133, 1500, 700])
def test_tf_variant_against_the_restatement(n_samples):
    """Ragged lengths: the tail frames read zeros past the end (pad_end=True); 700 samples = no whole frame at all."""
    from oracle import tf_variant as O
    from seld_b200 import data_loader as DL
    from seld_b200.synth import make_clips
    wav = make_clips([301, 302, 303], n_samples)
    got = DL.get_preprocessed_x_tf(wav, 24000, max_label_length=12, multiplier=5)
    assert got.shape == (3, 60, 64, 7) and got.is_cuda
    for i in range(3):
        want = O.get_preprocessed_x_tf(wav[i].numpy().astype(np.float64), 24000, max_label_length=12, multiplier=5)
        _check(got[i].cpu().numpy(), want, f'L={n_samples} clip {i}')
    one = DL.get_preprocessed_x_tf(wav[1], 24000, max_label_length=12, multiplier=5)
    assert torch.equal(one, got[1])


def test_tf_variant_full_clip_and_silence():
    from oracle import tf_variant as O
    from seld_b200 import data_loader as DL, pipeline
    from seld_b200.synth import make_clip
    wav = make_clip(1000)                                            # 60 s: 3000 frames, 10 s at 1e-6 (clamp active), 1 s of exact zeros
    got = DL.get_preprocessed_x_tf(wav, 24000).cpu().numpy()
    want = O.get_preprocessed_x_tf(wav.numpy().astype(np.float64), 24000)
    assert got.shape == (3000, 64, 7)
    _check(got, want, 'full clip')
    floor = want[..., :4].max() - 80.0
    assert abs(got[..., :4].min() - floor) <= 2e-4 and (got[..., :4] <= floor + 2e-4).mean() > 0.05
    # an all-zero clip: log(0) = -inf survives the clamp (max - 80 = -inf), IV = 0 -- as tfio's dbscale gives
    z = DL.get_preprocessed_x_tf(torch.zeros(4, 4800), 24000, max_label_length=4, multiplier=5).cpu().numpy()
    assert np.isneginf(z[:10, :, :4]).all() and not z[:10, :, 4:].any() and not z[10:].any()
    with pytest.raises(ValueError):
        DL.get_preprocessed_x_tf(wav, 24000, mode='mic')
    # interleaved input gives the same bits
    w3 = wav[:, :48000].unsqueeze(0).cuda()
    a, ka = pipeline.extract_batch_tf(w3, 24000)
    b, kb = pipeline.extract_batch_tf(w3.transpose(1, 2).contiguous(), 24000, layout='interleaved')
    assert torch.equal(a, b) and torch.equal(ka, kb)
