"""The oracle is pinned: both restatements against the committed outputs of the UNMODIFIED reference
(tests/golden, made by oracle/make_golden.py) and against the live reference when /root/reference exists."""
import numpy as np
import pytest
import torch

from cases import CASES, case_input, input_matches_golden, load_golden
from oracle import extractor as O
from oracle.ref_shim import load_reference_extractor, reference_available


@pytest.mark.parametrize('name', list(CASES))
@pytest.mark.parametrize('mode', ['foa', 'mic'])
def test_port_equals_golden(name, mode):
    wav, sr, n_mels, kw = case_input(name)
    g = load_golden(name)
    if not input_matches_golden(wav, g):
        pytest.skip('torch RNG stream differs from the one the fixture was generated with')
    got = O.extract_features_port(wav, sr, mode=mode, n_mels=n_mels, **kw)
    assert got.shape == g[mode].shape and got.dtype == np.float32
    # same library kernels as the reference -> identical up to summation-order noise
    assert np.abs(got - g[mode]).max() <= 2e-5


@pytest.mark.parametrize('name', ['prod', 'default', 'ragged', 'nfft256', 'nfft2048', 'loud', 'zeros'])
@pytest.mark.parametrize('mode', ['foa', 'mic'])
def test_float64_restatement_vs_golden(name, mode):
    """Independent method (numpy float64, explicit framing): bounds the reference's own float32 noise floor."""
    wav, sr, n_mels, kw = case_input(name)
    g = load_golden(name)
    if not input_matches_golden(wav, g):
        pytest.skip('torch RNG stream differs')
    got = O.extract_features_f64(wav.numpy(), sr, mode=mode, n_mels=n_mels, **kw)
    assert got.shape == g[mode].shape
    assert np.abs(got[..., :4] - g[mode][..., :4]).max() <= 1e-4
    assert np.abs(got[..., 4:] - g[mode][..., 4:]).max() <= 1e-3


def test_reference_known_answers_from_its_own_test_input():
    """reference feature_extractor_test.py:24-34: zeros(4, 32000) @ 16 kHz, default kwargs."""
    g = load_golden('zeros')
    assert g['foa'].shape == (126, 64, 7) and g['mic'].shape == (126, 64, 10)
    assert np.all(g['foa'][..., :4] == -100.0) and np.all(g['foa'][..., 4:] == 0.0)
    assert np.allclose(g['mic'][:, 32, 4:], 1.0, atol=1e-6)
    z = torch.zeros(4, 32000)
    for mode in ('foa', 'mic'):
        assert np.array_equal(O.extract_features_port(z, 16000, mode=mode), g[mode])


def test_sub_stage_goldens():
    for name in ('prod', 'default'):
        wav, sr, n_mels, kw = case_input(name)
        g = load_golden(name)
        if not input_matches_golden(wav, g):
            pytest.skip('torch RNG stream differs')
        spec = O.complex_spec_port(wav, **kw)
        assert tuple(spec.shape) == tuple(g['spec_shape'])
        assert np.abs(spec.numpy() - g['spec']).max() <= 1e-4 * max(1.0, np.abs(g['spec']).max())
        assert np.abs(O.foa_intensity_vectors_port(spec).numpy() - g['iv']).max() <= 2e-3
        assert np.abs(O.gcc_features_port(spec, n_mels).numpy() - g['gcc']).max() <= 1e-5
        s64 = O.complex_spec_f64(wav.numpy(), **kw)
        assert np.abs(s64 - g['spec']).max() <= 1e-4 * max(1.0, np.abs(g['spec']).max())


def test_stats_normalizer_pad_goldens(golden_dir):
    import os
    g = np.load(os.path.join(golden_dir, 'stats_norm.npz'))
    clips = g['clips']                                       # [3, 40, 8, 7]
    mean64, std64 = O.statistics_f64(clips)
    assert np.abs(mean64 - g['mean']).max() <= 1e-5 and np.abs(std64 - g['std']).max() <= 1e-5
    mean32, std32 = O.statistics_port(clips)
    assert np.array_equal(mean32, g['mean']) and np.array_equal(std32, g['std'])
    for i in range(3):
        assert np.array_equal(O.normalize_port(clips[i], g['mean'], g['std']), g['normed'][i])
    assert np.array_equal(O.preprocess_features_port(g['feats'], 4, 5), g['f_pad'])
    assert np.array_equal(O.preprocess_features_port(g['feats'], 2, 5), g['f_cut'])


@pytest.mark.skipif(not reference_available(), reason='/root/reference only exists in the authoring container')
@pytest.mark.parametrize('mode', ['foa', 'mic'])
def test_port_equals_live_reference(mode):
    fe = load_reference_extractor()
    wav, sr, n_mels, kw = case_input('prod')
    ref = fe.extract_features(wav, sr, mode=mode, n_mels=n_mels, **kw)
    assert np.array_equal(O.extract_features_port(wav, sr, mode=mode, n_mels=n_mels, **kw), np.ascontiguousarray(ref))
    with pytest.raises(ValueError):
        fe.extract_features(wav, sr, mode='bad')


def test_mel_table_is_torchaudio_bit_for_bit():
    torchaudio = pytest.importorskip('torchaudio')
    from seld_b200.melscale import melscale_fbanks_htk, sparsify
    for n_freqs, sr, n_mels in [(513, 24000, 64), (257, 16000, 64), (129, 8000, 32), (1025, 48000, 64), (1025, 48000, 128)]:
        ours = melscale_fbanks_htk(n_freqs, sr, n_mels)
        theirs = torchaudio.functional.melscale_fbanks(n_freqs, 0.0, float(sr // 2), n_mels, sr, norm=None, mel_scale='htk')
        assert torch.equal(ours, theirs)
        seg, w0, w1 = sparsify(ours.numpy())
        dense = np.zeros_like(ours.numpy())
        for k in range(n_freqs):
            if seg[k] >= 0:
                dense[k, seg[k]] = w0[k]
                if w1[k] != 0:
                    dense[k, seg[k] + 1] = w1[k]
        assert np.array_equal(dense, ours.numpy())
        assert np.all(np.diff(seg[seg >= 0]) >= 0)
