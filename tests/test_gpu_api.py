"""Stand-alone public helpers and the file-level driver of the reference surface, on the GPU."""
import os
import struct

import numpy as np
import pytest
import torch

from cases import PROD, case_input, check_features, input_matches_golden, load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name', ['prod', 'default'])
def test_complex_spec_iv_gcc_vs_reference_golden(name):
    from seld_b200 import feature_extractor as fe
    wav, sr, n_mels, kw = case_input(name)
    g = load_golden(name)
    if not input_matches_golden(wav, g):
        pytest.skip('torch RNG stream differs')
    spec = fe.complex_spec(wav, **kw)
    assert spec.is_cuda and torch.is_complex(spec) and tuple(spec.shape) == tuple(g['spec_shape'])
    scale = max(1.0, np.abs(g['spec']).max())
    assert np.abs(spec.cpu().numpy() - g['spec']).max() <= 2e-6 * scale * 32
    iv = fe.foa_intensity_vectors(torch.from_numpy(g['spec']))       # same spectra in -> isolates the IV kernel
    assert tuple(iv.shape) == g['iv'].shape and np.abs(iv.cpu().numpy() - g['iv']).max() <= 1e-5
    gcc = fe.gcc_features(torch.from_numpy(g['spec']), n_mels)
    assert tuple(gcc.shape) == g['gcc'].shape and np.abs(gcc.cpu().numpy() - g['gcc']).max() <= 1e-5
    # chained through our own spectra as the reference chains them
    assert np.abs(fe.gcc_features(spec, n_mels).cpu().numpy() - g['gcc']).max() <= 1e-3


def test_complex_spec_variants():
    from oracle import extractor as O
    from seld_b200 import feature_extractor as fe
    wav = case_input('default')[0]
    for kw in (dict(n_fft=512, normalized=True), dict(n_fft=256, pad=100), dict(n_fft=1024, win_length=600, hop_length=200)):
        got = fe.complex_spec(wav[:3], **kw).cpu().numpy()            # odd channel count
        want = O.complex_spec_port(wav[:3], **kw).numpy()
        assert got.shape == want.shape and np.abs(got - want).max() <= 1e-4 * max(1.0, np.abs(want).max())
    got = fe.extract_features(wav, 16000, mode='foa', n_fft=512, normalized=True, pad=64)
    want = O.extract_features_port(wav, 16000, mode='foa', n_fft=512, normalized=True, pad=64)
    check_features(got, want, 'foa', 'normalized+pad')
    with pytest.raises(ValueError):
        fe.complex_spec(wav, n_fft=300)                               # unsupported FFT size is an error, not a fallback


def _write_wav(path, x, rate):
    pcm = np.clip(np.round(x.T * 32768.0), -32768, 32767).astype('<i2')
    body = pcm.tobytes()
    with open(path, 'wb') as fh:
        fh.write(b'RIFF' + struct.pack('<I', 36 + len(body)) + b'WAVE' + b'fmt ' +
                 struct.pack('<IHHIIHH', 16, 1, x.shape[0], rate, rate * 2 * x.shape[0], 2 * x.shape[0], 16))
        fh.write(b'data' + struct.pack('<I', len(body)) + body)
    return torch.from_numpy(pcm.T.astype(np.float32) / 32768.0)


@pytest.mark.parametrize('mode', ['foa', 'mic'])
def test_extract_seldnet_data_end_to_end(tmp_path, mode):
    """reference feature_extractor.py:15-50 + :304-307 on a tiny synthetic dataset: on-disk contract and values."""
    from oracle import extractor as O
    from seld_b200 import feature_extractor as fe
    from seld_b200.synth import make_clip
    wdir, ldir = tmp_path / 'wav', tmp_path / 'meta'
    wdir.mkdir(); ldir.mkdir()
    wavs = {}
    for i in range(2):
        name = f'fold{i + 1}_room1_mix00{i}'
        wavs[name] = _write_wav(wdir / f'{name}.wav', make_clip(40 + i, 24000 * 3).numpy(), 24000)
        (ldir / f'{name}.csv').write_text(f'0,{i},0,10,5\n12,3,0,-30,20\n')
    fo, lo, no = tmp_path / f'{mode}_dev', tmp_path / f'{mode}_dev_label', tmp_path / f'{mode}_dev_norm'
    fe.extract_seldnet_data(str(wdir), str(fo), str(ldir), str(lo), mode=mode, **PROD)
    c = 7 if mode == 'foa' else 10
    for name, wav in wavs.items():
        f = np.load(fo / f'{name}.npy'); l = np.load(lo / f'{name}.npy')
        assert f.shape == (3000, 64, c) and f.dtype == np.float32 and f.flags['C_CONTIGUOUS'] and l.shape == (600, 56)
        ref = O.preprocess_features_port(O.extract_features_port(wav, 24000, mode=mode, **PROD))
        check_features(f, ref, mode, name)
    mean, std = fe.calculate_statistics(str(fo))
    fe.apply_normalizer(str(fo), str(no), mean, std)
    allf = np.concatenate([np.load(fo / f'{n}.npy') for n in sorted(wavs)], 0).astype(np.float64)
    assert np.abs(mean - allf.mean(0, keepdims=True)).max() <= 1e-4
    n0 = np.load(no / f'{sorted(wavs)[0]}.npy')
    assert n0.shape == (3000, 64, c) and abs(float(n0.mean())) < 1.0
