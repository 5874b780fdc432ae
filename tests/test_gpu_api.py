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


@pytest.mark.parametrize('mode', ['foa', 'mic'])
def test_dataset_step_eager_and_cuda_graph_match_the_plain_calls(mode):
    """pipeline.DatasetStep (static buffers; the whole extract -> statistics -> normalise step as ONE CUDA graph) gives the
    same bits as the individual calls, replay after replay."""
    from seld_b200 import pipeline
    from seld_b200.synth import make_clips
    wav = make_clips(range(40, 46), 480 * 120).cuda()
    want, mean, std = pipeline.extract_normalized_dataset(wav, 24000, mode=mode, t_out=100, **PROD)
    step = pipeline.DatasetStep(wav, 24000, mode=mode, t_out=100, **PROD)
    feat, m, s = step.run()
    assert torch.equal(feat, want) and torch.equal(m, mean) and torch.equal(s, std)
    step.capture()
    for _ in range(3):
        feat.zero_()
        feat, m, s = step.replay()
        torch.cuda.synchronize()
        assert torch.equal(feat, want) and torch.equal(m, mean) and torch.equal(s, std)


@pytest.mark.parametrize('dtype', ['float32', 'int16'])
def test_host_dataset_extractor_pipelined_submits(dtype):
    """End-to-end form with pinned host buffers: one synchronous run == the device-resident step, and two datasets in
    flight (upload of the second overlapping the download of the first) both arrive intact."""
    from seld_b200 import pipeline
    from seld_b200.synth import make_clips
    n, L = 7, 480 * 90
    wav = make_clips(range(60, 60 + n), L)
    if dtype == 'int16':
        pcm = torch.clamp(torch.round(wav * 32768.0), -32768, 32767).to(torch.int16)
        host_in = pcm.transpose(1, 2).contiguous().pin_memory()
        dev_in = (pcm.to(torch.float32) / 32768.0).cuda()
        ex = pipeline.HostDatasetExtractor(n, L, 24000, mode='foa', t_out=80, chunk_clips=3, dtype=torch.int16, **PROD)
    else:
        host_in = wav.pin_memory()
        dev_in = wav.cuda()
        ex = pipeline.HostDatasetExtractor(n, L, 24000, mode='foa', t_out=80, chunk_clips=3, **PROD)
    want, mean, std = pipeline.extract_normalized_dataset(dev_in, 24000, mode='foa', t_out=80, **PROD)
    outs = [torch.empty(n, 80, 64, 7).pin_memory() for _ in range(3)]
    m, s = ex.run(host_in, outs[0])
    assert torch.equal(outs[0], want.cpu()) and torch.equal(m, mean) and torch.equal(s, std)
    host_b = (host_in // 2) if dtype == 'int16' else (host_in * 0.5)
    host_b = host_b.contiguous().pin_memory()
    pend = [ex.submit(host_in, outs[1]), ex.submit(host_b, outs[2])]
    for _, _, done in pend:
        done.synchronize()
    assert torch.equal(outs[1], want.cpu())
    dev_b = (host_b.to(torch.float32) / 32768.0).transpose(1, 2).contiguous().cuda() if dtype == 'int16' else host_b.cuda()
    want_b, _, _ = pipeline.extract_normalized_dataset(dev_b, 24000, mode='foa', t_out=80, **PROD)
    assert torch.equal(outs[2], want_b.cpu())
