"""top_db clamp, dataset statistics (float64, deterministic) and normalisation vs the oracle, through the C ABI."""
import os

import numpy as np
import pytest
import torch

from cases import PROD

pytestmark = pytest.mark.gpu


def _features(mode, n=6, L=40000, t_out=None):
    from seld_b200 import pipeline
    from seld_b200.synth import make_clips
    wav = make_clips(range(50, 50 + n), L)
    feat, key = pipeline.extract_batch(wav.cuda(), 24000, mode=mode, t_out=t_out, **PROD)
    return wav, feat, key, 1 + L // 480


@pytest.mark.parametrize('mode', ['foa', 'mic'])
def test_statistics_and_normalise_vs_float64_oracle(mode):
    from oracle import extractor as O
    from seld_b200 import pipeline
    wav, feat, key, t_raw = _features(mode, t_out=90)                 # 84 real frames + 6 zero rows
    ref = np.stack([O.preprocess_features_port(O.extract_features_port(w, 24000, mode=mode, **PROD), 90, 1) for w in wav])
    acc = pipeline.partial_statistics(feat, key, t_raw)
    mean, std = pipeline.finish_statistics(acc, 64, feat.shape[3])
    m64, s64 = O.statistics_f64(ref)
    # vs float64 oracle: <= 1e-5 relative; inputs differ by the kernel's own <= 1e-4 feature error
    assert np.abs(mean.cpu().numpy() - m64).max() <= 2e-5 * max(1.0, np.abs(m64).max())
    assert np.abs(std.cpu().numpy() - s64).max() <= 2e-5 * max(1.0, np.abs(s64).max())
    assert mean.shape == (1, 64, feat.shape[3]) and mean.dtype == torch.float32
    # the reference's own float32 numpy statistics are within its documented ~5e-3 of ours
    m32, s32 = O.statistics_port(ref)
    assert np.abs(mean.cpu().numpy() - m32).max() <= 5e-3 and np.abs(std.cpu().numpy() - s32).max() <= 5e-3
    # run-to-run determinism (fixed reduction order, no float atomics)
    acc2 = pipeline.partial_statistics(feat, key, t_raw)
    assert torch.equal(acc, acc2)
    # normalise with IDENTICAL mean/std fed to both sides
    pipeline.finalize_(feat, key, t_raw, mean, std)
    want = O.normalize_port(ref, mean.cpu().numpy(), std.cpu().numpy())
    got = feat.cpu().numpy()
    tol = 1e-4 / np.maximum(std.cpu().numpy(), 1e-8) + 1e-5
    assert np.all(np.abs(got - want)[..., :4] <= tol[..., :4] * 1.5)
    assert np.all(np.abs(got - want)[..., 4:] <= 1e-3 / np.maximum(std.cpu().numpy()[..., 4:], 1e-8) + 1e-5)


def test_clamp_only_and_zero_rows():
    from oracle import extractor as O
    from seld_b200 import pipeline
    wav, feat, key, t_raw = _features('foa', n=3, t_out=100)
    raw = feat.clone()
    pipeline.finalize_(feat, key, t_raw)
    got = feat.cpu().numpy()
    cmax = pipeline.clip_max_db(key).cpu().numpy()
    r = raw.cpu().numpy()
    for i in range(3):
        want = r[i].copy()
        want[:t_raw, :, :4] = np.maximum(want[:t_raw, :, :4], cmax[i] - 80.0)
        assert np.array_equal(got[i], want)                       # clamp is exact; IV channels and pad rows untouched
        assert np.all(got[i, t_raw:] == 0.0)
    # out-of-place form leaves the input alone
    out = torch.empty_like(raw)
    pipeline.finalize_(raw, key, t_raw, out=out)
    assert torch.equal(out, feat) and torch.equal(raw.cpu(), torch.from_numpy(r))


def test_reference_file_api_statistics_and_normalizer(tmp_path, golden_dir):
    """calculate_statistics / apply_normalizer (reference feature_extractor.py:218-234) on .npy folders, vs the outputs
    of the reference's own functions (tests/golden/stats_norm.npz)."""
    from seld_b200 import feature_extractor as fe
    g = np.load(os.path.join(golden_dir, 'stats_norm.npz'))
    src, dst = tmp_path / 'feat', tmp_path / 'norm'
    src.mkdir()
    for i, c in enumerate(g['clips']):
        np.save(src / f'fold1_room1_mix00{i}.npy', c)
    mean, std = fe.calculate_statistics(str(src))
    assert mean.shape == g['mean'].shape == (1, 8, 7) and mean.dtype == np.float32
    assert np.abs(mean - g['mean']).max() <= 1e-5 and np.abs(std - g['std']).max() <= 1e-5
    fe.apply_normalizer(str(src), str(dst), g['mean'], g['std'])
    for i in range(3):
        got = np.load(dst / f'fold1_room1_mix00{i}.npy')
        assert got.shape == (40, 8, 7) and np.abs(got - g['normed'][i]).max() <= 1e-5
    with pytest.raises(ValueError):
        fe.calculate_statistics(str(tmp_path / 'empty'))


def test_full_pipeline_matches_reference_main():
    """extract -> statistics -> normalise for a small shard == the reference's __main__ arithmetic on the same clips."""
    from oracle import extractor as O
    from seld_b200 import pipeline
    from seld_b200.synth import make_clips
    wav = make_clips(range(900, 904), 60000)
    feat, mean, std = pipeline.extract_normalized_dataset(wav.cuda(), 24000, mode='foa', t_out=120, **PROD)
    ref = np.stack([O.preprocess_features_port(O.extract_features_port(w, 24000, mode='foa', **PROD), 120, 1) for w in wav])
    m64, s64 = O.statistics_f64(ref)
    want = O.normalize_port(ref, m64, s64)
    got = feat.cpu().numpy()
    assert got.shape == (4, 120, 64, 7)
    assert np.abs(got[..., :4] - want[..., :4]).max() <= 1e-4 / s64[..., :4].min() + 1e-4
    assert np.abs(got[..., 4:] - want[..., 4:]).max() <= 1e-3 / s64[..., 4:].min() + 1e-4


def test_peer_allreduce_kernel_single_rank_and_repeated_epochs():
    """seld_stats_peer_allreduce with one rank (its own buffer is the only peer): publish -> flag -> wait -> rank-ordered sum must
    return the input unchanged, epoch after epoch (two alternating slots), also when replayed inside a CUDA graph.  The real
    multi-GPU exchange is checked by bench.py's stats_check at N > 1 against single-GPU statistics."""
    import torch
    from seld_b200 import pipeline
    n = 2 * 64 * 10 + 1
    red = pipeline.PeerStatisticsAllReduce(n)
    for it in range(5):
        acc = torch.arange(n, dtype=torch.float64, device='cuda') * (it + 1) + 0.25
        want = acc.clone()
        red(acc)
        assert torch.equal(acc, want)
    acc = torch.full((n,), 3.5, dtype=torch.float64, device='cuda')
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        red(acc)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        red(acc)
    for _ in range(4):
        g.replay()
    torch.cuda.synchronize()
    assert bool((acc == 3.5).all())
    # DatasetStep with the peer exchange == the plain step
    from seld_b200.synth import make_clips
    wav = make_clips(range(5), 480 * 60).cuda()
    a = pipeline.DatasetStep(wav, 24000, mode='foa', t_out=50, win_length=960, hop_length=480, n_fft=1024)
    b = pipeline.DatasetStep(wav, 24000, mode='foa', t_out=50, peer_allreduce=True, win_length=960, hop_length=480, n_fft=1024)
    fa, ma, sa = a.run()
    fb, mb, sb = b.run()
    assert torch.equal(fa, fb) and torch.equal(ma, mb) and torch.equal(sa, sb)
