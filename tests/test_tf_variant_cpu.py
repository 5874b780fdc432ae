"""The TensorFlow-variant on-the-fly path: host logic and the oracle restatement.  PARITY UNPINNED -- TensorFlow and
tensorflow_io are absent here and the reference has no test for these functions (oracle/tf_variant.py header); these tests pin
the restatement against closed forms and the product's host code against the restatement."""
import numpy as np
import pytest
import torch

from oracle import tf_variant as O
from seld_b200 import data_loader as DL, tables


def test_tf_tables_match_their_closed_forms():
    w = tables.tf_hann_window(1024)
    k = np.arange(1024)
    assert w.dtype == np.float32 and np.abs(w - (0.5 - 0.5 * np.cos(2 * np.pi * k / 1024))).max() < 2e-7 and w[0] == 0.0
    assert np.array_equal(w, O.hann_window_tf(1024))
    m = tables.tf_mel_weight_matrix(64, 513, 24000)
    assert m.shape == (513, 64) and m.dtype == np.float32 and np.array_equal(m, O.linear_to_mel_weight_matrix_tf())
    assert not m[0].any() and m[512].max() < 1e-6                       # DC row zeroed, the Nyquist bin sits on the last upper edge
    exact = O.linear_to_mel_weight_matrix_tf(dtype=np.float64)
    assert np.abs(m - exact).max() < 2e-5                               # float32 vs float64 arithmetic of the same formula
    for row in m:                                                       # <= 2 adjacent non-zeros per bin: the extractor's piece form applies
        nz = np.nonzero(row)[0]
        assert len(nz) <= 2 and (len(nz) < 2 or nz[1] - nz[0] == 1)
    # interior bins: the two weights of a bin sum to one only in MEL-linear interpolation -- check one bin by hand
    mel = lambda f: 1127.0 * np.log1p(f / 700.0)                        # noqa: E731
    edges = np.linspace(mel(0.0), mel(12000.0), 66)
    b, f = 200, 200 * 12000.0 / 512
    j = np.searchsorted(edges, mel(f)) - 1                              # mel(f) lies between edges j and j + 1
    assert abs(exact[b, j - 1] - (edges[j + 1] - mel(f)) / (edges[j + 1] - edges[j])) < 1e-12
    assert abs(exact[b, j] - (mel(f) - edges[j]) / (edges[j + 1] - edges[j])) < 1e-12


def test_stft_tf_framing_and_known_answers():
    x = np.zeros((4, 1000))
    x[:, 0] = 1.0
    s = O.stft_tf(x, 1024, 480, 1024)
    assert s.shape == (4, 3, 513)                                       # ceil(1000 / 480) frames, zero-padded tail
    assert np.abs(s[0, 0]).max() == 0.0                                 # the impulse sits on the window's zero tap
    t = np.arange(4800)
    tone = np.cos(2 * np.pi * 64 * t / 1024)[None].repeat(4, 0)
    s = O.stft_tf(tone, 1024, 480, 1024)
    assert s.shape[1] == 10 and abs(np.abs(s[0, 2, 64]) - 256.0) < 1e-4 and np.abs(s[0, 2, 70:]).max() < 1e-4
    f = O.get_preprocessed_x_tf(tone * 0.1, 24000)
    assert f.shape == (3000, 64, 7) and not f[10:].any()                # padded to max_label_length * multiplier
    assert f[:10, :, :4].max() - f[:10, :, :4].min() <= 80.0 + 1e-9      # tfio dbscale clamp
    iv = f[:8, :, 4:]                                                   # identical channels: conj(W) X is real and positive
    assert np.abs(iv[..., 0] - iv[..., 1]).max() < 1e-12


def test_gcc_features_tf_slices_frames_like_the_reference_code():
    rng = np.random.default_rng(0)
    spec = rng.standard_normal((4, 100, 513)) + 1j * rng.standard_normal((4, 100, 513))
    want = O.gcc_features_tf(spec, 64)
    got = DL.gcc_features_tf(torch.as_tensor(spec), 64).numpy()
    assert want.shape == (6, 64, 1024) and np.abs(got - want).max() < 1e-9
    iv = DL.foa_intensity_vectors_tf(torch.as_tensor(spec)).numpy()
    assert np.abs(iv - O.foa_intensity_vectors_tf(spec)).max() < 1e-12 and np.abs((iv ** 2).sum(0) - 1).max() < 1e-9
    with pytest.raises(ValueError):
        O.get_preprocessed_x_tf(np.zeros((4, 4800)), 24000, mode='mic')


def test_tdm_aug_matches_the_restated_mixing_rule():
    g = torch.Generator().manual_seed(5)
    n_cls, spf = 3, 2400
    x = [torch.randn(4, 60 * spf, generator=g) * 0.1 for _ in range(3)]
    y = [(torch.rand(60, 4 * n_cls, generator=g) < 0.15).float() for _ in range(3)]
    tdm_x = [torch.randn(4, (80 + 10 * c) * spf, generator=g) for c in range(n_cls)]
    tdm_y = [torch.rand(80 + 10 * c, 4 * n_cls, generator=g) for c in range(n_cls)]
    x0, y0 = [t.clone().numpy() for t in x], [t.clone().numpy() for t in y]
    xa, ya, draws = DL.TDM_aug(x, y, tdm_x, tdm_y, seed=3, return_draws=True)
    assert xa is x and ya is y and len(draws) == 3 and all(len(d) == 5 for d in draws)
    for d in draws:
        for cls, st, off, tdo in d:
            assert 0 <= cls < n_cls and 10 <= st < 50 and 0 <= off <= 60 - st and 0 <= tdo
    xr, yr = O.tdm_aug(x0, y0, [t.numpy() for t in tdm_x], [t.numpy() for t in tdm_y], draws)
    assert all(np.allclose(a.numpy(), b, atol=1e-6) for a, b in zip(x, xr))
    assert all(np.allclose(a.numpy(), b, atol=1e-6) for a, b in zip(y, yr))
    z = DL.normalize_over_clips(torch.stack([t[:, :100] for t in x]))
    assert float(z.mean(0).abs().max()) < 1e-5
