"""Generate tests/golden/*.npz from the UNMODIFIED reference -- TEST INFRASTRUCTURE ONLY.

Run in the authoring container (the only place /root/reference exists):

    python -m oracle.make_golden

Each fixture stores the reference's output for a seeded synthetic input plus the
sha256 of the input bytes; tests regenerate the input from the seed, check the
hash (same image => same torch RNG stream) and compare.  Inputs are not stored
(random floats do not compress).
"""
import hashlib
import os
import sys
import tempfile

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from oracle.ref_shim import load_reference_extractor  # noqa: E402
from seld_b200.synth import make_clip  # noqa: E402

PROD = dict(win_length=960, hop_length=480, n_fft=1024)

# name -> (seed or None for zeros, n_samples, sample_rate, n_mels, kwargs)
CASES = {
    'prod':     (1000, 36000, 24000, 64, PROD),
    'prodB':    (2000, 36000, 24000, 64, PROD),
    'zeros':    (None, 32000, 16000, 64, {}),            # reference feature_extractor_test.py:24-34
    'default':  (7,    8000,  16000, 64, {}),
    'ragged':   (11,   30007, 24000, 64, PROD),
    'nfft256':  (13,   5000,  8000,  32, dict(n_fft=256)),
    'nfft2048': (17,   20000, 48000, 64, dict(n_fft=2048, win_length=1200, hop_length=600)),
    'loud':     (19,   12000, 24000, 64, PROD),          # scaled x300 below: top_db floor > 0 dB
}


def case_input(name):
    seed, n, sr, n_mels, kw = CASES[name]
    if seed is None:
        wav = torch.zeros(4, n)
    else:
        wav = make_clip(seed, n, sr)
        if name == 'loud':
            wav = wav * 300.0
    return wav, sr, n_mels, kw


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    os.environ['CUDA_VISIBLE_DEVICES'] = '-1'
    fe = load_reference_extractor()
    out_dir = os.path.join(REPO, 'tests', 'golden')
    os.makedirs(out_dir, exist_ok=True)
    for name in CASES:
        wav, sr, n_mels, kw = case_input(name)
        rec = {'input_sha256': np.array(sha(wav.numpy()))}
        for mode in ('foa', 'mic'):
            rec[mode] = np.ascontiguousarray(fe.extract_features(wav, sr, mode=mode, n_mels=n_mels, **kw))
        spec = fe.complex_spec(wav, **kw)
        rec['spec_shape'] = np.array(spec.shape)
        if name in ('prod', 'default'):
            rec['spec'] = spec.numpy()
            rec['iv'] = fe.foa_intensity_vectors(spec).numpy()
            rec['gcc'] = fe.gcc_features(spec, n_mels).numpy()
        np.savez_compressed(os.path.join(out_dir, f'extract_{name}.npz'), **rec)
        print(name, {k: v.shape for k, v in rec.items()})

    # statistics / normaliser / pad-truncate through the reference's file-based functions
    rng = np.random.default_rng(5)
    with tempfile.TemporaryDirectory() as d:
        src, dst = os.path.join(d, 'feat'), os.path.join(d, 'norm')
        os.makedirs(src)
        clips = [(rng.standard_normal((40, 8, 7)) * (1 + i) + i).astype(np.float32) for i in range(3)]
        for i, c in enumerate(clips):
            np.save(os.path.join(src, f'fold1_room1_mix00{i}.npy'), c)
        cwd = os.getcwd()
        os.chdir(d)
        try:
            mean, std = fe.calculate_statistics(src)
            fe.apply_normalizer(src, dst, mean, std)
        finally:
            os.chdir(cwd)
        normed = [np.load(os.path.join(dst, f'fold1_room1_mix00{i}.npy')) for i in range(3)]
    feats = (rng.standard_normal((13, 4, 7))).astype(np.float32)
    labels = (rng.standard_normal((7, 8))).astype(np.float32)
    f_pad, l_pad = fe.preprocess_features_labels(feats, labels, max_label_length=4, multiplier=5)
    f_cut, l_cut = fe.preprocess_features_labels(feats, labels, max_label_length=2, multiplier=5)
    np.savez_compressed(os.path.join(out_dir, 'stats_norm.npz'),
                        clips=np.stack(clips), mean=mean, std=std, normed=np.stack(normed),
                        feats=feats, labels=labels, f_pad=f_pad, l_pad=l_pad, f_cut=f_cut, l_cut=l_cut,
                        polar=np.array([[0, 90, 1], [-90, 0, 1], [0, 0, 1], [135, 0, np.sqrt(8)], [0, 0, 0]], float),
                        cart=fe.polar_to_cartesian(np.array([[0, 90, 1], [-90, 0, 1], [0, 0, 1],
                                                             [135, 0, np.sqrt(8)], [0, 0, 0]], float)))
    print('stats_norm', mean.shape, std.shape)


if __name__ == '__main__':
    main()
