"""Oracle + host logic of the batch-level spatial augmentations, pinned on the reference's own tests:
transforms_test.py:32-52 (x and y flipped equally), :54-62 (label split shapes), :64-73 (exact mic_gcc_perm table)."""
import numpy as np
import torch

from oracle import augment as A
from oracle.tf_random import philox4x32_10
from seld_b200 import philox, transforms as T

MIC_PERM = np.array([[1, 3, 0, 2], [3, 1, 2, 0], [1, 0, 3, 2], [2, 0, 3, 1], [0, 2, 1, 3], [3, 2, 1, 0]])
GCC_PERM = np.array([[4, 0, 3, 2, 5, 1], [4, 5, 2, 3, 0, 1], [0, 4, 3, 2, 1, 5], [1, 5, 3, 2, 0, 4], [1, 0, 2, 3, 5, 4],
                     [5, 4, 2, 3, 1, 0]])      # transforms_test.py:66-72


def test_mic_gcc_perm_known_answers():
    assert np.array_equal(A.mic_gcc_perm_ref(MIC_PERM), GCC_PERM)
    assert np.array_equal(T.mic_gcc_perm(MIC_PERM), GCC_PERM)
    assert torch.equal(T.mic_gcc_perm(torch.from_numpy(MIC_PERM)).long(), torch.from_numpy(GCC_PERM))


def test_channel_list_matches_oracle_table():
    assert np.array_equal(np.array(T.channel_list), A.CHANNEL_LIST)
    for mic, foa in T.channel_list:           # every entry is a permutation (signed for FOA), W untouched
        assert sorted(mic) == [0, 1, 2, 3] and foa[0] == 0 and sorted(abs(v) for v in foa[1:]) == [1, 2, 3]


def test_split_total_labels_shapes():
    batch, time, n_classes = 32, 10, 14
    _, (sed, doa) = T.split_total_labels_to_sed_doa(None, torch.zeros(batch, time, n_classes * 4))
    assert tuple(sed.shape) == (batch, time, n_classes) and tuple(doa.shape) == (batch, time, n_classes * 3)


def _flip_fractions(x, nx, y, ny):
    xf = (x[..., -3:] != nx[..., -3:]).astype(np.float32).mean(axis=(1, 2))
    s = y.shape[:-1] + (4, -1)
    yf = (y.reshape(s)[..., -3:, :] != ny.reshape(s)[..., -3:, :]).astype(np.float32).mean(axis=(1, 3))
    return xf, yf


def test_oracle_iv_aug_flips_x_and_y_equally():
    rng = np.random.default_rng(2022)
    x = rng.random((8, 10, 32, 7), dtype=np.float32)
    y = rng.random((8, 2, 12), dtype=np.float32)
    for trial in range(4):
        flip = rng.integers(0, 2, (8, 3))
        swap = rng.integers(0, 2, 8)
        nx, ny = A.foa_intensity_vec_aug_ref(x, y, flip, swap)
        assert nx.shape == x.shape and ny.shape == y.shape
        xf, yf = _flip_fractions(x, nx, y, ny)
        assert np.array_equal(xf, yf)
        # log-mel W channel never changes; |values| are only permuted
        assert np.array_equal(nx[..., 0], x[..., 0])
        assert np.array_equal(np.sort(np.abs(nx[..., 4:7]), -1), np.sort(np.abs(x[..., 4:7]), -1))
    nx, ny = A.foa_intensity_vec_aug_ref(x, y, np.zeros((8, 3), int), np.zeros(8, int))
    assert np.array_equal(nx, x) and np.array_equal(ny, y)


def test_oracle_acs_identity_and_group_structure():
    rng = np.random.default_rng(7)
    x = rng.standard_normal((8, 6, 16, 17)).astype(np.float32)
    y = rng.standard_normal((8, 3, 12)).astype(np.float32)
    nx, ny = A.acs_aug_ref(x, y, np.full(8, 2))        # entry 2 is the identity
    assert np.array_equal(nx, x) and np.array_equal(ny, y)
    nx, ny = A.acs_aug_ref(x, y, np.arange(8))
    assert np.array_equal(nx[..., 0], x[..., 0])
    for b in range(8):                                   # every block is a (signed) permutation of itself
        for lo, hi in ((1, 4), (4, 7), (7, 11), (11, 17)):
            assert np.array_equal(np.sort(np.abs(nx[b, ..., lo:hi]), -1), np.sort(np.abs(x[b, ..., lo:hi]), -1))
    # involutions of the table (entries 3, 5, 6, 7 are their own inverse)
    for i in (3, 5, 6, 7):
        a, b2 = A.acs_aug_ref(*A.acs_aug_ref(x, y, np.full(8, i)), np.full(8, i))
        assert np.array_equal(a, x) and np.array_equal(b2, y)


def test_host_philox_matches_oracle_generator():
    seed = 0x1234_5678_9ABC_DEF0
    w = philox.sample_words(seed, (1 << 32) - 2, 5, T.STREAM_IV_AUG, draw=3)
    for i in range(5):
        s = (1 << 32) - 2 + i
        want = philox4x32_10((s & 0xFFFFFFFF, s >> 32, T.STREAM_IV_AUG, 3), (seed & 0xFFFFFFFF, seed >> 32))
        assert tuple(int(v) for v in w[i]) == want


def test_level_jitter_restatement_is_normal():
    """The jitter itself is drawn on the device (seld_augment_batch); its numpy restatement -- Box-Muller in float64 on Philox
    words 0, 1 of (seed, sample, STREAM_LEVEL_JITTER) -- is what tests/test_gpu_augment.py compares it with."""
    a = A.level_offsets_ref(11, 5, 100000, 0.2)
    assert a.dtype == np.float32 and np.array_equal(a, A.level_offsets_ref(11, 5, 100000, 0.2))
    assert abs(float(a.mean())) < 3e-3 and abs(float(a.std()) - 0.2) < 2e-3
    assert abs(float(np.mean(np.abs(a) < 0.2)) - 0.6827) < 5e-3              # one sigma
    assert np.array_equal(A.level_offsets_ref(11, 15, 10, 0.2), a[10:20])    # keyed by global sample index
