"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): numpy restatement of the reference's batch-level spatial
augmentations with the random draws passed in explicitly.

  foa_intensity_vec_aug_ref   reference transforms.py:78-114
  mic_gcc_perm_ref            reference transforms.py:122-139
  acs_aug_ref                 reference transforms.py:155-199  (CHANNEL_LIST = :143-152)

Pinned by the reference's own tests: the exact `mic_gcc_perm` table (transforms_test.py:64-73) and the "x and y are
flipped equally" property (transforms_test.py:46-52).  The reference draws with tf.random.uniform on vector shapes; no
test pins those values, so the product defines its own counter-based stream and this oracle takes the draws as input."""
import numpy as np

CHANNEL_LIST = np.array([
    [[1, 3, 0, 2], [0, -3, -2, 1]],
    [[3, 1, 2, 0], [0, -3, 2, -1]],
    [[0, 1, 2, 3], [0, 1, 2, 3]],
    [[1, 0, 3, 2], [0, -1, -2, 3]],
    [[2, 0, 3, 1], [0, 3, -2, -1]],
    [[0, 2, 1, 3], [0, 3, 2, 1]],
    [[3, 2, 1, 0], [0, -1, 2, -3]],
    [[2, 3, 0, 1], [0, 1, -2, -3]]], dtype=np.int64)


def _gather_last(a, idx):
    """tf.gather(a, idx, axis=-1, batch_dims=1): a [B, ..., C], idx [B, K] -> [B, ..., K]."""
    return np.stack([a[b][..., idx[b]] for b in range(a.shape[0])])


def _gather_m2(a, idx):
    """tf.gather(a, idx, axis=-2, batch_dims=1)."""
    return np.stack([a[b][..., idx[b], :] for b in range(a.shape[0])])


def foa_intensity_vec_aug_ref(x, y, flip, swap):
    """x [B, T, F, 7], y [B, T, 4*n_classes]; flip int [B, 3] in {0, 1}; swap int [B] in {0, 1} (the reference's
    `tf.random.uniform([B, 1], maxval=2)` before the *2)."""
    x = np.array(x, dtype=np.float32)
    y = np.array(y, dtype=np.float32)
    b = x.shape[0]
    y4 = y.reshape(y.shape[:-1] + (4, y.shape[-1] // 4))
    iv = x[..., -3:]
    cart = y4[..., -3:, :]
    f = np.asarray(flip, dtype=np.float32)
    iv = (1 - 2 * f.reshape(-1, 1, 1, 3)) * iv
    cart = (1 - 2 * f.reshape(-1, 1, 3, 1)) * cart
    correct = np.tile([[0, 1, 2]], (b, 1))
    perm = 2 * np.asarray(swap, dtype=np.int64).reshape(b, 1)
    perm = np.concatenate([perm, np.ones_like(perm), 2 - perm], -1)
    check = (perm != correct).astype(np.int64).sum(-1, keepdims=True)
    feat_perm = (perm + check) % 3
    iv = _gather_last(iv, feat_perm)
    cart = _gather_m2(cart, feat_perm)
    x = np.concatenate([x[..., :1], _gather_last(x[..., 1:4], perm), iv], -1)
    y4 = np.concatenate([y4[..., :-3, :], cart], -2)
    return x.astype(np.float32), y4.reshape(y.shape).astype(np.float32)


def mic_gcc_perm_ref(mic_perm):
    """mic_perm int [B, 4] -> [B, 6]."""
    mic_perm = np.asarray(mic_perm, dtype=np.int64)
    decode = np.array([[0, 0, 1, 2], [0, 0, 3, 4], [1, 3, 0, 5], [2, 4, 5, 0]])
    out = []
    for m in mic_perm:
        # pairs (i, j), i < j, of the permuted channel list, each decoded to the index of the original pair
        pairs = [(m[i], m[j]) for i in range(4) for j in range(i + 1, 4)]
        out.append([decode[a, c] for a, c in pairs])
    return np.array(out, dtype=np.int64)


def acs_aug_ref(x, y, idx):
    """x [B, T, F, 17], y [B, T, 4*n_classes]; idx int [B] in [0, 8)."""
    x = np.array(x, dtype=np.float32)
    y = np.array(y, dtype=np.float32)
    y4 = y.reshape(y.shape[:-1] + (4, y.shape[-1] // 4))
    iv = x[..., 4:7]
    cart = y4[..., -3:, :]
    flip = CHANNEL_LIST[np.asarray(idx, dtype=np.int64)]
    foa_flip = flip[..., 1, 1:]
    foa_sign = np.sign(foa_flip)
    foa_perm = foa_sign * foa_flip - 1
    check = (foa_perm != np.array([0, 1, 2])).astype(np.int64).sum(-1, keepdims=True)
    feat_perm = (foa_perm + check) % 3
    foa_x = _gather_last(x[..., 1:4], foa_perm)
    s = foa_sign.astype(np.float32)
    iv = _gather_last(iv, feat_perm) * s[:, None, None, :]
    cart = _gather_m2(cart, feat_perm) * s[:, None, :, None]
    mic_flip = flip[..., 0, :]
    gcc = _gather_last(x[..., 11:], mic_gcc_perm_ref(mic_flip))
    mic_x = _gather_last(x[..., 7:11], mic_flip)
    x = np.concatenate([x[..., :1], foa_x, iv, mic_x, gcc], -1)
    y4 = np.concatenate([y4[..., :-3, :], cart], -2)
    return x.astype(np.float32), y4.reshape(y.shape).astype(np.float32)


def level_offsets_ref(seed, first_sample, n, stddev, stream_id=0x102):
    """random_ups_and_downs (reference trainv2.py:120-124) as the product draws it: one N(0, stddev^2) float32 per global sample
    index -- Box-Muller in float64 on Philox4x32-10 words 0, 1 of counter (lo32(sample), hi32(sample), stream_id, 0), key = seed."""
    from oracle.tf_random import philox4x32_10
    out = np.empty(n, dtype=np.float32)
    stddev = float(np.float32(stddev))               # the C ABI carries the standard deviation as a float
    for i in range(n):
        s = first_sample + i
        w = philox4x32_10((s & 0xFFFFFFFF, s >> 32, stream_id, 0), (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))
        u1, u2 = (w[0] + 0.5) / 4294967296.0, (w[1] + 0.5) / 4294967296.0
        out[i] = stddev * np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)
    return out
