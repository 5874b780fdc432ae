"""Restatement of TensorFlow-2 eager random seeding + Philox4x32-10 -- TEST INFRASTRUCTURE ONLY.

TensorFlow is not installed here and is not part of /root/reference, so this follows
TF's published behaviour, anchored on the reference's own known-answer test
(reference transforms_test.py:8-30), which tests/test_masking_cpu.py reproduces:

  tf.random.set_seed(s)  -> graph seed g = s and a fresh ``random.Random(s)`` (CPython
      Mersenne Twister) in the eager context;
  every random op created without an op seed draws  op = rng.randint(0, 2**31 - 1);
  kernel seeds (seed, seed2) = (g % (2**31-1), op % (2**31-1)), and (0, 2**31-1) if both are 0;
  PhiloxRandom(seed, seed2): key = (lo32(seed), hi32(seed)), counter = (0, 0, lo32(seed2), hi32(seed2));
  a scalar int32 uniform in [minval, maxval) is  minval + out[0] % (maxval - minval).

Call sites in the reference: transforms.py:21-22 and :60-61 (size, then offset, per mask).
"""
import random

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK32 = 0xFFFFFFFF
MAXINT32 = 2 ** 31 - 1


def philox4x32_10(counter, key):
    """One Philox4x32-10 block. ``counter`` = 4 uint32, ``key`` = 2 uint32 -> 4 uint32."""
    c0, c1, c2, c3 = (int(v) & MASK32 for v in counter)
    k0, k1 = (int(v) & MASK32 for v in key)
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK32, p1 & MASK32, \
                         ((p0 >> 32) ^ c3 ^ k1) & MASK32, p0 & MASK32
        k0 = (k0 + W0) & MASK32
        k1 = (k1 + W1) & MASK32
    return c0, c1, c2, c3


def tf_philox_first_u32(seed: int, seed2: int) -> int:
    key = (seed & MASK32, (seed >> 32) & MASK32)
    counter = (0, 0, seed2 & MASK32, (seed2 >> 32) & MASK32)
    return philox4x32_10(counter, key)[0]


class TFEagerRandom:
    """Stream of scalar int32 ``tf.random.uniform`` draws after ``tf.random.set_seed(seed)``."""

    def __init__(self, seed: int):
        self.graph_seed = int(seed)
        self._rng = random.Random(int(seed))

    def op_seeds(self):
        op = self._rng.randint(0, MAXINT32)
        s, s2 = self.graph_seed % MAXINT32, op % MAXINT32
        if s == 0 and s2 == 0:
            s, s2 = 0, MAXINT32
        return s, s2

    def uniform_int(self, maxval: int, minval: int = 0) -> int:
        if maxval <= minval:
            raise ValueError('maxval must be > minval')
        s, s2 = self.op_seeds()
        return minval + tf_philox_first_u32(s, s2) % (maxval - minval)


class CounterRandom:
    """The product's own stream (rng_mode PHILOX_COUNTER of seld_mask, seld_b200/csrc/mask.cu): Philox4x32-10
    keyed by the 64-bit seed, counter = (lo32(s), hi32(s), chunk, axis << 24 | mask << 1 | draw) with
    s = global sample index, axis 0 = time / 1 = freq, draw 0 = size / 1 = offset."""

    def __init__(self, seed: int):
        self.key = (seed & MASK32, (seed >> 32) & MASK32)

    def u32(self, sample: int, axis: int, chunk: int, mask_i: int, draw: int) -> int:
        c3 = ((axis & 0xFF) << 24) | ((mask_i << 1) & 0xFFFFFF) | (draw & 1)
        return philox4x32_10((sample & MASK32, (sample >> 32) & MASK32, chunk & MASK32, c3), self.key)[0]

    def drawer(self, sample: int, axis: int, chunk: int):
        """-> draw(maxval) callable yielding size, offset, size, offset, ... for successive masks."""
        state = {'i': 0}

        def draw(maxval: int) -> int:
            i = state['i']
            state['i'] += 1
            return self.u32(sample, axis, chunk, i // 2, i % 2) % maxval
        return draw
