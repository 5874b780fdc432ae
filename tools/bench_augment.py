"""Time the batch-level spatial augmentations on the reference's batch shape (train.py:163-165: [256, 300, 64, 7])."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from seld_b200 import transforms as T


def timed(fn, iters=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


if __name__ == '__main__':
    for c, fn in ((7, T.foa_intensity_vec_aug), (17, T.acs_aug)):
        x = torch.rand(256, 300, 64, c, device='cuda')
        y = torch.rand(256, 60, 56, device='cuda')
        us = timed(lambda: fn(x, y, seed=1))
        us_copy = timed(lambda: (x.clone(), y.clone()))
        nbytes = 2 * x.numel() * 4
        print(json.dumps({'op': fn.__name__, 'shape': list(x.shape), 'us_per_batch': round(us, 1),
                          'includes': 'one fused launch over x (device draws) + one over the labels', 'us_plain_copy': round(us_copy, 1),
                          'effective_GBps': round(nbytes / (us * 1e-6) / 1e9, 1)}))
