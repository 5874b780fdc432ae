"""Small driver for ncu: a few extract launches on a reduced shard (same kernel, same per-CTA work pattern).

    python tools/profile_extract.py --mode foa --clips 74 --iters 3 [--layout interleaved]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from seld_b200 import pipeline  # noqa: E402
from seld_b200.synth import make_clip  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--mode', default='foa')
ap.add_argument('--layout', default='planar')
ap.add_argument('--clips', type=int, default=74)
ap.add_argument('--iters', type=int, default=3)
ap.add_argument('--post', action='store_true', help='also run stats + finalize')
a = ap.parse_args()
kw = dict(win_length=960, hop_length=480, n_fft=1024)
base = [make_clip(1000 + i, device='cuda') for i in range(4)]
wav = torch.stack([base[i % 4] for i in range(a.clips)])
if a.layout == 'interleaved':
    wav = wav.transpose(1, 2).contiguous()
for _ in range(a.iters):
    feat, key = pipeline.extract_batch(wav, 24000, mode=a.mode, t_out=3000, layout=a.layout, **kw)
    if a.post:
        acc = pipeline.partial_statistics(feat, key, 3001)
        mean, std = pipeline.finish_statistics(acc, 64, feat.shape[3])
        pipeline.finalize_(feat, key, 3001, mean, std)
torch.cuda.synchronize()
print('ok', feat.shape, float(feat[0, 100, 10, 0]))
