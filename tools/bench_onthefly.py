"""Config 5(ii) of BASELINE.json: on-the-fly training batches, 256 chunks of 300 frames (6 s) with context ->
fused extract -> clamp + normalise -> time/frequency masking -> [256, 300, 64, 7].  One JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from seld_b200 import pipeline  # noqa: E402

kw = dict(win_length=960, hop_length=480, n_fft=1024)
B, T = 256, 300
Lc = (T - 1) * 480 + 1024
n_batches = 12
chunks = (torch.rand(n_batches, B, 4, Lc, device='cuda') - 0.5) * 0.2          # 12 x 594 MB
cmax = torch.full((B,), 20.0, device='cuda')
mean = torch.zeros(1, 64, 7, device='cuda')
std = torch.ones(1, 64, 7, device='cuda')
for i in range(3):
    pipeline.training_batch(chunks[i], 24000, cmax, mean, std, seed=1, sample_offset=i * B, **kw)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(n_batches):
    out = pipeline.training_batch(chunks[i], 24000, cmax, mean, std, seed=2, sample_offset=i * B, **kw)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n_batches
alg = 4 * B * 4 * Lc + 4 * out.numel()
print(json.dumps({'workload': 'on-the-fly batch: 256 chunks x 6 s FOA -> extract + normalise + mask -> [256,300,64,7]',
                  'ms_per_batch': ms, 'batches_per_s': 1000.0 / ms, 'audio_hours_per_s': B * 6.0 / 3600.0 / (ms / 1000.0),
                  'algorithmic_GBps': alg / ms / 1e6}))
