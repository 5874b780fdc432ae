"""Config 5(i) of BASELINE.json: fused time + frequency masking of a training batch (256 x [300, 64, 7] float32, 6 s chunks)
with the parameters of reference train.py:157-160 (24x1, 16x1) and trainv2.py:136-137 (6x10, 8x6).  Prints one JSON line
per parameter set: batches/s, ms, GB/s against the reference's full read + write (275 MB/batch) and the bytes the kernel
actually touches (masked elements only)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from seld_b200 import transforms as T  # noqa: E402

B, FR, M, C = 256, 300, 64, 7
n_batches = 24                                     # 24 x 138 MB = 3.3 GB of distinct batches >> 126 MB L2
x = torch.randn(n_batches, B, FR, M, C, device='cuda')
alg_bytes = 2 * 4 * B * FR * M * C
for name, tm, fm in (('train.py (time 24x1, freq 16x1)', (24, 1), (16, 1)), ('trainv2.py (time 6x10, freq 8x6)', (6, 10), (8, 6))):
    for i in range(3):
        T.mask_batch_(x[i], tm, fm, seed=1, sample_offset=i * B)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n_batches):
        T.mask_batch_(x[i], tm, fm, seed=2, sample_offset=i * B)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n_batches
    frac = float((x[0] == 0).float().mean())
    print(json.dumps({'workload': f'mask_batch_ 256x[300,64,7] {name}', 'ms_per_batch': ms, 'batches_per_s': 1000.0 / ms,
                      'samples_per_s': B * 1000.0 / ms, 'algorithmic_GBps_vs_full_read_write': alg_bytes / ms / 1e6,
                      'touched_fraction_measured': frac, 'touched_GBps': 2 * frac * alg_bytes / 2 / ms / 1e6}))
    x.normal_()
