"""Diagnostic: 40 back-to-back extract launches on one resident shard, per-launch CUDA-event times + clocks."""
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from seld_b200 import pipeline  # noqa: E402
from seld_b200.synth import make_clip  # noqa: E402

kw = dict(win_length=960, hop_length=480, n_fft=1024)
layout = sys.argv[1] if len(sys.argv) > 1 else 'planar'
clips = int(sys.argv[2]) if len(sys.argv) > 2 else 600
base = [make_clip(1000 + i, device='cuda') for i in range(8)]
wav = torch.stack([base[i % 8] for i in range(clips)])
if layout == 'interleaved':
    wav = wav.transpose(1, 2).contiguous()
out = torch.empty(clips, 3000, 64, 7, device='cuda')
use_smi = os.environ.get('DIAG_SMI', '0') == '1'
smi = None if not use_smi else subprocess.Popen(['nvidia-smi', '--query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.active',
                        '--format=csv,noheader', '-lms', '50'], stdout=subprocess.PIPE, text=True)
time.sleep(0.3)
n = 40
ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
ev[0].record()
for i in range(n):
    pipeline.extract_batch(wav, 24000, mode='foa', t_out=3000, layout=layout, out=out, **kw)
    ev[i + 1].record()
torch.cuda.synchronize()
time.sleep(0.2)
print(layout, clips, 'ms per launch:', ' '.join(f'{ev[i].elapsed_time(ev[i + 1]):.1f}' for i in range(n)))
if smi is None:
    sys.exit(0)
smi.terminate()
lines = smi.stdout.read().strip().splitlines()
print('smi samples:', len(lines))
for ln in lines[::max(1, len(lines) // 12)]:
    print('  ', ln)
