"""Diagnostic matrix for the planar-layout instability: per-variant median / max of 16 launches (run in subprocesses so
each variant gets its own env)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
child = '''
import os, sys, statistics
sys.path.insert(0, %r)
import torch
from seld_b200 import pipeline
from seld_b200.synth import make_clip
kw = dict(win_length=960, hop_length=480, n_fft=1024)
base = [make_clip(1000 + i, device='cuda') for i in range(8)]
wav = torch.stack([base[i %% 8] for i in range(600)])
out = torch.empty(600, 3000, 64, 7, device='cuda')
def run(layout, env):
    for k in ('SELD_FPW', 'SELD_ASSIGN', 'SELD_PLANAR_BURST', 'SELD_ODD', 'SELD_SKIP'):
        os.environ.pop(k, None)
    os.environ.update(env)
    w = wav if layout == 'planar' else wav.transpose(1, 2).contiguous()
    n = 20
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    for _ in range(4):
        pipeline.extract_batch(w, 24000, mode='foa', t_out=3000, layout=layout, out=out, **kw)
    ev[0].record()
    for i in range(n):
        pipeline.extract_batch(w, 24000, mode='foa', t_out=3000, layout=layout, out=out, **kw)
        ev[i + 1].record()
    torch.cuda.synchronize()
    t = [ev[i].elapsed_time(ev[i + 1]) for i in range(n)]
    print(f'{layout:12s} {str(env):60s} median {statistics.median(t):7.2f} min {min(t):7.2f} max {max(t):7.2f}', flush=True)
''' % os.path.dirname(HERE)
variants = [('planar', {}), ('interleaved', {}), ('planar', {'SELD_FPW': '2'}), ('planar', {'SELD_FPW': '4'}), ('planar', {'SELD_FPW': '16'}),
            ('interleaved', {'SELD_FPW': '2'}), ('planar', {}), ('interleaved', {})]
code = child + '\n' + '\n'.join(f'run({l!r}, {e!r})' for l, e in variants)
subprocess.run([sys.executable, '-c', code], check=False)
