"""Time the extract launch of every library variant built by tools/build_variants.py (600 planar clips, CUDA events)."""
import glob
import os
import subprocess
import sys

here = os.path.dirname(os.path.abspath(__file__))
mode = sys.argv[1] if len(sys.argv) > 1 else 'mic'
code = r'''
import sys, torch
sys.path.insert(0, %r)
from seld_b200 import pipeline
from seld_b200.synth import make_clip
kw = dict(win_length=960, hop_length=480, n_fft=1024)
mode = %r
base = [make_clip(1000 + i, device='cuda') for i in range(8)]
wav = torch.stack([base[i %% 8] for i in range(600)])
out = torch.empty(600, 3000, 64, 7 if mode == 'foa' else 10, device='cuda')
for _ in range(3):
    pipeline.extract_batch(wav, 24000, mode=mode, t_out=3000, out=out, **kw)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    pipeline.extract_batch(wav, 24000, mode=mode, t_out=3000, out=out, **kw)
e1.record()
torch.cuda.synchronize()
print('%%.3f ms  checksum %%.6f' %% (e0.elapsed_time(e1) / 5, float(out[::37, ::101].double().sum())))
''' % (os.path.dirname(here), mode)
libs = sorted(glob.glob(os.path.join(os.path.dirname(here), 'seld_b200', 'build', 'variants', 'lib_*.so')))
for lib in libs:
    env = dict(os.environ, SELD_B200_LIB=lib)
    r = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True, timeout=150)
    print(f'{mode} {os.path.basename(lib):40s} {r.stdout.strip() or r.stderr.strip()[-300:]}', flush=True)
