"""Time the statistics pass and the clamp + normalise pass of every library variant in seld_b200/build/variants
(600 x [3000, 64, C] resident, CUDA events, 20 repetitions)."""
import glob
import os
import subprocess
import sys

here = os.path.dirname(os.path.abspath(__file__))
code = r'''
import sys, torch
sys.path.insert(0, %r)
from seld_b200 import pipeline
out = []
for C in (7, 10):
    feat = torch.randn(600, 3000, 64, C, device='cuda') * 10 - 40
    key = torch.zeros(600, dtype=torch.int32, device='cuda')          # (the clamp floor value does not change the traffic)
    ws = None
    acc = pipeline.new_accumulator(64, C, feat.device)
    def t(fn, n=20):
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    ts = t(lambda: pipeline.partial_statistics(feat, key, 3001, acc=acc))
    mean, std = pipeline.finish_statistics(acc, 64, C)
    tf = t(lambda: pipeline.finalize_(feat, key, 3001, mean, std))
    nbytes = feat.numel() * 4
    out.append('C=%%d stats %%.3f ms (%%.0f GB/s)  finalize %%.3f ms (%%.0f GB/s)' %% (C, ts, nbytes / ts / 1e6, tf, 2 * nbytes / tf / 1e6))
    del feat
print(' | '.join(out))
''' % (os.path.dirname(here),)
libs = sorted(glob.glob(os.path.join(os.path.dirname(here), 'seld_b200', 'build', 'variants', 'lib_*.so')))
for lib in libs:
    env = dict(os.environ, SELD_B200_LIB=lib)
    r = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True, timeout=200)
    print(f'{os.path.basename(lib):32s} {r.stdout.strip() or r.stderr.strip()[-400:]}', flush=True)
