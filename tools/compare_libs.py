"""Bitwise comparison of the MIC / FOA features two builds of the library produce on ordinary and degenerate inputs:
    python tools/compare_libs.py seld_b200/build/variants/lib_prev.so            (against the in-tree library)"""
import os
import subprocess
import sys

here = os.path.dirname(os.path.abspath(__file__))
root = os.path.dirname(here)
code = r'''
import sys, torch, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r + '/tests')
from cases import PROD
from seld_b200 import pipeline
from seld_b200.synth import make_clips
base = make_clips([31, 32], 48000)
cases = {'plain': base, 'quiet1e-18': base * 1e-18, 'quiet1e-21': base * 1e-21, 'dc': torch.full((1, 4, 24000), 0.5)}
w = base.clone(); w[:, :, :20000] = 0; cases['half_silent'] = w
w = base.clone(); w[0, 2] = 0; cases['dead_ch2'] = w
w = torch.zeros(1, 4, 24000); w[0, :, 10000] = 1.0; cases['impulse'] = w
w = base.clone() * 1e-19; w[0, 1] *= 1e-3; cases['mixed_tiny'] = w
out = {}
for name, wav in cases.items():
    for mode in ('mic', 'foa'):
        feat, key = pipeline.extract_batch(wav.cuda(), 24000, mode=mode, **PROD)
        out[name + '_' + mode] = feat.cpu().numpy()
np.savez(sys.argv[1], **out)
''' % (root, root)
outs = []
for i, lib in enumerate([sys.argv[1], None]):
    env = dict(os.environ)
    if lib:
        env['SELD_B200_LIB'] = os.path.abspath(lib)
    path = f'/tmp/cmp_{i}.npz'
    r = subprocess.run([sys.executable, '-c', code, path], env=env, capture_output=True, text=True)
    if r.returncode:
        raise SystemExit(r.stderr[-2000:])
    outs.append(path)
import numpy as np
a, b = np.load(outs[0]), np.load(outs[1])
for k in a.files:
    x, y = a[k], b[k]
    same = np.array_equal(x, y, equal_nan=True)
    d = np.nanmax(np.abs(x - y)) if not same else 0.0
    print(f'{k:24s} identical {same}  max |diff| {d:.3e}  nan {int(np.isnan(x).sum())} / {int(np.isnan(y).sum())}')
