"""Degenerate / extreme inputs through the CUDA extractor next to the oracle port (prints max errors; test infrastructure)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from cases import PROD
from oracle import extractor as O
from seld_b200 import pipeline
from seld_b200.synth import make_clips
def run(name, wav, modes=('foa','mic')):
    for mode in modes:
        feat, key = pipeline.extract_batch(wav.cuda(), 24000, mode=mode, **PROD)
        pipeline.finalize_(feat, key, feat.shape[1])
        for i in range(wav.shape[0]):
            ref = O.extract_features_port(wav[i], 24000, mode=mode, **PROD)
            got = feat[i].cpu().numpy()
            print(name, mode, i, 'logmel err %.2e' % np.abs(got[..., :4]-ref[..., :4]).max(), 'rest err %.2e' % np.abs(got[..., 4:]-ref[..., 4:]).max(), 'finite', np.isfinite(got).all(), flush=True)
base = make_clips([31], 24000)
w = base.clone(); w[0,1] = w[0,0]; run('identical ch0==ch1', w)
w = base.clone(); w[0,3] = w[0,2]; run('identical ch2==ch3', w)
w = base.clone(); w[0,2] = w[0,0]; run('identical ch0==ch2', w)
run('loud x1e4', base*1e4); run('quiet x1e-6', base*1e-6); run('quiet x1e-3', base*1e-3)
run('dc 0.5', torch.full((1,4,24000), 0.5))
w = base.clone(); w[0,:, :12000] = 0; run('half silent', w)
w = torch.zeros(1,4,24000); w[0,:,10000] = 1.0; run('impulse', w)
# CUDA-core GCC path on the quiet noise-free case (per-channel unit phasors rescale, no |R|^2 window)
for scale in (1e-3, 1e-6):
    w = base * scale
    feat, key = pipeline.extract_batch(w.cuda(), 24000, mode='mic', use_tensor_cores=False, **PROD)
    pipeline.finalize_(feat, key, feat.shape[1])
    ref = O.extract_features_port(w[0], 24000, mode='mic', **PROD)
    print('quiet x%g mic CUDA-core path: gcc err %.2e' % (scale, np.abs(feat[0].cpu().numpy()[..., 4:] - ref[..., 4:]).max()), flush=True)
