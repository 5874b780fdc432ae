import sys, torch
sys.path.insert(0, '.')
from seld_b200 import pipeline
from seld_b200.synth import make_clip
kw = dict(win_length=960, hop_length=480, n_fft=1024)
mode = sys.argv[1] if len(sys.argv) > 1 else 'foa'
base = [make_clip(1000 + i, device='cuda') for i in range(8)]
wav = torch.stack([base[i % 8] for i in range(600)])
out = torch.empty(600, 3000, 64, 7 if mode == 'foa' else 10, device='cuda')
for _ in range(3):
    pipeline.extract_batch(wav, 24000, mode=mode, t_out=3000, out=out, **kw)
ts = []
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        pipeline.extract_batch(wav, 24000, mode=mode, t_out=3000, out=out, **kw)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) / 10)
print(mode, ' '.join('%.3f' % t for t in ts), 'ms  checksum %.6f' % float(out[::37, ::101].double().sum()))
