"""Diagnostic: is the planar-layout extract time allocation (physical placement) dependent?"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from seld_b200 import pipeline  # noqa: E402

kw = dict(win_length=960, hop_length=480, n_fft=1024)
clips = 600


def timeit(w, layout, out, reps=3):
    pipeline.extract_batch(w, 24000, mode='foa', t_out=3000, layout=layout, out=out, **kw)
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pipeline.extract_batch(w, 24000, mode='foa', t_out=3000, layout=layout, out=out, **kw)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return ' '.join(f'{t:6.2f}' for t in ts)


out = torch.empty(clips, 3000, 64, 7, device='cuda')
keep = []
for L in (1_440_000, 1_441_792, 1_444_000):
    for trial in range(4):
        w = (torch.rand(clips, 4, L, device='cuda') - 0.5) * 0.2
        print(f'L {L} alloc {trial} ptr {w.data_ptr():#x} planar      ', timeit(w, 'planar', out), flush=True)
        if trial == 0:
            wi = w.transpose(1, 2).contiguous()
            print(f'L {L} alloc {trial} ptr {wi.data_ptr():#x} interleaved ', timeit(wi, 'interleaved', out), flush=True)
            del wi
        keep.append(w) if trial % 2 == 0 and len(keep) < 2 else None     # perturb the allocator state
        del w
    keep.clear()
    torch.cuda.empty_cache()
