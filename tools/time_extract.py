"""Event-timed extract launches for a sweep of shard sizes / layouts (diagnostic)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from seld_b200 import pipeline  # noqa: E402
from seld_b200.synth import make_clip  # noqa: E402

kw = dict(win_length=960, hop_length=480, n_fft=1024)
mode = sys.argv[1] if len(sys.argv) > 1 else 'foa'
base = [make_clip(1000 + i, device='cuda') for i in range(8)]
for clips in (74, 148, 296, 600):
    wav = torch.stack([base[i % 8] for i in range(clips)])
    for layout in ('planar', 'interleaved', 'pcm16'):
        if layout == 'planar':
            w = wav
        elif layout == 'interleaved':
            w = wav.transpose(1, 2).contiguous()
        else:                                   # 16-bit PCM in WAV frame order, decoded by the kernel
            w = (wav.transpose(1, 2) * 32767.0).round().clamp(-32768, 32767).to(torch.int16).contiguous()
        out = torch.empty(clips, 3000, 64, 7 if mode == 'foa' else 10, device='cuda')
        for _ in range(2):
            pipeline.extract_batch(w, 24000, mode=mode, t_out=3000, layout='interleaved' if layout == 'pcm16' else layout, out=out, **kw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            pipeline.extract_batch(w, 24000, mode=mode, t_out=3000, layout='interleaved' if layout == 'pcm16' else layout, out=out, **kw)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        print(f'{mode} clips {clips:4d} {layout:12s} {ms:8.3f} ms  {1000 * ms / clips:7.2f} us/clip', flush=True)
        del w, out
    del wav
