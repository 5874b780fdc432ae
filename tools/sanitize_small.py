"""Small extract / post / mask / remap invocations for compute-sanitizer (memcheck, racecheck):

    compute-sanitizer --tool memcheck python tools/sanitize_small.py      (where the pool allows the sanitizer)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from seld_b200 import pipeline, transforms  # noqa: E402

kw = dict(win_length=960, hop_length=480, n_fft=1024)
g = torch.Generator().manual_seed(0)
wav = (torch.rand(3, 4, 24000, generator=g) - 0.5).cuda()
for mode in ('foa', 'mic'):
    for layout in ('planar', 'interleaved'):
        w = wav if layout == 'planar' else wav.transpose(1, 2).contiguous()
        feat, key = pipeline.extract_batch(w, 24000, mode=mode, t_out=50, layout=layout, **kw)
        if mode == 'mic':
            pipeline.extract_batch(w, 24000, mode=mode, t_out=50, layout=layout, use_tensor_cores=False, **kw)
        acc = pipeline.partial_statistics(feat, key, 51)
        mean, std = pipeline.finish_statistics(acc, 64, feat.shape[3])
        pipeline.finalize_(feat, key, 51, mean, std)
pcm = (wav.transpose(1, 2) * 32767).round().to(torch.int16).contiguous()
pipeline.extract_batch(pcm, 24000, mode='foa', t_out=50, layout='interleaved', **kw)
pipeline.extract_batch(wav[:, :, :4000], 24000, mode='foa', n_fft=512)            # reference default geometry (R = 16)
pipeline.extract_batch(wav[:, :, :4000], 24000, mode='mic', n_fft=256)
x = torch.rand(4, 300, 64, 7, device='cuda')
transforms.mask_batch_(x, (24, 1), (16, 1), seed=1)
transforms.foa_intensity_vec_aug(x, torch.rand(4, 60, 56, device='cuda'), seed=2)
torch.cuda.synchronize()
print('ok')
