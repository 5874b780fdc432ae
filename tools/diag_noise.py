"""Diagnostic: does this box show sporadic slow launches for a plain device copy too? (environment noise check)"""
import torch

a = torch.empty(3_456_000_000, dtype=torch.float32, device='cuda').normal_()     # 13.8 GB
b = torch.empty_like(a[:806_400_000])
n = 40
for label, fn in (('copy 3.2GB', lambda: b.copy_(a[:806_400_000])), ('sum 13.8GB', lambda: a.sum())):
    fn()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    for i in range(n):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    print(label, ' '.join(f'{ev[i].elapsed_time(ev[i + 1]):.2f}' for i in range(n)))
