"""Environment-noise check with long kernels: per-launch event times of 60 bf16 GEMMs (~9 ms each) and 60 large copies."""
import torch

a = torch.randn(20480, 20480, device='cuda', dtype=torch.bfloat16)
b = torch.randn(20480, 20480, device='cuda', dtype=torch.bfloat16)
src = torch.empty(6_000_000_000, dtype=torch.float32, device='cuda').normal_()      # 24 GB
dst = torch.empty_like(src)
for label, fn in (('gemm', lambda: a @ b), ('copy 24GB', lambda: dst.copy_(src))):
    for _ in range(3):
        fn()
    n = 60
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    for i in range(n):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    t = [ev[i].elapsed_time(ev[i + 1]) for i in range(n)]
    print(label, 'median %.2f min %.2f max %.2f' % (sorted(t)[n // 2], min(t), max(t)), ' '.join(f'{x:.1f}' for x in t))
