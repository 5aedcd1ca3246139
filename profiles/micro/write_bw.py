"""Micro-measurement: pure-write, pure-read and copy bandwidth of this B200 (context for the RoI kernel's roofline:
the path's dominant kernel is ~85% writes).  Run under gpurun; prints GB/s."""
import torch

def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3

for mb in (482, 1024, 4096):
    n = mb * 1024 * 1024 // 4
    a = torch.empty(n, device='cuda'); b = torch.empty(n, device='cuda')
    t = timeit(lambda: a.fill_(1.0)); print('fill   %5d MB: %7.1f GB/s' % (mb, n * 4 / t / 1e9))
    t = timeit(lambda: a.zero_()); print('zero   %5d MB: %7.1f GB/s' % (mb, n * 4 / t / 1e9))
    t = timeit(lambda: b.copy_(a)); print('copy   %5d MB: %7.1f GB/s (read+write bytes)' % (mb, 2 * n * 4 / t / 1e9))
    t = timeit(lambda: a.sum()); print('read   %5d MB: %7.1f GB/s' % (mb, n * 4 / t / 1e9))
    t = timeit(lambda: torch.add(a, 1.0, out=b)); print('add    %5d MB: %7.1f GB/s (read+write bytes)' % (mb, 2 * n * 4 / t / 1e9))
