import torch, time
def ev(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s,e=torch.cuda.Event(True),torch.cuda.Event(True); s.record()
    for _ in range(it): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e)/it
n = 5*1024**3
a = torch.empty(n, dtype=torch.uint8, device='cuda'); b = torch.empty(n, dtype=torch.uint8, device='cuda')
ms = ev(lambda: a.fill_(7)); print('fill_ %.1f GB in %.3f ms -> %.0f GB/s write-only'%(n/1e9, ms, n/ms/1e6))
af = a.view(torch.float32)
ms = ev(lambda: af.fill_(1.5)); print('fill_ f32 -> %.0f GB/s'%(n/ms/1e6))
ms = ev(lambda: b.copy_(a)); print('copy -> %.0f GB/s (read+write)'%(2*n/ms/1e6))
ms = ev(lambda: a.sum()); print('sum(read-only) u8 -> %.0f GB/s'%(n/ms/1e6))
ms = ev(lambda: af.sum()); print('sum(read-only) f32 -> %.0f GB/s'%(n/ms/1e6))
