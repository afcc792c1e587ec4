import torch
dev = torch.device("cuda:0")
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(n):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
for mb in (90, 512, 2048):
    n = mb * 1024 * 1024 // 2
    x = torch.empty(n, dtype=torch.float16, device=dev)
    y = torch.empty(n, dtype=torch.float16, device=dev)
    ms = t(lambda: x.fill_(1.0))
    print(f"fill  {mb:5d} MB: {ms*1e3:8.1f} us  {mb*1.048576/ms:8.1f} GB/s (write only)")
    ms = t(lambda: y.copy_(x))
    print(f"copy  {mb:5d} MB: {ms*1e3:8.1f} us  {2*mb*1.048576/ms:8.1f} GB/s (read+write)")
    ms = t(lambda: x.sum())
    print(f"sum   {mb:5d} MB: {ms*1e3:8.1f} us  {mb*1.048576/ms:8.1f} GB/s (read only)")
