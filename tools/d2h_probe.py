"""raw pinned-memory D2H / H2D bandwidth of the box (context for bench.py's e2e number)"""
import torch, time
n = 64 * 1024 * 1024
d = torch.empty(n, dtype=torch.uint8, device="cuda"); h = torch.empty(n, dtype=torch.uint8).pin_memory()
for name, fn in (("D2H", lambda: h.copy_(d, non_blocking=True)), ("H2D", lambda: d.copy_(h, non_blocking=True))):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    print("%s 64 MiB pinned: %.1f GB/s" % (name, 20 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9))
