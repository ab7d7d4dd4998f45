"""measures the pinned host -> device copy rate of this box (the bound of bench.py's end-to-end number)"""
import torch
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for size in (n, 358 << 20, 64 << 20):
    best = 0.0
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        d[:size].copy_(h[:size], non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, size / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    print(f"H2D pinned {size >> 20} MiB: {best:.1f} GB/s")
