"""One-off: pinned host->device copy rate of one 4K 10-bit window (15 x 24.9 MB) on this box."""
import torch, time
n = 15 * 24883200
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(3):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    d.copy_(h, non_blocking=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"H2D pinned: {n / 1e6:.0f} MB in {ms:.2f} ms = {n / ms / 1e6:.1f} GB/s")
