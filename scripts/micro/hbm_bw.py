"""HBM bandwidth by access mix (development aid): copy (1 read : 1 write, what MEASURED_PEAKS.json quotes), write-only, and a
1 : 13 read : write mix like the Jacobian kernels (read z, write J)."""
import torch
n = 1 << 30
a = torch.empty(n, dtype=torch.uint8, device="cuda"); b = torch.empty(n, dtype=torch.uint8, device="cuda")
def t(f, reps=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best
print(f"copy  (read + write bytes): {2 * n / t(lambda: b.copy_(a)) / 1e9:.0f} GB/s")
print(f"write only (fill)         : {n / t(lambda: b.zero_()) / 1e9:.0f} GB/s")
af = a.view(torch.float32)
print(f"read only (sum)           : {n / t(lambda: af.sum()) / 1e9:.0f} GB/s")
