"""Raw pinned-memory PCIe bandwidth (development aid): D2H alone, H2D alone, both directions at once."""
import torch, time
n = 256 << 20
hp_in = torch.empty(n, dtype=torch.uint8).pin_memory(); hp_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(hp_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): hp_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
for _ in range(2): run(True, True, 2)
t = run(True, False); print(f"H2D alone  : {n / t / 1e9:.1f} GB/s")
t = run(False, True); print(f"D2H alone  : {n / t / 1e9:.1f} GB/s")
t = run(True, True); print(f"both       : {n / t / 1e9:.1f} GB/s each direction ({2 * n / t / 1e9:.1f} total)")
