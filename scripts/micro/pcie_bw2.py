"""Two processes, one GPU each, raw pinned D2H + H2D at the same time (does the host side scale across GPUs?)."""
import os, sys, time, torch, torch.multiprocessing as mp

def work(rank, aff):
    if aff == "numa":
        # bind to the CPUs of the GPU's NUMA node before allocating pinned memory
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(rank)
        try:
            mask = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
            cpus = [i for i in range(os.cpu_count()) if (mask[i // 64] >> (i % 64)) & 1]
            os.sched_setaffinity(0, cpus)
        except Exception as e:
            print("affinity failed", e)
    torch.cuda.set_device(rank)
    n = 256 << 20
    hp_in = torch.empty(n, dtype=torch.uint8).pin_memory(); hp_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def run(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(reps):
            with torch.cuda.stream(s1): d_in.copy_(hp_in, non_blocking=True)
            with torch.cuda.stream(s2): hp_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
    run(3)
    t = run(40)
    print(f"[{aff}] rank {rank}: {n / t / 1e9:.1f} GB/s each direction, cpus={len(os.sched_getaffinity(0))}", flush=True)

if __name__ == "__main__":
    for aff in ("none", "numa"):
        mp.spawn(work, args=(aff,), nprocs=2, join=True)
