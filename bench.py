#!/usr/bin/env python
"""bench.py — knot-point Jacobian evals/sec (discrete_jacobian! RK4), the BASELINE.json metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cartpole|quadrotor|satellite]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one batched discrete_jacobian! call over one batch of synthetic knot points (default workload: BASELINE
configs[1], Cartpole RK4, 2^20 knot points, fp64).  One process per GPU; every rank evaluates its own batch of the full
size (weak scaling, no data-path collective: knot points are independent), the timed region is bracketed by a barrier
and a device synchronize, timed with CUDA events on the launching stream, and the MAX over ranks is reported.

  value      evals/s with inputs and outputs resident in HBM (kernel path through the C ABI with device pointers)
  e2e        the same metric through the same C-ABI call with HOST (pinned) buffers: H2D of [x;u], kernel, D2H of J inside
             the timed region, every step
  roofline   algorithmic bytes (read [x;u], write J: SURVEY.md §8d) / measured launch duration vs the measured HBM peak
  cpu_baseline  the CPU oracle (a port of the reference's ForwardAD path; Julia is not installed) on the host cores

`--impl reference` times that CPU path alone, as the reference arm (the reference is pure Julia and cannot run here).
The oracle is only ever the baseline / checker here, never the measured product path.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "knot-point Jacobian evals/sec (discrete_jacobian! RK4)"
UNIT = "evals/s"

WORKLOADS = {
    # name: (description, n, m, N, numpy dtype name, dt)
    "cartpole": ("BASELINE configs[1]: Cartpole (n=4,m=1) RK4 discrete_jacobian!, 2^20 knot points, fp64", 4, 1, 1 << 20, "float64", 0.01),
    "quadrotor": ("BASELINE configs[2]: Quadrotor RigidBody{QuatRotation} (n=13,m=4) RK4, 262144 knot points, fp32", 13, 4, 262144, "float32", 0.01),
    "satellite": ("BASELINE configs[3]: Satellite RigidBody{MRP} (n=12,m=6) RK2, 2^20 knot points, fp64", 12, 6, 1 << 20, "float64", 0.1),
}
SWEEP_DESC = ("BASELINE configs[4]: mixed trajectory sweep, 4096 trajectories x 256 knot points, first half Cartpole (fp64), second half "
              "Quadrotor (fp32), RK4, per-trajectory dt, contiguous trajectory blocks per rank (strong scaling)")


def make_inputs(n, m, N, dtype, seed):
    rng = np.random.default_rng(seed)
    Z = rng.random((N, n + m))
    if n >= 12:
        q = rng.standard_normal((N, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
        if n == 13:
            Z[:, 3:7] = q
        else:
            Z[:, 3:6] = q[:, 1:] / (1.0 + np.abs(q[:, :1]))
    return Z.astype(dtype)


def oracle_model(name):
    from oracle import rd_oracle as o
    return {"cartpole": (o.cartpole, o.RK4), "quadrotor": (o.quadrotor, o.RK4), "satellite": (lambda: o.satellite(o.ROT_MRP), o.RK2)}[name]


def gpu_model(name, rd):
    return {"cartpole": (rd.Cartpole, rd.RK4), "quadrotor": (rd.Quadrotor, rd.RK4), "satellite": (lambda: rd.Satellite(rd.MRP), rd.RK2)}[name]


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_rate(name, budget_s, threads):
    """evals/s of the CPU oracle (forward-mode Dual path == the reference's default ForwardAD path) on `threads` host threads;
    returns (rate, sample description)."""
    from oracle import rd_oracle as o
    mk, Q = oracle_model(name)
    _, n, m, N, _, dt = WORKLOADS[name]
    model = mk()
    Zc = make_inputs(n, m, 1 << 14, "float64", 99)
    out = np.empty((Zc.shape[0], n + m, n))
    o.discrete_jacobian(model, Q, Zc, dt, nthreads=threads, out=out)                # warm-up (OpenMP pool, page faults)
    t0 = time.perf_counter(); o.discrete_jacobian(model, Q, Zc, dt, nthreads=threads, out=out); cal = time.perf_counter() - t0
    per_step = max(1 << 14, min(N, int((1 << 14) / max(cal, 1e-6) * min(budget_s, 1.0))))
    Z = make_inputs(n, m, per_step, "float64", 100)
    out = np.empty((per_step, n + m, n))
    o.discrete_jacobian(model, Q, Z, dt, nthreads=threads, out=out)
    reps, t0 = 0, time.perf_counter()
    while True:
        o.discrete_jacobian(model, Q, Z, dt, nthreads=threads, out=out)
        reps += 1
        el = time.perf_counter() - t0
        if el >= budget_s or reps >= 1000:
            break
    return per_step * reps / el, f"{reps} passes over {per_step} of the workload's knot points ({el:.1f} s of CPU time), method=ForwardAD port", per_step


def run_reference(args):
    """Reference arm: the reference's own algorithm on the host CPU (oracle port; Julia absent), all host threads."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    from oracle import rd_oracle as o
    name = args.workload
    desc, n, m, N, dtn, dt = WORKLOADS[name]
    threads = host_cores()
    mk, Q = oracle_model(name)
    model = mk()
    # size one step so that (steps + warmup) passes stay within ~2 minutes
    rate, _, _ = cpu_rate(name, 2.0, threads)
    per_step = int(min(N, max(1 << 12, rate * min(1.0, 100.0 / max(1, args.steps + args.warmup)))))
    Z = make_inputs(n, m, per_step, "float64", 100)
    out = np.empty((per_step, n + m, n))
    for _ in range(args.warmup):
        o.discrete_jacobian(model, Q, Z, dt, nthreads=threads, out=out)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.discrete_jacobian(model, Q, Z, dt, nthreads=threads, out=out)
    el = time.perf_counter() - t0
    val = per_step * args.steps / el
    sample = f"each step = {per_step} of the workload's {N} knot points, fp64, ForwardAD (Dual<{n + m}>) port of the reference path"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": el / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": desc, "note": "reference is pure Julia (not installed); CPU oracle port timed on host cores"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


class ClockSampler(threading.Thread):
    """Polls SM clock and clock-event reasons through NVML while the timed region runs."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap", 0x80: "hw_power_brake"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz, self.ok = index, [], set(), False, None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        while self.ok and not self.stop_flag:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.005)

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def physical_gpu_index(local):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local])
        except (ValueError, IndexError):
            return local
    return local


def run_sweep(args):
    """BASELINE configs[4]: the 4096 x 256 mixed sweep, sharded by contiguous trajectory blocks (strong scaling: total work fixed)."""
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import rdb200 as rd
    from rdb200 import sharding as sh
    ntraj, K = 4096, 256
    segs = sh.partition_segments({"cartpole": ntraj // 2, "quadrotor": ntraj // 2}, world, rank)
    work = []
    for name, (lo, hi) in segs.items():
        _, n, m, _, dtn, _ = WORKLOADS[name]
        mk, Q = gpu_model(name, rd)
        model = mk()
        cnt = (hi - lo) * K
        nsets = 6
        Zs = [torch.from_numpy(make_inputs(n, m, cnt, dtn, 17 * rank + i)).cuda() for i in range(nsets)]
        dt = torch.from_numpy(np.repeat(0.01 * (1 + np.arange(lo, hi) % 4), K)).cuda()
        Js = [torch.empty((cnt, n + m, n), dtype=Zs[0].dtype, device="cuda") for _ in range(nsets)]
        work.append((model._h, Q.code, Zs, dt, Js, cnt, np.dtype(dtn).itemsize * ((n + m) + n * (n + m))))

    # The two model segments are independent: one stream each, so the tail of one kernel overlaps the head of the other and, at
    # 4-8 GPUs, the small launch-bound kernels overlap (measured at 1 GPU: two streams 122.6 us, one stream 131.9 us per sweep).
    nstreams = args.sweep_streams or 2
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    streams = [streams[i % nstreams] for i in range(len(work))]
    main = torch.cuda.current_stream()

    def step(i):
        for (h, qc, Zs, dt, Js, _, _), st in zip(work, streams):
            with torch.cuda.stream(st):
                h.discrete_jacobian(qc, Zs[i % len(Zs)], dt, J=Js[i % len(Zs)])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for st in streams:
        st.wait_stream(main)
    for i in range(args.steps):
        step(i)
    for st in streams:
        main.wait_stream(st)
    e1.record(main)
    barrier()
    ms = sh.barrier_max_ms(e0.elapsed_time(e1), device=torch.device("cuda", local))
    if rank == 0:
        total = ntraj * K
        byts = sum(c * b for *_, c, b in work) * world        # every rank holds an equal share of both segments
        line = {"metric": METRIC, "value": total * args.steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64+f32", "data": "synthetic",
                "config": {"workload": SWEEP_DESC, "l2": "6 rotating buffer sets per segment", "parallelism": f"trajectory-block sharded x{world}",
                           "streams": nstreams},
                "roofline": {"bound": "hbm", "achieved": byts / (ms * 1e-3 / args.steps) / 1e9 / world, "peak": 6551.4, "unit": "GB/s per GPU",
                             "frac": byts / (ms * 1e-3 / args.steps) / 1e9 / world / 6551.4, "traffic": None},
                "gpu_launches": 2 * args.steps}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_ours(args):
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible — the product path has no CPU fallback")
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import rdb200 as rd

    name = args.workload
    desc, n, m, N, dtn, dt = WORKLOADS[name]
    mk, Q = gpu_model(name, rd)
    model = mk()
    h = model._h
    es = np.dtype(dtn).itemsize
    alg_bytes = es * ((n + m) + n * (n + m))                     # read z, write J
    tdt = torch.float64 if dtn == "float64" else torch.float32
    # rotate over enough buffer sets that the bytes touched between two uses of a set exceed the 126 MB L2 several times
    nsets = max(2, int(np.ceil(600e6 / (N * alg_bytes))) + 1)
    Zs = [torch.from_numpy(make_inputs(n, m, N, dtn, 1000 * rank + i)).cuda() for i in range(nsets)]
    Js = [torch.empty((N, n + m, n), dtype=tdt, device="cuda") for _ in range(nsets)]
    qc = Q.code

    def step(i):
        h.discrete_jacobian(qc, Zs[i % nsets], dt, J=Js[i % nsets])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(physical_gpu_index(local))
    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    # keep the identical kernel running a little longer (untimed) if the timed region was too short for the clock sampler
    t_extra = time.perf_counter()
    while sampler.ok and len(sampler.samples) < 20 and time.perf_counter() - t_extra < 2.0:
        for i in range(50):
            step(i)
        torch.cuda.synchronize()
    sampler.stop_flag = True
    from rdb200 import sharding as sh
    ms_max = sh.barrier_max_ms(ms, device=torch.device("cuda", local))
    value = world * N * args.steps / (ms_max * 1e-3)

    # ---- end to end through the C ABI with HOST buffers (pinned): H2D + kernel + D2H inside the timed region ----------
    Zp = rd.PinnedArray((N, n + m), dtn); Jp = rd.PinnedArray((N, n + m, n), dtn)
    Zp.array[...] = make_inputs(n, m, N, dtn, 7 + rank)
    e2e_steps = max(3, min(args.steps, 20))
    for _ in range(2):
        h.discrete_jacobian(qc, Zp.array, dt, J=Jp.array)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        h.discrete_jacobian(qc, Zp.array, dt, J=Jp.array)
        _ = float(Jp.array[-1, 0, 0])                            # read a result on the host
    t_e2e = time.perf_counter() - t0
    t_e2e = sh.barrier_max_ms(t_e2e * 1e3, device=torch.device("cuda", local)) * 1e-3
    e2e_val = world * N * e2e_steps / t_e2e
    # light parity guard on what was just timed (device vs host path agree bit-for-bit on the same inputs)
    Jd = h.discrete_jacobian(qc, torch.from_numpy(Zp.array).cuda(), dt)
    assert np.array_equal(Jd.cpu().numpy(), Jp.array), "device and host paths disagree"

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        launch_s = ms_max * 1e-3 / args.steps
        achieved = N * alg_bytes / launch_s / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(name)
        cpu = {"value": None, "unit": UNIT, "cores": host_cores(), "kind": "port", "sample": "skipped (N>1)"}
        if world == 1 and not args.no_cpu_baseline:
            r, sample, _ = cpu_rate(name, args.cpu_budget, host_cores())
            cpu = {"value": r, "unit": UNIT, "cores": host_cores(), "kind": "port", "sample": sample}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64" if dtn == "float64" else "f32", "data": "synthetic",
            "config": {"workload": desc, "knot_points_per_gpu": N, "layout": "reference AoS (Z (n+m,N), J (n,n+m,N) column-major)",
                       "l2": f"inputs larger than L2: {nsets} rotating buffer sets x {N * alg_bytes / 1e6:.0f} MB", "parallelism": f"knot-sharded x{world}"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "kernel": "rdb::knot_kernel", "algorithmic_bytes_per_eval": alg_bytes, "peak_source": peak_src},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": N * (n + m) * es, "d2h_bytes_per_step": N * n * (n + m) * es,
                    "steps": e2e_steps, "path": "rdb_discrete_jacobian with pinned host pointers (3-stream chunked pipeline)"},
            "gpu_launches": args.steps,
            "clocks": sampler.summary(),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cartpole", choices=sorted(WORKLOADS) + ["sweep"])
    ap.add_argument("--cpu-budget", type=float, default=10.0, help="seconds of CPU time for the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sweep-streams", type=int, default=0, help="sweep workload: streams for the two model segments (0 = auto)")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.workload == "sweep":
            args.workload = "cartpole"
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun when called directly with --gpus N
        import subprocess
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    if args.workload == "sweep":
        return run_sweep(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
