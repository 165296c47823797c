#!/usr/bin/env python
"""bench.py — knot-point Jacobian evals/sec (discrete_jacobian! RK4), the BASELINE.json metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cartpole|quadrotor|satellite|sweep]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one batched discrete_jacobian! call over one batch of synthetic knot points (default workload: BASELINE
configs[1], Cartpole RK4, 2^20 knot points, fp64).  One process per GPU; every rank evaluates its own batch of the full
size (weak scaling, no data-path collective: knot points are independent).  The timed region is EXACTLY K steps, bracketed by a
barrier and a device synchronize on both sides, timed with CUDA events on the launching stream; the MAX over ranks is reported
and every rank's own time is listed (`per_rank_ms_per_step`).

  value      evals/s with inputs and outputs resident in HBM (one pre-validated C-ABI launch per step, rdb_plan_launch)
  e2e        the same metric through the C-ABI call with HOST (pinned) buffers: H2D of [x;u], kernel, D2H of J inside the timed
             region, every step
  roofline   algorithmic bytes (read [x;u], write J: SURVEY.md §8d) / measured launch duration vs the measured HBM peak
  cpu_baseline  the CPU oracle (a port of the reference's ForwardAD path; Julia is not installed) on the host cores
  extra      the other BASELINE configurations measured the same way in the same process: configs[2] (Quadrotor RK4 fp32, plus its
             error-state form), configs[3] (Satellite{MRP} RK2 fp64) — weak, one full batch per GPU — and configs[4], the mixed
             4096 x 256 sweep, STRONG-scaled over the ranks by contiguous trajectory blocks

Launch gate: before the first event a ~2 ms device-side sleep is enqueued, so that the host has queued all K launches before the
GPU reaches the timed region — the events then bracket K back-to-back kernel executions and no host launch latency (the quantity
the metric names; without the gate a 20-step region of 0.8 ms carries ~2 % of host latency of the first launch).

`--impl reference` times the reference's own CPU path alone (a Julia installation with RobotDynamics.jl if the box has one, else
the C++ port of its ForwardAD path).  The oracle is only ever the baseline / checker here, never the measured product path.
"""
import argparse
import json
import os
import shutil
import subprocess
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
GATE_CYCLES = 4_000_000          # ~2 ms of device-side sleep in front of the timed region (see module docstring)


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


def julia_reference(name, steps, warmup):
    """The REAL reference on this box, if it has one: `julia` on PATH and RobotDynamics.jl loadable in oracle/ref_julia's project
    (or a driver-installed copy under baseline/_ref).  Returns the parsed JSON of oracle/ref_julia/bench_ref.jl or None."""
    exe = shutil.which("julia")
    if not exe:
        return None
    env = dict(os.environ)
    ref = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(ref):
        env["RD_REF"] = ref
    try:
        p = subprocess.run([exe, "--project=" + os.path.join(ROOT, "oracle", "ref_julia"), "-t", "auto",
                            os.path.join(ROOT, "oracle", "ref_julia", "bench_ref.jl"), name, str(steps), str(warmup)],
                           capture_output=True, text=True, timeout=900, env=env)
        for ln in reversed(p.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)
    except Exception:
        return None
    return None


def run_reference(args):
    """Reference arm: the reference's own algorithm on the host CPU, all host threads.  Probes for a Julia installation first (the
    real RobotDynamics.jl, `kind: reference`); the images used so far have none, then the C++ port of its ForwardAD path is timed."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    name = args.workload
    desc, n, m, N, dtn, dt = WORKLOADS[name]
    threads = host_cores()
    jl = julia_reference(name, args.steps, args.warmup)
    if jl and "value" in jl:
        val, ms, kind, cores = float(jl["value"]), float(jl["ms_per_step"]), "reference", int(jl.get("threads", threads))
        sample = jl.get("sample", "RobotDynamics.jl jacobian!(StaticReturn(), ForwardAD(), ...) loop, Threads.@threads")
        note = "RobotDynamics.jl v0.4.8 run in Julia on the host cores"
    else:
        from oracle import rd_oracle as o
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
        val, ms, kind, cores = per_step * args.steps / el, el / args.steps * 1e3, "port", threads
        sample = f"each step = {per_step} of the workload's {N} knot points, fp64, ForwardAD (Dual<{n + m}>) port of the reference path"
        note = "reference is pure Julia (probed: no `julia` on this box); CPU oracle port timed on host cores"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": desc, "note": note},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


class ClockSampler(threading.Thread):
    """Polls SM clock and clock-event reasons through NVML while the measurement runs; every sample carries a host timestamp so that
    the ones taken inside the timed region can be told apart from the ones taken during the warm-up steps before it."""
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
                mhz = self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)
                r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.perf_counter(), mhz))
                for bit, nm in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.0002)

    def summary(self, load_window, timed_window):
        under = [m for (ts, m) in self.samples if load_window[0] <= ts <= load_window[1]]
        timed = [m for (ts, m) in self.samples if timed_window[0] <= ts <= timed_window[1]]
        if not self.ok or not under:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(under)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(under),
                "samples_in_timed_region": len(timed), "sm_mhz_min": float(min(under)),
                "how": "NVML polled every ~0.2 ms from the first warm-up step to the barrier after the timed region; samples_in_timed_region "
                       "= those taken between the launch gate and the closing synchronize"}


def physical_gpu_index(local):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local])
        except (ValueError, IndexError):
            return local
    return local


class Bench:
    """Shared plumbing of the timed regions: rank / world, barrier, gate, per-rank gather."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device visible — the product path has no CPU fallback")
        self.torch, self.dist = torch, dist
        self.rank, self.world, self.local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
        torch.cuda.set_device(self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        import rdb200 as rd
        self.rd = rd
        self.launches = 0

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def gather(self, x):
        """every rank's value of a python float, as a list (rank order)"""
        if self.world == 1:
            return [float(x)]
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device="cuda")
        out = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [float(o.item()) for o in out]

    def timed(self, step, steps, warmup, streams=None):
        """W untimed + EXACTLY `steps` timed calls of step(i); returns (local ms, host window of the timed region).  `streams`: side
        streams the step uses (forked from / joined into the current stream around the timed region)."""
        torch = self.torch
        main = torch.cuda.current_stream()
        for i in range(max(warmup, 3)):
            step(i)
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(GATE_CYCLES)                    # launch gate: the GPU sleeps while the host queues the whole timed region
        t_host0 = time.perf_counter()
        e0.record(main)
        for st in streams or ():
            st.wait_stream(main)
        for i in range(steps):
            step(i)
        for st in streams or ():
            main.wait_stream(st)
        e1.record(main)
        self.barrier()
        t_host1 = time.perf_counter()
        return e0.elapsed_time(e1), (t_host0, t_host1)

    def close(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


class KnotWorkload:
    """One BASELINE configuration on this rank: rotating device buffer sets (more bytes between two uses of a set than the 126 MB L2
    holds several times over) and one pre-validated plan per set."""

    def __init__(self, B, name, N=None, error_state=False, seed0=0):
        rd, torch = B.rd, B.torch
        self.B, self.name = B, name
        self.desc, self.n, self.m, Nw, self.dtn, self.dt = WORKLOADS[name]
        self.N = Nw if N is None else N
        mk, Q = gpu_model(name, rd)
        self.model, self.Q = mk(), Q
        h = self.model._h
        self.es = np.dtype(self.dtn).itemsize
        jr, jc = (h.nerr, h.nerr + h.m) if error_state else (h.n, h.n + h.m)
        self.alg_bytes = self.es * ((h.n + h.m) + jr * jc)                      # read z, write J
        tdt = torch.float64 if self.dtn == "float64" else torch.float32
        self.nsets = max(2, int(np.ceil(600e6 / (self.N * self.alg_bytes))) + 1)
        self.Zs = [torch.from_numpy(make_inputs(self.n, self.m, self.N, self.dtn, seed0 + 1000 * B.rank + i)).cuda() for i in range(self.nsets)]
        self.Js = [torch.empty((self.N, jc, jr), dtype=tdt, device="cuda") for _ in range(self.nsets)]
        op = rd._abi.OP_DISCRETE_ERROR_JACOBIAN if error_state else rd._abi.OP_DISCRETE_JACOBIAN
        self.plans = [rd._abi.Plan(h, op, Q.code, Z, self.dt, J=J) for Z, J in zip(self.Zs, self.Js)]

    def step(self, i):
        self.plans[i % self.nsets].launch()
        self.B.launches += 1

    def measure(self, steps, warmup, peak):
        ms, window = self.B.timed(self.step, steps, warmup)
        per_rank = self.B.gather(ms / steps)
        ms_step = max(per_rank)
        achieved = self.N * self.alg_bytes / (ms_step * 1e-3) / 1e9
        return {"ms_per_step": ms_step, "per_rank_ms_per_step": per_rank, "value": self.B.world * self.N / (ms_step * 1e-3),
                "achieved_GBs": achieved, "frac": achieved / peak, "algorithmic_bytes_per_eval": self.alg_bytes,
                "knot_points_per_gpu": self.N, "buffer_sets": self.nsets}, window


SWEEP_STREAMS = 2


class SweepWorkload:
    """BASELINE configs[4]: 4096 trajectories x 256 knot points, first half Cartpole fp64, second half Quadrotor fp32, per-trajectory
    dt; STRONG scaling — every rank takes its contiguous share of both segments (sharding.partition_segments).  The two segments are
    independent: each runs on its own stream, joined at the end of the timed region."""

    def __init__(self, B):
        rd, torch = B.rd, B.torch
        from rdb200 import sharding as sh
        self.B = B
        self.ntraj, self.K = 4096, 256
        segs = sh.partition_segments({"cartpole": self.ntraj // 2, "quadrotor": self.ntraj // 2}, B.world, B.rank)
        self.work, self.meta, self.bytes_local, self.knots_local = [], [], 0, 0
        # RDB_SWEEP_STREAMS=1: both segments back to back on one stream (experiments); default one stream per segment
        shared = torch.cuda.Stream() if os.environ.get("RDB_SWEEP_STREAMS", str(SWEEP_STREAMS)) == "1" else None
        if os.environ.get("RDB_SWEEP_ORDER") == "quadfirst":
            segs = dict(reversed(list(segs.items())))
        for name, (lo, hi) in segs.items():
            _, n, m, _, dtn, _ = WORKLOADS[name]
            mk, Q = gpu_model(name, rd)
            model = mk()
            cnt = (hi - lo) * self.K
            if cnt == 0:
                continue
            per = np.dtype(dtn).itemsize * ((n + m) + n * (n + m))
            nsets = max(2, int(np.ceil(600e6 / (cnt * per))) + 1)
            nsets = min(nsets, 48)
            Zs = [torch.from_numpy(make_inputs(n, m, cnt, dtn, 17 * B.rank + i)).cuda() for i in range(nsets)]
            dt = torch.from_numpy(np.repeat(0.01 * (1 + np.arange(lo, hi) % 4), self.K)).cuda()
            Js = [torch.empty((cnt, n + m, n), dtype=Zs[0].dtype, device="cuda") for _ in range(nsets)]
            # the two segments overlap on the GPU (one stream each): their plans say so (rdb_plan_set_shared)
            plans = [rd._abi.Plan(model._h, rd._abi.OP_DISCRETE_JACOBIAN, Q.code, Z, dt, J=J, shared_gpu=shared is None) for Z, J in zip(Zs, Js)]
            self.work.append((model, plans, shared if shared is not None else torch.cuda.Stream()))
            self.meta.append((Q.code, n, m, self.ntraj // 2))
            self.bytes_local += cnt * per
            self.knots_local += cnt
        self.streams = list({id(w[2]): w[2] for w in self.work}.values())

    def step(self, i):
        for _, plans, st in self.work:
            plans[i % len(plans)].launch(st.cuda_stream)
            self.B.launches += 1

    def measure_with_allgather(self, steps, warmup):
        """SURVEY.md §8e's optional epilogue, reported separately: every rank ends the step holding the Jacobians of ALL trajectories
        (ncclAllGather of each segment's J on that segment's stream, right behind its kernel).  Not part of `value`: the path itself has
        no exchange step, and the number shows why a consumer should stay sharded (or all-gather the 40 / 68-byte inputs and recompute)."""
        torch, dist = self.B.torch, self.B.dist
        full = [torch.empty((self.B.world,) + tuple(plans[0].J.shape), dtype=plans[0].J.dtype, device="cuda") for _, plans, _ in self.work]

        def step(i):
            for (_, plans, st), out in zip(self.work, full):
                pl = plans[i % len(plans)]
                pl.launch(st.cuda_stream)
                with torch.cuda.stream(st):
                    dist.all_gather_into_tensor(out, pl.J)
                self.B.launches += 1

        ms, _ = self.B.timed(step, steps, warmup, streams=self.streams)
        per_rank = self.B.gather(ms / steps)
        ms_step = max(per_rank)
        # parity of the gathered copy: this rank's slice of the last step equals its own J
        ok = all(bool(torch.equal(out[self.B.rank], plans[(steps - 1) % len(plans)].J)) for (_, plans, _), out in zip(self.work, full))
        return {"ms_per_step": ms_step, "per_rank_ms_per_step": per_rank, "value": self.ntraj * self.K / (ms_step * 1e-3),
                "gathered_bytes_per_rank_per_step": int(sum(o.numel() * o.element_size() for o in full)), "gathered_equals_local": ok,
                "collective": "ncclAllGather (torch.distributed all_gather_into_tensor) per segment, on the segment's stream"}

    def measure_replicated(self, steps, warmup):
        """The alternative to gathering J: all-gather the INPUTS (40 / 68 bytes per knot instead of 160 / 884) and let every rank evaluate
        the Jacobians of all trajectories itself — recomputing is cheaper than moving them over NVLink."""
        torch, dist, rd, W = self.B.torch, self.B.dist, self.B.rd, self.B.world
        full = []
        for (model, plans, st), (qc, n, m, ntr) in zip(self.work, self.meta):
            Zf = torch.empty((W,) + tuple(plans[0].Z.shape), dtype=plans[0].Z.dtype, device="cuda")
            Jf = torch.empty((W * plans[0].Z.shape[0], n + m, n), dtype=Zf.dtype, device="cuda")
            dt = torch.from_numpy(np.repeat(0.01 * (1 + np.arange(0, ntr) % 4), self.K)).cuda()
            full.append((Zf, Jf, rd._abi.Plan(model._h, rd._abi.OP_DISCRETE_JACOBIAN, qc, Zf.view(-1, n + m), dt, J=Jf)))

        def step(i):
            for (_, plans, st), (Zf, Jf, pl) in zip(self.work, full):
                with torch.cuda.stream(st):
                    dist.all_gather_into_tensor(Zf, plans[i % len(plans)].Z)
                pl.launch(st.cuda_stream)
                self.B.launches += 1

        ms, _ = self.B.timed(step, steps, warmup, streams=self.streams)
        per_rank = self.B.gather(ms / steps)
        ms_step = max(per_rank)
        # this rank's slice of the replicated result equals what its own shard computes from the same inputs
        ok = True
        for (_, plans, st), (Zf, Jf, pl) in zip(self.work, full):
            mine = plans[(steps - 1) % len(plans)]
            mine.launch(st.cuda_stream); st.synchronize()
            cnt = mine.Z.shape[0]
            ok = ok and bool(torch.equal(Jf[self.B.rank * cnt:(self.B.rank + 1) * cnt], mine.J))
        return {"ms_per_step": ms_step, "per_rank_ms_per_step": per_rank, "value": self.ntraj * self.K / (ms_step * 1e-3),
                "gathered_bytes_per_rank_per_step": int(sum(f[0].numel() * f[0].element_size() for f in full)), "replica_equals_shard": ok,
                "collective": "ncclAllGather of Z per segment, then the full-size Jacobian launch on every rank"}

    def measure(self, steps, warmup, peak):
        ms, _ = self.B.timed(self.step, steps, warmup, streams=self.streams)
        per_rank = self.B.gather(ms / steps)
        ms_step = max(per_rank)
        tot_bytes = sum(self.B.gather(self.bytes_local))
        achieved = tot_bytes / self.B.world / (ms_step * 1e-3) / 1e9
        return {"workload": SWEEP_DESC, "scaling": "strong", "ms_per_step": ms_step, "per_rank_ms_per_step": per_rank,
                "value": self.ntraj * self.K / (ms_step * 1e-3), "achieved_GBs_per_gpu": achieved, "frac": achieved / peak,
                "launches_per_step_per_rank": len(self.work), "streams": len(self.streams)}


def hbm_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def run_ours(args):
    B = Bench()
    torch, rd = B.torch, B.rd
    peak, peak_src = hbm_peak()
    name = args.workload
    if name == "sweep":
        sw = SweepWorkload(B)
        res = sw.measure(args.steps, args.warmup, peak)
        if B.rank == 0:
            line = {"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": B.world, "steps": args.steps, "warmup": max(args.warmup, 3),
                    "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64+f32",
                    "data": "synthetic", "config": {"workload": SWEEP_DESC, "parallelism": f"trajectory-block sharded x{B.world}"},
                    "roofline": {"bound": "hbm", "achieved": res["achieved_GBs_per_gpu"], "peak": peak, "unit": "GB/s per GPU", "frac": res["frac"], "traffic": None},
                    "per_rank_ms_per_step": res["per_rank_ms_per_step"], "gpu_launches": B.launches}
            print(json.dumps(line), flush=True)
        B.close()
        return 0

    W = KnotWorkload(B, name)
    h, qc, N, n, m, dtn, dt, es = W.model._h, W.Q.code, W.N, W.n, W.m, W.dtn, W.dt, W.es

    # ---- clocks: NVML polled (~5 kHz) while the warm-up, the launch gate and the timed region run --------------------------------------
    # (no pre-roll: tens of milliseconds of this kernel back to back take the board into sw_power_cap and ~6 % lower clocks — measured
    #  41.9 us per step after an 80 ms pre-roll against 39.5 us without; the number reported is the one of the K timed steps as launched)
    sampler = ClockSampler(physical_gpu_index(B.local))
    sampler.start()
    launches_before = B.launches
    t_load0 = time.perf_counter()
    main_res, timed_window = W.measure(args.steps, args.warmup, peak)
    t_load1 = time.perf_counter()
    timed_launches = B.launches - launches_before - max(args.warmup, 3)
    sampler.stop_flag = True
    clocks = sampler.summary((t_load0, t_load1), timed_window)

    # ---- end to end through the C ABI with HOST buffers (pinned): H2D + kernel + D2H inside the timed region ----------
    from rdb200 import sharding as sh
    Zp = rd.PinnedArray((N, n + m), dtn); Jp = rd.PinnedArray((N, n + m, n), dtn)
    Zp.array[...] = make_inputs(n, m, N, dtn, 7 + B.rank)
    e2e_steps = max(3, min(args.steps, 20))
    for _ in range(2):
        h.discrete_jacobian(qc, Zp.array, dt, J=Jp.array)
    B.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        h.discrete_jacobian(qc, Zp.array, dt, J=Jp.array)
        _ = float(Jp.array[-1, 0, 0])                            # read a result on the host
    t_e2e = time.perf_counter() - t0
    e2e_per_rank = B.gather(t_e2e * 1e3 / e2e_steps)
    e2e_val = B.world * N / (max(e2e_per_rank) * 1e-3)
    e2e_launches = e2e_steps * int(np.ceil(N / 65536))           # the host pipeline evaluates 65536-knot chunks
    # light parity guard on what was just timed (device vs host path agree bit-for-bit on the same inputs)
    Jd = h.discrete_jacobian(qc, torch.from_numpy(Zp.array).cuda(), dt)
    assert np.array_equal(Jd.cpu().numpy(), Jp.array), "device and host paths disagree"
    del Zp, Jp, Jd

    # ---- the other BASELINE configurations, same process, same method --------------------------------------------------------------
    extra = {}
    if not args.no_extra and name == "cartpole":
        del W.plans, W.Zs, W.Js
        torch.cuda.empty_cache()
        for key, wl, err in (("quadrotor_c3", "quadrotor", False), ("quadrotor_c3_error_state", "quadrotor", True), ("satellite_c4", "satellite", False)):
            X = KnotWorkload(B, wl, error_state=err, seed0=50)
            r, _ = X.measure(args.steps, args.warmup, peak)
            r["workload"] = WORKLOADS[wl][0] + (" — error-state form G(x+)' [A B] blkdiag(G(x), I), 12 x 16 per knot" if err else "")
            r["scaling"] = "weak"
            extra[key] = r
            del X
            torch.cuda.empty_cache()
        sw = SweepWorkload(B)
        extra["sweep_c5"] = sw.measure(args.steps, args.warmup, peak)
        if B.world > 1:
            extra["sweep_c5_with_allgather_of_J"] = sw.measure_with_allgather(min(args.steps, 20), args.warmup)
            extra["sweep_c5_replicated_via_allgather_of_Z"] = sw.measure_replicated(min(args.steps, 20), args.warmup)
        del sw
        torch.cuda.empty_cache()

    if B.rank == 0:
        traffic, tsrc = None, None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            tj = json.load(open(tp))
            traffic, tsrc = tj.get(name), tj.get("_source", "profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel (not measured in this run)")
        cpu = {"value": None, "unit": UNIT, "cores": host_cores(), "kind": "port", "sample": "skipped (N>1)"}
        if B.world == 1 and not args.no_cpu_baseline:
            r, sample, _ = cpu_rate(name, args.cpu_budget, host_cores())
            cpu = {"value": r, "unit": UNIT, "cores": host_cores(), "kind": "port", "sample": sample}
        line = {
            "metric": METRIC, "value": main_res["value"], "unit": UNIT, "n_gpus": B.world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": main_res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64" if dtn == "float64" else "f32", "data": "synthetic",
            "config": {"workload": W.desc, "knot_points_per_gpu": N, "layout": "reference AoS (Z (n+m,N), J (n,n+m,N) column-major)",
                       "l2": f"inputs larger than L2: {main_res['buffer_sets']} rotating buffer sets x {N * W.alg_bytes / 1e6:.0f} MB",
                       "parallelism": f"knot-sharded x{B.world}",
                       "launch": "one rdb_plan_launch per step; ~2 ms device-side launch gate before the first event (all K launches queued "
                                 "before the timed region starts)"},
            "per_rank_ms_per_step": main_res["per_rank_ms_per_step"],
            "roofline": {"bound": "hbm", "achieved": main_res["achieved_GBs"], "peak": peak, "unit": "GB/s", "frac": main_res["frac"],
                         "traffic": traffic, "traffic_source": tsrc, "kernel": "rdb::knot_kernel", "algorithmic_bytes_per_eval": W.alg_bytes,
                         "peak_source": peak_src},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": N * (n + m) * es, "d2h_bytes_per_step": N * n * (n + m) * es,
                    "steps": e2e_steps, "per_rank_ms_per_step": e2e_per_rank, "gpu_launches": e2e_launches,
                    "path": "rdb_discrete_jacobian with pinned host pointers (3-stream chunked pipeline)"},
            "gpu_launches": timed_launches,
            "gpu_launches_total": B.launches + e2e_launches,
            "clocks": clocks,
            "extra": extra,
        }
        print(json.dumps(line), flush=True)
    B.close()
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
    ap.add_argument("--no-extra", action="store_true", help="skip the other BASELINE configurations (extra)")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.workload == "sweep":
            args.workload = "cartpole"
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun when called directly with --gpus N
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
