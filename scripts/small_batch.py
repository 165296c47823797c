"""Small-batch regime (what Altro / TrajectoryOptimization actually call: 10^2..10^3 knot points per solve iteration).

    python scripts/small_batch.py [--out profiles/small_batch_rNN.md]

For N = 64 .. 16384 and each operation of the path, device pointers:
  stream   us per call, 300 calls enqueued back to back on one stream, CUDA events (launch-throughput bound)
  sync     us per call, each call followed by a device synchronize, host wall clock (what a solver loop that needs the result sees)
  graph    us per replay of a CUDA graph holding ONE call (captured through the same C-ABI entry point), events
  host     us per call with HOST (numpy) pointers: staging H2D + kernel + D2H inside the library, wall clock — pageable arrays, and
           the same arrays page-locked in place with rdb_host_register
  plan     us per rdb_plan_launch of the same operation (validated once), back to back on one stream, events; and with a sync after each
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--reps", type=int, default=300)
    args = ap.parse_args()
    import torch
    import rdb200 as rd
    import bench

    cp, qd = rd.Cartpole(), rd.Quadrotor()
    ops = [
        ("discrete_jacobian! RK4 Cartpole fp64", cp._h, "float64", lambda h, Z, o: h.discrete_jacobian(rd.RK4.code, Z, 0.01, J=o), lambda h, N: (N, 5, 4)),
        ("discrete_jacobian! RK4 Quadrotor fp32", qd._h, "float32", lambda h, Z, o: h.discrete_jacobian(rd.RK4.code, Z, 0.01, J=o), lambda h, N: (N, 17, 13)),
        ("error-state Jacobian RK4 Quadrotor fp32", qd._h, "float32", lambda h, Z, o: h.discrete_error_jacobian(rd.RK4.code, Z, 0.01, J=o), lambda h, N: (N, 16, 12)),
        ("errstate_jacobian! Quadrotor fp32", qd._h, "float32", lambda h, Z, o: h.errstate_jacobian(Z, G=o), lambda h, N: (N, 12, 13)),
        ("discrete_dynamics RK4 Quadrotor fp32", qd._h, "float32", lambda h, Z, o: h.discrete_dynamics(rd.RK4.code, Z, 0.01, out=o), lambda h, N: (N, 13)),
    ]
    PLAN_OPS = {"discrete_jacobian! RK4 Cartpole fp64": rd._abi.OP_DISCRETE_JACOBIAN, "discrete_jacobian! RK4 Quadrotor fp32": rd._abi.OP_DISCRETE_JACOBIAN,
                "error-state Jacobian RK4 Quadrotor fp32": rd._abi.OP_DISCRETE_ERROR_JACOBIAN, "discrete_dynamics RK4 Quadrotor fp32": rd._abi.OP_DISCRETE_DYNAMICS}
    lines = ["| op | N | stream us | sync us | graph us | host-pointer us (pageable) | host-pointer us (rdb_host_register) | plan stream us | plan sync us |", "|---|---|---|---|---|---|---|---|---|"]
    for name, h, dtn, call, oshape in ops:
        for N in (64, 256, 1024, 4096, 16384):
            Zh = bench.make_inputs(h.n, h.m, N, dtn, 5)
            Z = torch.from_numpy(Zh).cuda()
            out = torch.empty(oshape(h, N), dtype=Z.dtype, device="cuda")
            outh = np.empty(oshape(h, N), dtype=Zh.dtype)
            for _ in range(10):
                call(h, Z, out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.reps):
                call(h, Z, out)
            e1.record(); torch.cuda.synchronize()
            us_stream = e0.elapsed_time(e1) / args.reps * 1e3
            t0 = time.perf_counter()
            for _ in range(args.reps):
                call(h, Z, out); torch.cuda.synchronize()
            us_sync = (time.perf_counter() - t0) / args.reps * 1e6
            g = torch.cuda.CUDAGraph()
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                call(h, Z, out)
                torch.cuda.synchronize()
                with torch.cuda.graph(g, stream=st):
                    call(h, Z, out)
            torch.cuda.synchronize()
            for _ in range(5):
                g.replay()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(args.reps):
                g.replay()
            e1.record(); torch.cuda.synchronize()
            us_graph = e0.elapsed_time(e1) / args.reps * 1e3
            for _ in range(3):
                call(h, Zh, outh)
            t0 = time.perf_counter()
            for _ in range(50):
                call(h, Zh, outh)
            us_host = (time.perf_counter() - t0) / 50 * 1e6
            with rd.RegisteredArray(Zh), rd.RegisteredArray(outh):            # the same arrays page-locked in place (rdb_host_register)
                for _ in range(3):
                    call(h, Zh, outh)
                t0 = time.perf_counter()
                for _ in range(50):
                    call(h, Zh, outh)
                us_host_reg = (time.perf_counter() - t0) / 50 * 1e6
            us_plan = us_plan_sync = float("nan")
            if name in PLAN_OPS:
                op = PLAN_OPS[name]
                kw = dict(out=out) if op == rd._abi.OP_DISCRETE_DYNAMICS else dict(J=out)
                plan = rd._abi.Plan(h, op, rd.RK4.code, Z, 0.01, **kw)
                for _ in range(10):
                    plan.launch()
                torch.cuda.synchronize()
                e0.record()
                for _ in range(args.reps):
                    plan.launch()
                e1.record(); torch.cuda.synchronize()
                us_plan = e0.elapsed_time(e1) / args.reps * 1e3
                t0 = time.perf_counter()
                for _ in range(args.reps):
                    plan.launch(); torch.cuda.synchronize()
                us_plan_sync = (time.perf_counter() - t0) / args.reps * 1e6
            lines.append(f"| {name} | {N} | {us_stream:.1f} | {us_sync:.1f} | {us_graph:.1f} | {us_host:.1f} | {us_host_reg:.1f} | {us_plan:.1f} | {us_plan_sync:.1f} |")
            print(lines[-1], flush=True)
    if args.out:
        with open(args.out, "w") as f:
            f.write("# Small-batch latency (B200, device pointers unless noted; scripts/small_batch.py)\n\n" + "\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
