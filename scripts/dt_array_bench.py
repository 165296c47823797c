"""Cost of per-knot dt arrays (KnotPoint.dt as an array: trajectories, the mixed sweep) against a scalar step: back-to-back plan launches
over rotating buffer sets, CUDA events (development aid)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import rdb200 as rd
import bench

def run(name, mk, dtn, N, nsets=3, steps=40):
    h = mk()._h
    n, m = h.n, h.m
    Zs = [torch.from_numpy(bench.make_inputs(n, m, N, dtn, i)).cuda() for i in range(nsets)]
    Js = [torch.empty((N, n + m, n), dtype=Zs[0].dtype, device="cuda") for _ in range(nsets)]
    dta = torch.full((N,), 0.01, dtype=torch.float64, device="cuda")
    out = {}
    for label, dt in (("scalar dt", 0.01), ("dt array", dta)):
        plans = [rd._abi.Plan(h, rd._abi.OP_DISCRETE_JACOBIAN, rd.RK4.code, Z, dt, J=J) for Z, J in zip(Zs, Js)]
        for i in range(5): plans[i % nsets].launch()
        torch.cuda.synchronize()
        ts = []
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(2000000); e0.record()
            for i in range(steps): plans[i % nsets].launch()
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / steps * 1e3)
        out[label] = min(ts)
    print(f"{name} N={N}: scalar dt {out['scalar dt']:.1f} us, dt array {out['dt array']:.1f} us", flush=True)

if __name__ == "__main__":
    run("quadrotor fp32", rd.Quadrotor, "float32", 1 << 20)
    run("quadrotor fp32", rd.Quadrotor, "float32", 1 << 19)
    run("cartpole fp64", rd.Cartpole, "float64", 1 << 20)
    run("cartpole fp64", rd.Cartpole, "float64", 1 << 19)
