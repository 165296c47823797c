"""Forward pass + linearisation of a batch of trajectories (BASELINE configs[4] shape: 4096 x 256):
rollout alone, linearisation alone, the two back to back on one stream, and the chunk-pipelined two-stream form
(rdb_trajectory_rollout_linearize) for several chunk counts.  CUDA events, device-resident, median of `reps`."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import rdb200 as rd
import bench

def med(fn, reps=20):
    ts = []
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts))

ntraj, K = 4096, 256
print("| model | rollout us | linearize us | back to back us | pipelined us (chunks: 2 / 4 / 8 / 16 / 32) | error-state pipelined (8) us |")
print("|---|---|---|---|---|---|")
for name, mk, dtn, err in (("Quadrotor fp32", rd.Quadrotor, "float32", True), ("Cartpole fp64", rd.Cartpole, "float64", False), ("Quadrotor fp64", rd.Quadrotor, "float64", True)):
    dm = rd.DiscretizedDynamics(mk(), rd.RK4)
    h = dm._h
    tb = rd.TrajectoryBatch(dm, ntraj, K, np.dtype(dtn))
    Z0 = bench.make_inputs(h.n, h.m, ntraj, dtn, 3)
    tb.set_initial_state(np.ascontiguousarray(Z0[:, :h.n]))
    tb.set_controls((0.5 * np.random.default_rng(4).random((K - 1, ntraj, h.m)) + (1.0 if h.m == 4 else 0.0)).astype(dtn))
    tb.set_timesteps(0.01)
    J = torch.empty((K, ntraj, h.n + h.m, h.n), dtype=getattr(torch, dtn), device="cuda")
    Jb = torch.empty((K, ntraj, h.nerr + h.m, h.nerr), dtype=getattr(torch, dtn), device="cuda")
    t_roll = med(lambda: tb.rollout())
    t_lin = med(lambda: tb.linearize(J=J))
    t_b2b = med(lambda: (tb.rollout(), tb.linearize(J=J)))
    pipes = [med(lambda c=c: tb.rollout_linearize(J=J, chunks=c)) for c in (2, 4, 8, 16, 32)]
    t_err = med(lambda: tb.rollout_linearize(error_state=True, J=Jb, chunks=8)) if err else float("nan")
    X = tb.states()
    assert np.isfinite(X).all()
    print(f"| {name} | {t_roll:.1f} | {t_lin:.1f} | {t_b2b:.1f} | {' / '.join(f'{p:.1f}' for p in pipes)} | {t_err:.1f} |", flush=True)
