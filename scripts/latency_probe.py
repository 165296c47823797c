"""Kernel latency at small N (what a solver calls per iteration): CUDA-graph replay of 20 serialised discrete_jacobian! launches,
time per launch.  Development aid for scripts/tune.py variants (RDB200_LIB selects the library).

    python scripts/latency_probe.py [quadrotor|cartpole] [float32|float64]
"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import rdb200 as rd
import bench
from oracle import rd_oracle as o

wl = sys.argv[1] if len(sys.argv) > 1 else "quadrotor"
dtn = sys.argv[2] if len(sys.argv) > 2 else "float32"
mk, Q = bench.gpu_model(wl, rd)
omk, oQ = bench.oracle_model(wl)
h = mk()._h
n, m = h.n, h.m
out = {}
for N in (64, 256, 1024, 4096, 16384):
    Z = torch.from_numpy(bench.make_inputs(n, m, N, dtn, 3)).cuda()
    J = torch.empty((N, n + m, n), dtype=Z.dtype, device="cuda")
    plan = rd._abi.Plan(h, rd._abi.OP_DISCRETE_JACOBIAN, Q.code, Z, 0.01, J=J)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(3):
            plan.launch(st.cuda_stream)
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(20):
                plan.launch(st.cuda_stream)
        g.replay(); st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(10):
            g.replay()
        e1.record(st); st.synchronize()
    us = e0.elapsed_time(e1) / 200 * 1e3
    ref = o.discrete_jacobian(omk(), oQ, Z.cpu().numpy().astype(np.float64)[:64], 0.01)
    err = float(np.abs(J.cpu().numpy()[:64] - ref).max())
    out[N] = round(us, 2)
    assert err < (1e-4 if dtn == "float32" else 1e-10), err
print(json.dumps(out))
