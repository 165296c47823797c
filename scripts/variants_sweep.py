"""discrete_jacobian! RK4 for every RigidBody{R} x velocity frame x dtype (and the small models): back-to-back plan launches over
rotating buffer sets larger than L2, CUDA events, launch gate as in bench.py; fraction of the measured HBM peak.

    python scripts/variants_sweep.py [N] > profiles/variants_rNN.md

Under ncu (per-launch instruction counts and pipe utilisation of the same kernels; scripts/variants_roofline.py merges both):
    RDB_SWEEP_STEPS=1 RDB_SWEEP_WARM=1 ncu --metrics ... -k regex:knot_kernel --csv --log-file variants_ncu.csv python scripts/variants_sweep.py
"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import rdb200 as rd
import bench
from common import rand_inputs, zoo
from oracle import rd_oracle as o

N = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
peak, _ = bench.hbm_peak()
print(f"| model | dtype | rule | N | us | evals/s | algorithmic GB/s | of {peak:.0f} GB/s | max err vs checker |")
print("|---|---|---|---|---|---|---|---|---|")
names = [k for k in sorted(zoo()) if k.startswith(("quad", "body"))] + ["cartpole", "di3"]
for name in names:
    for dtn in ("float32", "float64"):
        for Q, qn in ((rd.RK4, "RK4"),) + (((rd.RK2, "RK2"), (rd.RK3, "RK3")) if name in ("quad_quat_world", "body_mrp_world") else ()):
            om, gm = zoo()[name][0](), zoo()[name][1](rd)
            h = gm._h
            n, m = h.n, h.m
            Nn = N * (4 if n < 12 else 1)
            es = np.dtype(dtn).itemsize
            per = es * ((n + m) + n * (n + m))
            nsets = max(2, int(np.ceil(600e6 / (Nn * per))) + 1)
            Zs = [torch.from_numpy(rand_inputs(n, m, Nn, np.random.default_rng(i)).astype(dtn)).cuda() for i in range(nsets)]
            Js = [torch.empty((Nn, n + m, n), dtype=Zs[0].dtype, device="cuda") for _ in range(nsets)]
            plans = [rd._abi.Plan(h, rd._abi.OP_DISCRETE_JACOBIAN, Q.code, Z, 0.01, J=J) for Z, J in zip(Zs, Js)]
            for i in range(int(os.environ.get("RDB_SWEEP_WARM", "5"))):
                plans[i % nsets].launch()
            torch.cuda.synchronize()
            steps = int(os.environ.get("RDB_SWEEP_STEPS", "50"))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(2_000_000)
            e0.record()
            for i in range(steps):
                plans[i % nsets].launch()
            e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / steps * 1e3
            idx = np.arange(0, Nn, 997)
            ref = o.discrete_jacobian(om, Q.code, Zs[0].cpu().numpy()[idx].astype(np.float64), 0.01)
            err = float(np.abs(Js[0].cpu().numpy()[idx] - ref).max())
            gbs = Nn * per / (us * 1e-6) / 1e9
            print(f"| {name} | {dtn} | {qn} | {Nn} | {us:.1f} | {Nn / (us * 1e-6):.3e} | {gbs:.0f} | {gbs / peak:.2f} | {err:.1e} |", flush=True)
            del plans, Zs, Js
            torch.cuda.empty_cache()
