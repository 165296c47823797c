"""fp32 ImplicitMidpoint error against the CPU checker, per model (development aid for the 1e-4 bound)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rdb200 as rd
from oracle import rd_oracle as o
from common import rand_inputs, zoo
for name in ["cartpole", "quad_quat_world", "quad_mrp_world", "quad_quat_body", "body_mrp_body", "satellite_mrp", "body_quat_world", "di3"]:
    om, gm = zoo()[name][0](), zoo()[name][1](rd)
    N = 4000
    Z = rand_inputs(om.n, om.m, N, np.random.default_rng(111))
    dt = np.random.default_rng(112).uniform(0.005, 0.1, N)
    Z32 = Z.astype(np.float32)
    ref = o.discrete_jacobian(om, o.IMPLICIT_MIDPOINT, Z32.astype(np.float64), dt)
    xr = o.discrete_dynamics(om, o.IMPLICIT_MIDPOINT, Z32.astype(np.float64), dt)
    xn = np.empty((N, om.n), dtype=np.float32)
    J32 = gm._h.discrete_jacobian(rd._abi.IMPLICIT_MIDPOINT, Z32, dt, xn=xn)
    e = np.abs(J32 - ref)
    k = np.unravel_index(np.argmax(e), e.shape)
    print(f"{name:18s} max|dJ|={e.max():.3e} at {k} (|Jref|={abs(ref[k]):.3f}, dt={dt[k[0]]:.3f})  p99.9={np.quantile(e.max(axis=(1,2)), 0.999):.3e}  max|dx|={np.abs(xn - xr).max():.3e}", flush=True)
