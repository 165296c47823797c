"""Launches the error-state Jacobian kernel or the warp-cooperative ImplicitMidpoint kernel a few times (target of ncu captures).

    python scripts/prof_extra.py err|implicit|soa|body|body64
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import torch
import rdb200 as rd
import gpu_quick as g

if __name__ == "__main__":
    what = sys.argv[1]
    qd = rd.Quadrotor()
    N = 262144
    Z = torch.from_numpy(g.rand_inputs(qd._h, N, np.random.default_rng(2)).astype(np.float32)).cuda()
    if what == "soa":                    # C2 in the component-major layout (2-D tensor-map loads and stores)
        cp = rd.Cartpole()
        Nc = 1 << 20
        Zc = torch.rand((5, Nc), dtype=torch.float64, device="cuda")
        Jc = torch.empty((20, Nc), dtype=torch.float64, device="cuda")
        for _ in range(8):
            cp._h.discrete_jacobian(3, Zc, 0.01, J=Jc, layout=rd.SOA)
    elif what == "err":
        J = torch.empty((N, 16, 12), dtype=torch.float32, device="cuda")
        for _ in range(8):
            qd._h.discrete_error_jacobian(3, Z, 0.01, J=J)
    elif what in ("body", "body64"):   # body-frame quadrotor RK4 Jacobian (split force), fp32 / fp64: the issue- / FP64-latency-bound kernels
        qb = rd.Quadrotor(bodyframe=True)
        Zb = Z if what == "body" else Z.double()
        J = torch.empty((N, 17, 13), dtype=Zb.dtype, device="cuda")
        for _ in range(8):
            qb._h.discrete_jacobian(3, Zb, 0.01, J=J)
    else:
        J = torch.empty((N, 17, 13), dtype=torch.float32, device="cuda")
        for _ in range(4):
            qd._h.discrete_jacobian(4, Z, 0.01, J=J)
    torch.cuda.synchronize()
