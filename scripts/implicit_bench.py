"""ImplicitMidpoint timings (development aid): python scripts/implicit_bench.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import gpu_quick as g
import rdb200 as rd

if __name__ == "__main__":
    qd = rd.Quadrotor()
    g.bench("cartpole implicit-midpoint", rd.Cartpole(), 4, np.float64, 1 << 20)
    g.bench("quadrotor implicit-midpoint", qd, 4, np.float32, 262144)
    g.bench("quadrotor implicit-midpoint", qd, 4, np.float64, 262144)
    g.bench("quadrotor{MRP} implicit-midpoint", rd.Quadrotor(rd.MRP), 4, np.float32, 262144)
    g.bench("satellite{MRP} implicit-midpoint", rd.Satellite(rd.MRP), 4, np.float64, 262144, dt=0.1)
    # the same quadrotor as a USER rigid body (NVRTC): conforming wrench -> block-triangular kernel as well
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    from test_abi_host import QUAD_WRENCH
    uq = rd.CustomRigidBody(rd.QuatRotation, 4, QUAD_WRENCH, mass=0.5, J=(0.0023, 0.0023, 0.004), params=[1.0, 0.0245, 0.175, 0.0, 0.0, -9.81])
    g.bench("user-wrench quadrotor implicit-midpoint", uq, 4, np.float32, 262144)
