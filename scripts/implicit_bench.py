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
