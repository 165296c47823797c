"""Component-major ("SoA") layout timings next to the knot-major ones (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import gpu_quick as g
import rdb200 as rd

if __name__ == "__main__":
    cp, qd, sat = rd.Cartpole(), rd.Quadrotor(), rd.Satellite(rd.MRP)
    g.bench_soa("cartpole", cp, 3, np.float64, 1 << 20)
    g.bench_soa("quadrotor", qd, 3, np.float32, 262144)
    g.bench_soa("satellite rk2", sat, 1, np.float64, 1 << 20)
    g.bench("cartpole", cp, 3, np.float64, 1 << 20)
    g.bench("quadrotor", qd, 3, np.float32, 262144)
