"""Shared helpers for the test-suite: seeded inputs mirroring rand(model) and the model zoo in both worlds."""
import os

import numpy as np

from oracle import rd_oracle as o

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = {np.float64: 1e-10, np.float32: 1e-4}      # north_star: 1e-10 (fp64) / 1e-4 (fp32), test/integration_tests.jl:7-18


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def rand_inputs(n, m, N, rng):
    """rand(model): U[0,1) everywhere; rigid bodies get a unit quaternion / its MRP (reference: src/rigidbody.jl:50-59)."""
    Z = rng.random((N, n + m))
    if n >= 12:
        q = rng.standard_normal((N, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
        if n == 13:
            Z[:, 3:7] = q
        else:
            Z[:, 3:6] = q[:, 1:] / (1.0 + np.abs(q[:, :1]))
    return Z


# name -> (oracle model factory, rdb200 model factory taking the rdb200 module)
ROTS = {"quat": (o.ROT_QUAT, "QuatRotation"), "mrp": (o.ROT_MRP, "MRP"), "rp": (o.ROT_RP, "RodriguesParam")}


def zoo():
    z = {"cartpole": (lambda: o.cartpole(), lambda rd: rd.Cartpole())}
    for D in (1, 2, 3):
        z[f"di{D}"] = (lambda D=D: o.double_integrator(D), lambda rd, D=D: rd.DoubleIntegrator(D))
    for rn, (rc, rcls) in ROTS.items():
        for fn, fc in (("world", o.WORLD), ("body", o.BODYFRAME)):
            z[f"quad_{rn}_{fn}"] = (lambda rc=rc, fc=fc: o.quadrotor(rc, fc),
                                    lambda rd, rcls=rcls, fc=fc: rd.Quadrotor(getattr(rd, rcls), bodyframe=bool(fc)))
            z[f"body_{rn}_{fn}"] = (lambda rc=rc, fc=fc: o.body(rc, fc),
                                    lambda rd, rcls=rcls, fc=fc: rd.Body(getattr(rd, rcls), bodyframe=bool(fc)))
    z["satellite_mrp"] = (lambda: o.satellite(o.ROT_MRP), lambda rd: rd.Satellite(rd.MRP))
    return z
