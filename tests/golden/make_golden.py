"""Generates tests/golden/*.npz — seeded input/output vectors for the BASELINE.json configurations.

The reference (Julia) cannot run in this image and ships no golden vectors, so these come from
the CPU oracle (oracle/rd_oracle.cpp, forward-mode path == the reference's default ForwardAD path) and are accepted
only if the oracle's second, independent path (integrator chain rule, reference: src/integration.jl:302-337) and the
independent numpy complex-step restatement (oracle/independent.py) agree with them to 1e-11.
Parity status therefore stays "unpinned against Julia output"; the vectors pin the oracle and the GPU path to each
other and across commits.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import rd_oracle as o          # noqa: E402
from oracle import independent as ind      # noqa: E402

QN = {o.EULER: "euler", o.RK2: "rk2", o.RK3: "rk3", o.RK4: "rk4"}


def rigid_inputs(rng, N, nrot, m):
    """rand(model) for a rigid body: r, v, w ~ U[0,1)^3, unit quaternion (or its MRP), u ~ U[0,1)^m."""
    q = rng.standard_normal((N, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    att = q if nrot == 4 else q[:, 1:] / (1.0 + q[:, :1])
    return np.concatenate([rng.random((N, 3)), att, rng.random((N, 6)), rng.random((N, m))], axis=1)


def verify(name, model, imodel, Q, Z, dt, J):
    Jc = o.discrete_jacobian(model, Q, Z, dt, method=o.CHAIN)
    assert np.abs(J - Jc).max() < 1e-11, name
    dts = np.broadcast_to(dt, (Z.shape[0],))
    for k in range(0, Z.shape[0], max(1, Z.shape[0] // 16)):
        Ji = ind.discrete_jacobian(imodel, QN[Q], Z[k], dts[k])
        assert np.abs(o.as_matrix(J)[k] - Ji).max() < 1e-11, (name, k)


def main():
    out = {}
    # C1: Cartpole RK3, N=1024, seed 1, x,u ~ U[0,1), dt = 0.01 (test/integration_tests.jl:21-25)
    rng = np.random.default_rng(1)
    Z = rng.random((1024, 5)); m = o.cartpole()
    J = o.discrete_jacobian(m, o.RK3, Z, 0.01); verify("c1", m, ind.Cartpole(), o.RK3, Z, 0.01, J)
    out["c1_cartpole_rk3"] = dict(Z=Z, dt=0.01, J=J, xn=o.discrete_dynamics(m, o.RK3, Z, 0.01))
    # C2 (sample): Cartpole RK4 fp64, seed 2
    rng = np.random.default_rng(2)
    Z = rng.random((512, 5))
    J = o.discrete_jacobian(m, o.RK4, Z, 0.01); verify("c2", m, ind.Cartpole(), o.RK4, Z, 0.01, J)
    out["c2_cartpole_rk4"] = dict(Z=Z, dt=0.01, J=J, xn=o.discrete_dynamics(m, o.RK4, Z, 0.01))
    # C3 (sample): Quadrotor{QuatRotation} RK4 on fp32-rounded inputs, seed 3, plus the LieState maps
    rng = np.random.default_rng(3)
    Z = rigid_inputs(rng, 256, 4, 4).astype(np.float32).astype(np.float64); m = o.quadrotor()
    J = o.discrete_jacobian(m, o.RK4, Z, 0.01); verify("c3", m, ind.RigidBody("quadrotor", mass=0.5, J=np.diag([.0023, .0023, .004])), o.RK4, Z, 0.01, J)
    X0 = rigid_inputs(rng, 256, 4, 4)[:, :13]
    out["c3_quadrotor_rk4"] = dict(Z=Z, dt=0.01, J=J, xn=o.discrete_dynamics(m, o.RK4, Z, 0.01), G=o.errstate_jacobian(m, Z[:, :13]),
                                   X0=X0, dX=o.state_diff(m, Z[:, :13], X0), H=o.grad_errstate_jacobian(m, Z[:, :13], X0))
    # C4 (sample): Satellite RigidBody{MRP} RK2, dt = 0.1 (examples/single_satellite.jl:39), seed 4
    rng = np.random.default_rng(4)
    Z = rigid_inputs(rng, 256, 3, 6); m = o.satellite(o.ROT_MRP)
    J = o.discrete_jacobian(m, o.RK2, Z, 0.1); verify("c4", m, ind.RigidBody("body", rot="mrp", mass=1.0, J=np.eye(3)), o.RK2, Z, 0.1, J)
    out["c4_satellite_mrp_rk2"] = dict(Z=Z, dt=0.1, J=J, xn=o.discrete_dynamics(m, o.RK2, Z, 0.1), G=o.errstate_jacobian(m, Z[:, :12]))
    # C5 (sample): mixed sweep, 8 trajectories x 32 knots, per-trajectory dt = 0.01 (1 + traj mod 4), last knot terminal
    rng = np.random.default_rng(5)
    ntraj, K = 8, 32
    dt = np.repeat(0.01 * (1 + np.arange(ntraj) % 4), K).reshape(ntraj, K); dt[:, -1] = 0.0
    Zc = rng.random((ntraj // 2 * K, 5)); Zq = rigid_inputs(rng, ntraj // 2 * K, 4, 4)
    mc, mq = o.cartpole(), o.quadrotor()
    dtc, dtq = dt[:ntraj // 2].reshape(-1), dt[ntraj // 2:].reshape(-1)
    Jc = o.discrete_jacobian(mc, o.RK4, Zc, dtc); verify("c5c", mc, ind.Cartpole(), o.RK4, Zc, dtc, Jc)
    Jq = o.discrete_jacobian(mq, o.RK4, Zq, dtq)
    out["c5_mixed_sweep"] = dict(Zc=Zc, dtc=dtc, Jc=Jc, Zq=Zq, dtq=dtq, Jq=Jq)
    # rollout: 16 Cartpole trajectories x 24 knots (src/trajectories.jl:436-441)
    rng = np.random.default_rng(6)
    x0 = rng.random((16, 4)); U = rng.random((16, 23, 1))
    out["rollout_cartpole_rk4"] = dict(x0=x0, U=U, dt=0.02, X=o.rollout(mc, o.RK4, x0, U, 0.02))
    for name, d in out.items():
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(name, {k: np.shape(v) for k, v in d.items()})


if __name__ == "__main__":
    main()
