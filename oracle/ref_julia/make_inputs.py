"""Seeded INPUT vectors for the Julia reference run (oracle/ref_julia/gen_fixtures.jl) — committed under tests/golden/julia_in/
as plain .npy files (Julia reads them without any package).  Deterministic: tests/test_julia_fixtures.py regenerates them and
checks the committed files did not drift.

    python oracle/ref_julia/make_inputs.py

Layout: Z (N, n+m) C-order == Julia Matrix{Float64}(n+m, N): column k is getdata(z_k) (reference: src/knotpoint.jl:196).
"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(os.path.dirname(HERE)), "tests", "golden", "julia_in")

# case -> (quadrature rule name, dt); shared verbatim with gen_fixtures.jl and tests/test_julia_fixtures.py
CASES = {
    "c1_cartpole": ("RK3", 0.01), "c2_cartpole": ("RK4", 0.01), "c3_quadrotor": ("RK4", 0.01), "c3_quadrotor_dt01": ("RK4", 0.1),
    "c4_satellite_mrp": ("RK2", 0.1),
}
ROTS = ("quat", "mrp", "rp")
FRAMES = ("world", "body")


def unit_quats(rng, N):
    q = rng.standard_normal((N, 4))
    return q / np.linalg.norm(q, axis=1, keepdims=True)


def rigid(rng, N, rot, m, q=None):
    """rand(model) of a RigidBody{R}: r, v, w ~ U[0,1)^3, attitude from a random unit quaternion, u ~ U[0,1)^m
    (reference: src/rigidbody.jl:50-59, src/liestate.jl:175-205)."""
    q = unit_quats(rng, N) if q is None else q
    if rot == "quat":
        att = q
    else:                                   # a 3-parameter attitude of norm < 1 (MRP of the quaternion with w >= 0; also used as the
        att = q[:, 1:] / (1.0 + np.abs(q[:, :1]))    # Rodrigues vector: any 3-vector is a valid one, and this keeps it well-conditioned)
    return np.concatenate([rng.random((N, 3)), att, rng.random((N, 6)), rng.random((N, m))], axis=1)


def inputs():
    out = {}
    N = 32
    out["c1_cartpole_Z"] = np.random.default_rng(1).random((N, 5))               # seed 1, U[0,1)
    out["c2_cartpole_Z"] = np.random.default_rng(2).random((N, 5))
    out["c3_quadrotor_Z"] = rigid(np.random.default_rng(3), N, "quat", 4)
    out["c3_quadrotor_dt01_Z"] = out["c3_quadrotor_Z"]
    out["c4_satellite_mrp_Z"] = rigid(np.random.default_rng(4), N, "mrp", 6)
    # off-manifold quaternions: |q| in [0.8, 1.2] (the state quaternion is never renormalised inside dynamics, src/rigidbody.jl:101-105)
    rng = np.random.default_rng(21)
    Z = rigid(rng, N, "quat", 4)
    Z[:, 3:7] *= rng.uniform(0.8, 1.2, (N, 1))
    out["quad_offmanifold_Z"] = Z
    # thrust clamp max(0, kf w): exact ties and negative controls (test/quadrotor.jl:67-70)
    Z = rigid(np.random.default_rng(22), 16, "quat", 4)
    pat = np.array([[0.0, 0.3, -0.2, 0.7], [-0.5, 0.0, 0.0, 0.9], [0.0, 0.0, 0.0, 0.0], [-0.1, -0.2, 0.4, 0.0]])
    Z[:, 13:] = pat[np.arange(16) % 4]
    out["quad_tie_Z"] = Z
    # every RigidBody{R} x velocity frame: Quadrotor (m = 4) and Body (m = 6, mass 2, J = diag(2,3,1), test/rigidbody_test.jl:23-56)
    for i, rot in enumerate(ROTS):
        out[f"quad_{rot}_Z"] = rigid(np.random.default_rng(30 + i), 16, rot, 4)
        out[f"body_{rot}_Z"] = rigid(np.random.default_rng(40 + i), 16, rot, 6)
        # LieState maps: a second state per knot (x0 for state_diff, xbar for the derivative of the error-state Jacobian)
        out[f"lie_{rot}_X0"] = rigid(np.random.default_rng(50 + i), 16, rot, 4)[:, :(13 if rot == "quat" else 12)]
    out["implicit_cartpole_Z"] = np.random.default_rng(60).random((16, 5))
    out["implicit_quadrotor_Z"] = rigid(np.random.default_rng(61), 16, "quat", 4)
    return out


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    for name, a in inputs().items():
        np.save(os.path.join(OUT, name + ".npy"), np.ascontiguousarray(a, dtype=np.float64))
    print(f"wrote {len(inputs())} arrays to {OUT}")
