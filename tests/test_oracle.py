"""Pins the CPU oracle (test infrastructure) — runs without a GPU.

The reference ships no golden vectors and Julia is unavailable, so the oracle is pinned by
 (1) the relational assertions the reference's own tests make on this path, re-run here on the oracle;
 (2) an independent numpy complex-step restatement (oracle/independent.py) that shares no code with it;
 (3) the hand-verified known answers of SURVEY.md Appendix C;
 (4) the committed fixtures in tests/golden/ (regression pin);
 (5) sympy / 50-digit mpmath restatements (oracle/highprec.py): the symbolic Cartpole Jacobian, the reference's own hand-derived
     analytic Jacobian (test/cartpole_model.jl:57-96), and high-precision differences of the composed RK maps;
 (6) scipy.spatial.transform.Rotation — an independent third-party implementation — for everything the path takes from Rotations.jl on
     the manifold: q*v, q\\F, kinematics, ∇differential, ∇²differential, the Cayley error (all three attitude parameterisations).
"""
import numpy as np
import pytest

from oracle import rd_oracle as o
from oracle import independent as ind
from common import golden, rand_inputs, zoo

QS = {o.EULER: "euler", o.RK2: "rk2", o.RK3: "rk3", o.RK4: "rk4"}


def ind_model(name):
    if name == "cartpole":
        return ind.Cartpole()
    if name.startswith("di"):
        return ind.DoubleIntegrator(int(name[2]))
    if name == "satellite_mrp":
        return ind.RigidBody("body", rot="mrp", mass=1.0, J=np.eye(3))
    kind, rot, frame = name.split("_")
    if kind == "quad":
        return ind.RigidBody("quadrotor", rot=rot, frame=frame, mass=0.5, J=np.diag([0.0023, 0.0023, 0.004]))
    return ind.RigidBody("body", rot=rot, frame=frame)


# ---- (3) known answers, SURVEY.md Appendix C ------------------------------------------------------------------------
def test_cartpole_known_answers():
    m = o.cartpole()
    z = np.array([[0.1, 0.2, 0.3, 0.4, 0.5]])
    assert np.allclose(o.dynamics(m, z)[0], [0.3, 0.4, 0.8782651651833607, -5.619408939955959], atol=1e-14)
    xn = {o.EULER: [0.103, 0.204, 0.3087826516518336, 0.3438059106004404],
          o.RK2: [0.10304391325825918, 0.20371902955300222, 0.30881310601518436, 0.34336875420270035],
          o.RK3: [0.10304401477280367, 0.20371757236500976, 0.308811491738707, 0.34338962242102583],
          o.RK4: [0.10304401065773387, 0.2037176246055777, 0.3088114828472635, 0.3433897139235887]}
    for Q, ref in xn.items():
        assert np.allclose(o.discrete_dynamics(m, Q, z, 0.01)[0], ref, atol=1e-14)
    J = o.as_matrix(o.discrete_jacobian(m, o.RK4, z, 0.01))[0]
    Jref = np.array([[1, 8.6756343027589333e-05, 9.9999999999999985e-03, 1.0448126477784444e-06, 4.9597243518762175e-05],
                     [0, 9.9888657059330599e-01, 0, 9.9948082258736765e-03, -9.7176440940498530e-05],
                     [0, 1.7323086579775784e-02, 1, 2.3441048181776361e-04, 9.9182435944503398e-03],
                     [0, -2.2255256446242244e-01, 0, 9.9859778390010401e-01, -1.9427409451867484e-02]])
    assert np.abs(J - Jref).max() < 1e-13
    inv = {o.EULER: (2.012474668101985, 3.804810437010853, 3.999690905289077),
           o.RK2: (2.011344109181849, 3.802943360695978, 3.997482116503141),
           o.RK3: (2.011337012474398, 3.803014784739832, 3.997483940351493),
           o.RK4: (2.011337220498831, 3.803015151419291, 3.997484354493410)}
    for Q, (fro, tot, tr) in inv.items():
        J = o.as_matrix(o.discrete_jacobian(m, Q, z, 0.01))[0]
        assert np.allclose([np.linalg.norm(J), J.sum(), np.trace(J[:, :4])], [fro, tot, tr], atol=1e-12)


def test_quadrotor_known_answers():
    m = o.quadrotor()
    z = np.array([[.1, .2, .3, .8, .2, -.4, .4, .4, -.5, .6, .7, .8, -.9, 1, 1.5, 1.25, 1.75]])
    xd = [0.4, -0.5, 0.6, 0.27, 0.3, 0.55, -0.14, -5.28, -7.04, -3.21, -18.489565217391302, 18.55608695652174, -6.125]
    assert np.allclose(o.dynamics(m, z)[0], xd, atol=1e-13)
    xn = [0.1037375220479742, 0.19464802561115327, 0.30584073621895413, 0.8030187350157577, 0.20250115547279274,
          -0.394271190376744, 0.3983772355640589, 0.3476534470164253, -0.5703831897018942, 0.5682771583095719,
          0.5159317379858185, 0.9860433879861684, -0.96125]
    assert np.allclose(o.discrete_dynamics(m, o.RK4, z, 0.01)[0], xn, atol=1e-13)
    J = o.as_matrix(o.discrete_jacobian(m, o.RK4, z, 0.01))[0]
    assert np.allclose([np.linalg.norm(J), J.sum(), np.trace(J[:, :13])], [3.934201769388178, 13.31063583261720, 12.99985104390944], atol=1e-11)
    assert np.allclose(J[12, 13:], [0.06125, -0.06125, 0.06125, -0.06125], atol=1e-14)      # km dt / J33
    G = o.errstate_jacobian(m, z[:, :13])[0].T
    assert np.allclose(G[3:7, 3:6], [[-.2, .4, -.4], [.8, -.4, -.4], [.4, .8, -.2], [.4, .2, .8]], atol=1e-14)


def test_satellite_mrp_known_answers():
    m = o.satellite(o.ROT_MRP)
    z = np.array([[.1, .2, .3, 1 / 9, -2 / 9, 2 / 9, .4, -.5, .6, .7, .8, -.9, .1, .2, .3, .4, .5, .6]])
    xd = [0.4, -0.5, 0.6, 0.15, 0.3388888888888889, -0.1111111111111111, -0.268, -0.024, 0.26, 0.4, 0.5, 0.6]
    assert np.allclose(o.dynamics(m, z)[0], xd, atol=1e-14)
    xn = [0.13866, 0.14988, 0.3613, 0.1256306537615741, -0.18796870972222224, 0.2120468998842593, 0.37508276490760817,
          -0.5024963801754104, 0.6278010697882347, 0.74, 0.85, -0.84]
    assert np.allclose(o.discrete_dynamics(m, o.RK2, z, 0.1)[0], xn, atol=1e-13)
    J = o.as_matrix(o.discrete_jacobian(m, o.RK2, z, 0.1))[0]
    assert np.allclose([np.linalg.norm(J), J.sum(), np.trace(J[:, :12])], [3.471030047880510, 12.80119860912479, 11.95425437702546], atol=1e-11)


# ---- (1) the reference's relational assertions ----------------------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(zoo()))
def test_forward_ad_equals_chain_rule(name):
    """test/integration_tests.jl:7-18: all sig x diff paths agree to atol 1e-10."""
    m = zoo()[name][0]()
    Z = rand_inputs(m.n, m.m, 64, np.random.default_rng(7))
    for Q in QS:
        for dt in (0.01, 0.1):
            a = o.discrete_jacobian(m, Q, Z, dt, method=o.AD)
            b = o.discrete_jacobian(m, Q, Z, dt, method=o.CHAIN)
            assert np.abs(a - b).max() < 1e-11
    assert np.abs(o.jacobian(m, Z, method=o.AD) - o.jacobian(m, Z, method=o.CHAIN)).max() < 1e-11   # :27-33


def test_rk_stage_formulas():
    """test/integration_tests.jl:35-38,47-59,80-99 — steps equal the hand-written stage formulas."""
    for name in ("cartpole", "quad_quat_world", "body_mrp_body"):
        m = zoo()[name][0]()
        Z = rand_inputs(m.n, m.m, 16, np.random.default_rng(8))
        x, u, h = Z[:, :m.n], Z[:, m.n:], 0.05
        f = lambda xx: o.dynamics(m, np.concatenate([xx, u], axis=1))
        k1 = f(x) * h
        assert np.abs(o.discrete_dynamics(m, o.EULER, Z, h) - (x + k1)).max() < 1e-14
        k2 = f(x + k1 / 2) * h
        assert np.abs(o.discrete_dynamics(m, o.RK2, Z, h) - (x + k2)).max() < 1e-14
        k3 = f(x - k1 + 2 * k2) * h
        assert np.abs(o.discrete_dynamics(m, o.RK3, Z, h) - (x + (k1 + 4 * k2 + k3) / 6)).max() < 1e-14
        k3 = f(x + k2 / 2) * h
        k4 = f(x + k3) * h
        assert np.abs(o.discrete_dynamics(m, o.RK4, Z, h) - (x + (k1 + 2 * k2 + 2 * k3 + k4) / 6)).max() < 1e-14


def test_rk2_linear_known_answer():
    """test/old_tests/linear_tests.jl:135-145 (disabled v0.3 test): A = I + A_(I + A_ dt/2) dt, B = (A_ B_ dt/2 + B_) dt."""
    D, dt = 2, 0.1
    m = o.double_integrator(D)
    A_ = np.block([[np.zeros((D, D)), np.eye(D)], [np.zeros((D, 2 * D))]])
    B_ = np.vstack([np.zeros((D, D)), np.eye(D)])
    J = o.as_matrix(o.discrete_jacobian(m, o.RK2, np.random.default_rng(0).random((1, 3 * D)), dt))[0]
    assert np.allclose(J[:, :2 * D], np.eye(2 * D) + A_ @ (np.eye(2 * D) + A_ * dt / 2) * dt, atol=1e-15)
    assert np.allclose(J[:, 2 * D:], (A_ @ B_ * dt / 2 + B_) * dt, atol=1e-15)


def test_rigid_body_first_principles():
    """test/rigidbody_test.jl:135-162: xdot re-derived from Newton-Euler with rotation matrices."""
    rng = np.random.default_rng(9)
    for frame in (o.WORLD, o.BODYFRAME):
        m = o.body(o.ROT_QUAT, frame)
        Z = rand_inputs(13, 6, 8, rng)
        xd = o.dynamics(m, Z)
        for k in range(8):
            r, q, v, w, u = Z[k, :3], Z[k, 3:7], Z[k, 7:10], Z[k, 10:13], Z[k, 13:]
            R = ind.quat_matrix(q)           # unit q -> rotation matrix
            assert np.allclose(R @ R.T, np.eye(3), atol=1e-12)
            F, tau, J = R @ u[:3], u[3:], np.diag([2.0, 3.0, 1.0])
            assert np.allclose(xd[k, 3:7], 0.5 * ind.lmult(q) @ np.r_[0, w], atol=1e-13)
            assert np.allclose(xd[k, 10:], np.linalg.solve(J, tau - np.cross(w, J @ w)), atol=1e-13)
            if frame == o.WORLD:
                assert np.allclose(xd[k, :3], v) and np.allclose(xd[k, 7:10], F / 2.0, atol=1e-13)
            else:
                assert np.allclose(xd[k, :3], R @ v, atol=1e-13)
                assert np.allclose(xd[k, 7:10], R.T @ (F / 2.0) - np.cross(w, v), atol=1e-13)


def test_errstate_structure_and_state_diff():
    """test/liestate.jl:73-99, test/rigid_body_jacobians.jl:64-83: G = cat(I3, ∇differential(q), I6); state_diff = ⊖."""
    rng = np.random.default_rng(10)
    for rn, rc in (("quat", o.ROT_QUAT), ("mrp", o.ROT_MRP), ("rp", o.ROT_RP)):
        m = o.body(rc)
        Z = rand_inputs(m.n, m.m, 8, rng)
        X, X0 = Z[:, :m.n], rand_inputs(m.n, m.m, 8, rng)[:, :m.n]
        G = o.errstate_jacobian(m, X)
        d = o.state_diff(m, X, X0)
        for k in range(8):
            assert np.allclose(G[k].T, ind.errstate_jacobian(rn, X[k]), atol=1e-14)
            assert np.allclose(d[k], ind.state_diff(rn, X[k], X0[k]), atol=1e-13)
        assert np.abs(o.state_diff(m, X, X)).max() < 1e-15
    m = o.cartpole()                                   # Euclidean: G = I, dx = x - x0 (src/statevectortype.jl:144-155)
    X = rng.random((4, 4))
    assert np.allclose(o.errstate_jacobian(m, X), np.eye(4)[None])
    assert np.allclose(o.state_diff(m, X, X[::-1]), X - X[::-1])


def test_grad_errstate_is_derivative_of_Gt_b():
    """∇errstate_jacobian = ∇²differential(q, b) = -(q·b) I3 for quaternions (test/liestate.jl:96-99); zeros outside the rotation block.  The
    MRP / RodriguesParam forms are checked against second differences of scipy's composition in test_rotation_conventions_vs_scipy."""
    rng = np.random.default_rng(11)
    m = o.body(o.ROT_QUAT)
    X = rand_inputs(13, 6, 4, rng)[:, :13]
    B = rng.random((4, 13))
    H = o.grad_errstate_jacobian(m, X, B)
    for k in range(4):
        q = X[k, 3:7]
        assert np.allclose(H[k].T[3:6, 3:6], -(q @ B[k, 3:7]) * np.eye(3), atol=1e-13)      # -(q.b) I3, test/liestate.jl:96-99
        assert np.abs(H[k]).sum() - np.abs(H[k].T[3:6, 3:6]).sum() < 1e-15


# ---- (2) independent complex-step restatement -------------------------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(zoo()))
def test_oracle_vs_independent_complex_step(name):
    m = zoo()[name][0]()
    im = ind_model(name)
    Z = rand_inputs(m.n, m.m, 6, np.random.default_rng(12))
    for Q, qn in QS.items():
        J = o.as_matrix(o.discrete_jacobian(m, Q, Z, 0.05))
        xn = o.discrete_dynamics(m, Q, Z, 0.05)
        for k in range(Z.shape[0]):
            assert np.abs(J[k] - ind.discrete_jacobian(im, qn, Z[k], 0.05)).max() < 1e-11
            assert np.abs(xn[k] - np.real(ind.step(im, qn, Z[k, :m.n], Z[k, m.n:], 0.05))).max() < 1e-13
    Jc = o.as_matrix(o.jacobian(m, Z))
    for k in range(Z.shape[0]):
        assert np.abs(Jc[k] - ind.continuous_jacobian(im, Z[k])).max() < 1e-11


def test_quadrotor_thrust_clamp_derivative():
    """max(0, kf w) (test/quadrotor.jl:67-70): zero derivative when clamped (w < 0); AT the tie w == 0 ForwardDiff keeps the
    partials (Base.max(x, y) = ifelse(isless(y, x), x, y) on the promoted Duals returns y); the yaw moment km*w is NOT clamped."""
    m = o.quadrotor()
    z = rand_inputs(13, 4, 1, np.random.default_rng(13))
    z[0, 13:] = [-0.5, 0.0, 0.7, 0.9]
    J = o.as_matrix(o.jacobian(m, z))[0]
    assert np.all(J[7:10, 13] == 0) and np.any(J[7:10, 14] != 0) and np.any(J[7:10, 15] != 0)
    assert np.allclose(J[7:10, 14], J[7:10, 15])                      # tie: same derivative as an active motor
    assert np.allclose(J[12, 13:], np.array([1, -1, 1, -1]) * 0.0245 / 0.004)


def test_terminal_knot_is_identity():
    m = o.cartpole()
    J = o.as_matrix(o.discrete_jacobian(m, o.RK4, np.random.default_rng(0).random((3, 5)), 0.0))
    assert np.allclose(J, np.concatenate([np.eye(4), np.zeros((4, 1))], axis=1)[None])


# ---- (4) committed fixtures -----------------------------------------------------------------------------------------------------
def test_golden_fixtures_reproduce():
    g = golden("c1_cartpole_rk3")
    assert np.abs(o.discrete_jacobian(o.cartpole(), o.RK3, g["Z"], float(g["dt"])) - g["J"]).max() < 1e-13
    g = golden("c2_cartpole_rk4")
    assert np.abs(o.discrete_jacobian(o.cartpole(), o.RK4, g["Z"], float(g["dt"])) - g["J"]).max() < 1e-13
    g = golden("c3_quadrotor_rk4")
    m = o.quadrotor()
    assert np.abs(o.discrete_jacobian(m, o.RK4, g["Z"], float(g["dt"])) - g["J"]).max() < 1e-13
    assert np.abs(o.errstate_jacobian(m, g["Z"][:, :13]) - g["G"]).max() < 1e-14
    assert np.abs(o.state_diff(m, g["Z"][:, :13], g["X0"]) - g["dX"]).max() < 1e-13
    g = golden("c4_satellite_mrp_rk2")
    assert np.abs(o.discrete_jacobian(o.satellite(o.ROT_MRP), o.RK2, g["Z"], float(g["dt"])) - g["J"]).max() < 1e-13
    g = golden("c5_mixed_sweep")
    assert np.abs(o.discrete_jacobian(o.cartpole(), o.RK4, g["Zc"], g["dtc"]) - g["Jc"]).max() < 1e-13
    assert np.abs(o.discrete_jacobian(m, o.RK4, g["Zq"], g["dtq"]) - g["Jq"]).max() < 1e-13
    g = golden("rollout_cartpole_rk4")
    assert np.abs(o.rollout(o.cartpole(), o.RK4, g["x0"], g["U"], float(g["dt"])) - g["X"]).max() < 1e-13


def test_implicit_midpoint_oracle_vs_independent():
    """ImplicitMidpoint (src/integration.jl:422-463,524-543,620-694): the oracle's Newton + LU + IFT Jacobian against an independent
    numpy solve differentiated THROUGH the converged Newton iteration by complex step; the residual vanishes at the solution
    (test/implicit_dynamics_test.jl checks the same relation)."""
    for name in ("cartpole", "quad_quat_world", "satellite_mrp", "di2"):
        m, im = zoo()[name][0](), ind_model(name)
        Z = rand_inputs(m.n, m.m, 5, np.random.default_rng(14))
        for h in (0.01, 0.1):
            xn = o.discrete_dynamics(m, o.IMPLICIT_MIDPOINT, Z, h)
            J = o.as_matrix(o.discrete_jacobian(m, o.IMPLICIT_MIDPOINT, Z, h))
            for k in range(Z.shape[0]):
                x, u = Z[k, :m.n], Z[k, m.n:]
                res = x + h * o.dynamics(m, np.r_[(x + xn[k]) / 2, u][None])[0] - xn[k]
                assert np.abs(res).max() < 1e-12
                assert np.abs(xn[k] - np.real(ind.implicit_midpoint_step(im, x, u, h))).max() < 1e-12
                assert np.abs(J[k] - ind.discrete_jacobian(im, "implicit_midpoint", Z[k], h)).max() < 1e-10


# ---- (5) symbolic / 50-digit pins (oracle/highprec.py) ----------------------------------------------
def test_cartpole_symbolic_jacobian_and_reference_analytic_jacobian():
    """sympy d f / d z of the Cartpole written as the reference writes it  ==  the reference's hand-derived jacobian!
    (test/cartpole_model.jl:57-96)  ==  the oracle's continuous Jacobian (mirrors test/integration_tests.jl:27-33)."""
    import sympy as sp
    from oracle import highprec as hp
    z, f, Jf = hp.cartpole_symbolic()
    fn, Jn = sp.lambdify(z, f, "numpy"), sp.lambdify(z, Jf, "numpy")
    rng = np.random.default_rng(5)
    Z = np.vstack([rng.random((6, 5)), 3.0 * rng.standard_normal((6, 5))])
    m = o.cartpole()
    Jo = o.as_matrix(o.jacobian(m, Z))
    fo = o.dynamics(m, Z)
    for k, zk in enumerate(Z):
        Jsym = np.array(Jn(*zk), dtype=float)
        scale = max(1.0, np.abs(Jsym).max())
        assert np.abs(np.array(fn(*zk), dtype=float).ravel() - fo[k]).max() < 1e-12 * max(1.0, np.abs(fo[k]).max())
        assert np.abs(Jsym - hp.cartpole_reference_analytic_jacobian(zk)).max() < 1e-11 * scale
        assert np.abs(Jsym - Jo[k]).max() < 1e-11 * scale


@pytest.mark.parametrize("Q", sorted(QS))
def test_cartpole_discrete_jacobian_vs_50_digit_differences(Q):
    """x+ and d x+ / d [x;u] of every explicit rule against 50-digit arithmetic differentiated by central differences."""
    from oracle import highprec as hp
    f = hp.cartpole_f_mp()
    rng = np.random.default_rng(6)
    Z = rng.random((4, 5))
    Z[3] = [0.3, 2.9, -1.2, 4.0, -7.5]                                     # far from the bench distribution
    m = o.cartpole()
    for k, h in enumerate((0.01, 0.05, 0.1, 0.02)):
        xn, J = hp.discrete_jacobian(f, QS[Q], Z[k], 4, h)
        assert np.abs(o.discrete_dynamics(m, Q, Z[k:k + 1], h)[0] - xn).max() < 1e-13 * max(1.0, np.abs(xn).max())
        assert np.abs(o.as_matrix(o.discrete_jacobian(m, Q, Z[k:k + 1], h))[0] - J).max() < 1e-12 * max(1.0, np.abs(J).max())


def test_quadrotor_rk4_jacobian_vs_50_digit_differences():
    """The n = 13 rigid body, including an off-manifold quaternion (the dynamics never renormalises)."""
    from oracle import highprec as hp
    f = hp.quadrotor_f()
    Z = rand_inputs(13, 4, 3, np.random.default_rng(7))
    Z[:, 13:] += 0.5                                                       # thrusts away from the max(0, .) kink
    Z[2, 3:7] *= 1.2
    m = o.quadrotor()
    for k, (Q, h) in enumerate(((o.RK4, 0.01), (o.RK3, 0.05), (o.RK4, 0.1))):
        xn, J = hp.discrete_jacobian(f, QS[Q], Z[k], 13, h)
        assert np.abs(o.discrete_dynamics(m, Q, Z[k:k + 1], h)[0] - xn).max() < 1e-13 * max(1.0, np.abs(xn).max())
        assert np.abs(o.as_matrix(o.discrete_jacobian(m, Q, Z[k:k + 1], h))[0] - J).max() < 1e-12 * max(1.0, np.abs(J).max())


def test_mrp_and_rodrigues_formulas_from_the_quaternion_group():
    """MRP / RodriguesParam `kinematics` and `∇differential` are restated from Rotations.jl's published formulas and no reference
    test pins them.  First principles do: through p <-> q both must agree with the QUATERNION statements the
    reference does pin (kinematics 1/2 L(q) H w, test/liemodel.jl:13-20; composition = Hamilton product):
      kinematics(p, w) = d/dt param(q(t)) with qdot = 1/2 q (x) [0; w],     ∇differential(p) = d param(q(p) (x) q(delta)) / d delta at 0."""
    def qmul(a, b):
        return np.r_[a[0] * b[0] - a[1:] @ b[1:], a[0] * b[1:] + b[0] * a[1:] + np.cross(a[1:], b[1:])]
    maps = {"mrp": (lambda p: np.r_[1 - p @ p, 2 * p] / (1 + p @ p), lambda q: q[1:] / (1 + q[0]), o.ROT_MRP),
            "rp": (lambda g: np.r_[1.0, g] / np.sqrt(1 + g @ g), lambda q: q[1:] / q[0], o.ROT_RP)}
    rng = np.random.default_rng(12)
    eps = 1e-6
    for name, (to_q, from_q, rc) in maps.items():
        m = o.body(rc)
        Z = rand_inputs(12, 6, 5, rng)
        xd = o.dynamics(m, Z)
        G = o.errstate_jacobian(m, Z[:, :12])                      # (N, nerr, n) column-major per knot: G[k].T is n x nerr
        for k in range(5):
            p, w = Z[k, 3:6], Z[k, 9:12]
            q = to_q(p)
            qd = 0.5 * qmul(q, np.r_[0.0, w])
            unit = lambda v: v / np.linalg.norm(v)
            pdot = (from_q(unit(q + eps * qd)) - from_q(unit(q - eps * qd))) / (2 * eps)
            assert np.abs(xd[k, 3:6] - pdot).max() < 1e-8
            D = np.zeros((3, 3))
            for j in range(3):
                d = np.zeros(3); d[j] = eps
                D[:, j] = (from_q(qmul(q, to_q(d))) - from_q(qmul(q, to_q(-d)))) / (2 * eps)
            assert np.abs(G[k].T[3:6, 3:6] - D).max() < 1e-8


def test_satellite_mrp_rk2_jacobian_vs_50_digit_differences():
    """BASELINE config C4: Satellite RigidBody{MRP}, RK2 (explicit midpoint), dt = 0.1 (examples/single_satellite.jl:39)."""
    from oracle import highprec as hp
    f = hp.satellite_mrp_f()
    Z = rand_inputs(12, 6, 3, np.random.default_rng(8))
    m = o.satellite(o.ROT_MRP)
    for k, (Q, h) in enumerate(((o.RK2, 0.1), (o.RK2, 0.01), (o.RK4, 0.1))):
        xn, J = hp.discrete_jacobian(f, QS[Q], Z[k], 12, h)
        assert np.abs(o.discrete_dynamics(m, Q, Z[k:k + 1], h)[0] - xn).max() < 1e-13 * max(1.0, np.abs(xn).max())
        assert np.abs(o.as_matrix(o.discrete_jacobian(m, Q, Z[k:k + 1], h))[0] - J).max() < 1e-12 * max(1.0, np.abs(J).max())


def test_implicit_midpoint_vs_50_digit_root_and_differences():
    """DiscretizedDynamics{L, ImplicitMidpoint} (src/integration.jl:620-694): x2 solves x1 + h f((x1 + x2)/2, u) - x2 = 0.  The root is
    found here at 50 digits (mpmath.findroot) and the Jacobian of the solution map by central differences — no Newton loop of ours, no
    implicit-function-theorem formula."""
    import mpmath as mpm
    from oracle import highprec as hp
    f = hp.cartpole_f_mp()
    m = o.cartpole()
    rng = np.random.default_rng(13)
    Z = rng.random((3, 5))

    def solve(z, h):
        x1, u = z[:4], z[4:]
        g = lambda *x2: [a + h * fi - b for a, fi, b in zip(x1, f([(a + b) / 2 for a, b in zip(x1, x2)], u), x2)]
        return list(mpm.findroot(g, x1, tol=mpm.mpf(10) ** -40))

    for k, h in enumerate((0.01, 0.05, 0.1)):
        zz, hh = [mpm.mpf(float(v)) for v in Z[k]], mpm.mpf(h)
        xn = np.array([float(v) for v in solve(zz, hh)])
        J = np.zeros((4, 5))
        eps = mpm.mpf(10) ** -15
        for j in range(5):
            zp, zm = list(zz), list(zz)
            zp[j] += eps
            zm[j] -= eps
            J[:, j] = [float((a - b) / (2 * eps)) for a, b in zip(solve(zp, hh), solve(zm, hh))]
        assert np.abs(o.discrete_dynamics(m, o.IMPLICIT_MIDPOINT, Z[k:k + 1], h)[0] - xn).max() < 1e-12
        assert np.abs(o.as_matrix(o.discrete_jacobian(m, o.IMPLICIT_MIDPOINT, Z[k:k + 1], h))[0] - J).max() < 1e-10


# ---- (6) the Rotations.jl boundary against an independent third-party implementation (scipy.spatial.transform) ------------------------
def _scipy_rot(rn, att):
    """The rotation Rotations.jl builds for an attitude: QuatRotation [w,x,y,z] (Hamilton, active), MRP p = v / (1 + w), RodriguesParam
    g = v / w  (SURVEY.md §8c).  scipy: quaternions are [x,y,z,w]; MRPs are tan(theta/4) axis; a Gibbs vector is tan(theta/2) axis."""
    from scipy.spatial.transform import Rotation as R
    if rn == "quat":
        return R.from_quat(np.r_[att[1:], att[0]])
    if rn == "mrp":
        return R.from_mrp(att)
    nrm = np.linalg.norm(att)
    return R.from_rotvec(2.0 * np.arctan(nrm) * att / nrm)


def _scipy_att(rn, rot, like=None):
    if rn == "quat":
        q = rot.as_quat()
        q = np.r_[q[3], q[:3]]
        return -q if like is not None and q @ like < 0 else q
    if rn == "mrp":
        return rot.as_mrp()
    v = rot.as_rotvec()
    th = np.linalg.norm(v)
    return np.tan(th / 2.0) * v / th


@pytest.mark.parametrize("rn,rc", [("quat", o.ROT_QUAT), ("mrp", o.ROT_MRP), ("rp", o.ROT_RP)])
def test_rotation_conventions_vs_scipy(rn, rc):
    """What the reference takes from Rotations.jl — `q * v`, `q \\ F`, `kinematics(q, w)`, `rotation_error(.., CayleyMap())`, `∇differential`
    (src/rigidbody.jl:224-230, src/liestate.jl:216-217,291) — compared with scipy's Rotation class on the manifold: active Hamilton
    rotations, body-frame angular velocity (right multiplication), MRP = tan(theta/4) axis, Rodrigues = tan(theta/2) axis, Cayley error =
    Gibbs vector of R0' R.  An independent implementation, not the reference: it pins conventions, not off-manifold behaviour."""
    from scipy.spatial.transform import Rotation as R
    rng = np.random.default_rng(77)
    m = o.body(rc, o.BODYFRAME)                      # mass 2, J = diag(2,3,1): r' = q*v, v' = q\(F/m) - w x v, F_world = q*u[:3]
    n, na = m.n, m.n - 9
    Z = rand_inputs(n, 6, 6, rng)
    xd = o.dynamics(m, Z)
    G = o.errstate_jacobian(m, Z[:, :n])
    X0 = rand_inputs(n, 6, 6, rng)[:, :n]
    d = o.state_diff(m, Z[:, :n], X0)
    eps = 1e-6
    for k in range(Z.shape[0]):
        att, v, w, u = Z[k, 3:3 + na], Z[k, 3 + na:6 + na], Z[k, 6 + na:9 + na], Z[k, n:]
        rot = _scipy_rot(rn, att)
        assert np.allclose(xd[k, :3], rot.apply(v), atol=1e-12)                                        # r' = q * v
        Fw = rot.apply(u[:3])                                                                         # F_world = q * u[1:3]
        assert np.allclose(xd[k, 3 + na:6 + na], rot.inv().apply(Fw / 2.0) - np.cross(w, v), atol=1e-12)   # v' = q \ (F/m) - w x v
        # kinematics: d/dt of the attitude parameters when R(t + e) = R(t) exp(e w^)   (central difference on the manifold)
        ap = _scipy_att(rn, rot * R.from_rotvec(eps * w), like=att if rn == "quat" else None)
        am = _scipy_att(rn, rot * R.from_rotvec(-eps * w), like=att if rn == "quat" else None)
        assert np.allclose(xd[k, 3:3 + na], (ap - am) / (2 * eps), atol=1e-8)
        # errstate_jacobian attitude block = ∇differential = d att(R * dR(delta)) / d delta at 0, where delta parameterises the small right
        # perturbation dR the way the attitude itself is parameterised: the vector part of a quaternion [1, delta] (= a Gibbs vector) for
        # QuatRotation, a Gibbs vector for RodriguesParam, an MRP for MRP
        Gk = G[k].T[3:3 + na, 3:6]
        for j in range(3):
            dl = np.zeros(3); dl[j] = eps
            small = (lambda s: R.from_mrp(s * dl)) if rn == "mrp" else (lambda s: R.from_rotvec(2.0 * np.arctan(eps) * s * dl / eps))
            col = (_scipy_att(rn, rot * small(+1), like=att if rn == "quat" else None) -
                   _scipy_att(rn, rot * small(-1), like=att if rn == "quat" else None)) / (2 * eps)
            assert np.allclose(Gk[:, j], col, atol=1e-6)
        # ∇errstate_jacobian attitude block = ∇²differential(att, b) = sum_k b_k d²att_k(R * dR(delta)) / d delta² at 0: the second-order
        # term of the retraction (for quaternions -(q.b) I3, test/liestate.jl:96-99), by second central differences of scipy's composition
        b = rng.random(n)
        H = o.grad_errstate_jacobian(m, Z[k:k + 1, :n], b[None])[0].T[3:6, 3:6]
        e2 = 1e-4
        def comp(dv):
            nd = np.linalg.norm(dv)
            dR = R.from_mrp(dv) if rn == "mrp" else R.from_rotvec(2.0 * np.arctan(nd) * dv / nd)
            return _scipy_att(rn, rot * dR, like=att if rn == "quat" else None) @ b[3:3 + na]
        for i in range(3):
            for j in range(3):
                ei, ej = np.eye(3)[i] * e2, np.eye(3)[j] * e2
                hij = (comp(ei + ej) - comp(ei - ej) - comp(-ei + ej) + comp(-ei - ej)) / (4 * e2 * e2) if i != j else \
                      (comp(ei) - 2 * (att @ b[3:3 + na]) + comp(-ei)) / (e2 * e2)
                assert abs(H[i, j] - hij) < 2e-6, (rn, i, j, H[i, j], hij)
        # state_diff attitude part = Cayley map of R0' R = its Gibbs vector
        rel = _scipy_rot(rn, X0[k, 3:3 + na]).inv() * rot
        rv = rel.as_rotvec(); th = np.linalg.norm(rv)
        assert np.allclose(d[k, 3:6], np.tan(th / 2.0) * rv / th, atol=1e-11)


def test_unnormalised_rotation_identity_behind_the_split_body_frame_force():
    """models.cuh (SPLIT) evaluates the body-frame q \\ (q * Fb + Gw) as |q|^4 Fb + q \\ Gw.  Symbolically, for the polynomial rotation the
    reference path uses (q * r = (w^2 - v.v) r + 2 v (v.r) + 2 w (v x r), q \\ r = conj(q) * r, no normalisation: SURVEY §8c), the identity
    q \\ (q * F) = |q|^4 F holds for EVERY quaternion — hence also for every derivative with respect to q and F."""
    sympy = pytest.importorskip("sympy")
    w, x, y, z, f0, f1, f2 = sympy.symbols("w x y z f0 f1 f2", real=True)

    def rot(qw, v, r):
        v, r = sympy.Matrix(v), sympy.Matrix(r)
        return (qw**2 - v.dot(v)) * r + 2 * v * v.dot(r) + 2 * qw * v.cross(r)

    F = sympy.Matrix([f0, f1, f2])
    back = rot(w, [-x, -y, -z], rot(w, [x, y, z], F))
    n2 = w**2 + x**2 + y**2 + z**2
    assert all(sympy.expand(e) == 0 for e in (back - n2**2 * F))
