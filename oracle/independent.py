"""Independent numpy restatement + complex-step differentiation.  TEST INFRASTRUCTURE ONLY.

A second, separately written statement of the reference's models and integrators (pure numpy, works on
complex inputs), used to cross-check oracle/rd_oracle.cpp without sharing any code with it:
Jacobians here come from complex-step differentiation  J[:,j] = Im f(z + i·h·e_j) / h,  h = 1e-30,
which is exact to rounding for analytic functions and involves no hand-derived derivative.
`max(0, ·)` (quadrotor thrust clamp, test/quadrotor.jl:67-70) is applied on the real part.

Reference anchors: test/cartpole_model.jl:11-30, src/rigidbody.jl:213-236, test/quadrotor.jl:56-96,
test/rigidbody_test.jl:26-31, src/integration.jl:73-76,130-135,280-286, src/liestate.jl:210-298.
"""
import numpy as np


def skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]], dtype=np.result_type(v, float))


def quat_matrix(q):
    """Un-normalised 'rotation matrix' of q*r = (w²-v'v) r + 2 v (v'r) + 2 w (v×r)."""
    w, v = q[0], np.asarray(q[1:4])
    return (w * w - v @ v) * np.eye(3) + 2 * np.outer(v, v) + 2 * w * skew(v)


def lmult(q):
    w, x, y, z = q
    return np.array([[w, -x, -y, -z], [x, w, -z, y], [y, z, w, -x], [z, -y, x, w]])


def to_quat(rot, p):
    p = np.asarray(p)
    if rot == "quat":
        return p
    n2 = p @ p
    if rot == "mrp":
        return np.concatenate([[(1 - n2) / (1 + n2)], 2 * p / (1 + n2)])
    return np.concatenate([[1.0 + 0 * n2], p]) / np.sqrt(1 + n2)


def kinematics(rot, p, w):
    p = np.asarray(p)
    if rot == "quat":
        return 0.5 * lmult(p)[:, 1:] @ w
    n2 = p @ p
    if rot == "mrp":
        return 0.25 * (((1 - n2) * np.eye(3) + 2 * skew(p) + 2 * np.outer(p, p)) @ w)
    return 0.5 * ((np.eye(3) + skew(p) + np.outer(p, p)) @ w)


def relu(a):
    return 0 * a if np.real(a) < 0 else a          # Base.max(0, a) = ifelse(a < 0, 0, a): the tie a == 0 returns a (partials kept)


class Cartpole:
    n, m = 4, 1

    def __init__(self, mc=1.0, mp=0.2, l=0.5, g=9.81):
        self.mc, self.mp, self.l, self.g = mc, mp, l, g

    def f(self, x, u):
        mc, mp, l, g = self.mc, self.mp, self.l, self.g
        qd = x[2:4]
        s, c = np.sin(x[1]), np.cos(x[1])
        H = np.array([[mc + mp + 0 * c, mp * l * c], [mp * l * c, mp * l * l + 0 * c]])
        C = np.array([[0 * s, -mp * qd[1] * l * s], [0 * s, 0 * s]])
        G = np.array([0 * s, mp * g * l * s])
        B = np.array([1.0, 0.0])
        qdd = -np.linalg.solve(H, C @ qd + G - B * u[0])
        return np.concatenate([qd, qdd])


class RigidBody:
    def __init__(self, kind, rot="quat", frame="world", mass=2.0, J=np.diag([2.0, 3.0, 1.0]),
                 gravity=(0, 0, -9.81), motor_dist=0.175, kf=1.0, km=0.0245):
        self.kind, self.rot, self.frame = kind, rot, frame
        self.mass, self.J = mass, np.asarray(J, float)
        self.gravity = np.asarray(gravity, float)
        self.L, self.kf, self.km = motor_dist, kf, km
        self.np_ = 4 if rot == "quat" else 3
        self.n = 9 + self.np_
        self.m = 4 if kind == "quadrotor" else 6

    def f(self, x, u):
        k = self.np_
        q, v, w = x[3:3 + k], x[3 + k:6 + k], x[6 + k:9 + k]
        R = quat_matrix(to_quat(self.rot, q))
        if self.kind == "quadrotor":
            F_ = [relu(self.kf * ui) for ui in u]
            F = self.mass * self.gravity + R @ np.array([0 * F_[0], 0 * F_[0], F_[0] + F_[1] + F_[2] + F_[3]])
            tau = np.array([self.L * (F_[1] - F_[3]), self.L * (F_[2] - F_[0]),
                            self.km * (u[0] - u[1] + u[2] - u[3])])
        else:
            F = R @ u[0:3]
            tau = u[3:6]
        qdot = kinematics(self.rot, q, w)
        if self.frame == "world":
            rdot, vdot = v, F / self.mass
        else:
            qc = to_quat(self.rot, q) * np.array([1, -1, -1, -1])
            rdot = R @ v
            vdot = quat_matrix(qc) @ (F / self.mass) - np.cross(w, v)
        wdot = np.linalg.solve(self.J.astype(x.dtype), tau - np.cross(w, self.J @ w))
        return np.concatenate([rdot, qdot, vdot, wdot])


class DoubleIntegrator:
    def __init__(self, D):
        self.D, self.n, self.m = D, 2 * D, D

    def f(self, x, u):
        return np.concatenate([x[self.D:], u])


def step(model, Q, x, u, h):
    f = model.f
    if Q == "euler":
        return x + h * f(x, u)
    k1 = f(x, u) * h
    k2 = f(x + k1 / 2, u) * h
    if Q == "rk2":
        return x + k2
    if Q == "rk3":
        k3 = f(x - k1 + 2 * k2, u) * h
        return x + (k1 + 4 * k2 + k3) / 6
    k3 = f(x + k2 / 2, u) * h
    k4 = f(x + k3, u) * h
    return x + (k1 + 2 * k2 + 2 * k3 + k4) / 6


def implicit_midpoint_step(model, x, u, h, iters=30):
    """x2 solving x + h f((x + x2)/2, u) - x2 = 0 by Newton with a complex-step Jacobian of the residual; works on complex
    inputs, so differentiating THROUGH the converged solve by complex step gives the implicit-function-theorem Jacobian
    (reference: src/integration.jl:422-463, 524-543, 620-694) without forming it."""
    x = np.asarray(x, dtype=complex); u = np.asarray(u, dtype=complex)
    x2 = x.copy()
    n = len(x)
    for _ in range(iters):
        r = x + h * model.f((x + x2) / 2, u) - x2
        # d r / d x2 at the REAL part (the imaginary parts are O(1e-30) perturbations carried along linearly)
        xr, ur, x2r = np.real(x), np.real(u), np.real(x2)
        A = np.stack([np.imag(xr + h * model.f((xr + x2r + 1e-30j * np.eye(n)[j]) / 2, ur.astype(complex)) - (x2r + 1e-30j * np.eye(n)[j])) / 1e-30
                      for j in range(n)], axis=1)
        x2 = x2 - np.linalg.solve(A.astype(complex), r)
    return x2


def complex_step_jacobian(fun, z, h=1e-30):
    z = np.asarray(z, dtype=complex)
    cols = []
    for j in range(len(z)):
        zp = z.copy()
        zp[j] += 1j * h
        cols.append(np.imag(fun(zp)) / h)
    return np.stack(cols, axis=1)


def discrete_jacobian(model, Q, z, h):
    n = model.n
    if Q == "implicit_midpoint":
        return complex_step_jacobian(lambda zz: implicit_midpoint_step(model, zz[:n], zz[n:], h), z)
    return complex_step_jacobian(lambda zz: step(model, Q, zz[:n], zz[n:], h), z)


def continuous_jacobian(model, z):
    n = model.n
    return complex_step_jacobian(lambda zz: model.f(zz[:n], zz[n:]), z)


def errstate_jacobian(rot, x):
    """cat(I3, ∇differential(q), I6) — test/rigid_body_jacobians.jl:72-77."""
    k = 4 if rot == "quat" else 3
    p = np.asarray(x[3:3 + k], float)
    if rot == "quat":
        p = p / np.linalg.norm(p)
        D = lmult(p)[:, 1:]
    elif rot == "mrp":
        D = (1 - p @ p) * np.eye(3) + 2 * (skew(p) + np.outer(p, p))
    else:
        D = np.eye(3) + skew(p) + np.outer(p, p)
    G = np.zeros((9 + k, 12))
    G[0:3, 0:3] = np.eye(3)
    G[3:3 + k, 3:6] = D
    G[3 + k:, 6:] = np.eye(6)
    return G


def state_diff(rot, x, x0):
    k = 4 if rot == "quat" else 3
    q, q0 = to_quat(rot, np.asarray(x[3:3 + k], float)), to_quat(rot, np.asarray(x0[3:3 + k], float))
    q, q0 = q / np.linalg.norm(q), q0 / np.linalg.norm(q0)
    e = lmult(q0 * np.array([1, -1, -1, -1])) @ q
    return np.concatenate([x[0:3] - x0[0:3], e[1:] / e[0], x[3 + k:] - x0[3 + k:]])
