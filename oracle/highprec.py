"""Third independent pin of the oracle: symbolic (sympy) and 50-digit (mpmath) restatements.  TEST INFRASTRUCTURE ONLY.

SURVEY.md §7 step 1 asks for cross-checks that need no Julia: (ii) a symbolic Jacobian for Cartpole and (iii) high-precision
differentiation at a handful of points.  Nothing here shares code with oracle/rd_oracle.cpp or oracle/independent.py:

  * `cartpole_symbolic()`   the Cartpole dynamics exactly as the reference writes them — H, C, G, B and  qdd = -H \\ (C qd + G - B u)
                            (reference: test/cartpole_model.jl:11-30) — as sympy expressions, and their symbolic Jacobian;
  * `cartpole_reference_analytic_jacobian()`  the reference's HAND-DERIVED continuous Jacobian (test/cartpole_model.jl:57-96),
                            restated numerically; the reference asserts it equals ForwardDiff (test/integration_tests.jl:27-33);
  * `rk_step()` / `discrete_jacobian()`  Euler / RK2 / RK3 / RK4 (reference: src/integration.jl:73-76,130-135,280-286; RK2 = explicit
                            midpoint) composed in mpmath at 50 digits, differentiated by mpmath's high-precision central differences
                            (no derivative formula of any kind is involved);
  * `quadrotor_f()`         RigidBody{QuatRotation} with the Quadrotor wrench (reference: src/rigidbody.jl:213-236,
                            test/quadrotor.jl:56-96; rotation formulas of SURVEY.md §8c) in plain mpmath arithmetic;
  * `satellite_mrp_f()`     RigidBody{MRP} with the Satellite wrench (examples/single_satellite.jl:17-35), BASELINE config C4.
"""
import mpmath as mp
import numpy as np

mp.mp.dps = 50


# ---- Cartpole, symbolic -------------------------------------------------------------------------------------------------------
def cartpole_symbolic(mc=1.0, mp_=0.2, l=0.5, g=9.81):
    """(z symbols, f(z) as a sympy column, df/dz as a sympy matrix); parameters enter as exact rationals of their decimal strings."""
    import sympy as sp
    x1, x2, x3, x4, u1 = z = sp.symbols("x1 x2 x3 x4 u1", real=True)
    mc, mp_, l, g = (sp.Rational(str(v)) for v in (mc, mp_, l, g))
    s, c = sp.sin(x2), sp.cos(x2)
    qd = sp.Matrix([x3, x4])
    H = sp.Matrix([[mc + mp_, mp_ * l * c], [mp_ * l * c, mp_ * l ** 2]])
    C = sp.Matrix([[0, -mp_ * x4 * l * s], [0, 0]])
    G = sp.Matrix([0, mp_ * g * l * s])
    B = sp.Matrix([1, 0])
    qdd = -H.LUsolve(C * qd + G - B * u1)
    f = sp.Matrix([x3, x4, qdd[0], qdd[1]])
    return z, f, f.jacobian(sp.Matrix(z))


def cartpole_reference_analytic_jacobian(zv, mc=1.0, mp_=0.2, l=0.5, g=9.81):
    """The reference's hand-derived jacobian!(model::Cartpole, J, xdot, x, u, t), test/cartpole_model.jl:57-96, in numpy."""
    x = np.asarray(zv[:4], dtype=float)
    u = float(zv[4])
    qd = x[2:4]
    s, c = np.sin(x[1]), np.cos(x[1])
    H = np.array([[mc + mp_, mp_ * l * c], [mp_ * l * c, mp_ * l ** 2]])
    C = np.array([[0, -mp_ * qd[1] * l * s], [0, 0]])
    G = np.array([0, mp_ * g * l * s])
    B = np.array([1.0, 0.0])
    qdd = -np.linalg.solve(H, C @ qd + G - B * u)
    dH = np.array([[0, -mp_ * l * s * qdd[1], 0, 0, 0], [0, -mp_ * l * s * qdd[0], 0, 0, 0]])
    dC = np.array([[0, -mp_ * l * c * qd[1] ** 2, 0, -2 * mp_ * l * qd[1] * s, 0], [0, 0, 0, 0, 0]])
    dG = np.array([[0, 0, 0, 0, 0], [0, mp_ * g * l * c, 0, 0, 0]])
    dB = np.array([[0, 0, 0, 0, 1.0], [0, 0, 0, 0, 0]])
    J = np.zeros((4, 5))
    J[0, 2] = 1.0
    J[1, 3] = 1.0
    J[2:4] = np.linalg.solve(H, -dH - dC - dG + dB)
    return J


def cartpole_f_mp():
    """f(x, u) -> list of mpf, lambdified from the symbolic expressions."""
    import sympy as sp
    z, f, _ = cartpole_symbolic()
    fn = sp.lambdify(z, list(f), modules="mpmath")
    return lambda x, u: list(fn(*x, *u))


# ---- Quadrotor, plain mpmath ----------------------------------------------------------------------------------------------------
def _rotate(q, r):
    """q*r = (w^2 - v.v) r + 2 v (v.r) + 2 w (v x r), un-normalised."""
    w, v = q[0], q[1:4]
    vv = sum(a * a for a in v)
    vr = sum(a * b for a, b in zip(v, r))
    cr = [v[1] * r[2] - v[2] * r[1], v[2] * r[0] - v[0] * r[2], v[0] * r[1] - v[1] * r[0]]
    return [(w * w - vv) * r[i] + 2 * v[i] * vr + 2 * w * cr[i] for i in range(3)]


def quadrotor_f(mass="0.5", J=("0.0023", "0.0023", "0.004"), gz="-9.81", L="0.175", kf="1.0", km="0.0245"):
    mass, gz, L, kf, km = (mp.mpf(v) for v in (mass, gz, L, kf, km))
    J = [mp.mpf(v) for v in J]

    def f(x, u):
        q, v, w = x[3:7], x[7:10], x[10:13]
        F = [mp.mpf(0) if ui < 0 else kf * ui for ui in u]                      # max(0, kf w_i), test/quadrotor.jl:67-70
        Fw = _rotate(q, [mp.mpf(0), mp.mpf(0), sum(F)])
        Fw[2] += mass * gz
        tau = [L * (F[1] - F[3]), L * (F[2] - F[0]), km * (u[0] - u[1] + u[2] - u[3])]   # moments use the unclamped inputs
        qw, qx, qy, qz = q
        qdot = [(-qx * w[0] - qy * w[1] - qz * w[2]) / 2, (qw * w[0] - qz * w[1] + qy * w[2]) / 2,
                (qz * w[0] + qw * w[1] - qx * w[2]) / 2, (-qy * w[0] + qx * w[1] + qw * w[2]) / 2]   # 1/2 L(q) H w
        Jw = [J[i] * w[i] for i in range(3)]
        wxJw = [w[1] * Jw[2] - w[2] * Jw[1], w[2] * Jw[0] - w[0] * Jw[2], w[0] * Jw[1] - w[1] * Jw[0]]
        wdot = [(tau[i] - wxJw[i]) / J[i] for i in range(3)]
        return list(v) + qdot + [Fw[i] / mass for i in range(3)] + wdot
    return f


def satellite_mrp_f(mass="1.0", J=("1.0", "1.0", "1.0")):
    """RigidBody{MRP} with the Satellite / Body wrench F_world = q*u[1:3], M_body = u[4:6] (reference: src/rigidbody.jl:213-236,
    examples/single_satellite.jl:17-35); state [r(3) p(3) v(3) w(3)], world-frame velocity, diagonal inertia."""
    mass = mp.mpf(mass)
    J = [mp.mpf(v) for v in J]

    def f(x, u):
        p, v, w = x[3:6], x[6:9], x[9:12]
        n2 = sum(a * a for a in p)
        q = [(1 - n2) / (1 + n2)] + [2 * a / (1 + n2) for a in p]                 # the unit quaternion of the MRP
        F = _rotate(q, list(u[0:3]))
        pw = sum(a * b for a, b in zip(p, w))
        cr = [p[1] * w[2] - p[2] * w[1], p[2] * w[0] - p[0] * w[2], p[0] * w[1] - p[1] * w[0]]
        pdot = [((1 - n2) * w[i] + 2 * cr[i] + 2 * p[i] * pw) / 4 for i in range(3)]   # 1/4 [(1-|p|^2) I + 2 skew(p) + 2 p p'] w
        Jw = [J[i] * w[i] for i in range(3)]
        wxJw = [w[1] * Jw[2] - w[2] * Jw[1], w[2] * Jw[0] - w[0] * Jw[2], w[0] * Jw[1] - w[1] * Jw[0]]
        wdot = [(u[3 + i] - wxJw[i]) / J[i] for i in range(3)]
        return list(v) + pdot + [F[i] / mass for i in range(3)] + wdot
    return f


# ---- integrators and differentiation, 50 digits -----------------------------------------------------------------------------------
def rk_step(f, rule, x, u, h):
    ax = lambda a, s, b: [ai + s * bi for ai, bi in zip(a, b)]
    if rule == "euler":                       # src/integration.jl:73-76
        return ax(x, h, f(x, u))
    if rule == "rk2":                         # explicit midpoint, test/old_tests/linear_tests.jl:135-141
        return ax(x, h, f(ax(x, h / 2, f(x, u)), u))
    if rule == "rk3":                         # src/integration.jl:130-135
        k1 = [h * a for a in f(x, u)]
        k2 = [h * a for a in f(ax(x, mp.mpf(1) / 2, k1), u)]
        k3 = [h * a for a in f(ax(ax(x, -1, k1), 2, k2), u)]
        return [xi + (a + 4 * b + c) / 6 for xi, a, b, c in zip(x, k1, k2, k3)]
    if rule == "rk4":                         # src/integration.jl:280-286
        k1 = [h * a for a in f(x, u)]
        k2 = [h * a for a in f(ax(x, mp.mpf(1) / 2, k1), u)]
        k3 = [h * a for a in f(ax(x, mp.mpf(1) / 2, k2), u)]
        k4 = [h * a for a in f(ax(x, 1, k3), u)]
        return [xi + (a + 2 * b + 2 * c + d) / 6 for xi, a, b, c, d in zip(x, k1, k2, k3, k4)]
    raise ValueError(rule)


def discrete_jacobian(f, rule, z, n, h):
    """(x+, d x+ / d [x;u]) at the fp64 point z, both rounded to fp64 from 50-digit arithmetic.  Central differences with a
    1e-20 step at 50 digits: truncation ~1e-40 relative, rounding ~1e-30 — far below fp64."""
    zz = [mp.mpf(float(v)) for v in z]
    hh = mp.mpf(float(h))
    step = lambda zc: rk_step(f, rule, zc[:n], zc[n:], hh)
    xn = step(zz)
    J = np.zeros((n, len(zz)))
    eps = mp.mpf(10) ** -20
    for j in range(len(zz)):
        zp, zm = list(zz), list(zz)
        zp[j] += eps
        zm[j] -= eps
        fp, fm = step(zp), step(zm)
        J[:, j] = [float((a - b) / (2 * eps)) for a, b in zip(fp, fm)]
    return np.array([float(v) for v in xn]), J
