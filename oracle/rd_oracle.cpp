// rd_oracle.cpp — CPU ORACLE for the RobotDynamics.jl hot path.  TEST INFRASTRUCTURE ONLY.
//
// This file is a plain C++17 restatement of the reference's algorithm.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
// The product (robotdynamics.jl_b200/) never links, imports or calls anything in oracle/.
//
// PARITY STATUS: "parity unpinned" against Julia output.  The reference ships no golden vectors
//  and Julia is not installed here, so this restatement is pinned by
//   (1) every relational assertion the reference's own tests make on this path
//       (test/integration_tests.jl:7-18,27-40,47-61,80-111; test/rigidbody_test.jl:135-162;
//        test/liestate.jl:73-99; test/rigid_body_jacobians.jl:55-83) — re-run in tests/test_oracle.py;
//   (2) an independent numpy complex-step restatement (oracle/independent.py);
//   (3) the hand-verified known-answer vectors of SURVEY.md Appendix C.
// The rotation arithmetic lives in Rotations.jl 1.x / Quaternions.jl 0.7 (Project.toml:11-13,21-23),
// which are NOT under /root/reference; their published formulas are restated below.
//
// Two independent Jacobian paths, like the reference:
//   method 0: forward-mode Dual numbers through the whole integrator   (src/jacobian_gen.jl:485-507)
//   method 1: per-stage continuous Jacobians + integrator chain rule   (src/integration.jl:85-93,149-177,302-337)
//
// Layout: everything is fp64, "AoS": z_k = [x;u] contiguous per knot, J_k = n x (n+m) column-major
// per knot (src/jacobian.jl:26-37).

#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace rdo {

// ---------------------------------------------------------------------------------------------
// Forward-mode dual number: value + NP partials (ForwardDiff.Dual semantics).
// ---------------------------------------------------------------------------------------------
template <int NP>
struct Dual {
    double v;
    double d[NP];
    Dual() : v(0) { for (int i = 0; i < NP; ++i) d[i] = 0; }
    Dual(double a) : v(a) { for (int i = 0; i < NP; ++i) d[i] = 0; }
};
template <int NP> inline Dual<NP> operator+(const Dual<NP>& a, const Dual<NP>& b) { Dual<NP> r; r.v = a.v + b.v; for (int i = 0; i < NP; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
template <int NP> inline Dual<NP> operator-(const Dual<NP>& a, const Dual<NP>& b) { Dual<NP> r; r.v = a.v - b.v; for (int i = 0; i < NP; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
template <int NP> inline Dual<NP> operator-(const Dual<NP>& a) { Dual<NP> r; r.v = -a.v; for (int i = 0; i < NP; ++i) r.d[i] = -a.d[i]; return r; }
template <int NP> inline Dual<NP> operator*(const Dual<NP>& a, const Dual<NP>& b) { Dual<NP> r; r.v = a.v * b.v; for (int i = 0; i < NP; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
template <int NP> inline Dual<NP> operator/(const Dual<NP>& a, const Dual<NP>& b) { Dual<NP> r; double ib = 1.0 / b.v; r.v = a.v * ib; for (int i = 0; i < NP; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * ib; return r; }
template <int NP> inline Dual<NP> operator+(const Dual<NP>& a, double b) { Dual<NP> r = a; r.v += b; return r; }
template <int NP> inline Dual<NP> operator+(double b, const Dual<NP>& a) { return a + b; }
template <int NP> inline Dual<NP> operator-(const Dual<NP>& a, double b) { Dual<NP> r = a; r.v -= b; return r; }
template <int NP> inline Dual<NP> operator-(double b, const Dual<NP>& a) { return (-a) + b; }
template <int NP> inline Dual<NP> operator*(const Dual<NP>& a, double b) { Dual<NP> r; r.v = a.v * b; for (int i = 0; i < NP; ++i) r.d[i] = a.d[i] * b; return r; }
template <int NP> inline Dual<NP> operator*(double b, const Dual<NP>& a) { return a * b; }
template <int NP> inline Dual<NP> operator/(const Dual<NP>& a, double b) { return a * (1.0 / b); }
template <int NP> inline Dual<NP> operator/(double a, const Dual<NP>& b) { return Dual<NP>(a) / b; }

inline double sin_(double a) { return std::sin(a); }
inline double cos_(double a) { return std::cos(a); }
inline double sqrt_(double a) { return std::sqrt(a); }
// max(0, a): under ForwardDiff the 0 is promoted and Base.max(x, y) = ifelse(isless(y, x), x, y) returns y = a unless a < 0, so at the
// tie a == 0 the partials of a are KEPT (DiffRules' max rule agrees: d/dy = x > y ? 0 : 1)   (reference: test/quadrotor.jl:67-70)
inline double relu_(double a) { return a < 0 ? 0.0 : a; }
inline double val(double a) { return a; }
template <int NP> inline Dual<NP> sin_(const Dual<NP>& a) { Dual<NP> r; r.v = std::sin(a.v); double c = std::cos(a.v); for (int i = 0; i < NP; ++i) r.d[i] = c * a.d[i]; return r; }
template <int NP> inline Dual<NP> cos_(const Dual<NP>& a) { Dual<NP> r; r.v = std::cos(a.v); double s = -std::sin(a.v); for (int i = 0; i < NP; ++i) r.d[i] = s * a.d[i]; return r; }
template <int NP> inline Dual<NP> sqrt_(const Dual<NP>& a) { Dual<NP> r; r.v = std::sqrt(a.v); double s = 0.5 / r.v; for (int i = 0; i < NP; ++i) r.d[i] = s * a.d[i]; return r; }
// ForwardDiff: max(0, Dual) compares values; derivative is 0 when the clamp is active or at exactly 0
// (test/quadrotor.jl:67-70).
template <int NP> inline Dual<NP> relu_(const Dual<NP>& a) { return a.v < 0 ? Dual<NP>(0.0) : a; }
template <int NP> inline double val(const Dual<NP>& a) { return a.v; }

// ---------------------------------------------------------------------------------------------
// Model descriptor (same enums and parameter packing as include/rdb200.h).
// ---------------------------------------------------------------------------------------------
enum Kind { CARTPOLE = 0, QUADROTOR = 1, BODY = 2, DOUBLE_INTEGRATOR = 3 };
enum Rot { ROT_NONE = 0, ROT_QUAT = 1, ROT_MRP = 2, ROT_RP = 3 };
enum Frame { FRAME_WORLD = 0, FRAME_BODY = 1 };
enum Quad { EULER = 0, RK2 = 1, RK3 = 2, RK4 = 3, IMPLICIT_MIDPOINT = 4 };

struct Model {
    int kind, rot, frame;
    int n, m, nerr;
    // cartpole
    double mc, mp, l, g;
    // rigid bodies
    double mass, J[9], Jinv[9], grav[3], motor_dist, kf, km;
    // double integrator
    int D;
};

static int rot_params(int rot) { return rot == ROT_QUAT ? 4 : 3; }

static void inv3(const double* A, double* Ai) {
    double a = A[0], b = A[1], c = A[2], d = A[3], e = A[4], f = A[5], g = A[6], h = A[7], i = A[8];
    double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    double id = 1.0 / det;
    Ai[0] = (e * i - f * h) * id; Ai[1] = (c * h - b * i) * id; Ai[2] = (b * f - c * e) * id;
    Ai[3] = (f * g - d * i) * id; Ai[4] = (a * i - c * g) * id; Ai[5] = (c * d - a * f) * id;
    Ai[6] = (d * h - e * g) * id; Ai[7] = (b * g - a * h) * id; Ai[8] = (a * e - b * d) * id;
}

// params packing (see include/rdb200.h):
//  CARTPOLE: [mc, mp, l, g]                                  test/cartpole_model.jl:9
//  QUADROTOR: [mass, J(9 row-major), gravity(3), motor_dist, kf, km]   test/quadrotor.jl:36-46
//  BODY: [mass, J(9 row-major)]                              test/rigidbody_test.jl:53-54, examples/single_satellite.jl:31-35
//  DOUBLE_INTEGRATOR: [D]                                    test/double_integrator.jl:97-100
static int make_model(int kind, int rot, int frame, const double* p, int np, Model& M) {
    std::memset(&M, 0, sizeof(M));
    M.kind = kind; M.rot = rot; M.frame = frame;
    switch (kind) {
    case CARTPOLE:
        if (np < 4) return -1;
        M.mc = p[0]; M.mp = p[1]; M.l = p[2]; M.g = p[3];
        M.n = 4; M.m = 1; M.nerr = 4; M.rot = ROT_NONE; break;
    case QUADROTOR:
        if (np < 16 || rot == ROT_NONE) return -1;
        M.mass = p[0]; for (int i = 0; i < 9; ++i) M.J[i] = p[1 + i];
        for (int i = 0; i < 3; ++i) M.grav[i] = p[10 + i];
        M.motor_dist = p[13]; M.kf = p[14]; M.km = p[15];
        inv3(M.J, M.Jinv);
        M.n = 9 + rot_params(rot); M.m = 4; M.nerr = 12; break;
    case BODY:
        if (np < 10 || rot == ROT_NONE) return -1;
        M.mass = p[0]; for (int i = 0; i < 9; ++i) M.J[i] = p[1 + i];
        inv3(M.J, M.Jinv);
        M.n = 9 + rot_params(rot); M.m = 6; M.nerr = 12; break;
    case DOUBLE_INTEGRATOR:
        if (np < 1) return -1;
        M.D = (int)p[0]; if (M.D < 1 || M.D > 3) return -1;
        M.n = 2 * M.D; M.m = M.D; M.nerr = M.n; M.rot = ROT_NONE; break;
    default: return -1;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Rotation arithmetic (Rotations.jl conventions; Hamilton quaternion [w,x,y,z], active rotation).
// ---------------------------------------------------------------------------------------------
template <class S> inline void cross3(const S* a, const S* b, S* c) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
// QuatRotation * r  — NOT normalised (degree-2 polynomial form): (w^2 - v'v) r + 2 v (v'r) + 2 w (v x r).
// The state quaternion is built with renorm=false (src/rigidbody.jl:101-105).
template <class S> inline void quat_rotate(const S* q, const S* r, S* out) {
    S w = q[0]; const S* v = q + 1;
    S vv = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    S vr = v[0] * r[0] + v[1] * r[1] + v[2] * r[2];
    S c[3]; cross3(v, r, c);
    S a = w * w - vv;
    for (int i = 0; i < 3; ++i) out[i] = a * r[i] + 2.0 * (v[i] * vr) + 2.0 * (w * c[i]);
}
// q \ r = inv(q) * r with inv = conjugate (no normalisation).
template <class S> inline void quat_rotate_inv(const S* q, const S* r, S* out) {
    S qc[4] = {q[0], -q[1], -q[2], -q[3]};
    quat_rotate(qc, r, out);
}
// Convert a 3-parameter attitude to the unit quaternion Rotations.jl builds for it.
template <class S> inline void mrp_to_quat(const S* p, S* q) {
    S n2 = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
    S M = 2.0 / (1.0 + n2);
    q[0] = (1.0 - n2) / (1.0 + n2); q[1] = M * p[0]; q[2] = M * p[1]; q[3] = M * p[2];
}
template <class S> inline void rp_to_quat(const S* g, S* q) {
    S M = 1.0 / sqrt_(1.0 + (g[0] * g[0] + g[1] * g[1] + g[2] * g[2]));
    q[0] = M; q[1] = M * g[0]; q[2] = M * g[1]; q[3] = M * g[2];
}
template <class S> inline void rot_rotate(int rot, const S* p, const S* r, S* out) {
    if (rot == ROT_QUAT) { quat_rotate(p, r, out); return; }
    S q[4]; if (rot == ROT_MRP) mrp_to_quat(p, q); else rp_to_quat(p, q);
    quat_rotate(q, r, out);
}
template <class S> inline void rot_rotate_inv(int rot, const S* p, const S* r, S* out) {
    if (rot == ROT_QUAT) { quat_rotate_inv(p, r, out); return; }
    S q[4]; if (rot == ROT_MRP) mrp_to_quat(p, q); else rp_to_quat(p, q);
    quat_rotate_inv(q, r, out);
}
// Rotations.kinematics(R, ω): time derivative of the attitude parameters for body rate ω.
//   quat: ½ q ⊗ [0; ω]  (bilinear, no normalisation)          test/liemodel.jl:13-20
//   MRP : ¼ [(1-|p|²) I + 2 skew(p) + 2 p pᵀ] ω
//   RP  : ½ [I + skew(g) + g gᵀ] ω
template <class S> inline void rot_kinematics(int rot, const S* p, const S* w, S* out) {
    if (rot == ROT_QUAT) {
        S qw = p[0], qx = p[1], qy = p[2], qz = p[3];
        out[0] = 0.5 * (-(qx * w[0]) - qy * w[1] - qz * w[2]);
        out[1] = 0.5 * (qw * w[0] + qy * w[2] - qz * w[1]);
        out[2] = 0.5 * (qw * w[1] + qz * w[0] - qx * w[2]);
        out[3] = 0.5 * (qw * w[2] + qx * w[1] - qy * w[0]);
    } else {
        S n2 = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
        S pw = p[0] * w[0] + p[1] * w[1] + p[2] * w[2];
        S c[3]; cross3(p, w, c);
        if (rot == ROT_MRP) {
            for (int i = 0; i < 3; ++i) out[i] = 0.25 * ((1.0 - n2) * w[i] + 2.0 * c[i] + 2.0 * (p[i] * pw));
        } else {
            for (int i = 0; i < 3; ++i) out[i] = 0.5 * (w[i] + c[i] + p[i] * pw);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Continuous dynamics  xdot = f(x,u,t)
// ---------------------------------------------------------------------------------------------
// Cartpole — test/cartpole_model.jl:11-30.
template <class S> inline void cartpole_dynamics(const Model& M, const S* x, const S* u, S* xd) {
    double mc = M.mc, mp = M.mp, l = M.l, g = M.g;
    S qd0 = x[2], qd1 = x[3];
    S s = sin_(x[1]), c = cos_(x[1]);
    // H = [mc+mp  mp l c; mp l c  mp l^2]
    S H00 = S(mc + mp), H01 = (mp * l) * c, H11 = S(mp * l * l);
    // C*qd = [-mp qd1 l s * qd1, 0];  G = [0, mp g l s];  B u = [u, 0]
    S r0 = -(mp * l) * (qd1 * s) * qd1 - u[0];
    S r1 = (mp * g * l) * s;
    // qdd = -H \ r    (2x2 closed-form solve, StaticArrays)
    S det = H00 * H11 - H01 * H01;
    S qdd0 = -((H11 * r0 - H01 * r1) / det);
    S qdd1 = -((H00 * r1 - H01 * r0) / det);
    xd[0] = qd0; xd[1] = qd1; xd[2] = qdd0; xd[3] = qdd1;
}

// RigidBody — src/rigidbody.jl:213-236 with the wrench models of test/quadrotor.jl:56-96 (QUADROTOR) and
// test/rigidbody_test.jl:26-31 == examples/single_satellite.jl:17-27 (BODY: F = q*u[1:3], M = u[4:6]).
template <class S> inline void rigidbody_dynamics(const Model& M, const S* x, const S* u, S* xd) {
    const int np = rot_params(M.rot);
    const S* r = x; (void)r;
    const S* q = x + 3;
    const S* v = x + 3 + np;
    const S* w = x + 6 + np;
    S F[3], tau[3];
    if (M.kind == QUADROTOR) {
        double kf = M.kf, km = M.km, L = M.motor_dist;
        S F1 = relu_(kf * u[0]), F2 = relu_(kf * u[1]), F3 = relu_(kf * u[2]), F4 = relu_(kf * u[3]);
        S Fb[3] = {S(0.0), S(0.0), F1 + F2 + F3 + F4};
        S qF[3]; rot_rotate(M.rot, q, Fb, qF);
        for (int i = 0; i < 3; ++i) F[i] = (M.mass * M.grav[i]) + qF[i];
        tau[0] = L * (F2 - F4);
        tau[1] = L * (F3 - F1);
        tau[2] = km * u[0] - km * u[1] + km * u[2] - km * u[3];
    } else {  // BODY
        S Fb[3] = {u[0], u[1], u[2]};
        rot_rotate(M.rot, q, Fb, F);
        tau[0] = u[3]; tau[1] = u[4]; tau[2] = u[5];
    }
    S qdot[4]; rot_kinematics(M.rot, q, w, qdot);
    S rdot[3], vdot[3];
    if (M.frame == FRAME_WORLD) {
        for (int i = 0; i < 3; ++i) { rdot[i] = v[i]; vdot[i] = F[i] / M.mass; }
    } else {
        rot_rotate(M.rot, q, v, rdot);
        S Fm[3] = {F[0] / M.mass, F[1] / M.mass, F[2] / M.mass};
        S qiF[3]; rot_rotate_inv(M.rot, q, Fm, qiF);
        S wv[3]; cross3(w, v, wv);
        for (int i = 0; i < 3; ++i) vdot[i] = qiF[i] - wv[i];
    }
    // ωdot = Jinv (τ - ω × (J ω))
    S Jw[3], wJw[3], rhs[3];
    for (int i = 0; i < 3; ++i) Jw[i] = M.J[3 * i] * w[0] + M.J[3 * i + 1] * w[1] + M.J[3 * i + 2] * w[2];
    cross3(w, Jw, wJw);
    for (int i = 0; i < 3; ++i) rhs[i] = tau[i] - wJw[i];
    for (int i = 0; i < 3; ++i) xd[6 + np + i] = M.Jinv[3 * i] * rhs[0] + M.Jinv[3 * i + 1] * rhs[1] + M.Jinv[3 * i + 2] * rhs[2];
    for (int i = 0; i < 3; ++i) xd[i] = rdot[i];
    for (int i = 0; i < np; ++i) xd[3 + i] = qdot[i];
    for (int i = 0; i < 3; ++i) xd[3 + np + i] = vdot[i];
}

// Double integrator — test/double_integrator.jl:101-106.
template <class S> inline void di_dynamics(const Model& M, const S* x, const S* u, S* xd) {
    for (int i = 0; i < M.D; ++i) { xd[i] = x[M.D + i]; xd[M.D + i] = u[i]; }
}

// dynamics(model, x, u, t): all shipped models are time-invariant (src/dynamics.jl:83).
template <class S> inline void dynamics(const Model& M, const S* x, const S* u, double /*t*/, S* xd) {
    switch (M.kind) {
    case CARTPOLE: cartpole_dynamics(M, x, u, xd); break;
    case QUADROTOR: case BODY: rigidbody_dynamics(M, x, u, xd); break;
    default: di_dynamics(M, x, u, xd); break;
    }
}

// ---------------------------------------------------------------------------------------------
// Explicit integrators — src/integration.jl:73-76 (Euler), :130-135 (RK3), :280-286 (RK4);
// RK2 = explicit midpoint (test/old_tests/linear_tests.jl:135-141).
// Written like the reference: k_i = f(...) * h.
// ---------------------------------------------------------------------------------------------
constexpr int NMAX = 13;
template <class S> inline void integrate(const Model& M, int Q, const S* x, const S* u, double t, double h, S* xn) {
    const int n = M.n;
    S k1[NMAX], k2[NMAX], k3[NMAX], k4[NMAX], xt[NMAX];
    dynamics(M, x, u, t, k1);
    for (int i = 0; i < n; ++i) k1[i] = k1[i] * h;
    if (Q == EULER) { for (int i = 0; i < n; ++i) xn[i] = x[i] + k1[i]; return; }
    for (int i = 0; i < n; ++i) xt[i] = x[i] + k1[i] / 2.0;
    dynamics(M, xt, u, t + h / 2, k2);
    for (int i = 0; i < n; ++i) k2[i] = k2[i] * h;
    if (Q == RK2) { for (int i = 0; i < n; ++i) xn[i] = x[i] + k2[i]; return; }
    if (Q == RK3) {
        for (int i = 0; i < n; ++i) xt[i] = x[i] - k1[i] + 2.0 * k2[i];
        dynamics(M, xt, u, t + h, k3);
        for (int i = 0; i < n; ++i) k3[i] = k3[i] * h;
        for (int i = 0; i < n; ++i) xn[i] = x[i] + (k1[i] + 4.0 * k2[i] + k3[i]) / 6.0;
        return;
    }
    for (int i = 0; i < n; ++i) xt[i] = x[i] + k2[i] / 2.0;
    dynamics(M, xt, u, t + h / 2, k3);
    for (int i = 0; i < n; ++i) k3[i] = k3[i] * h;
    for (int i = 0; i < n; ++i) xt[i] = x[i] + k3[i];
    dynamics(M, xt, u, t + h, k4);
    for (int i = 0; i < n; ++i) k4[i] = k4[i] * h;
    for (int i = 0; i < n; ++i) xn[i] = x[i] + (k1[i] + 2.0 * k2[i] + 2.0 * k3[i] + k4[i]) / 6.0;
}

// ---------------------------------------------------------------------------------------------
// Jacobians.
// ---------------------------------------------------------------------------------------------
// method 0: ForwardDiff.jacobian over the whole [x;u]  (src/jacobian_gen.jl:485-507)
template <int NZ> static void jac_ad(const Model& M, int Q /* -1 = continuous */, const double* z, double t, double h, double* J) {
    using D = Dual<NZ>;
    const int n = M.n;
    D zz[NZ];
    for (int i = 0; i < NZ; ++i) { zz[i] = D(z[i]); zz[i].d[i] = 1.0; }
    D out[NMAX];
    if (Q < 0) dynamics(M, zz, zz + n, t, out);
    else integrate(M, Q, zz, zz + n, t, h, out);
    for (int j = 0; j < NZ; ++j)
        for (int i = 0; i < n; ++i) J[i + n * j] = out[i].d[j];
}
static void jac_ad_dispatch(const Model& M, int Q, const double* z, double t, double h, double* J) {
    switch (M.n + M.m) {
    case 3: jac_ad<3>(M, Q, z, t, h, J); break;
    case 5: jac_ad<5>(M, Q, z, t, h, J); break;
    case 6: jac_ad<6>(M, Q, z, t, h, J); break;
    case 9: jac_ad<9>(M, Q, z, t, h, J); break;
    case 16: jac_ad<16>(M, Q, z, t, h, J); break;
    case 17: jac_ad<17>(M, Q, z, t, h, J); break;
    case 18: jac_ad<18>(M, Q, z, t, h, J); break;
    case 19: jac_ad<19>(M, Q, z, t, h, J); break;
    }
}

// Cartpole analytic continuous Jacobian — test/cartpole_model.jl:57-96.
static void cartpole_jacobian(const Model& M, const double* x, const double* u, double* J) {
    double mc = M.mc, mp = M.mp, l = M.l, g = M.g;
    double qd1 = x[3];
    double s = std::sin(x[1]), c = std::cos(x[1]);
    double H00 = mc + mp, H01 = mp * l * c, H11 = mp * l * l;
    double xd[4]; cartpole_dynamics<double>(M, x, u, xd);
    double qdd0 = xd[2], qdd1 = xd[3];
    // rows of (-dH - dC - dG + dB), 2 x 5
    double R[2][5] = {{0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}};
    R[0][1] = -(-mp * l * s * qdd1) - (-mp * l * c * qd1 * qd1);
    R[1][1] = -(-mp * l * s * qdd0) - (mp * g * l * c);
    R[0][3] = -(-2 * mp * l * qd1 * s);
    R[0][4] = 1.0;
    double det = H00 * H11 - H01 * H01;
    for (int i = 0; i < 20; ++i) J[i] = 0;
    J[0 + 4 * 2] = 1.0; J[1 + 4 * 3] = 1.0;
    for (int j = 0; j < 5; ++j) {
        J[2 + 4 * j] = (H11 * R[0][j] - H01 * R[1][j]) / det;
        J[3 + 4 * j] = (H00 * R[1][j] - H01 * R[0][j]) / det;
    }
}

// continuous Jacobian used by the chain-rule path: analytic for Cartpole and the double integrator
// (the reference's UserDefined methods), ForwardAD of `dynamics` for rigid bodies (what @autodiff generates).
static void cont_jacobian(const Model& M, const double* x, const double* u, double t, double* J) {
    const int n = M.n, m = M.m;
    if (M.kind == CARTPOLE) { cartpole_jacobian(M, x, u, J); return; }
    if (M.kind == DOUBLE_INTEGRATOR) {           // test/double_integrator.jl:119-127
        for (int i = 0; i < n * (n + m); ++i) J[i] = 0;
        for (int i = 0; i < n; ++i) J[i + n * (i + M.D)] = 1.0;
        return;
    }
    double z[NMAX + 6];
    for (int i = 0; i < n; ++i) z[i] = x[i];
    for (int i = 0; i < m; ++i) z[n + i] = u[i];
    jac_ad_dispatch(M, -1, z, t, 0.0, J);
}

static void matmul_acc(int n, int p, const double* A /*n x n*/, const double* X /*n x p*/, double alpha, double* Y /*n x p*/) {
    for (int j = 0; j < p; ++j)
        for (int k = 0; k < n; ++k) {
            double xkj = alpha * X[k + n * j];
            for (int i = 0; i < n; ++i) Y[i + n * j] += A[i + n * k] * xkj;
        }
}

// method 1: integrator chain rule — src/integration.jl:85-93 (Euler), :149-177 (RK3), :302-337 (RK4).
static void jac_chain(const Model& M, int Q, const double* z, double t, double h, double* J) {
    const int n = M.n, m = M.m, nn = n * n, nm = n * m;
    const double* x = z; const double* u = z + n;
    double k1[NMAX], k2[NMAX], k3[NMAX], xt[NMAX];
    double Jc[4][NMAX * (NMAX + 6)];
    double dA[4][NMAX * NMAX], dB[4][NMAX * 6];
    auto A = [&](int s) { return Jc[s]; };
    auto B = [&](int s) { return Jc[s] + nn; };
    cont_jacobian(M, x, u, t, Jc[0]);
    for (int i = 0; i < nn; ++i) dA[0][i] = A(0)[i] * h;
    for (int i = 0; i < nm; ++i) dB[0][i] = B(0)[i] * h;
    auto finish = [&](const double* wts, int ns) {
        for (int i = 0; i < nn; ++i) { double s = 0; for (int q = 0; q < ns; ++q) s += wts[q] * dA[q][i]; J[i] = s; }
        for (int i = 0; i < n; ++i) J[i + n * i] += 1.0;
        for (int i = 0; i < nm; ++i) { double s = 0; for (int q = 0; q < ns; ++q) s += wts[q] * dB[q][i]; J[nn + i] = s; }
    };
    // dA_s = A_s (I + X) h,  dB_s = (B_s + A_s Y) h   with X, Y given
    auto stage = [&](int s, const double* X, const double* Y) {
        for (int i = 0; i < nn; ++i) dA[s][i] = A(s)[i];
        matmul_acc(n, n, A(s), X, 1.0, dA[s]);
        for (int i = 0; i < nn; ++i) dA[s][i] *= h;
        for (int i = 0; i < nm; ++i) dB[s][i] = B(s)[i];
        matmul_acc(n, m, A(s), Y, 1.0, dB[s]);
        for (int i = 0; i < nm; ++i) dB[s][i] *= h;
    };
    if (Q == EULER) { double w[1] = {1.0}; finish(w, 1); return; }
    dynamics(M, x, u, t, k1);
    for (int i = 0; i < n; ++i) { k1[i] *= h; xt[i] = x[i] + k1[i] / 2; }
    cont_jacobian(M, xt, u, t + h / 2, Jc[1]);
    double X[NMAX * NMAX], Y[NMAX * 6];
    for (int i = 0; i < nn; ++i) X[i] = 0.5 * dA[0][i];
    for (int i = 0; i < nm; ++i) Y[i] = 0.5 * dB[0][i];
    stage(1, X, Y);
    if (Q == RK2) {   // x+ = x + k2  =>  A = I + dA2, B = dB2
        double w[2] = {0.0, 1.0}; finish(w, 2); return;
    }
    dynamics(M, xt, u, t + h / 2, k2);
    for (int i = 0; i < n; ++i) k2[i] *= h;
    if (Q == RK3) {
        for (int i = 0; i < n; ++i) xt[i] = x[i] - k1[i] + 2 * k2[i];
        cont_jacobian(M, xt, u, t + h, Jc[2]);
        for (int i = 0; i < nn; ++i) X[i] = -dA[0][i] + 2 * dA[1][i];
        for (int i = 0; i < nm; ++i) Y[i] = 2 * dB[1][i] - dB[0][i];
        stage(2, X, Y);
        double w[3] = {1.0 / 6, 4.0 / 6, 1.0 / 6}; finish(w, 3); return;
    }
    for (int i = 0; i < n; ++i) xt[i] = x[i] + k2[i] / 2;
    cont_jacobian(M, xt, u, t + h / 2, Jc[2]);
    for (int i = 0; i < nn; ++i) X[i] = 0.5 * dA[1][i];
    for (int i = 0; i < nm; ++i) Y[i] = 0.5 * dB[1][i];
    stage(2, X, Y);
    dynamics(M, xt, u, t + h / 2, k3);
    for (int i = 0; i < n; ++i) { k3[i] *= h; xt[i] = x[i] + k3[i]; }
    cont_jacobian(M, xt, u, t + h, Jc[3]);
    for (int i = 0; i < nn; ++i) X[i] = dA[2][i];
    for (int i = 0; i < nm; ++i) Y[i] = dB[2][i];
    stage(3, X, Y);
    double w[4] = {1.0 / 6, 2.0 / 6, 2.0 / 6, 1.0 / 6}; finish(w, 4);
}

// ---------------------------------------------------------------------------------------------
// ImplicitMidpoint — src/integration.jl:422-463 (Newton solve), :524-543 (implicit-function-theorem Jacobian),
// :620-694 (residual x1 + h f((x1+x2)/2, u1, t+h/2) - x2 and its Jacobians J1 = [I + h/2 A, h B], J2 = h/2 A - I).
// Newton: x2 <- x1; up to 10 iterations; Jacobians are evaluated BEFORE the convergence test ||r||_2 < 1e-12, so on exit they
// belong to the returned iterate; J = -(J2 \ J1).  LU with partial pivoting like LAPACK getrf (src/utils.jl:32-54).
// ---------------------------------------------------------------------------------------------
static void lu_solve_inplace(int n, double* A /* n x n col-major, destroyed */, int nrhs, double* B /* n x nrhs col-major */) {
    for (int k = 0; k < n; ++k) {
        int p = k; double best = std::fabs(A[k + n * k]);
        for (int i = k + 1; i < n; ++i) if (std::fabs(A[i + n * k]) > best) { best = std::fabs(A[i + n * k]); p = i; }
        if (p != k) {
            for (int j = 0; j < n; ++j) std::swap(A[k + n * j], A[p + n * j]);
            for (int j = 0; j < nrhs; ++j) std::swap(B[k + n * j], B[p + n * j]);
        }
        const double inv = 1.0 / A[k + n * k];
        for (int i = k + 1; i < n; ++i) {
            const double l = A[i + n * k] * inv;
            A[i + n * k] = l;
            for (int j = k + 1; j < n; ++j) A[i + n * j] -= l * A[k + n * j];
            for (int j = 0; j < nrhs; ++j) B[i + n * j] -= l * B[k + n * j];
        }
    }
    for (int j = 0; j < nrhs; ++j)
        for (int i = n - 1; i >= 0; --i) {
            double s = B[i + n * j];
            for (int c = i + 1; c < n; ++c) s -= A[i + n * c] * B[c + n * j];
            B[i + n * j] = s / A[i + n * i];
        }
}

static void implicit_midpoint(const Model& M, const double* z, double t, double h, double* xn, double* J /* may be null */) {
    const int n = M.n, m = M.m, nz = n + m;
    const double* x = z; const double* u = z + n;
    double x2[NMAX], xm[NMAX], f[NMAX], r[NMAX];
    double J1[NMAX * (NMAX + 6)], A2[NMAX * NMAX], W[NMAX * NMAX];
    for (int i = 0; i < n; ++i) x2[i] = x[i];
    for (int iter = 0; iter < 10; ++iter) {
        for (int i = 0; i < n; ++i) xm[i] = (x[i] + x2[i]) / 2;
        dynamics<double>(M, xm, u, t + h / 2, f);
        double nrm = 0;
        for (int i = 0; i < n; ++i) { r[i] = x[i] + h * f[i] - x2[i]; nrm += r[i] * r[i]; }
        cont_jacobian(M, xm, u, t + h / 2, J1);
        for (int i = 0; i < n * nz; ++i) J1[i] *= h;
        for (int i = 0; i < n * n; ++i) { J1[i] /= 2; A2[i] = J1[i]; }
        for (int i = 0; i < n; ++i) { J1[i + n * i] += 1.0; A2[i + n * i] -= 1.0; }
        if (std::sqrt(nrm) < 1e-12) break;
        for (int i = 0; i < n * n; ++i) W[i] = A2[i];
        lu_solve_inplace(n, W, 1, r);
        for (int i = 0; i < n; ++i) x2[i] -= r[i];
    }
    for (int i = 0; i < n; ++i) xn[i] = x2[i];
    if (J) {
        for (int i = 0; i < n * n; ++i) W[i] = A2[i];
        for (int i = 0; i < n * nz; ++i) J[i] = -J1[i];
        lu_solve_inplace(n, W, nz, J);
    }
}

// ---------------------------------------------------------------------------------------------
// LieState error-state maps — src/liestate.jl:210-320 for LieState(R, (3,6)) (src/rigidbody.jl:48),
// Euclidean fall-backs src/statevectortype.jl:144-155.
// ---------------------------------------------------------------------------------------------
static void normalize4(const double* q, double* o) {
    double nrm = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int i = 0; i < 4; ++i) o[i] = q[i] / nrm;
}
static void attitude_to_unit_quat(int rot, const double* p, double* q) {
    if (rot == ROT_QUAT) normalize4(p, q);           // default ctor normalises
    else if (rot == ROT_MRP) mrp_to_quat(p, q);
    else rp_to_quat(p, q);
}
// Rotations.∇differential(R): 4x3 (quat) or 3x3, column-major into D with leading dimension ld.
static void grad_differential(int rot, const double* p, double* D, int ld) {
    if (rot == ROT_QUAT) {
        double q[4]; normalize4(p, q);
        double w = q[0], x = q[1], y = q[2], z = q[3];
        double G[4][3] = {{-x, -y, -z}, {w, -z, y}, {z, w, -x}, {-y, x, w}};   // L(q) H
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 3; ++j) D[i + ld * j] = G[i][j];
    } else {
        double n2 = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
        double sk[3][3] = {{0, -p[2], p[1]}, {p[2], 0, -p[0]}, {-p[1], p[0], 0}};
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
            double I = (i == j) ? 1.0 : 0.0;
            if (rot == ROT_MRP) D[i + ld * j] = (1 - n2) * I + 2 * (sk[i][j] + p[i] * p[j]);
            else D[i + ld * j] = I + sk[i][j] + p[i] * p[j];
        }
    }
}

static void errstate_jacobian(const Model& M, const double* x, double* G /* n x nerr, column-major, fully written */) {
    const int n = M.n, ne = M.nerr;
    for (int i = 0; i < n * ne; ++i) G[i] = 0;
    if (M.rot == ROT_NONE) { for (int i = 0; i < n; ++i) G[i + n * i] = 1; return; }
    const int np = rot_params(M.rot);
    for (int i = 0; i < 3; ++i) G[i + n * i] = 1;
    grad_differential(M.rot, x + 3, G + 3 + n * 3, n);
    for (int i = 0; i < 6; ++i) G[(3 + np + i) + n * (6 + i)] = 1;
}

// ∇errstate_jacobian!: ∇²differential(R, b) on the rotation block, zeros elsewhere (nerr x nerr) — src/liestate.jl:300-320.
// Rotations.jl defines ∇²differential(R, b) = ∇²composition1(R, I, b): the Jacobian with respect to δ, at δ = 0, of
// [∂(R ∘ δ)/∂δ]ᵀ b, i.e. the Hessians of the components of the composition R ∘ δ contracted with b (the second-order term of the
// retraction that Altro adds to Gᵀ ∇²f G).  With δ parameterised like R:
//   quat: q ∘ φ(δ), φ = [1; δ]/sqrt(1 + |δ|²)         ->  -(q·b) I3   (q normalised; pinned by test/liestate.jl:96-99)
//   MRP : p ∘ δ = [(1-|p|²) δ + (1-|δ|²) p + 2 p×δ] / (1 + |p|²|δ|² - 2 p·δ)
//                                                   ->  -2 (1+|p|²)(p·b) I + 2 (a pᵀ + p aᵀ) + 8 (p·b) p pᵀ,   a = (1-|p|²) b - 2 p×b
//   RP  : g ∘ δ = (g + δ + g×δ) / (1 - g·δ)            ->  a gᵀ + g aᵀ + 2 (g·b) g gᵀ,                           a = b - g×b
// (round 2: the MRP / RP rows used to be d/dδ [∇differential(p ∘ δ)ᵀ b], a different object; the present formulas agree with second
// differences of scipy's Rotation composition to 1e-8, tests/test_oracle.py::test_rotation_conventions_vs_scipy.)
static void grad_errstate_jacobian(const Model& M, const double* x, const double* b, double* H) {
    const int ne = M.nerr;
    for (int i = 0; i < ne * ne; ++i) H[i] = 0;
    if (M.rot == ROT_NONE) return;
    const double* p = x + 3; const double* bb = b + 3;
    if (M.rot == ROT_QUAT) {
        double q[4]; normalize4(p, q);
        double d = -(q[0] * bb[0] + q[1] * bb[1] + q[2] * bb[2] + q[3] * bb[3]);
        for (int i = 0; i < 3; ++i) H[(3 + i) + ne * (3 + i)] = d;
        return;
    }
    const double pb = p[0] * bb[0] + p[1] * bb[1] + p[2] * bb[2], n2 = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
    const double pxb[3] = {p[1] * bb[2] - p[2] * bb[1], p[2] * bb[0] - p[0] * bb[2], p[0] * bb[1] - p[1] * bb[0]};
    double a[3];
    for (int i = 0; i < 3; ++i) a[i] = (M.rot == ROT_MRP) ? (1 - n2) * bb[i] - 2 * pxb[i] : bb[i] - pxb[i];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        const double I = (i == j) ? 1.0 : 0.0;
        H[(3 + i) + ne * (3 + j)] = (M.rot == ROT_MRP) ? -2 * (1 + n2) * pb * I + 2 * (a[i] * p[j] + p[i] * a[j]) + 8 * pb * p[i] * p[j]
                                                       : a[i] * p[j] + p[i] * a[j] + 2 * pb * p[i] * p[j];
    }
}

// state_diff(model, x, x0) with the CayleyMap — src/liestate.jl:210-260:  δ = vec(q0⁻¹ ⊗ q) / scalar(q0⁻¹ ⊗ q).
static void state_diff(const Model& M, const double* x, const double* x0, double* dx) {
    const int n = M.n;
    if (M.rot == ROT_NONE) { for (int i = 0; i < n; ++i) dx[i] = x[i] - x0[i]; return; }
    const int np = rot_params(M.rot);
    for (int i = 0; i < 3; ++i) dx[i] = x[i] - x0[i];
    double q[4], q0[4];
    attitude_to_unit_quat(M.rot, x + 3, q);
    attitude_to_unit_quat(M.rot, x0 + 3, q0);
    double c[4] = {q0[0], -q0[1], -q0[2], -q0[3]};
    double e[4];
    e[0] = c[0] * q[0] - c[1] * q[1] - c[2] * q[2] - c[3] * q[3];
    e[1] = c[0] * q[1] + c[1] * q[0] + c[2] * q[3] - c[3] * q[2];
    e[2] = c[0] * q[2] - c[1] * q[3] + c[2] * q[0] + c[3] * q[1];
    e[3] = c[0] * q[3] + c[1] * q[2] - c[2] * q[1] + c[3] * q[0];
    for (int i = 0; i < 3; ++i) dx[3 + i] = e[1 + i] / e[0];
    for (int i = 0; i < 6; ++i) dx[6 + i] = x[3 + np + i] - x0[3 + np + i];
}

}  // namespace rdo

// =============================================================================================
// C entry points (ctypes).  nthreads <= 1: serial; otherwise OpenMP over knot points.
// =============================================================================================
using namespace rdo;

#define RDO_LOOP(N, nthreads) _Pragma("omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)") for (int64_t k = 0; k < (N); ++k)

extern "C" {

int rdo_num_procs() {
#ifdef _OPENMP
    return omp_get_num_procs();
#else
    return 1;
#endif
}

int rdo_dims(int kind, int rot, int frame, const double* params, int np, int* n, int* m, int* nerr) {
    Model M; if (make_model(kind, rot, frame, params, np, M)) return -1;
    *n = M.n; *m = M.m; *nerr = M.nerr; return 0;
}

// t, dt: per-knot arrays (length N) or NULL (then t = 0 / dt = dt0).
int rdo_dynamics(int kind, int rot, int frame, const double* params, int np, int64_t N, const double* Z,
                 const double* t, double* xdot, int nthreads) {
    Model M; if (make_model(kind, rot, frame, params, np, M)) return -1;
    const int nz = M.n + M.m;
    RDO_LOOP(N, nthreads) {
        const double* z = Z + k * nz;
        dynamics<double>(M, z, z + M.n, t ? t[k] : 0.0, xdot + k * M.n);
    }
    return 0;
}

int rdo_discrete_dynamics(int kind, int rot, int frame, const double* params, int np, int Q, int64_t N,
                          const double* Z, const double* t, const double* dt, double dt0, double* xn, int nthreads) {
    Model M; if (make_model(kind, rot, frame, params, np, M)) return -1;
    const int nz = M.n + M.m;
    RDO_LOOP(N, nthreads) {
        const double* z = Z + k * nz;
        if (Q == IMPLICIT_MIDPOINT) implicit_midpoint(M, z, t ? t[k] : 0.0, dt ? dt[k] : dt0, xn + k * M.n, nullptr);
        else integrate<double>(M, Q, z, z + M.n, t ? t[k] : 0.0, dt ? dt[k] : dt0, xn + k * M.n);
    }
    return 0;
}

// Continuous Jacobian  ∂f/∂[x;u]  (method 0: ForwardAD, 1: UserDefined/analytic where the reference has one).
int rdo_jacobian(int kind, int rot, int frame, const double* params, int np, int method, int64_t N,
                 const double* Z, const double* t, double* J, int nthreads) {
    Model M; if (make_model(kind, rot, frame, params, np, M)) return -1;
    const int nz = M.n + M.m, nj = M.n * nz;
    RDO_LOOP(N, nthreads) {
        const double* z = Z + k * nz;
        if (method == 0) jac_ad_dispatch(M, -1, z, t ? t[k] : 0.0, 0.0, J + k * nj);
        else cont_jacobian(M, z, z + M.n, t ? t[k] : 0.0, J + k * nj);
    }
    return 0;
}

// Discrete Jacobian (method 0: ForwardAD through the integrator; 1: chain rule).
int rdo_discrete_jacobian(int kind, int rot, int frame, const double* params, int np, int Q, int method, int64_t N,
                          const double* Z, const double* t, const double* dt, double dt0, double* J, int nthreads) {
    Model M; if (make_model(kind, rot, frame, params, np, M)) return -1;
    const int nz = M.n + M.m, nj = M.n * nz;
    RDO_LOOP(N, nthreads) {
        const double* z = Z + k * nz;
        double tk = t ? t[k] : 0.0, hk = dt ? dt[k] : dt0;
        if (Q == IMPLICIT_MIDPOINT) { double xn[NMAX]; implicit_midpoint(M, z, tk, hk, xn, J + k * nj); }
        else if (method == 0) jac_ad_dispatch(M, Q, z, tk, hk, J + k * nj);
        else jac_chain(M, Q, z, tk, hk, J + k * nj);
    }
    return 0;
}

int rdo_errstate_jacobian(int kind, int rot, int frame, const double* params, int np, int64_t N,
                          const double* X, int ldx, double* G, int nthreads) {
    Model M; if (make_model(kind, rot, frame, params, np, M)) return -1;
    RDO_LOOP(N, nthreads) errstate_jacobian(M, X + k * ldx, G + k * M.n * M.nerr);
    return 0;
}

int rdo_grad_errstate_jacobian(int kind, int rot, int frame, const double* params, int np, int64_t N,
                               const double* X, int ldx, const double* Bv, double* H, int nthreads) {
    Model M; if (make_model(kind, rot, frame, params, np, M)) return -1;
    RDO_LOOP(N, nthreads) grad_errstate_jacobian(M, X + k * ldx, Bv + k * M.n, H + k * M.nerr * M.nerr);
    return 0;
}

int rdo_state_diff(int kind, int rot, int frame, const double* params, int np, int64_t N,
                   const double* X, int ldx, const double* X0, int ldx0, double* dX, int nthreads) {
    Model M; if (make_model(kind, rot, frame, params, np, M)) return -1;
    RDO_LOOP(N, nthreads) state_diff(M, X + k * ldx, X0 + k * ldx0, dX + k * M.nerr);
    return 0;
}

// rollout!: X[:,0] = x0; x_{k+1} = discrete_dynamics(x_k, u_k, t_k, dt_k)  — src/trajectories.jl:436-441.
// X: ntraj x K x n, U: ntraj x (K-1) x m  (knot-major per trajectory), t/dt: ntraj x K or NULL.
int rdo_rollout(int kind, int rot, int frame, const double* params, int np, int Q, int64_t ntraj, int K,
                const double* x0, const double* U, const double* t, const double* dt, double dt0, double* X, int nthreads) {
    Model M; if (make_model(kind, rot, frame, params, np, M)) return -1;
    const int n = M.n, m = M.m;
    RDO_LOOP(ntraj, nthreads) {
        double* Xk = X + k * (int64_t)K * n;
        for (int i = 0; i < n; ++i) Xk[i] = x0[k * n + i];
        for (int j = 0; j + 1 < K; ++j) {
            double tj = t ? t[k * K + j] : 0.0, hj = dt ? dt[k * K + j] : dt0;
            integrate<double>(M, Q, Xk + j * n, U + (k * (int64_t)(K - 1) + j) * m, tj, hj, Xk + (j + 1) * n);
        }
    }
    return 0;
}

}  // extern "C"
