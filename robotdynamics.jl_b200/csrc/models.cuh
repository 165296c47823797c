// models.cuh — continuous dynamics xdot = f(x,u) of the models the reference's hot path is exercised with,
// written once, generically over the scalar kind (plain T, or SD<T,MASK> sparse duals from sdual.cuh).
//
//   Cartpole           reference: test/cartpole_model.jl:9-30           (n=4,  m=1)
//   RigidBody{R}       reference: src/rigidbody.jl:213-236              (n=13 quat / 12 MRP,RP)
//     Quadrotor wrench reference: test/quadrotor.jl:56-96               (m=4)
//     Body/Satellite   reference: test/rigidbody_test.jl:26-31, examples/single_satellite.jl:17-27  (m=6)
//   DoubleIntegrator   reference: test/double_integrator.jl:97-106      (n=2D, m=D)
//
// Rotation arithmetic restates the published formulas of Rotations.jl 1.x (not vendored in the reference;
// SURVEY.md §8c): Hamilton quaternion [w,x,y,z], q*r = (w^2 - v.v) r + 2 v (v.r) + 2 w (v x r) WITHOUT
// normalisation (the state quaternion is built with renorm=false, src/rigidbody.jl:101-105).
#pragma once
#include "sdual.cuh"

namespace rdb {

enum ModelKind { KIND_CARTPOLE = 0, KIND_QUADROTOR = 1, KIND_BODY = 2, KIND_DOUBLE_INTEGRATOR = 3 };
enum RotKind { ROT_NONE = 0, ROT_QUAT = 1, ROT_MRP = 2, ROT_RP = 3 };
enum FrameKind { FRAME_WORLD = 0, FRAME_BODY = 1 };
enum QuadRule { Q_EULER = 0, Q_RK2 = 1, Q_RK3 = 2, Q_RK4 = 3, Q_CONTINUOUS = 4, Q_IMPLICIT_MIDPOINT = 5 };

// Parameter block shared by host and device (passed to kernels by value).
template <class T>
struct ModelParams {
    // cartpole (+ constants derived on the host so that no thread divides: 1/(mp l), (mc+mp)/(mp l), -(mp l)/(mc+mp))
    T mc, mp, l, g;
    T cp_ia, cp_H00, cp_nH00i;
    // rigid bodies
    T mass, inv_mass;
    T J[9], Jinv[9];
    T mg[3];             // mass * gravity
    T motor_dist, kf, km;
};

// ------------------------------------------------------------------------------------------------
// Cartpole
// ------------------------------------------------------------------------------------------------
template <class T>
struct Cartpole {
    static constexpr int n = 4, m = 1, nerr = 4, rot = ROT_NONE;
    ModelParams<T> p;
    // sin/cos of the angle of the first evaluation since reset(): the later RK stages evaluate at angles O(h) away and get
    // their sin/cos from the angle-addition formula (sincos_near) instead of a second full-range sincos.
    mutable T th0, s0, c0;
    mutable bool cached;
    RDB_HD void reset() const { cached = false; }
    // ordering under which df/dx is block lower-triangular (implicit_block.cuh): [theta, w | v | p]
    static constexpr int mp_nblocks = 3;
    RDB_HD static constexpr int mp_size(int b) { return b == 0 ? 2 : 1; }
    RDB_HD static constexpr int mp_idx(int b, int k) { return b == 0 ? (k == 0 ? 1 : 3) : (b == 1 ? 2 : 0); }
    // qdd(theta, w, u) is ONE elemental operation: values on plain scalars, its 2 x 3 local Jacobian by hand (the reference ships the
    // same thing as its UserDefined analytic Jacobian, test/cartpole_model.jl:57-96), partials by chain() — 14 FMAs per stage for the
    // three live columns {theta, w, u} where forward mode through every intermediate of the 2x2 solve spends ~70 (the C2 kernel is
    // co-limited by the FP64 pipe).  Equal to ForwardDiff to rounding (parity tests: 1e-10 against a dense Dual-number restatement of the reference path).
    template <class X, class U>
    RDB_HD auto f(const X& x, const U& u) const {
        const auto& th = get<1>(x);
        const auto& qd0 = get<2>(x);
        const auto& qd1 = get<3>(x);
        const auto& uu = get<0>(u);
        const T thv = val(th), w = val(qd1), uv = val(uu);
#ifdef RDB_TUNE_NULLMODEL   // tuning experiments only: the kernel skeleton with (almost) no arithmetic = the data-movement ceiling
        { const T d0[3] = {T(1), T(0.5), T(2)}, d1[3] = {T(-1), T(0.25), T(3)}; const auto in = vec(th, qd1, uu);
          return vec(qd0, qd1, chain<T>(thv + w, d0, in), chain<T>(uv - w, d1, in)); }
#endif
        T s, c;
        if (!cached) { sincos_(thv, s, c); th0 = thv; s0 = s; c0 = c; cached = true; }
        else sincos_near(thv, th0, s0, c0, s, c);
        // H = [mc+mp  mp l c; mp l c  mp l^2];  C qd + G - B u = [-mp l s w^2 - u, mp g l s];  qdd = -H \ r  (closed-form 2x2 solve, like
        // StaticArrays).  Both sides divided by mp*l:  H' = [H00  c; c  l],  r' = [-s w^2 - u/(mp l), g s],  H00 = (mc+mp)/(mp l).
        const T ia = p.cp_ia, H00 = p.cp_H00, k = p.cp_nH00i, g = p.g;
        // (every multiply-add is an explicit fma_: see sdual.cuh)
        const T w2 = w * w;
        const T r0 = fma_(-ia, uv, -(s * w2));
        const T r1 = g * s;
        const T iD = T(1) / fma_(-c, c, H00 * p.l);
        const T q1 = fma_(c, r0, -(H00 * r1)) * iD;              // qdd1
        const T q0 = k * fma_(c, q1, r0);                        // qdd0: first row of H' qdd = -r'
        using In = decltype(vec(th, qd1, uu));
        if constexpr (!has_partials<T, In>()) return vec(qd0, qd1, q0, q1);
        else {
            // d/dtheta: s' = c, c' = -s;  d/dw: r0' = -2 s w;  d/du: r0' = -1/(mp l)
            const T r0t = -(c * w2);
            const T q1t = fma_(-q1, (T(2) * c) * s, fma_(c, r0t, fma_(-s, r0, -((H00 * g) * c)))) * iD;
            const T q0t = k * fma_(c, q1t, fma_(-s, q1, r0t));
            const T r0w = (T(-2) * s) * w;
            const T ciD = c * iD;
            const T q1w = ciD * r0w;
            const T q0w = k * fma_(c, q1w, r0w);
            const T q1u = -(ia * ciD);
            const T q0u = k * fma_(c, q1u, -ia);
            const T d0[3] = {q0t, q0w, q0u}, d1[3] = {q1t, q1w, q1u};
            const auto in = vec(th, qd1, uu);
            return vec(qd0, qd1, chain<T>(q0, d0, in), chain<T>(q1, d1, in));
        }
    }
};

// ------------------------------------------------------------------------------------------------
// Double integrator
// ------------------------------------------------------------------------------------------------
template <class T, int D>
struct DoubleIntegrator {
    static constexpr int n = 2 * D, m = D, nerr = 2 * D, rot = ROT_NONE;
    ModelParams<T> p;
    RDB_HD void reset() const {}
    static constexpr int mp_nblocks = 2;                       // [v | p]  (implicit_block.cuh)
    RDB_HD static constexpr int mp_size(int) { return D; }
    RDB_HD static constexpr int mp_idx(int b, int k) { return b == 0 ? D + k : k; }
    template <class X, class U>
    RDB_HD auto f(const X& x, const U& u) const { return cat(slice<D, D>(x), u); }
};

// ------------------------------------------------------------------------------------------------
// Rotation helpers (generic over element kinds)
// ------------------------------------------------------------------------------------------------
// q * r, un-normalised polynomial form.  q = Vec<w,x,y,z>, r = Vec<3>.
template <class T, class Q, class R>
RDB_HD auto quat_rotate(const Q& q, const R& r) {
    const auto& w = get<0>(q);
    auto v = slice<1, 3>(q);
    auto a = sq_(w) - norm2_3<T>(v);
    auto vr2 = T(2) * dot3<T>(v, r);
    auto w2 = T(2) * w;
    auto c = cross3<T>(v, r);
    return vec(fmadd<T>(w2, get<0>(c), fmadd<T>(get<0>(v), vr2, a * get<0>(r))),
               fmadd<T>(w2, get<1>(c), fmadd<T>(get<1>(v), vr2, a * get<1>(r))),
               fmadd<T>(w2, get<2>(c), fmadd<T>(get<2>(v), vr2, a * get<2>(r))));
}
// q * [0, 0, s]: s times the third column of the (un-normalised) rotation matrix — 4 squares + 4 + 3 products instead of the
// 9 products the general formula spends on a vector with two structural zeros
template <class T, class Q, class S>
RDB_HD auto quat_rotate_z(const Q& q, const S& s) {
    const auto& w = get<0>(q); const auto& x = get<1>(q); const auto& y = get<2>(q); const auto& z = get<3>(q);
    auto c0 = fmadd<T>(w, y, x * z);                 // half of R[0][2], R[1][2]: the 2 goes onto the scalar s (few partials)
    auto c1 = fmadd<T, -1>(w, x, y * z);
    auto c2 = sqadd<T, -1>(y, sqadd<T, -1>(x, sqadd<T>(z, sq_(w))));
    auto s2 = T(2) * s;
    return vec(c0 * s2, c1 * s2, c2 * s);
}
template <class Q> RDB_HD auto quat_conj(const Q& q) { return vec(get<0>(q), -get<1>(q), -get<2>(q), -get<3>(q)); }

// 3-parameter attitude -> the unit quaternion Rotations.jl builds for it.
template <class T, int ROT, class P>
RDB_HD auto to_quat(const P& p) {
    if constexpr (ROT == ROT_QUAT) return p;
    else if constexpr (ROT == ROT_MRP) {
        auto n2 = norm2_3<T>(p);
        auto i1 = T(1) / (T(1) + n2);
        auto M = T(2) * i1;
        return vec((T(1) - n2) * i1, M * get<0>(p), M * get<1>(p), M * get<2>(p));
    } else {
        auto M = rsqrt_(T(1) + norm2_3<T>(p));
        return vec(M, M * get<0>(p), M * get<1>(p), M * get<2>(p));
    }
}

// c * Rotations.kinematics(R, w): (c times) the time derivative of the attitude parameters for body rate w.  The factor goes
// onto the constants / onto w (3 elements with few partials), never onto the results.
template <class T, int ROT, class P, class W>
RDB_HD auto rot_kinematics(const P& p, const W& w, T c = T(1)) {
    if constexpr (ROT == ROT_QUAT) {   // 1/2 q (x) [0; w], bilinear, no normalisation (reference: test/liemodel.jl:13-20)
        const auto& qw = get<0>(p); const auto& qx = get<1>(p); const auto& qy = get<2>(p); const auto& qz = get<3>(p);
        // the factor 1/2 is applied to w once (3 scalars) rather than to the 4 results (exact: a power of two)
        const T hc = T(0.5) * c;
        auto w0 = hc * get<0>(w); auto w1 = hc * get<1>(w); auto w2 = hc * get<2>(w);
        return vec(fmadd<T, -1>(qz, w2, fmadd<T, -1>(qy, w1, -(qx * w0))),
                   fmadd<T, -1>(qz, w1, fmadd<T>(qy, w2, qw * w0)),
                   fmadd<T, -1>(qx, w2, fmadd<T>(qz, w0, qw * w1)),
                   fmadd<T, -1>(qy, w0, fmadd<T>(qx, w1, qw * w2)));
    } else {
        auto pw = dot3<T>(p, w);
        auto cr = cross3<T>(p, w);
        if constexpr (ROT == ROT_MRP) {  // 1/4 [(1-|p|^2) I + 2 skew(p) + 2 p p'] w
            auto a = T(1) - norm2_3<T>(p);
            const T qc = T(0.25) * c;
            return vec(qc * (a * get<0>(w) + T(2) * (get<0>(cr) + get<0>(p) * pw)),
                       qc * (a * get<1>(w) + T(2) * (get<1>(cr) + get<1>(p) * pw)),
                       qc * (a * get<2>(w) + T(2) * (get<2>(cr) + get<2>(p) * pw)));
        } else {                          // 1/2 [I + skew(g) + g g'] w
            const T hc = T(0.5) * c;
            return vec(hc * (get<0>(w) + get<0>(cr) + get<0>(p) * pw),
                       hc * (get<1>(w) + get<1>(cr) + get<1>(p) * pw),
                       hc * (get<2>(w) + get<2>(cr) + get<2>(p) * pw));
        }
    }
}

// ---- rotations as elemental operations ---------------------------------------------------------------------------------------------
// AttRot holds the VALUES of the quaternion an attitude stands for (the state quaternion itself, un-normalised, for QuatRotation;
// the unit quaternion Rotations.jl builds for an MRP / Rodrigues vector) and, for the 3-parameter attitudes, dq/dp.  rot<INV>(r)
// evaluates  y = R(q) r  (INV: R(q)' r = q \ r)  on plain scalars, writes the 3 x (np + 3) local Jacobian by hand and lets chain()
// build the partials: 7 (6) FMAs per output and partial, where forward mode through the polynomial  (w^2 - v.v) r + 2 v (v.r) +
// 2 w (v x r)  spends ~14 — and ~18 more through to_quat for the 3-parameter attitudes.  The rigid-body kernels are bound by FP32 /
// FP64 instruction issue, so this is where the body-frame and MRP / Rodrigues models lost their time in round 1
// (0.31 - 0.52 of the HBM roofline).  Same map as quat_rotate / to_quat, hence the same derivatives (up to rounding).
template <class T, class X> RDB_HD T pval(const X& x) { return opnd<T, X>::v(x); }
template <class T, int ROT, class Att>
struct AttRot {
    static constexpr int np = (ROT == ROT_QUAT) ? 4 : 3;
    const Att& att;
    T w, x, y, z;
    T dq[4][3];                      // d(w, x, y, z) / dp   (3-parameter attitudes only)
    RDB_HD explicit AttRot(const Att& a) : att(a) {
        if constexpr (ROT == ROT_QUAT) {
            w = pval<T>(get<0>(a)); x = pval<T>(get<1>(a)); y = pval<T>(get<2>(a)); z = pval<T>(get<3>(a));
        } else {
            const T p[3] = {pval<T>(get<0>(a)), pval<T>(get<1>(a)), pval<T>(get<2>(a))};
            const T n2 = fma_(p[2], p[2], fma_(p[1], p[1], p[0] * p[0]));
            if constexpr (ROT == ROT_MRP) {          // q = ((1 - |p|^2), 2 p) / (1 + |p|^2)
                const T i1 = T(1) / (T(1) + n2), M = T(2) * i1;
                w = (T(1) - n2) * i1; x = M * p[0]; y = M * p[1]; z = M * p[2];
                const T qv[3] = {x, y, z};
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const T g = -(M * p[j]);             // -2 p_j / (1 + |p|^2)
                    dq[0][j] = g * (T(1) + w);
#pragma unroll
                    for (int i = 0; i < 3; ++i) dq[1 + i][j] = (i == j) ? fma_(g, qv[i], M) : g * qv[i];
                }
            } else {                                  // q = (1, g) / sqrt(1 + |g|^2)
                const T M = rsqrt_(T(1) + n2), M3 = M * M * M;
                w = M; x = M * p[0]; y = M * p[1]; z = M * p[2];
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const T g = -(p[j] * M3);
                    dq[0][j] = g;
#pragma unroll
                    for (int i = 0; i < 3; ++i) dq[1 + i][j] = (i == j) ? fma_(g, p[i], M) : g * p[i];
                }
            }
        }
    }
    template <bool INV, class R3>
    RDB_HD auto rot(const R3& r) const {
        const T v[3] = {x, y, z};
        const T rr[3] = {pval<T>(get<0>(r)), pval<T>(get<1>(r)), pval<T>(get<2>(r))};
        const T sw = INV ? -w : w;                    // conj(q) = (w, -v): only the cross-product term changes sign
        const T a = fma_(w, w, -fma_(z, z, fma_(y, y, x * x)));
        const T d = fma_(z, rr[2], fma_(y, rr[1], x * rr[0]));
        const T c[3] = {fma_(y, rr[2], -(z * rr[1])), fma_(z, rr[0], -(x * rr[2])), fma_(x, rr[1], -(y * rr[0]))};
        const T sw2 = T(2) * sw, d2 = T(2) * d;
        T yv[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) yv[i] = fma_(sw2, c[i], fma_(d2, v[i], a * rr[i]));
        using In = decltype(cat(att, r));
        if constexpr (!has_partials<T, In>()) return vec(yv[0], yv[1], yv[2]);
        else {
            // (e_j x r)_i and (v x e_j)_i
            const T exr[3][3] = {{T(0), rr[2], -rr[1]}, {-rr[2], T(0), rr[0]}, {rr[1], -rr[0], T(0)}};      // [i][j] = (e_j x r)_i
            const T vxe[3][3] = {{T(0), -v[2], v[1]}, {v[2], T(0), -v[0]}, {-v[1], v[0], T(0)}};            // [i][j] = (v x e_j)_i
            T cf[3][np + 3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const T cw = T(2) * fma_(w, rr[i], INV ? -c[i] : c[i]);
                T cv[3];
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const T sym = fma_(rr[j], v[i], fma_(-v[j], rr[i], i == j ? d : T(0)));
                    cv[j] = (i == j) ? T(2) * sym : fma_(sw2, exr[i][j], T(2) * sym);         // (e_i x r)_i = 0
                }
                if constexpr (ROT == ROT_QUAT) {
                    cf[i][0] = cw; cf[i][1] = cv[0]; cf[i][2] = cv[1]; cf[i][3] = cv[2];
                } else {
#pragma unroll
                    for (int j = 0; j < 3; ++j) cf[i][j] = fma_(cv[2], dq[3][j], fma_(cv[1], dq[2][j], fma_(cv[0], dq[1][j], cw * dq[0][j])));
                }
#pragma unroll
                for (int j = 0; j < 3; ++j) cf[i][np + j] = (i == j) ? fma_(T(2) * v[j], v[i], a) : fma_(sw2, vxe[i][j], (T(2) * v[j]) * v[i]);
            }
            const auto in = cat(att, r);
            return vec(chain<T>(yv[0], cf[0], in), chain<T>(yv[1], cf[1], in), chain<T>(yv[2], cf[2], in));
        }
    }
};

// c * Rotations.kinematics for the 3-parameter attitudes as ONE elemental operation (3 x 6 local Jacobian; the quaternion form is
// bilinear and already costs what its local Jacobian would).
template <class T, int ROT, class P, class W>
RDB_HD auto rot_kinematics_el(const P& pp, const W& ww, T c) {
    static_assert(ROT == ROT_MRP || ROT == ROT_RP, "3-parameter attitudes");
    const T p[3] = {pval<T>(get<0>(pp)), pval<T>(get<1>(pp)), pval<T>(get<2>(pp))};
    const T w[3] = {pval<T>(get<0>(ww)), pval<T>(get<1>(ww)), pval<T>(get<2>(ww))};
    const T pw = fma_(p[2], w[2], fma_(p[1], w[1], p[0] * w[0]));
    const T cr[3] = {fma_(p[1], w[2], -(p[2] * w[1])), fma_(p[2], w[0], -(p[0] * w[2])), fma_(p[0], w[1], -(p[1] * w[0]))};
    const T n2 = fma_(p[2], p[2], fma_(p[1], p[1], p[0] * p[0]));
    // MRP: 1/4 [(1-|p|^2) w + 2 p x w + 2 p (p.w)];  RP: 1/2 [w + g x w + g (g.w)]  ==  k [a w + b (p x w + p (p.w))]
    const T k = (ROT == ROT_MRP ? T(0.25) : T(0.5)) * c, a = ROT == ROT_MRP ? T(1) - n2 : T(1), b = ROT == ROT_MRP ? T(2) : T(1);
    const T kb = k * b, ka = k * a;
    T yv[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) yv[i] = fma_(kb, fma_(p[i], pw, cr[i]), ka * w[i]);
    using In = decltype(cat(pp, ww));
    if constexpr (!has_partials<T, In>()) return vec(yv[0], yv[1], yv[2]);
    else {
        const T exw[3][3] = {{T(0), w[2], -w[1]}, {-w[2], T(0), w[0]}, {w[1], -w[0], T(0)}};      // [i][j] = (e_j x w)_i
        const T pxe[3][3] = {{T(0), -p[2], p[1]}, {p[2], T(0), -p[0]}, {-p[1], p[0], T(0)}};      // [i][j] = (p x e_j)_i
        T cf[3][6];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                // d/dp_j: k [da/dp_j w_i + b ((e_j x w)_i + delta_ij (p.w) + p_i w_j)],  da/dp_j = -2 p_j (MRP) / 0 (RP)
                const T inner = fma_(p[i], w[j], i == j ? pw : exw[i][j]);                   // (e_i x w)_i = 0
                cf[i][j] = (ROT == ROT_MRP) ? fma_((-(T(2) * k)) * p[j], w[i], kb * inner) : kb * inner;
                // d/dw_j: k [a delta_ij + b ((p x e_j)_i + p_i p_j)]
                const T sym = (i == j) ? p[i] * p[j] : fma_(p[i], p[j], pxe[i][j]);
                cf[i][3 + j] = (i == j) ? fma_(kb, sym, ka) : kb * sym;
            }
        const auto in = cat(pp, ww);
        return vec(chain<T>(yv[0], cf[0], in), chain<T>(yv[1], cf[1], in), chain<T>(yv[2], cf[2], in));
    }
}

// y = diag(d) x
template <class T, class A> RDB_HD auto diag3_mul(T d0, T d1, T d2, const A& x) { return vec(d0 * get<0>(x), d1 * get<1>(x), d2 * get<2>(x)); }

// ------------------------------------------------------------------------------------------------
// RigidBody{R} with the Quadrotor or Body/Satellite wrench
// ------------------------------------------------------------------------------------------------
// The part every RigidBody{R} shares (reference: src/rigidbody.jl:213-236): kinematics, Newton and Euler equations around a
// wrench supplied by the concrete model — `wrench(R, q, r, v, w, u)` (R: AttRot, the attitude as an elemental rotation; q: the same
// attitude as a quaternion of duals, for user code) returns vec(F/m in the world frame (3), tau in the body frame (3))
// (reference: forces / moments / wrenches, src/rigidbody.jl:244-257).
//
// SCALED: return c * f(x,u) instead of f(x,u) (the integrators ask for the stage increment h a_s f(X_s) directly, integrators.cuh).
// The factor is folded into constants and into elements with few partials — the inverse inertia, w before the kinematics, v before
// a body-frame rotation — so that almost no partial is multiplied by it; the WRENCH MUST ALREADY RETURN c * F/m (and tau unscaled).
//
// SPLIT (body-frame built-in models): the wrench returns vec(Fb (3), Gw (3), tau (3)) with F/m = q * Fb' + Gw — a body-frame part and a
// world-frame part (plain constants or Zero) — where Fb = |q|^4 Fb' already.  The reference evaluates q \ (q * Fb' + Gw)
// (test/quadrotor.jl:74 `m*g + q*F`, then src/rigidbody.jl:229-230 `q \ (F ./ m)`); for the un-normalised polynomial rotation
// R(q)' R(q) = |q|^4 I identically (R(q) = |q|^2 x a rotation), so q \ (q * Fb') == |q|^4 Fb' as polynomials in q: same value and same
// derivatives up to rounding, without pushing ~17 columns of partials through two rotations.  (Unit quaternions built from MRP /
// Rodrigues vectors have |q| = 1 by construction: the factor is 1 and its derivative 0.)
template <class T, int ROT, int FRAME, bool DIAG_INERTIA, bool SCALED = false, bool SPLIT = false, class X, class U, class Wrench>
RDB_HD auto rigid_body_f(const ModelParams<T>& p, const X& x, const U& u, const Wrench& wrench, T c = T(1)) {
    static_assert(!SPLIT || FRAME == FRAME_BODY, "the split force is a body-frame form");
    constexpr int np = (ROT == ROT_QUAT) ? 4 : 3;
    auto r = slice<0, 3>(x);
    auto att = slice<3, np>(x);
    auto v = slice<3 + np, 3>(x);
    auto w = slice<6 + np, 3>(x);
    auto q = to_quat<T, ROT>(att);          // identity for quaternions (never renormalised); dead code unless the wrench reads it
    const AttRot<T, ROT, decltype(att)> R(att);
    auto xi = wrench(R, q, r, v, w, u);
    auto Fm = slice<0, 3>(xi);
    auto tau = slice<(SPLIT ? 6 : 3), 3>(xi);
    auto qdot = [&]() {
        if constexpr (ROT == ROT_QUAT) { if constexpr (SCALED) return rot_kinematics<T, ROT>(att, w, c); else return rot_kinematics<T, ROT>(att, w); }
        else return rot_kinematics_el<T, ROT>(att, w, SCALED ? c : T(1));
    }();
    // omega_dot = Jinv (tau - w x (J w))
    auto wdot = [&]() {
        if constexpr (DIAG_INERTIA) {      // w x (J w) = ((J3-J2) wy wz, (J1-J3) wz wx, (J2-J1) wx wy): three products instead of six
            const auto& wx = get<0>(w); const auto& wy = get<1>(w); const auto& wz = get<2>(w);
            const T J1 = p.J[0], J2 = p.J[4], J3 = p.J[8];
            const T i0 = SCALED ? c * p.Jinv[0] : p.Jinv[0], i1 = SCALED ? c * p.Jinv[4] : p.Jinv[4], i2 = SCALED ? c * p.Jinv[8] : p.Jinv[8];
            return vec(i0 * fmadd<T, -1>((J3 - J2) * wy, wz, get<0>(tau)),
                       i1 * fmadd<T, -1>((J1 - J3) * wz, wx, get<1>(tau)),
                       i2 * fmadd<T, -1>((J2 - J1) * wx, wy, get<2>(tau)));
        } else {
            if constexpr (SCALED) {
                T Ji[9];
#pragma unroll
                for (int i = 0; i < 9; ++i) Ji[i] = c * p.Jinv[i];
                return mat3_mul(Ji, vsub(tau, cross3<T>(w, mat3_mul(p.J, w))));
            } else {
                return mat3_mul(p.Jinv, vsub(tau, cross3<T>(w, mat3_mul(p.J, w))));
            }
        }
    }();
    if constexpr (FRAME == FRAME_WORLD) {
        if constexpr (SCALED) return cat(vscale(c, v), qdot, Fm, wdot);
        else return cat(v, qdot, Fm, wdot);
    } else {
        // q \ (F/m): the generic form rotates the world-frame force back; the split form only its world-frame part (if any)
        auto qF = [&]() {
            if constexpr (SPLIT) {
                using G0 = rstd::remove_cv_t<rstd::remove_reference_t<decltype(get<3>(xi))>>;
                if constexpr (rstd::is_same<G0, Zero>::value) return Fm;
                else return vadd(Fm, R.template rot<true>(slice<3, 3>(xi)));
            } else return R.template rot<true>(Fm);
        }();
        if constexpr (SCALED) {            // c (q*v) = q*(c v),  c (q\F/m - w x v) = q\(c F/m) - w x (c v)
            auto cv = vscale(c, v);
            return cat(R.template rot<false>(cv), qdot, vsub(qF, cross3<T>(w, cv)), wdot);
        } else {
            auto rdot = R.template rot<false>(v);
            auto vdot = vsub(qF, cross3<T>(w, v));
            return cat(rdot, qdot, vdot, wdot);
        }
    }
}

template <class T, int KIND, int ROT, int FRAME>
struct RigidBody {
    static constexpr int np = (ROT == ROT_QUAT) ? 4 : 3;
    static constexpr int n = 9 + np, m = (KIND == KIND_QUADROTOR) ? 4 : 6, nerr = 12, rot = ROT, frame = FRAME;
    // ordering under which df/dx is block lower-triangular (implicit_block.cuh): [w | att | v | r]  (src/rigidbody.jl:213-236: w' reads
    // w and u, att' reads att and w, v' reads att, v, w, u, r' reads att and v)
    static constexpr int mp_nblocks = 4;
    RDB_HD static constexpr int mp_size(int b) { return b == 1 ? np : 3; }
    RDB_HD static constexpr int mp_idx(int b, int k) { return b == 0 ? 6 + np + k : b == 1 ? 3 + k : b == 2 ? 3 + np + k : k; }
    // The reference Quadrotor stores its inertia as Diagonal{Float64} (test/quadrotor.jl:25-26): only the diagonal exists.
    // Body/Satellite carry a full SMatrix{3,3} (test/rigidbody_test.jl:24, examples/single_satellite.jl:9).
    static constexpr bool diag_inertia = (KIND == KIND_QUADROTOR);
    ModelParams<T> p;
    RDB_HD void reset() const {}

    // c * f(x,u) with the factor folded into the model's constants (integrators.cuh: stage increments)
    static constexpr bool folds_scale = true;
    template <class X, class U> RDB_HD auto f(const X& x, const U& u) const { return eval<false>(x, u, T(1)); }
    template <class X, class U> RDB_HD auto fs(const X& x, const U& u, T c) const { return eval<true>(x, u, c); }

    template <bool SCALED, class X, class U>
    RDB_HD auto eval(const X& x, const U& u, T c) const {
        // wrench: F/m in the world frame (the 1/m of vdot = F/m — and the stage factor c — are folded into the few scalars that
        // build F, instead of scaling every partial of the rotated vector), tau in the body frame
        const T im = SCALED ? c * p.inv_mass : p.inv_mass;
        constexpr bool SPLIT = (FRAME == FRAME_BODY);
        return rigid_body_f<T, ROT, FRAME, diag_inertia, SCALED, SPLIT>(p, x, u, [&](const auto& R, const auto& q, const auto&, const auto&, const auto&, const auto& uu) {
            // |q|^4 for the state quaternion (rigid_body_f: SPLIT); 1 for the unit quaternions of the 3-parameter attitudes
            auto s4 = [&]() {
                if constexpr (SPLIT && ROT == ROT_QUAT) return sq_(sqadd<T>(get<3>(q), sqadd<T>(get<2>(q), sqadd<T>(get<1>(q), sq_(get<0>(q))))));
                else return T(1);
            }();
            if constexpr (KIND == KIND_QUADROTOR) {
                auto F1 = relu_(p.kf * get<0>(uu));
                auto F2 = relu_(p.kf * get<1>(uu));
                auto F3 = relu_(p.kf * get<2>(uu));
                auto F4 = relu_(p.kf * get<3>(uu));
                auto tau = vec(p.motor_dist * (F2 - F4), p.motor_dist * (F3 - F1),
                               p.km * (get<0>(uu) - get<1>(uu) + get<2>(uu) - get<3>(uu)));
                const T g0 = p.mg[0] * im, g1 = p.mg[1] * im, g2 = p.mg[2] * im;
                if constexpr (SPLIT) {      // thrust stays in the body frame, gravity is the world-frame part
                    auto Fz = im * (F1 + F2 + F3 + F4);
                    if constexpr (ROT == ROT_QUAT) return cat(vec(Zero{}, Zero{}, s4 * Fz), vec(g0, g1, g2), tau);
                    else return cat(vec(Zero{}, Zero{}, Fz), vec(g0, g1, g2), tau);
                } else {
                    // thrust along the body z axis: the third column of R(q) for the state quaternion (quat_rotate_z: 18 products per partial),
                    // the elemental rotation of [0, 0, s] for the 3-parameter attitudes (12 FMAs per partial instead of to_quat + rotate)
                    auto qF = [&]() {
                        if constexpr (ROT == ROT_QUAT) return quat_rotate_z<T>(q, im * (F1 + F2 + F3 + F4));
                        else return R.template rot<false>(vec(Zero{}, Zero{}, im * (F1 + F2 + F3 + F4)));
                    }();
                    auto Fm = vec(g0 + get<0>(qF), g1 + get<1>(qF), g2 + get<2>(qF));
                    return cat(Fm, tau);
                }
            } else {
                if constexpr (SPLIT) {      // the force input is a body-frame vector and there is no world-frame part
                    auto Fb = vscale(im, slice<0, 3>(uu));
                    if constexpr (ROT == ROT_QUAT) return cat(vscale(s4, Fb), vec(Zero{}, Zero{}, Zero{}), slice<3, 3>(uu));
                    else return cat(Fb, vec(Zero{}, Zero{}, Zero{}), slice<3, 3>(uu));
                } else return cat(R.template rot<false>(vscale(im, slice<0, 3>(uu))), slice<3, 3>(uu));
            }
        }, c);
    }
};

}  // namespace rdb
