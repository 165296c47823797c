// integrators.cuh — explicit Runge-Kutta steps x+ = integrate(Q, f, x, u, h), generic over the scalar kind.
//
//   Euler  reference: src/integration.jl:73-76        x + h f(x,u)
//   RK2    explicit midpoint (not in src/ at v0.4.8; pinned by test/old_tests/linear_tests.jl:135-141)
//   RK3    reference: src/integration.jl:130-135      k3 at x - k1 + 2 k2, weights (1,4,1)/6
//   RK4    reference: src/integration.jl:280-286      classic, weights (1,2,2,1)/6
//
// The reference forms k_i = f(.)*h and sums them; here h is folded into the stage coefficients
// (x + (h/2) f1 instead of x + (f1*h)/2), which is the same map up to rounding (~1 ulp) and saves one
// multiply per partial when the scalars are sparse duals.  Called with seeded duals this IS the discrete
// Jacobian (forward mode through the integrator, reference: src/jacobian_gen.jl:485-507), and it equals the
// chain rule of src/integration.jl:302-337 to rounding.  Time: stage s is evaluated at t + c_s h exactly as the reference does
// (RK4: t, t+h/2, t+h/2, t+h, src/integration.jl:281-284; RK3: t, t+h/2, t+h, :131-133; midpoint rules: t, t+h/2); t is a plain
// scalar (never differentiated) and reaches only models that declare `time_varying` (user models, custom.cu) — every shipped model
// is time-invariant (src/dynamics.jl:83) and the stage times are dead code for it.
#pragma once
#include "models.cuh"

namespace rdb {

// dynamics(model, x, u, t) (reference: src/dynamics.jl:81-83): models that declare `static constexpr bool time_varying = true`
// provide f(x, u, t) / fs(x, u, c, t); the others f(x, u) / fs(x, u, c)
template <class Model, class = void> struct uses_time : rstd::false_type {};
template <class Model> struct uses_time<Model, rstd::void_t<decltype(Model::time_varying)>> : rstd::integral_constant<bool, Model::time_varying> {};
template <class T, class Model, class X, class U>
RDB_HD auto feval(const Model& model, const X& x, const U& u, T t) {
    if constexpr (uses_time<Model>::value) return model.f(x, u, t); else return model.f(x, u);
}
template <class T, class Model, class X, class U>
RDB_HD auto fseval(const Model& model, const X& x, const U& u, T c, T t) {
    if constexpr (uses_time<Model>::value) return model.fs(x, u, c, t); else return model.fs(x, u, c);
}

// ---- saturated stage types -----------------------------------------------------------------------------------------
// The partial masks of the stage points grow from stage to stage (fill-in) until they reach a fixed point XS with
// type(x + c f(XS, u)) == XS.  Carrying XS through a ROLLED loop over the stages makes every stage execute the same
// code: the kernel body shrinks ~3x (a fully unrolled RK4 Jacobian of a rigid body is ~85 KB of straight-line SASS and
// stalls on instruction fetch), at the price of explicit zeros in the early stages.
template <class Model, class T, class X0, class U, class X>
struct saturate_impl {
    using F = decltype(feval<T>(rstd::declval<const Model&>(), rstd::declval<const X&>(), rstd::declval<const U&>(), rstd::declval<T>()));
    using Next = decltype(axpy(rstd::declval<const X0&>(), rstd::declval<T>(), rstd::declval<const F&>()));
    using type = typename rstd::conditional_t<rstd::is_same<Next, X>::value, ident<X>, saturate_impl<Model, T, X0, U, Next>>::type;
};
template <class Model, class T, class X0, class U> using saturated_t = typename saturate_impl<Model, T, X0, U, X0>::type;

// ---- stage increments ------------------------------------------------------------------------------------------------------
// Models that can fold a scalar factor into their constants (Model::folds_scale, e.g. RigidBody::fs) are asked for the stage
// INCREMENT G_s = (h a_s) f(X_s) directly: the next stage point is then x + G_s — no multiplication of a partial at all, where
// x + (h a_s) f_s costs one multiply per partial — and x+ = x + sum_s (b_s / a_s) G_s.  The kernels are bound by FP32 instruction
// issue, so the ~130 multiplies per stage of a rigid body count.  Models without fs() keep the
// classical form below (for Cartpole the two forms cost the same).
#ifndef RDB_INCREMENT_FORM
#define RDB_INCREMENT_FORM 1     // 0: never; 1: straight-line rules only; 2: rolled stage loops too (tuning experiments)
#endif
template <class Model, class = void> struct folds_scale : rstd::false_type {};
template <class Model> struct folds_scale<Model, rstd::void_t<decltype(Model::folds_scale)>> : rstd::integral_constant<bool, Model::folds_scale> {};

// RK4 in increment form:  G1 = h/2 f(x), G2 = h/2 f(x+G1), G3 = h f(x+G2), G4 = h f(x+G3);  x+ = x + 1/3 (G1 + 2 G2 + G3 + 1/2 G4).
// ROLL == 0: unrolled; 1: stage 1 on the sparse seed types, stages 2..4 rolled over the saturated types; 2: all four rolled.
template <int ROLL, class T, class Model, class X, class U>
RDB_HD auto rk4_increments(const Model& model, const X& x, const U& u, T h, T t) {
    const T hh = T(0.5) * h;
    if constexpr (ROLL == 0) {
        auto g1 = fseval<T>(model, x, u, hh, t);
        auto g2 = fseval<T>(model, vadd(x, g1), u, hh, t + hh);
        auto g3 = fseval<T>(model, vadd(x, g2), u, h, t + hh);
        auto g4 = fseval<T>(model, vadd(x, g3), u, h, t + h);
        return axpy(x, T(1.0 / 3.0), axpy(vadd(axpy(g1, T(2), g2), g3), T(0.5), g4));
    } else {
        using XS = saturated_t<Model, T, X, U>;
        using FS = decltype(fseval<T>(model, rstd::declval<const XS&>(), u, h, t));
        XS Xs;
        FS acc;
        int s0;
        if constexpr (ROLL == 1) {
            auto g1 = fseval<T>(model, x, u, hh, t);               // stage 1 on the sparse seed types
            Xs = widen_vec<XS>(vadd(x, g1));
            acc = widen_vec<FS>(g1);
            s0 = 1;
        } else {
            Xs = widen_vec<XS>(x);
            acc = zero_vec<FS, T>();
            s0 = 0;
        }
#pragma unroll 1
        for (int s = s0; s < 4; ++s) {
            const FS g = fseval<T>(model, Xs, u, s < 2 ? hh : h, t + (s == 0 ? T(0) : (s < 3 ? hh : h)));
            const T w = (s == 1) ? T(2) : (s == 3 ? T(0.5) : T(1));
            acc = axpy(acc, w, g);
            if (s < 3) Xs = widen_vec<XS>(vadd(x, g));
        }
        return axpy(x, T(1.0 / 3.0), acc);
    }
}

// RK3 in increment form:  G1 = h/2 f(x), G2 = 2h f(x+G1), G3 = h f(x - 2 G1 + G2);  x+ = x + 1/3 (G1 + G2 + 1/2 G3).
template <int ROLL, class T, class Model, class X, class U>
RDB_HD auto rk3_increments(const Model& model, const X& x, const U& u, T h, T t) {
    const auto g1 = fseval<T>(model, x, u, T(0.5) * h, t);
    const auto P = axpy(x, T(-2), g1);                              // x - h f1, stays in the sparse stage-1 types
    if constexpr (ROLL == 0) {
        auto g2 = fseval<T>(model, vadd(x, g1), u, T(2) * h, t + T(0.5) * h);
        auto g3 = fseval<T>(model, vadd(P, g2), u, h, t + h);
        return axpy(x, T(1.0 / 3.0), axpy(vadd(g1, g2), T(0.5), g3));
    } else {
        using XS = saturated_t<Model, T, X, U>;
        using FS = decltype(fseval<T>(model, rstd::declval<const XS&>(), u, h, t));
        XS Xs = widen_vec<XS>(vadd(x, g1));
        FS acc = widen_vec<FS>(g1);
#pragma unroll 1
        for (int s = 0; s < 2; ++s) {
            const FS g = fseval<T>(model, Xs, u, s == 0 ? T(2) * h : h, t + (s == 0 ? T(0.5) * h : h));
            acc = axpy(acc, s == 0 ? T(1) : T(0.5), g);
            if (s == 0) Xs = widen_vec<XS>(vadd(P, g));
        }
        return axpy(x, T(1.0 / 3.0), acc);
    }
}

// RK4 with stages 2..4 (ROLL == 1) or all four stages (ROLL == 2) rolled over the saturated types.
template <int ROLL, class T, class Model, class X, class U>
RDB_HD auto rk4_rolled(const Model& model, const X& x, const U& u, T h, T t) {
    using XS = saturated_t<Model, T, X, U>;
    using FS = decltype(feval<T>(model, rstd::declval<const XS&>(), u, t));
    XS Xs;
    FS acc;
    int s0;
    if constexpr (ROLL == 1) {
        auto f1 = feval<T>(model, x, u, t);                    // stage 1 on the sparse seed types
        Xs = widen_vec<XS>(axpy(x, T(0.5) * h, f1));
        acc = widen_vec<FS>(f1);
        s0 = 1;
    } else {
        Xs = widen_vec<XS>(x);
        acc = zero_vec<FS, T>();
        s0 = 0;
    }
#pragma unroll 1
    for (int s = s0; s < 4; ++s) {
        const FS f = feval<T>(model, Xs, u, t + (s == 0 ? T(0) : (s < 3 ? T(0.5) * h : h)));
        const T w = (s == 1 || s == 2) ? T(2) : T(1);
        acc = axpy(acc, w, f);
        const T c = (s == 2) ? h : T(0.5) * h;                 // next stage point: x + h/2 f1, x + h/2 f2, x + h f3
        if (s < 3) Xs = axpy(x, c, f);
    }
    return axpy(x, h * T(1.0 / 6.0), acc);
}

// RK3 (k3 at x - k1 + 2 k2, weights (1,4,1)/6) with stages 2 and 3 rolled over the saturated types.  P = x - h f1 stays in
// its sparse stage-1 type; only the stage point, the accumulator and the current stage derivative are saturated.
template <class T, class Model, class X, class U>
RDB_HD auto rk3_rolled(const Model& model, const X& x, const U& u, T h, T t) {
    using XS = saturated_t<Model, T, X, U>;
    using FS = decltype(feval<T>(model, rstd::declval<const XS&>(), u, t));
    const auto f1 = feval<T>(model, x, u, t);
    const auto P = axpy(x, -h, f1);
    XS Xs = widen_vec<XS>(axpy(x, T(0.5) * h, f1));
    FS acc = widen_vec<FS>(f1);
#pragma unroll 1
    for (int s = 0; s < 2; ++s) {
        const FS f = feval<T>(model, Xs, u, t + (s == 0 ? T(0.5) * h : h));
        acc = axpy(acc, s == 0 ? T(4) : T(1), f);
        if (s == 0) Xs = widen_vec<XS>(axpy(P, T(2) * h, f));
    }
    return axpy(x, h * T(1.0 / 6.0), acc);
}

template <int Q, class T, int ROLL = 0, class Model, class X, class U>
RDB_HD auto integrate(const Model& model, const X& x, const U& u, T h, T t = T(0)) {
    if constexpr (Q == Q_CONTINUOUS) {
        return feval<T>(model, x, u, t);
    } else if constexpr (folds_scale<Model>::value && RDB_INCREMENT_FORM && (ROLL == 0 || RDB_INCREMENT_FORM > 1)) {
        // (rolled stage loops keep the classical form: there the stage point x + G is a loop-carried copy of G, one MOV per partial
        // where the classical form spends its one FMUL — measured no gain, profiles/tuning_r01.md)
        if constexpr (Q == Q_EULER) return vadd(x, fseval<T>(model, x, u, h, t));
        else if constexpr (Q == Q_RK2) return vadd(x, fseval<T>(model, vadd(x, fseval<T>(model, x, u, T(0.5) * h, t)), u, h, t + T(0.5) * h));
        else if constexpr (Q == Q_RK3) return rk3_increments<ROLL, T>(model, x, u, h, t);
        else { static_assert(Q == Q_RK4, "unknown quadrature rule"); return rk4_increments<ROLL, T>(model, x, u, h, t); }
    } else if constexpr (Q == Q_RK4 && ROLL != 0) {
        return rk4_rolled<ROLL, T>(model, x, u, h, t);
    } else if constexpr (Q == Q_RK3 && ROLL != 0) {
        return rk3_rolled<T>(model, x, u, h, t);
    } else if constexpr (Q == Q_EULER) {
        return axpy(x, h, feval<T>(model, x, u, t));
    } else if constexpr (Q == Q_RK2) {
        auto f1 = feval<T>(model, x, u, t);
        auto f2 = feval<T>(model, axpy(x, T(0.5) * h, f1), u, t + T(0.5) * h);
        return axpy(x, h, f2);
    } else if constexpr (Q == Q_RK3) {
        auto f1 = feval<T>(model, x, u, t);
        auto f2 = feval<T>(model, axpy(x, T(0.5) * h, f1), u, t + T(0.5) * h);
        auto f3 = feval<T>(model, axpy(axpy(x, -h, f1), T(2) * h, f2), u, t + h);
        auto acc = axpy(vadd(f1, f3), T(4), f2);
        return axpy(x, h * T(1.0 / 6.0), acc);
    } else {
        static_assert(Q == Q_RK4, "unknown quadrature rule");
        auto f1 = feval<T>(model, x, u, t);
        auto f2 = feval<T>(model, axpy(x, T(0.5) * h, f1), u, t + T(0.5) * h);
        auto acc1 = axpy(f1, T(2), f2);
        auto f3 = feval<T>(model, axpy(x, T(0.5) * h, f2), u, t + T(0.5) * h);
        auto acc2 = axpy(acc1, T(2), f3);
        auto f4 = feval<T>(model, axpy(x, h, f3), u, t + h);
        return axpy(x, h * T(1.0 / 6.0), vadd(acc2, f4));
    }
}

}  // namespace rdb
