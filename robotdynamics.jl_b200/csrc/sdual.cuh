// sdual.cuh — compile-time-sparse forward-mode numbers and heterogeneous fixed vectors (device side).
//
// The reference's default Jacobian path is forward-mode AD over the whole [x;u] vector
// (reference: src/jacobian_gen.jl:485-507, ForwardDiff.jacobian).  On the GPU we keep that meaning
// but make the set of non-zero partials part of the TYPE: SD<T,MASK> carries a value and one partial
// per set bit of MASK (bit j <-> column j of [x;u]).  Every operation returns the union mask, so
// structural zeros of the stage Jacobians are never computed, never stored,
// and never occupy a register — without anybody hand-deriving a sparsity pattern per model.
// Because element types differ, vectors are heterogeneous tuples (Vec<...>) indexed at compile time.
#pragma once
// Compiles under nvcc (library build) and under NVRTC (user-defined models, custom.cu): `rstd` is std or cuda::std.
#ifdef __CUDACC_RTC__
#include <cuda/std/utility>
#include <cuda/std/type_traits>
namespace rstd = cuda::std;
typedef unsigned int uint32_t;
typedef unsigned long long uintptr_t;
#else
#include <cstdint>
#include <utility>
#include <type_traits>
namespace rstd = std;
#endif

#ifndef RDB_HD
#define RDB_HD __host__ __device__ __forceinline__
#endif

namespace rdb {

using mask_t = uint32_t;

__host__ __device__ constexpr int cpopc(mask_t m) { int c = 0; while (m) { c += int(m & 1u); m >>= 1; } return c; }
__host__ __device__ constexpr bool chas(mask_t m, int j) { return (m >> j) & 1u; }
__host__ __device__ constexpr int cslot(mask_t m, int j) { return cpopc(m & ((mask_t(1) << j) - 1u)); }

template <class T> struct ident { using type = T; };
template <class T> using ident_t = typename ident<T>::type;

// ---------------------------------------------------------------------------------------------
// Structural zero: 0 * x == Zero, 0 + x == x.  Lets generic code (e.g. rotate([0,0,F])) prune itself.
// ---------------------------------------------------------------------------------------------
struct Zero {};

// ---------------------------------------------------------------------------------------------
// Partial storage.  double: one partial per slot.  float: TWO partials per slot (columns 2p, 2p+1 share a float2) so that
// all partial arithmetic issues as packed fp32x2 instructions (Blackwell `fma.rn.f32x2` / `add.f32x2` / `mul.f32x2`, SASS
// FFMA2/FADD2/FMUL2): same FMA-pipe flops as scalar FFMA (measured 125 vs 126 FMA/clk/SM, scripts/micro/fma_rate.cu) for HALF
// the issue slots — and the rigid-body Jacobian kernels are issue-bound.  A slot whose other column is structurally zero
// carries an explicit 0 in that lane (lane-wise linear arithmetic keeps it 0; it is never read).
// ---------------------------------------------------------------------------------------------
#ifndef RDB_PACK_F32
#define RDB_PACK_F32 0   // measured slower on the quadrotor RK4 kernel (59.3 vs 51.1 us): padded lanes + pair alignment, see profiles/tuning_r01.md
#endif
// a * b + c as ONE rounding, spelled out: the compiler may contract `p*q + r*s` into an FMA around either product, and it need not
// choose the same one in two instantiations of a kernel — results are promised to be bit-identical across layouts and pointer kinds
// (tests: test_layouts_and_pointer_kinds_agree_bitwise), so every multiply-add of the partial arithmetic and of the hand-written
// elemental operations is an explicit fma
RDB_HD float fma_(float a, float b, float c) { return fmaf(a, b, c); }
RDB_HD double fma_(double a, double b, double c) { return fma(a, b, c); }
template <class T> struct PK {
    using vec = T;
    static constexpr int W = 1;
    RDB_HD static vec splat(T s) { return s; }
    RDB_HD static vec zero() { return T(0); }
    RDB_HD static vec add(vec a, vec b) { return a + b; }
    RDB_HD static vec sub(vec a, vec b) { return a - b; }
    RDB_HD static vec mul(vec a, vec b) { return a * b; }
    RDB_HD static vec fma(vec a, vec b, vec c) { return fma_(a, b, c); }
    RDB_HD static vec neg(vec a) { return -a; }
    RDB_HD static vec sel(bool on, vec a) { return on ? a : T(0); }
    template <int L> RDB_HD static T lane(vec a) { return a; }
    template <int L> RDB_HD static void set_lane(vec& a, T v) { a = v; }
};
#if RDB_PACK_F32 && defined(__CUDA_ARCH__)
template <> struct PK<float> {
    using vec = float2;
    static constexpr int W = 2;
    RDB_HD static vec splat(float s) { return make_float2(s, s); }
    RDB_HD static vec zero() { return make_float2(0.0f, 0.0f); }
    RDB_HD static vec add(vec a, vec b) { return __fadd2_rn(a, b); }
    RDB_HD static vec sub(vec a, vec b) { return __ffma2_rn(b, make_float2(-1.0f, -1.0f), a); }
    RDB_HD static vec mul(vec a, vec b) { return __fmul2_rn(a, b); }
    RDB_HD static vec fma(vec a, vec b, vec c) { return __ffma2_rn(a, b, c); }
    RDB_HD static vec neg(vec a) { return make_float2(-a.x, -a.y); }
    RDB_HD static vec sel(bool on, vec a) { return on ? a : make_float2(0.0f, 0.0f); }
    template <int L> RDB_HD static float lane(vec a) { return L == 0 ? a.x : a.y; }
    template <int L> RDB_HD static void set_lane(vec& a, float v) { if (L == 0) a.x = v; else a.y = v; }
};
#endif
// slot mask of a column mask: bit p <-> slot p
template <class T> __host__ __device__ constexpr mask_t smask(mask_t m) {
    if (PK<T>::W == 1) return m;
    mask_t r = 0;
    for (int p = 0; p < 16; ++p) if ((m >> (2 * p)) & 3u) r |= mask_t(1) << p;
    return r;
}

// ---------------------------------------------------------------------------------------------
// Sparse dual number.
// ---------------------------------------------------------------------------------------------
template <class T, mask_t M>
struct SD {
    static_assert(M != 0, "SD needs at least one partial; use plain T otherwise");
    using P = PK<T>;
    static constexpr mask_t mask = M;
    static constexpr mask_t slots = smask<T>(M);
    static constexpr int NS = cpopc(smask<T>(M));
    T v;
    typename P::vec d[NS];
    template <int J> RDB_HD T part() const {
        if constexpr (chas(M, J)) return P::template lane<J % P::W>(d[cslot(slots, J / P::W)]); else return T(0);
    }
    // overwrite partial J (J must belong to M)
    template <int J> RDB_HD void set_part(T val) {
        static_assert(chas(M, J), "set_part: column not in mask");
        P::template set_lane<J % P::W>(d[cslot(slots, J / P::W)], val);
    }
    RDB_HD void zero_parts() { for (int i = 0; i < NS; ++i) d[i] = P::zero(); }
};

template <class X> struct is_sd : rstd::false_type {};
template <class T, mask_t M> struct is_sd<SD<T, M>> : rstd::true_type {};
template <class X> struct mask_of { static constexpr mask_t value = 0; };
template <class T, mask_t M> struct mask_of<SD<T, M>> { static constexpr mask_t value = M; };

template <class T> RDB_HD T val(const T& a) { return a; }
template <class T, mask_t M> RDB_HD T val(const SD<T, M>& a) { return a.v; }

// ---- slot-wise kernels of the binary operations (SA, SB: slot masks of the operands; result slots = SA | SB) -----------
template <class T, mask_t A, mask_t B, int J = 0>
RDB_HD void add_parts(SD<T, (A | B)>& r, const SD<T, A>& a, const SD<T, B>& b) {
    constexpr mask_t SA = smask<T>(A), SB = smask<T>(B), SR = SA | SB;
    if constexpr ((SR >> J) != 0) {
        if constexpr (chas(SA, J) && chas(SB, J)) r.d[cslot(SR, J)] = PK<T>::add(a.d[cslot(SA, J)], b.d[cslot(SB, J)]);
        else if constexpr (chas(SA, J)) r.d[cslot(SR, J)] = a.d[cslot(SA, J)];
        else if constexpr (chas(SB, J)) r.d[cslot(SR, J)] = b.d[cslot(SB, J)];
        add_parts<T, A, B, J + 1>(r, a, b);
    }
}
template <class T, mask_t A, mask_t B, int J = 0>
RDB_HD void sub_parts(SD<T, (A | B)>& r, const SD<T, A>& a, const SD<T, B>& b) {
    constexpr mask_t SA = smask<T>(A), SB = smask<T>(B), SR = SA | SB;
    if constexpr ((SR >> J) != 0) {
        if constexpr (chas(SA, J) && chas(SB, J)) r.d[cslot(SR, J)] = PK<T>::sub(a.d[cslot(SA, J)], b.d[cslot(SB, J)]);
        else if constexpr (chas(SA, J)) r.d[cslot(SR, J)] = a.d[cslot(SA, J)];
        else if constexpr (chas(SB, J)) r.d[cslot(SR, J)] = PK<T>::neg(b.d[cslot(SB, J)]);
        sub_parts<T, A, B, J + 1>(r, a, b);
    }
}
// r.d = a.d * bv + b.d * av     (av, bv pre-splatted)
template <class T, mask_t A, mask_t B, int J = 0>
RDB_HD void mul_parts(SD<T, (A | B)>& r, const SD<T, A>& a, const SD<T, B>& b, typename PK<T>::vec av, typename PK<T>::vec bv) {
    constexpr mask_t SA = smask<T>(A), SB = smask<T>(B), SR = SA | SB;
    if constexpr ((SR >> J) != 0) {
        if constexpr (chas(SA, J) && chas(SB, J)) r.d[cslot(SR, J)] = PK<T>::fma(b.d[cslot(SB, J)], av, PK<T>::mul(a.d[cslot(SA, J)], bv));
        else if constexpr (chas(SA, J)) r.d[cslot(SR, J)] = PK<T>::mul(a.d[cslot(SA, J)], bv);
        else if constexpr (chas(SB, J)) r.d[cslot(SR, J)] = PK<T>::mul(b.d[cslot(SB, J)], av);
        mul_parts<T, A, B, J + 1>(r, a, b, av, bv);
    }
}
// r.d = (a.d - q * b.d) * ib  =  a.d * ib + b.d * (-q * ib)        (ib = 1/b, q = a/b; ibv, nqib pre-splatted)
template <class T, mask_t A, mask_t B, int J = 0>
RDB_HD void div_parts(SD<T, (A | B)>& r, const SD<T, A>& a, const SD<T, B>& b, typename PK<T>::vec ibv, typename PK<T>::vec nqib) {
    constexpr mask_t SA = smask<T>(A), SB = smask<T>(B), SR = SA | SB;
    if constexpr ((SR >> J) != 0) {
        if constexpr (chas(SA, J) && chas(SB, J)) r.d[cslot(SR, J)] = PK<T>::fma(b.d[cslot(SB, J)], nqib, PK<T>::mul(a.d[cslot(SA, J)], ibv));
        else if constexpr (chas(SA, J)) r.d[cslot(SR, J)] = PK<T>::mul(a.d[cslot(SA, J)], ibv);
        else if constexpr (chas(SB, J)) r.d[cslot(SR, J)] = PK<T>::mul(b.d[cslot(SB, J)], nqib);
        div_parts<T, A, B, J + 1>(r, a, b, ibv, nqib);
    }
}

template <class T, mask_t A, mask_t B>
RDB_HD SD<T, (A | B)> operator+(const SD<T, A>& a, const SD<T, B>& b) { SD<T, (A | B)> r; r.v = a.v + b.v; add_parts<T, A, B>(r, a, b); return r; }
template <class T, mask_t A, mask_t B>
RDB_HD SD<T, (A | B)> operator-(const SD<T, A>& a, const SD<T, B>& b) { SD<T, (A | B)> r; r.v = a.v - b.v; sub_parts<T, A, B>(r, a, b); return r; }
template <class T, mask_t A, mask_t B>
RDB_HD SD<T, (A | B)> operator*(const SD<T, A>& a, const SD<T, B>& b) {
    SD<T, (A | B)> r; r.v = a.v * b.v; mul_parts<T, A, B>(r, a, b, PK<T>::splat(a.v), PK<T>::splat(b.v)); return r;
}
template <class T, mask_t A, mask_t B>
RDB_HD SD<T, (A | B)> operator/(const SD<T, A>& a, const SD<T, B>& b) {
    SD<T, (A | B)> r; const T ib = T(1) / b.v; r.v = a.v * ib;
    div_parts<T, A, B>(r, a, b, PK<T>::splat(ib), PK<T>::splat(-(r.v * ib))); return r;
}

// ---- SD with scalar -----------------------------------------------------------------------------
template <class T, mask_t A> RDB_HD SD<T, A> scale_parts(const SD<T, A>& a, T value, T s) {
    SD<T, A> r; r.v = value; const auto sv = PK<T>::splat(s); for (int i = 0; i < SD<T, A>::NS; ++i) r.d[i] = PK<T>::mul(a.d[i], sv); return r;
}
template <class T, mask_t A> RDB_HD SD<T, A> operator-(const SD<T, A>& a) { SD<T, A> r; r.v = -a.v; for (int i = 0; i < SD<T, A>::NS; ++i) r.d[i] = PK<T>::neg(a.d[i]); return r; }
template <class T, mask_t A> RDB_HD SD<T, A> operator+(const SD<T, A>& a, ident_t<T> b) { SD<T, A> r = a; r.v = a.v + b; return r; }
template <class T, mask_t A> RDB_HD SD<T, A> operator+(ident_t<T> b, const SD<T, A>& a) { SD<T, A> r = a; r.v = b + a.v; return r; }
template <class T, mask_t A> RDB_HD SD<T, A> operator-(const SD<T, A>& a, ident_t<T> b) { SD<T, A> r = a; r.v = a.v - b; return r; }
template <class T, mask_t A> RDB_HD SD<T, A> operator-(ident_t<T> b, const SD<T, A>& a) { return scale_parts<T, A>(a, b - a.v, T(-1)); }
template <class T, mask_t A> RDB_HD SD<T, A> operator*(const SD<T, A>& a, ident_t<T> b) { return scale_parts<T, A>(a, a.v * b, b); }
template <class T, mask_t A> RDB_HD SD<T, A> operator*(ident_t<T> b, const SD<T, A>& a) { return scale_parts<T, A>(a, a.v * b, b); }
template <class T, mask_t A> RDB_HD SD<T, A> operator/(const SD<T, A>& a, ident_t<T> b) { const T ib = T(1) / b; return scale_parts<T, A>(a, a.v * ib, ib); }
template <class T, mask_t A> RDB_HD SD<T, A> operator/(ident_t<T> b, const SD<T, A>& a) {
    const T ib = T(1) / a.v; const T q = b * ib; return scale_parts<T, A>(a, q, -q * ib);
}

// ---- fused  c + SGN * a * b  and  c + SGN * a^2 ---------------------------------------------------------------------------
// The rigid-body kernels are bound by FP32 instruction issue; a product followed by an add costs FMUL + FFMA + FADD per partial
// through the operators above (IEEE forbids re-associating it into an FMA chain).  Written as ONE accumulation the same value
// costs two FFMAs per partial (one for a square): dot and cross products, matrix-vector products and the model formulas use it.
template <class T, class X> struct opnd {                        // plain scalar operand
    static constexpr mask_t mask = 0;
    RDB_HD static T v(const X& x) { return x; }
};
template <class T> struct opnd<T, Zero> {                        // structural zero operand
    static constexpr mask_t mask = 0;
    RDB_HD static T v(const Zero&) { return T(0); }
};
template <class T, mask_t M> struct opnd<T, SD<T, M>> {
    static constexpr mask_t mask = M;
    RDB_HD static T v(const SD<T, M>& x) { return x.v; }
    template <int J> RDB_HD static typename PK<T>::vec slot(const SD<T, M>& x) { return x.d[cslot(smask<T>(M), J)]; }
};
template <class T, class A, class B, class C, mask_t R, int J = 0>
RDB_HD void fmadd_parts(SD<T, R>& r, const A& a, const B& b, const C& c, typename PK<T>::vec sav, typename PK<T>::vec sbv) {
    constexpr mask_t SA = smask<T>(opnd<T, A>::mask), SB = smask<T>(opnd<T, B>::mask), SC = smask<T>(opnd<T, C>::mask), SR = smask<T>(R);
    if constexpr ((SR >> J) != 0) {
        if constexpr (chas(SR, J)) {
            constexpr bool ha = chas(SA, J), hb = chas(SB, J), hc = chas(SC, J);
            typename PK<T>::vec acc = PK<T>::zero();
            if constexpr (hc) acc = opnd<T, C>::template slot<J>(c);
            if constexpr (ha) { if constexpr (hc) acc = PK<T>::fma(opnd<T, A>::template slot<J>(a), sbv, acc); else acc = PK<T>::mul(opnd<T, A>::template slot<J>(a), sbv); }
            if constexpr (hb) { if constexpr (hc || ha) acc = PK<T>::fma(opnd<T, B>::template slot<J>(b), sav, acc); else acc = PK<T>::mul(opnd<T, B>::template slot<J>(b), sav); }
            r.d[cslot(SR, J)] = acc;
        }
        fmadd_parts<T, A, B, C, R, J + 1>(r, a, b, c, sav, sbv);
    }
}
// c + SGN * a * b   (SGN = +1 or -1); operands: plain T, SD<T,.> or Zero
template <class T, int SGN = 1, class A, class B, class C>
RDB_HD auto fmadd(const A& a, const B& b, const C& c) {
    if constexpr (rstd::is_same<A, Zero>::value || rstd::is_same<B, Zero>::value) return c;
    else if constexpr (rstd::is_same<C, Zero>::value) { if constexpr (SGN > 0) return a * b; else return -(a * b); }
    else {
        constexpr mask_t R = opnd<T, A>::mask | opnd<T, B>::mask | opnd<T, C>::mask;
        const T av = opnd<T, A>::v(a), bv = opnd<T, B>::v(b), cv = opnd<T, C>::v(c);
        const T sav = SGN > 0 ? av : -av, sbv = SGN > 0 ? bv : -bv;
        if constexpr (R == 0) return sav * bv + cv;
        else { SD<T, R> r; r.v = sav * bv + cv; fmadd_parts<T, A, B, C, R>(r, a, b, c, PK<T>::splat(sav), PK<T>::splat(sbv)); return r; }
    }
}
template <class T, class A, class C, mask_t R, int J = 0>
RDB_HD void sqadd_parts(SD<T, R>& r, const A& a, const C& c, typename PK<T>::vec s2a) {
    constexpr mask_t SA = smask<T>(opnd<T, A>::mask), SC = smask<T>(opnd<T, C>::mask), SR = smask<T>(R);
    if constexpr ((SR >> J) != 0) {
        if constexpr (chas(SR, J)) {
            if constexpr (chas(SA, J) && chas(SC, J)) r.d[cslot(SR, J)] = PK<T>::fma(opnd<T, A>::template slot<J>(a), s2a, opnd<T, C>::template slot<J>(c));
            else if constexpr (chas(SA, J)) r.d[cslot(SR, J)] = PK<T>::mul(opnd<T, A>::template slot<J>(a), s2a);
            else r.d[cslot(SR, J)] = opnd<T, C>::template slot<J>(c);
        }
        sqadd_parts<T, A, C, R, J + 1>(r, a, c, s2a);
    }
}
// c + SGN * a^2
template <class T, int SGN = 1, class A, class C>
RDB_HD auto sqadd(const A& a, const C& c) {
    if constexpr (rstd::is_same<A, Zero>::value) return c;
    else {
        constexpr mask_t R = opnd<T, A>::mask | opnd<T, C>::mask;
        const T av = opnd<T, A>::v(a), cv = opnd<T, C>::v(c);
        const T sa = SGN > 0 ? av : -av;
        if constexpr (R == 0) return sa * av + cv;
        else { SD<T, R> r; r.v = sa * av + cv; sqadd_parts<T, A, C, R>(r, a, c, PK<T>::splat(sa + sa)); return r; }
    }
}

// ---- Zero algebra -------------------------------------------------------------------------------
RDB_HD Zero operator+(Zero, Zero) { return {}; }
RDB_HD Zero operator-(Zero, Zero) { return {}; }
RDB_HD Zero operator*(Zero, Zero) { return {}; }
RDB_HD Zero operator-(Zero) { return {}; }
template <class X, class = rstd::enable_if_t<!rstd::is_same<X, Zero>::value>> RDB_HD X operator+(Zero, const X& x) { return x; }
template <class X, class = rstd::enable_if_t<!rstd::is_same<X, Zero>::value>> RDB_HD X operator+(const X& x, Zero) { return x; }
template <class X, class = rstd::enable_if_t<!rstd::is_same<X, Zero>::value>> RDB_HD X operator-(const X& x, Zero) { return x; }
template <class X, class = rstd::enable_if_t<!rstd::is_same<X, Zero>::value>> RDB_HD auto operator-(Zero, const X& x) { return -x; }
template <class X, class = rstd::enable_if_t<!rstd::is_same<X, Zero>::value>> RDB_HD Zero operator*(Zero, const X&) { return {}; }
template <class X, class = rstd::enable_if_t<!rstd::is_same<X, Zero>::value>> RDB_HD Zero operator*(const X&, Zero) { return {}; }

// ---- elementary functions -----------------------------------------------------------------------
RDB_HD void sincos_(float a, float& s, float& c) { sincosf(a, &s, &c); }
RDB_HD void sincos_(double a, double& s, double& c) { sincos(a, &s, &c); }
template <class T, mask_t A>
RDB_HD void sincos_(const SD<T, A>& a, SD<T, A>& s, SD<T, A>& c) {
    T sv, cv;
    sincos_(a.v, sv, cv);
    const auto cs = PK<T>::splat(cv), ns = PK<T>::splat(-sv);
    for (int i = 0; i < SD<T, A>::NS; ++i) { s.d[i] = PK<T>::mul(a.d[i], cs); c.d[i] = PK<T>::mul(a.d[i], ns); }
    s.v = sv; c.v = cv;
}
// value-level sincos of an angle NEAR one whose sine and cosine are known:  sin(a0 + d) = s0 cos d + c0 sin d  with
// degree-11/12 Taylor polynomials in d (|d| <= 1/4: truncation < 3e-18) — about half the FP64 work of a full-range sincos.
// Used for RK stages 2..4, whose angles differ from the stage-1 angle by O(h).  Falls back to the full sincos otherwise.
template <class T> __host__ __device__ __noinline__ void sincos_far(T a, T& s, T& c) { sincos_(a, s, c); }   // rare path, kept out of line
template <class T>
RDB_HD void sincos_near(T a, T a0, T s0, T c0, T& s, T& c) {
    const T d = a - a0;
    if (d > T(0.25) || d < T(-0.25)) { sincos_far(a, s, c); return; }
    const T d2 = d * d;
#ifndef RDB_TUNE_NO_NEAR8
    if (d2 < T(1.0 / 1024)) {                                 // |d| < 1/32 (h w of a typical step): degree 7 / 8 truncate below 1e-19
        T ps = T(-1.9841269841269841e-04);
        ps = ps * d2 + T(8.3333333333333332e-03);
        ps = ps * d2 + T(-1.6666666666666666e-01);
        const T sd = d + d * (d2 * ps);
        T pc = T(2.4801587301587302e-05);
        pc = pc * d2 + T(-1.3888888888888889e-03);
        pc = pc * d2 + T(4.1666666666666664e-02);
        pc = pc * d2 + T(-0.5);
        const T cdm1 = d2 * pc;
        s = s0 + (s0 * cdm1 + c0 * sd);
        c = c0 + (c0 * cdm1 - s0 * sd);
        return;
    }
#endif
    T ps = T(-2.5052108385441720e-08);                       // -1/11!
    ps = ps * d2 + T(2.7557319223985893e-06);                //  1/9!
    ps = ps * d2 + T(-1.9841269841269841e-04);               // -1/7!
    ps = ps * d2 + T(8.3333333333333332e-03);                //  1/5!
    ps = ps * d2 + T(-1.6666666666666666e-01);               // -1/3!
    const T sd = d + d * (d2 * ps);
    T pc = T(2.0876756987868100e-09);                        //  1/12!
    pc = pc * d2 + T(-2.7557319223985888e-07);               // -1/10!
    pc = pc * d2 + T(2.4801587301587302e-05);                //  1/8!
    pc = pc * d2 + T(-1.3888888888888889e-03);               // -1/6!
    pc = pc * d2 + T(4.1666666666666664e-02);                //  1/4!
    pc = pc * d2 + T(-0.5);
    const T cdm1 = d2 * pc;                                   // cos d - 1
    s = s0 + (s0 * cdm1 + c0 * sd);
    c = c0 + (c0 * cdm1 - s0 * sd);
}
// duals of sin/cos from their values
template <class T> RDB_HD void sincos_with(const T&, T sv, T cv, T& s, T& c) { s = sv; c = cv; }
template <class T, mask_t A>
RDB_HD void sincos_with(const SD<T, A>& a, T sv, T cv, SD<T, A>& s, SD<T, A>& c) {
    const auto cs = PK<T>::splat(cv), ns = PK<T>::splat(-sv);
    for (int i = 0; i < SD<T, A>::NS; ++i) { s.d[i] = PK<T>::mul(a.d[i], cs); c.d[i] = PK<T>::mul(a.d[i], ns); }
    s.v = sv; c.v = cv;
}

// x^2 with ONE multiply per partial (the generic product rule spends two on a*a)
RDB_HD float sq_(float a) { return a * a; }
RDB_HD double sq_(double a) { return a * a; }
template <class T, mask_t A> RDB_HD SD<T, A> sq_(const SD<T, A>& a) { return scale_parts<T, A>(a, a.v * a.v, a.v + a.v); }
RDB_HD Zero sq_(Zero) { return {}; }

// convenience forms for user-written models
template <class S> RDB_HD S sin_(const S& a) { S s = a, c = a; sincos_(a, s, c); return s; }
template <class S> RDB_HD S cos_(const S& a) { S s = a, c = a; sincos_(a, s, c); return c; }
RDB_HD float exp_(float a) { return expf(a); }
RDB_HD double exp_(double a) { return exp(a); }
template <class T, mask_t A> RDB_HD SD<T, A> exp_(const SD<T, A>& a) { const T e = exp_(a.v); return scale_parts<T, A>(a, e, e); }
RDB_HD float sqrt_(float a) { return sqrtf(a); }
RDB_HD double sqrt_(double a) { return sqrt(a); }
template <class T, mask_t A> RDB_HD SD<T, A> sqrt_(const SD<T, A>& a) { const T r = sqrt_(a.v); return scale_parts<T, A>(a, r, T(0.5) / r); }
RDB_HD float rsqrt_(float a) { return 1.0f / sqrtf(a); }
RDB_HD double rsqrt_(double a) { return 1.0 / sqrt(a); }
template <class T, mask_t A>
RDB_HD SD<T, A> rsqrt_(const SD<T, A>& a) {   // a^(-1/2);  d = -1/2 a^(-3/2) da
    const T r = rsqrt_(a.v); return scale_parts<T, A>(a, r, T(-0.5) * r / a.v);
}
// max(0, a) as the reference's Quadrotor writes it (test/quadrotor.jl:67-70).  Under ForwardDiff `max(0, d)` promotes the 0 to a
// Dual and returns `ifelse(d < 0, 0, d)` (Base.max(x::T, y::T) = ifelse(isless(y, x), x, y); DiffRules' rule for max(x, y) gives the
// same: d/dy = (x > y ? 0 : 1)): the clamp is active only for a < 0, and AT the tie a == 0 the result is `a` itself, partials kept.
// A zero control (the common initial guess) therefore has d relu/da = 1, not 0.
RDB_HD float relu_(float a) { return a < 0.0f ? 0.0f : a; }
RDB_HD double relu_(double a) { return a < 0.0 ? 0.0 : a; }
template <class T, mask_t A>
RDB_HD SD<T, A> relu_(const SD<T, A>& a) {
    SD<T, A> r; const bool on = !(a.v < T(0)); r.v = on ? a.v : T(0); for (int i = 0; i < SD<T, A>::NS; ++i) r.d[i] = PK<T>::sel(on, a.d[i]); return r;
}

// ---------------------------------------------------------------------------------------------
// Heterogeneous fixed vector.
// ---------------------------------------------------------------------------------------------
template <int I, class E> struct Leaf { E e; };
template <class Seq, class... Es> struct VecBase;
template <size_t... Is, class... Es>
struct VecBase<rstd::index_sequence<Is...>, Es...> : Leaf<int(Is), Es>... {
    RDB_HD VecBase() {}
    RDB_HD VecBase(const Es&... es) : Leaf<int(Is), Es>{es}... {}
};
template <class... Es>
struct Vec : VecBase<rstd::index_sequence_for<Es...>, Es...> {
    using Base = VecBase<rstd::index_sequence_for<Es...>, Es...>;
    static constexpr int size = int(sizeof...(Es));
    RDB_HD Vec() {}
    RDB_HD Vec(const Es&... es) : Base(es...) {}
};
template <int I, class E> RDB_HD const E& get(const Leaf<I, E>& l) { return l.e; }
template <int I, class E> RDB_HD E& get(Leaf<I, E>& l) { return l.e; }
template <class... Es> RDB_HD Vec<Es...> vec(const Es&... es) { return Vec<Es...>(es...); }
template <class V> using iseq = rstd::make_index_sequence<size_t(V::size)>;

template <int S, class V, size_t... Is> RDB_HD auto slice_impl(const V& v, rstd::index_sequence<Is...>) { return vec(get<S + int(Is)>(v)...); }
template <int S, int L, class V> RDB_HD auto slice(const V& v) { return slice_impl<S>(v, rstd::make_index_sequence<size_t(L)>{}); }

template <class... As, class... Bs, size_t... Is, size_t... Js>
RDB_HD auto cat_impl(const Vec<As...>& a, const Vec<Bs...>& b, rstd::index_sequence<Is...>, rstd::index_sequence<Js...>) { return vec(get<int(Is)>(a)..., get<int(Js)>(b)...); }
template <class... As, class... Bs> RDB_HD auto cat(const Vec<As...>& a, const Vec<Bs...>& b) { return cat_impl(a, b, rstd::index_sequence_for<As...>{}, rstd::index_sequence_for<Bs...>{}); }
template <class A, class B, class C, class... R> RDB_HD auto cat(const A& a, const B& b, const C& c, const R&... r) { return cat(cat(a, b), c, r...); }

// elementwise a + s*b, a + b, s*a, a - b  (s scalar of any numeric kind)
template <class A, class S, class B, size_t... Is> RDB_HD auto axpy_impl(const A& a, const S& s, const B& b, rstd::index_sequence<Is...>) { return vec(fmadd<S>(s, get<int(Is)>(b), get<int(Is)>(a))...); }
template <class A, class S, class B> RDB_HD auto axpy(const A& a, const S& s, const B& b) { return axpy_impl(a, s, b, iseq<A>{}); }
template <class A, class B, size_t... Is> RDB_HD auto vadd_impl(const A& a, const B& b, rstd::index_sequence<Is...>) { return vec((get<int(Is)>(a) + get<int(Is)>(b))...); }
template <class A, class B> RDB_HD auto vadd(const A& a, const B& b) { return vadd_impl(a, b, iseq<A>{}); }
template <class A, class B, size_t... Is> RDB_HD auto vsub_impl(const A& a, const B& b, rstd::index_sequence<Is...>) { return vec((get<int(Is)>(a) - get<int(Is)>(b))...); }
template <class A, class B> RDB_HD auto vsub(const A& a, const B& b) { return vsub_impl(a, b, iseq<A>{}); }
template <class S, class A, size_t... Is> RDB_HD auto vscale_impl(const S& s, const A& a, rstd::index_sequence<Is...>) { return vec((s * get<int(Is)>(a))...); }
template <class S, class A> RDB_HD auto vscale(const S& s, const A& a) { return vscale_impl(s, a, iseq<A>{}); }

// 3-vector algebra on heterogeneous triples
template <class T, class A, class B> RDB_HD auto dot3(const A& a, const B& b) {
    return fmadd<T>(get<2>(a), get<2>(b), fmadd<T>(get<1>(a), get<1>(b), get<0>(a) * get<0>(b)));
}
template <class T, class V> RDB_HD auto norm2_3(const V& a) { return sqadd<T>(get<2>(a), sqadd<T>(get<1>(a), sq_(get<0>(a)))); }
template <class T, class A, class B> RDB_HD auto cross3(const A& a, const B& b) {
    return vec(fmadd<T, -1>(get<2>(a), get<1>(b), get<1>(a) * get<2>(b)),
               fmadd<T, -1>(get<0>(a), get<2>(b), get<2>(a) * get<0>(b)),
               fmadd<T, -1>(get<1>(a), get<0>(b), get<0>(a) * get<1>(b)));
}
// y = M x for a row-major 3x3 of plain scalars
template <class T, class A> RDB_HD auto mat3_mul(const T* M, const A& x) {
    return vec(fmadd<T>(M[2], get<2>(x), fmadd<T>(M[1], get<1>(x), M[0] * get<0>(x))),
               fmadd<T>(M[5], get<2>(x), fmadd<T>(M[4], get<1>(x), M[3] * get<0>(x))),
               fmadd<T>(M[8], get<2>(x), fmadd<T>(M[7], get<1>(x), M[6] * get<0>(x))));
}

// ---------------------------------------------------------------------------------------------
// Elemental operations with hand-written local derivatives.
//   chain<T>(v, c, in):  the number with value v and partials  sum_i c[i] * d(in_i)   (in: a Vec of plain / SD / Zero elements).
// A model can evaluate a block of its formulas on plain values, compute the block's local Jacobian by hand (like the reference's
// UserDefined analytic Jacobians, test/cartpole_model.jl:57-96) and hand both to chain(): one FMA per (output, input, partial)
// instead of forward mode through every intermediate.  With plain inputs chain() returns v and the local derivatives are dead code.
// ---------------------------------------------------------------------------------------------
template <class V, int I> using elem_t = rstd::remove_cv_t<rstd::remove_reference_t<decltype(get<I>(rstd::declval<const V&>()))>>;
template <class T, class V, int I = 0> __host__ __device__ constexpr mask_t vec_mask() {
    if constexpr (I == V::size) return 0; else return opnd<T, elem_t<V, I>>::mask | vec_mask<T, V, I + 1>();
}
template <class T, int J, int I, bool HAVE, class V>
RDB_HD typename PK<T>::vec chain_acc(const T* c, const V& in, typename PK<T>::vec acc) {
    if constexpr (I == V::size) return acc;
    else {
        using E = elem_t<V, I>;
        if constexpr (chas(smask<T>(opnd<T, E>::mask), J)) {
            const auto s = opnd<T, E>::template slot<J>(get<I>(in));
            if constexpr (HAVE) return chain_acc<T, J, I + 1, true>(c, in, PK<T>::fma(s, PK<T>::splat(c[I]), acc));
            else return chain_acc<T, J, I + 1, true>(c, in, PK<T>::mul(s, PK<T>::splat(c[I])));
        } else return chain_acc<T, J, I + 1, HAVE>(c, in, acc);
    }
}
template <class T, mask_t R, class V, int J = 0>
RDB_HD void chain_parts(SD<T, R>& r, const T* c, const V& in) {
    constexpr mask_t SR = smask<T>(R);
    if constexpr ((SR >> J) != 0) {
        if constexpr (chas(SR, J)) r.d[cslot(SR, J)] = chain_acc<T, J, 0, false>(c, in, PK<T>::zero());
        chain_parts<T, R, V, J + 1>(r, c, in);
    }
}
template <class T, class V>
RDB_HD auto chain(T v, const T* c, const V& in) {
    constexpr mask_t R = vec_mask<T, V>();
    if constexpr (R == 0) return v;
    else { SD<T, R> r; r.v = v; chain_parts<T, R, V>(r, c, in); return r; }
}
template <class T, class V> __host__ __device__ constexpr bool has_partials() { return vec_mask<T, V>() != 0; }

// ---------------------------------------------------------------------------------------------
// Seeding and extraction.
// ---------------------------------------------------------------------------------------------
// element I of [x;u]: a seeded dual if column I belongs to this thread's column chunk, a plain T otherwise
template <class T, int I, mask_t CHUNK>
RDB_HD auto seed(T z) {
    if constexpr (chas(CHUNK, I)) { SD<T, (mask_t(1) << I)> r; r.v = z; r.zero_parts(); r.template set_part<I>(T(1)); return r; }
    else return z;
}
template <int J, class T> RDB_HD T partial(const T&) { return T(0); }
template <int J, class T, mask_t M> RDB_HD T partial(const SD<T, M>& a) { return a.template part<J>(); }

// ---------------------------------------------------------------------------------------------
// Widening: re-type an element / vector to a superset mask (new partials are explicit zeros), so that a rolled
// loop over integrator stages can carry ONE type (see integrators.cuh: saturated stage types).
// ---------------------------------------------------------------------------------------------
template <class To> struct widen_to;
template <class T> struct widen_to {   // plain target
    RDB_HD static T from(const T& a) { return a; }
};
template <class T, mask_t B> struct widen_to<SD<T, B>> {
    RDB_HD static SD<T, B> from(const T& a) { SD<T, B> r; r.v = a; r.zero_parts(); return r; }
    template <mask_t A, int J = 0>
    RDB_HD static void fill(SD<T, B>& r, const SD<T, A>& a) {
        constexpr mask_t SA = smask<T>(A), SB = smask<T>(B);
        if constexpr ((SB >> J) != 0) {
            if constexpr (chas(SB, J)) { if constexpr (chas(SA, J)) r.d[cslot(SB, J)] = a.d[cslot(SA, J)]; else r.d[cslot(SB, J)] = PK<T>::zero(); }
            fill<A, J + 1>(r, a);
        }
    }
    template <mask_t A> RDB_HD static SD<T, B> from(const SD<T, A>& a) {
        static_assert((A & ~B) == 0, "widen: target mask must contain the source mask");
        SD<T, B> r; r.v = a.v; fill<A>(r, a); return r;
    }
};
template <class To, class From, size_t... Is>
RDB_HD To widen_vec_impl(const From& a, rstd::index_sequence<Is...>) {
    return To(widen_to<rstd::remove_cv_t<rstd::remove_reference_t<decltype(get<int(Is)>(rstd::declval<const To&>()))>>>::from(get<int(Is)>(a))...);
}
template <class To, class From> RDB_HD To widen_vec(const From& a) { return widen_vec_impl<To>(a, iseq<To>{}); }
template <class To, class T, size_t... Is>
RDB_HD To zero_vec_impl(rstd::index_sequence<Is...>) {
    return To(widen_to<rstd::remove_cv_t<rstd::remove_reference_t<decltype(get<int(Is)>(rstd::declval<const To&>()))>>>::from(T(0))...);
}
template <class To, class T> RDB_HD To zero_vec() { return zero_vec_impl<To, T>(iseq<To>{}); }

}  // namespace rdb
