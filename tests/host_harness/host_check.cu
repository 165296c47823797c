// host_check.cu — TEST INFRASTRUCTURE: instantiates the product's own model / integrator / sparse-dual templates (the headers the
// CUDA kernels are compiled from: csrc/sdual.cuh, models.cuh, integrators.cuh) in a plain HOST program, so that the -m "not gpu"
// suite can compare the arithmetic the kernels execute — elemental operations, rolled stage loops, increment forms, time-varying
// stage times — with the CPU checker without a GPU.  Nothing in the package uses this; the product path has no CPU fallback.
//
//   nvcc -std=c++17 -O1 -I robotdynamics.jl_b200/csrc -DHK=<kind> -DHR=<rot> -DHF=<frame> -DHQ=<rule> -DHROLL=<0|1|2> -o hc host_check.cu
//   stdin : 16 parameters (cartpole: mc mp l g then zeros; rigid: mass J(9) m*g(3) motor_dist kf km), N, h, then N rows of [x;u]
//   stdout: per row the n x (n+m) column-major Jacobian followed by x+ (n values)
#include <cstdio>
#include <utility>
#include "integrators.cuh"
using namespace rdb;
using T = double;
#if HK == 0
using Model = Cartpole<T>;
#else
using Model = RigidBody<T, HK, HR, HF>;
#endif
template <int N_, int J, class XN, size_t... Is> void put1(const XN& xn, double* Jm, std::index_sequence<Is...>) { ((Jm[int(Is) + N_ * J] = partial<J>(get<int(Is)>(xn))), ...); }
template <int N_, class XN, size_t... Js> void putall(const XN& xn, double* Jm, std::index_sequence<Js...>) { (put1<N_, int(Js)>(xn, Jm, std::make_index_sequence<size_t(N_)>{}), ...); }
template <class XN, size_t... Is> void putvals(const XN& xn, double* v, std::index_sequence<Is...>) { ((v[Is] = val(get<int(Is)>(xn))), ...); }
template <class T_, mask_t CH, size_t... Is> auto seeded(const T_* z, std::index_sequence<Is...>) { return vec(seed<T_, int(Is), CH>(z[Is])...); }
int main() {
    constexpr int n = Model::n, m = Model::m, NZ = n + m;
    constexpr mask_t ALL = (mask_t(1) << NZ) - 1u;
    Model model;
    ModelParams<T>& p = model.p;
    double pr[16];
    for (int i = 0; i < 16; ++i) if (scanf("%lf", &pr[i]) != 1) return 1;
#if HK == 0
    p.mc = pr[0]; p.mp = pr[1]; p.l = pr[2]; p.g = pr[3];
    p.cp_ia = 1.0 / (p.mp * p.l); p.cp_H00 = (p.mc + p.mp) * p.cp_ia; p.cp_nH00i = -1.0 / p.cp_H00;       // as rdb_model_create does
#else
    p.mass = pr[0]; p.inv_mass = 1.0 / pr[0];
    for (int i = 0; i < 9; ++i) { p.J[i] = pr[1 + i]; p.Jinv[i] = 0; }
    p.Jinv[0] = 1 / p.J[0]; p.Jinv[4] = 1 / p.J[4]; p.Jinv[8] = 1 / p.J[8];                                // the test inertias are diagonal
    for (int i = 0; i < 3; ++i) p.mg[i] = pr[10 + i];
    p.motor_dist = pr[13]; p.kf = pr[14]; p.km = pr[15];
#endif
    int N; double h;
    if (scanf("%d %lf", &N, &h) != 2) return 1;
    for (int k = 0; k < N; ++k) {
        T z[NZ];
        for (int i = 0; i < NZ; ++i) if (scanf("%lf", &z[i]) != 1) return 1;
        model.reset();
        auto zz = seeded<T, ALL>(z, std::make_index_sequence<size_t(NZ)>{});
        auto xn = integrate<HQ, T, HROLL>(model, slice<0, n>(zz), slice<n, m>(zz), h);
        double Jm[n * NZ], xv[n];
        putall<n>(xn, Jm, std::make_index_sequence<size_t(NZ)>{});
        putvals(xn, xv, std::make_index_sequence<size_t(n)>{});
        for (int i = 0; i < n * NZ; ++i) printf("%.17g ", Jm[i]);
        for (int i = 0; i < n; ++i) printf("%.17g ", xv[i]);
        printf("\n");
    }
    return 0;
}
