// launch.cuh — per-(model, dtype) tiling configuration and the host-side launcher of knot_kernel.
#pragma once
#include <cstdio>
#include "kernels.cuh"

namespace rdb {

// ---- column-chunk generation ----------------------------------------------------------------------------------
__host__ __device__ constexpr mask_t range_mask(int lo, int hi) { return (hi >= 32 ? ~mask_t(0) : ((mask_t(1) << hi) - 1u)) & ~((mask_t(1) << lo) - 1u); }
__host__ __device__ constexpr mask_t chunk_mask(int NZ, int cpr, int r) { return range_mask(r * cpr, (r + 1) * cpr < NZ ? (r + 1) * cpr : NZ); }
template <int NZ, int CPR, class Seq> struct gen_chunks;
template <int NZ, int CPR, size_t... Rs> struct gen_chunks<NZ, CPR, std::index_sequence<Rs...>> { using type = MaskList<chunk_mask(NZ, CPR, int(Rs))...>; };
template <int NZ, int CPR> using uniform_chunks = typename gen_chunks<NZ, CPR, std::make_index_sequence<size_t((NZ + CPR - 1) / CPR)>>::type;

// ---- tiling configuration ---------------------------------------------------------------------------------------
// TILE  knots per CTA tile (multiple of 32);  Chunks  column chunks = roles;  MINB  CTAs/SM promised to ptxas.
template <class Model, class T, bool WITH_J, class Enable = void>
struct KnotConfig {   // small models (Cartpole, double integrators) and every value-only kernel: one role
    static constexpr int TILE = 128, MINB = (sizeof(T) == 8 ? 3 : 4);
    using Chunks = MaskList<WITH_J ? range_mask(0, Model::n + Model::m) : mask_t(0)>;
};
// rigid bodies with Jacobians, fp32
template <class Model>
struct KnotConfig<Model, float, true, std::enable_if_t<(Model::n >= 12)>> {
    static constexpr int NZ = Model::n + Model::m;
    static constexpr int TILE = 64, MINB = 2;
    using Chunks = std::conditional_t<NZ == 17, MaskList<0x7Fu, 0x1F80u, 0x1E000u>,            // {r,q} {v,w} {u}
                   std::conditional_t<NZ == 16, MaskList<0x3Fu, 0xFC0u, 0xF000u>,             // {r,p} {v,w} {u}
                   std::conditional_t<NZ == 19, MaskList<0x7Fu, 0x1F80u, 0xE000u, 0x70000u>,  // {r,q} {v,w} {u0-2} {u3-5}
                                                MaskList<0x3Fu, 0xFC0u, 0x7000u, 0x38000u>>>>; // NZ == 18
};
// rigid bodies with Jacobians, fp64 (twice the registers per value: narrower chunks)
template <class Model>
struct KnotConfig<Model, double, true, std::enable_if_t<(Model::n >= 12)>> {
    static constexpr int NZ = Model::n + Model::m;
    static constexpr int TILE = 32, MINB = 1;
    using Chunks = uniform_chunks<NZ, 3>;
};

struct DeviceInfo { int device; int sm_count; };

template <class Model, int Q, class T, bool WITH_J>
struct KnotLaunch {
    using Cfg = KnotConfig<Model, T, WITH_J>;
    using S = KnotSmem<Model, Cfg::TILE, WITH_J, T>;
    static constexpr int NTHR = Cfg::TILE * Cfg::Chunks::count;
    static int run(const Model& model, const KnotArgs<T>& a, const DeviceInfo& dev, cudaStream_t st) {
        auto kern = knot_kernel<Model, Q, T, Cfg::TILE, WITH_J, typename Cfg::Chunks, Cfg::MINB>;
        static int occ_cache[64];   // CTAs/SM per device id; 0 = not yet configured on that device
        const int d = dev.device & 63;
        if (occ_cache[d] == 0) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(S::total));
            if (e != cudaSuccess) return int(e);
            int occ = 0;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NTHR, S::total);
            if (e != cudaSuccess) return int(e);
            occ_cache[d] = occ > 0 ? occ : 1;
        }
        if (a.N <= 0) return 0;
        const long long ntiles = (a.N + Cfg::TILE - 1) / Cfg::TILE;
        const long long cap = (long long)dev.sm_count * occ_cache[d];
        const unsigned grid = unsigned(ntiles < cap ? ntiles : cap);
        kern<<<grid, NTHR, S::total, st>>>(model, a);
        return int(cudaGetLastError());
    }
};

// ---- request passed from the C ABI to a per-model compilation unit ------------------------------------------------
enum UnitOp { OP_KNOT = 0, OP_ROLLOUT = 1 };
struct KnotRequest {
    int op;              // UnitOp
    int Q;               // QuadRule (Q_CONTINUOUS for dynamics / continuous Jacobian)
    int dtype;           // 0 = f32, 1 = f64
    int with_j;
    ModelParams<double> params;
    const void* Z; const double* dt; double dt0; void* J; void* out; long long N; int layout;
    // OP_ROLLOUT: x0 (n, ntraj), U (m, K-1, ntraj), dt (K, ntraj) or null, X (n, K, ntraj)
    const void* x0; const void* U; void* X; long long ntraj; int K;
    DeviceInfo dev;
    cudaStream_t stream;
};

// rollout!: x_{k+1} = discrete_dynamics(x_k, u_k, t_k, dt_k), sequential in k, one thread per trajectory
// (reference: src/trajectories.jl:436-441, src/discrete_dynamics.jl:217-235).
template <class T, size_t... Is> __device__ __forceinline__ auto load_plain(const T* p, std::index_sequence<Is...>) { return vec(p[Is]...); }
template <class Model, int Q, class T>
__global__ void __launch_bounds__(64) rollout_kernel(const Model model, const T* __restrict__ x0, const T* __restrict__ U,
                                                     const double* __restrict__ dt, double dt0, T* __restrict__ X,
                                                     long long ntraj, int K) {
    constexpr int n = Model::n, m = Model::m;
    const long long tr = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (tr >= ntraj) return;
    auto x = load_plain(x0 + tr * n, std::make_index_sequence<size_t(n)>{});
    T* Xt = X + tr * (long long)K * n;
    put_vals(x, Xt, std::make_index_sequence<size_t(n)>{});
    for (int k = 0; k + 1 < K; ++k) {
        auto u = load_plain(U + (tr * (long long)(K - 1) + k) * m, std::make_index_sequence<size_t(m)>{});
        const T h = T(dt ? dt[tr * K + k] : dt0);
        x = integrate<Q, T>(model, x, u, h);
        put_vals(x, Xt + (long long)(k + 1) * n, std::make_index_sequence<size_t(n)>{});
    }
}

template <class T>
inline ModelParams<T> cast_params(const ModelParams<double>& p) {
    ModelParams<T> q;
    q.mc = T(p.mc); q.mp = T(p.mp); q.l = T(p.l); q.g = T(p.g);
    q.mass = T(p.mass); q.inv_mass = T(p.inv_mass);
    for (int i = 0; i < 9; ++i) { q.J[i] = T(p.J[i]); q.Jinv[i] = T(p.Jinv[i]); }
    for (int i = 0; i < 3; ++i) q.mg[i] = T(p.mg[i]);
    q.motor_dist = T(p.motor_dist); q.kf = T(p.kf); q.km = T(p.km);
    return q;
}

template <template <class> class ModelT, class T, int Q, bool WITH_J>
inline int run_one(const KnotRequest& r) {
    ModelT<T> model; model.p = cast_params<T>(r.params);
    KnotArgs<T> a;
    a.Z = static_cast<const T*>(r.Z); a.dt = r.dt; a.dt0 = r.dt0;
    a.J = static_cast<T*>(r.J); a.out = static_cast<T*>(r.out); a.N = r.N; a.layout = r.layout;
    return KnotLaunch<ModelT<T>, Q, T, WITH_J>::run(model, a, r.dev, r.stream);
}
template <template <class> class ModelT, class T, int Q>
inline int run_rollout(const KnotRequest& r) {
    if constexpr (Q == Q_CONTINUOUS) return -2;
    else {
        ModelT<T> model; model.p = cast_params<T>(r.params);
        if (r.ntraj <= 0 || r.K <= 0) return 0;
        const unsigned grid = unsigned((r.ntraj + 63) / 64);
        rollout_kernel<ModelT<T>, Q, T><<<grid, 64, 0, r.stream>>>(model, static_cast<const T*>(r.x0), static_cast<const T*>(r.U),
                                                                  r.dt, r.dt0, static_cast<T*>(r.X), r.ntraj, r.K);
        return int(cudaGetLastError());
    }
}
template <template <class> class ModelT, class T, int Q>
inline int run_q(const KnotRequest& r) {
    if (r.op == OP_ROLLOUT) return run_rollout<ModelT, T, Q>(r);
    return r.with_j ? run_one<ModelT, T, Q, true>(r) : run_one<ModelT, T, Q, false>(r);
}
template <template <class> class ModelT, class T>
inline int run_t(const KnotRequest& r) {
    switch (r.Q) {
        case Q_EULER: return run_q<ModelT, T, Q_EULER>(r);
        case Q_RK2: return run_q<ModelT, T, Q_RK2>(r);
        case Q_RK3: return run_q<ModelT, T, Q_RK3>(r);
        case Q_RK4: return run_q<ModelT, T, Q_RK4>(r);
        case Q_CONTINUOUS: return run_q<ModelT, T, Q_CONTINUOUS>(r);
    }
    return -2;
}
template <template <class> class ModelT>
inline int run_model(const KnotRequest& r) { return r.dtype == 0 ? run_t<ModelT, float>(r) : run_t<ModelT, double>(r); }

}  // namespace rdb
