// launch.cuh — per-(model, dtype) tiling configuration and the host-side launcher of knot_kernel.
#pragma once
#include <atomic>
#include <cstdlib>
#include <cuda.h>            // CUtensorMap types only; the encoder is fetched through the runtime (no libcuda link dependency)
#include "kernels.cuh"
#include "implicit_block.cuh"

namespace rdb {

// ---- column-chunk generation ----------------------------------------------------------------------------------
__host__ __device__ constexpr mask_t range_mask(int lo, int hi) { return (hi >= 32 ? ~mask_t(0) : ((mask_t(1) << hi) - 1u)) & ~((mask_t(1) << lo) - 1u); }
__host__ __device__ constexpr mask_t chunk_mask(int NZ, int cpr, int r) { return range_mask(r * cpr, (r + 1) * cpr < NZ ? (r + 1) * cpr : NZ); }
template <int NZ, int CPR, class Seq> struct gen_chunks;
template <int NZ, int CPR, size_t... Rs> struct gen_chunks<NZ, CPR, std::index_sequence<Rs...>> { using type = MaskList<chunk_mask(NZ, CPR, int(Rs))...>; };
template <int NZ, int CPR> using uniform_chunks = typename gen_chunks<NZ, CPR, std::make_index_sequence<size_t((NZ + CPR - 1) / CPR)>>::type;

// ---- tiling configuration ---------------------------------------------------------------------------------------
// TILE   knots per CTA tile (multiple of 32)
// Chunks column chunks = roles (one warp-uniform code path each)
// MINB   CTAs/SM promised to ptxas (register cap = 65536 / (threads * MINB))
// ROLL   RK4 stage loop: 0 fully unrolled, 1 stage 1 sparse + stages 2..4 rolled, 2 all four rolled (integrators.cuh)
// Tuning overrides for experiments (scripts/tune.py): -DRDB_TUNE_TILE / _MINB / _ROLL / _C0.._C8 (chunk masks) apply to the
// Jacobian kernels of the unit being compiled.
template <class Model, class T, bool WITH_J, int Q, class Enable = void>
struct KnotConfigDefault {   // small models (Cartpole, double integrators) and every value-only kernel: one role.
    // Jacobians of small models: single-warp CTAs (TILE 32), 12 per SM — warps drift apart instead of hitting the FP64-heavy and
    // FP64-free phases of a tile in lockstep (Cartpole RK4 fp64: 47.5 -> 45.1 us, profiles/tuning_r01.md).
    // Round 2 (elemental Cartpole, 112 registers): 64-knot tiles, 8 CTAs per SM measured best (38.3 us against 39.4 us for the
    // single-warp form, 45 us with 18 CTAs of 96 registers + spills; profiles/tuning_r02.md).
    static constexpr int TILE = WITH_J ? 64 : 128, MINB = WITH_J ? 8 : 4, ROLL = 0;
    using Chunks = MaskList<WITH_J ? range_mask(0, Model::n + Model::m) : mask_t(0)>;
};
// Rigid bodies with Jacobians.  Measured on B200 (profiles/tuning_r01.md): few wide roles beat many narrow ones (every role
// recomputes the stage values), as long as the role's live partials fit the register file; RK4 is rolled (ROLL=1).
template <class Model, int Q>
struct KnotConfigDefault<Model, float, true, Q, std::enable_if_t<(Model::n >= 12)>> {
    static constexpr int n = Model::n, m = Model::m, NZ = n + m;
    static constexpr bool heavy = (Q == Q_RK3 || Q == Q_RK4);
    static constexpr int ROLL = heavy ? 1 : 0;
    static constexpr int TILE = heavy ? 128 : 64;
    static constexpr int MINB = heavy ? 1 : 2;
    // RK3 / RK4: two wide roles.  World-frame quaternion models split after w1 ({r,q,v,w0,w1} {w2,u}: C3 51 us, Body 67 -> 58 us);
    // body-frame and 3-parameter attitudes couple more rows to v and w and balance better split after v ({r,att,v} {w,u}:
    // quadrotor body frame 162 -> 137 us, quadrotor{MRP} 78 -> 74 us; profiles/tuning_r01.md).
    // After the split body-frame force (models.cuh) the body-frame rows are lighter and every body-frame model balances best split after
    // w1 as well: quadrotor body frame 85.8 -> 76.2 us, quadrotor{MRP} body frame 80.8 -> 78.9 us, Body{Quat} body frame 71.6 -> 71.7 us
    // (profiles/tuning_r02.md); only the world-frame 3-parameter attitudes keep {r,att,v} {w,u}.
    static constexpr int split = (Model::rot == ROT_QUAT || Model::frame == FRAME_BODY) ? n - 1 : n - 3;
    using Heavy = MaskList<range_mask(0, split), range_mask(split, NZ)>;
    using Light = std::conditional_t<m == 4, MaskList<range_mask(0, n - 6), range_mask(n - 6, n), range_mask(n, NZ)>,            // {r,att} {v,w} {u}
                                              MaskList<range_mask(0, n - 6), range_mask(n - 6, n), range_mask(n, n + 3), range_mask(n + 3, NZ)>>;
    using Chunks = std::conditional_t<heavy, Heavy, Light>;
};
// fp64: twice the registers per value -> narrower roles for the long RK3/RK4 chains
template <class Model, int Q>
struct KnotConfigDefault<Model, double, true, Q, std::enable_if_t<(Model::n >= 12)>> {
    static constexpr int n = Model::n, m = Model::m, NZ = n + m;
    static constexpr bool heavy = (Q == Q_RK3 || Q == Q_RK4);
    static constexpr int ROLL = heavy ? 1 : 0;
    static constexpr int TILE = 64;
    static constexpr int MINB = heavy ? 1 : 2;
    // four roles balanced by the number of non-trivial entries they carry (position and velocity columns are nearly free):
    // {r, att[0..np-2]} {att[np-1], v, w0} {w1, w2, u0(,u1)} {rest of u}   (quadrotor fp64 RK4: 126 -> 118 us)
    static constexpr int c3 = n + (m > 4 ? (m - 2) / 2 : 1);
    using Heavy = std::conditional_t<n == 13,
        MaskList<range_mask(0, n - 7), range_mask(n - 7, n - 2), range_mask(n - 2, c3), range_mask(c3, NZ)>,
        MaskList<range_mask(0, n - 6), range_mask(n - 6, n), range_mask(n, n + m / 2), range_mask(n + m / 2, NZ)>>;   // 3-parameter attitudes: {r,att} {v,w} {u lo} {u hi} measured better
    using Light = MaskList<range_mask(0, NZ / 2), range_mask(NZ / 2, NZ)>;
    using Chunks = std::conditional_t<heavy, Heavy, Light>;
};

template <class Model, class T, bool WITH_J, int Q>
struct KnotConfig : KnotConfigDefault<Model, T, WITH_J, Q> {};
#if defined(RDB_TUNE_V_TILE) || defined(RDB_TUNE_V_MINB)      // tuning experiments on the value-only kernels (scripts/tune.py)
template <class Model, class T, int Q>
struct KnotConfig<Model, T, false, Q> {
    using D = KnotConfigDefault<Model, T, false, Q>;
#ifdef RDB_TUNE_V_TILE
    static constexpr int TILE = RDB_TUNE_V_TILE;
#else
    static constexpr int TILE = D::TILE;
#endif
#ifdef RDB_TUNE_V_MINB
    static constexpr int MINB = RDB_TUNE_V_MINB;
#else
    static constexpr int MINB = D::MINB;
#endif
    static constexpr int ROLL = 0;
    using Chunks = typename D::Chunks;
};
#endif

// Small batches (a solver's 10^2 - 10^3 knots per call): one wave of CTAs, the launch lasts as long as ONE thread's dependent chain.
// The long rigid-body RK3 / RK4 Jacobian chains then run faster on 32-knot tiles — four times as many CTAs, each warp on a scheduler of
// its own instead of two per scheduler on few SMs — with the same roles (graph-replayed launch, Quadrotor RK4 fp32: 5.3 -> 3.2 us at
// N <= 1024, 5.6 -> 4.3 us at 4096, equal at 16384; profiles/small_batch_r02.md).  Above RDB_SMALL_N the wide tiles return, unless the
// family prefers the narrow ones at every size (below).  `distinct` = this configuration exists as its own instantiation.
#ifndef RDB_SMALL_N
#define RDB_SMALL_N 8192
#endif
template <class Model, class T, bool WITH_J, int Q, class Enable = void>
struct KnotConfigSmall : KnotConfig<Model, T, WITH_J, Q> { static constexpr bool distinct = false; };
// ... and some of them at EVERY size (measured at 262144 knots, all 12 variants x dtype, profiles/tuning_r02.md: four 64-thread CTAs per SM
// overlap their load / compute / store phases better than one 256-thread CTA whose roles meet at one barrier): world-frame quaternion
// models in fp32 (C3 45.5 -> 42.6 us, its error-state form 45.3 -> 42.8 us, Body{Quat} 51.0 -> 50.0 us) and the Body / Satellite family
// with a 3-parameter attitude in fp64 or in the body frame (Body{MRP} body frame 74.0 -> 69.1 us fp32, 211 -> 191 us fp64; Body{MRP} fp64
// 146.7 -> 139.6 us).  Every other variant is faster on the wide tiles from ~16k knots on (quadrotor{MRP} fp64 141.9 vs 157.7 us).
// The fused error-state kernels (ERR) were measured separately (profiles/err_tiles_r02.md): narrow wins everywhere except the quadrotor in
// the body frame (fp32: 79.0 vs 84.3 us) and, in fp64, the quadrotor's body-frame and 3-parameter variants (150.5 vs 179.6 us).
template <class Model, class T, bool ERR>
constexpr bool prefers_narrow_tiles() {
    constexpr bool quad = (Model::m == 4), quat_world = (Model::rot == ROT_QUAT && Model::frame == FRAME_WORLD);
    if (ERR) {
        if (sizeof(T) == 4) return !(quad && Model::frame == FRAME_BODY);
        return !quad || quat_world;
    }
    if (quat_world) return sizeof(T) == 4;
    if (!quad && Model::rot != ROT_QUAT) return sizeof(T) == 8 || Model::frame == FRAME_BODY;
    return false;
}
#ifndef RDB_TUNE_TILE
template <class Model, class T, int Q>
struct KnotConfigSmall<Model, T, true, Q, std::enable_if_t<(Model::n >= 12) && (Q == Q_RK3 || Q == Q_RK4)>> {
    using D = KnotConfig<Model, T, true, Q>;
    static constexpr int TILE = 32, MINB = 1, ROLL = D::ROLL;
    using Chunks = typename D::Chunks;
    static constexpr bool distinct = (D::TILE != 32);
};
#endif

template <mask_t... Acc> struct nz_build {
    template <mask_t M> using add = std::conditional_t<M != 0, nz_build<Acc..., M>, nz_build<Acc...>>;
    using type = MaskList<Acc...>;
};
#ifndef RDB_TUNE_C1
#define RDB_TUNE_C1 0
#endif
#ifndef RDB_TUNE_C2
#define RDB_TUNE_C2 0
#endif
#ifndef RDB_TUNE_C3
#define RDB_TUNE_C3 0
#endif
#ifndef RDB_TUNE_C4
#define RDB_TUNE_C4 0
#endif
#ifndef RDB_TUNE_C5
#define RDB_TUNE_C5 0
#endif
#ifndef RDB_TUNE_C6
#define RDB_TUNE_C6 0
#endif
#ifndef RDB_TUNE_C7
#define RDB_TUNE_C7 0
#endif
#ifndef RDB_TUNE_C8
#define RDB_TUNE_C8 0
#endif
#if defined(RDB_TUNE_TILE) || defined(RDB_TUNE_MINB) || defined(RDB_TUNE_ROLL) || defined(RDB_TUNE_C0)
template <class Model, class T, int Q>
struct KnotConfig<Model, T, true, Q> {
    using D = KnotConfigDefault<Model, T, true, Q>;
#ifdef RDB_TUNE_TILE
    static constexpr int TILE = RDB_TUNE_TILE;
#else
    static constexpr int TILE = D::TILE;
#endif
#ifdef RDB_TUNE_MINB
    static constexpr int MINB = RDB_TUNE_MINB;
#else
    static constexpr int MINB = D::MINB;
#endif
#ifdef RDB_TUNE_ROLL
    static constexpr int ROLL = RDB_TUNE_ROLL;
#else
    static constexpr int ROLL = D::ROLL;
#endif
#ifdef RDB_TUNE_C0
    using Chunks = typename nz_build<>::add<RDB_TUNE_C0>::add<RDB_TUNE_C1>::add<RDB_TUNE_C2>::add<RDB_TUNE_C3>::add<RDB_TUNE_C4>
        ::add<RDB_TUNE_C5>::add<RDB_TUNE_C6>::add<RDB_TUNE_C7>::add<RDB_TUNE_C8>::type;
#else
    using Chunks = typename D::Chunks;
#endif
};
#endif

struct DeviceInfo { int device; int sm_count; int pdl; };   // pdl: launch with programmatic stream serialization

// ---- tensor map of J for the padded-image store (kernels.cuh: tensor_store_2d) -------------------------------------------------
// J is described as a 2-D tensor: inner extent E = rows x cols of one knot's Jacobian, outer extent N knots, dense.  The box is
// pitch x tile with pitch > E: the pad of every smem row lies outside the tensor and is not written.
#ifndef RDB_SOA_KERNELS
#define RDB_SOA_KERNELS 1        // component-major kernels (tensor-map loads / stores); 0 leaves only the transposing path
#endif
#ifndef RDB_TUNE_JMAP
#define RDB_TUNE_JMAP 1          // 0: tuning experiments only (one bulk store per knot row instead)
#endif
typedef CUresult (*rdb_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline rdb_encode_tiled_fn encode_tiled_entry() {
    static rdb_encode_tiled_fn fn = []() -> rdb_encode_tiled_fn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        return reinterpret_cast<rdb_encode_tiled_fn>(p);
    }();
    return fn;
}
static_assert(sizeof(TensorMap) == sizeof(CUtensorMap) && alignof(TensorMap) >= alignof(CUtensorMap), "TensorMap must mirror CUtensorMap");
// a dense 2-D tensor (inner extent d0 contiguous, outer extent d1 at a pitch of ld elements) with a box of b0 x b1 elements
inline int encode_map2d(TensorMap* tm, const void* base, long long d0, long long d1, long long ld, int b0, int b1, int es) {
    if (!base || b0 > 256 || b1 > 256 || (reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld * es) % 16 != 0) return 0;
    rdb_encode_tiled_fn enc = encode_tiled_entry();
    if (!enc) return 0;
    const cuuint64_t gdim[2] = {cuuint64_t(d0), cuuint64_t(d1)};
    const cuuint64_t gstride[1] = {cuuint64_t(ld) * cuuint64_t(es)};
    const cuuint32_t box[2] = {cuuint32_t(b0), cuuint32_t(b1)};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult rc = enc(reinterpret_cast<CUtensorMap*>(tm), es == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2,
                            const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return rc == CUDA_SUCCESS ? 1 : 0;
}
// returns 1 and fills *tm when J (16-byte aligned, at least one full tile) can leave through a tensor map, 0 otherwise
inline int encode_jmap(TensorMap* tm, void* J, long long N, int E, int pitch, int tile, int es) {
    if (!RDB_TUNE_JMAP || N < tile) return 0;
    return encode_map2d(tm, J, E, N, E, pitch, tile, es);
}
// can component-major arrays of leading dimension ld (knots per component row) go through the tensor-map kernels?
inline bool soa_tma_ok(const void* Z, const void* J, const void* out, long long ld, int es) {
    return encode_tiled_entry() != nullptr && (ld * es) % 16 == 0 &&
           ((reinterpret_cast<uintptr_t>(Z) | reinterpret_cast<uintptr_t>(J) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
}

// shape seen by the tiling rules in error-state mode: nerr rows, nerr + m columns (like an n = 12 model)
template <class Model> struct ErrShape { static constexpr int n = Model::nerr, m = Model::m, rot = ROT_QUAT, frame = FRAME_WORLD; };

template <class Model, int Q, class T, bool WITH_J, bool ERR = false, bool SOA = false, bool SMALL = false>
struct KnotLaunch {
    using Shape = std::conditional_t<ERR, ErrShape<Model>, Model>;
    using Cfg = std::conditional_t<SMALL, KnotConfigSmall<Shape, T, WITH_J, Q>, KnotConfig<Shape, T, WITH_J, Q>>;
    using S = KnotSmem<Model, Cfg::TILE, WITH_J, T, ERR>;
    static constexpr int NTHR = Cfg::TILE * Cfg::Chunks::count;
    // ld: knots per component row of the caller's arrays (component-major kernels only)
    static int run(const Model& model, const KnotArgs<T>& a, const DeviceInfo& dev, cudaStream_t st, long long ld = 0) {
        auto kern = knot_kernel<Model, Q, T, Cfg::TILE, WITH_J, typename Cfg::Chunks, Cfg::MINB, Cfg::ROLL, ERR, SOA>;
        // CTAs/SM per device id; 0 = not yet configured on that device.  Atomic: two host threads may race on the first use of a
        // kernel — both then set the same attribute and store the same value (idempotent), and nobody reads a torn entry.
        static std::atomic<int> occ_cache[64];
        const int d = dev.device & 63;
        int occ_d = occ_cache[d].load(std::memory_order_acquire);
        if (occ_d == 0) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(S::total));
            if (e != cudaSuccess) return int(e);
            int occ = 0;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NTHR, S::total);
            if (e != cudaSuccess) return int(e);
            occ_d = occ > 0 ? occ : 1;
            occ_cache[d].store(occ_d, std::memory_order_release);
        }
        if (a.N <= 0) return 0;
        const long long ntiles = (a.N + Cfg::TILE - 1) / Cfg::TILE;
        const long long cap = (long long)dev.sm_count * occ_d;
        const unsigned grid = unsigned(ntiles < cap ? ntiles : cap);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(NTHR); cfg.dynamicSmemBytes = S::total; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = dev.pdl ? 1 : 0;
        if constexpr (SOA) {
            KnotArgs<T> b = a;
            constexpr int es = int(sizeof(T));
            int ok = encode_map2d(&b.zmap, b.Z, b.N, S::NZ, ld, Cfg::TILE, S::NZ, es);
            if (b.J) ok &= encode_map2d(&b.jmap, b.J, b.N, S::E, ld, Cfg::TILE, S::E, es);
            if (b.out) ok &= encode_map2d(&b.omap, b.out, b.N, S::n, ld, Cfg::TILE, S::n, es);
            if (!ok) return -2;
            return int(cudaLaunchKernelEx(&cfg, kern, model, b));
        } else if constexpr (S::ROWSTORE) {
            KnotArgs<T> b = a;
            b.use_jmap = encode_jmap(&b.jmap, b.J, b.N, S::E, S::PJ, Cfg::TILE, int(sizeof(T)));
            return int(cudaLaunchKernelEx(&cfg, kern, model, b));
        }
        return int(cudaLaunchKernelEx(&cfg, kern, model, a));
    }
};

// ---- request passed from the C ABI to a per-model compilation unit ------------------------------------------------
enum UnitOp { OP_KNOT = 0, OP_ROLLOUT = 1, OP_DYNERR = 2 };
struct KnotRequest {
    int op;              // UnitOp
    int Q;               // QuadRule (Q_CONTINUOUS for dynamics / continuous Jacobian)
    int dtype;           // 0 = f32, 1 = f64
    int with_j;
    int err;             // error-state Jacobian  G(x+)' [A B] blkdiag(G(x), I)  (rigid bodies; needs with_j)
    ModelParams<double> params;
    const void* Z; const double* dt; double dt0; void* J; void* out; long long N;   // knot-major, or component-major when soa != 0
    const double* t;         // per-knot times (N) or null; only time-varying (user) models read them
    int whole_sm;            // plans that share the GPU with other kernels (rdb_plan_set_shared): keep the wide, SM-filling tiles
    int soa; long long ld;   // component-major arrays with ld knots per component row, served by the tensor-map kernels
    // OP_ROLLOUT: x0 (n, ntraj), U (m, K-1, ntraj), dt (K, ntraj) or null, X (n, K, ntraj)   [trajectory-major, zmode == 0]
    //             zmode != 0: X is the knot-major batch Z (K, ntraj, n+m) holding the controls; steps [kb, ke)   (kernels.cuh)
    const void* x0; const void* U; void* X; long long ntraj; int K;
    int zmode, kb, ke;
    int rollout_block;   // threads per CTA of the rollout kernel (0 -> 32: one warp per CTA, spread over the SMs)
    // OP_DYNERR (ImplicitMidpoint): dynamics_error / dynamics_error_jacobian! of the pairs (Z[k], Z2[k]): Z2 (N, ld2) holds x2, J2 -> de/dz2,
    // J -> de/dz1, out -> e
    const void* Z2; int ld2; void* J2;
    DeviceInfo dev;
    cudaStream_t stream;
};

template <class T>
inline ModelParams<T> cast_params(const ModelParams<double>& p) {
    ModelParams<T> q;
    q.mc = T(p.mc); q.mp = T(p.mp); q.l = T(p.l); q.g = T(p.g);
    q.cp_ia = T(p.cp_ia); q.cp_H00 = T(p.cp_H00); q.cp_nH00i = T(p.cp_nH00i);
    q.mass = T(p.mass); q.inv_mass = T(p.inv_mass);
    for (int i = 0; i < 9; ++i) { q.J[i] = T(p.J[i]); q.Jinv[i] = T(p.Jinv[i]); }
    for (int i = 0; i < 3; ++i) q.mg[i] = T(p.mg[i]);
    q.motor_dist = T(p.motor_dist); q.kf = T(p.kf); q.km = T(p.km);
    return q;
}

template <template <class> class ModelT, class T, int Q, bool WITH_J, bool ERR = false>
inline int run_one(const KnotRequest& r) {
    ModelT<T> model; model.p = cast_params<T>(r.params);
    KnotArgs<T> a;
    a.Z = static_cast<const T*>(r.Z); a.dt = r.dt; a.dt0 = r.dt0; a.t = r.t;
    a.J = static_cast<T*>(r.J); a.out = static_cast<T*>(r.out); a.N = r.N; a.use_jmap = 0;
    {
        constexpr long long JR = ERR ? ModelT<T>::nerr : ModelT<T>::n, JC = JR + ModelT<T>::m;
        a.stream_out = knot_stream_out(r.N, (long long)sizeof(T) * ((WITH_J && r.J ? JR * JC : 0) + (r.out ? ModelT<T>::n : 0))) ? 1 : 0;
    }
    if (r.soa) {
        if constexpr (RDB_SOA_KERNELS) return KnotLaunch<ModelT<T>, Q, T, WITH_J, ERR, true>::run(model, a, r.dev, r.stream, r.ld);
        else return -2;
    }
    if constexpr (KnotConfigSmall<typename KnotLaunch<ModelT<T>, Q, T, WITH_J, ERR>::Shape, T, WITH_J, Q>::distinct) {
        // RDB200_SMALL_N overrides the threshold (experiments: scripts/tile_threshold.py)
        static const long long small_n = []() {
            const char* e = std::getenv("RDB200_SMALL_N");
            return e ? std::atoll(e) : (prefers_narrow_tiles<ModelT<T>, T, ERR>() ? (long long)1 << 62 : (long long)RDB_SMALL_N);
        }();
        if (r.N <= small_n && !(r.whole_sm && r.N > RDB_SMALL_N))
            return KnotLaunch<ModelT<T>, Q, T, WITH_J, ERR, false, true>::run(model, a, r.dev, r.stream);
    }
    return KnotLaunch<ModelT<T>, Q, T, WITH_J, ERR>::run(model, a, r.dev, r.stream);
}
#ifndef RDB_IMPLICIT_WARP_MIN_N
#define RDB_IMPLICIT_WARP_MIN_N 8     // models with at least this many states use the warp-cooperative ImplicitMidpoint kernel
#endif
template <template <class> class ModelT, class T>
inline int run_implicit(const KnotRequest& r) {
    if (r.op != OP_KNOT || r.err) return -2;
    ModelT<T> model; model.p = cast_params<T>(r.params);
    KnotArgs<T> a;
    a.Z = static_cast<const T*>(r.Z); a.dt = r.dt; a.dt0 = r.dt0; a.t = r.t;
    a.J = static_cast<T*>(r.J); a.out = static_cast<T*>(r.out); a.N = r.N; a.use_jmap = 0;
    if (r.N <= 0) return 0;
    if constexpr (mp_block_ok<ModelT<T>, T>()) {
        // df/dx is block lower-triangular under the model's declared ordering: block forward substitution, one thread per knot
        // (implicit_block.cuh).  RDB200_IMPLICIT_BLOCK=0 selects the dense kernels below for comparison.
        static const bool dense = []() { const char* e = std::getenv("RDB200_IMPLICIT_BLOCK"); return e && e[0] == '0'; }();
        if (!dense) return r.with_j ? MidpointBlockLaunch<ModelT<T>, T, true>::run(model, a, r.dev.sm_count, r.stream)
                                    : MidpointBlockLaunch<ModelT<T>, T, false>::run(model, a, r.dev.sm_count, r.stream);
    }
    if constexpr (ModelT<T>::n >= RDB_IMPLICIT_WARP_MIN_N && ModelT<T>::n + ModelT<T>::m <= 32) {
        // rigid bodies: a group of lanes per knot, the columns of [A B] dealt round-robin to the lanes, Gauss-Jordan across the group
        // (kernels.cuh: implicit_midpoint_group_kernel).  RDB200_IMPLICIT_WARP=1 selects the round-1 kernel (one warp-wide group per knot,
        // one column per lane) for comparison.
        static const bool use_warp = []() { const char* e = std::getenv("RDB200_IMPLICIT_WARP"); return e && e[0] == '1'; }();
        if (!use_warp) {
            // lanes per knot: 4 (8 knots per warp) measured best for both dtypes (quadrotor fp32 612 vs 710 us with 8 lanes; satellite{MRP}
            // fp64 1230 vs 1656 us; quadrotor fp64 equal); RDB200_IMPLICIT_L=4|8 overrides it for experiments
            static const int forced = []() { const char* e = std::getenv("RDB200_IMPLICIT_L"); return e ? std::atoi(e) : 0; }();
            const int L = forced == 4 || forced == 8 ? forced : 4;
            const unsigned grid = unsigned((r.N * L + IMG_THREADS - 1) / IMG_THREADS);
            if (L == 4) {
                if (r.with_j) implicit_midpoint_group_kernel<ModelT<T>, T, true, 4><<<grid, IMG_THREADS, 0, r.stream>>>(model, a);
                else implicit_midpoint_group_kernel<ModelT<T>, T, false, 4><<<grid, IMG_THREADS, 0, r.stream>>>(model, a);
            } else {
                if (r.with_j) implicit_midpoint_group_kernel<ModelT<T>, T, true, 8><<<grid, IMG_THREADS, 0, r.stream>>>(model, a);
                else implicit_midpoint_group_kernel<ModelT<T>, T, false, 8><<<grid, IMG_THREADS, 0, r.stream>>>(model, a);
            }
            return int(cudaGetLastError());
        }
        constexpr int GS = (ModelT<T>::n + ModelT<T>::m <= 16) ? 16 : 32;
        const unsigned grid = unsigned((r.N * GS + 127) / 128);
        if (r.with_j) implicit_midpoint_warp_kernel<ModelT<T>, T, true, GS><<<grid, 128, 0, r.stream>>>(model, a);
        else implicit_midpoint_warp_kernel<ModelT<T>, T, false, GS><<<grid, 128, 0, r.stream>>>(model, a);
    } else {
        const unsigned grid = unsigned((r.N + 127) / 128);
        if (r.with_j) implicit_midpoint_kernel<ModelT<T>, T, true><<<grid, 128, 0, r.stream>>>(model, a);
        else implicit_midpoint_kernel<ModelT<T>, T, false><<<grid, 128, 0, r.stream>>>(model, a);
    }
    return int(cudaGetLastError());
}
template <class T>
inline RolloutArgs<T> rollout_args(const KnotRequest& r, int n, int m) {
    RolloutArgs<T> a;
    a.x0 = static_cast<const T*>(r.x0); a.U = static_cast<const T*>(r.U); a.dt = r.dt; a.t = r.t; a.dt0 = r.dt0;
    a.X = static_cast<T*>(r.X); a.ntraj = r.ntraj; a.K = r.K;
    if (r.zmode) { a.kb = r.kb; a.ke = r.ke; a.sj = 1; a.sk = r.ntraj; a.uj = 0; a.uk = 0; a.ldx = n + m; a.U = nullptr; }
    else { a.kb = 0; a.ke = r.K - 1; a.sj = r.K; a.sk = 1; a.uj = r.K - 1; a.uk = 1; a.ldx = n; }
    return a;
}
template <template <class> class ModelT, class T, int Q>
inline int run_rollout(const KnotRequest& r) {
    if constexpr (Q == Q_CONTINUOUS) return -2;
    else {
        ModelT<T> model; model.p = cast_params<T>(r.params);
        if (r.ntraj <= 0 || r.K <= 0) return 0;
        // single-warp CTAs by default: 4096 trajectories = 128 warps, one per SM (the sweep is latency-bound: every warp gets a scheduler
        // to itself).  The pipelined rollout + linearisation packs 8 warps per CTA instead, so that the rollout occupies FEW SMs and the
        // Jacobian kernel — whose CTAs fill an SM's register file — keeps the rest (abi.cu: rdb_trajectory_rollout_linearize).
        const int block = r.rollout_block > 0 ? r.rollout_block : 32;
        const unsigned grid = unsigned((r.ntraj + block - 1) / block);
        rollout_kernel<ModelT<T>, Q, T><<<grid, block, 0, r.stream>>>(model, rollout_args<T>(r, ModelT<T>::n, ModelT<T>::m));
        return int(cudaGetLastError());
    }
}
template <template <class> class ModelT, class T, int Q>
inline int run_q(const KnotRequest& r) {
    if (r.op == OP_ROLLOUT) return run_rollout<ModelT, T, Q>(r);
    if (r.err) {
        if constexpr (ModelT<T>::rot != ROT_NONE && Q != Q_CONTINUOUS) return run_one<ModelT, T, Q, true, true>(r);
        else return -2;
    }
    return r.with_j ? run_one<ModelT, T, Q, true>(r) : run_one<ModelT, T, Q, false>(r);
}
template <class T>
inline DynErrArgs<T> dynerr_args(const KnotRequest& r) {
    DynErrArgs<T> a;
    a.Z1 = static_cast<const T*>(r.Z); a.Z2 = static_cast<const T*>(r.Z2); a.ld2 = r.ld2; a.t = r.t; a.dt = r.dt; a.dt0 = r.dt0;
    a.J2 = static_cast<T*>(r.J2); a.J1 = static_cast<T*>(r.J); a.e = static_cast<T*>(r.out); a.N = r.N;
    return a;
}
template <template <class> class ModelT, class T>
inline int run_dynerr(const KnotRequest& r) {
    if (r.Q != Q_IMPLICIT_MIDPOINT) return -2;          // explicit rules are composed from the knot kernel by the caller (abi.cu)
    if (r.N <= 0) return 0;
    ModelT<T> model; model.p = cast_params<T>(r.params);
    const unsigned grid = unsigned((r.N + 127) / 128);
    if (r.with_j) midpoint_error_kernel<ModelT<T>, T, true><<<grid, 128, 0, r.stream>>>(model, dynerr_args<T>(r));
    else midpoint_error_kernel<ModelT<T>, T, false><<<grid, 128, 0, r.stream>>>(model, dynerr_args<T>(r));
    return int(cudaGetLastError());
}
template <template <class> class ModelT, class T>
inline int run_t(const KnotRequest& r) {
    if (r.op == OP_DYNERR) return run_dynerr<ModelT, T>(r);
    switch (r.Q) {
        case Q_EULER: return run_q<ModelT, T, Q_EULER>(r);
        case Q_RK2: return run_q<ModelT, T, Q_RK2>(r);
        case Q_RK3: return run_q<ModelT, T, Q_RK3>(r);
        case Q_RK4: return run_q<ModelT, T, Q_RK4>(r);
        case Q_CONTINUOUS: return run_q<ModelT, T, Q_CONTINUOUS>(r);
        case Q_IMPLICIT_MIDPOINT: return run_implicit<ModelT, T>(r);
    }
    return -2;
}
template <template <class> class ModelT>
inline int run_model(const KnotRequest& r) { return r.dtype == 0 ? run_t<ModelT, float>(r) : run_t<ModelT, double>(r); }

}  // namespace rdb
