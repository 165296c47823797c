// kernels.cuh — the batched knot-point kernel (K1/K2/K3 of SURVEY.md §2) for sm_100a.
//
// One template serves dynamics(), discrete_dynamics() and their Jacobians:
//   knot_kernel<Model, Q, T, TILE, WITH_J, Chunks, MINB, ROLL, ERR>
//     Q in {EULER,RK2,RK3,RK4,CONTINUOUS};  WITH_J: also produce d out / d [x;u]  (n x (n+m), column-major [A B],
//     reference layout: src/jacobian.jl:26-37).
//
// Mapping (B200-first):
//   * persistent CTAs, one tile = TILE consecutive knot points; CTA = TILE x NROLES threads.
//   * a ROLE is a warp-uniform compile-time column chunk of [x;u]: the threads of role r carry forward-mode
//     partials only for the columns in CHUNKS[r] (sdual.cuh), so register pressure is split across warps
//     without any intra-warp divergence and without communication (columns of a Jacobian are independent).
//   * HBM <-> SM traffic is two streams of contiguous bytes per tile when the caller uses the reference's own
//     knot-major layout (Julia Array{T,2}(n+m,N) in, Array{T,3}(n,n+m,N) out): the input tile arrives by one
//     TMA bulk copy (cp.async.bulk + mbarrier complete_tx, double-buffered, prefetched one tile ahead), results
//     are assembled as the exact output image in shared memory and leave by one TMA bulk store per tile that
//     drains while the next tile computes.  No per-element global address arithmetic exists on that path.
//   * unaligned pointers and the ragged last tile use a cooperative coalesced copy between the same shared-memory images
//     and global memory.
//   * component-major ("SoA") callers are served by the same kernel with SOA = true: 2-D tensor maps let the TMA unit load the
//     tile as the image [component][knot] and store the results from [entry][knot] images — the layout change costs nothing
//     (coalesced transposes around the knot-major kernel, layout.cu, remain as the fallback for rows that are not whole
//     16-byte units).
#pragma once
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#endif
#include "integrators.cuh"

namespace rdb {

enum Layout { LAYOUT_AOS = 0, LAYOUT_SOA = 1 };

// CUtensorMap (cuda.h) by another name, so that this header also compiles under NVRTC: 128 opaque bytes the host encodes with
// cuTensorMapEncodeTiled (launch.cuh: encode_jmap) and the TMA unit reads from the kernel's parameter space.
struct alignas(64) TensorMap { unsigned long long opaque[16]; };

template <class T>
struct KnotArgs {
    const T* Z;          // [x;u] per knot, knot-major (N, n+m)
    const double* dt;    // per-knot step (N) or nullptr -> dt0      (KnotPoint.dt is Float64: src/knotpoint.jl:148-153)
    double dt0;
    const double* t;     // per-knot time (N) or nullptr -> 0; read only by models that declare time_varying (KnotPoint.t)
    T* J;                // (N, n+m, n): per knot an n x (n+m) column-major matrix;  may be nullptr
                         // (error-state mode: nerr x (nerr+m) per knot instead)
    T* out;              // xdot or x+, (N, n);  may be nullptr
    long long N;
    int use_jmap;        // jmap describes J as a 2-D tensor (E x N); used by the padded-image store of the n = 12 models
    int stream_out = 0;  // outputs of this launch exceed what the L2 can keep for a consumer (knot_stream_out): their stores carry the
                         // L2 evict_first hint — written lines leave the 126 MB write-back L2 promptly and in order instead of by LRU
                         // among the reads (C2 skeleton, scripts/micro/stream_mix.cu: 39.7 -> 37.7 us); small batches keep J in L2
    TensorMap jmap;      // (component-major kernels, SOA = true: J as the tensor N x E, inner extent = knots)
    TensorMap zmap;      // component-major kernels only: Z as the tensor N x (n+m), out as N x n
    TensorMap omap;
};

// ---- PTX helpers: mbarrier + 1-D bulk async copies (TMA engine; SASS: UBLKCP / SYNCS) ------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) { while (!mbar_try_wait(bar, parity)) {} }
__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, uint32_t src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}
// 2-D tiled TMA store (SASS UTMASTG): the box is the whole PADDED smem image (row pitch PJ elements); the tensor's inner extent is
// E < PJ, so the pad elements fall outside the tensor and the TMA unit drops them — dense rows in HBM from a conflict-free image,
// one instruction per tile.
__device__ __forceinline__ void tensor_store_2d(const TensorMap* tm, uint32_t src_smem, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(tm), "r"(src_smem), "r"(c0), "r"(c1) : "memory");
}
// 2-D tiled TMA load (SASS UTMALDG): box lands densely in shared memory, elements outside the tensor arrive as zeros, the mbarrier
// receives the byte count of the whole box
__device__ __forceinline__ void tensor_load_2d(uint32_t dst_smem, const TensorMap* tm, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst_smem), "l"(tm), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
// the same stores with an L2 eviction-priority hint (64-bit policy from createpolicy), and L2 prefetches of a tile's input rows
__device__ __forceinline__ uint64_t l2_policy_evict_first() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ void bulk_store(void* dst, uint32_t src_smem, uint32_t bytes, uint64_t pol) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst), "r"(src_smem), "r"(bytes), "l"(pol) : "memory");
}
__device__ __forceinline__ void tensor_store_2d(const TensorMap* tm, uint32_t src_smem, int c0, int c1, uint64_t pol) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
                 ::"l"(tm), "r"(src_smem), "r"(c0), "r"(c1), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tensor_prefetch_2d_l2(const TensorMap* tm, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(tm), "r"(c0), "r"(c1) : "memory");
}
// Outputs larger than this do not survive in the L2 for a consumer anyway (126 MB, shared with the input stream): stream them out.
__host__ __device__ constexpr bool knot_stream_out(long long N, long long out_bytes_per_knot) { return N * out_bytes_per_knot > (64ll << 20); }
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- compile-time helpers ---------------------------------------------------------------------------------
template <mask_t... Ms> struct MaskList { static constexpr int count = int(sizeof...(Ms)); };
template <class T, mask_t CHUNK, size_t... Is>
__device__ __forceinline__ auto load_seeded(const T* zrow, rstd::index_sequence<Is...>) {
    return vec(seed<T, int(Is), CHUNK>(zrow[Is])...);
}
template <class T> __device__ __forceinline__ T plain(const T& a) { return a; }
template <class T, mask_t M> __device__ __forceinline__ T plain(const SD<T, M>& a) { return a.v; }

// 16-byte shared-memory stores of 4 floats / 2 doubles (used when a Jacobian column is a whole number of 16-byte units)
__device__ __forceinline__ void st16(float* p, float a, float b, float c, float d) { *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d); }
__device__ __forceinline__ void st16(double* p, double a, double b) { *reinterpret_cast<double2*>(p) = make_double2(a, b); }
template <int N_, int J, class XN, size_t... Ks>
__device__ __forceinline__ void put_col_v(const XN& xn, float* col, rstd::index_sequence<Ks...>) {
    (st16(col + 4 * int(Ks), partial<J>(get<4 * int(Ks)>(xn)), partial<J>(get<4 * int(Ks) + 1>(xn)), partial<J>(get<4 * int(Ks) + 2>(xn)),
          partial<J>(get<4 * int(Ks) + 3>(xn))), ...);
}
template <int N_, int J, class XN, size_t... Ks>
__device__ __forceinline__ void put_col_v(const XN& xn, double* col, rstd::index_sequence<Ks...>) {
    (st16(col + 2 * int(Ks), partial<J>(get<2 * int(Ks)>(xn)), partial<J>(get<2 * int(Ks) + 1>(xn))), ...);
}
// ES = element stride of the image: 1 for the knot-major images, TILE for the component-major ones (element e of knot kt at
// e * TILE + kt: consecutive lanes on consecutive words)
template <int N_, mask_t CHUNK, int J, bool VEC, int ES, class T, class XN, size_t... Is>
__device__ __forceinline__ void put_col(const XN& xn, T* jrow, rstd::index_sequence<Is...>) {
    if constexpr (chas(CHUNK, J)) {
        if constexpr (VEC) put_col_v<N_, J>(xn, jrow + N_ * J, rstd::make_index_sequence<size_t(N_ * sizeof(T) / 16)>{});
        else ((jrow[(int(Is) + N_ * J) * ES] = partial<J>(get<int(Is)>(xn))), ...);
    }
}
template <int N_, mask_t CHUNK, bool VEC, int ES = 1, class T, class XN, size_t... Js>
__device__ __forceinline__ void put_cols(const XN& xn, T* jrow, rstd::index_sequence<Js...>) {
    static_assert(!VEC || ES == 1, "16-byte column stores need unit element stride");
    (put_col<N_, CHUNK, int(Js), VEC, ES>(xn, jrow, rstd::make_index_sequence<size_t(N_)>{}), ...);
}
template <int ES = 1, class T, class XN, size_t... Is>
__device__ __forceinline__ void put_vals(const XN& xn, T* orow, rstd::index_sequence<Is...>) {
    ((orow[int(Is) * ES] = plain(get<int(Is)>(xn))), ...);
}

// ---- error-state ("LieState") mode --------------------------------------------------------------------------------------
// Altro / TrajectoryOptimization consume  Abar = G(x+)' A G(x),  Bbar = G(x+)' B  for RotationState models, with
// G = errstate_jacobian (reference: src/liestate.jl:262-298; jacobian_width = errstate_dim + m, src/functionbase.jl:135).
// Here the product never exists as a product: the attitude entries of x are SEEDED with the rows of the attitude block of
// G(x) (3 tangent directions instead of 4 unit vectors), forward mode then yields [A G(x), B] directly, and the attitude rows
// of the result are contracted with G(x+)'.  Output: nerr x (nerr + m), column-major.
template <class T, int ROT>
__device__ __forceinline__ void att_grad(const T* p, T (&G)[4][3]) {   // Rotations.∇differential, values only
    if constexpr (ROT == ROT_QUAT) {                                    // L(q) H with q normalised (default constructor)
        const T s = rsqrt_(p[0] * p[0] + p[1] * p[1] + p[2] * p[2] + p[3] * p[3]);
        const T w = p[0] * s, x = p[1] * s, y = p[2] * s, z = p[3] * s;
        G[0][0] = -x; G[0][1] = -y; G[0][2] = -z;
        G[1][0] = w;  G[1][1] = -z; G[1][2] = y;
        G[2][0] = z;  G[2][1] = w;  G[2][2] = -x;
        G[3][0] = -y; G[3][1] = x;  G[3][2] = w;
    } else {
        const T sk[3][3] = {{T(0), -p[2], p[1]}, {p[2], T(0), -p[0]}, {-p[1], p[0], T(0)}};
        const T a = (ROT == ROT_MRP) ? T(1) - (p[0] * p[0] + p[1] * p[1] + p[2] * p[2]) : T(1);
        const T c = (ROT == ROT_MRP) ? T(2) : T(1);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) G[i][j] = (i == j ? a : T(0)) + c * (sk[i][j] + p[i] * p[j]);
    }
}
// value + partials g3,g4,g5 on the error columns 3,4,5 that belong to this role's chunk
template <class T, mask_t CHUNK>
__device__ __forceinline__ auto seed_att(T v, T g3, T g4, T g5) {
    constexpr mask_t M = CHUNK & mask_t(0x38);
    if constexpr (M == 0) return v;
    else {
        SD<T, M> r; r.v = v; r.zero_parts();
        if constexpr (chas(M, 3)) r.template set_part<3>(g3);
        if constexpr (chas(M, 4)) r.template set_part<4>(g4);
        if constexpr (chas(M, 5)) r.template set_part<5>(g5);
        return r;
    }
}
template <class Model, class T, mask_t CHUNK, int I>
__device__ __forceinline__ auto seed_err(const T* z, const T (&G)[4][3]) {
    constexpr int np = Model::n - 9;
    if constexpr (I < 3) return seed<T, I, CHUNK>(z[I]);
    else if constexpr (I < 3 + np) return seed_att<T, CHUNK>(z[I], G[I - 3][0], G[I - 3][1], G[I - 3][2]);
    else return seed<T, I - (np - 3), CHUNK>(z[I]);
}
template <class Model, class T, mask_t CHUNK, size_t... Is>
__device__ __forceinline__ auto load_seeded_err(const T* z, rstd::index_sequence<Is...>) {
    T G[4][3];
    att_grad<T, Model::rot>(z + 3, G);
    return vec(seed_err<Model, T, CHUNK, int(Is)>(z, G)...);
}
template <class T, class A, size_t... Is> __device__ __forceinline__ void plain_vals(const A& a, T* out, rstd::index_sequence<Is...>) { ((out[Is] = plain(get<int(Is)>(a))), ...); }
template <int J, class T, class A, size_t... Is>
__device__ __forceinline__ auto contract_col(const T (&G)[4][3], const A& att, rstd::index_sequence<Is...>) { return ((G[Is][J] * get<int(Is)>(att)) + ...); }
// rows of the result in error coordinates: [r+; G(x+)' att+; v+; w+]
template <class Model, class T, class XN>
__device__ __forceinline__ auto project_err(const XN& xn) {
    constexpr int np = Model::n - 9;
    auto att = slice<3, np>(xn);
    T p[4], G[4][3];
    plain_vals(att, p, rstd::make_index_sequence<size_t(np)>{});
    att_grad<T, Model::rot>(p, G);
    using Seq = rstd::make_index_sequence<size_t(np)>;
    return cat(slice<0, 3>(xn), vec(contract_col<0>(G, att, Seq{}), contract_col<1>(G, att, Seq{}), contract_col<2>(G, att, Seq{})),
               slice<3 + np, 6>(xn));
}

// all threads of the CTA meet here between "results are in registers" and "results go to the smem images":
// thread 0 first waits until the previous tile's bulk stores have finished reading those images.
template <int NTHR, int ISSUERS>
__device__ __forceinline__ void images_free_barrier(int tid) {
    if (tid < ISSUERS) bulk_wait_read0();      // bulk groups are per thread: every thread that issued stores waits for its own
    __syncwarp();
    // The roles (warp-uniform code paths) reach this barrier from role-specific program counters, so it is the NON-aligned form:
    // `barrier.sync` lets the threads of a CTA arrive from different instructions (bar.sync == barrier.sync.aligned promises that every
    // thread executes the same one, which compute-sanitizer's synccheck rightly reported for the round-1 form).
    asm volatile("barrier.sync 1, %0;" ::"n"(NTHR) : "memory");
}

// this thread's [x;u] row from the smem image into registers.  Rows that are whole 16-byte units (n + m = 16, 18, ...) are read with
// 16-byte loads: their pitch is a multiple of 4 words, so 4-byte loads put 8..16 lanes of a warp on one bank.
__device__ __forceinline__ void ld16(const float* p, float* z) { const float4 v = *reinterpret_cast<const float4*>(p); z[0] = v.x; z[1] = v.y; z[2] = v.z; z[3] = v.w; }
__device__ __forceinline__ void ld16(const double* p, double* z) { const double2 v = *reinterpret_cast<const double2*>(p); z[0] = v.x; z[1] = v.y; }
template <class T, int NZ, int ES = 1>
__device__ __forceinline__ void read_row(const T* zrow, T (&z)[NZ]) {
    constexpr int PER = int(16 / sizeof(T));
    if constexpr (ES != 1) {
#pragma unroll
        for (int i = 0; i < NZ; ++i) z[i] = zrow[i * ES];
    } else if constexpr (NZ % PER == 0) {
#pragma unroll
        for (int i = 0; i < NZ; i += PER) ld16(zrow + i, z + i);
    } else {
#pragma unroll
        for (int i = 0; i < NZ; ++i) z[i] = zrow[i];
    }
}

// One role: evaluate the map for one knot with partials for the columns in CHUNK; write this role's share.
template <class Model, int Q, class T, bool WITH_J, bool ERR, mask_t CHUNK, bool WRITE_OUT, int NTHR, int ROLL, int ISSUERS, bool VEC, int ES>
__device__ __forceinline__ void role_body(const Model& model, const T* zrow, T h, T t, T* jrow, T* orow, int tid) {
    constexpr int n = Model::n, m = Model::m, NZ = n + m;
    model.reset();                 // per-knot evaluation caches (e.g. Cartpole's stage-1 sincos) start empty
    T zreg[NZ];
    read_row<T, NZ, ES>(zrow, zreg);
    if constexpr (ERR) {
        auto zz = load_seeded_err<Model, T, CHUNK>(zreg, rstd::make_index_sequence<size_t(NZ)>{});
        auto xn = integrate<Q, T, ROLL>(model, slice<0, n>(zz), slice<n, m>(zz), h, t);
        auto e = project_err<Model, T>(xn);
        images_free_barrier<NTHR, ISSUERS>(tid);
        put_cols<Model::nerr, CHUNK, VEC, ES>(e, jrow, rstd::make_index_sequence<size_t(Model::nerr + m)>{});
        if constexpr (WRITE_OUT) { if (orow) put_vals<ES>(xn, orow, rstd::make_index_sequence<size_t(n)>{}); }
    } else {
        auto zz = load_seeded<T, (WITH_J ? CHUNK : mask_t(0))>(zreg, rstd::make_index_sequence<size_t(NZ)>{});
        auto xn = integrate<Q, T, (WITH_J ? ROLL : 0)>(model, slice<0, n>(zz), slice<n, m>(zz), h, t);
        images_free_barrier<NTHR, ISSUERS>(tid);
        if constexpr (WITH_J) put_cols<n, CHUNK, VEC, ES>(xn, jrow, rstd::make_index_sequence<size_t(NZ)>{});
        if constexpr (WRITE_OUT) { if (orow) put_vals<ES>(xn, orow, rstd::make_index_sequence<size_t(n)>{}); }
    }
}

template <int R, class L> struct list_at;
template <int R, mask_t M0, mask_t... Ms> struct list_at<R, MaskList<M0, Ms...>> { static constexpr mask_t value = list_at<R - 1, MaskList<Ms...>>::value; };
template <mask_t M0, mask_t... Ms> struct list_at<0, MaskList<M0, Ms...>> { static constexpr mask_t value = M0; };

template <class Model, int Q, class T, bool WITH_J, bool ERR, class Chunks, int NTHR, int ROLL, int ISSUERS, bool VEC, int ES, int R = 0>
__device__ __forceinline__ void dispatch_role(int role, const Model& model, const T* zrow, T h, T t, T* jrow, T* orow, int tid) {
    if constexpr (R + 1 == Chunks::count) {
        role_body<Model, Q, T, WITH_J, ERR, list_at<R, Chunks>::value, R == 0, NTHR, ROLL, ISSUERS, VEC, ES>(model, zrow, h, t, jrow, orow, tid);
    } else {
        if (role == R) role_body<Model, Q, T, WITH_J, ERR, list_at<R, Chunks>::value, R == 0, NTHR, ROLL, ISSUERS, VEC, ES>(model, zrow, h, t, jrow, orow, tid);
        else dispatch_role<Model, Q, T, WITH_J, ERR, Chunks, NTHR, ROLL, ISSUERS, VEC, ES, R + 1>(role, model, zrow, h, t, jrow, orow, tid);
    }
}

// cooperative copies between a knot-major smem image [cnt][W] (row pitch P >= W) and knot-major global memory: the fallback for
// unaligned pointers and the ragged last tile
template <class T>
__device__ __forceinline__ void coop_load(T* img, int P, const T* g, long long k0, int cnt, int W, int tid, int nthr) {
    if (P == W) {
        const T* src = g + k0 * W;
        for (int i = tid; i < cnt * W; i += nthr) img[i] = src[i];
    } else {
        const int lane = tid & 31, nw = nthr >> 5;
        for (int r = tid >> 5; r < cnt; r += nw) { const T* src = g + (k0 + r) * W; for (int e = lane; e < W; e += 32) img[r * P + e] = src[e]; }
    }
}
template <class T>
__device__ __forceinline__ void coop_store(const T* img, int P, T* g, long long k0, int cnt, int W, int tid, int nthr) {
    if (P == W) {
        T* dst = g + k0 * W;
        for (int i = tid; i < cnt * W; i += nthr) dst[i] = img[i];
    } else {   // one warp per knot row: conflict-free smem reads (unit stride), fully coalesced global writes
        const int lane = tid & 31, nw = nthr >> 5;
        for (int r = tid >> 5; r < cnt; r += nw) { T* dst = g + (k0 + r) * W; for (int e = lane; e < W; e += 32) dst[e] = img[r * P + e]; }
    }
}

__host__ __device__ constexpr int cgcd(int a, int b) { return b == 0 ? a : cgcd(b, a % b); }
// Dynamic shared memory of knot_kernel as a function of plain integers (jr x jc = Jacobian image shape, es = sizeof(T)); used by
// KnotSmem below (static_assert) and by custom.cu, which only knows the dimensions of a user model at run time.
#ifndef RDB_TUNE_ROWSTORE
#define RDB_TUNE_ROWSTORE 1      // 0: tuning experiments only (dense image even where it is bank-conflicted)
#endif
#ifndef RDB_ROWSTORE_MINWAY
#define RDB_ROWSTORE_MINWAY 16   // pad the image when the dense one would put at least this many lanes of a warp on one bank
#endif
__host__ __device__ constexpr bool knot_rowstore(int jr, int jc, bool with_j, int es) {
    // rigid bodies: pad when the dense image would put >= 16 lanes on one bank (the 8-way case E = 216 fp32 measured equal).
    // The small fp64 models (Cartpole: rows of 20 doubles, 8 lanes per bank) were padded too for a while (C2 39.9 -> 38.7 us); since
    // the streamed-out stores carry the L2 evict_first hint the dense image with its ONE 1-D bulk store is the faster one again
    // (37.8 against 38.2 us, three alternating runs each) and saves the per-launch tensor-map encode on the host.
    const int ways = cgcd(jr * jc * es / 4, 32);
    return RDB_TUNE_ROWSTORE && with_j && ((jr * jc) % 2 == 0) && ((jr * es) % 16 == 0) && jr >= 12 && ways >= RDB_ROWSTORE_MINWAY;
}
__host__ __device__ constexpr int knot_pitch(int jr, int jc, bool with_j, int es) {
    const int E = jr * jc, u = (E * es + 15) / 16;
    return knot_rowstore(jr, jc, with_j, es) ? ((u % 2 == 0) ? u + 1 : u) * 16 / es : E;
}
// nscal: per-knot scalar arrays staged next to the [x;u] rows (dt; dt and t for time-varying models), double-buffered like them
__host__ __device__ constexpr size_t knot_smem_total(int n, int m, int jr, int jc, int TILE, bool with_j, int es, int nscal) {
    const size_t a16 = 15;
    const size_t in_b = (size_t(TILE) * size_t(n + m) * es + a16) & ~a16;
    const size_t j_b = with_j ? ((size_t(TILE) * size_t(knot_pitch(jr, jc, with_j, es)) * es + a16) & ~a16) : 0;
    const size_t o_b = (size_t(TILE) * size_t(n) * es + a16) & ~a16;
    return 2 * in_b + j_b + o_b + size_t(2) * size_t(nscal) * size_t(TILE) * 8 + 16;
}
template <class Model, int TILE, bool WITH_J, class T, bool ERR = false>
struct KnotSmem {
    static constexpr int n = Model::n, NZ = Model::n + Model::m;
    static constexpr int JR = ERR ? Model::nerr : n, JC = ERR ? Model::nerr + Model::m : NZ, E = JR * JC;   // Jacobian image shape
    // Row pitch of the Jacobian image.  Thread kt writes element e of its knot to word kt*PJ + e, so a pitch that is a multiple
    // of 32 words puts all lanes of a warp on ONE bank (E = 192 for n=12, m=4).  The quaternion models have odd E (221, 247):
    // dense image, conflict-free, leaves by ONE TMA bulk store per tile.  The n = 12 models (E even, columns = whole 16-byte
    // units) get a pitch of an ODD number of 16-byte units, write their columns with 16-byte stores (conflict-free) and leave
    // by ONE 2-D tensor store per tile whose box is the padded image (the pad lies outside the tensor and is dropped);
    // fallback when no tensor map could be encoded: one 1-D bulk store PER KNOT ROW, issued by TILE threads in parallel.
    // (only when the dense image would be badly conflicted: >= 16 lanes per bank; the 8-way case E = 216 fp32 measured faster dense)
    static constexpr bool ROWSTORE = knot_rowstore(JR, JC, WITH_J, int(sizeof(T)));
    static constexpr int PJ = knot_pitch(JR, JC, WITH_J, int(sizeof(T)));
    static constexpr int ISSUERS = ROWSTORE ? TILE : 1;
    static constexpr size_t in_bytes = size_t(TILE) * NZ * sizeof(T);
    static constexpr size_t j_bytes = WITH_J ? size_t(TILE) * PJ * sizeof(T) : 0;
    static constexpr size_t j_dense_bytes = WITH_J ? size_t(TILE) * E * sizeof(T) : 0;
    static constexpr size_t o_bytes = size_t(TILE) * n * sizeof(T);
    __host__ __device__ static constexpr size_t align16(size_t b) { return (b + 15) & ~size_t(15); }
    static constexpr size_t off_in0 = 0;
    static constexpr size_t off_in1 = align16(in_bytes);
    static constexpr size_t off_j = off_in1 + align16(in_bytes);
    static constexpr size_t off_o = off_j + align16(j_bytes);
    // per-knot steps (and times, for models that read them): [buffer][array][knot] doubles, staged by the same TMA transaction as the rows
    static constexpr int NSCAL = uses_time<Model>::value ? 2 : 1;
    static constexpr size_t sc_bytes = size_t(TILE) * 8;
    static constexpr size_t off_sc = off_o + align16(o_bytes);
    static constexpr size_t off_bar = off_sc + 2 * NSCAL * sc_bytes;
    static constexpr size_t total = off_bar + 16;
    static_assert(total == knot_smem_total(n, Model::m, JR, JC, TILE, WITH_J, int(sizeof(T)), NSCAL), "smem layout and knot_smem_total disagree");
};


// SOA = true: component-major caller arrays (Z (n+m, N), J (n(n+m), N), out (n, N); N * sizeof(T) a multiple of 16).  The layout
// change is done by the TMA unit: the tile of Z arrives through a 2-D tensor map as the box TILE x (n+m) — image [component][knot],
// read conflict-free — and the results, assembled as [entry][knot], leave through 2-D tensor stores of TILE x E / TILE x n boxes.
// The ragged last tile needs no special case: loads beyond N arrive as zeros, stores beyond N are dropped.
template <class Model, int Q, class T, int TILE, bool WITH_J, class Chunks, int MINB, int ROLL, bool ERR = false, bool SOA = false>
__global__ void __launch_bounds__(TILE * Chunks::count, MINB)
knot_kernel(const Model model, const __grid_constant__ KnotArgs<T> a) {
    using S = KnotSmem<Model, TILE, WITH_J, T, ERR>;
    constexpr int n = Model::n, NZ = Model::n + Model::m, E = S::E;
    constexpr int NTHR = TILE * Chunks::count;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* in_img[2] = {reinterpret_cast<T*>(smem_raw + S::off_in0), reinterpret_cast<T*>(smem_raw + S::off_in1)};
    T* j_img = reinterpret_cast<T*>(smem_raw + S::off_j);
    T* o_img = reinterpret_cast<T*>(smem_raw + S::off_o);
    const uint32_t bar0 = smem_u32(smem_raw + S::off_bar);

    const int tid = threadIdx.x;
    const int role = tid / TILE;            // warp-uniform (TILE % 32 == 0)
    const int kt = tid - role * TILE;
    const long long N = a.N;
    const long long ntiles = (N + TILE - 1) / TILE;
    const bool want_j = WITH_J && a.J != nullptr;
    const bool want_o = a.out != nullptr;
    // TMA path needs the reference's knot-major layout and 16-byte aligned streams
    const bool tma_ok = SOA || ((reinterpret_cast<uintptr_t>(a.Z) | reinterpret_cast<uintptr_t>(a.J) |
                                                           reinterpret_cast<uintptr_t>(a.out)) & 15) == 0;
    auto tile_tma = [&](long long tile) { return tma_ok && (SOA || (tile + 1) * TILE <= N); };
    // Per-knot steps / times (KnotPoint.dt / .t as arrays: trajectories, the mixed sweep) ride on the tile's TMA transaction when their
    // tile is a whole 16-byte-aligned range: one more bulk copy of TILE doubles per array into sc_img[buffer][array][knot].  As plain
    // per-thread global loads they cost 14 % of the kernel time for 0.8 - 4 % more bytes (quadrotor fp32 2^20: 172.9 -> 198.0 us,
    // Cartpole fp64 37.4 -> 42.6 us), even when requested a tile ahead; they remain the fallback for ragged tiles and odd alignments.
    double* sc_img = reinterpret_cast<double*>(smem_raw + S::off_sc);
    const bool dt_tma = (Q != Q_CONTINUOUS) && a.dt != nullptr && (reinterpret_cast<uintptr_t>(a.dt) & 15) == 0;
    const bool t_tma = uses_time<Model>::value && a.t != nullptr && (reinterpret_cast<uintptr_t>(a.t) & 15) == 0;
    auto sc_tma = [&](long long tile) { return (tile + 1) * TILE <= N && tile_tma(tile); };   // scalars staged for this tile?
    auto load_tile = [&](long long tile, int buf) {           // thread 0: one TMA copy of the tile's [x;u] rows into image `buf`
        const bool sc = sc_tma(tile);
        const uint32_t extra = sc ? uint32_t(S::sc_bytes) * ((dt_tma ? 1u : 0u) + (t_tma ? 1u : 0u)) : 0u;
        mbar_expect_tx(bar0 + 8 * buf, uint32_t(S::in_bytes) + extra);
        if constexpr (SOA) tensor_load_2d(smem_u32(in_img[buf]), &a.zmap, int(tile * TILE), 0, bar0 + 8 * buf);
        else bulk_load(smem_u32(in_img[buf]), a.Z + tile * TILE * NZ, uint32_t(S::in_bytes), bar0 + 8 * buf);
        if (sc && dt_tma) bulk_load(smem_u32(sc_img + (buf * S::NSCAL) * TILE), a.dt + tile * TILE, uint32_t(S::sc_bytes), bar0 + 8 * buf);
        if constexpr (uses_time<Model>::value) {
            if (sc && t_tma) bulk_load(smem_u32(sc_img + (buf * S::NSCAL + 1) * TILE), a.t + tile * TILE, uint32_t(S::sc_bytes), bar0 + 8 * buf);
        }
    };

    long long tile = blockIdx.x;
    if (tid == 0) {
        mbar_init(bar0, 1); mbar_init(bar0 + 8, 1); fence_mbar_init();
        // L2 prefetch of this CTA's first tile, issued BEFORE the dependency wait below: a prefetch only moves lines into the L2 — the
        // point of coherence, which a still-running producer kernel writes through — so it cannot expose stale data, and the DRAM
        // latency of the first tile overlaps the tail of the previous kernel in the stream.
#ifndef RDB_TUNE_NO_PREFETCH
        if (tile < ntiles && tile_tma(tile)) {
            if constexpr (SOA) tensor_prefetch_2d_l2(&a.zmap, int(tile * TILE), 0);
            else bulk_prefetch_l2(a.Z + tile * TILE * NZ, uint32_t(S::in_bytes));
        }
#endif
    }
    __syncthreads();
    // Programmatic dependent launch: everything above overlaps the tail of the previous kernel in the stream; nothing
    // below (first global access that returns data) may start before that kernel's memory is visible.  No-ops without the launch attribute.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");

    if (tile < ntiles && tile_tma(tile) && tid == 0) load_tile(tile, 0);
    // Per-knot steps and times (KnotPoint.dt / .t as arrays: trajectories, the mixed sweep) are plain global loads, one per thread.
    // Issued where they are used they cost every warp of the CTA a DRAM round trip at the top of EVERY tile (all warps of a
    // register-filling rigid-body CTA stall together: ~10 % of the tile time); they are therefore requested ONE TILE AHEAD — the volatile
    // load is issued here, its register is first read a whole tile of arithmetic later.
    auto knot_scalar = [&](const double* p, long long k) -> double { return (p && k < N) ? *reinterpret_cast<const volatile double*>(p + k) : 0.0; };
    double h_next = 0.0, t_next = 0.0;
    if (tile < ntiles) {
        if constexpr (Q != Q_CONTINUOUS) { if (a.dt && !(dt_tma && sc_tma(tile))) h_next = knot_scalar(a.dt, tile * TILE + kt); }
        if constexpr (uses_time<Model>::value) { if (!(t_tma && sc_tma(tile))) t_next = knot_scalar(a.t, tile * TILE + kt); }
    }
    uint32_t phase[2] = {0, 0};
#ifdef RDB_TUNE_NO_STREAMOUT
    const bool stream_out = false;
#else
    const bool stream_out = a.stream_out != 0;
#endif
    const uint64_t pol_out = l2_policy_evict_first();
    for (int it = 0; tile < ntiles; ++it, tile += gridDim.x) {
        const int s = it & 1;
        const long long k0 = tile * TILE;
        const int cnt = int((N - k0) < TILE ? (N - k0) : TILE);
        const double h_cur = h_next, t_cur = t_next;
        if (tile + gridDim.x < ntiles) {                       // (only tiles whose scalars do not come through the TMA transaction)
            const long long tn = tile + gridDim.x;
            if constexpr (Q != Q_CONTINUOUS) { if (a.dt && !(dt_tma && sc_tma(tn))) h_next = knot_scalar(a.dt, tn * TILE + kt); }
            if constexpr (uses_time<Model>::value) { if (!(t_tma && sc_tma(tn))) t_next = knot_scalar(a.t, tn * TILE + kt); }
        }
        const long long nxt = tile + gridDim.x;
        // (1) prefetch the next tile's [x;u] rows (buffer s^1 was last read before the previous iteration's barriers)
        if (tid == 0 && nxt < ntiles && tile_tma(nxt)) load_tile(nxt, s ^ 1);
        // (2) this tile's inputs
        const bool tma = tile_tma(tile);
        if (tma) { mbar_wait(bar0 + 8 * s, phase[s]); phase[s] ^= 1; }
        else { coop_load(in_img[s], NZ, a.Z, k0, cnt, NZ, tid, NTHR); __syncthreads(); }
        // (3) compute in registers
        const T* zrow = SOA ? in_img[s] + kt : in_img[s] + kt * NZ;
        constexpr int ES = SOA ? TILE : 1;                     // element stride of the images
        const bool sc = tma && sc_tma(tile);
        T h = T(0);
        if constexpr (Q != Q_CONTINUOUS) h = T(a.dt ? ((sc && dt_tma) ? sc_img[(s * S::NSCAL) * TILE + kt] : h_cur) : a.dt0);
        T tk = T(0);                                           // KnotPoint.t: only time-varying (user) models read it
        if constexpr (uses_time<Model>::value) tk = T((sc && t_tma) ? sc_img[(s * S::NSCAL + 1) * TILE + kt] : t_cur);
        (void)cnt; (void)t_cur; (void)h_cur;
        // (4) evaluate; inside, all threads meet at images_free_barrier() before touching the output images
        //     (rows past the ragged end compute on stale smem and are never copied out)
        dispatch_role<Model, Q, T, WITH_J, ERR, Chunks, NTHR, ROLL, S::ISSUERS, S::ROWSTORE && !SOA, ES>(
            role, model, zrow, h, tk, SOA ? j_img + kt : j_img + kt * S::PJ, want_o ? (SOA ? o_img + kt : o_img + kt * n) : nullptr, tid);
        // (5) publish
        if (tma) {
            fence_proxy_async();
            __syncthreads();
            if constexpr (SOA) {
                if (tid == 0) {
                    if (stream_out) {
                        if (want_j) tensor_store_2d(&a.jmap, smem_u32(j_img), int(k0), 0, pol_out);
                        if (want_o) tensor_store_2d(&a.omap, smem_u32(o_img), int(k0), 0, pol_out);
                    } else {
                        if (want_j) tensor_store_2d(&a.jmap, smem_u32(j_img), int(k0), 0);
                        if (want_o) tensor_store_2d(&a.omap, smem_u32(o_img), int(k0), 0);
                    }
                    bulk_commit();
                }
            } else if constexpr (S::ROWSTORE) {
                if (a.use_jmap) {
                    if (tid == 0) {
                        if (stream_out) {
                            if (want_j) tensor_store_2d(&a.jmap, smem_u32(j_img), 0, int(k0), pol_out);
                            if (want_o) bulk_store(a.out + k0 * n, smem_u32(o_img), uint32_t(S::o_bytes), pol_out);
                        } else {
                            if (want_j) tensor_store_2d(&a.jmap, smem_u32(j_img), 0, int(k0));
                            if (want_o) bulk_store(a.out + k0 * n, smem_u32(o_img), uint32_t(S::o_bytes));
                        }
                        bulk_commit();
                    }
                } else if (tid < TILE) {
                    if (want_j) bulk_store(a.J + (k0 + tid) * E, smem_u32(j_img + tid * S::PJ), uint32_t(E * sizeof(T)));
                    if (tid == 0 && want_o) bulk_store(a.out + k0 * n, smem_u32(o_img), uint32_t(S::o_bytes));
                    bulk_commit();
                }
            } else if (tid == 0) {
                if (stream_out) {
                    if (want_j) bulk_store(a.J + k0 * E, smem_u32(j_img), uint32_t(S::j_dense_bytes), pol_out);
                    if (want_o) bulk_store(a.out + k0 * n, smem_u32(o_img), uint32_t(S::o_bytes), pol_out);
                } else {
                    if (want_j) bulk_store(a.J + k0 * E, smem_u32(j_img), uint32_t(S::j_dense_bytes));
                    if (want_o) bulk_store(a.out + k0 * n, smem_u32(o_img), uint32_t(S::o_bytes));
                }
                bulk_commit();
            }
        } else {
            __syncthreads();
            if (want_j) coop_store(j_img, S::PJ, a.J, k0, cnt, E, tid, NTHR);
            if (want_o) coop_store(o_img, n, a.out, k0, cnt, n, tid, NTHR);
        }
    }
    // the CTA may retire once its last stores have READ the shared-memory images (the writes themselves complete, like every other
    // outstanding memory operation, before the grid does)
    if (tid < S::ISSUERS) bulk_wait_read0();
}

// ---- ImplicitMidpoint ----------------------------------------------------------------------------------------------------
// x2 solves  x1 + h f((x1 + x2)/2, u1) - x2 = 0  (reference: src/integration.jl:620-694): Newton from x2 = x1, at most 10 iterations,
// residual and Jacobians evaluated before the convergence test ||r||_2 < tol, step dx = (h/2 A - I) \ r by LU with partial
// pivoting (src/integration.jl:422-463, src/utils.jl:32-54); Jacobian by the implicit function theorem, J = -(h/2 A - I) \ [I + h/2 A, h B]
// (src/integration.jl:524-543).  One thread per knot: the continuous Jacobian comes from ONE forward-mode evaluation with all n+m
// columns seeded (it is sparse, so it fits registers); the small dense solves run on per-thread local arrays.
template <class T, int N_>
__device__ __forceinline__ void lu_solve_inplace(T* A /* N_ x N_ col-major, destroyed */, int nrhs, T* B /* N_ x nrhs col-major */) {
    for (int k = 0; k < N_; ++k) {
        int p = k; T best = fabs(A[k + N_ * k]);
        for (int i = k + 1; i < N_; ++i) { const T v = fabs(A[i + N_ * k]); if (v > best) { best = v; p = i; } }
        if (p != k) {
            for (int j = 0; j < N_; ++j) { const T t = A[k + N_ * j]; A[k + N_ * j] = A[p + N_ * j]; A[p + N_ * j] = t; }
            for (int j = 0; j < nrhs; ++j) { const T t = B[k + N_ * j]; B[k + N_ * j] = B[p + N_ * j]; B[p + N_ * j] = t; }
        }
        const T inv = T(1) / A[k + N_ * k];
        for (int i = k + 1; i < N_; ++i) {
            const T l = A[i + N_ * k] * inv;
            for (int j = k + 1; j < N_; ++j) A[i + N_ * j] -= l * A[k + N_ * j];
            for (int j = 0; j < nrhs; ++j) B[i + N_ * j] -= l * B[k + N_ * j];
        }
    }
    for (int j = 0; j < nrhs; ++j)
        for (int i = N_ - 1; i >= 0; --i) {
            T s = B[i + N_ * j];
            for (int c = i + 1; c < N_; ++c) s -= A[i + N_ * c] * B[c + N_ * j];
            B[i + N_ * j] = s / A[i + N_ * i];
        }
}

template <class Model, class T, bool WITH_J>
__global__ void __launch_bounds__(128) implicit_midpoint_kernel(const Model model, const KnotArgs<T> a) {
    constexpr int n = Model::n, m = Model::m, NZ = n + m;
    constexpr mask_t ALL = (NZ >= 32) ? ~mask_t(0) : ((mask_t(1) << NZ) - 1u);
    const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (k >= a.N) return;
    const T* z = a.Z + k * NZ;
    const T h = T(a.dt ? a.dt[k] : a.dt0);
    // the midpoint time t + h/2 of dynamics_error (src/integration.jl:652-653); the reference's Jacobian evaluates at t (:693), an
    // inconsistency with its own residual that is not reproduced
    T tm = T(0);
    if constexpr (uses_time<Model>::value) tm = T(a.t ? a.t[k] : 0.0) + T(0.5) * h;
    const T tol = sizeof(T) == 8 ? T(1e-12) : T(1e-5);
    T zm[NZ], x2[n], r[n], J1[n * NZ], A2[n * n], W[n * n];
#pragma unroll
    for (int i = 0; i < n; ++i) x2[i] = z[i];
#pragma unroll
    for (int i = 0; i < m; ++i) zm[n + i] = z[n + i];
    for (int iter = 0; iter < 10; ++iter) {
#pragma unroll
        for (int i = 0; i < n; ++i) zm[i] = (z[i] + x2[i]) * T(0.5);
        model.reset();
        auto zz = load_seeded<T, ALL>(zm, rstd::make_index_sequence<size_t(NZ)>{});
        auto f = feval<T>(model, slice<0, n>(zz), slice<n, m>(zz), tm);
        put_vals(f, r, rstd::make_index_sequence<size_t(n)>{});
        put_cols<n, ALL, false>(f, J1, rstd::make_index_sequence<size_t(NZ)>{});
        T nrm = T(0);
#pragma unroll
        for (int i = 0; i < n; ++i) { r[i] = z[i] + h * r[i] - x2[i]; nrm += r[i] * r[i]; }
        for (int i = 0; i < n * NZ; ++i) J1[i] *= h;
        for (int i = 0; i < n * n; ++i) { J1[i] *= T(0.5); A2[i] = J1[i]; }
        for (int i = 0; i < n; ++i) { J1[i + n * i] += T(1); A2[i + n * i] -= T(1); }
        if (sqrt(nrm) < tol) break;
        for (int i = 0; i < n * n; ++i) W[i] = A2[i];
        lu_solve_inplace<T, n>(W, 1, r);
        for (int i = 0; i < n; ++i) x2[i] -= r[i];
    }
    if (a.out) { T* o = a.out + k * n; for (int i = 0; i < n; ++i) o[i] = x2[i]; }
    if constexpr (WITH_J) {
        if (a.J) {
            for (int i = 0; i < n * n; ++i) W[i] = A2[i];
            for (int i = 0; i < n * NZ; ++i) J1[i] = -J1[i];
            lu_solve_inplace<T, n>(W, NZ, J1);
            T* Jo = a.J + k * (long long)(n * NZ);
            for (int i = 0; i < n * NZ; ++i) Jo[i] = J1[i];
        }
    }
}

// ---- ImplicitMidpoint, warp-cooperative (rigid bodies: n >= 8) -----------------------------------------------------------------
// The per-thread kernel above keeps ~600 words of matrices per knot in local memory.  Here a group of GS lanes (GS = 32 for the
// rigid bodies: n + m = 16..19) shares ONE knot and lane j owns COLUMN j of every matrix, in registers:
//   * [A B] by forward mode with a single runtime-seeded partial per lane (lane j seeds z_j): one evaluation of f yields the
//     whole continuous Jacobian, one column per lane;
//   * lane j < n holds column j of  M = h/2 A - I,  lane j < n+m column j of the right-hand side  [I + h/2 A, h B];  the Newton
//     residual is replicated on every lane;
//   * LU with partial pivoting runs across the lanes: lane k scans its column for the pivot and computes the multipliers, which
//     reach the other lanes by warp shuffles; every lane eliminates its own column and its own right-hand side; in the back
//     substitution the entries of U are broadcast from their owning lanes the same way.
// Same arithmetic as the reference loop (src/integration.jl:620-694, 524-543): Newton from x2 = x1, at most 10 iterations, residual
// and Jacobians evaluated before the test, Jacobian from the last evaluated iterate.
template <class T> __device__ __forceinline__ SD<T, 1u> lane_dual(T v, bool mine) { SD<T, 1u> r; r.v = v; r.d[0] = PK<T>::splat(mine ? T(1) : T(0)); return r; }
template <class T, size_t... Is>
__device__ __forceinline__ auto load_lane_seeded(const T* z, int col, rstd::index_sequence<Is...>) { return vec(lane_dual<T>(z[Is], col == int(Is))...); }

// (p == i) ? a : b as an opaque select: written with ?: the compiler recognises the chain over i as the dynamically indexed
// access W[p] and moves the whole register array to local memory
__device__ __forceinline__ float sel_eq(int p, int i, float a, float b) {
    float r;
    asm("{\n\t.reg .pred q;\n\tsetp.eq.s32 q, %3, %4;\n\tselp.f32 %0, %1, %2, q;\n\t}" : "=f"(r) : "f"(a), "f"(b), "r"(p), "r"(i));
    return r;
}
__device__ __forceinline__ double sel_eq(int p, int i, double a, double b) {
    double r;
    asm("{\n\t.reg .pred q;\n\tsetp.eq.s32 q, %3, %4;\n\tselp.f64 %0, %1, %2, q;\n\t}" : "=d"(r) : "d"(a), "d"(b), "r"(p), "r"(i));
    return r;
}

// B <- M^{-1} B where lane j < N_ of each GS-lane group holds column j of M in W (destroyed) and B is the lane's own vector
template <class T, int N_, int GS>
__device__ __forceinline__ void warp_lu_solve(T (&W)[N_], T (&B)[N_]) {
    constexpr unsigned FULL = 0xffffffffu;
    T dinv[N_];
#pragma unroll
    for (int k = 0; k < N_; ++k) {
        // partial pivoting.  Common case first: the diagonal entry already has the largest magnitude of its column (M = h/2 A - I
        // is close to -I) — one max per row on lane k, one shuffle and a warp-uniform branch; the full search and the row swap
        // (selects over the register rows: p is only known at run time) run only when some group of the warp needs them.
        T mx = T(0);
#pragma unroll
        for (int i = k + 1; i < N_; ++i) mx = fmax(mx, fabs(W[i]));
        const int need = __shfl_sync(FULL, int(mx > fabs(W[k])), k, GS);
        if (__any_sync(FULL, need)) {
            int p = k;
            T best = fabs(W[k]);
#pragma unroll
            for (int i = k + 1; i < N_; ++i) { const T v = fabs(W[i]); if (v > best) { best = v; p = i; } }
            p = __shfl_sync(FULL, p, k, GS);
            const T wk = W[k], bk = B[k];
            T wp = wk, bp = bk;
#pragma unroll
            for (int i = k + 1; i < N_; ++i) {
                wp = sel_eq(p, i, W[i], wp); bp = sel_eq(p, i, B[i], bp);
                W[i] = sel_eq(p, i, wk, W[i]); B[i] = sel_eq(p, i, bk, B[i]);
            }
            W[k] = wp; B[k] = bp;
        }
        dinv[k] = __shfl_sync(FULL, T(1) / W[k], k, GS);
#pragma unroll
        for (int i = k + 1; i < N_; ++i) {
            const T l = __shfl_sync(FULL, W[i], k, GS) * dinv[k];
            W[i] -= l * W[k];
            B[i] -= l * B[k];
        }
    }
#pragma unroll
    for (int i = N_ - 1; i >= 0; --i) {           // back substitution: U(i,c) lives on lane c
        T s = B[i];
#pragma unroll
        for (int c = i + 1; c < N_; ++c) s -= __shfl_sync(FULL, W[i], c, GS) * B[c];
        B[i] = s * dinv[i];
    }
}

template <class Model, class T, bool WITH_J, int GS>
__global__ void __launch_bounds__(128) implicit_midpoint_warp_kernel(const Model model, const KnotArgs<T> a) {
    constexpr int n = Model::n, m = Model::m, NZ = n + m;
    static_assert(NZ <= GS && GS <= 32 && (32 % GS) == 0, "one lane per column of [x;u]");
    constexpr unsigned FULL = 0xffffffffu;
    const long long gtid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int col = int(threadIdx.x) % GS;                     // the column of [x;u] this lane owns
    const bool valid = gtid / GS < a.N;
    const long long k = valid ? gtid / GS : a.N - 1;           // lanes past the end shadow the last knot (shuffles stay convergent)
    const T* zg = a.Z + k * NZ;
    const T h = T(a.dt ? a.dt[k] : a.dt0);
    T tm = T(0);
    if constexpr (uses_time<Model>::value) tm = T(a.t ? a.t[k] : 0.0) + T(0.5) * h;
    const T tol = sizeof(T) == 8 ? T(1e-12) : T(1e-5);
    T z[NZ], zm[NZ], x2[n], Mf[n], Rf[n];
#pragma unroll
    for (int i = 0; i < NZ; ++i) { z[i] = zg[i]; zm[i] = z[i]; }
#pragma unroll
    for (int i = 0; i < n; ++i) { x2[i] = z[i]; Mf[i] = T(0); Rf[i] = T(0); }
    bool done = false;
#pragma unroll 1
    for (int iter = 0; iter < 10; ++iter) {
#pragma unroll
        for (int i = 0; i < n; ++i) zm[i] = (z[i] + x2[i]) * T(0.5);
        model.reset();
        auto zz = load_lane_seeded<T>(zm, col, rstd::make_index_sequence<size_t(NZ)>{});
        auto f = feval<T>(model, slice<0, n>(zz), slice<n, m>(zz), tm);
        T r[n], acol[n];
        put_vals(f, r, rstd::make_index_sequence<size_t(n)>{});
        put_cols<n, 1u, false>(f, acol, rstd::make_index_sequence<size_t(1)>{});     // column `col` of [A B]
        T nrm = T(0);
#pragma unroll
        for (int i = 0; i < n; ++i) { r[i] = z[i] + h * r[i] - x2[i]; nrm += r[i] * r[i]; }
        T W[n];
#pragma unroll
        for (int i = 0; i < n; ++i) {
            const T ha = h * acol[i], hha = T(0.5) * ha, e = (i == col) ? T(1) : T(0);
            W[i] = hha - e;                                    // column of  h/2 A - I            (col < n)
            if (!done) { Mf[i] = W[i]; Rf[i] = col < n ? hha + e : ha; }     // column of  [I + h/2 A, h B]
        }
        const bool conv = sqrt(nrm) < tol;                     // the residual is replicated: uniform within the group
        if (!done && conv) done = true;
        if (__all_sync(FULL, done)) break;
        warp_lu_solve<T, n, GS>(W, r);                         // all lanes take part; groups that are done discard the step
        if (!done) {
#pragma unroll
            for (int i = 0; i < n; ++i) x2[i] -= r[i];
        }
    }
    if (valid && col == 0 && a.out) { T* o = a.out + k * n;
#pragma unroll
        for (int i = 0; i < n; ++i) o[i] = x2[i]; }
    if constexpr (WITH_J) {
        if (a.J) {                                             // uniform over the grid
            warp_lu_solve<T, n, GS>(Mf, Rf);                   // J = -(h/2 A - I) \ [I + h/2 A, h B], one column per lane
            if (valid && col < NZ) { T* Jo = a.J + k * (long long)(n * NZ) + n * col;
#pragma unroll
                for (int i = 0; i < n; ++i) Jo[i] = -Rf[i]; }
        }
    }
}

// ---- ImplicitMidpoint, group-cooperative (round 2) -------------------------------------------------------------------------------------
// The warp kernel above spends one warp on ONE knot and broadcasts every multiplier of the elimination by a shuffle that serves one
// knot: ~730 shuffles per knot, and 32 lanes evaluating f with one partial each.  Here a group of L lanes (4 for fp32, 8 for fp64:
// 8 / 4 knots per warp) shares a knot and lane g owns the columns c = g, g + L, g + 2L, ... of every matrix, in registers:
//   * [A B] by ONE forward-mode evaluation of f per lane with S = ceil((n+m)/L) run-time-seeded partials (slot s of lane g is column
//     s L + g);
//   * M = h/2 A - I and the right-hand sides are eliminated by Gauss-Jordan with partial pivoting: the lane that owns pivot column k
//     finds the pivot and the multipliers, ONE shuffle per multiplier serves all the knots of the warp, and every lane updates its own
//     columns (the pivot-row entries it needs are its own);  the Newton residual is replicated on the lanes and eliminated along.
// Same iteration as the reference (src/integration.jl:422-463: Newton from x2 = x1, at most 10 iterations, residual and Jacobians
// evaluated before the test ||r||_2 < tol) and the same implicit-function-theorem Jacobian (src/integration.jl:524-543).
template <class T, int S, int L, int I>
__device__ __forceinline__ auto group_dual(T v, int g) {
    SD<T, ((mask_t(1) << S) - 1u)> r; r.v = v;
#pragma unroll
    for (int s = 0; s < S; ++s) r.d[s] = PK<T>::splat((s * L + g == I) ? T(1) : T(0));
    return r;
}
template <class T, int S, int L, size_t... Is>
__device__ __forceinline__ auto load_group_seeded(const T* z, int g, rstd::index_sequence<Is...>) { return vec(group_dual<T, S, L, int(Is)>(z[Is], g)...); }

// Gauss-Jordan on [M | R | b]: lane g of each L-lane group holds the columns s L + g of M (SA slots) and of R (SR slots, may be 0) and a
// replicated copy of b.  On return R and b hold M^{-1} R and M^{-1} b.  Rows are swapped only when some group of the warp needs it.
template <class T, int N_, int SA, int SR, int L, bool WITH_B>
__device__ __forceinline__ void group_gauss_jordan(T (&M)[SA][N_], T (&R)[SR > 0 ? SR : 1][N_], T (&b)[N_], int g) {
    constexpr unsigned FULL = 0xffffffffu;
    T pinv[N_];
#pragma unroll
    for (int k = 0; k < N_; ++k) {
        constexpr int dummy = 0; (void)dummy;
        const int sk = k / L, gk = k % L;              // slot and owner lane of pivot column k (compile-time after unrolling)
        T mx = T(0);
#pragma unroll
        for (int i = k + 1; i < N_; ++i) mx = fmax(mx, fabs(M[sk][i]));
        const int need = __shfl_sync(FULL, int(mx > fabs(M[sk][k])), gk, L);
        if (__any_sync(FULL, need)) {
            int p = k;
            T best = fabs(M[sk][k]);
#pragma unroll
            for (int i = k + 1; i < N_; ++i) { const T v = fabs(M[sk][i]); if (v > best) { best = v; p = i; } }
            p = __shfl_sync(FULL, p, gk, L);
#pragma unroll
            for (int s = 0; s < SA; ++s) {
                const T wk = M[s][k]; T wp = wk;
#pragma unroll
                for (int i = k + 1; i < N_; ++i) { wp = sel_eq(p, i, M[s][i], wp); M[s][i] = sel_eq(p, i, wk, M[s][i]); }
                M[s][k] = wp;
            }
            if constexpr (SR > 0) {
#pragma unroll
                for (int s = 0; s < SR; ++s) {
                    const T wk = R[s][k]; T wp = wk;
#pragma unroll
                    for (int i = k + 1; i < N_; ++i) { wp = sel_eq(p, i, R[s][i], wp); R[s][i] = sel_eq(p, i, wk, R[s][i]); }
                    R[s][k] = wp;
                }
            }
            if constexpr (WITH_B) {
                const T wk = b[k]; T wp = wk;
#pragma unroll
                for (int i = k + 1; i < N_; ++i) { wp = sel_eq(p, i, b[i], wp); b[i] = sel_eq(p, i, wk, b[i]); }
                b[k] = wp;
            }
        }
        const T inv = T(1) / M[sk][k];
        pinv[k] = __shfl_sync(FULL, inv, gk, L);
#pragma unroll
        for (int i = 0; i < N_; ++i) {
            if (i == k) continue;
            const T l = __shfl_sync(FULL, M[sk][i] * inv, gk, L);
#pragma unroll
            for (int s = sk; s < SA; ++s) M[s][i] = fma_(-l, M[s][k], M[s][i]);       // columns left of the pivot are already e_c D_c
            if constexpr (SR > 0) {
#pragma unroll
                for (int s = 0; s < SR; ++s) R[s][i] = fma_(-l, R[s][k], R[s][i]);
            }
            if constexpr (WITH_B) b[i] = fma_(-l, b[k], b[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < N_; ++i) {
        if constexpr (WITH_B) b[i] *= pinv[i];
        if constexpr (SR > 0) {
#pragma unroll
            for (int s = 0; s < SR; ++s) R[s][i] *= pinv[i];
        }
    }
}

// threads per CTA of the group kernel: two warps, so that the per-warp output image (below) stays a static allocation for fp64
constexpr int IMG_THREADS = 64;
template <class Model, class T, bool WITH_J, int L>
__global__ void __launch_bounds__(IMG_THREADS) implicit_midpoint_group_kernel(const Model model, const KnotArgs<T> a) {
    constexpr int n = Model::n, m = Model::m, NZ = n + m;
    constexpr int SA = (n + L - 1) / L, SZ = (NZ + L - 1) / L;
    static_assert(SZ <= 8 && (32 % L) == 0, "group size too small for this model");
    // The Jacobians of the 32 / L knots of a warp are ONE contiguous byte range of J.  Each lane owns scattered columns of them, so the
    // columns are assembled in a per-warp shared-memory image and leave by full-line coalesced stores (the direct form — 4-byte stores
    // 52 bytes apart within a knot and 884 bytes apart between knots — showed up as lg_throttle / long_scoreboard stalls in ncu).
    constexpr int KPW = 32 / L, E = n * NZ;
    constexpr bool STAGE = WITH_J && size_t(IMG_THREADS / 32) * KPW * E * sizeof(T) <= 40 * 1024;
    __shared__ T jimg[STAGE ? (IMG_THREADS / 32) * KPW * E : 1];
    constexpr mask_t DENSE = (mask_t(1) << SZ) - 1u;
    constexpr unsigned FULL = 0xffffffffu;
    const long long gtid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int g = int(threadIdx.x) % L;
    const bool valid = gtid / L < a.N;
    const long long k = valid ? gtid / L : a.N - 1;           // lanes past the end shadow the last knot (shuffles stay convergent)
    const T* zg = a.Z + k * NZ;
    const T h = T(a.dt ? a.dt[k] : a.dt0);
    T tm = T(0);
    if constexpr (uses_time<Model>::value) tm = T(a.t ? a.t[k] : 0.0) + T(0.5) * h;
    const T tol = sizeof(T) == 8 ? T(1e-12) : T(1e-5);
    T z[NZ], zm[NZ], x2[n], W[SZ][n], r[n];
#pragma unroll
    for (int i = 0; i < NZ; ++i) { z[i] = zg[i]; zm[i] = z[i]; }
#pragma unroll
    for (int i = 0; i < n; ++i) x2[i] = z[i];
    bool done = false;
#pragma unroll 1
    for (int iter = 0; iter < 10; ++iter) {
#pragma unroll
        for (int i = 0; i < n; ++i) zm[i] = (z[i] + x2[i]) * T(0.5);
        model.reset();
        {
            auto zz = load_group_seeded<T, SZ, L>(zm, g, rstd::make_index_sequence<size_t(NZ)>{});
            auto f = feval<T>(model, slice<0, n>(zz), slice<n, m>(zz), tm);
            put_vals(f, r, rstd::make_index_sequence<size_t(n)>{});
            put_cols<n, DENSE, false>(f, &W[0][0], rstd::make_index_sequence<size_t(SZ)>{});      // W[s][i] = d f_i / d z_{s L + g}
        }
        T nrm = T(0);
#pragma unroll
        for (int i = 0; i < n; ++i) { r[i] = z[i] + h * r[i] - x2[i]; nrm += r[i] * r[i]; }
        const bool conv = sqrt(nrm) < tol;                     // the residual is replicated: uniform within the group
        if (!done && conv) done = true;
        if (__all_sync(FULL, done)) break;
        T Mc[SA][n], none[1][n];
#pragma unroll
        for (int s = 0; s < SA; ++s)
#pragma unroll
            for (int i = 0; i < n; ++i) Mc[s][i] = (s * L + g < n) ? T(0.5) * h * W[s][i] - ((s * L + g == i) ? T(1) : T(0)) : T(0);
        group_gauss_jordan<T, n, SA, 0, L, true>(Mc, none, r, g);   // all lanes take part; groups that are done discard the step
        if (!done) {
#pragma unroll
            for (int i = 0; i < n; ++i) x2[i] -= r[i];
        }
    }
    if (valid && g == 0 && a.out) { T* o = a.out + k * n;
#pragma unroll
        for (int i = 0; i < n; ++i) o[i] = x2[i]; }
    if constexpr (WITH_J) {
        if (a.J) {                                             // uniform over the grid
            // J = -(h/2 A - I) \ [I + h/2 A, h B] from the last evaluated iterate, one column per (lane, slot)
            T Mc[SA][n];
#pragma unroll
            for (int s = 0; s < SA; ++s)
#pragma unroll
                for (int i = 0; i < n; ++i) Mc[s][i] = (s * L + g < n) ? T(0.5) * h * W[s][i] - ((s * L + g == i) ? T(1) : T(0)) : T(0);
#pragma unroll
            for (int s = 0; s < SZ; ++s)
#pragma unroll
                for (int i = 0; i < n; ++i) {
                    const int c = s * L + g;
                    const T ha = h * W[s][i];
                    W[s][i] = c < n ? T(0.5) * ha + ((c == i) ? T(1) : T(0)) : (c < NZ ? ha : T(0));
                }
            group_gauss_jordan<T, n, SA, SZ, L, false>(Mc, W, r, g);
            if constexpr (STAGE) {
                const int lane = int(threadIdx.x) & 31, warp = int(threadIdx.x) >> 5;
                T* img = jimg + warp * (KPW * E);
#pragma unroll
                for (int s = 0; s < SZ; ++s) {
                    const int c = s * L + g;
                    if (c < NZ) { T* col = img + (lane / L) * E + n * c;
#pragma unroll
                        for (int i = 0; i < n; ++i) col[i] = -W[s][i]; }
                }
                __syncwarp();
                const long long k0w = (gtid - lane) / L;                       // first knot of this warp
                const long long left = a.N - k0w;
                const int cnt = left >= KPW ? KPW * E : (left > 0 ? int(left) * E : 0);
                T* Jo = a.J + k0w * (long long)E;
                for (int e = lane; e < cnt; e += 32) Jo[e] = img[e];
            } else {
#pragma unroll
                for (int s = 0; s < SZ; ++s) {
                    const int c = s * L + g;
                    if (valid && c < NZ) { T* Jo = a.J + k * (long long)(n * NZ) + n * c;
#pragma unroll
                        for (int i = 0; i < n; ++i) Jo[i] = -W[s][i]; }
                }
            }
        }
    }
}

// ---- dynamics_error / dynamics_error_jacobian! for ImplicitMidpoint ---------------------------------------------------------------
// e = x1 + h f((x1 + x2)/2, u1, t + h/2) - x2   (reference: src/integration.jl:640-654);  J1 = de/dz1 = [I + h/2 A, h B],
// J2 = de/dz2 = [h/2 A - I, 0]  (src/integration.jl:674-700), A, B the continuous Jacobian at the midpoint (evaluated at t + h/2 like
// the residual; the reference's Jacobian passes t there, an inconsistency with its own residual that is not reproduced).
// One thread per knot pair, ONE forward-mode evaluation of f with every column seeded (sparse: it fits registers).
template <class T>
struct DynErrArgs {
    const T* Z1;         // (N, n+m): z1 = [x1; u1]
    const T* Z2;         // (N, ld2): x2 in the first n entries of every row
    int ld2;
    const double* t; const double* dt; double dt0;
    T* J2; T* J1;        // (N, n+m, n) each: column-major n x (n+m) per knot; may be nullptr
    T* e;                // (N, n); may be nullptr
    long long N;
};
template <class Model, class T, bool WITH_J>
__global__ void __launch_bounds__(128) midpoint_error_kernel(const Model model, const DynErrArgs<T> a) {
    constexpr int n = Model::n, m = Model::m, NZ = n + m;
    constexpr mask_t ALL = (NZ >= 32) ? ~mask_t(0) : ((mask_t(1) << NZ) - 1u);
    const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (k >= a.N) return;
    const T* z1 = a.Z1 + k * NZ;
    const T* x2 = a.Z2 + k * (long long)a.ld2;
    const T h = T(a.dt ? a.dt[k] : a.dt0);
    T tm = T(0);
    if constexpr (uses_time<Model>::value) tm = T(a.t ? a.t[k] : 0.0) + T(0.5) * h;
    T zm[NZ], f[n];
#pragma unroll
    for (int i = 0; i < n; ++i) zm[i] = (z1[i] + x2[i]) * T(0.5);
#pragma unroll
    for (int i = 0; i < m; ++i) zm[n + i] = z1[n + i];
    model.reset();
    if constexpr (WITH_J) {
        T AB[n * NZ];
        auto zz = load_seeded<T, ALL>(zm, rstd::make_index_sequence<size_t(NZ)>{});
        auto fd = feval<T>(model, slice<0, n>(zz), slice<n, m>(zz), tm);
        put_vals(fd, f, rstd::make_index_sequence<size_t(n)>{});
        put_cols<n, ALL, false>(fd, AB, rstd::make_index_sequence<size_t(NZ)>{});
        T* J1 = a.J1 ? a.J1 + k * (long long)(n * NZ) : nullptr;
        T* J2 = a.J2 ? a.J2 + k * (long long)(n * NZ) : nullptr;
#pragma unroll
        for (int j = 0; j < NZ; ++j)
#pragma unroll
            for (int i = 0; i < n; ++i) {
                const T ha = h * AB[i + n * j], hha = T(0.5) * ha, d = (i == j) ? T(1) : T(0);
                if (J1) J1[i + n * j] = j < n ? hha + d : ha;
                if (J2) J2[i + n * j] = j < n ? hha - d : T(0);
            }
    } else {
        auto zz = load_seeded<T, mask_t(0)>(zm, rstd::make_index_sequence<size_t(NZ)>{});
        auto fd = feval<T>(model, slice<0, n>(zz), slice<n, m>(zz), tm);
        put_vals(fd, f, rstd::make_index_sequence<size_t(n)>{});
    }
    if (a.e) {
        T* e = a.e + k * n;
#pragma unroll
        for (int i = 0; i < n; ++i) e[i] = z1[i] + h * f[i] - x2[i];
    }
}
// explicit rules: e = discrete_dynamics(z1) - x2 and J2 = [-I 0] (reference: src/discrete_dynamics.jl:137-138,181-182); the knot kernel
// has already written discrete_dynamics(z1) into e and its Jacobian into J1
template <class T>
__global__ void __launch_bounds__(256) explicit_error_fixup_kernel(int n, int m, long long N, const T* __restrict__ Z2, int ld2, T* __restrict__ e,
                                                                   T* __restrict__ J2) {
    const int per = n * (n + m);
    const long long total = N * (long long)(J2 ? per : n);
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        if (J2) {
            const long long k = idx / per;
            const int r = int(idx - k * per), j = r / n, i = r - j * n;
            J2[idx] = (i == j) ? T(-1) : T(0);
            if (e && j == 0) e[k * n + i] -= Z2[k * (long long)ld2 + i];
        } else {
            const long long k = idx / n;
            e[idx] -= Z2[k * (long long)ld2 + int(idx - k * n)];
        }
    }
}

// rollout!: x_{k+1} = discrete_dynamics(x_k, u_k, t_k, dt_k), sequential in k, one thread per trajectory
// (reference: src/trajectories.jl:436-441, src/discrete_dynamics.jl:217-235).
//
// Rows: knot k of trajectory j lives in row  j * sj + k * sk  of X (row width ldx), of dt and of t.
//   trajectory-major (rdb_rollout):      sj = K, sk = 1, ldx = n, controls from U (rows j * uj + k * uk, width m)
//   knot-major batch (rdb_trajectory_*): sj = 1, sk = ntraj, ldx = n + m, U == nullptr: the controls are read from the SAME rows
//     (Z = [x; u] per knot, like the reference's rollout! that reads control(Z[k]) and writes state(Z[k+1])); adjacent threads own
//     adjacent rows, so every step's stores and loads of a warp are one contiguous range.
// Steps k in [kb, ke) are computed (x_{k+1} written for each); kb == 0 with x0 != nullptr also writes row 0.  Time: t[row] when
// given, else accumulated from 0 (times of a SampledTrajectory built from dt, src/trajectories.jl:82-110).
template <class T>
struct RolloutArgs {
    const T* x0;         // (ntraj, n) or nullptr: the states already in row 0 of X
    const T* U;          // controls, or nullptr (Z mode)
    const double* dt;    // per-row steps or nullptr -> dt0
    const double* t;     // per-row times or nullptr
    double dt0;
    T* X;
    long long ntraj;
    int K, kb, ke;
    long long sj, sk, uj, uk;
    int ldx;
};
template <class T, size_t... Is> __device__ __forceinline__ auto load_plain(const T* p, rstd::index_sequence<Is...>) { return vec(p[Is]...); }
template <class Model, int Q, class T>
__global__ void __launch_bounds__(256) rollout_kernel(const Model model, const RolloutArgs<T> a) {
    constexpr int n = Model::n, m = Model::m;
    const long long tr = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (tr >= a.ntraj) return;
    const long long base = tr * a.sj;
    T* X = a.X;
    const T* Usrc = a.U ? a.U + tr * a.uj * m : X + base * a.ldx + n;          // controls of knot 0
    const long long ustep = a.U ? a.uk * m : a.sk * (long long)a.ldx;
    auto x = load_plain((a.kb == 0 && a.x0) ? a.x0 + tr * n : X + (base + a.kb * a.sk) * a.ldx, rstd::make_index_sequence<size_t(n)>{});
    if (a.kb == 0 && a.x0) put_vals(x, X + base * a.ldx, rstd::make_index_sequence<size_t(n)>{});
    T tacc = T(0);
    if constexpr (uses_time<Model>::value) {
        if (!a.t) { for (int k = 0; k < a.kb; ++k) tacc += T(a.dt ? a.dt[base + k * a.sk] : a.dt0); }
    }
    // software prefetch: the controls and step of knot k + 1 are requested before the arithmetic of knot k
    T ucur[m];
    double hcur = a.dt0;
    if (a.kb < a.ke) {
#pragma unroll
        for (int i = 0; i < m; ++i) ucur[i] = Usrc[a.kb * ustep + i];
        if (a.dt) hcur = a.dt[base + a.kb * a.sk];
    }
    for (int k = a.kb; k < a.ke; ++k) {
        T unext[m];
        double hnext = a.dt0;
        const bool more = k + 1 < a.ke;
#pragma unroll
        for (int i = 0; i < m; ++i) unext[i] = more ? Usrc[(k + 1) * ustep + i] : T(0);
        if (a.dt && more) hnext = a.dt[base + (k + 1) * a.sk];
        auto u = load_plain(ucur, rstd::make_index_sequence<size_t(m)>{});
        const T h = T(hcur);
        T tk = tacc;
        if constexpr (uses_time<Model>::value) { if (a.t) tk = T(a.t[base + k * a.sk]); }
        model.reset();
        x = integrate<Q, T>(model, x, u, h, tk);
        put_vals(x, X + (base + (k + 1) * a.sk) * a.ldx, rstd::make_index_sequence<size_t(n)>{});
        tacc += h;
#pragma unroll
        for (int i = 0; i < m; ++i) ucur[i] = unext[i];
        hcur = hnext;
    }
}

}  // namespace rdb
