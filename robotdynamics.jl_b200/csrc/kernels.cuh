// kernels.cuh — the batched knot-point kernel (K1/K2/K3 of SURVEY.md §2) for sm_100a.
//
// One template serves dynamics(), discrete_dynamics() and their Jacobians:
//   knot_kernel<Model, Q, T, TILE, WITH_J, CHUNKS...>
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
//   * component-major ("SoA") callers, unaligned pointers and the ragged last tile use a cooperative
//     coalesced copy between the same shared-memory images and global memory.
#pragma once
#include <cuda_runtime.h>
#include "integrators.cuh"

namespace rdb {

enum Layout { LAYOUT_AOS = 0, LAYOUT_SOA = 1 };

template <class T>
struct KnotArgs {
    const T* Z;          // [x;u] per knot: AOS (N, n+m) knot-major  |  SOA (n+m, N) component-major
    const double* dt;    // per-knot step (N) or nullptr -> dt0      (KnotPoint.dt is Float64: src/knotpoint.jl:148-153)
    double dt0;
    T* J;                // AOS (N, n+m, n) == per knot n x (n+m) column-major | SOA (n*(n+m), N);  may be nullptr
    T* out;              // xdot or x+ : AOS (N, n) | SOA (n, N);  may be nullptr
    long long N;
    int layout;
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
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- compile-time helpers ---------------------------------------------------------------------------------
template <mask_t... Ms> struct MaskList { static constexpr int count = int(sizeof...(Ms)); };
template <class T, mask_t CHUNK, size_t... Is>
__device__ __forceinline__ auto load_seeded(const T* zrow, std::index_sequence<Is...>) {
    return vec(seed<T, int(Is), CHUNK>(zrow[Is])...);
}
template <class T> __device__ __forceinline__ T plain(const T& a) { return a; }
template <class T, mask_t M> __device__ __forceinline__ T plain(const SD<T, M>& a) { return a.v; }

template <int N_, mask_t CHUNK, int J, class T, class XN, size_t... Is>
__device__ __forceinline__ void put_col(const XN& xn, T* jrow, std::index_sequence<Is...>) {
    if constexpr (chas(CHUNK, J)) { ((jrow[int(Is) + N_ * J] = partial<J>(get<int(Is)>(xn))), ...); }
}
template <int N_, mask_t CHUNK, class T, class XN, size_t... Js>
__device__ __forceinline__ void put_cols(const XN& xn, T* jrow, std::index_sequence<Js...>) {
    (put_col<N_, CHUNK, int(Js)>(xn, jrow, std::make_index_sequence<size_t(N_)>{}), ...);
}
template <class T, class XN, size_t... Is>
__device__ __forceinline__ void put_vals(const XN& xn, T* orow, std::index_sequence<Is...>) {
    ((orow[Is] = plain(get<int(Is)>(xn))), ...);
}

// all threads of the CTA meet here between "results are in registers" and "results go to the smem images":
// thread 0 first waits until the previous tile's bulk stores have finished reading those images.
template <int NTHR>
__device__ __forceinline__ void images_free_barrier(int tid) {
    if (tid == 0) bulk_wait_read0();
    asm volatile("bar.sync 1, %0;" ::"n"(NTHR) : "memory");
}

// One role: evaluate the map for one knot with partials for the columns in CHUNK; write this role's share.
template <class Model, int Q, class T, bool WITH_J, mask_t CHUNK, bool WRITE_OUT, int NTHR, int ROLL>
__device__ __forceinline__ void role_body(const Model& model, const T* zrow, T h, T* jrow, T* orow, int tid) {
    constexpr int n = Model::n, m = Model::m, NZ = n + m;
    auto zz = load_seeded<T, (WITH_J ? CHUNK : mask_t(0))>(zrow, std::make_index_sequence<size_t(NZ)>{});
    auto xn = integrate<Q, T, (WITH_J ? ROLL : 0)>(model, slice<0, n>(zz), slice<n, m>(zz), h);
    images_free_barrier<NTHR>(tid);
    if constexpr (WITH_J) put_cols<n, CHUNK>(xn, jrow, std::make_index_sequence<size_t(NZ)>{});
    if constexpr (WRITE_OUT) { if (orow) put_vals(xn, orow, std::make_index_sequence<size_t(n)>{}); }
}

template <int R, class L> struct list_at;
template <int R, mask_t M0, mask_t... Ms> struct list_at<R, MaskList<M0, Ms...>> { static constexpr mask_t value = list_at<R - 1, MaskList<Ms...>>::value; };
template <mask_t M0, mask_t... Ms> struct list_at<0, MaskList<M0, Ms...>> { static constexpr mask_t value = M0; };

template <class Model, int Q, class T, bool WITH_J, class Chunks, int NTHR, int ROLL, int R = 0>
__device__ __forceinline__ void dispatch_role(int role, const Model& model, const T* zrow, T h, T* jrow, T* orow, int tid) {
    if constexpr (R + 1 == Chunks::count) {
        role_body<Model, Q, T, WITH_J, list_at<R, Chunks>::value, R == 0, NTHR, ROLL>(model, zrow, h, jrow, orow, tid);
    } else {
        if (role == R) role_body<Model, Q, T, WITH_J, list_at<R, Chunks>::value, R == 0, NTHR, ROLL>(model, zrow, h, jrow, orow, tid);
        else dispatch_role<Model, Q, T, WITH_J, Chunks, NTHR, ROLL, R + 1>(role, model, zrow, h, jrow, orow, tid);
    }
}

// cooperative copies between a dense knot-major smem image [cnt][W] and global memory
template <class T>
__device__ __forceinline__ void coop_load(T* img, const T* g, long long k0, int cnt, int W, long long N, int layout, int tid, int nthr) {
    const int total = cnt * W;
    if (layout == LAYOUT_AOS) {
        const T* src = g + k0 * W;
        for (int i = tid; i < total; i += nthr) img[i] = src[i];
    } else {
        for (int i = tid; i < total; i += nthr) { const int c = i / cnt, kt = i - c * cnt; img[kt * W + c] = g[(long long)c * N + k0 + kt]; }
    }
}
template <class T>
__device__ __forceinline__ void coop_store(const T* img, T* g, long long k0, int cnt, int W, long long N, int layout, int tid, int nthr) {
    const int total = cnt * W;
    if (layout == LAYOUT_AOS) {
        T* dst = g + k0 * W;
        for (int i = tid; i < total; i += nthr) dst[i] = img[i];
    } else {
        for (int i = tid; i < total; i += nthr) { const int c = i / cnt, kt = i - c * cnt; g[(long long)c * N + k0 + kt] = img[kt * W + c]; }
    }
}

template <class Model, int TILE, bool WITH_J, class T>
struct KnotSmem {
    static constexpr int n = Model::n, NZ = Model::n + Model::m, E = Model::n * NZ;
    static constexpr size_t in_bytes = size_t(TILE) * NZ * sizeof(T);
    static constexpr size_t j_bytes = WITH_J ? size_t(TILE) * E * sizeof(T) : 0;
    static constexpr size_t o_bytes = size_t(TILE) * n * sizeof(T);
    static constexpr size_t align16(size_t b) { return (b + 15) & ~size_t(15); }
    static constexpr size_t off_in0 = 0;
    static constexpr size_t off_in1 = align16(in_bytes);
    static constexpr size_t off_j = off_in1 + align16(in_bytes);
    static constexpr size_t off_o = off_j + align16(j_bytes);
    static constexpr size_t off_bar = off_o + align16(o_bytes);
    static constexpr size_t total = off_bar + 16;
};

template <class Model, int Q, class T, int TILE, bool WITH_J, class Chunks, int MINB, int ROLL>
__global__ void __launch_bounds__(TILE * Chunks::count, MINB)
knot_kernel(const Model model, const KnotArgs<T> a) {
    constexpr int n = Model::n, NZ = Model::n + Model::m, E = n * NZ;
    constexpr int NTHR = TILE * Chunks::count;
    using S = KnotSmem<Model, TILE, WITH_J, T>;
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
    const bool tma_ok = a.layout == LAYOUT_AOS && ((reinterpret_cast<uintptr_t>(a.Z) | reinterpret_cast<uintptr_t>(a.J) |
                                                    reinterpret_cast<uintptr_t>(a.out)) & 15) == 0;
    auto tile_tma = [&](long long tile) { return tma_ok && (tile + 1) * TILE <= N; };

    if (tid == 0) { mbar_init(bar0, 1); mbar_init(bar0 + 8, 1); fence_mbar_init(); }
    __syncthreads();
    // Programmatic dependent launch: everything above overlaps the tail of the previous kernel in the stream; nothing
    // below (first global access) may start before that kernel's memory is visible.  No-ops without the launch attribute.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");

    long long tile = blockIdx.x;
    if (tile < ntiles && tile_tma(tile) && tid == 0) {
        mbar_expect_tx(bar0, uint32_t(S::in_bytes));
        bulk_load(smem_u32(in_img[0]), a.Z + tile * TILE * NZ, uint32_t(S::in_bytes), bar0);
    }
    uint32_t phase[2] = {0, 0};
    for (int it = 0; tile < ntiles; ++it, tile += gridDim.x) {
        const int s = it & 1;
        const long long k0 = tile * TILE;
        const int cnt = int((N - k0) < TILE ? (N - k0) : TILE);
        const long long nxt = tile + gridDim.x;
        // (1) prefetch the next tile's [x;u] rows (buffer s^1 was last read before the previous iteration's barriers)
        if (tid == 0 && nxt < ntiles && tile_tma(nxt)) {
            mbar_expect_tx(bar0 + 8 * (s ^ 1), uint32_t(S::in_bytes));
            bulk_load(smem_u32(in_img[s ^ 1]), a.Z + nxt * TILE * NZ, uint32_t(S::in_bytes), bar0 + 8 * (s ^ 1));
        }
        // (2) this tile's inputs
        const bool tma = tile_tma(tile);
        if (tma) { mbar_wait(bar0 + 8 * s, phase[s]); phase[s] ^= 1; }
        else { coop_load(in_img[s], a.Z, k0, cnt, NZ, N, a.layout, tid, NTHR); __syncthreads(); }
        // (3) compute in registers
        const T* zrow = in_img[s] + kt * NZ;
        T h = T(0);
        if constexpr (Q != Q_CONTINUOUS) h = T(a.dt ? (kt < cnt ? a.dt[k0 + kt] : 0.0) : a.dt0);
        (void)cnt;
        // (4) evaluate; inside, all threads meet at images_free_barrier() before touching the output images
        //     (rows past the ragged end compute on stale smem and are never copied out)
        dispatch_role<Model, Q, T, WITH_J, Chunks, NTHR, ROLL>(role, model, zrow, h, j_img + kt * E, want_o ? o_img + kt * n : nullptr, tid);
        // (5) publish
        if (tma) {
            fence_proxy_async();
            __syncthreads();
            if (tid == 0) {
                if (want_j) bulk_store(a.J + k0 * E, smem_u32(j_img), uint32_t(S::j_bytes));
                if (want_o) bulk_store(a.out + k0 * n, smem_u32(o_img), uint32_t(S::o_bytes));
                bulk_commit();
            }
        } else {
            __syncthreads();
            if (want_j) coop_store(j_img, a.J, k0, cnt, E, N, a.layout, tid, NTHR);
            if (want_o) coop_store(o_img, a.out, k0, cnt, n, N, a.layout, tid, NTHR);
        }
    }
    if (tid == 0) bulk_wait0();
}

}  // namespace rdb
