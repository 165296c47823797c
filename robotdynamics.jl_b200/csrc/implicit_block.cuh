// implicit_block.cuh — ImplicitMidpoint (reference: src/integration.jl:422-463 Newton loop, :524-543 implicit-function-theorem Jacobian,
// :620-694 residual) for models whose continuous Jacobian A = df/dx is BLOCK LOWER-TRIANGULAR under an ordering of the state that the
// model declares (mp_nblocks / mp_size / mp_idx):
//
//   RigidBody        [w | att | v | r]   w' = f(w,u);  att' = f(att,w);  v' = f(att,v,w,u);  r' = f(att,v)      (src/rigidbody.jl:213-236)
//   Cartpole         [theta,w | v | p]                                                                          (test/cartpole_model.jl:11-30)
//   DoubleIntegrator [v | p]
//
// M = h/2 A - I then has the same block structure with diagonal blocks of at most 4 x 4, so  M \ b  is a block forward substitution: one
// small pivoted LU per diagonal block (in registers, fully unrolled; blocks on which A has no entry are -I and cost nothing) and one FMA per
// structural non-zero of A below the diagonal blocks.  The non-zeros are not a hand-written pattern: f is evaluated ONCE per Newton
// iteration in forward mode with all columns seeded (sdual.cuh), the masks of the resulting sparse duals are the pattern — checked at
// compile time against the declared ordering (mp_block_ok; a model that does not conform falls back to the dense group kernel,
// kernels.cuh) — and every product below is guarded by `if constexpr` on those masks.  Right-hand-side columns of the Jacobian solve
// J = -M \ [I + h/2 A, h B] are structurally sparse as well: blocks of the solution that are zero for a column are skipped at compile time.
//
// One thread per knot, one warp per CTA (persistent): the continuous Jacobian (~100 partials for a rigid body) and the LU factors live in
// registers; the columns of J are assembled in a per-warp shared-memory image (odd pitch: conflict-free) and leave by coalesced stores.
// Same Newton arithmetic as the dense kernels to rounding (same iterates: the linear systems are solved exactly up to rounding either way).
#pragma once
#include "kernels.cuh"

namespace rdb {

template <int B, int E, class Fn>
__device__ __forceinline__ void sfor(Fn&& fn) {
    if constexpr (B < E) { fn(rstd::integral_constant<int, B>{}); sfor<B + 1, E>(fn); }
}

// mask of row I of the forward-mode result F (a Vec of SD / plain elements)
template <class F, int I>
__host__ __device__ constexpr mask_t mp_rmask() { return mask_of<rstd::remove_cv_t<rstd::remove_reference_t<decltype(get<I>(rstd::declval<const F&>()))>>>::value; }
template <class F, size_t... Is>
__host__ __device__ constexpr mask_t mp_rmask_at(int i, rstd::index_sequence<Is...>) { mask_t r = 0; ((int(Is) == i ? (r = mp_rmask<F, int(Is)>(), 0) : 0), ...); return r; }

template <class Model, class F>
struct MidpointStructure {
    static constexpr int n = Model::n, m = Model::m, NZ = n + m, NB = Model::mp_nblocks;
    using Seq = rstd::make_index_sequence<size_t(n)>;
    __host__ __device__ static constexpr mask_t rmask(int i) { return mp_rmask_at<F>(i, Seq{}); }
    __host__ __device__ static constexpr int block_of(int col) {
        for (int b = 0; b < NB; ++b) for (int k = 0; k < Model::mp_size(b); ++k) if (Model::mp_idx(b, k) == col) return b;
        return -1;
    }
    __host__ __device__ static constexpr mask_t block_cols(int b) { mask_t r = 0; for (int k = 0; k < Model::mp_size(b); ++k) r |= mask_t(1) << Model::mp_idx(b, k); return r; }
    // every state index in exactly one block, every row free of columns of LATER blocks
    __host__ __device__ static constexpr bool ok() {
        mask_t seen = 0;
        for (int b = 0; b < NB; ++b) {
            if (Model::mp_size(b) < 1 || Model::mp_size(b) > 4) return false;
            if (seen & block_cols(b)) return false;
            seen |= block_cols(b);
        }
        if (seen != ((mask_t(1) << n) - 1u)) return false;
        for (int b = 0; b < NB; ++b) {
            mask_t later = 0;
            for (int c = b + 1; c < NB; ++c) later |= block_cols(c);
            for (int k = 0; k < Model::mp_size(b); ++k) if (rmask(Model::mp_idx(b, k)) & later) return false;
        }
        return true;
    }
    // does A have any entry inside diagonal block b?  (otherwise that block of M is -I)
    __host__ __device__ static constexpr bool diag_nz(int b) { for (int k = 0; k < Model::mp_size(b); ++k) if (rmask(Model::mp_idx(b, k)) & block_cols(b)) return true; return false; }
    // does block b (rows) depend on block c (columns)?
    __host__ __device__ static constexpr bool dep(int b, int c) { for (int k = 0; k < Model::mp_size(b); ++k) if (rmask(Model::mp_idx(b, k)) & block_cols(c)) return true; return false; }
    // blocks in which the right-hand-side column c of [I + h/2 A, h B] can be non-zero / in which M \ that column can be non-zero
    __host__ __device__ static constexpr unsigned rhs_blocks(int c) {
        unsigned r = 0;
        for (int b = 0; b < NB; ++b) for (int k = 0; k < Model::mp_size(b); ++k) {
            const int i = Model::mp_idx(b, k);
            if (i == c || (rmask(i) >> c) & 1u) r |= 1u << b;
        }
        return r;
    }
    __host__ __device__ static constexpr unsigned sol_blocks(int c) {
        const unsigned rb = rhs_blocks(c);
        unsigned y = 0;
        for (int b = 0; b < NB; ++b) {
            bool nz = (rb >> b) & 1u;
            for (int e = 0; e < b && !nz; ++e) if (((y >> e) & 1u) && dep(b, e)) nz = true;
            if (nz) y |= 1u << b;
        }
        return y;
    }
};

// LU with partial pivoting of a block of at most 4 x 4, fully unrolled, in registers (run-time pivot indices act through selects)
template <class T>
struct SmallLU {
    T a[4][4];
    T dinv[4];
    int piv[4];
    template <int N>
    __device__ __forceinline__ void factor() {
#pragma unroll
        for (int k = 0; k < N; ++k) {
            int p = k;
            T best = fabs(a[k][k]);
#pragma unroll
            for (int i = k + 1; i < N; ++i) { const T v = fabs(a[i][k]); if (v > best) { best = v; p = i; } }
            piv[k] = p;
#pragma unroll
            for (int j = 0; j < N; ++j) {
                const T wk = a[k][j]; T wp = wk;
#pragma unroll
                for (int i = k + 1; i < N; ++i) { wp = sel_eq(p, i, a[i][j], wp); a[i][j] = sel_eq(p, i, wk, a[i][j]); }
                a[k][j] = wp;
            }
            dinv[k] = T(1) / a[k][k];
#pragma unroll
            for (int i = k + 1; i < N; ++i) {
                const T l = a[i][k] * dinv[k];
                a[i][k] = l;
#pragma unroll
                for (int j = k + 1; j < N; ++j) a[i][j] = fma_(-l, a[k][j], a[i][j]);
            }
        }
    }
    template <int N>
    __device__ __forceinline__ void solve(T (&b)[4]) const {
#pragma unroll
        for (int k = 0; k < N; ++k) {
            const T bk = b[k]; T bp = bk;
#pragma unroll
            for (int i = k + 1; i < N; ++i) { bp = sel_eq(piv[k], i, b[i], bp); b[i] = sel_eq(piv[k], i, bk, b[i]); }
            b[k] = bp;
#pragma unroll
            for (int i = k + 1; i < N; ++i) b[i] = fma_(-a[i][k], b[k], b[i]);
        }
#pragma unroll
        for (int i = N - 1; i >= 0; --i) {
            T s = b[i];
#pragma unroll
            for (int c = i + 1; c < N; ++c) s = fma_(-a[i][c], b[c], s);
            b[i] = s * dinv[i];
        }
    }
};

template <class Model, class T, class F>
struct MidpointBlockSolver {
    using S = MidpointStructure<Model, F>;
    static constexpr int n = S::n, NB = S::NB;
    SmallLU<T> lu[NB];
    // diagonal blocks of M = hh A - I from the partials of fd, factored
    __device__ __forceinline__ void factor(const F& fd, T hh) {
        sfor<0, NB>([&](auto bc) {
            constexpr int B = decltype(bc)::value, SZ = Model::mp_size(B);
            if constexpr (S::diag_nz(B)) {
                sfor<0, SZ>([&](auto kc) {
                    constexpr int K = decltype(kc)::value, I = Model::mp_idx(B, K);
                    sfor<0, SZ>([&](auto lc) {
                        constexpr int L = decltype(lc)::value, Jc = Model::mp_idx(B, L);
                        lu[B].a[K][L] = hh * partial<Jc>(get<I>(fd)) - (K == L ? T(1) : T(0));
                    });
                });
                lu[B].template factor<SZ>();
            }
        });
    }
    // y <- M \ y for a vector whose blocks outside YNZ are structurally zero (and stay so); yt = hh * y
    template <unsigned YNZ = ~0u>
    __device__ __forceinline__ void solve(const F& fd, T hh, T (&y)[n]) const {
        T yt[n];
        sfor<0, NB>([&](auto bc) {
            constexpr int B = decltype(bc)::value, SZ = Model::mp_size(B);
            if constexpr ((YNZ >> B) & 1u) {
                T rb[4];
                sfor<0, SZ>([&](auto kc) {
                    constexpr int K = decltype(kc)::value, I = Model::mp_idx(B, K);
                    T s = y[I];
                    sfor<0, n>([&](auto jc) {
                        constexpr int Jc = decltype(jc)::value;
                        constexpr int BJ = S::block_of(Jc);
                        if constexpr (BJ < B && ((YNZ >> BJ) & 1u) && ((S::rmask(I) >> Jc) & 1u)) s = fma_(-partial<Jc>(get<I>(fd)), yt[Jc], s);
                    });
                    rb[K] = s;
                });
                if constexpr (S::diag_nz(B)) lu[B].template solve<SZ>(rb);
                else sfor<0, SZ>([&](auto kc) { rb[decltype(kc)::value] = -rb[decltype(kc)::value]; });
                sfor<0, SZ>([&](auto kc) {
                    constexpr int K = decltype(kc)::value, I = Model::mp_idx(B, K);
                    y[I] = rb[K]; yt[I] = hh * rb[K];
                });
            }
        });
    }
};

// every column of [x;u] seeded (host + device: the result TYPE is also needed by the host-side dispatch)
template <class T, mask_t ALL, size_t... Is>
RDB_HD auto mp_seed_all(const T* z, rstd::index_sequence<Is...>) { return vec(seed<T, int(Is), ALL>(z[Is])...); }

template <class Model, class T>
struct MidpointF {
    static constexpr int n = Model::n, m = Model::m, NZ = n + m;
    static constexpr mask_t ALL = (NZ >= 32) ? ~mask_t(0) : ((mask_t(1) << NZ) - 1u);
    using Z = decltype(mp_seed_all<T, ALL>(static_cast<const T*>(nullptr), rstd::make_index_sequence<size_t(NZ)>{}));
    using type = decltype(feval<T>(rstd::declval<const Model&>(), slice<0, n>(rstd::declval<const Z&>()), slice<n, m>(rstd::declval<const Z&>()), T(0)));
};
template <class Model, class = void> struct has_mp_blocks : rstd::false_type {};
template <class Model> struct has_mp_blocks<Model, rstd::void_t<decltype(Model::mp_nblocks)>> : rstd::true_type {};
template <class Model, class T, bool = has_mp_blocks<Model>::value> struct mp_block_check { static constexpr bool value = false; };
template <class Model, class T> struct mp_block_check<Model, T, true> {
    static constexpr bool value = MidpointStructure<Model, typename MidpointF<Model, T>::type>::ok();
};
template <class Model, class T> constexpr bool mp_block_ok() { return mp_block_check<Model, T>::value; }

#ifndef RDB_IMPLICIT_IMG_BUDGET
#define RDB_IMPLICIT_IMG_BUDGET 0      // tuning experiments: bytes of per-warp image (0 = default rule below)
#endif
// CG: Jacobian columns staged per flush of the per-warp image (all n+m when the image fits; fewer for the fp64 rigid bodies)
template <class Model, class T, bool WITH_J, int CG>
__global__ void __launch_bounds__(32) implicit_midpoint_block_kernel(const Model model, const KnotArgs<T> a) {
    constexpr int n = Model::n, m = Model::m, NZ = n + m, E = n * NZ;
    using MF = MidpointF<Model, T>;
    using F = typename MF::type;
    using S = MidpointStructure<Model, F>;
    static_assert(S::ok(), "df/dx of this model is not block lower-triangular under its declared ordering (mp_nblocks / mp_size / mp_idx)");
    constexpr mask_t ALL = MF::ALL;
    constexpr int NG = (NZ + CG - 1) / CG;                       // flushes per tile
    constexpr int GW = n * CG, PJ = (GW % 2) ? GW : GW + 1;      // image row: one knot's columns of a group, odd pitch
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* in_img = reinterpret_cast<T*>(smem_raw);                  // [32][NZ]
    T* o_img = in_img + 32 * NZ;                                 // [32][n]
    T* j_img = o_img + 32 * n;                                   // [32][PJ]
    const int lane = threadIdx.x;
    const long long N = a.N, ntiles = (N + 31) / 32;
    const T tol = sizeof(T) == 8 ? T(1e-12) : T(1e-5);
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long k0 = tile * 32;
        const int cnt = int((N - k0) < 32 ? (N - k0) : 32);
        __syncwarp();
        { const T* src = a.Z + k0 * NZ; for (int i = lane; i < cnt * NZ; i += 32) in_img[i] = src[i]; }
        __syncwarp();
        const long long k = k0 + (lane < cnt ? lane : cnt - 1);   // lanes past the ragged end shadow the last knot
        T z[NZ], zm[NZ], x2[n], r[n];
#pragma unroll
        for (int i = 0; i < NZ; ++i) { z[i] = in_img[(lane < cnt ? lane : cnt - 1) * NZ + i]; zm[i] = z[i]; }
#pragma unroll
        for (int i = 0; i < n; ++i) x2[i] = z[i];
        const T h = T(a.dt ? a.dt[k] : a.dt0), hh = T(0.5) * h;
        T tm = T(0);
        if constexpr (uses_time<Model>::value) tm = T(a.t ? a.t[k] : 0.0) + hh;
        F fd;
        MidpointBlockSolver<Model, T, F> solver;
#pragma unroll 1
        for (int iter = 0; iter < 10; ++iter) {
#pragma unroll
            for (int i = 0; i < n; ++i) zm[i] = (z[i] + x2[i]) * T(0.5);
            model.reset();
            {
                auto zz = mp_seed_all<T, ALL>(zm, rstd::make_index_sequence<size_t(NZ)>{});
                fd = feval<T>(model, slice<0, n>(zz), slice<n, m>(zz), tm);
            }
            put_vals(fd, r, rstd::make_index_sequence<size_t(n)>{});
            T nrm = T(0);
#pragma unroll
            for (int i = 0; i < n; ++i) { r[i] = z[i] + h * r[i] - x2[i]; nrm = fma_(r[i], r[i], nrm); }
            if (sqrt(nrm) < tol) break;
            solver.factor(fd, hh);
            solver.template solve<>(fd, hh, r);
#pragma unroll
            for (int i = 0; i < n; ++i) x2[i] -= r[i];
        }
        if (a.out) {
#pragma unroll
            for (int i = 0; i < n; ++i) o_img[lane * n + i] = x2[i];
            __syncwarp();
            T* dst = a.out + k0 * n;
            for (int i = lane; i < cnt * n; i += 32) dst[i] = o_img[i];
        }
        if constexpr (WITH_J) {
            if (a.J) {
                solver.factor(fd, hh);          // Jacobian from the last evaluated iterate (src/integration.jl:524-543)
                sfor<0, NG>([&](auto gc) {
                    constexpr int G = decltype(gc)::value, C0 = G * CG, C1 = (C0 + CG < NZ) ? C0 + CG : NZ;
                    if constexpr (G > 0) __syncwarp();           // the previous group's copy-out has read the image
                    sfor<C0, C1>([&](auto cc) {
                        constexpr int C = decltype(cc)::value;
                        constexpr unsigned YB = S::sol_blocks(C);
                        T y[n];
                        sfor<0, n>([&](auto ic) {
                            constexpr int I = decltype(ic)::value;
                            if constexpr ((S::rmask(I) >> C) & 1u) {
                                const T ai = (C < n ? hh : h) * partial<C>(get<I>(fd));
                                y[I] = (I == C) ? ai + T(1) : ai;
                            } else y[I] = (I == C) ? T(1) : T(0);
                        });
                        solver.template solve<YB>(fd, hh, y);
                        T* col = j_img + lane * PJ + n * (C - C0);
                        sfor<0, n>([&](auto ic) {
                            constexpr int I = decltype(ic)::value;
                            if constexpr ((YB >> S::block_of(I)) & 1u) col[I] = -y[I]; else col[I] = T(0);
                        });
                    });
                    __syncwarp();
                    constexpr int W = n * (C1 - C0);
                    for (int rr = 0; rr < cnt; ++rr) {           // one knot row per step: unit-stride smem reads, coalesced global writes
                        T* dst = a.J + (k0 + rr) * (long long)E + n * C0;
                        for (int e = lane; e < W; e += 32) dst[e] = j_img[rr * PJ + e];
                    }
                });
            }
        }
    }
}

// Columns staged per flush, as plain integers (shared by the launcher below and by custom.cu, which only knows a user model's dimensions
// at run time).  The kernel is latency-bound (ncu: 29 % issue utilisation at 6 warps per SM with the whole 28 KB image staged), so the
// image is kept small enough for the REGISTER-limited number of warps per SM — 16 at <= 128 registers (fp32), 8 at <= 255 (fp64) — at
// the price of two or three flushes per tile.
__host__ __device__ constexpr size_t mp_img_bytes(int n, int cg, int es) { return size_t(32) * size_t(n * cg + 1) * size_t(es); }
__host__ __device__ constexpr int mp_cols_per_flush(int n, int m, int es, bool with_j) {
    if (!with_j) return 1;
    const size_t budget = RDB_IMPLICIT_IMG_BUDGET > 0 ? size_t(RDB_IMPLICIT_IMG_BUDGET) : (es == 4 ? 10 * 1024 + 512 : 22 * 1024);
    const int NZ = n + m;
    for (int g = 1; g < NZ; ++g) { const int cg = (NZ + g - 1) / g; if (mp_img_bytes(n, cg, es) <= budget) return cg; }
    return 1;
}
__host__ __device__ constexpr size_t mp_smem_bytes(int n, int m, int es, bool with_j) {
    return size_t(32) * size_t(2 * n + m) * size_t(es) + (with_j ? mp_img_bytes(n, mp_cols_per_flush(n, m, es, with_j), es) : 0) + 16;
}

#ifndef __CUDACC_RTC__
template <class Model, class T, bool WITH_J>
struct MidpointBlockLaunch {
    static constexpr int n = Model::n, m = Model::m, NZ = n + m;
    static constexpr int CG = mp_cols_per_flush(n, m, int(sizeof(T)), WITH_J);
    static constexpr size_t smem = mp_smem_bytes(n, m, int(sizeof(T)), WITH_J);
    static int run(const Model& model, const KnotArgs<T>& a, int sm_count, cudaStream_t st) {
        auto kern = implicit_midpoint_block_kernel<Model, T, WITH_J, CG>;
        static std::atomic<int> occ_cache[64];
        int dev = 0;
        cudaGetDevice(&dev);
        int occ = occ_cache[dev & 63].load(std::memory_order_acquire);
        if (occ == 0) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
            if (e != cudaSuccess) return int(e);
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32, smem);
            if (e != cudaSuccess) return int(e);
            if (occ < 1) occ = 1;
            occ_cache[dev & 63].store(occ, std::memory_order_release);
        }
        const long long ntiles = (a.N + 31) / 32, cap = (long long)sm_count * occ;
        kern<<<unsigned(ntiles < cap ? ntiles : cap), 32, smem, st>>>(model, a);
        return int(cudaGetLastError());
    }
};
#endif   // !__CUDACC_RTC__

}  // namespace rdb
