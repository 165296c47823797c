// lie.cu — LieState error-state maps for RigidBody{R} states (K4/K5 of SURVEY.md §2), one thread per OUTPUT
// element so every store is coalesced; the rotation block entries are recomputed from the knot's attitude.
//
//   errstate_jacobian!   reference: src/liestate.jl:262-298 (block-diag I / Rotations.∇differential), Euclidean
//                        fall-back src/statevectortype.jl:149-155.  Unlike the reference the whole matrix is written.
//   ∇errstate_jacobian!  reference: src/liestate.jl:300-320 (Rotations.∇²differential on the rotation block)
//   state_diff           reference: src/liestate.jl:210-260 (vector parts x - x0, rotations Cayley error of q0\q)
//
// Here the attitude IS normalised (default R(w,x,y,z) constructor), unlike inside dynamics.
#include <atomic>
#include "lie.h"
#include "kernels.cuh"   // mbarrier / bulk-copy PTX helpers

namespace rdb {

template <class T> __device__ __forceinline__ T rsq(T a);
template <> __device__ __forceinline__ float rsq<float>(float a) { return 1.0f / sqrtf(a); }
template <> __device__ __forceinline__ double rsq<double>(double a) { return 1.0 / sqrt(a); }

// unit quaternion of an attitude parameterisation
template <class T>
__device__ __forceinline__ void unit_quat(int rot, const T* p, T* q) {
    if (rot == ROT_QUAT) {
        const T s = rsq(p[0] * p[0] + p[1] * p[1] + p[2] * p[2] + p[3] * p[3]);
        q[0] = p[0] * s; q[1] = p[1] * s; q[2] = p[2] * s; q[3] = p[3] * s;
    } else if (rot == ROT_MRP) {
        const T n2 = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
        const T i1 = T(1) / (T(1) + n2), M = T(2) * i1;
        q[0] = (T(1) - n2) * i1; q[1] = M * p[0]; q[2] = M * p[1]; q[3] = M * p[2];
    } else {
        const T M = rsq(T(1) + p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
        q[0] = M; q[1] = M * p[0]; q[2] = M * p[1]; q[3] = M * p[2];
    }
}

// entry (i,j) of Rotations.∇differential(R):  quat 4x3 L(q)H,  MRP (1-|p|^2) I + 2(skew(p) + p p'),  RP I + skew(g) + g g'
template <class T>
__device__ __forceinline__ T grad_differential(int rot, const T* p, int i, int j) {
    if (rot == ROT_QUAT) {
        T q[4]; unit_quat(rot, p, q);
        const T w = q[0], x = q[1], y = q[2], z = q[3];
        const T G[4][3] = {{-x, -y, -z}, {w, -z, y}, {z, w, -x}, {-y, x, w}};
        return G[i][j];
    }
    const T sk[3][3] = {{T(0), -p[2], p[1]}, {p[2], T(0), -p[0]}, {-p[1], p[0], T(0)}};
    const T I = (i == j) ? T(1) : T(0);
    if (rot == ROT_MRP) {
        const T n2 = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
        return (T(1) - n2) * I + T(2) * (sk[i][j] + p[i] * p[j]);
    }
    return I + sk[i][j] + p[i] * p[j];
}

// ---- index algebra of LieState{R,P} (reference: src/liestate.jl:127-131: rot_inds / vec_inds) -------------------------------------------
// where state index i (or error-state index, ERR = true) falls: block b, offset o within it, and whether the block is a rotation
struct LiePos { int rot; int b; int o; int start; };
template <bool ERR>
__device__ __forceinline__ LiePos lie_locate(const LieParts& parts, int np, int i) {
    const int w = ERR ? 3 : np;
    int s = 0;
    for (int k = 0; k < parts.nv; ++k) {
        if (i < s + parts.P[k]) return {0, k, i - s, s};
        s += parts.P[k];
        if (k + 1 < parts.nv) {
            if (i < s + w) return {1, k, i - s, s};
            s += w;
        }
    }
    return {0, parts.nv, 0, s};
}
__device__ __forceinline__ int lie_rot_start(const LieParts& parts, int w, int b) {      // first index of rotation b (width w per rotation)
    int s = 0;
    for (int k = 0; k <= b; ++k) s += parts.P[k];
    return s + b * w;
}

// errstate_jacobian! for ANY LieState (and Euclidean states, nv = 1): block diagonal, I on the vector blocks, Rotations.∇differential on
// every rotation (reference: src/liestate.jl:262-298).  One thread per output element, coalesced stores.
template <class T>
__global__ void __launch_bounds__(256) errstate_jacobian_kernel(int rot, LieParts parts, int n, int ne, long long N, const T* __restrict__ X, int ldx,
                                                                T* __restrict__ G) {
    const long long total = N * (long long)n * ne;
    const int per = n * ne;
    const int np = (rot == ROT_QUAT) ? 4 : 3;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long k = idx / per;
        const int e = int(idx - k * per), j = e / n, i = e - j * n;
        const LiePos pi = lie_locate<false>(parts, np, i), pj = lie_locate<true>(parts, np, j);
        T v = T(0);
        if (pi.rot == pj.rot && pi.b == pj.b) {
            if (!pi.rot) v = (pi.o == pj.o) ? T(1) : T(0);
            else v = grad_differential(rot, X + k * ldx + pi.start, pi.o, pj.o);
        }
        G[idx] = v;
    }
}

// Rigid bodies (the hot one: G is n x 12 per knot, 624 B in fp32, and only its 4x3 / 3x3 attitude block depends on the state).
// The output of a tile of TILE knots is one contiguous byte range, so the kernel keeps two shared-memory IMAGES of it: the
// structural 0/1 pattern is written once per CTA, each tile only overwrites its knots' attitude-block entries (12 or 9 stores per
// knot) and one TMA bulk store ships the image; the two images alternate so a store drains while the next tile is prepared.
// Pure store stream: HBM-write bound.
template <class T, int NP, int TILE>
__global__ void __launch_bounds__(TILE) errstate_jacobian_tma_kernel(int rot, long long N, const T* __restrict__ X, int ldx, T* __restrict__ G, int stream_out) {
    constexpr int n = 9 + NP, ne = 12, PER = n * ne;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* img[2] = {reinterpret_cast<T*>(smem_raw), reinterpret_cast<T*>(smem_raw) + TILE * PER};
    const long long ntiles = (N + TILE - 1) / TILE;
    // structural 0/1 pattern: every thread writes its own knot's row with compile-time indices (no div/mod; the round-1 form, a
    // strided loop over both images with two integer divisions per element, cost ~30 us per launch whatever N was); the second
    // image only if this CTA has a second tile
    const int nimg = (blockIdx.x + (long long)gridDim.x < ntiles) ? 2 : 1;
    for (int b = 0; b < nimg; ++b) {
        T* row = img[b] + threadIdx.x * PER;
#pragma unroll
        for (int j = 0; j < ne; ++j)
#pragma unroll
            for (int i = 0; i < n; ++i) row[i + n * j] = (i < 3) ? T(j == i) : (i < 3 + NP) ? T(0) : T(j == i - NP + 3);
    }
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    int it = 0;
    const uint64_t pol_out = l2_policy_evict_first();
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const long long k0 = tile * TILE;
        const int cnt = int((N - k0) < TILE ? (N - k0) : TILE);
        T* im = img[it & 1];
        if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the store that last used this image is done reading
        __syncthreads();
        if (threadIdx.x < cnt) {
            const T* p = X + (k0 + threadIdx.x) * ldx + 3;
            T pp[4] = {p[0], p[1], p[2], NP == 4 ? p[NP - 1] : T(0)};
            T* row = im + threadIdx.x * PER;
#pragma unroll
            for (int i = 0; i < NP; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) row[(3 + i) + n * (3 + j)] = grad_differential(rot, pp, i, j);
        }
        fence_proxy_async();
        __syncthreads();
        if (threadIdx.x == 0) {      // outputs larger than the L2 keeps for a consumer leave with the evict_first hint (kernels.cuh: knot_stream_out)
            if (stream_out) bulk_store(G + k0 * PER, smem_u32(im), uint32_t(cnt * PER * sizeof(T)), pol_out);
            else bulk_store(G + k0 * PER, smem_u32(im), uint32_t(cnt * PER * sizeof(T)));
            bulk_commit();
        }
    }
    if (threadIdx.x == 0) bulk_wait_read0();
}

// ∇²differential(R, b) = ∇²composition1(R, I, b): the Hessians of the components of R ∘ δ with respect to δ at 0, contracted with b (δ
// parameterised like R) — quat: -(q.b) I3;  MRP: -2 (1+|p|²)(p.b) I + 2 (a p' + p a') + 8 (p.b) p p',  a = (1-|p|²) b - 2 p x b;
// RP: a g' + g a' + 2 (g.b) g g',  a = b - g x b.  One 3x3 block per rotation on the diagonal, zeros elsewhere (reference:
// src/liestate.jl:300-320).  (Round 2: the MRP / RP rows were d/dδ [∇differential(p∘δ)' b] before — not the package's definition.)
template <class T>
__global__ void __launch_bounds__(256) grad_errstate_jacobian_kernel(int rot, LieParts parts, int n, int ne, long long N, const T* __restrict__ X, int ldx,
                                                                     const T* __restrict__ B, int ldb, T* __restrict__ H) {
    const long long total = N * (long long)ne * ne;
    const int per = ne * ne;
    const int np = (rot == ROT_QUAT) ? 4 : 3;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long k = idx / per;
        const int e = int(idx - k * per), j = e / ne, i = e - j * ne;
        const LiePos pi = lie_locate<true>(parts, np, i), pj = lie_locate<true>(parts, np, j);
        T v = T(0);
        if (pi.rot && pj.rot && pi.b == pj.b) {
            const int xs = lie_rot_start(parts, np, pi.b);
            const T* p = X + k * ldx + xs;
            const T* b = B + k * ldb + xs;
            const int a = pi.o, c = pj.o;
            if (rot == ROT_QUAT) {
                if (a == c) { T q[4]; unit_quat(rot, p, q); v = -(q[0] * b[0] + q[1] * b[1] + q[2] * b[2] + q[3] * b[3]); }
            } else {
                const T pb = p[0] * b[0] + p[1] * b[1] + p[2] * b[2], n2 = p[0] * p[0] + p[1] * p[1] + p[2] * p[2];
                const T pxb[3] = {p[1] * b[2] - p[2] * b[1], p[2] * b[0] - p[0] * b[2], p[0] * b[1] - p[1] * b[0]};
                const bool mrp = rot == ROT_MRP;
                const T aa = mrp ? (T(1) - n2) * b[a] - T(2) * pxb[a] : b[a] - pxb[a];
                const T ac = mrp ? (T(1) - n2) * b[c] - T(2) * pxb[c] : b[c] - pxb[c];
                const T I = (a == c) ? T(1) : T(0);
                v = mrp ? T(-2) * (T(1) + n2) * pb * I + T(2) * (aa * p[c] + p[a] * ac) + T(8) * pb * p[a] * p[c]
                        : aa * p[c] + p[a] * ac + T(2) * pb * p[a] * p[c];
            }
        }
        H[idx] = v;
    }
}

// state_diff for ANY LieState: vector blocks x - x0, every rotation the Cayley error of q0 \ q (reference: src/liestate.jl:210-260)
template <class T>
__global__ void __launch_bounds__(256) state_diff_kernel(int rot, LieParts parts, int n, int ne, long long N, const T* __restrict__ X, int ldx,
                                                         const T* __restrict__ X0, int ldx0, T* __restrict__ dX) {
    const long long total = N * (long long)ne;
    const int np = (rot == ROT_QUAT) ? 4 : 3;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long k = idx / ne;
        const int i = int(idx - k * ne);
        const T* x = X + k * ldx;
        const T* x0 = X0 + k * ldx0;
        const LiePos pe = lie_locate<true>(parts, np, i);
        T v;
        if (!pe.rot) {
            const int xi = pe.start + pe.b * (np - 3) + pe.o;          // every rotation before this block is np wide in x, 3 wide in dx
            v = x[xi] - x0[xi];
        } else {
            const int xs = lie_rot_start(parts, np, pe.b);
            T q[4], q0[4];
            unit_quat(rot, x + xs, q);
            unit_quat(rot, x0 + xs, q0);
            // e = conj(q0) (x) q ;  Cayley map: vec(e) / scalar(e)
            const T c0 = q0[0], c1 = -q0[1], c2 = -q0[2], c3 = -q0[3];
            const T e0 = c0 * q[0] - c1 * q[1] - c2 * q[2] - c3 * q[3];
            T ev;
            if (pe.o == 0) ev = c0 * q[1] + c1 * q[0] + c2 * q[3] - c3 * q[2];
            else if (pe.o == 1) ev = c0 * q[2] - c1 * q[3] + c2 * q[0] + c3 * q[1];
            else ev = c0 * q[3] + c1 * q[2] - c2 * q[1] + c3 * q[0];
            v = ev / e0;
        }
        dX[idx] = v;
    }
}

// Error-state projection for ANY LieState:  Jbar = G(x+)' [A B] blkdiag(G(x), I)  (what Altro / TrajectoryOptimization build from
// jacobian! and errstate_jacobian!, src/liestate.jl:262-298, src/functionbase.jl:135).  G is block diagonal, so entry (a, c) of Jbar only
// sums over the rows of a's block and the columns of c's block (1 term for vector blocks, np x np for two rotations).
template <class T>
__global__ void __launch_bounds__(256) project_error_jacobian_kernel(int rot, LieParts parts, int n, int m, int ne, long long N, const T* __restrict__ Z, int ldz,
                                                                     const T* __restrict__ Xn, const T* __restrict__ J, T* __restrict__ Jbar) {
    const int per = ne * (ne + m);
    const long long total = N * (long long)per;
    const int np = (rot == ROT_QUAT) ? 4 : 3;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long k = idx / per;
        const int e = int(idx - k * per), c = e / ne, a = e - c * ne;
        const T* Jk = J + k * (long long)(n * (n + m));
        const LiePos pa = lie_locate<true>(parts, np, a);
        // rows of J that meet error row a, with weights G(x+)[i][a]
        int i0, ni; T wr[4];
        if (!pa.rot) { i0 = pa.start + pa.b * (np - 3) + pa.o; ni = 1; wr[0] = T(1); }
        else {
            i0 = lie_rot_start(parts, np, pa.b); ni = np;
            for (int i = 0; i < np; ++i) wr[i] = grad_differential(rot, Xn + k * n + i0, i, pa.o);
        }
        T v = T(0);
        if (c >= ne) {                                                   // control column: Bbar = G(x+)' B
            for (int i = 0; i < ni; ++i) v += wr[i] * Jk[(i0 + i) + n * (n + (c - ne))];
        } else {
            const LiePos pc = lie_locate<true>(parts, np, c);
            if (!pc.rot) {
                const int j = pc.start + pc.b * (np - 3) + pc.o;
                for (int i = 0; i < ni; ++i) v += wr[i] * Jk[(i0 + i) + n * j];
            } else {
                const int j0 = lie_rot_start(parts, np, pc.b);
                for (int jj = 0; jj < np; ++jj) {
                    const T g = grad_differential(rot, Z + k * ldz + j0, jj, pc.o);
                    T col = T(0);
                    for (int i = 0; i < ni; ++i) col += wr[i] * Jk[(i0 + i) + n * (j0 + jj)];
                    v += col * g;
                }
            }
        }
        Jbar[idx] = v;
    }
}

// Tiled state_diff for rigid bodies: one thread per knot computes all 12 error components (one quaternion product instead of
// three), the tile leaves through shared memory with unit-stride stores.
template <class T, int NP, int TILE>
__global__ void __launch_bounds__(TILE) state_diff_tile_kernel(int rot, long long N, const T* __restrict__ X, int ldx, const T* __restrict__ X0, int ldx0,
                                                                T* __restrict__ dX) {
    constexpr int ne = 12;
    __shared__ T img[TILE][ne + 1];
    const long long ntiles = (N + TILE - 1) / TILE;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long k0 = tile * TILE;
        const int cnt = int((N - k0) < TILE ? (N - k0) : TILE);
        __syncthreads();
        if (threadIdx.x < cnt) {
            const T* x = X + (k0 + threadIdx.x) * ldx;
            const T* x0 = X0 + (k0 + threadIdx.x) * ldx0;
            T* o = img[threadIdx.x];
#pragma unroll
            for (int i = 0; i < 3; ++i) o[i] = x[i] - x0[i];
#pragma unroll
            for (int i = 0; i < 6; ++i) o[6 + i] = x[3 + NP + i] - x0[3 + NP + i];
            T q[4], q0[4];
            unit_quat(rot, x + 3, q);
            unit_quat(rot, x0 + 3, q0);
            const T c0 = q0[0], c1 = -q0[1], c2 = -q0[2], c3 = -q0[3];     // conj(q0) (x) q, then the Cayley map vec / scalar
            const T ie0 = T(1) / (c0 * q[0] - c1 * q[1] - c2 * q[2] - c3 * q[3]);
            o[3] = (c0 * q[1] + c1 * q[0] + c2 * q[3] - c3 * q[2]) * ie0;
            o[4] = (c0 * q[2] - c1 * q[3] + c2 * q[0] + c3 * q[1]) * ie0;
            o[5] = (c0 * q[3] + c1 * q[2] - c2 * q[1] + c3 * q[0]) * ie0;
        }
        __syncthreads();
        T* out = dX + k0 * ne;
        for (int e = threadIdx.x; e < cnt * ne; e += TILE) { const int kt = e / ne; out[e] = img[kt][e - kt * ne]; }
    }
}

static unsigned grid_for(long long total, int sm_count) {
    long long g = (total + 255) / 256;
    const long long cap = (long long)sm_count * 8;
    if (g > cap) g = cap;
    return unsigned(g < 1 ? 1 : g);
}

int lie_errstate_jacobian(int dtype, int rot, const LieParts& parts, int n, int ne, long long N, const void* X, int ldx, void* G, int sm_count, cudaStream_t st) {
    if (N <= 0) return 0;
    if (rot != ROT_NONE && lie_is_rigid(parts) && (reinterpret_cast<uintptr_t>(G) & 15) == 0) {      // rigid bodies, 16-byte aligned output: TMA image kernel
        const int np = rot == ROT_QUAT ? 4 : 3;
        const int TILE = dtype == 0 ? 64 : 32;
        const size_t smem = size_t(2) * TILE * (9 + np) * 12 * (dtype == 0 ? 4 : 8);
        const long long ntiles = (N + TILE - 1) / TILE, cap = (long long)sm_count * 2;
        const unsigned g = unsigned(ntiles < cap ? ntiles : cap);
        // the dynamic-smem attribute is set once per (kernel, device), not per call.  `kid` tells the four instantiations apart: two of
        // them share one function-pointer TYPE, so a static inside this generic lambda alone would be shared between them.
        static std::atomic<unsigned long long> configured[4];              // one bit per device
        auto go = [&](int kid, auto kern, auto* x, auto* out) {
            int dev = 0;
            cudaGetDevice(&dev);
            const unsigned long long bit = 1ull << (dev & 63);
            if (!(configured[kid].load(std::memory_order_acquire) & bit)) {
                cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
                if (e != cudaSuccess) return int(e);
                configured[kid].fetch_or(bit, std::memory_order_release);
            }
            kern<<<g, TILE, smem, st>>>(rot, N, x, ldx, out, knot_stream_out(N, (long long)(9 + np) * 12 * (dtype == 0 ? 4 : 8)) ? 1 : 0);
            return int(cudaGetLastError());
        };
        if (dtype == 0) return np == 4 ? go(0, errstate_jacobian_tma_kernel<float, 4, 64>, (const float*)X, (float*)G)
                                       : go(1, errstate_jacobian_tma_kernel<float, 3, 64>, (const float*)X, (float*)G);
        return np == 4 ? go(2, errstate_jacobian_tma_kernel<double, 4, 32>, (const double*)X, (double*)G)
                       : go(3, errstate_jacobian_tma_kernel<double, 3, 32>, (const double*)X, (double*)G);
    }
    const unsigned g = grid_for(N * (long long)n * ne, sm_count);
    if (dtype == 0) errstate_jacobian_kernel<float><<<g, 256, 0, st>>>(rot, parts, n, ne, N, (const float*)X, ldx, (float*)G);
    else errstate_jacobian_kernel<double><<<g, 256, 0, st>>>(rot, parts, n, ne, N, (const double*)X, ldx, (double*)G);
    return int(cudaGetLastError());
}
int lie_grad_errstate_jacobian(int dtype, int rot, const LieParts& parts, int n, int ne, long long N, const void* X, int ldx, const void* B, int ldb, void* H,
                               int sm_count, cudaStream_t st) {
    if (N <= 0) return 0;
    const unsigned g = grid_for(N * (long long)ne * ne, sm_count);
    if (dtype == 0) grad_errstate_jacobian_kernel<float><<<g, 256, 0, st>>>(rot, parts, n, ne, N, (const float*)X, ldx, (const float*)B, ldb, (float*)H);
    else grad_errstate_jacobian_kernel<double><<<g, 256, 0, st>>>(rot, parts, n, ne, N, (const double*)X, ldx, (const double*)B, ldb, (double*)H);
    return int(cudaGetLastError());
}
int lie_state_diff(int dtype, int rot, const LieParts& parts, int n, int ne, long long N, const void* X, int ldx, const void* X0, int ldx0, void* dX,
                   int sm_count, cudaStream_t st) {
    if (N <= 0) return 0;
    if (rot != ROT_NONE && lie_is_rigid(parts)) {
        constexpr int TILE = 128;
        const long long ntiles = (N + TILE - 1) / TILE, cap = (long long)sm_count * 12;
        const unsigned g = unsigned(ntiles < cap ? ntiles : cap);
        if (dtype == 0) {
            if (rot == ROT_QUAT) state_diff_tile_kernel<float, 4, TILE><<<g, TILE, 0, st>>>(rot, N, (const float*)X, ldx, (const float*)X0, ldx0, (float*)dX);
            else state_diff_tile_kernel<float, 3, TILE><<<g, TILE, 0, st>>>(rot, N, (const float*)X, ldx, (const float*)X0, ldx0, (float*)dX);
        } else {
            if (rot == ROT_QUAT) state_diff_tile_kernel<double, 4, TILE><<<g, TILE, 0, st>>>(rot, N, (const double*)X, ldx, (const double*)X0, ldx0, (double*)dX);
            else state_diff_tile_kernel<double, 3, TILE><<<g, TILE, 0, st>>>(rot, N, (const double*)X, ldx, (const double*)X0, ldx0, (double*)dX);
        }
        return int(cudaGetLastError());
    }
    const unsigned g = grid_for(N * (long long)ne, sm_count);
    if (dtype == 0) state_diff_kernel<float><<<g, 256, 0, st>>>(rot, parts, n, ne, N, (const float*)X, ldx, (const float*)X0, ldx0, (float*)dX);
    else state_diff_kernel<double><<<g, 256, 0, st>>>(rot, parts, n, ne, N, (const double*)X, ldx, (const double*)X0, ldx0, (double*)dX);
    return int(cudaGetLastError());
}
int lie_project_error_jacobian(int dtype, int rot, const LieParts& parts, int n, int m, int ne, long long N, const void* Z, int ldz, const void* Xn,
                               const void* J, void* Jbar, int sm_count, cudaStream_t st) {
    if (N <= 0) return 0;
    const unsigned g = grid_for(N * (long long)ne * (ne + m), sm_count);
    if (dtype == 0) project_error_jacobian_kernel<float><<<g, 256, 0, st>>>(rot, parts, n, m, ne, N, (const float*)Z, ldz, (const float*)Xn, (const float*)J, (float*)Jbar);
    else project_error_jacobian_kernel<double><<<g, 256, 0, st>>>(rot, parts, n, m, ne, N, (const double*)Z, ldz, (const double*)Xn, (const double*)J, (double*)Jbar);
    return int(cudaGetLastError());
}

}  // namespace rdb
