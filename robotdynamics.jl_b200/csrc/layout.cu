// layout.cu — coalesced transposes between the component-major ("SoA") caller layout and the knot-major layout: the FALLBACK for
// component-major calls the tensor-map kernels cannot take (rows that are not whole 16-byte units, user models, ImplicitMidpoint).
// SoA: W unit-stride streams of N values (stream c starts at base + c*ld);  knot-major: [cnt][W].  Classic 32x32 shared-memory
// tile transpose: both sides read and write full 128-byte lines.
#include "layout.h"

namespace rdb {

template <class T>
__global__ void __launch_bounds__(256) soa_to_aos_kernel(const T* __restrict__ src, long long ld, T* __restrict__ dst, int W, long long cnt) {
    __shared__ T tile[32][33];
    const long long k0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += 8) {                 // r: component within the tile, threadIdx.x: knot
        const int c = c0 + r; const long long k = k0 + threadIdx.x;
        if (c < W && k < cnt) tile[r][threadIdx.x] = src[(long long)c * ld + k];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {                 // r: knot within the tile, threadIdx.x: component
        const long long k = k0 + r; const int c = c0 + threadIdx.x;
        if (c < W && k < cnt) dst[k * W + c] = tile[threadIdx.x][r];
    }
}
template <class T>
__global__ void __launch_bounds__(256) aos_to_soa_kernel(const T* __restrict__ src, T* __restrict__ dst, long long ld, int W, long long cnt) {
    __shared__ T tile[32][33];
    const long long k0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += 8) {                 // r: knot, threadIdx.x: component
        const long long k = k0 + r; const int c = c0 + threadIdx.x;
        if (c < W && k < cnt) tile[r][threadIdx.x] = src[k * W + c];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) {                 // r: component, threadIdx.x: knot
        const int c = c0 + r; const long long k = k0 + threadIdx.x;
        if (c < W && k < cnt) dst[(long long)c * ld + k] = tile[threadIdx.x][r];
    }
}

// narrow arrays (W <= 32, e.g. the 5 x N inputs and 20 x N Jacobians of a Cartpole): one thread per knot walks its row; every
// store / load instruction of a warp touches 32 consecutive elements of one stream, the row side is served by L1.
template <class T>
__global__ void __launch_bounds__(256) soa_to_aos_narrow_kernel(const T* __restrict__ src, long long ld, T* __restrict__ dst, int W, long long cnt) {
    const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (k >= cnt) return;
    for (int c = 0; c < W; ++c) dst[k * W + c] = src[(long long)c * ld + k];
}
template <class T>
__global__ void __launch_bounds__(256) aos_to_soa_narrow_kernel(const T* __restrict__ src, T* __restrict__ dst, long long ld, int W, long long cnt) {
    const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (k >= cnt) return;
    for (int c = 0; c < W; ++c) dst[(long long)c * ld + k] = src[k * W + c];
}

int soa_to_aos(int dtype, const void* src, long long ld, void* dst, int W, long long cnt, cudaStream_t st) {
    if (cnt <= 0) return 0;
    if (W <= 32) {
        const unsigned g = unsigned((cnt + 255) / 256);
        if (dtype == 0) soa_to_aos_narrow_kernel<float><<<g, 256, 0, st>>>((const float*)src, ld, (float*)dst, W, cnt);
        else soa_to_aos_narrow_kernel<double><<<g, 256, 0, st>>>((const double*)src, ld, (double*)dst, W, cnt);
        return int(cudaGetLastError());
    }
    const dim3 grid(unsigned((cnt + 31) / 32), unsigned((W + 31) / 32)), block(32, 8);
    if (dtype == 0) soa_to_aos_kernel<float><<<grid, block, 0, st>>>((const float*)src, ld, (float*)dst, W, cnt);
    else soa_to_aos_kernel<double><<<grid, block, 0, st>>>((const double*)src, ld, (double*)dst, W, cnt);
    return int(cudaGetLastError());
}
int aos_to_soa(int dtype, const void* src, void* dst, long long ld, int W, long long cnt, cudaStream_t st) {
    if (cnt <= 0) return 0;
    if (W <= 32) {
        const unsigned g = unsigned((cnt + 255) / 256);
        if (dtype == 0) aos_to_soa_narrow_kernel<float><<<g, 256, 0, st>>>((const float*)src, (float*)dst, ld, W, cnt);
        else aos_to_soa_narrow_kernel<double><<<g, 256, 0, st>>>((const double*)src, (double*)dst, ld, W, cnt);
        return int(cudaGetLastError());
    }
    const dim3 grid(unsigned((cnt + 31) / 32), unsigned((W + 31) / 32)), block(32, 8);
    if (dtype == 0) aos_to_soa_kernel<float><<<grid, block, 0, st>>>((const float*)src, (float*)dst, ld, W, cnt);
    else aos_to_soa_kernel<double><<<grid, block, 0, st>>>((const double*)src, (double*)dst, ld, W, cnt);
    return int(cudaGetLastError());
}

// ---- device trajectory helpers (traj.cu): column-block copies between dense arrays and the rows of Z, and the time grid ------------
// dst[r * dld + dcol + c] = src[r * sld + scol + c]  for r < rows, c < width: one thread per element, consecutive threads on
// consecutive elements of a row (the dense side is fully coalesced, the strided side touches `width` consecutive elements per row).
template <class T>
__global__ void __launch_bounds__(256) copy_cols_kernel(const T* __restrict__ src, long long sld, int scol, T* __restrict__ dst, long long dld,
                                                        int dcol, int width, long long rows, int zero_tail /* columns after the block, or 0 */) {
    const long long total = rows * (long long)(width + zero_tail);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / (width + zero_tail);
        const int c = int(i - r * (width + zero_tail));
        dst[r * dld + dcol + c] = c < width ? src[r * sld + scol + c] : T(0);
    }
}
int copy_cols(int dtype, const void* src, long long sld, int scol, void* dst, long long dld, int dcol, int width, long long rows, cudaStream_t st) {
    if (rows <= 0 || width <= 0) return 0;
    const long long total = rows * (long long)width;
    const unsigned g = unsigned(total + 255 < (1ll << 20) * 256 ? (total + 255) / 256 : (1ll << 20));
    if (dtype == 0) copy_cols_kernel<float><<<g, 256, 0, st>>>((const float*)src, sld, scol, (float*)dst, dld, dcol, width, rows, 0);
    else copy_cols_kernel<double><<<g, 256, 0, st>>>((const double*)src, sld, scol, (double*)dst, dld, dcol, width, rows, 0);
    return int(cudaGetLastError());
}
// zero `width` columns of every row in [r0, r1)
template <class T>
__global__ void __launch_bounds__(256) zero_cols_kernel(T* __restrict__ dst, long long dld, int dcol, int width, long long r0, long long r1) {
    const long long total = (r1 - r0) * (long long)width;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = r0 + i / width;
        dst[r * dld + dcol + int(i % width)] = T(0);
    }
}
int zero_cols(int dtype, void* dst, long long dld, int dcol, int width, long long r0, long long r1, cudaStream_t st) {
    if (r1 <= r0 || width <= 0) return 0;
    const unsigned g = unsigned(((r1 - r0) * width + 255) / 256);
    if (dtype == 0) zero_cols_kernel<float><<<g, 256, 0, st>>>((float*)dst, dld, dcol, width, r0, r1);
    else zero_cols_kernel<double><<<g, 256, 0, st>>>((double*)dst, dld, dcol, width, r0, r1);
    return int(cudaGetLastError());
}
// time grid of a batch of trajectories stored knot-major (row k * ntraj + j): dt[K-1][j] = 0 (terminal knot,
// src/trajectories.jl:82-83,110), t[k][j] = t0 + sum_{i<k} dt[i][j]; dt_in == nullptr -> every step is dt0.  One thread per trajectory.
__global__ void __launch_bounds__(128) time_grid_kernel(const double* __restrict__ dt_in, double dt0, double t0, double* __restrict__ dt,
                                                        double* __restrict__ t, long long ntraj, int K) {
    const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j >= ntraj) return;
    double acc = t0;
    for (int k = 0; k < K; ++k) {
        const double h = (k == K - 1) ? 0.0 : (dt_in ? dt_in[k * ntraj + j] : dt0);
        dt[k * ntraj + j] = h;
        t[k * ntraj + j] = acc;
        acc += h;
    }
}
int time_grid(const double* dt_in, double dt0, double t0, double* dt, double* t, long long ntraj, int K, cudaStream_t st) {
    if (ntraj <= 0 || K <= 0) return 0;
    time_grid_kernel<<<unsigned((ntraj + 127) / 128), 128, 0, st>>>(dt_in, dt0, t0, dt, t, ntraj, K);
    return int(cudaGetLastError());
}

}  // namespace rdb
