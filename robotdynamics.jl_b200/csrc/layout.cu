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

}  // namespace rdb
