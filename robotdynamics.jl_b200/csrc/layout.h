// layout.h — host entry points of layout.cu (component-major <-> knot-major transposes).
#pragma once
#include <cuda_runtime.h>

namespace rdb {
// src: W streams with leading dimension ld (elements);  dst: [cnt][W]
int soa_to_aos(int dtype, const void* src, long long ld, void* dst, int W, long long cnt, cudaStream_t st);
// src: [cnt][W];  dst: W streams with leading dimension ld
int aos_to_soa(int dtype, const void* src, void* dst, long long ld, int W, long long cnt, cudaStream_t st);
// device-trajectory helpers (traj.cu): dst[r][dcol..dcol+width) = src[r][scol..scol+width) for `rows` rows of leading dimensions dld / sld
int copy_cols(int dtype, const void* src, long long sld, int scol, void* dst, long long dld, int dcol, int width, long long rows, cudaStream_t st);
int zero_cols(int dtype, void* dst, long long dld, int dcol, int width, long long r0, long long r1, cudaStream_t st);
// knot-major time grid of ntraj trajectories x K knots: terminal dt = 0, t = t0 + cumsum(dt); dt_in (K, ntraj) or nullptr -> dt0
int time_grid(const double* dt_in, double dt0, double t0, double* dt, double* t, long long ntraj, int K, cudaStream_t st);
}  // namespace rdb
