// layout.h — host entry points of layout.cu (component-major <-> knot-major transposes).
#pragma once
#include <cuda_runtime.h>

namespace rdb {
// src: W streams with leading dimension ld (elements);  dst: [cnt][W]
int soa_to_aos(int dtype, const void* src, long long ld, void* dst, int W, long long cnt, cudaStream_t st);
// src: [cnt][W];  dst: W streams with leading dimension ld
int aos_to_soa(int dtype, const void* src, void* dst, long long ld, int W, long long cnt, cudaStream_t st);
}  // namespace rdb
