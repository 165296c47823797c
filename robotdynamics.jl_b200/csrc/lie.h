// lie.h — host-side entry points of lie.cu (LieState error-state kernels).
#pragma once
#include <cuda_runtime.h>
#include "models.cuh"

namespace rdb {
int lie_errstate_jacobian(int dtype, int rot, int n, int ne, long long N, const void* X, int ldx, void* G, int sm_count, cudaStream_t st);
int lie_grad_errstate_jacobian(int dtype, int rot, int n, int ne, long long N, const void* X, int ldx, const void* B, int ldb, void* H,
                               int sm_count, cudaStream_t st);
int lie_state_diff(int dtype, int rot, int n, int ne, long long N, const void* X, int ldx, const void* X0, int ldx0, void* dX,
                   int sm_count, cudaStream_t st);
}  // namespace rdb
