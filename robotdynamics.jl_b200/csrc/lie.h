// lie.h — host-side entry points of lie.cu (LieState error-state kernels).
#pragma once
#include <cuda_runtime.h>
#include "models.cuh"

namespace rdb {
// LieState{R,P} (reference: src/liestate.jl:75-132): nv vector blocks of lengths P[0..nv), with ONE rotation of type R between
// consecutive blocks (nv - 1 rotations).  n = sum(P) + (nv-1) params(R), errstate_dim = sum(P) + 3 (nv-1).  Euclidean states are the
// special case nv = 1; RigidBody{R} is P = (3, 6) (src/rigidbody.jl:48).
struct LieParts {
    int nv;
    int P[8];
};
inline LieParts lie_parts_euclidean(int n) { LieParts p{}; p.nv = 1; p.P[0] = n; return p; }
inline LieParts lie_parts_rigid() { LieParts p{}; p.nv = 2; p.P[0] = 3; p.P[1] = 6; return p; }
inline bool lie_is_rigid(const LieParts& p) { return p.nv == 2 && p.P[0] == 3 && p.P[1] == 6; }

int lie_errstate_jacobian(int dtype, int rot, const LieParts& parts, int n, int ne, long long N, const void* X, int ldx, void* G, int sm_count, cudaStream_t st);
int lie_grad_errstate_jacobian(int dtype, int rot, const LieParts& parts, int n, int ne, long long N, const void* X, int ldx, const void* B, int ldb, void* H,
                               int sm_count, cudaStream_t st);
int lie_state_diff(int dtype, int rot, const LieParts& parts, int n, int ne, long long N, const void* X, int ldx, const void* X0, int ldx0, void* dX,
                   int sm_count, cudaStream_t st);
// Jbar = G(x+)' [A B] blkdiag(G(x), I) for any LieState: J (n, n+m, N), Z (ldz >= n, N) holds x, Xn (n, N) holds x+, Jbar (ne, ne+m, N)
int lie_project_error_jacobian(int dtype, int rot, const LieParts& parts, int n, int m, int ne, long long N, const void* Z, int ldz, const void* Xn,
                               const void* J, void* Jbar, int sm_count, cudaStream_t st);
}  // namespace rdb
