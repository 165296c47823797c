/* c_abi_example.c — the C ABI from plain C (what a Julia `ccall` does, without Julia).
 *
 *   gcc -std=c99 -I include examples/c_abi_example.c -o /tmp/rdb_example -L robotdynamics.jl_b200 -lrdb200 \
 *       -Wl,-rpath,$PWD/robotdynamics.jl_b200 && /tmp/rdb_example
 *
 * Evaluates the RK4 discrete Jacobians of 8 Cartpole knot points through HOST pointers (the library stages them through the GPU)
 * and prints the first one.  Without a GPU it prints the library's error and exits with status 3 — there is no CPU fallback. */
#include <stdio.h>
#include <stdlib.h>

#include "rdb200.h"

#define CHECK(call)                                                                 \
    do {                                                                            \
        int rc_ = (call);                                                           \
        if (rc_ != 0) {                                                             \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, rdb_strerror(rc_));       \
            return rc_ == RDB_ERR_NO_DEVICE ? 3 : 1;                                \
        }                                                                           \
    } while (0)

int main(void) {
    enum { N = 8, n = 4, m = 1 };
    rdb_context* ctx = NULL;
    rdb_model* cartpole = NULL;
    const double params[4] = {1.0, 0.2, 0.5, 9.81}; /* mc, mp, l, g  (test/cartpole_model.jl:9) */
    double Z[N][n + m], J[N][n + m][n], xn[N][n];   /* J[k] is a column-major n x (n+m) matrix: J[k][col][row] */
    int k, i, j, dn, dm, dne;

    printf("rdb200 version %d\n", rdb_version());
    CHECK(rdb_create(0, &ctx));
    CHECK(rdb_model_create(ctx, RDB_CARTPOLE, RDB_ROT_NONE, RDB_FRAME_WORLD, params, 4, &cartpole));
    CHECK(rdb_model_dims(cartpole, &dn, &dm, &dne));
    for (k = 0; k < N; ++k)
        for (i = 0; i < n + m; ++i) Z[k][i] = 0.1 * (i + 1) + 0.01 * k;
    CHECK(rdb_discrete_jacobian(cartpole, RDB_RK4, RDB_F64, RDB_AOS, N, Z, NULL, NULL, 0.01, J, xn, NULL));
    printf("n = %d, m = %d, nerr = %d; [A B] of knot 0:\n", dn, dm, dne);
    for (i = 0; i < n; ++i) {
        for (j = 0; j < n + m; ++j) printf(" % .6e", J[0][j][i]);
        printf("\n");
    }
    rdb_model_destroy(cartpole);
    rdb_destroy(ctx);
    return 0;
}
