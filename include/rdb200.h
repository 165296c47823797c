/* rdb200.h — C ABI of librdb200.so: B200-native batched dynamics / Jacobian evaluation for the
 * RobotDynamics.jl hot path.  Plain pointers and sizes only; what a Julia `ccall` (or ctypes / cgo) binds.
 *
 * Every entry point cites the reference interface it replaces (file:line in RobotDynamics.jl v0.4.8).
 * The reference evaluates ONE knot point per call inside the caller's `for k in 1:N` loop; each entry point
 * here evaluates the whole batch of N independent knot points in one call.
 *
 * Conventions
 *   - return 0 on success, a negative rdb_status on argument errors, a positive cudaError_t on CUDA failures;
 *     rdb_strerror() explains either.  Nothing is thrown, nothing is allocated for the caller.
 *   - data pointers may be DEVICE pointers (the call only enqueues work on `stream`; the caller synchronises)
 *     or HOST pointers (the call copies in, computes and copies out through an internal pinned, double-buffered
 *     pipeline and returns when the outputs are valid).  All data pointers of one call must be of one kind.
 *   - dtype selects the arithmetic AND the storage type of Z/X/J/... (float or double).  t and dt are always
 *     double, like KnotPoint.t / KnotPoint.dt (src/knotpoint.jl:148-153).
 *   - layout RDB_AOS is the reference's own memory image: Z is the Julia Matrix{T}(n+m, N) of `getdata(z)` columns,
 *     J the Array{T,3}(n, n+m, N) of column-major [A B] DynamicsJacobian data (src/jacobian.jl:26-37), x+ the
 *     Matrix{T}(n, N).  RDB_SOA is component-major: Z is (N, n+m) as a Julia matrix, i.e. each component of [x;u] is a
 *     unit-stride stream of N values; J has n*(n+m) such streams (entry i + n*j), x+ has n.  Both layouts run at the same
 *     speed when N * sizeof(T) is a multiple of 16 (the kernel reads and writes component-major arrays through 2-D TMA
 *     tensor maps); other N are transposed around the knot-major kernel.
 *   - t (KnotPoint.t per knot point) may be NULL (= 0).  It reaches dynamics(model, x, u, t) (src/dynamics.jl:81-83) of models that
 *     can depend on it — user models (rdb_model_create_custom*), whose body may read `t` — with the reference's stage times
 *     t, t+h/2, t+h/2, t+h for RK4 (src/integration.jl:281-284), t, t+h/2, t+h for RK3 (:131-133), t, t+h/2 for the midpoint rules;
 *     it is never differentiated.  The shipped model families are time-invariant (src/dynamics.jl:83): for them t is ignored and
 *     not even transferred.  dt may be NULL (then dt0 is used for every knot point).  Knot points with dt == 0 (terminal,
 *     src/knotpoint.jl:57-67) yield J = [I 0].
 *   - a call changes neither the caller's current CUDA device nor anything else outside its arguments; device pointers must
 *     belong to the context's GPU (RDB_ERR_POINTER_MIX otherwise).  Inputs may be device-resident while outputs are HOST arrays
 *     (RDB_AOS only): the kernel reads Z in place and only J / x+ cross PCIe.
 */
#ifndef RDB200_H
#define RDB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rdb_context rdb_context; /* one per (process, GPU): device id, SM count, staging pipeline */
typedef struct rdb_model rdb_model;     /* immutable model description; replaces an AbstractModel instance */
typedef struct rdb_trajectory rdb_trajectory; /* persistent device mirror of (a batch of) SampledTrajectory, src/trajectories.jl:40-50 */
typedef struct rdb_plan rdb_plan;       /* a validated knot operation on device pointers: launching it is one kernel launch */

typedef enum { RDB_F32 = 0, RDB_F64 = 1 } rdb_dtype;
typedef enum { RDB_AOS = 0, RDB_SOA = 1 } rdb_layout;
/* QuadratureRule subtypes: src/integration.jl:69 (Euler), :109 (RK3), :258 (RK4); RK2 = explicit midpoint
 * (v0.3 name kept by BASELINE; semantics pinned by test/old_tests/linear_tests.jl:135-141).  ImplicitMidpoint: src/integration.jl:620
 * (Newton solve :422-463, implicit-function-theorem Jacobian :524-543); accepted by rdb_discrete_dynamics / rdb_discrete_jacobian /
 * rdb_dynamics_error* for built-in and user models (not by rollout or the error-state form). */
typedef enum { RDB_EULER = 0, RDB_RK2 = 1, RDB_RK3 = 2, RDB_RK4 = 3, RDB_IMPLICIT_MIDPOINT = 4 } rdb_integrator;
/* model families: test/cartpole_model.jl, test/quadrotor.jl, test/rigidbody_test.jl:23-56 (= the Satellite of
 * examples/single_satellite.jl:7-35 with other parameters), test/double_integrator.jl:97-127 */
typedef enum { RDB_CARTPOLE = 0, RDB_QUADROTOR = 1, RDB_BODY = 2, RDB_DOUBLE_INTEGRATOR = 3, RDB_CUSTOM = 4 } rdb_model_kind;
/* rotation parameterisation R of RigidBody{R} (src/liestate.jl:42-46) and velocity_frame (src/rigidbody.jl:258) */
typedef enum { RDB_ROT_NONE = 0, RDB_ROT_QUAT = 1, RDB_ROT_MRP = 2, RDB_ROT_RP = 3 } rdb_rot;
typedef enum { RDB_FRAME_WORLD = 0, RDB_FRAME_BODY = 1 } rdb_frame;

/* operations a plan can hold (the batch entry points below, in order) */
typedef enum { RDB_OP_DYNAMICS = 0, RDB_OP_DISCRETE_DYNAMICS = 1, RDB_OP_JACOBIAN = 2, RDB_OP_DISCRETE_JACOBIAN = 3,
               RDB_OP_DISCRETE_ERROR_JACOBIAN = 4 } rdb_op;

typedef enum {
    RDB_OK = 0,
    RDB_ERR_ARG = -1,            /* bad enum / NULL / negative size        -> Julia ArgumentError           */
    RDB_ERR_NOT_IMPLEMENTED = -2, /* combination has no kernel              -> RobotDynamics.NotImplementedError (src/utils.jl:1-8) */
    RDB_ERR_POINTER_MIX = -3,    /* host and device data pointers mixed in one call                         */
    RDB_ERR_NO_DEVICE = -4,      /* no CUDA device / library built without the kernels                      */
    RDB_ERR_COMPILE = -5         /* user model source does not compile (rdb_last_log() has the NVRTC log)   */
} rdb_status;

int rdb_version(void);
const char* rdb_strerror(int code);

int rdb_create(int device, rdb_context** ctx);
int rdb_destroy(rdb_context* ctx);
/* pinned host memory for the HOST-pointer path (optional; pageable memory also works, slower) */
void* rdb_host_alloc(size_t bytes);
void rdb_host_free(void* p);
/* page-lock memory the caller already owns (a Julia Array{T,3} of Jacobians that a solver reuses every iteration,
 * src/trajectories.jl:40-50 / Altro's per-knot DynamicsJacobian storage) so that the HOST-pointer path copies by DMA at PCIe
 * speed instead of through the driver's pageable staging (~3x slower); register once, unregister before the array is freed.
 * The range need not be page-aligned.  Returns 0, RDB_ERR_ARG for a null / empty range, or the CUDA error as a positive code. */
int rdb_host_register(void* p, size_t bytes);
int rdb_host_unregister(void* p);

/* params packing per kind (doubles):
 *   CARTPOLE          [mc, mp, l, g]                                          test/cartpole_model.jl:9
 *   QUADROTOR         [mass, J(9, row-major), gravity(3), motor_dist, kf, km]  test/quadrotor.jl:36-46
 *   BODY              [mass, J(9, row-major)]                                  test/rigidbody_test.jl:53-54
 *   DOUBLE_INTEGRATOR [D]  (1..3)                                              test/double_integrator.jl:97-100 */
int rdb_model_create(rdb_context* ctx, int kind, int rot, int frame, const double* params, int nparams, rdb_model** model);
/* User-defined model: any `dynamics(model, x, u)` (src/dynamics.jl:81-83), differentiated exactly by forward mode like the
 * reference's `@autodiff`-generated ForwardAD methods (src/jacobian_gen.jl:485-529).  f_body is the BODY of
 *     template <class X, class U> auto f(const X& x, const U& u) const
 * in CUDA C++ against csrc/sdual.cuh: read inputs with get<i>(x), get<j>(u), the time with `t` (a plain scalar: dynamics(model, x,
 * u, t), src/dynamics.jl:81-83), parameters with p[k], write constants as T(0.5),
 * use sin_/cos_/sincos_/exp_/sqrt_/relu_ and + - * /, and `return vec(xdot_0, ..., xdot_{n-1});`.  EuclideanState only
 * (errstate maps are identities), n + m <= 32.  Compiled with NVRTC for sm_100a at creation (syntax) and on first use per
 * (operation, integrator, dtype); works with every batch entry point below except rdb_discrete_error_jacobian's G-seeding
 * (which degenerates to rdb_discrete_jacobian for Euclidean states anyway). */
int rdb_model_create_custom(rdb_context* ctx, int n, int m, const char* f_body, const double* params, int nparams, rdb_model** model);
/* User-defined RigidBody{R}: the reference's rigid-body extension interface is forces(model,x,u) / moments(model,x,u) / mass /
 * inertia (src/rigidbody.jl:244-257, e.g. examples/single_satellite.jl:17-35).  wrench_body is the body of a function that sees
 *     q (attitude as the quaternion Rotations.jl builds for it, Vec4 [w,x,y,z]; use quat_rotate<T>(q, vec3) for q*r),
 *     r, v, w (Vec3), u (m controls), p[k] (user parameters), mass
 * and returns vec(Fx, Fy, Fz, tx, ty, tz): force in the WORLD frame, torque in the BODY frame.  The library supplies the
 * kinematics, Newton / Euler equations, velocity frame, LieState (R, (3,6)) error maps and the error-state Jacobian.
 * J: 3x3 inertia, row-major.  n = 13 (quaternion) or 12 (MRP, Rodrigues). */
int rdb_model_create_custom_rigid(rdb_context* ctx, int rot, int frame, int m, const char* wrench_body, double mass, const double* J,
                                  const double* params, int nparams, rdb_model** model);
/* User-defined model whose state is a general LieState{R,P} (src/liestate.jl:75-132): `nparts` vector blocks of lengths parts[0..nparts)
 * with ONE rotation of type `rot` between consecutive blocks, e.g. parts = (3, 2, 3) with RDB_ROT_QUAT is the [v3, q, v2, q, v3] state of
 * the reference's QuatState(16, (4, 10)) example (src/liestate.jl:93-100); n = sum(parts) + (nparts-1) params(R), errstate_dim =
 * sum(parts) + 3 (nparts-1).  f_body as for rdb_model_create_custom (it sees the full state x, rotations NOT renormalised).
 * rdb_errstate_jacobian / rdb_grad_errstate_jacobian / rdb_state_diff use the partition (block-diagonal I / ∇differential per
 * rotation, src/liestate.jl:210-320) and rdb_discrete_error_jacobian returns G(x+)' [A B] blkdiag(G(x), I) in errstate coordinates. */
int rdb_model_create_custom_lie(rdb_context* ctx, int rot, int nparts, const int* parts, int m, const char* f_body, const double* params,
                                int nparams, rdb_model** model);
/* compile-only check of a user model body (needs no GPU); on failure rdb_last_log() returns the compiler log */
int rdb_custom_check(int n, int m, const char* f_body, int nparams, int dtype);
int rdb_custom_rigid_check(int rot, int frame, int m, const char* wrench_body, int nparams, int dtype);
const char* rdb_last_log(void);
int rdb_model_destroy(rdb_model* model);
/* state_dim / control_dim / errstate_dim  (src/functionbase.jl:124-135, src/liestate.jl:124) */
int rdb_model_dims(const rdb_model* model, int* n, int* m, int* nerr);

/* dynamics(model, x, u, t)                                   src/dynamics.jl:81-83 */
int rdb_dynamics(const rdb_model* model, int dtype, int layout, int64_t N, const void* Z, const double* t,
                 void* xdot, void* stream);
/* discrete_dynamics(DiscretizedDynamics{L,Q}, x, u, t, dt)   src/discrete_dynamics.jl:80-81, src/discretized_dynamics.jl:203-206 */
int rdb_discrete_dynamics(const rdb_model* model, int integrator, int dtype, int layout, int64_t N, const void* Z,
                          const double* t, const double* dt, double dt0, void* xn, void* stream);
/* jacobian!(sig, diff, model::ContinuousDynamics, J, xdot, z)  src/functionbase.jl:242, src/jacobian_gen.jl:485-529
 * xdot may be NULL */
int rdb_jacobian(const rdb_model* model, int dtype, int layout, int64_t N, const void* Z, const double* t, void* J,
                 void* xdot, void* stream);
/* jacobian!(sig, diff, DiscretizedDynamics{L,Q}, J, y, z)  == v0.3 discrete_jacobian!(Q, J, model, z)
 * src/discretized_dynamics.jl:169-219, src/integration.jl:85-93,149-177,302-337.  xn (= x+) may be NULL. */
int rdb_discrete_jacobian(const rdb_model* model, int integrator, int dtype, int layout, int64_t N, const void* Z,
                          const double* t, const double* dt, double dt0, void* J, void* xn, void* stream);
/* Error-state ("LieState") discrete Jacobian, what Altro / TrajectoryOptimization consume for RotationState models:
 *   Jbar = G(x+)' [A B] blkdiag(G(x), I),  G = errstate_jacobian (src/liestate.jl:262-298), nerr x (nerr + m) column-major per
 * knot (jacobian_width = errstate_dim + control_dim, src/functionbase.jl:135).  Computed in one pass by seeding forward mode with
 * the columns of G(x); never forms the n x (n+m) Jacobian.  For EuclideanState models G = I and this equals
 * rdb_discrete_jacobian.  Layouts as for J with (nerr, nerr+m) in place of (n, n+m); xn (= x+, n values) may be NULL. */
int rdb_discrete_error_jacobian(const rdb_model* model, int integrator, int dtype, int layout, int64_t N, const void* Z,
                                const double* t, const double* dt, double dt0, void* Jbar, void* xn, void* stream);
/* errstate_jacobian!(model, G, x)                            src/statevectortype.jl:119-122, src/liestate.jl:262-298
 * X: N states with leading dimension ldx (>= n; pass n+m to read the states straight out of Z).
 * G: (n, nerr, N) column-major per knot, fully written (zeros included). */
int rdb_errstate_jacobian(const rdb_model* model, int dtype, int64_t N, const void* X, int ldx, void* G, void* stream);
/* ∇errstate_jacobian!(model, ∇G, x, xbar)                    src/liestate.jl:300-320.  H: (nerr, nerr, N), fully written. */
int rdb_grad_errstate_jacobian(const rdb_model* model, int dtype, int64_t N, const void* X, int ldx, const void* Xbar,
                               int ldb, void* H, void* stream);
/* state_diff(model, x, x0)  (CayleyMap)                      src/liestate.jl:210-260.  dX: (nerr, N). */
int rdb_state_diff(const rdb_model* model, int dtype, int64_t N, const void* X, int ldx, const void* X0, int ldx0,
                   void* dX, void* stream);
/* rollout!(sig, dmodel, Z, x0)                               src/trajectories.jl:436-441, src/discrete_dynamics.jl:217-235
 * x0 (n, ntraj); U (m, K-1, ntraj); t, dt (K, ntraj) or NULL; X (n, K, ntraj) out.  Parallel over trajectories. */
int rdb_rollout(const rdb_model* model, int integrator, int dtype, int64_t ntraj, int K, const void* x0, const void* U,
                const double* t, const double* dt, double dt0, void* X, void* stream);

/* dynamics_error(dmodel, z2, z1) and dynamics_error_jacobian!(sig, diff, dmodel, J2, J1, y2, y1, z2, z1) for N pairs of knot points
 * (src/discrete_dynamics.jl:116-200): Z1 (n+m, N) holds z1 = [x1; u1], Z2 (ld2, N) holds x2 in its first n rows (pass the next knots'
 * z rows with ld2 = n+m).  Explicit rules: e = discrete_dynamics(z1) - x2, J1 = the discrete Jacobian, J2 = [-I 0]
 * (src/discrete_dynamics.jl:137-138,181-182).  RDB_IMPLICIT_MIDPOINT: e = x1 + h f((x1+x2)/2, u1, t + h/2) - x2, J1 = [I + h/2 A, h B],
 * J2 = [h/2 A - I, 0] (src/integration.jl:640-700).  e (n, N); J1, J2 (n, n+m, N); any of J2 / J1 / e may be NULL. */
int rdb_dynamics_error(const rdb_model* model, int integrator, int dtype, int64_t N, const void* Z1, const void* Z2, int ld2,
                       const double* t, const double* dt, double dt0, void* e, void* stream);
int rdb_dynamics_error_jacobian(const rdb_model* model, int integrator, int dtype, int64_t N, const void* Z1, const void* Z2, int ld2,
                                const double* t, const double* dt, double dt0, void* J2, void* J1, void* e, void* stream);

/* ---- pre-validated launches --------------------------------------------------------------------------------------------------
 * A solver evaluates the SAME batch (same buffers, 10^2..10^3 knot points) every iteration; there the per-call host work (pointer
 * classification, argument checks, tensor-map encoding) costs more than the kernel.  A plan does that work once:
 * rdb_plan_create validates one of the batch operations above on DEVICE pointers (op = rdb_op; J / out as for that entry point;
 * integrator ignored by the continuous ops), rdb_plan_launch enqueues it on `stream` — one kernel launch, nothing else — any
 * number of times (it reads whatever the buffers hold at execution time), also under CUDA-graph capture.  The reference has no
 * counterpart: its per-knot calls are plain Julia method calls (src/discretized_dynamics.jl:129-136). */
int rdb_plan_create(const rdb_model* model, int op, int integrator, int dtype, int layout, int64_t N, const void* Z, const double* t,
                    const double* dt, double dt0, void* J, void* out, rdb_plan** plan);
int rdb_plan_launch(const rdb_plan* plan, void* stream);
/* shared != 0: this plan's launches will run WHILE other kernels use the same GPU (another model's plan on a second stream, as in a
 * mixed-model sweep).  Some rigid-body kernels are fastest on small CTAs, four to an SM, when they have the GPU to themselves; next to
 * foreign CTAs those interleave SM by SM and both kernels slow down.  A shared plan keeps the wide tiling whose CTAs each fill an SM
 * (mixed Cartpole + Quadrotor sweep: 106.7 us against 124.7 us).  Results are identical either way. */
int rdb_plan_set_shared(rdb_plan* plan, int shared);
int rdb_plan_destroy(rdb_plan* plan);

/* ---- persistent device trajectory -------------------------------------------------------------------------------------------------
 * The device mirror of SampledTrajectory (src/trajectories.jl:40-50) for `ntraj` independent trajectories of K knot points each,
 * stored knot-major across the batch: row k * ntraj + j holds z = [x;u] of knot k of trajectory j; times and steps (double) are
 * stored the same way, the terminal knot has dt = 0 and zero controls (src/trajectories.jl:82-83,110; src/knotpoint.jl:57-67).
 * For ntraj == 1 — one solve of Altro / TrajectoryOptimization — this is exactly the gathered Matrix{T}(n+m, K) of the host's
 * Vector{KnotPoint} (an array of pointers to mutable structs, src/knotpoint.jl:213-217), uploaded once instead of every call.
 * All setters / getters take HOST or DEVICE pointers of dense arrays in the same knot-major order:
 *   X (n, ntraj, K) as a Julia array == rows k * ntraj + j of n values;  U (m, ntraj, K-1 or K);  dt (ntraj, K) doubles. */
int rdb_trajectory_create(const rdb_model* model, int dtype, int64_t ntraj, int K, rdb_trajectory** traj);
int rdb_trajectory_destroy(rdb_trajectory* traj);
int rdb_trajectory_dims(const rdb_trajectory* traj, int64_t* ntraj, int* K, int* n, int* m, int* dtype);
/* device pointers of the mirror itself (zero-copy: wrap them as CuArrays / tensors): Z (n+m, ntraj, K), t and dt (ntraj, K) */
int rdb_trajectory_data(const rdb_trajectory* traj, void** Z, double** t, double** dt);
/* setstates!(Z, X) / setcontrols!(Z, U) / setinitialstate / settimes   src/trajectories.jl:215-250 */
int rdb_trajectory_set_states(rdb_trajectory* traj, const void* X, void* stream);
int rdb_trajectory_set_initial_state(rdb_trajectory* traj, const void* x0 /* (n, ntraj) */, void* stream);
int rdb_trajectory_set_controls(rdb_trajectory* traj, const void* U, int knots /* K-1 (terminal control := 0) or K */, void* stream);
/* steps dt (ntraj, K) or NULL (every step dt0); the terminal step is forced to 0 and t[k] = t0 + sum_{i<k} dt[i] */
int rdb_trajectory_set_timesteps(rdb_trajectory* traj, const double* dt, double dt0, double t0, void* stream);
/* states(Z) / controls(Z)   src/trajectories.jl:181-199 */
int rdb_trajectory_get_states(rdb_trajectory* traj, void* X, void* stream);
int rdb_trajectory_get_controls(rdb_trajectory* traj, void* U, void* stream);
/* rollout!(sig, dmodel, Z, x0): x_{k+1} = discrete_dynamics(dmodel, Z[k]) from the state of knot 0 and the stored controls, times
 * and steps   src/trajectories.jl:436-441, src/discrete_dynamics.jl:217-235 */
int rdb_trajectory_rollout(rdb_trajectory* traj, int integrator, void* stream);
/* jacobian!(sig, diff, dmodel, J[k], y[k], Z[k]) for every knot of every trajectory (src/discretized_dynamics.jl:129-136):
 * J (n, n+m, ntraj, K) — or, error_state != 0, Jbar (nerr, nerr+m, ntraj, K) as rdb_discrete_error_jacobian — host or device;
 * xn (n, ntraj, K) or NULL.  Terminal knots (dt = 0) give [I 0]. */
int rdb_trajectory_linearize(rdb_trajectory* traj, int integrator, int error_state, void* J, void* xn, void* stream);
/* Forward pass + linearisation in one call.  chunks <= 1 (default): two phases on `stream` — the rollout writes
 * z = [x;u] rows in place, one Jacobian launch then streams all K * ntraj knots at full-GPU rate.  chunks > 1: the measured
 * alternative, a pipeline on two internal streams (rollout in chunks of knots, Jacobians of each finished chunk — a contiguous row
 * range of the knot-major batch — meanwhile), joined back into `stream`; on B200 it is slower than the two phases. */
int rdb_trajectory_rollout_linearize(rdb_trajectory* traj, int integrator, int error_state, int chunks, void* J, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RDB200_H */
