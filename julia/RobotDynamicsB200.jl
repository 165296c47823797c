# RobotDynamicsB200.jl — the reference-side binding of librdb200.so (include/rdb200.h).
#
# NOT executed in this repository's image (Julia is not installed there); this is the binding a RobotDynamics.jl
# maintainer would add (`julia/runtests.jl` is its test-suite, mirroring the reference's own test/cartpole_test.jl:48-72).  It introduces one new DiffMethod subtype, `B200`, the extension point the reference documents
# (docs/src/autodiff.md:20-21), and batched methods of the reference's own generic functions.  Pure ccall + gather/scatter.
module RobotDynamicsB200

using RobotDynamics
const RD = RobotDynamics
using StaticArrays, Rotations, LinearAlgebra

const LIB = get(ENV, "RDB200_LIB", joinpath(@__DIR__, "..", "robotdynamics.jl_b200", "librdb200.so"))

"Batched forward-mode evaluation on a B200 through librdb200 (a new `DiffMethod`, docs/src/autodiff.md:20-21)."
struct B200 <: RD.DiffMethod end

# ---- status codes -> the exceptions the reference throws (src/utils.jl:1-8, src/discretized_dynamics.jl:243-249) ----
function check(rc::Cint, what)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:rdb_strerror, LIB), Cstring, (Cint,), rc))
    rc == -2 && throw(RD.NotImplementedError(what))
    rc == -1 && throw(ArgumentError("$what: $msg"))
    error("$what: rdb200 status $rc: $msg")
end

# ---- context and model handles ----------------------------------------------------------------------------------------
mutable struct Context
    ptr::Ptr{Cvoid}
    function Context(device::Integer=0)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:rdb_create, LIB), Cint, (Cint, Ref{Ptr{Cvoid}}), device, r), "rdb_create")
        finalizer(c -> ccall((:rdb_destroy, LIB), Cint, (Ptr{Cvoid},), c.ptr), new(r[]))
    end
end
const CTX = Ref{Context}()
context() = isassigned(CTX) ? CTX[] : (CTX[] = Context(0))

"""
    pin!(A::Array) / unpin!(A)

Page-lock an array the solver already owns (the `Array{T,3}` of Jacobians it reuses every iteration, a `SampledTrajectory`'s gathered
`(n+m) x N` buffer) so the host-pointer path moves it by DMA at PCIe speed; `unpin!` before the array is freed.
"""
pin!(A::Array) = (check(ccall((:rdb_host_register, LIB), Cint, (Ptr{Cvoid}, Csize_t), pointer(A), sizeof(A)), "rdb_host_register"); A)
unpin!(A::Array) = (check(ccall((:rdb_host_unregister, LIB), Cint, (Ptr{Cvoid},), pointer(A)), "rdb_host_unregister"); A)

const KIND_CARTPOLE, KIND_QUADROTOR, KIND_BODY, KIND_DI = Cint(0), Cint(1), Cint(2), Cint(3)
rotcode(::Type{<:QuatRotation}) = Cint(1); rotcode(::Type{<:MRP}) = Cint(2); rotcode(::Type{<:RodriguesParam}) = Cint(3)
framecode(model) = RD.velocity_frame(model) == :body ? Cint(1) : Cint(0)          # src/rigidbody.jl:258

mutable struct Handle
    ptr::Ptr{Cvoid}
end
function Handle(kind, rot, frame, params::Vector{Float64})
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:rdb_model_create, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Cdouble}, Cint, Ref{Ptr{Cvoid}}),
                context().ptr, kind, rot, frame, params, length(params), r), "rdb_model_create")
    finalizer(h -> ccall((:rdb_model_destroy, LIB), Cint, (Ptr{Cvoid},), h.ptr), Handle(r[]))
end

"""
    handle(model)

Describe a reference model to the library (parameter packing as in include/rdb200.h).  The model TYPES the reference's hot path
is exercised with live in its test / example scripts (test/cartpole_model.jl, test/quadrotor.jl, examples/single_satellite.jl), not
in the package, so they are recognised structurally — by the fields / methods those definitions have — instead of by name:

* a `ContinuousDynamics` with fields `mc, mp, l, g`                          → the Cartpole of test/cartpole_model.jl:1-30
* a `RigidBody{R}` with fields `gravity, motor_dist, kf, km`                → the Quadrotor of test/quadrotor.jl:20-96
* a `RigidBody{R}` whose owner declared `wrench_kind(::MyBody) = :force_moment`  → F_world = q*u[1:3], M_body = u[4:6] (mass, inertia
  from `RD.mass`, `RD.inertia`): the Satellite of examples/single_satellite.jl:7-35 and the Body of test/rigidbody_test.jl:23-56
* anything else: `custom_handle` / `custom_rigid_handle` below (user dynamics compiled by NVRTC), or `NotImplementedError`.
"""
function handle(m::RD.ContinuousDynamics)
    T = typeof(m)
    if all(f -> hasfield(T, f), (:mc, :mp, :l, :g))
        return Handle(KIND_CARTPOLE, Cint(0), Cint(0), Float64[m.mc, m.mp, m.l, m.g])
    end
    throw(RD.NotImplementedError("no B200 kernel for $(T): describe it with custom_handle(n, m, f_body)"))
end
"Opt-in trait for rigid bodies whose wrench is `F_world = q*u[1:3]`, `M_body = u[4:6]`."
wrench_kind(::RD.RigidBody) = :unknown
rowmajor(J) = collect(Float64, vec(Matrix(J)'))
function handle(m::RD.RigidBody{R}) where {R}
    T = typeof(m)
    if all(f -> hasfield(T, f), (:gravity, :motor_dist, :kf, :km))
        return Handle(KIND_QUADROTOR, rotcode(R), framecode(m),
                      Float64[RD.mass(m); rowmajor(RD.inertia(m)); collect(Float64, m.gravity); m.motor_dist; m.kf; m.km])
    elseif wrench_kind(m) == :force_moment
        return Handle(KIND_BODY, rotcode(R), framecode(m), Float64[RD.mass(m); rowmajor(RD.inertia(m))])
    end
    throw(RD.NotImplementedError("no B200 kernel for $(T): declare RobotDynamicsB200.wrench_kind(::$(nameof(T))) = :force_moment " *
                                 "or describe its wrench with custom_rigid_handle"))
end

"""
    custom_handle(n, m, f_body; params=Float64[])

Any user `dynamics(model, x, u)`: `f_body` is the CUDA C++ body of `f(x, u)` (see include/rdb200.h, rdb_model_create_custom); the
library compiles it with NVRTC and differentiates it by forward mode, like `@autodiff` does on the CPU (src/jacobian_gen.jl:64-82).
"""
function custom_handle(n::Integer, m::Integer, f_body::AbstractString; params::Vector{Float64}=Float64[])
    r = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:rdb_model_create_custom, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Cstring, Ptr{Cdouble}, Cint, Ref{Ptr{Cvoid}}),
               context().ptr, n, m, f_body, params, length(params), r)
    rc == -5 && error("user model does not compile:\n" * unsafe_string(ccall((:rdb_last_log, LIB), Cstring, ())))
    check(rc, "rdb_model_create_custom")
    finalizer(h -> ccall((:rdb_model_destroy, LIB), Cint, (Ptr{Cvoid},), h.ptr), Handle(r[]))
end

"RigidBody{R} with user forces / moments (src/rigidbody.jl:244-257): `wrench_body` returns vec(F_world..., tau_body...)."
function custom_rigid_handle(::Type{R}, m::Integer, wrench_body::AbstractString, mass::Real, J::AbstractMatrix;
                             params::Vector{Float64}=Float64[], bodyframe::Bool=false) where {R}
    r = Ref{Ptr{Cvoid}}(C_NULL)
    Jrow = collect(Float64, vec(Matrix(J)'))
    check(ccall((:rdb_model_create_custom_rigid, LIB), Cint,
                (Ptr{Cvoid}, Cint, Cint, Cint, Cstring, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ref{Ptr{Cvoid}}),
                context().ptr, rotcode(R), Cint(bodyframe), m, wrench_body, mass, Jrow, params, length(params), r), "rdb_model_create_custom_rigid")
    finalizer(h -> ccall((:rdb_model_destroy, LIB), Cint, (Ptr{Cvoid},), h.ptr), Handle(r[]))
end

const HANDLES = IdDict{Any,Handle}()
gethandle(m) = get!(() -> handle(m), HANDLES, m)

# RK2 (explicit midpoint) is the v0.3 name BASELINE.json keeps; it is absent from src/ at v0.4.8 (test/old_tests/linear_tests.jl:135-141
# pins its meaning), so it is (re)defined here — as a full QuadratureRule, usable on the CPU path as well.
if isdefined(RD, :RK2)
    const RK2 = RD.RK2
else
    "Explicit midpoint: x + h f(x + h/2 f(x, u, t), u, t + h/2)."
    struct RK2 <: RD.Explicit end
    RK2(n::Integer, m::Integer) = RK2()
    RD.integrate(::RK2, model, x, u, t, h) = x + h * RD.dynamics(model, x + (h / 2) * RD.dynamics(model, x, u, t), u, t + h / 2)
end
const INTEGRATOR = Dict{Any,Cint}(RD.Euler => 0, RK2 => 1, RD.RK3 => 2, RD.RK4 => 3, RD.ImplicitMidpoint => 4)
dtypecode(::Type{Float32}) = Cint(0); dtypecode(::Type{Float64}) = Cint(1)

# ---- gather / scatter between the reference containers and the batched images ---------------------------------------
"Z as Matrix{T}(n+m, N): column k is `getdata(Z[k])` (src/knotpoint.jl:196); also the Float64 dt and t vectors."
function gather(Z::RD.SampledTrajectory{n,m,T}) where {n,m,T}
    N = length(Z)
    data = Matrix{T}(undef, n + m, N)
    dts = Vector{Float64}(undef, N)
    ts = Vector{Float64}(undef, N)
    @inbounds for k in 1:N
        data[:, k] .= RD.getdata(Z[k])
        dts[k] = RD.timestep(Z[k])
        ts[k] = RD.time(Z[k])
    end
    data, dts, ts
end

# ---- the hot path: batched methods of the reference's own generic functions -----------------------------------------------
"""
    RD.jacobian!(sig, ::B200, dmodel::DiscretizedDynamics{L,Q}, J::Array{T,3}, y::Matrix{T}, Z::SampledTrajectory)

Discrete Jacobians of every knot point of `Z` in one call: `J[:, :, k]` is the `n×(n+m)` `[A B]` of knot k (the memory of a
`DynamicsJacobian`, src/jacobian.jl:26-37), `y[:, k] = x⁺_k`.  Replaces the caller's
`for k in 1:N; jacobian!(sig, diff, dmodel, J[k], y[k], Z[k]); end` (src/discretized_dynamics.jl:129-136).
"""
function RD.jacobian!(sig::RD.FunctionSignature, ::B200, dmodel::RD.DiscretizedDynamics{L,Q}, J::Array{T,3}, y::Matrix{T},
                      Z::RD.SampledTrajectory) where {L,Q,T}
    data, dts, ts = gather(Z)
    discrete_jacobian_batch!(dmodel, J, y, data, dts; times=ts)
end

"Same, on pre-gathered (or device-resident CuArray) data: no per-knot host work at all."
function discrete_jacobian_batch!(dmodel::RD.DiscretizedDynamics{L,Q}, J, y, data, dts; times=nothing, stream=C_NULL) where {L,Q}
    T = eltype(data)
    N = size(data, 2)
    h = gethandle(dmodel.continuous_dynamics)
    GC.@preserve data dts times J y begin
        check(ccall((:rdb_discrete_jacobian, LIB), Cint,
                    (Ptr{Cvoid}, Cint, Cint, Cint, Int64, Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                    h.ptr, INTEGRATOR[Q], dtypecode(T), 0 #= RDB_AOS =#, N, pointer(data), times === nothing ? C_NULL : pointer(times),
                    pointer(dts), 0.0, pointer(J), y === nothing ? C_NULL : pointer(y), stream), "rdb_discrete_jacobian")
    end
    nothing
end

"Single knot point, for API fidelity only (one kernel launch per call — use the batched method in loops)."
function RD.jacobian!(sig::RD.FunctionSignature, ::B200, dmodel::RD.DiscretizedDynamics, J, y, z::RD.AbstractKnotPoint)
    n, m = RD.dims(dmodel)
    data = reshape(Vector(RD.getdata(z)), n + m, 1)
    Jb = Array{eltype(data),3}(undef, n, n + m, 1); yb = Matrix{eltype(data)}(undef, n, 1)
    discrete_jacobian_batch!(dmodel, Jb, yb, data, [RD.timestep(z)]; times=[Float64(RD.time(z))])
    J .= @view Jb[:, :, 1]
    y .= @view yb[:, 1]
    nothing
end

"v0.3 spelling kept by BASELINE.json: `discrete_jacobian!(RK4, ∇f, model, z)` (README.md:81-82)."
discrete_jacobian!(::Type{Q}, ∇f, model::RD.AbstractModel, z) where {Q<:RD.QuadratureRule} =
    RD.jacobian!(RD.StaticReturn(), B200(), RD.DiscretizedDynamics{Q}(model), ∇f, zeros(RD.state_dim(model)), z)

"Error-state expansion `Jbar[:, :, k] = G(x⁺)' [A B] blkdiag(G(x), I)` (n̄ × (n̄+m)) for every knot point, one pass."
function discrete_error_jacobian_batch!(dmodel::RD.DiscretizedDynamics{L,Q}, Jbar::Array{T,3}, y, data::Matrix{T}, dts; stream=C_NULL) where {L,Q,T}
    check(ccall((:rdb_discrete_error_jacobian, LIB), Cint,
                (Ptr{Cvoid}, Cint, Cint, Cint, Int64, Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                gethandle(dmodel.continuous_dynamics).ptr, INTEGRATOR[Q], dtypecode(T), 0, size(data, 2), pointer(data), C_NULL, pointer(dts), 0.0,
                pointer(Jbar), y === nothing ? C_NULL : pointer(y), stream), "rdb_discrete_error_jacobian")
end

"`errstate_jacobian!` for every knot point: `G[:, :, k]` is `n×n̄`, fully written (src/liestate.jl:262-298 writes non-zeros only)."
function errstate_jacobian_batch!(model::RD.AbstractModel, G::Array{T,3}, X::Matrix{T}; stream=C_NULL) where T
    check(ccall((:rdb_errstate_jacobian, LIB), Cint, (Ptr{Cvoid}, Cint, Int64, Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}),
                gethandle(model).ptr, dtypecode(T), size(X, 2), pointer(X), size(X, 1), pointer(G), stream), "rdb_errstate_jacobian")
end

"`state_diff(model, x, x0)` for every knot point (src/liestate.jl:210-260)."
function state_diff_batch!(model::RD.AbstractModel, dX::Matrix{T}, X::Matrix{T}, X0::Matrix{T}; stream=C_NULL) where T
    check(ccall((:rdb_state_diff, LIB), Cint, (Ptr{Cvoid}, Cint, Int64, Ptr{Cvoid}, Cint, Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}),
                gethandle(model).ptr, dtypecode(T), size(X, 2), pointer(X), size(X, 1), pointer(X0), size(X0, 1), pointer(dX), stream),
          "rdb_state_diff")
end

"`rollout!` for many independent trajectories: X (n, K, ntraj), U (m, K-1, ntraj) (src/trajectories.jl:436-441)."
function rollout_batch!(dmodel::RD.DiscretizedDynamics{L,Q}, X::Array{T,3}, x0::Matrix{T}, U::Array{T,3}, dt::Float64; stream=C_NULL) where {L,Q,T}
    check(ccall((:rdb_rollout, LIB), Cint,
                (Ptr{Cvoid}, Cint, Cint, Int64, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Ptr{Cvoid}, Ptr{Cvoid}),
                gethandle(dmodel.continuous_dynamics).ptr, INTEGRATOR[Q], dtypecode(T), size(x0, 2), size(X, 2), pointer(x0), pointer(U),
                C_NULL, C_NULL, dt, pointer(X), stream), "rdb_rollout")
end

# ---- pre-validated launches (rdb_plan_*): what a solver loop calls every iteration on device-resident data --------------------------------
"""
    LaunchPlan(dmodel, Z, dts, J; y=nothing, times=nothing, error_state=false, shared_gpu=false)

Validates `jacobian!(sig, B200(), dmodel, J, y, Z)` once for DEVICE arrays (anything `pointer` works on: `CuArray`s — `Z` is `(n+m) x N`,
`J` is `n x (n+m) x N`, or `n̄ x (n̄+m) x N` with `error_state`); `launch!(plan; stream)` is then a single kernel launch that evaluates
whatever the arrays hold (6–7 µs of host time, capturable in a CUDA graph).  `shared_gpu=true` for plans whose launches overlap other
kernels on the same GPU (rdb_plan_set_shared).  The reference has no counterpart (per-knot method calls, src/discretized_dynamics.jl:129-136).
"""
mutable struct LaunchPlan
    ptr::Ptr{Cvoid}
    keep::Any             # the arrays and the model handle the plan points at
end
function LaunchPlan(dmodel::RD.DiscretizedDynamics{L,Q}, Z, dts, J; y=nothing, times=nothing, error_state::Bool=false, shared_gpu::Bool=false) where {L,Q}
    h = gethandle(dmodel.continuous_dynamics)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    op = error_state ? Cint(4) : Cint(3)          # RDB_OP_DISCRETE_ERROR_JACOBIAN / RDB_OP_DISCRETE_JACOBIAN
    check(ccall((:rdb_plan_create, LIB), Cint,
                (Ptr{Cvoid}, Cint, Cint, Cint, Cint, Int64, Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Ptr{Cvoid}, Ptr{Cvoid}, Ref{Ptr{Cvoid}}),
                h.ptr, op, INTEGRATOR[Q], dtypecode(eltype(Z)), 0 #= RDB_AOS =#, size(Z, 2), pointer(Z), times === nothing ? C_NULL : pointer(times),
                pointer(dts), 0.0, pointer(J), y === nothing ? C_NULL : pointer(y), r), "rdb_plan_create")
    shared_gpu && check(ccall((:rdb_plan_set_shared, LIB), Cint, (Ptr{Cvoid}, Cint), r[], 1), "rdb_plan_set_shared")
    finalizer(p -> ccall((:rdb_plan_destroy, LIB), Cint, (Ptr{Cvoid},), p.ptr), LaunchPlan(r[], (Z, dts, J, y, times, h)))
end
launch!(p::LaunchPlan; stream=C_NULL) = check(ccall((:rdb_plan_launch, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), p.ptr, stream), "rdb_plan_launch")

# ---- persistent device trajectory (rdb_trajectory_*): the device mirror of SampledTrajectory ------------------------------------------
"""
    DeviceTrajectory(dmodel, Z::SampledTrajectory)            # one trajectory, uploaded once
    DeviceTrajectory(dmodel, T, ntraj, K)                     # a batch of `ntraj` trajectories of K knot points

Keeps `[x;u]`, `t`, `dt` of every knot point in GPU memory (knot-major across the batch, include/rdb200.h), so that solver iterations
only send what changed (`setcontrols!`, `setstates!`: src/trajectories.jl:215-250) and receive what they need (Jacobians into host
`Array`s): no per-iteration gather of the host's Vector{KnotPoint}, no re-upload of the inputs.
"""
mutable struct DeviceTrajectory{T}
    ptr::Ptr{Cvoid}
    n::Int; m::Int; ntraj::Int; K::Int
    integrator::Cint
    model::Any            # keeps the model handle alive
end
function DeviceTrajectory(dmodel::RD.DiscretizedDynamics{L,Q}, ::Type{T}, ntraj::Integer, K::Integer) where {L,Q,T}
    h = gethandle(dmodel.continuous_dynamics)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:rdb_trajectory_create, LIB), Cint, (Ptr{Cvoid}, Cint, Int64, Cint, Ref{Ptr{Cvoid}}), h.ptr, dtypecode(T), ntraj, K, r),
          "rdb_trajectory_create")
    n, m = RD.dims(dmodel)
    finalizer(t -> ccall((:rdb_trajectory_destroy, LIB), Cint, (Ptr{Cvoid},), t.ptr), DeviceTrajectory{T}(r[], n, m, ntraj, K, INTEGRATOR[Q], h))
end
function DeviceTrajectory(dmodel::RD.DiscretizedDynamics, Z::RD.SampledTrajectory{n,m,T}) where {n,m,T}
    D = DeviceTrajectory(dmodel, T, 1, length(Z))
    data, dts, ts = gather(Z)
    RD.setstates!(D, data[1:n, :]); RD.setcontrols!(D, data[n+1:end, :])
    check(ccall((:rdb_trajectory_set_timesteps, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Cdouble, Cdouble, Ptr{Cvoid}), D.ptr, dts, 0.0, ts[1], C_NULL),
          "rdb_trajectory_set_timesteps")
    D
end
"setstates!(Z, X): X is (n, ntraj, K) — or (n, K) for one trajectory — host Array or device pointer holder (src/trajectories.jl:215-230)"
function RD.setstates!(D::DeviceTrajectory{T}, X::Array{T}) where {T}
    @assert length(X) == D.n * D.ntraj * D.K
    check(ccall((:rdb_trajectory_set_states, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), D.ptr, X, C_NULL), "rdb_trajectory_set_states")
end
"setcontrols!(Z, U): U is (m, ntraj, K-1) or (m, ntraj, K) (src/trajectories.jl:232-250); without a terminal control it is zero"
function RD.setcontrols!(D::DeviceTrajectory{T}, U::Array{T}) where {T}
    knots = length(U) ÷ (D.m * D.ntraj)
    check(ccall((:rdb_trajectory_set_controls, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Cvoid}), D.ptr, U, knots, C_NULL), "rdb_trajectory_set_controls")
end
function settimesteps!(D::DeviceTrajectory, dt::Real; t0::Real=0.0)
    check(ccall((:rdb_trajectory_set_timesteps, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Cdouble, Cdouble, Ptr{Cvoid}), D.ptr, C_NULL, dt, t0, C_NULL),
          "rdb_trajectory_set_timesteps")
end
function RD.states(D::DeviceTrajectory{T}) where {T}
    X = Array{T,3}(undef, D.n, D.ntraj, D.K)
    check(ccall((:rdb_trajectory_get_states, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), D.ptr, X, C_NULL), "rdb_trajectory_get_states")
    X
end
"rollout!(sig, dmodel, Z, x0) on the device (src/trajectories.jl:436-441): x0 is (n, ntraj)"
function RD.rollout!(D::DeviceTrajectory{T}, x0::Array{T}) where {T}
    check(ccall((:rdb_trajectory_set_initial_state, LIB), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), D.ptr, x0, C_NULL), "rdb_trajectory_set_initial_state")
    check(ccall((:rdb_trajectory_rollout, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}), D.ptr, D.integrator, C_NULL), "rdb_trajectory_rollout")
end
"""
    linearize!(D, J; error_state=false, rollout=false)

`J[:, :, j, k]` = discrete Jacobian of knot k of trajectory j (`n×(n+m)`, or the error-state `n̄×(n̄+m)`), a host `Array{T,4}`:
the kernels read the device mirror in place, only J crosses PCIe.  `rollout=true` first rolls the states out from knot 1 (pipelined).
"""
function linearize!(D::DeviceTrajectory{T}, J::Array{T,4}; error_state::Bool=false, rollout::Bool=false) where {T}
    rc = rollout ?
        ccall((:rdb_trajectory_rollout_linearize, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Cvoid}, Ptr{Cvoid}), D.ptr, D.integrator, error_state, 0, J, C_NULL) :
        ccall((:rdb_trajectory_linearize, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), D.ptr, D.integrator, error_state, J, C_NULL, C_NULL)
    check(rc, "rdb_trajectory_linearize")
end

end # module
