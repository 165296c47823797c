# RobotDynamicsB200.jl — the reference-side binding of librdb200.so (include/rdb200.h).
#
# NOT executed in this repository's image (Julia is not installed there); this is the stub a RobotDynamics.jl
# maintainer would add.  It introduces one new DiffMethod subtype, `B200`, the extension point the reference documents
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

Describe a reference model to the library.  One method per model family; parameter packing as in include/rdb200.h.
Shown for the models the reference's tests/examples define (test/cartpole_model.jl, test/quadrotor.jl,
examples/single_satellite.jl); a generic `RigidBody{R}` with a user `forces/moments` is out of scope (SURVEY §8f row 4).
"""
handle(m) = throw(RD.NotImplementedError("no B200 kernel for $(typeof(m))"))
# handle(m::Cartpole)     = Handle(KIND_CARTPOLE, Cint(0), Cint(0), [m.mc, m.mp, m.l, m.g])
# handle(m::Quadrotor{R}) where R = Handle(KIND_QUADROTOR, rotcode(R), framecode(m),
#       [m.mass; vec(Matrix(m.J)'); m.gravity; m.motor_dist; m.kf; m.km])
# handle(m::Satellite{R}) where R = Handle(KIND_BODY, rotcode(R), framecode(m), [m.mass; vec(Matrix(m.J)')])

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

const INTEGRATOR = Dict{Any,Cint}(RD.Euler => 0, RD.RK3 => 2, RD.RK4 => 3, RD.ImplicitMidpoint => 4)   # RK2 = 1: the v0.3 name, add when it returns to src/
dtypecode(::Type{Float32}) = Cint(0); dtypecode(::Type{Float64}) = Cint(1)

# ---- gather / scatter between the reference containers and the batched images ---------------------------------------
"Z as Matrix{T}(n+m, N): column k is `getdata(Z[k])` (src/knotpoint.jl:196); also the Float64 dt vector."
function gather(Z::RD.SampledTrajectory{n,m,T}) where {n,m,T}
    N = length(Z)
    data = Matrix{T}(undef, n + m, N)
    dts = Vector{Float64}(undef, N)
    @inbounds for k in 1:N
        data[:, k] .= RD.getdata(Z[k])
        dts[k] = RD.timestep(Z[k])
    end
    data, dts
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
    data, dts = gather(Z)
    discrete_jacobian_batch!(dmodel, J, y, data, dts)
end

"Same, on pre-gathered (or device-resident CuArray) data: no per-knot host work at all."
function discrete_jacobian_batch!(dmodel::RD.DiscretizedDynamics{L,Q}, J, y, data, dts; stream=C_NULL) where {L,Q}
    T = eltype(data)
    N = size(data, 2)
    h = gethandle(dmodel.continuous_dynamics)
    check(ccall((:rdb_discrete_jacobian, LIB), Cint,
                (Ptr{Cvoid}, Cint, Cint, Cint, Int64, Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}),
                h.ptr, INTEGRATOR[Q], dtypecode(T), 0 #= RDB_AOS =#, N, pointer(data), C_NULL, pointer(dts), 0.0,
                pointer(J), y === nothing ? C_NULL : pointer(y), stream), "rdb_discrete_jacobian")
    nothing
end

"Single knot point, for API fidelity only (one kernel launch per call — use the batched method in loops)."
function RD.jacobian!(sig::RD.FunctionSignature, ::B200, dmodel::RD.DiscretizedDynamics, J, y, z::RD.AbstractKnotPoint)
    n, m = RD.dims(dmodel)
    data = reshape(Vector(RD.getdata(z)), n + m, 1)
    Jb = Array{eltype(data),3}(undef, n, n + m, 1); yb = Matrix{eltype(data)}(undef, n, 1)
    discrete_jacobian_batch!(dmodel, Jb, yb, data, [RD.timestep(z)])
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

end # module
