# gen_fixtures.jl — pins the CPU oracle (oracle/rd_oracle.cpp) and the CUDA path to the REAL reference: runs RobotDynamics.jl
# v0.4.8 (with Rotations.jl 1.x, ForwardDiff.jl 0.10) on the committed seeded inputs tests/golden/julia_in/*.npy and writes the
# reference's own outputs to tests/golden/julia_out/*.npy, which tests/test_julia_fixtures.py consumes automatically when present.
#
#     julia --project=oracle/ref_julia -e 'using Pkg; Pkg.instantiate()'          # once (network), or:
#     RD_REF=/path/to/RobotDynamics.jl julia --project=oracle/ref_julia oracle/ref_julia/gen_fixtures.jl
#     julia --project=oracle/ref_julia oracle/ref_julia/gen_fixtures.jl
#
# TEST INFRASTRUCTURE ONLY.  Julia is not installed in the build image of this repository (probed: `julia` not found, no network),
# so this script has NOT been executed there; until somebody runs it the oracle's parity status stays "unpinned against Julia output"
# (DESIGN.md §3).  It deliberately uses only the reference's public API and the model definitions of the reference's own test suite
# (test/cartpole_model.jl, test/quadrotor.jl); the Body / Satellite models restate test/rigidbody_test.jl:23-56 and
# examples/single_satellite.jl:7-35 (those files cannot be included without running their tests / benchmarks).
#
# What is pinned (each case is independent: a failure is reported and the others still run):
#   c1/c2    Cartpole RK3 / RK4 discrete Jacobians   jacobian!(StaticReturn(), ForwardAD(), DiscretizedDynamics{Q}(model), J, y, z)
#            (src/jacobian_gen.jl:485-507) + the UserDefined chain rule (src/integration.jl:149-177,302-337) + x⁺
#   c3       Quadrotor{QuatRotation} RK4 (dt = 0.01, 0.1), errstate_jacobian!, ∇errstate_jacobian!, state_diff
#   c4       Satellite RigidBody{MRP} RK2 (explicit midpoint; not in src/ at v0.4.8, written out below from dynamics())
#   quad_offmanifold   continuous dynamics + Jacobian with NON-unit state quaternions: fixes the off-manifold form of q*r, q\r
#   quad_tie           max(0, kf w) at exact ties / negative controls (ForwardDiff's derivative at the kink)
#   quad_{rot}_{frame}, body_{rot}_{frame}   every RigidBody{R} × velocity frame, RK4
#   lie_{rot}          G = errstate_jacobian!, ∇G = ∇errstate_jacobian!, state_diff for QuatRotation / MRP / RodriguesParam
#                      (Rotations.∇differential / ∇²differential / rotation_error(CayleyMap) at the MRP / RP boundary)
#   implicit_*         DiscretizedDynamics{ImplicitMidpoint}: x⁺ and the implicit-function-theorem Jacobian

import Pkg
if haskey(ENV, "RD_REF")
    Pkg.develop(path=ENV["RD_REF"])
end

using RobotDynamics, Rotations, StaticArrays, ForwardDiff, FiniteDiff, LinearAlgebra   # (@autodiff-generated methods name ForwardDiff / FiniteDiff in the caller's module)
const RD = RobotDynamics

const REPO = normpath(joinpath(@__DIR__, "..", ".."))
const IN = joinpath(REPO, "tests", "golden", "julia_in")
const OUT = joinpath(REPO, "tests", "golden", "julia_out")
const REF = get(ENV, "RD_REF", pkgdir(RobotDynamics))

include(joinpath(REF, "test", "cartpole_model.jl"))     # Cartpole + its analytic jacobian!
include(joinpath(REF, "test", "quadrotor.jl"))          # Quadrotor{R}

# ---- .npy I/O without packages (C-order (N, a, b) on disk == Julia column-major (b, a, N)) ----------------------------------------
function read_npy(path)
    open(path) do io
        magic = read(io, 6)
        @assert magic == UInt8[0x93, UInt8('N'), UInt8('U'), UInt8('M'), UInt8('P'), UInt8('Y')] "not an .npy file: $path"
        major = read(io, UInt8); read(io, UInt8)
        hlen = major == 1 ? Int(read(io, UInt16)) : Int(read(io, UInt32))
        header = String(read(io, hlen))
        descr = match(r"'descr':\s*'([^']+)'", header).captures[1]
        @assert !occursin(r"'fortran_order':\s*True", header) "fortran_order arrays are not supported"
        shp = match(r"'shape':\s*\(([^)]*)\)", header).captures[1]
        dims = [parse(Int, strip(s)) for s in split(shp, ",") if !isempty(strip(s))]
        T = descr == "<f8" ? Float64 : descr == "<f4" ? Float32 : error("unsupported dtype $descr")
        data = Array{T}(undef, reverse(dims)...)
        read!(io, data)
        data
    end
end
function write_npy(path, A::Array{T}) where {T<:Union{Float32,Float64}}
    descr = T == Float64 ? "<f8" : "<f4"
    dims = reverse(size(A))
    shape = join(dims, ", ") * (length(dims) == 1 ? "," : "")
    header = "{'descr': '$descr', 'fortran_order': False, 'shape': ($shape), }"
    pad = 64 - mod(10 + length(header) + 1, 64)
    header = header * " "^pad * "\n"
    open(path, "w") do io
        write(io, UInt8(0x93)); write(io, "NUMPY"); write(io, UInt8(1)); write(io, UInt8(0))
        write(io, UInt16(length(header))); write(io, header); write(io, A)
    end
end
input(name) = read_npy(joinpath(IN, name * ".npy"))
output(name, A) = write_npy(joinpath(OUT, name * ".npy"), Array{Float64}(A))

# ---- models that the reference defines only inside test / example scripts ----------------------------------------------------------
# test/rigidbody_test.jl:23-56 (mass 2, J = diag(2,3,1)); the velocity frame is a field here so that both frames can be generated
RD.@autodiff struct PinBody{R} <: RD.RigidBody{R}
    bodyframe::Bool
end
RD.control_dim(::PinBody) = 6
function RD.wrenches(model::PinBody, x::StaticVector, u::StaticVector, t)
    q = RD.orientation(model, x)
    F = q * SA[u[1], u[2], u[3]]
    SA[F[1], F[2], F[3], u[4], u[5], u[6]]
end
RD.mass(::PinBody) = 2.0
RD.inertia(::PinBody) = Diagonal(SA[2, 3, 1.0])
RD.velocity_frame(model::PinBody) = model.bodyframe ? :body : :world

# examples/single_satellite.jl:7-35 (mass 1, J = I)
RD.@autodiff struct PinSatellite{R} <: RD.RigidBody{R}
    mass::Float64
    J::Diagonal{Float64,SVector{3,Float64}}
end
RD.control_dim(::PinSatellite) = 6
RD.mass(model::PinSatellite) = model.mass
RD.inertia(model::PinSatellite) = model.J
RD.forces(model::PinSatellite, x::StaticVector, u::StaticVector) = RD.orientation(model, x) * SA[u[1], u[2], u[3]]
RD.moments(model::PinSatellite, x::StaticVector, u::StaticVector) = SA[u[4], u[5], u[6]]

# ---- helpers ----------------------------------------------------------------------------------------------------------------------
rottype(rot) = rot == "quat" ? QuatRotation{Float64} : rot == "mrp" ? MRP{Float64} : RodriguesParam{Float64}

"explicit midpoint (the v0.3 `RK2`, test/old_tests/linear_tests.jl:135-141): x + h f(x + h/2 f(x,u,t), u, t + h/2)"
rk2(model, x, u, t, h) = x + h * RD.dynamics(model, x + (h / 2) * RD.dynamics(model, x, u, t), u, t + h / 2)

"J (n, n+m, N) and x⁺ (n, N) of a discretized model over the columns of Z, through the reference's own jacobian! / discrete_dynamics"
function discrete_case(dmodel, Z, dt; sig=RD.StaticReturn(), diff=RD.ForwardAD())
    n, m = RD.dims(dmodel)
    N = size(Z, 2)
    J = zeros(n, n + m, N); Xn = zeros(n, N)
    Jk = zeros(n, n + m); y = zeros(n)
    for k in 1:N
        z = RD.KnotPoint(SVector{n}(Z[1:n, k]), SVector{m}(Z[n+1:n+m, k]), 0.0, dt)
        RD.jacobian!(sig, diff, dmodel, Jk, y, z)
        J[:, :, k] .= Jk
        Xn[:, k] .= RD.discrete_dynamics(dmodel, z)
    end
    J, Xn
end

"the same Jacobian straight from ForwardDiff over [x;u] (what the generated method does, src/jacobian_gen.jl:485-507)"
function forwarddiff_case(step, n, m, Z, dt)
    N = size(Z, 2)
    ix, iu = SVector{n}(1:n), SVector{m}(n+1:n+m)
    J = zeros(n, n + m, N); Xn = zeros(n, N)
    for k in 1:N
        z = SVector{n + m}(Z[:, k])
        J[:, :, k] .= ForwardDiff.jacobian(v -> step(v[ix], v[iu], 0.0, dt), z)
        Xn[:, k] .= step(z[ix], z[iu], 0.0, dt)
    end
    J, Xn
end

function continuous_case(model, Z)
    n, m = RD.dims(model)
    N = size(Z, 2)
    J = zeros(n, n + m, N); Xd = zeros(n, N)
    Jk = zeros(n, n + m); y = zeros(n)
    for k in 1:N
        z = RD.KnotPoint(SVector{n}(Z[1:n, k]), SVector{m}(Z[n+1:n+m, k]), 0.0, 0.0)
        RD.jacobian!(RD.StaticReturn(), RD.ForwardAD(), model, Jk, y, z)
        J[:, :, k] .= Jk
        Xd[:, k] .= RD.dynamics(model, z)
    end
    J, Xd
end

function lie_case(model, X, X0)
    n = RD.state_dim(model); ne = RD.errstate_dim(model)
    N = size(X, 2)
    G = zeros(n, ne, N); dG = zeros(ne, ne, N); dX = zeros(ne, N)
    for k in 1:N
        x, x0 = SVector{n}(X[1:n, k]), SVector{n}(X0[1:n, k])
        Gk = zeros(n, ne); RD.errstate_jacobian!(model, Gk, x); G[:, :, k] .= Gk
        Hk = zeros(ne, ne); RD.∇errstate_jacobian!(model, Hk, x, x0); dG[:, :, k] .= Hk
        dX[:, k] .= RD.state_diff(model, x, x0)
    end
    G, dG, dX
end

failures = String[]
macro case(name, body)
    quote
        try
            $(esc(body))
            println("ok      ", $(esc(name)))
        catch err
            push!(failures, $(esc(name)))
            println("FAILED  ", $(esc(name)), ": ", sprint(showerror, err))
        end
    end
end

mkpath(OUT)

# ---- C1 / C2: Cartpole -------------------------------------------------------------------------------------------------------------
for (name, Q, dt) in (("c1_cartpole", RD.RK3, 0.01), ("c2_cartpole", RD.RK4, 0.01))
    @case name begin
        dmodel = RD.DiscretizedDynamics{Q}(Cartpole())
        Z = input(name * "_Z")
        J, Xn = discrete_case(dmodel, Z, dt)
        Jc, _ = discrete_case(dmodel, Z, dt; diff=RD.UserDefined())          # chain rule over the analytic continuous Jacobian
        Ji, _ = discrete_case(dmodel, Z, dt; sig=RD.InPlace())
        @assert maximum(abs.(J - Jc)) < 1e-10 && maximum(abs.(J - Ji)) < 1e-10     # test/integration_tests.jl:7-18
        output(name * "_J", J); output(name * "_J_userdefined", Jc); output(name * "_xn", Xn)
    end
end

# ---- C3: Quadrotor{QuatRotation} RK4 + LieState maps ------------------------------------------------------------------------------
for (name, dt) in (("c3_quadrotor", 0.01), ("c3_quadrotor_dt01", 0.1))
    @case name begin
        dmodel = RD.DiscretizedDynamics{RD.RK4}(Quadrotor())
        J, Xn = discrete_case(dmodel, input(name * "_Z"), dt)
        output(name * "_J", J); output(name * "_xn", Xn)
    end
end

# ---- C4: Satellite RigidBody{MRP}, RK2 ----------------------------------------------------------------------------------------------
@case "c4_satellite_mrp" begin
    model = PinSatellite{MRP{Float64}}(1.0, Diagonal(@SVector ones(3)))
    J, Xn = forwarddiff_case((x, u, t, h) -> rk2(model, x, u, t, h), 12, 6, input("c4_satellite_mrp_Z"), 0.1)
    output("c4_satellite_mrp_J", J); output("c4_satellite_mrp_xn", Xn)
end

# ---- off-manifold quaternions and the thrust clamp: continuous dynamics + Jacobian ----------------------------------------------------
for name in ("quad_offmanifold", "quad_tie")
    @case name begin
        J, Xd = continuous_case(Quadrotor(), input(name * "_Z"))
        output(name * "_J", J); output(name * "_xdot", Xd)
        Jd, Xn = discrete_case(RD.DiscretizedDynamics{RD.RK4}(Quadrotor()), input(name * "_Z"), 0.05)
        output(name * "_Jd", Jd); output(name * "_xn", Xn)
    end
end

# ---- every RigidBody{R} × velocity frame ---------------------------------------------------------------------------------------------
for rot in ("quat", "mrp", "rp"), (frame, bf) in (("world", false), ("body", true))
    R = rottype(rot)
    @case "quad_$(rot)_$(frame)" begin
        dmodel = RD.DiscretizedDynamics{RD.RK4}(Quadrotor{R}(bodyframe=bf))
        J, Xn = discrete_case(dmodel, input("quad_$(rot)_Z"), 0.05)
        output("quad_$(rot)_$(frame)_J", J); output("quad_$(rot)_$(frame)_xn", Xn)
    end
    @case "body_$(rot)_$(frame)" begin
        dmodel = RD.DiscretizedDynamics{RD.RK4}(PinBody{R}(bf))
        J, Xn = discrete_case(dmodel, input("body_$(rot)_Z"), 0.05)
        output("body_$(rot)_$(frame)_J", J); output("body_$(rot)_$(frame)_xn", Xn)
    end
end

# ---- LieState maps at the Rotations.jl boundary ----------------------------------------------------------------------------------------
for rot in ("quat", "mrp", "rp")
    @case "lie_$(rot)" begin
        model = Quadrotor{rottype(rot)}()
        G, dG, dX = lie_case(model, input("quad_$(rot)_Z"), input("lie_$(rot)_X0"))
        output("lie_$(rot)_G", G); output("lie_$(rot)_dG", dG); output("lie_$(rot)_dx", dX)
    end
end

# ---- ImplicitMidpoint ----------------------------------------------------------------------------------------------------------------------
for (name, mk) in (("implicit_cartpole", () -> Cartpole()), ("implicit_quadrotor", () -> Quadrotor()))
    @case name begin
        dmodel = RD.DiscretizedDynamics{RD.ImplicitMidpoint}(mk())
        J, Xn = discrete_case(dmodel, input(name * "_Z"), 0.05; sig=RD.InPlace())
        output(name * "_J", J); output(name * "_xn", Xn)
    end
end

# ---- provenance -----------------------------------------------------------------------------------------------------------------------------
open(joinpath(OUT, "VERSIONS.txt"), "w") do io
    println(io, "julia ", VERSION)
    for (uuid, info) in Pkg.dependencies()
        info.name in ("RobotDynamics", "Rotations", "ForwardDiff", "StaticArrays", "Quaternions") && println(io, info.name, " ", info.version)
    end
    println(io, "failures: ", isempty(failures) ? "none" : join(failures, ", "))
end
println(isempty(failures) ? "all cases written to $OUT" : "FAILED cases: $(join(failures, ", "))")
exit(isempty(failures) ? 0 : 1)
