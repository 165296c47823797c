# bench_ref.jl — the reference arm of bench.py when the box has Julia: times RobotDynamics.jl's own
#     jacobian!(StaticReturn(), ForwardAD(), DiscretizedDynamics{Q}(model), J, y, z)          (src/jacobian_gen.jl:485-507)
# in the caller's `for k in 1:N` loop (src/discretized_dynamics.jl:129-136) over StaticKnotPoints, Threads.@threads over the knot
# points with one `copy(dmodel)` per thread (test/integration_tests.jl:73-77), and prints ONE JSON line for bench.py to parse.
#
#     julia --project=oracle/ref_julia -t auto oracle/ref_julia/bench_ref.jl <cartpole|quadrotor|satellite> <steps> <warmup>
#
# TEST / BASELINE INFRASTRUCTURE ONLY; not executed in this repository's build image (no Julia there).
import Pkg
haskey(ENV, "RD_REF") && Pkg.develop(path=ENV["RD_REF"])
using RobotDynamics, Rotations, StaticArrays, ForwardDiff, FiniteDiff, LinearAlgebra, Random   # (@autodiff-generated methods name ForwardDiff / FiniteDiff in the caller's module)
const RD = RobotDynamics
const REF = get(ENV, "RD_REF", pkgdir(RobotDynamics))
include(joinpath(REF, "test", "cartpole_model.jl"))
include(joinpath(REF, "test", "quadrotor.jl"))

RD.@autodiff struct BenchSatellite{R} <: RD.RigidBody{R}
    mass::Float64
    J::Diagonal{Float64,SVector{3,Float64}}
end
RD.control_dim(::BenchSatellite) = 6
RD.mass(model::BenchSatellite) = model.mass
RD.inertia(model::BenchSatellite) = model.J
RD.forces(model::BenchSatellite, x::StaticVector, u::StaticVector) = RD.orientation(model, x) * SA[u[1], u[2], u[3]]
RD.moments(model::BenchSatellite, x::StaticVector, u::StaticVector) = SA[u[4], u[5], u[6]]

name, steps, warmup = ARGS[1], parse(Int, ARGS[2]), parse(Int, ARGS[3])
rk2(model, x, u, t, h) = x + h * RD.dynamics(model, x + (h / 2) * RD.dynamics(model, x, u, t), u, t + h / 2)

model, dt, Ntotal = name == "cartpole" ? (Cartpole(), 0.01, 1 << 20) : name == "quadrotor" ? (Quadrotor(), 0.01, 262144) :
                    (BenchSatellite{MRP{Float64}}(1.0, Diagonal(@SVector ones(3))), 0.1, 1 << 20)
n, m = RD.dims(model)
Random.seed!(100)

function one_pass!(Js, zs, dmodels, model, dt, rk2mode)
    Threads.@threads for k in eachindex(zs)
        dm = dmodels[Threads.threadid()]
        if rk2mode
            z = zs[k]
            ix, iu = SVector{length(RD.state(z))}(1:length(RD.state(z))), SVector{length(RD.control(z))}(length(RD.state(z)) .+ (1:length(RD.control(z))))
            Js[k] .= ForwardDiff.jacobian(v -> rk2(model, v[ix], v[iu], 0.0, dt), RD.getdata(z))
        else
            RD.jacobian!(RD.StaticReturn(), RD.ForwardAD(), dm, Js[k], zeros(length(RD.state(zs[k]))), zs[k])
        end
    end
end

rk2mode = name == "satellite"
dmodel = RD.DiscretizedDynamics{RD.RK4}(model)
dmodels = [copy(dmodel) for _ in 1:Threads.nthreads()]
# bounded sample: size one step so that the whole run stays within a few minutes
probe = [RD.KnotPoint(rand(model)..., 0.0, dt) for _ in 1:4096]
Jp = [zeros(n, n + m) for _ in 1:4096]
one_pass!(Jp, probe, dmodels, model, dt, rk2mode)
tprobe = @elapsed one_pass!(Jp, probe, dmodels, model, dt, rk2mode)
per_step = clamp(round(Int, 4096 / tprobe * min(1.0, 100.0 / max(1, steps + warmup))), 4096, Ntotal)
zs = [RD.KnotPoint(rand(model)..., 0.0, dt) for _ in 1:per_step]
Js = [zeros(n, n + m) for _ in 1:per_step]
for _ in 1:warmup
    one_pass!(Js, zs, dmodels, model, dt, rk2mode)
end
el = @elapsed for _ in 1:steps
    one_pass!(Js, zs, dmodels, model, dt, rk2mode)
end
println("{\"value\": $(per_step * steps / el), \"ms_per_step\": $(el / steps * 1e3), \"threads\": $(Threads.nthreads()), ",
        "\"sample\": \"each step = $(per_step) of the workload's $(Ntotal) knot points, RobotDynamics.jl v0.4.8 jacobian!(StaticReturn(), ForwardAD(), ...) per knot, Threads.@threads\"}")
