# runtests.jl — test-suite of the Julia binding (julia/RobotDynamicsB200.jl), mirroring the reference's own tests for the hot path:
#   test/cartpole_test.jl:40-72      jacobian! into a Matrix and a DynamicsJacobian, every signature, discretized model
#   test/integration_tests.jl:7-18   all Jacobian paths agree to 1e-10
#   test/rigid_body_jacobians.jl     Quadrotor discrete Jacobian, errstate_jacobian!, state_diff
# with the B200 diff method in place of ForwardAD, plus the batched / device-trajectory forms the binding adds.
#
#     RD_REF=/path/to/RobotDynamics.jl julia --project=oracle/ref_julia julia/runtests.jl        (on a machine with a B200 and Julia)
#
# NOT executed in this repository's build image (no Julia there; see DESIGN.md §3).  The Python mirror of the same interface
# (robotdynamics.jl_b200/api.py) runs the equivalent assertions in tests/test_gpu_parity.py::test_reference_api_*.
using Test
using RobotDynamics, Rotations, StaticArrays, ForwardDiff, FiniteDiff, LinearAlgebra, Random
const RD = RobotDynamics
const REF = get(ENV, "RD_REF", pkgdir(RobotDynamics))
include(joinpath(REF, "test", "cartpole_model.jl"))
include(joinpath(REF, "test", "quadrotor.jl"))
include(joinpath(@__DIR__, "RobotDynamicsB200.jl"))
using .RobotDynamicsB200: B200, RK2, DeviceTrajectory, linearize!, discrete_jacobian_batch!, discrete_error_jacobian_batch!,
                          errstate_jacobian_batch!, state_diff_batch!, custom_handle
const B = RobotDynamicsB200

Random.seed!(1)

@testset "Cartpole: jacobian!(sig, B200(), ...) == ForwardAD (test/cartpole_test.jl:40-72)" begin
    model = Cartpole()
    n, m = RD.dims(model)
    x, u = rand(model)
    t, dt = 0.0, 0.1
    z = RD.KnotPoint(x, u, t, dt)
    for Q in (RD.Euler, RK2, RD.RK3, RD.RK4)
        dmodel = RD.DiscretizedDynamics{Q}(model)
        F, F0, y, y0 = zeros(n, n + m), zeros(n, n + m), zeros(n), zeros(n)
        if Q === RK2
            F0 .= ForwardDiff.jacobian(v -> RD.discrete_dynamics(dmodel, v[SVector{n}(1:n)], v[SVector{m}(n+1:n+m)], t, dt), RD.getdata(z))
        else
            RD.jacobian!(RD.StaticReturn(), RD.ForwardAD(), dmodel, F0, y0, z)
        end
        for sig in (RD.StaticReturn(), RD.InPlace())
            RD.jacobian!(sig, B200(), dmodel, F, y, z)
            @test F ≈ F0 atol = 1e-10
            @test y ≈ RD.discrete_dynamics(dmodel, z) atol = 1e-12
        end
        D = RD.DynamicsJacobian(n, m)                      # any AbstractMatrix, incl. DynamicsJacobian (test/cartpole_test.jl:48-51)
        RD.jacobian!(RD.StaticReturn(), B200(), dmodel, D, y, z)
        @test D.A ≈ F0[:, 1:n] atol = 1e-10
        @test D.B ≈ F0[:, n .+ (1:m)] atol = 1e-10
    end
    # v0.3 spelling kept by BASELINE.json (README.md:81-82)
    ∇f = zeros(n, n + m)
    B.discrete_jacobian!(RD.RK4, ∇f, model, z)
    F0 = zeros(n, n + m)
    RD.jacobian!(RD.StaticReturn(), RD.ForwardAD(), RD.DiscretizedDynamics{RD.RK4}(model), F0, zeros(n), z)
    @test ∇f ≈ F0 atol = 1e-10
end

@testset "Batched: a whole SampledTrajectory in one call" begin
    model = Cartpole()
    dmodel = RD.DiscretizedDynamics{RD.RK4}(model)
    n, m = RD.dims(model)
    N = 101
    X = [@SVector rand(n) for _ in 1:N]
    U = [@SVector rand(m) for _ in 1:N-1]
    Z = RD.SampledTrajectory(X, U, dt=0.01)
    J, Y = zeros(n, n + m, N), zeros(n, N)
    RD.jacobian!(RD.StaticReturn(), B200(), dmodel, J, Y, Z)
    Jk, yk = zeros(n, n + m), zeros(n)
    for k in 1:N                                            # the loop the batched call replaces (src/discretized_dynamics.jl:129-136)
        RD.jacobian!(RD.StaticReturn(), RD.ForwardAD(), dmodel, Jk, yk, Z[k])
        @test J[:, :, k] ≈ Jk atol = 1e-10
    end
    @test J[:, :, N] ≈ [I zeros(n, m)]                     # terminal knot: dt = 0 (src/knotpoint.jl:57-67)
end

@testset "Quadrotor{QuatRotation}: Jacobian, LieState maps, error-state form (test/rigid_body_jacobians.jl)" begin
    model = Quadrotor()
    dmodel = RD.DiscretizedDynamics{RD.RK4}(model)
    n, m = RD.dims(model)
    N = 64
    data = zeros(n + m, N); X0 = zeros(n, N)
    for k in 1:N
        x, u = rand(model); data[:, k] .= [x; u]; X0[:, k] .= rand(model)[1]
    end
    dts = fill(0.05, N)
    J, Y = zeros(n, n + m, N), zeros(n, N)
    discrete_jacobian_batch!(dmodel, J, Y, data, dts)
    G = zeros(n, 12, N); errstate_jacobian_batch!(model, G, data[1:n, :])
    dX = zeros(12, N); state_diff_batch!(model, dX, data[1:n, :], X0)
    Jbar = zeros(12, 12 + m, N); discrete_error_jacobian_batch!(dmodel, Jbar, nothing, data, dts)
    Jk, yk, Gk, Gn = zeros(n, n + m), zeros(n), zeros(n, 12), zeros(n, 12)
    for k in 1:N
        z = RD.KnotPoint(SVector{n}(data[1:n, k]), SVector{m}(data[n+1:end, k]), 0.0, 0.05)
        RD.jacobian!(RD.StaticReturn(), RD.ForwardAD(), dmodel, Jk, yk, z)
        @test J[:, :, k] ≈ Jk atol = 1e-10
        @test Y[:, k] ≈ RD.discrete_dynamics(dmodel, z) atol = 1e-12
        Gk .= 0; RD.errstate_jacobian!(model, Gk, RD.state(z))
        @test G[:, :, k] ≈ Gk atol = 1e-12
        @test dX[:, k] ≈ RD.state_diff(model, RD.state(z), SVector{n}(X0[:, k])) atol = 1e-12
        Gn .= 0; RD.errstate_jacobian!(model, Gn, SVector{n}(Y[:, k]))
        @test Jbar[:, :, k] ≈ [Gn' * Jk[:, 1:n] * Gk  Gn' * Jk[:, n+1:end]] atol = 1e-10
    end
end

@testset "Device trajectory: upload once, update controls, linearize into host arrays" begin
    model = Cartpole()
    dmodel = RD.DiscretizedDynamics{RD.RK4}(model)
    n, m = RD.dims(model)
    K = 51
    Z = RD.SampledTrajectory([@SVector rand(n) for _ in 1:K], [@SVector rand(m) for _ in 1:K-1], dt=0.02)
    D = DeviceTrajectory(dmodel, Z)
    J = zeros(n, n + m, 1, K)
    linearize!(D, J)
    Jk, yk = zeros(n, n + m), zeros(n)
    for k in 1:K
        RD.jacobian!(RD.StaticReturn(), RD.ForwardAD(), dmodel, Jk, yk, Z[k])
        @test J[:, :, 1, k] ≈ Jk atol = 1e-10
    end
    # forward pass on the device == rollout! on the host (src/trajectories.jl:436-441)
    x0 = RD.state(Z[1])
    RD.rollout!(RD.StaticReturn(), dmodel, Z, x0)
    RD.rollout!(D, reshape(Vector(x0), n, 1))
    Xd = RD.states(D)
    for k in 1:K
        @test Xd[:, 1, k] ≈ RD.state(Z[k]) atol = 1e-10
    end
end

@testset "User model with time-varying dynamics (dynamics(model, x, u, t), src/dynamics.jl:81-83)" begin
    h = custom_handle(2, 1, "return vec(get<1>(x), cos_(T(3) * t) * get<0>(u) - p[0] * sin_(get<0>(x)) + t * get<1>(x));"; params=[1.7])
    @test h.ptr != C_NULL
end

struct NoKernel <: RD.ContinuousDynamics end          # (type definitions must be at top level)
@testset "Errors map to the reference's exceptions" begin
    @test_throws RD.NotImplementedError B.handle(NoKernel())
end
