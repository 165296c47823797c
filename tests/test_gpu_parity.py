"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI (librdb200.so via ctypes);
the CPU oracle is only the checker.  Tolerances are the north_star's: 1e-10 (fp64) / 1e-4 (fp32) max-abs."""
import numpy as np
import pytest

from oracle import rd_oracle as o
from common import TOL, golden, rand_inputs, zoo

pytestmark = pytest.mark.gpu

QS = (o.EULER, o.RK2, o.RK3, o.RK4)


@pytest.fixture(scope="module")
def rd():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import rdb200
    return rdb200


@pytest.fixture(scope="module")
def torch_():
    import torch
    return torch


def dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# ---- every model x integrator x dtype against the oracle (device pointers, reference layout) ------------------------------
@pytest.mark.parametrize("name", sorted(zoo()))
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_discrete_jacobian_all_models(rd, torch_, name, dtype):
    om, gm = zoo()[name][0](), zoo()[name][1](rd)
    assert (gm.n, gm.m, rd.errstate_dim(gm)) == (om.n, om.m, om.nerr)
    N = 777                                                   # several tiles + a ragged tail for every TILE in use
    Z = rand_inputs(om.n, om.m, N, np.random.default_rng(21)).astype(dtype)
    Z64 = Z.astype(np.float64)
    Zd = dev(torch_, Z)
    for Q in QS:
        for dt in (0.01, 0.1):
            xn = torch_.empty((N, om.n), dtype=Zd.dtype, device="cuda")
            J = gm._h.discrete_jacobian(Q, Zd, dt, xn=xn)
            torch_.cuda.synchronize()
            assert np.abs(J.cpu().numpy() - o.discrete_jacobian(om, Q, Z64, dt)).max() < TOL[dtype]
            assert np.abs(xn.cpu().numpy() - o.discrete_dynamics(om, Q, Z64, dt)).max() < TOL[dtype]
    # above the small-batch threshold (launch.cuh: RDB_SMALL_N = 8192) the rigid-body RK3 / RK4 Jacobians run on their wide tiles:
    # the same comparison there, every knot, ragged tail
    Nb = 8192 + 4133
    Zb = rand_inputs(om.n, om.m, Nb, np.random.default_rng(22)).astype(dtype)
    for Q in (o.RK3, o.RK4):
        Jb = gm._h.discrete_jacobian(Q, dev(torch_, Zb), 0.05)
        torch_.cuda.synchronize()
        assert np.abs(Jb.cpu().numpy() - o.discrete_jacobian(om, Q, Zb.astype(np.float64), 0.05)).max() < TOL[dtype]
    # continuous dynamics and Jacobian (dynamics / jacobian!)
    xd = torch_.empty((N, om.n), dtype=Zd.dtype, device="cuda")
    Jc = gm._h.jacobian(Zd, xdot=xd)
    f = gm._h.dynamics(Zd)
    xn2 = gm._h.discrete_dynamics(o.RK4, Zd, 0.05)
    torch_.cuda.synchronize()
    scale = max(1.0, np.abs(o.jacobian(om, Z64)).max())       # quadrotor d(wdot)/du ~ L/J ~ 76: relative for fp32
    assert np.abs(Jc.cpu().numpy() - o.jacobian(om, Z64)).max() < TOL[dtype] * scale
    fs = max(1.0, np.abs(o.dynamics(om, Z64)).max())
    assert np.abs(xd.cpu().numpy() - o.dynamics(om, Z64)).max() < TOL[dtype] * fs
    assert np.abs(f.cpu().numpy() - o.dynamics(om, Z64)).max() < TOL[dtype] * fs
    assert np.abs(xn2.cpu().numpy() - o.discrete_dynamics(om, o.RK4, Z64, 0.05)).max() < TOL[dtype]


# ---- layouts, pointer kinds, alignment, ragged sizes ----------------------------------------------------------------------------
@pytest.mark.parametrize("name,dtype", [("cartpole", np.float64), ("quad_quat_world", np.float32), ("satellite_mrp", np.float64),
                                        ("quad_mrp_world", np.float32)])   # the last two: padded image + 2-D tensor store, and its fallbacks
def test_layouts_and_pointer_kinds_agree_bitwise(rd, torch_, name, dtype):
    om, gm = zoo()[name][0](), zoo()[name][1](rd)
    n, nz = om.n, om.n + om.m
    for N in (1, 31, 64, 65, 1000, 70001):                    # 70001 > one host-pipeline chunk (65536)
        Z = rand_inputs(om.n, om.m, N, np.random.default_rng(N)).astype(dtype)
        dt = np.random.default_rng(N + 1).uniform(0.0, 0.1, N)
        ref = o.discrete_jacobian(om, o.RK4, Z.astype(np.float64), dt)
        Zd = dev(torch_, Z)
        J_dev = gm._h.discrete_jacobian(o.RK4, Zd, dt).cpu().numpy()
        assert np.abs(J_dev - ref).max() < TOL[dtype]
        # host pointers (pinned pipeline inside the library)
        xn_h = np.empty((N, n), dtype=dtype)
        J_host = gm._h.discrete_jacobian(o.RK4, Z, dt, xn=xn_h)
        assert np.array_equal(J_host, J_dev)
        assert np.abs(xn_h - o.discrete_dynamics(om, o.RK4, Z.astype(np.float64), dt)).max() < TOL[dtype]
        # component-major layout, device and host
        Zs = np.ascontiguousarray(Z.T)
        Js = gm._h.discrete_jacobian(o.RK4, dev(torch_, Zs), dt, layout=rd.SOA).cpu().numpy()
        assert Js.shape == (n * nz, N) and np.array_equal(Js.T.reshape(N, nz, n), J_dev)
        Jsh = gm._h.discrete_jacobian(o.RK4, Zs, dt, layout=rd.SOA)
        assert np.array_equal(Jsh, Js)
        # value-only operations in the component-major layout (tensor-map kernels when N * sizeof(T) is a multiple of 16,
        # transposes around the knot-major kernel otherwise: both must reproduce the knot-major results bit for bit)
        xs = gm._h.discrete_dynamics(o.RK4, dev(torch_, Zs), dt, layout=rd.SOA).cpu().numpy()
        assert xs.shape == (n, N) and np.array_equal(xs.T, gm._h.discrete_dynamics(o.RK4, Zd, dt).cpu().numpy())
        fs = gm._h.dynamics(dev(torch_, Zs), layout=rd.SOA).cpu().numpy()
        assert np.array_equal(fs.T, gm._h.dynamics(Zd).cpu().numpy())
        # deliberately mis-aligned device buffers (element offset 1): cooperative-copy path instead of TMA
        if N > 1:
            esz = np.dtype(dtype).itemsize
            raw = torch_.empty(N * nz + 1, dtype=Zd.dtype, device="cuda")
            raw[1:].copy_(Zd.reshape(-1))
            Zu = raw[1:].view(N, nz)
            Jraw = torch_.empty(N * nz * n + 1, dtype=Zd.dtype, device="cuda")
            Ju = Jraw[1:].view(N, nz, n)
            assert Zu.data_ptr() % 16 == esz % 16 or esz == 16
            gm._h.discrete_jacobian(o.RK4, Zu, dt, J=Ju)
            assert np.array_equal(Ju.cpu().numpy(), J_dev)


def test_empty_batch_and_argument_errors(rd, torch_):
    gm = rd.Cartpole()
    Z0 = np.empty((0, 5))
    assert gm._h.discrete_jacobian(o.RK4, Z0, 0.01).shape == (0, 5, 4)
    assert gm._h.discrete_jacobian(o.RK4, dev(torch_, Z0), 0.01).shape == (0, 5, 4)
    with pytest.raises(ValueError):
        gm._h.discrete_jacobian(o.RK4, np.zeros((3, 4)), 0.01)           # wrong width
    with pytest.raises(rd.RDBError) as e:                                # host Z with device J
        gm._h.discrete_jacobian(o.RK4, np.zeros((4, 5)), 0.01, J=torch_.empty((4, 5, 4), dtype=torch_.float64, device="cuda"))
    assert e.value.code == rd._abi.ERR_POINTER_MIX
    with pytest.raises(rd.RDBError):
        gm._h.discrete_jacobian(7, np.zeros((4, 5)), 0.01)               # unknown integrator
    with pytest.raises(TypeError):
        gm._h.discrete_jacobian(o.RK4, np.zeros((4, 5), dtype=np.float16), 0.01)
    with pytest.raises(rd.NotImplementedModelError):
        rd.DoubleIntegrator(4)
    # caller-supplied buffers are validated before their bare pointers reach the C ABI (a float32 or short J would otherwise be
    # an out-of-bounds write of N * 20 * sizeof(double) bytes)
    Z = np.random.default_rng(0).random((8, 5))
    with pytest.raises(TypeError):
        gm._h.discrete_jacobian(o.RK4, Z, 0.01, J=np.empty((8, 5, 4), dtype=np.float32))
    with pytest.raises(ValueError):
        gm._h.discrete_jacobian(o.RK4, Z, 0.01, J=np.empty((7, 5, 4)))
    with pytest.raises(ValueError):
        gm._h.discrete_jacobian(o.RK4, Z, np.full(7, 0.01))              # dt vector shorter than the batch
    with pytest.raises(ValueError):
        gm._h.discrete_jacobian(o.RK4, Z, 0.01, xn=np.empty((8, 5)))
    with pytest.raises(ValueError):
        gm._h.discrete_jacobian(o.RK4, Z, 0.01, J=np.empty((8, 4, 5)).transpose(0, 2, 1))     # right shape, not contiguous
    with pytest.raises(TypeError):
        gm._h.rollout(o.RK4, Z[:, :4].copy(), np.zeros((8, 3, 1), dtype=np.float32), 0.01)    # U dtype differs from x0
    with pytest.raises(ValueError):
        gm._h.rollout(o.RK4, Z[:, :4].copy(), np.zeros((8, 3, 1)), np.full((8, 3), 0.01))     # dt must be (ntraj, K)
    with pytest.raises(TypeError):                                       # one DynamicsJacobian cannot receive a trajectory's Jacobians
        rd.jacobian_(rd.StaticReturn(), rd.ForwardAD(), rd.DiscretizedDynamics(gm, rd.RK4), rd.DynamicsJacobian(4, 1), None,
                     rd.SampledTrajectory(Z[:, :4], Z[:, 4:], 0.01))


def test_terminal_knots_and_per_knot_dt(rd, torch_):
    """dt == 0 (terminal knot point, src/knotpoint.jl:57-67) gives J = [I 0] and x+ = x."""
    gm, om = rd.Quadrotor(), o.quadrotor()
    N = 300
    Z = rand_inputs(13, 4, N, np.random.default_rng(5))
    dt = np.full(N, 0.02); dt[::7] = 0.0
    xn = np.empty((N, 13))
    J = gm._h.discrete_jacobian(o.RK4, Z, dt, xn=xn)
    I0 = np.concatenate([np.eye(13), np.zeros((13, 4))], axis=1)
    assert np.array_equal(o.as_matrix(J)[::7], np.broadcast_to(I0, (len(dt[::7]), 13, 17)))
    assert np.array_equal(xn[::7], Z[::7, :13])
    assert np.abs(J - o.discrete_jacobian(om, o.RK4, Z, dt)).max() < 1e-10


def test_quadrotor_clamp_kink(rd):
    """max(0, kf w): zero derivative when clamped (w < 0), partials KEPT at the tie w == 0 (ForwardDiff / Base.max semantics);
    yaw moment unclamped (test/quadrotor.jl:67-70,86-95)."""
    gm, om = rd.Quadrotor(), o.quadrotor()
    Z = rand_inputs(13, 4, 64, np.random.default_rng(6))
    Z[:, 13] = -0.3; Z[:, 14] = 0.0
    for dtype in (np.float64, np.float32):
        J = gm._h.jacobian(Z.astype(dtype))
        Jm = o.as_matrix(J)
        assert np.all(Jm[:, 7:10, 13] == 0) and np.all(Jm[:, 10:12, 13] == 0)
        assert np.any(Jm[:, 7:10, 14] != 0) and np.allclose(Jm[:, 7:10, 14], Jm[:, 7:10, 15], rtol=1e-5)     # tie == active
        assert np.abs(J - o.jacobian(om, Z.astype(dtype).astype(np.float64))).max() < TOL[dtype] * 100


# ---- committed golden fixtures (BASELINE configs) ------------------------------------------------------------------------------
def test_golden_c1_c2_cartpole(rd):
    gm = rd.Cartpole()
    g = golden("c1_cartpole_rk3")
    assert np.abs(gm._h.discrete_jacobian(o.RK3, g["Z"], float(g["dt"])) - g["J"]).max() < 1e-10
    g = golden("c2_cartpole_rk4")
    xn = np.empty_like(g["xn"])
    assert np.abs(gm._h.discrete_jacobian(o.RK4, g["Z"], float(g["dt"]), xn=xn) - g["J"]).max() < 1e-10
    assert np.abs(xn - g["xn"]).max() < 1e-10


def test_golden_c3_quadrotor_fp32_and_liestate(rd):
    gm = rd.Quadrotor()
    g = golden("c3_quadrotor_rk4")
    Z32 = g["Z"].astype(np.float32)
    assert np.array_equal(Z32.astype(np.float64), g["Z"])           # fixture inputs are exactly representable in fp32
    assert np.abs(gm._h.discrete_jacobian(o.RK4, Z32, float(g["dt"])) - g["J"]).max() < 1e-4
    assert np.abs(gm._h.discrete_jacobian(o.RK4, g["Z"], float(g["dt"])) - g["J"]).max() < 1e-10
    X = np.ascontiguousarray(g["Z"][:, :13])
    assert np.abs(gm._h.errstate_jacobian(g["Z"]) - g["G"]).max() < 1e-12          # ldx = n+m: states read in place
    assert np.abs(gm._h.errstate_jacobian(X.astype(np.float32)) - g["G"]).max() < 1e-6
    # Cayley error = vec/scalar of q0\\q: entries reach 1e2..1e3 for near-opposite random attitudes -> relative tolerance
    assert np.abs(gm._h.state_diff(X, g["X0"]) - g["dX"]).max() < 1e-10 * max(1.0, np.abs(g["dX"]).max())
    assert np.abs(gm._h.grad_errstate_jacobian(X, g["X0"]) - g["H"]).max() < 1e-12


def test_golden_c4_satellite_and_c5_sweep_and_rollout(rd):
    g = golden("c4_satellite_mrp_rk2")
    gm = rd.Satellite(rd.MRP)
    assert np.abs(gm._h.discrete_jacobian(o.RK2, g["Z"], float(g["dt"])) - g["J"]).max() < 1e-10
    assert np.abs(gm._h.discrete_jacobian(o.RK2, g["Z"].astype(np.float32), float(g["dt"])) - g["J"]).max() < 1e-4
    assert np.abs(gm._h.errstate_jacobian(g["Z"]) - g["G"]).max() < 1e-12
    g = golden("c5_mixed_sweep")
    assert np.abs(rd.Cartpole()._h.discrete_jacobian(o.RK4, g["Zc"], g["dtc"]) - g["Jc"]).max() < 1e-10
    assert np.abs(rd.Quadrotor()._h.discrete_jacobian(o.RK4, g["Zq"], g["dtq"]) - g["Jq"]).max() < 1e-10
    g = golden("rollout_cartpole_rk4")
    assert np.abs(rd.Cartpole()._h.rollout(o.RK4, g["x0"], g["U"], float(g["dt"])) - g["X"]).max() < 1e-10


# ---- LieState maps and rollouts for every rotation ---------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["body_quat_world", "body_mrp_world", "body_rp_body", "cartpole"])
def test_liestate_maps(rd, torch_, name):
    om, gm = zoo()[name][0](), zoo()[name][1](rd)
    rng = np.random.default_rng(31)
    N = 1001
    X = rand_inputs(om.n, om.m, N, rng)[:, :om.n].copy()
    X0 = X + 0.2 * rng.standard_normal(X.shape)                 # nearby reference states: the Cayley error vec/scalar of q0\\q
    if om.n >= 12:                                              # is singular at 180 deg, keep it well conditioned for fp32
        X0[:, 3:3 + om.n - 9] = X[:, 3:3 + om.n - 9] + 0.05 * rng.standard_normal((N, om.n - 9))
    X[:, 3:7] *= 1.7 if om.n == 13 else 1.0                     # un-normalised quaternions: the maps normalise (Appendix A.2)
    for dtype, tol in ((np.float64, 1e-12), (np.float32, 2e-5)):
        Xt, X0t = X.astype(dtype), X0.astype(dtype)
        Go, do_, Ho = o.errstate_jacobian(om, Xt.astype(np.float64)), o.state_diff(om, Xt.astype(np.float64), X0t.astype(np.float64)), \
            o.grad_errstate_jacobian(om, Xt.astype(np.float64), X0t.astype(np.float64))
        assert np.abs(gm._h.errstate_jacobian(Xt) - Go).max() < tol
        assert np.abs(gm._h.errstate_jacobian(dev(torch_, Xt)).cpu().numpy() - Go).max() < tol
        assert np.abs(gm._h.state_diff(Xt, X0t) - do_).max() < tol * max(1.0, np.abs(do_).max())
        assert np.abs(gm._h.state_diff(dev(torch_, Xt), dev(torch_, X0t)).cpu().numpy() - do_).max() < tol * max(1.0, np.abs(do_).max())
        assert np.abs(gm._h.grad_errstate_jacobian(Xt, X0t) - Ho).max() < tol * 10


@pytest.mark.parametrize("name", ["cartpole", "quad_quat_world", "body_mrp_body", "di2"])
def test_rollout(rd, torch_, name):
    om, gm = zoo()[name][0](), zoo()[name][1](rd)
    rng = np.random.default_rng(41)
    ntraj, K = 130, 40
    x0 = rand_inputs(om.n, om.m, ntraj, rng)[:, :om.n].copy()
    U = rng.random((ntraj, K - 1, om.m))
    dt = np.repeat(0.01 * (1 + np.arange(ntraj) % 4), K).reshape(ntraj, K)
    for Q in (o.RK2, o.RK4):
        ref = o.rollout(om, Q, x0, U, dt)
        assert np.abs(gm._h.rollout(Q, x0, U, dt) - ref).max() < 1e-10 * max(1.0, np.abs(ref).max())
        Xd = gm._h.rollout(Q, dev(torch_, x0), dev(torch_, U), dt)
        assert np.abs(Xd.cpu().numpy() - ref).max() < 1e-10 * max(1.0, np.abs(ref).max())
    X32 = gm._h.rollout(o.RK4, x0.astype(np.float32), U.astype(np.float32), 0.01)
    ref = o.rollout(om, o.RK4, x0.astype(np.float32).astype(np.float64), U.astype(np.float32).astype(np.float64), 0.01)
    assert np.abs(X32 - ref).max() < 1e-4 * max(1.0, np.abs(ref).max())


# ---- BASELINE full sizes: size-independent properties + strided oracle sample -------------------------------------------------------
def _full_size_check(rd, torch, gm, om, Q, N, dtype, dt, tol):
    n, m = om.n, om.m
    rng = np.random.default_rng(N % 1000 + n)
    Z = rand_inputs(n, m, N, rng)
    Z[:, n:] = 0.05 + 0.95 * Z[:, n:]                            # keep controls away from the quadrotor's max(0, .) kink for (b)
    Z = Z.astype(dtype)
    Zd = dev(torch, Z)
    xn = torch.empty((N, n), dtype=Zd.dtype, device="cuda")
    J = gm._h.discrete_jacobian(Q, Zd, dt, xn=xn)
    # (a) EVERY knot of the full-size batch against the oracle (all host threads: a few hundred milliseconds), x+ too
    Jo = o.discrete_jacobian(om, Q, Z.astype(np.float64), dt, nthreads=o.num_procs())
    assert np.abs(J.cpu().numpy() - Jo).max() < tol
    xo = o.discrete_dynamics(om, Q, Z.astype(np.float64), dt, nthreads=o.num_procs())
    assert np.abs(xn.cpu().numpy() - xo).max() < tol
    del Jo, xo
    # (b) directional finite difference of the GPU's own discrete_dynamics agrees with J d (fp64 arithmetic for the check)
    Z64 = Zd.double()
    d = torch.from_numpy(rng.standard_normal((N, n + m))).cuda()
    eps = 1e-6
    h64 = gm._h.discrete_dynamics
    fd = (h64(Q, (Z64 + eps * d).contiguous(), dt) - h64(Q, (Z64 - eps * d).contiguous(), dt)) / (2 * eps)
    Jd = torch.einsum("kji,kj->ki", J.double(), d)               # J stored (N, n+m, n)
    assert float((fd - Jd).abs().max()) < max(tol * 20, 1e-7) * max(1.0, float(Jd.abs().max()))
    # (c) x+ from the Jacobian call equals discrete_dynamics to rounding (the dual-number code path may contract FMAs
    #     differently); (d) a second call is bit-identical (no races, no stale shared memory)
    assert float((xn - gm._h.discrete_dynamics(Q, Zd, dt)).abs().max()) < (1e-13 if dtype == np.float64 else 1e-5)
    assert torch.equal(J, gm._h.discrete_jacobian(Q, Zd, dt))
    # (e) structural facts of the reference map
    if n == 4:                                                   # cartpole: column 1 of J is e1; d x1+/d x3 = dt
        Jm = J.transpose(1, 2)
        assert torch.all(Jm[:, :, 0] == torch.tensor([1.0, 0, 0, 0], dtype=J.dtype, device="cuda"))
    return J


def test_full_size_c2_cartpole_rk4_fp64(rd, torch_):
    _full_size_check(rd, torch_, rd.Cartpole(), o.cartpole(), o.RK4, 1 << 20, np.float64, 0.01, 1e-10)


def test_full_size_c3_quadrotor_rk4_fp32(rd, torch_):
    J = _full_size_check(rd, torch_, rd.Quadrotor(), o.quadrotor(), o.RK4, 262144, np.float32, 0.01, 1e-4)
    Jm = J.transpose(1, 2)                                         # position columns are unit vectors; omega rows ignore r, q, v
    assert torch_.all(Jm[:, :, 0:3] == torch_.eye(13, dtype=J.dtype, device="cuda")[:, 0:3])
    assert torch_.all(Jm[:, 10:13, 0:10] == 0)


def test_full_size_c4_satellite_mrp_rk2(rd, torch_):
    for dtype, tol in ((np.float64, 1e-10), (np.float32, 1e-4)):
        _full_size_check(rd, torch_, rd.Satellite(rd.MRP), o.satellite(o.ROT_MRP), o.RK2, 1 << 20, dtype, 0.1, tol)


def test_full_size_c5_mixed_sweep(rd, torch_):
    """4096 trajectories x 256 knots, first half Cartpole, second half Quadrotor, per-trajectory dt; this rank's shard only
    differs from the whole by the partition, so the single-GPU sweep checks the union of all shards."""
    from rdb200 import sharding as sh
    ntraj, K = 4096, 256
    segs = {"cartpole": ntraj // 2, "quadrotor": ntraj // 2}
    models = {"cartpole": (rd.Cartpole(), o.cartpole(), np.float64, 1e-10), "quadrotor": (rd.Quadrotor(), o.quadrotor(), np.float32, 1e-4)}
    rng = np.random.default_rng(5)
    for name, (gm, om, dtype, tol) in models.items():
        Z = rand_inputs(om.n, om.m, segs[name] * K, rng).astype(dtype)
        dt = np.repeat(0.01 * (1 + np.arange(segs[name]) % 4), K)
        whole = gm._h.discrete_jacobian(o.RK4, dev(torch_, Z), dt)
        parts = []
        for r in range(8):                                       # the 8-GPU partition, evaluated shard by shard
            lo, hi = sh.partition_segments(segs, 8, r)[name]
            parts.append(gm._h.discrete_jacobian(o.RK4, dev(torch_, Z[lo * K:hi * K]), dt[lo * K:hi * K]))
        assert torch_.equal(torch_.cat(parts), whole)
        idx = np.arange(0, Z.shape[0], 1499)
        assert np.abs(whole[torch_.from_numpy(idx).cuda()].cpu().numpy() - o.discrete_jacobian(om, o.RK4, Z[idx].astype(np.float64), dt[idx])).max() < tol


# ---- the reference-facing API mirror ---------------------------------------------------------------------------------------------------
def test_reference_api_single_knot_and_trajectory(rd):
    """Mirrors test/cartpole_test.jl:48-72 and test/integration_tests.jl:7-18 through the drop-in names."""
    model = rd.Cartpole()
    dmodel = rd.DiscretizedDynamics(model, rd.RK4)
    om = o.cartpole()
    rng = np.random.default_rng(51)
    x, u = rng.random(4), rng.random(1)
    z = rd.KnotPoint(x, u, 0.0, 0.01)
    zz = np.r_[x, u][None]
    assert rd.dims(dmodel) == (4, 1, 4) and isinstance(rd.default_diffmethod(dmodel), rd.ForwardAD)
    assert np.abs(rd.dynamics(model, z) - o.dynamics(om, zz)[0]).max() < 1e-12
    assert np.abs(rd.dynamics(model, x, u) - o.dynamics(om, zz)[0]).max() < 1e-12
    assert np.abs(rd.discrete_dynamics(dmodel, z) - o.discrete_dynamics(om, o.RK4, zz, 0.01)[0]).max() < 1e-12
    assert np.abs(rd.discrete_dynamics(dmodel, x, u, 0.0, 0.01) - rd.discrete_dynamics(rd.RK4, model, z)).max() == 0
    Jref = o.as_matrix(o.discrete_jacobian(om, o.RK4, zz, 0.01))[0]
    for J in (np.zeros((4, 5)), rd.DynamicsJacobian(model)):
        for sig in (rd.StaticReturn(), rd.InPlace()):
            for diff in (rd.ForwardAD(), rd.UserDefined(), rd.B200()):
                y = np.zeros(4)
                assert rd.jacobian_(sig, diff, dmodel, J, y, z) is None
                assert np.abs(np.asarray(J) - Jref).max() < 1e-10
                assert np.abs(y - o.discrete_dynamics(om, o.RK4, zz, 0.01)[0]).max() < 1e-12
    D = rd.DynamicsJacobian(model)
    rd.discrete_jacobian_(rd.RK3, D, model, z)                     # v0.3 spelling
    assert np.abs(D.A - o.as_matrix(o.discrete_jacobian(om, o.RK3, zz, 0.01))[0][:, :4]).max() < 1e-10
    with pytest.raises(rd.NotImplementedModelError):
        rd.jacobian_(rd.StaticReturn(), rd.FiniteDifference(), dmodel, D, None, z)
    Jc = np.zeros((4, 5))
    rd.jacobian_(rd.StaticReturn(), rd.UserDefined(), model, Jc, np.zeros(4), z)       # continuous Jacobian, test/integration_tests.jl:27-33
    assert np.abs(Jc - o.as_matrix(o.jacobian(om, zz))[0]).max() < 1e-10
    # whole trajectory in one call
    Z = rd.SampledTrajectory(rng.random((101, 4)), rng.random((100, 1)), dt=0.02)
    Js, ys = np.zeros((101, 5, 4)), np.zeros((101, 4))
    rd.jacobian_(rd.StaticReturn(), rd.B200(), dmodel, Js, ys, Z)
    assert np.abs(Js - o.discrete_jacobian(om, o.RK4, Z.data, Z.dts)).max() < 1e-10
    assert np.array_equal(Js[-1].T, np.concatenate([np.eye(4), np.zeros((4, 1))], axis=1))       # terminal knot
    x0 = rd.states(Z)[0].copy()
    rd.rollout_(rd.StaticReturn(), dmodel, Z)
    assert np.abs(rd.states(Z) - o.rollout(om, o.RK4, x0[None], rd.controls(Z)[None, :-1], Z.dts[None])[0]).max() < 1e-10


def test_reference_api_rigid_body_liestate(rd):
    """Mirrors test/rigid_body_jacobians.jl:55-83."""
    model = rd.Quadrotor()
    om = o.quadrotor()
    rng = np.random.default_rng(52)
    zz = rand_inputs(13, 4, 1, rng)
    z = rd.KnotPoint(zz[0, :13], zz[0, 13:], 0.0, 0.1)
    assert rd.errstate_dim(model) == 12 and rd.jacobian_width(model) == 16 and model.statevectortype is rd.RotationState
    J = rd.DynamicsJacobian(model)
    rd.jacobian_(rd.StaticReturn(), rd.ForwardAD(), rd.DiscretizedDynamics(model, rd.RK4), J, np.zeros(13), z)
    assert np.abs(np.asarray(J) - o.as_matrix(o.discrete_jacobian(om, o.RK4, zz, 0.1))[0]).max() < 1e-10
    G = np.zeros((13, 12))
    rd.errstate_jacobian_(model, G, z)
    assert np.abs(G - o.errstate_jacobian(om, zz[:, :13])[0].T).max() < 1e-12
    x0 = rand_inputs(13, 4, 1, rng)[0, :13]
    assert np.abs(rd.state_diff(model, zz[0, :13], x0) - o.state_diff(om, zz[:, :13], x0[None])[0]).max() < 1e-12
    dG = np.zeros((12, 12))
    rd.grad_errstate_jacobian_(model, dG, zz[0, :13], x0)
    assert np.abs(dG - o.grad_errstate_jacobian(om, zz[:, :13], x0[None])[0].T).max() < 1e-12


# ---- fused error-state Jacobian -------------------------------------------------------------------------------------
def _error_jacobian_ref(om, Q, Z, dt):
    """Gn' [A B] blkdiag(G, I) assembled from the oracle's jacobian / errstate_jacobian / discrete_dynamics."""
    n, m, ne = om.n, om.m, om.nerr
    J = o.as_matrix(o.discrete_jacobian(om, Q, Z, dt))                 # (N, n, n+m)
    xn = o.discrete_dynamics(om, Q, Z, dt)
    G = o.as_matrix(o.errstate_jacobian(om, Z[:, :n]))                 # (N, n, nerr)
    Gn = o.as_matrix(o.errstate_jacobian(om, xn))
    A = np.einsum("kia,kij,kjb->kab", Gn, J[:, :, :n], G)
    B = np.einsum("kia,kij->kaj", Gn, J[:, :, n:])
    return np.concatenate([A, B], axis=2)                              # (N, nerr, nerr+m)


@pytest.mark.parametrize("name", ["quad_quat_world", "quad_mrp_world", "body_quat_body", "body_rp_world", "satellite_mrp", "cartpole"])
def test_discrete_error_jacobian(rd, torch_, name):
    om, gm = zoo()[name][0](), zoo()[name][1](rd)
    N = 1500
    Z = rand_inputs(om.n, om.m, N, np.random.default_rng(61))
    if om.n == 13:
        Z[:, 3:7] *= 1.3                                                # off the unit sphere: G normalises, the dynamics does not
    dt = np.random.default_rng(62).uniform(0.0, 0.1, N)
    for dtype, tol in ((np.float64, 1e-10), (np.float32, 1e-4)):
        Zt = Z.astype(dtype)
        for Q in (o.RK4, o.RK2, o.EULER, o.RK3):
            ref = _error_jacobian_ref(om, Q, Zt.astype(np.float64), dt)
            xn = np.empty((N, om.n), dtype=dtype)
            Jb = gm._h.discrete_error_jacobian(Q, Zt, dt, xn=xn)
            assert Jb.shape == (N, om.nerr + om.m, om.nerr)
            assert np.abs(o.as_matrix(Jb) - ref).max() < tol
            assert np.abs(xn - o.discrete_dynamics(om, Q, Zt.astype(np.float64), dt)).max() < tol
        Jd = gm._h.discrete_error_jacobian(o.RK4, dev(torch_, Zt), dt)
        Js = gm._h.discrete_error_jacobian(o.RK4, dev(torch_, np.ascontiguousarray(Zt.T)), dt, layout=rd.SOA)
        torch_.cuda.synchronize()
        ref = _error_jacobian_ref(om, o.RK4, Zt.astype(np.float64), dt)
        assert np.abs(o.as_matrix(Jd.cpu().numpy()) - ref).max() < tol
        assert np.array_equal(Js.cpu().numpy().T.reshape(N, om.nerr + om.m, om.nerr), Jd.cpu().numpy())
        # wide-tile kernels (N above launch.cuh's RDB_SMALL_N)
        Nb = 8192 + 777
        Zb = np.tile(Zt, (Nb // N + 1, 1))[:Nb]
        Jbig = gm._h.discrete_error_jacobian(o.RK4, dev(torch_, Zb), 0.05)
        torch_.cuda.synchronize()
        assert np.abs(o.as_matrix(Jbig.cpu().numpy()) - _error_jacobian_ref(om, o.RK4, Zb.astype(np.float64), 0.05)).max() < tol
    # reference-facing spelling, one knot
    dm = rd.DiscretizedDynamics(gm, rd.RK4)
    z = rd.KnotPoint(Z[0, :om.n], Z[0, om.n:], 0.0, 0.05)
    Jbar, y = np.zeros((om.nerr, om.nerr + om.m)), np.zeros(om.n)
    rd.discrete_error_jacobian_(dm, Jbar, y, z)
    assert np.abs(Jbar - _error_jacobian_ref(om, o.RK4, Z[:1], 0.05)[0]).max() < 1e-10


def test_device_trajectory_and_rollout_linearize(rd, torch_):
    """persistent device trajectory (no per-call gather / PCIe) and rollout + linearisation on the device."""
    model = rd.Quadrotor()
    dmodel = rd.DiscretizedDynamics(model, rd.RK4)
    om = o.quadrotor()
    rng = np.random.default_rng(71)
    K = 65
    Zh = rd.SampledTrajectory(rand_inputs(13, 4, K, rng)[:, :13], 0.5 + rng.random((K - 1, 4)), dt=0.02)
    Zd = rd.DeviceTrajectory(model, Zh)
    J = torch_.zeros((K, 17, 13), dtype=torch_.float64, device="cuda")
    y = torch_.zeros((K, 13), dtype=torch_.float64, device="cuda")
    rd.jacobian_(rd.StaticReturn(), rd.B200(), dmodel, J, y, Zd)
    ref = o.discrete_jacobian(om, o.RK4, Zh.data, Zh.dts)
    assert np.abs(J.cpu().numpy() - ref).max() < 1e-10
    U2 = 0.5 + rng.random((K - 1, 4))
    rd.setcontrols_(Zd, U2)                                       # H2D update of the controls only
    rd.setcontrols_(Zh, U2)
    rd.jacobian_(rd.StaticReturn(), rd.B200(), dmodel, J, y, Zd)
    assert np.abs(J.cpu().numpy() - o.discrete_jacobian(om, o.RK4, Zh.data, Zh.dts)).max() < 1e-10
    assert np.array_equal(Zd.to_host().data, Zh.data)
    Jb = torch_.zeros((K, 16, 12), dtype=torch_.float64, device="cuda")
    rd.discrete_error_jacobian_(dmodel, Jb, None, Zd)
    assert np.abs(o.as_matrix(Jb.cpu().numpy()) - _error_jacobian_ref(om, o.RK4, Zh.data, Zh.dts)).max() < 1e-10
    # many trajectories: forward pass + linearisation, all on the device
    ntraj = 96
    x0 = rand_inputs(13, 4, ntraj, rng)[:, :13].copy()
    U = 0.5 + rng.random((ntraj, K - 1, 4))
    X, Jt = rd.rollout_and_linearize(dmodel, dev(torch_, x0), dev(torch_, U), 0.02)
    Xo = o.rollout(om, o.RK4, x0, U, 0.02)
    assert np.abs(X.cpu().numpy() - Xo).max() < 1e-10 * max(1.0, np.abs(Xo).max())
    Zo = np.concatenate([Xo[:, :-1], U], axis=2).reshape(-1, 17)
    assert np.abs(Jt.cpu().numpy().reshape(-1, 17, 13) - o.discrete_jacobian(om, o.RK4, Zo, 0.02)).max() < 1e-8
    Xh, Jh = rd.rollout_and_linearize(dmodel, x0, U, 0.02, error_state=True)        # host arrays, error-state form
    assert Jh.shape == (ntraj, K - 1, 16, 12) and np.abs(Xh - Xo).max() < 1e-10 * max(1.0, np.abs(Xo).max())


@pytest.mark.parametrize("name,dtype", [("cartpole", np.float64), ("quad_quat_world", np.float32), ("body_mrp_body", np.float64)])
def test_trajectory_batch_rollout_linearize(rd, torch_, name, dtype):
    """rdb_trajectory_*: the device trajectory of the C ABI.  Rollout in place (knot-major batch), Jacobians of every knot, and the
    chunk-pipelined rollout + linearisation on two streams — against the CPU checker, for host and device outputs, scalar and
    per-knot steps, full-state and error-state Jacobians."""
    om, gm = zoo()[name][0](), zoo()[name][1](rd)
    n, m = om.n, om.m
    dm = rd.DiscretizedDynamics(gm, rd.RK4)
    rng = np.random.default_rng(150)
    ntraj, K = 70, 37
    x0 = rand_inputs(n, m, ntraj, rng)[:, :n].astype(dtype)
    U = (0.4 * rng.random((K - 1, ntraj, m))).astype(dtype)
    dts = rng.uniform(0.005, 0.03, (K, ntraj))
    tol = TOL[dtype] * (1 if dtype == np.float64 else 3)
    for dt in (0.02, dts):
        tb = rd.TrajectoryBatch(dm, ntraj, K, dtype)
        tb.set_initial_state(x0); tb.set_controls(U); tb.set_timesteps(dt)
        Zv, tv, dv = tb.views()
        dt_ref = dv.cpu().numpy()                                              # (K, ntraj): terminal step forced to 0
        assert np.all(dt_ref[-1] == 0) and np.allclose(tv.cpu().numpy()[1:], np.cumsum(dt_ref[:-1], axis=0))
        assert np.all(Zv.cpu().numpy()[-1, :, n:] == 0)                       # no terminal control
        tb.rollout()
        X = tb.states()
        Xo = o.rollout(om, o.RK4, x0.astype(np.float64), np.transpose(U, (1, 0, 2)).astype(np.float64), np.ascontiguousarray(dt_ref.T))
        assert np.abs(np.transpose(X, (1, 0, 2)) - Xo).max() < tol * max(1.0, np.abs(Xo).max())
        Zall = Zv.cpu().numpy().reshape(K * ntraj, n + m)
        ref = o.discrete_jacobian(om, o.RK4, Zall.astype(np.float64), dt_ref.reshape(-1)).reshape(K, ntraj, n + m, n)
        Jd = tb.linearize()                                                   # device output
        xn = np.empty((K, ntraj, n), dtype=dtype)
        Jh = tb.linearize(J=np.empty((K, ntraj, n + m, n), dtype=dtype), xn=xn, device=False)     # device inputs, HOST outputs
        assert np.abs(Jd.cpu().numpy() - ref).max() < tol and np.array_equal(Jh, Jd.cpu().numpy())
        assert np.abs(xn[:-1] - X[1:]).max() < tol * max(1.0, np.abs(Xo).max())
        eye = np.concatenate([np.eye(n), np.zeros((m, n))])                    # terminal knots: J = [I 0] (src/knotpoint.jl:57-67)
        assert np.abs(Jh[-1] - eye).max() == 0
        # the pipelined forward pass + linearisation from scratch, several chunkings: same states, same Jacobians
        for chunks in (0, 1, 5, K + 3):
            tb2 = rd.TrajectoryBatch(dm, ntraj, K, dtype)
            tb2.set_initial_state(dev(torch_, x0)); tb2.set_controls(dev(torch_, U)); tb2.set_timesteps(dt if np.ndim(dt) == 0 else dev(torch_, dt))
            J2 = tb2.rollout_linearize(chunks=chunks)
            torch_.cuda.synchronize()
            assert np.array_equal(tb2.states(), X) and np.array_equal(J2.cpu().numpy(), Jh), chunks
        Jp = tb.rollout_linearize(device=False)                                # host Jacobians
        assert np.array_equal(Jp, Jh)
        if name != "cartpole":
            Jb = tb.rollout_linearize(error_state=True)
            refb = _error_jacobian_ref(om, o.RK4, Zall.astype(np.float64), dt_ref.reshape(-1))
            assert np.abs(o.as_matrix(Jb.cpu().numpy().reshape(K * ntraj, 12 + m, 12)) - refb).max() < tol * 3
    # setters: states / controls round trip, shape and dtype errors
    Xs = rng.random((K, ntraj, n)).astype(dtype)
    tb.set_states(Xs)
    assert np.array_equal(tb.states(), Xs) and np.array_equal(tb.controls()[:-1], U)
    with pytest.raises(ValueError):
        tb.set_states(Xs[:-1])
    with pytest.raises(rd.RDBError):
        rd._abi.check(rd._abi.lib().rdb_trajectory_set_controls(tb._t._p, Xs.ctypes.data, K - 2, None), "bad knot count")


def test_plans_and_time_varying_trajectory(rd, torch_):
    """rdb_plan_*: a pre-validated launch equals the direct call bit for bit, replays under CUDA-graph capture and follows the buffer
    contents; and a time-varying user model rolled out on a device trajectory sees the trajectory's own time grid."""
    gm = rd.Quadrotor()
    N = 777
    Z = dev(torch_, rand_inputs(13, 4, N, np.random.default_rng(160)).astype(np.float32))
    ref = gm._h.discrete_jacobian(rd.RK4.code, Z, 0.01)
    plan = rd._abi.Plan(gm._h, rd._abi.OP_DISCRETE_JACOBIAN, rd.RK4.code, Z, 0.01)
    assert torch_.equal(plan.launch(), ref)
    eplan = rd._abi.Plan(gm._h, rd._abi.OP_DISCRETE_ERROR_JACOBIAN, rd.RK4.code, Z, 0.01)
    assert torch_.equal(eplan.launch(), gm._h.discrete_error_jacobian(rd.RK4.code, Z, 0.01))
    vplan = rd._abi.Plan(gm._h, rd._abi.OP_DISCRETE_DYNAMICS, rd.RK4.code, Z, 0.01)
    assert torch_.equal(vplan.launch(), gm._h.discrete_dynamics(rd.RK4.code, Z, 0.01))
    g = torch_.cuda.CUDAGraph()
    st = torch_.cuda.Stream()
    with torch_.cuda.stream(st):
        plan.launch(); torch_.cuda.synchronize()
        with torch_.cuda.graph(g, stream=st):
            plan.launch()
    Z.mul_(0.5); Z[:, 3:7] *= 2.0                                           # new contents, same buffers
    plan.J.zero_(); g.replay(); torch_.cuda.synchronize()
    assert torch_.equal(plan.J, gm._h.discrete_jacobian(rd.RK4.code, Z, 0.01))
    with pytest.raises(rd.RDBError):
        rd._abi.Plan(gm._h, rd._abi.OP_DISCRETE_JACOBIAN, rd.RK4.code, Z.cpu().numpy(), 0.01)      # plans are for device data
    # time-varying user model on a device trajectory: stage times come from the trajectory's time grid (t0 + cumsum(dt))
    p0, t0, h = 1.7, 0.25, 0.04
    tv = rd.CustomModel(2, 1, TV_BODY, params=[p0])
    dm = rd.DiscretizedDynamics(tv, rd.RK4)
    rng = np.random.default_rng(161)
    ntraj, K = 9, 11
    x0, U = rng.random((ntraj, 2)), rng.random((K - 1, ntraj, 1))
    tb = rd.TrajectoryBatch(dm, ntraj, K)
    tb.set_initial_state(x0); tb.set_controls(U); tb.set_timesteps(h, t0=t0)
    J = tb.rollout_linearize(chunks=3, device=False)
    X = tb.states()
    for j in range(ntraj):
        x = x0[j].copy()
        for k in range(K - 1):
            z = np.r_[x, U[k, j]]
            assert np.abs(J[k, j].T - _cs_jac(lambda zz: _tv_step(o.RK4, zz, t0 + h * k, h, p0), z)).max() < 1e-10
            x = _tv_step(o.RK4, z, t0 + h * k, h, p0)
            assert np.abs(X[k + 1, j] - x).max() < 1e-12


# ---- user-defined models -----------------------------------------------------------------------------------------
CARTPOLE_BODY = """
        const T mc = p[0], mp = p[1], l = p[2], g = p[3];
        auto s = sin_(get<1>(x)); auto c = cos_(get<1>(x));
        auto qd0 = get<2>(x); auto qd1 = get<3>(x);
        auto H01 = (mp * l) * c;
        auto r0 = -((mp * l) * (qd1 * s) * qd1) - get<0>(u);
        auto r1 = (mp * g * l) * s;
        auto idet = T(1) / ((mc + mp) * (mp * l * l) - H01 * H01);
        return vec(qd0, qd1, (H01 * r1 - (mp * l * l) * r0) * idet, (H01 * r0 - (mc + mp) * r1) * idet);
"""
# a 7-state, 3-control toy with every elementary function: exercises multi-role tiling of custom models
TOY_BODY = """
        auto a = get<0>(x) * get<1>(x) + exp_(T(-0.5) * get<2>(x));
        auto b = sqrt_(T(1) + get<3>(x) * get<3>(x)) * get<0>(u);
        auto c = relu_(get<4>(x) - p[0]) + sin_(get<5>(x)) * cos_(get<6>(x));
        return vec(get<1>(x), a - b, get<3>(x) / (T(2) + get<2>(u) * get<2>(u)), b * c, get<1>(u) - p[1] * get<4>(x), c, a * get<2>(u));
"""


def _toy_f(z, p):
    x, u = z[:7], z[7:]
    a = x[0] * x[1] + np.exp(-0.5 * x[2])
    b = np.sqrt(1 + x[3] * x[3]) * u[0]
    c = (x[4] - p[0] if np.real(x[4] - p[0]) > 0 else 0 * x[4]) + np.sin(x[5]) * np.cos(x[6])
    return np.array([x[1], a - b, x[3] / (2 + u[2] * u[2]), b * c, u[1] - p[1] * x[4], c, a * u[2]])


def _rk4_cs(f, z, n, h):
    """complex-step Jacobian of one RK4 step of f (independent reference for a model the oracle does not know)."""
    def step(zz):
        x, u = zz[:n], zz[n:]
        k1 = f(np.r_[x, u]); k2 = f(np.r_[x + h / 2 * k1, u]); k3 = f(np.r_[x + h / 2 * k2, u]); k4 = f(np.r_[x + h * k3, u])
        return x + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
    zc = z.astype(complex)
    J = np.stack([np.imag(step(zc + 1e-30j * np.eye(len(z))[j])) / 1e-30 for j in range(len(z))], axis=1)
    return J, np.real(step(zc))


def test_user_defined_models(rd, torch_):
    # (1) the Cartpole re-written as a user model equals the built-in one and the oracle, for every rule and dtype
    um = rd.CustomModel(4, 1, CARTPOLE_BODY, params=[1.0, 0.2, 0.5, 9.81])
    assert rd.dims(um) == (4, 1, 4)
    om = o.cartpole()
    N = 3000 + 17
    Z = np.random.default_rng(81).random((N, 5))
    for dtype, tol in ((np.float64, 1e-10), (np.float32, 1e-4)):
        Zt = Z.astype(dtype)
        for Q in QS:
            xn = np.empty((N, 4), dtype=dtype)
            J = um._h.discrete_jacobian(Q, Zt, 0.02, xn=xn)
            assert np.abs(J - o.discrete_jacobian(om, Q, Zt.astype(np.float64), 0.02)).max() < tol
            assert np.abs(xn - o.discrete_dynamics(om, Q, Zt.astype(np.float64), 0.02)).max() < tol
        assert np.abs(um._h.jacobian(dev(torch_, Zt)).cpu().numpy() - o.jacobian(om, Zt.astype(np.float64))).max() < tol * 10
        assert np.abs(um._h.dynamics(Zt) - o.dynamics(om, Zt.astype(np.float64))).max() < tol * 10
    x0, U = Z[:64, :4].copy(), np.random.default_rng(82).random((64, 19, 1))
    assert np.abs(um._h.rollout(o.RK4, x0, U, 0.02) - o.rollout(om, o.RK4, x0, U, 0.02)).max() < 1e-10
    assert np.array_equal(um._h.errstate_jacobian(Z[:5, :4].copy()), np.broadcast_to(np.eye(4), (5, 4, 4)))
    # the reference-facing spelling works unchanged
    Jm, y = np.zeros((4, 5)), np.zeros(4)
    rd.jacobian_(rd.StaticReturn(), rd.ForwardAD(), rd.DiscretizedDynamics(um, rd.RK4), Jm, y, rd.KnotPoint(Z[0, :4], Z[0, 4:], 0.0, 0.02))
    assert np.abs(Jm - o.as_matrix(o.discrete_jacobian(om, o.RK4, Z[:1], 0.02))[0]).max() < 1e-10
    # (2) a model nobody built in: 7 states, 3 controls, exp / sqrt / relu / division; checked by complex step
    p = [0.3, 0.7]
    toy = rd.CustomModel(7, 3, TOY_BODY, params=p)
    Zt = np.random.default_rng(83).random((400, 10))
    xn = np.empty((400, 7))
    J = toy._h.discrete_jacobian(o.RK4, Zt, 0.05, xn=xn)
    J32 = toy._h.discrete_jacobian(o.RK4, Zt.astype(np.float32), 0.05)
    for k in range(0, 400, 37):
        Jr, xr = _rk4_cs(lambda zz: _toy_f(zz, p), Zt[k], 7, 0.05)
        assert np.abs(J[k].T - Jr).max() < 1e-10 and np.abs(xn[k] - xr).max() < 1e-12
        assert np.abs(J32[k].T - Jr).max() < 1e-4
    with pytest.raises(rd.RDBError) as e:
        rd.CustomModel(2, 1, "return vec(get<1>(x), nope);")
    assert e.value.code == rd._abi.ERR_COMPILE and "nope" in str(e.value)


# ---- time-varying user models: dynamics(model, x, u, t) with the reference's stage times ---------------------------------------------
TV_BODY = """
        auto drive = cos_(T(3) * t) * get<0>(u);
        return vec(get<1>(x), drive - p[0] * sin_(get<0>(x)) + t * get<1>(x));
"""


def _tv_f(x, u, t, p0):
    return np.array([x[1], np.cos(3 * t) * u[0] - p0 * np.sin(x[0]) + t * x[1]])


def _tv_step(rule, z, t, h, p0):
    """one step with the stage times of src/integration.jl:73-76 (Euler), :130-135 (RK3), :280-286 (RK4); RK2 = explicit midpoint"""
    x, u = z[:2], z[2:]
    f = lambda xx, tt: _tv_f(xx, u, tt, p0)
    if rule == o.EULER:
        return x + h * f(x, t)
    if rule == o.RK2:
        return x + h * f(x + h / 2 * f(x, t), t + h / 2)
    if rule == o.RK3:
        k1 = f(x, t) * h; k2 = f(x + k1 / 2, t + h / 2) * h; k3 = f(x - k1 + 2 * k2, t + h) * h
        return x + (k1 + 4 * k2 + k3) / 6
    k1 = f(x, t) * h; k2 = f(x + k1 / 2, t + h / 2) * h; k3 = f(x + k2 / 2, t + h / 2) * h; k4 = f(x + k3, t + h) * h
    return x + (k1 + 2 * k2 + 2 * k3 + k4) / 6


def _cs_jac(fun, z):
    zc = z.astype(complex)
    return np.stack([np.imag(fun(zc + 1e-30j * np.eye(len(z))[j])) / 1e-30 for j in range(len(z))], axis=1)


def test_time_varying_user_model(rd, torch_):
    """`t` reaches user models exactly as the reference feeds it: dynamics(model, x, u, t) (src/dynamics.jl:81-83), RK stages at
    t, t+h/2, t+h/2, t+h (src/integration.jl:281-284), never differentiated.  Checked by complex step of an independent numpy
    statement, for every explicit rule, host and device pointers, per-knot t and dt, and through rollout."""
    p0 = 1.7
    tv = rd.CustomModel(2, 1, TV_BODY, params=[p0])
    rng = np.random.default_rng(140)
    N = 333
    Z, t, dt = rng.random((N, 3)), rng.uniform(0.0, 5.0, N), rng.uniform(0.01, 0.1, N)
    # continuous dynamics and Jacobian at time t
    xd = tv._h.dynamics(Z, t=t)
    Jc = tv._h.jacobian(dev(torch_, Z), t=t).cpu().numpy()
    for k in range(0, N, 29):
        assert np.abs(xd[k] - _tv_f(Z[k, :2], Z[k, 2:], t[k], p0)).max() < 1e-13
        assert np.abs(Jc[k].T - _cs_jac(lambda zz: _tv_f(zz[:2], zz[2:], t[k], p0), Z[k])).max() < 1e-12
    assert np.abs(tv._h.dynamics(Z, t=None) - np.stack([_tv_f(Z[k, :2], Z[k, 2:], 0.0, p0) for k in range(N)])).max() < 1e-13   # no t == t = 0
    for Q in QS:
        xn = np.empty((N, 2))
        J = tv._h.discrete_jacobian(Q, Z, dt, t=t, xn=xn)
        Jd = tv._h.discrete_jacobian(Q, dev(torch_, Z), dt, t=t).cpu().numpy()
        assert np.array_equal(J, Jd)
        J32 = tv._h.discrete_jacobian(Q, Z.astype(np.float32), dt, t=t)
        for k in range(0, N, 29):
            ref = _cs_jac(lambda zz: _tv_step(Q, zz, t[k], dt[k], p0), Z[k])
            assert np.abs(J[k].T - ref).max() < 1e-10 and np.abs(xn[k] - _tv_step(Q, Z[k], t[k], dt[k], p0)).max() < 1e-12
            assert np.abs(J32[k].T - ref).max() < 1e-4
    # the time really matters: the same knots at another time give another Jacobian
    assert np.abs(tv._h.discrete_jacobian(o.RK4, Z, dt, t=t + 1.0) - tv._h.discrete_jacobian(o.RK4, Z, dt, t=t)).max() > 1e-3
    # reference-facing spelling: the knot point's own time is used
    dm = rd.DiscretizedDynamics(tv, rd.RK4)
    Jm, y = np.zeros((2, 3)), np.zeros(2)
    rd.jacobian_(rd.StaticReturn(), rd.ForwardAD(), dm, Jm, y, rd.KnotPoint(Z[0, :2], Z[0, 2:], float(t[0]), float(dt[0])))
    assert np.abs(Jm - _cs_jac(lambda zz: _tv_step(o.RK4, zz, t[0], dt[0], p0), Z[0])).max() < 1e-10
    assert np.abs(rd.discrete_dynamics(dm, Z[0, :2], Z[0, 2:], float(t[0]), float(dt[0])) - _tv_step(o.RK4, Z[0], t[0], dt[0], p0)).max() < 1e-12
    # rollout: explicit per-knot times, and times accumulated from 0 when none are given
    ntraj, K, h = 40, 12, 0.05
    x0, U = rng.random((ntraj, 2)), rng.random((ntraj, K - 1, 1))
    tt = np.broadcast_to(0.3 + h * np.arange(K), (ntraj, K)).copy()
    X = tv._h.rollout(o.RK4, x0, U, h, t=tt)
    X0 = tv._h.rollout(o.RK4, dev(torch_, x0), dev(torch_, U), h).cpu().numpy()
    for j in range(0, ntraj, 7):
        xa, xb = x0[j].copy(), x0[j].copy()
        for k in range(K - 1):
            xa = _tv_step(o.RK4, np.r_[xa, U[j, k]], tt[j, k], h, p0)
            xb = _tv_step(o.RK4, np.r_[xb, U[j, k]], h * k, h, p0)
            assert np.abs(X[j, k + 1] - xa).max() < 1e-12 and np.abs(X0[j, k + 1] - xb).max() < 1e-12


def test_user_rigid_body_wrench(rd, torch_):
    """A Quadrotor defined by the USER through forces/moments (reference: test/quadrotor.jl:56-96 on the RigidBody interface,
    src/rigidbody.jl:244-257) equals the oracle's quadrotor for every rotation / frame, including the LieState maps and the
    error-state Jacobian."""
    from test_abi_host import QUAD_WRENCH
    rng = np.random.default_rng(91)
    for Rname, rc, frame in (("QuatRotation", o.ROT_QUAT, o.WORLD), ("MRP", o.ROT_MRP, o.BODYFRAME), ("RodriguesParam", o.ROT_RP, o.WORLD)):
        om = o.quadrotor(rc, frame)
        um = rd.CustomRigidBody(getattr(rd, Rname), 4, QUAD_WRENCH, mass=0.5, J=(0.0023, 0.0023, 0.004),
                                params=[1.0, 0.0245, 0.175, 0.0, 0.0, -9.81], bodyframe=bool(frame))
        assert (um.n, um.m, rd.errstate_dim(um)) == (om.n, om.m, 12) and um.statevectortype is rd.RotationState
        N = 700
        Z = rand_inputs(om.n, om.m, N, rng)
        for dtype, tol in ((np.float64, 1e-10), (np.float32, 1e-4)):
            Zt = Z.astype(dtype)
            Z64 = Zt.astype(np.float64)
            for Q in (o.RK4, o.RK2):
                xn = np.empty((N, om.n), dtype=dtype)
                J = um._h.discrete_jacobian(Q, Zt, 0.03, xn=xn)
                assert np.abs(J - o.discrete_jacobian(om, Q, Z64, 0.03)).max() < tol
                assert np.abs(xn - o.discrete_dynamics(om, Q, Z64, 0.03)).max() < tol
            Jb = um._h.discrete_error_jacobian(o.RK4, dev(torch_, Zt), 0.03)
            assert np.abs(o.as_matrix(Jb.cpu().numpy()) - _error_jacobian_ref(om, o.RK4, Z64, 0.03)).max() < tol
            assert np.abs(um._h.dynamics(Zt) - o.dynamics(om, Z64)).max() < tol * max(1.0, np.abs(o.dynamics(om, Z64)).max())
        X = np.ascontiguousarray(Z[:, :om.n])
        assert np.abs(um._h.errstate_jacobian(X) - o.errstate_jacobian(om, X)).max() < 1e-12
        x0, U = X[:50].copy(), rng.random((50, 15, 4))
        ref = o.rollout(om, o.RK4, x0, U, 0.02)
        assert np.abs(um._h.rollout(o.RK4, x0, U, 0.02) - ref).max() < 1e-10 * max(1.0, np.abs(ref).max())


def test_sixteen_million_knots_64bit_indexing(rd, torch_):
    """N = 2^24 + 5 Cartpole knots in fp32 (1.3 GB of Jacobians): element offsets exceed 2^31; strided sample + ragged tail vs oracle."""
    N = (1 << 24) + 5
    gen = torch_.Generator(device="cuda").manual_seed(7)
    Z = torch_.rand((N, 5), dtype=torch_.float32, device="cuda", generator=gen)
    J = rd.Cartpole()._h.discrete_jacobian(o.RK4, Z, 0.01)
    idx = torch_.cat([torch_.arange(0, N, 65521, device="cuda"), torch_.arange(N - 40, N, device="cuda")])
    ref = o.discrete_jacobian(o.cartpole(), o.RK4, Z[idx].cpu().numpy().astype(np.float64), 0.01)
    assert np.abs(J[idx].cpu().numpy() - ref).max() < 1e-4


def test_calls_are_cuda_graph_capturable(rd, torch_):
    """Device-pointer calls only enqueue work on the caller's stream (no syncs, no allocations), so a solver can capture its
    whole linearisation step in a CUDA graph and replay it (launch-bound regimes: many small batches)."""
    cp, qd = rd.Cartpole(), rd.Quadrotor()
    rng = np.random.default_rng(101)
    Zc = dev(torch_, rng.random((5000, 5)))
    Zq = dev(torch_, rand_inputs(13, 4, 3000, rng).astype(np.float32))
    Jc = torch_.zeros((5000, 5, 4), dtype=torch_.float64, device="cuda")
    Jq = torch_.zeros((3000, 17, 13), dtype=torch_.float32, device="cuda")
    Gq = torch_.zeros((3000, 12, 13), dtype=torch_.float32, device="cuda")
    def step():
        cp._h.discrete_jacobian(o.RK4, Zc, 0.01, J=Jc)
        qd._h.discrete_jacobian(o.RK4, Zq, 0.01, J=Jq)
        qd._h.errstate_jacobian(Zq, G=Gq)
    step()                                                        # warm-up outside capture (kernel attributes, occupancy query)
    torch_.cuda.synchronize()
    ref_c, ref_q = Jc.clone(), Jq.clone()
    g = torch_.cuda.CUDAGraph()
    s = torch_.cuda.Stream()
    with torch_.cuda.stream(s):
        with torch_.cuda.graph(g, stream=s):
            step()
    Jc.zero_(); Jq.zero_(); Gq.zero_()
    Zc.mul_(0.5)                                                  # new inputs, same buffers
    g.replay()
    torch_.cuda.synchronize()
    assert np.abs(Jc.cpu().numpy() - o.discrete_jacobian(o.cartpole(), o.RK4, Zc.cpu().numpy(), 0.01)).max() < 1e-10
    assert torch_.equal(Jq, ref_q) and not torch_.equal(Jc, ref_c) and float(Gq.abs().sum()) > 0


@pytest.mark.parametrize("name", ["cartpole", "quad_quat_world", "body_mrp_body", "satellite_mrp", "di3"])
def test_implicit_midpoint(rd, torch_, name):
    """ImplicitMidpoint on the GPU (Newton + pivoted LU + implicit-function-theorem Jacobian; warp-cooperative kernel for the rigid
    bodies, per-thread kernel for the small models) against the oracle."""
    om, gm = zoo()[name][0](), zoo()[name][1](rd)
    N = 900
    Z = rand_inputs(om.n, om.m, N, np.random.default_rng(111))
    dt = np.random.default_rng(112).uniform(0.005, 0.1, N)
    ref_J = o.discrete_jacobian(om, o.IMPLICIT_MIDPOINT, Z, dt)
    ref_x = o.discrete_dynamics(om, o.IMPLICIT_MIDPOINT, Z, dt)
    xn = np.empty((N, om.n))
    J = gm._h.discrete_jacobian(rd._abi.IMPLICIT_MIDPOINT, Z, dt, xn=xn)
    assert np.abs(J - ref_J).max() < 1e-10 and np.abs(xn - ref_x).max() < 1e-10
    Jd = gm._h.discrete_jacobian(rd._abi.IMPLICIT_MIDPOINT, dev(torch_, Z), dt)
    xd = gm._h.discrete_dynamics(rd._abi.IMPLICIT_MIDPOINT, dev(torch_, Z), dt)
    assert np.array_equal(Jd.cpu().numpy(), J) and np.abs(xd.cpu().numpy() - ref_x).max() < 1e-10
    Z32 = Z.astype(np.float32)
    J32 = gm._h.discrete_jacobian(rd._abi.IMPLICIT_MIDPOINT, Z32, dt)
    assert np.abs(J32 - o.discrete_jacobian(om, o.IMPLICIT_MIDPOINT, Z32.astype(np.float64), dt)).max() < 1e-4        # north_star: 1e-4 (fp32)
    # reference-facing spelling
    dm = rd.DiscretizedDynamics(gm, rd.ImplicitMidpoint)
    Jm, y = np.zeros((om.n, om.n + om.m)), np.zeros(om.n)
    rd.jacobian_(rd.InPlace(), rd.ForwardAD(), dm, Jm, y, rd.KnotPoint(Z[0, :om.n], Z[0, om.n:], 0.0, float(dt[0])))
    assert np.abs(Jm - o.as_matrix(ref_J)[0]).max() < 1e-10 and np.abs(y - ref_x[0]).max() < 1e-10


# ---- dynamics_error / dynamics_error_jacobian! and ImplicitMidpoint for user models ---------------------------------------------------
def _cs_cols(fun, z):
    zc = z.astype(complex)
    return np.stack([np.imag(fun(zc + 1e-30j * np.eye(len(z))[j])) / 1e-30 for j in range(len(z))], axis=1)


@pytest.mark.parametrize("name", ["cartpole", "quad_quat_world", "body_mrp_body"])
def test_dynamics_error_and_its_jacobian(rd, torch_, name):
    """dynamics_error(dmodel, z2, z1) / dynamics_error_jacobian! (src/discrete_dynamics.jl:116-200): explicit rules give
    discrete_dynamics(z1) - x2 with J1 = the discrete Jacobian and J2 = [-I 0]; ImplicitMidpoint gives the midpoint residual with
    J1 = [I + h/2 A, h B], J2 = [h/2 A - I, 0] (src/integration.jl:640-700), A, B the continuous Jacobian at the midpoint."""
    om, gm = zoo()[name][0](), zoo()[name][1](rd)
    n, m = om.n, om.m
    N = 500
    rng = np.random.default_rng(170)
    Z1 = rand_inputs(n, m, N, rng)
    Z2 = Z1 + 0.05 * rng.standard_normal((N, n + m))                     # "next knots": only their states are read (ld2 = n + m)
    dt = rng.uniform(0.01, 0.1, N)
    # explicit
    e = gm._h.dynamics_error(o.RK4, Z1, Z2, dt)
    assert np.abs(e - (o.discrete_dynamics(om, o.RK4, Z1, dt) - Z2[:, :n])).max() < 1e-10
    J2, J1, e2 = gm._h.dynamics_error(o.RK4, dev(torch_, Z1), dev(torch_, Z2), dt, jacobian=True)
    assert np.abs(J1.cpu().numpy() - o.discrete_jacobian(om, o.RK4, Z1, dt)).max() < 1e-10 and np.abs(e2.cpu().numpy() - e).max() < 1e-12
    eye = np.concatenate([-np.eye(n), np.zeros((m, n))])
    assert np.array_equal(J2.cpu().numpy(), np.broadcast_to(eye, (N, n + m, n)))
    # ImplicitMidpoint
    IM = rd._abi.IMPLICIT_MIDPOINT
    Zm = np.concatenate([0.5 * (Z1[:, :n] + Z2[:, :n]), Z1[:, n:]], axis=1)
    f, AB = o.dynamics(om, Zm), o.as_matrix(o.jacobian(om, Zm))
    h = dt[:, None]
    ei = gm._h.dynamics_error(IM, Z1, Z2, dt)
    assert np.abs(ei - (Z1[:, :n] + h * f - Z2[:, :n])).max() < 1e-10
    J2, J1, e3 = gm._h.dynamics_error(IM, Z1, Z2, dt, jacobian=True)
    A, Bm = AB[:, :, :n], AB[:, :, n:]
    I = np.eye(n)[None]
    assert np.abs(o.as_matrix(J1) - np.concatenate([I + 0.5 * h[:, :, None] * A, h[:, :, None] * Bm], axis=2)).max() < 1e-10
    assert np.abs(o.as_matrix(J2) - np.concatenate([0.5 * h[:, :, None] * A - I, np.zeros((N, n, m))], axis=2)).max() < 1e-10
    assert np.abs(e3 - ei).max() < 1e-13                          # (plain and dual evaluations of f may round differently)
    # the residual vanishes at the implicit step, and the two Jacobians give the step's Jacobian by the implicit function theorem
    xn = gm._h.discrete_dynamics(IM, Z1, dt)
    assert np.abs(gm._h.dynamics_error(IM, Z1, np.ascontiguousarray(xn), dt)).max() < 1e-10
    J2s, J1s, _ = gm._h.dynamics_error(IM, Z1, np.ascontiguousarray(xn), dt, jacobian=True)
    Jstep = o.as_matrix(gm._h.discrete_jacobian(IM, Z1, dt))
    sol = -np.linalg.solve(o.as_matrix(J2s)[:, :, :n], o.as_matrix(J1s))
    assert np.abs(sol - Jstep).max() < 1e-8
    # reference-facing spelling, one knot pair
    dm = rd.DiscretizedDynamics(gm, rd.ImplicitMidpoint)
    z1, z2 = rd.KnotPoint(Z1[0, :n], Z1[0, n:], 0.0, float(dt[0])), rd.KnotPoint(Z2[0, :n], Z2[0, n:], float(dt[0]), float(dt[0]))
    assert np.abs(rd.dynamics_error(dm, z2, z1) - ei[0]).max() < 1e-12
    Ja, Jb, y2 = np.zeros((n, n + m)), np.zeros((n, n + m)), np.zeros(n)
    rd.dynamics_error_jacobian_(rd.StaticReturn(), rd.ForwardAD(), dm, Ja, Jb, y2, None, z2, z1)
    assert np.abs(Ja - o.as_matrix(J2)[0]).max() < 1e-12 and np.abs(Jb - o.as_matrix(J1)[0]).max() < 1e-12 and np.abs(y2 - ei[0]).max() < 1e-12


def test_implicit_midpoint_for_user_models(rd, torch_):
    """DiscretizedDynamics{L, ImplicitMidpoint} for NVRTC-compiled user models: the Cartpole written as a user model equals the built-in
    one; a time-varying model solves x1 + h f((x1+x2)/2, u, t + h/2) - x2 = 0 with the midpoint TIME; an 8-state model takes the
    warp-cooperative kernel (one column of [A B] per lane)."""
    IM = rd._abi.IMPLICIT_MIDPOINT
    um, om = rd.CustomModel(4, 1, CARTPOLE_BODY, params=[1.0, 0.2, 0.5, 9.81]), o.cartpole()
    N = 700
    rng = np.random.default_rng(180)
    Z, dt = rng.random((N, 5)), rng.uniform(0.01, 0.1, N)
    xn = np.empty((N, 4))
    J = um._h.discrete_jacobian(IM, Z, dt, xn=xn)
    assert np.abs(J - o.discrete_jacobian(om, o.IMPLICIT_MIDPOINT, Z, dt)).max() < 1e-10
    assert np.abs(xn - o.discrete_dynamics(om, o.IMPLICIT_MIDPOINT, Z, dt)).max() < 1e-10
    J32 = um._h.discrete_jacobian(IM, dev(torch_, Z.astype(np.float32)), dt).cpu().numpy()
    assert np.abs(J32 - o.discrete_jacobian(om, o.IMPLICIT_MIDPOINT, Z.astype(np.float32).astype(np.float64), dt)).max() < 1e-4
    # time-varying: residual at the returned x2 with the midpoint time, and the IFT Jacobian by complex step of a Newton solve
    p0 = 1.7
    tv = rd.CustomModel(2, 1, TV_BODY, params=[p0])
    Zt, tt, ht = rng.random((64, 3)), rng.uniform(0, 3, 64), rng.uniform(0.01, 0.1, 64)
    x2 = tv._h.discrete_dynamics(IM, Zt, ht, t=tt)
    Jt = tv._h.discrete_jacobian(IM, Zt, ht, t=tt)
    for k in range(0, 64, 5):
        xm = 0.5 * (Zt[k, :2] + x2[k])
        assert np.abs(Zt[k, :2] + ht[k] * _tv_f(xm, Zt[k, 2:], tt[k] + 0.5 * ht[k], p0) - x2[k]).max() < 1e-11

        def solve(zz):                                    # Newton on the complexified residual (analytic in z)
            x1, u = zz[:2], zz[2:]
            x = x1.copy()
            for _ in range(30):
                r = lambda xx: x1 + ht[k] * _tv_f(0.5 * (x1 + xx), u, tt[k] + 0.5 * ht[k], p0) - xx
                Jr = np.stack([(r(x + 1e-7 * np.eye(2)[j]) - r(x - 1e-7 * np.eye(2)[j])) / 2e-7 for j in range(2)], axis=1)
                x = x - np.linalg.solve(Jr, r(x))
            return x
        assert np.abs(Jt[k].T - _cs_cols(solve, Zt[k])).max() < 1e-7
    assert np.abs(tv._h.dynamics_error(IM, Zt, np.ascontiguousarray(x2), ht, t=tt)).max() < 1e-11
    # 8 states, 2 controls: chain of four coupled pendula-like oscillators -> warp-cooperative kernel (16 lanes per knot)
    body = """
        return vec(get<4>(x), get<5>(x), get<6>(x), get<7>(x),
                   -sin_(get<0>(x)) + p[0] * (get<1>(x) - get<0>(x)) + get<0>(u),
                   -sin_(get<1>(x)) + p[0] * (get<0>(x) - T(2) * get<1>(x) + get<2>(x)),
                   -sin_(get<2>(x)) + p[0] * (get<1>(x) - T(2) * get<2>(x) + get<3>(x)),
                   -sin_(get<3>(x)) + p[0] * (get<2>(x) - get<3>(x)) + get<1>(u) * cos_(get<3>(x)));
    """
    ch = rd.CustomModel(8, 2, body, params=[0.8])
    Zc, hc = rng.random((300, 10)), rng.uniform(0.01, 0.1, 300)
    x2 = ch._h.discrete_dynamics(IM, Zc, hc)
    J2, J1, e = ch._h.dynamics_error(IM, Zc, np.ascontiguousarray(x2), hc, jacobian=True)
    assert np.abs(e).max() < 1e-11                           # the step solves the midpoint equation
    Jc = ch._h.discrete_jacobian(IM, Zc, hc)
    sol = -np.linalg.solve(o.as_matrix(J2)[:, :, :8], o.as_matrix(J1))
    assert np.abs(o.as_matrix(Jc) - sol).max() < 1e-9        # and its Jacobian is the implicit-function-theorem one


def test_implicit_midpoint_for_user_rigid_bodies(rd, torch_):
    """RigidBody{R} with a user wrench under ImplicitMidpoint: a wrench with the built-in structure (forces read the attitude and the
    controls) takes the block-triangular kernel (implicit_block.cuh) and must equal the checker's quadrotor; a wrench that reads the
    POSITION makes v' depend on r, fails that kernel's compile-time structure check and falls back to the dense group kernel — its step
    must solve the midpoint equation and its Jacobian must be the implicit-function-theorem one."""
    from test_abi_host import QUAD_WRENCH
    IM = rd._abi.IMPLICIT_MIDPOINT
    rng = np.random.default_rng(190)
    for Rname, rc, frame in (("QuatRotation", o.ROT_QUAT, o.WORLD), ("MRP", o.ROT_MRP, o.BODYFRAME)):
        om = o.quadrotor(rc, frame)
        um = rd.CustomRigidBody(getattr(rd, Rname), 4, QUAD_WRENCH, mass=0.5, J=(0.0023, 0.0023, 0.004),
                                params=[1.0, 0.0245, 0.175, 0.0, 0.0, -9.81], bodyframe=bool(frame))
        N = 333
        Z, dt = rand_inputs(om.n, om.m, N, rng), rng.uniform(0.005, 0.1, N)
        xn = np.empty((N, om.n))
        J = um._h.discrete_jacobian(IM, Z, dt, xn=xn)
        assert np.abs(J - o.discrete_jacobian(om, o.IMPLICIT_MIDPOINT, Z, dt)).max() < 1e-10
        assert np.abs(xn - o.discrete_dynamics(om, o.IMPLICIT_MIDPOINT, Z, dt)).max() < 1e-10
        J32 = um._h.discrete_jacobian(IM, dev(torch_, Z.astype(np.float32)), dt).cpu().numpy()
        assert np.abs(J32 - o.discrete_jacobian(om, o.IMPLICIT_MIDPOINT, Z.astype(np.float32).astype(np.float64), dt)).max() < 1e-4
    spring = """
                return vec(-p[0] * get<0>(r) + get<0>(u), -p[0] * get<1>(r) + get<1>(u), -p[0] * get<2>(r) + get<2>(u) - T(9.81) * mass,
                           get<3>(u), get<4>(u), get<5>(u));
    """
    sm = rd.CustomRigidBody(rd.QuatRotation, 6, spring, mass=2.0, J=(2.0, 3.0, 1.0), params=[4.0])
    Zs, hs = rand_inputs(13, 6, 200, rng), rng.uniform(0.01, 0.1, 200)
    x2 = sm._h.discrete_dynamics(IM, Zs, hs)
    J2, J1, e = sm._h.dynamics_error(IM, Zs, np.ascontiguousarray(x2), hs, jacobian=True)
    assert np.abs(e).max() < 1e-10
    Js = sm._h.discrete_jacobian(IM, Zs, hs)
    sol = -np.linalg.solve(o.as_matrix(J2)[:, :, :13], o.as_matrix(J1))
    assert np.abs(o.as_matrix(Js) - sol).max() < 1e-8


# ---- general LieState{R,P}: several rotations, any partition (src/liestate.jl:75-132, 210-320) ---------------------------------------
TWO_BODY = """
        // state [w1 (3), q1 (4), s (2), q2 (4), w2 (3)]: two attitudes driven by their body rates, a planar slider driven by u
        auto kin = [&](const auto& q0, const auto& q1, const auto& q2, const auto& q3, const auto& wx, const auto& wy, const auto& wz) {
            return vec(T(-0.5) * (q1 * wx + q2 * wy + q3 * wz), T(0.5) * (q0 * wx + q2 * wz - q3 * wy),
                       T(0.5) * (q0 * wy - q1 * wz + q3 * wx), T(0.5) * (q0 * wz + q1 * wy - q2 * wx));
        };
        auto qa = kin(get<3>(x), get<4>(x), get<5>(x), get<6>(x), get<0>(x), get<1>(x), get<2>(x));
        auto qb = kin(get<9>(x), get<10>(x), get<11>(x), get<12>(x), get<13>(x), get<14>(x), get<15>(x));
        return vec(-p[0] * get<0>(x), get<0>(u) - get<1>(x), -get<2>(x) * get<1>(x),
                   get<0>(qa), get<1>(qa), get<2>(qa), get<3>(qa),
                   get<0>(u) * get<8>(x), get<1>(u) - get<7>(x),
                   get<0>(qb), get<1>(qb), get<2>(qb), get<3>(qb),
                   get<1>(u) * get<3>(x), -get<14>(x), p[1] * get<13>(x) * get<15>(x));
"""


def _two_body_f(z, p):
    x, u = z[:16], z[16:]

    def kin(q, w):
        return 0.5 * np.array([-(q[1] * w[0] + q[2] * w[1] + q[3] * w[2]), q[0] * w[0] + q[2] * w[2] - q[3] * w[1],
                               q[0] * w[1] - q[1] * w[2] + q[3] * w[0], q[0] * w[2] + q[1] * w[1] - q[2] * w[0]])
    return np.concatenate([[-p[0] * x[0], u[0] - x[1], -x[2] * x[1]], kin(x[3:7], x[0:3]), [u[0] * x[8], u[1] - x[7]], kin(x[9:13], x[13:16]),
                           [u[1] * x[3], -x[14], p[1] * x[13] * x[15]]])


def _LH(q):
    w, x, y, z = q / np.linalg.norm(q)
    return np.array([[-x, -y, -z], [w, -z, y], [z, w, -x], [-y, x, w]])


def test_general_liestate_two_rotations(rd, torch_):
    """LieState(QuatRotation, (3, 2, 3)) == the reference's QuatState(16, (4, 10)) example (src/liestate.jl:93-100): errstate_dim 14,
    G = blkdiag(I3, L(q1)H, I2, L(q2)H, I3), state_diff with one Cayley error per rotation, ∇G with one -(q.b) I3 block per rotation, and
    the error-state Jacobian G(x+)' [A B] blkdiag(G(x), I) of a user model living on that state."""
    ls = rd.QuatState(16, (4, 10))
    assert ls.P == (3, 2, 3) and len(ls) == 16
    p = [0.7, 0.3]
    model = rd.CustomLieModel(ls, 2, TWO_BODY, params=p)
    assert rd.dims(model) == (16, 2, 16) and rd.errstate_dim(model) == 14 and rd.jacobian_width(model) == 16
    rng = np.random.default_rng(190)
    N = 150

    def rand_states(M):
        X = rng.random((M, 16))
        for s0 in (3, 9):
            q = rng.standard_normal((M, 4)); X[:, s0:s0 + 4] = q / np.linalg.norm(q, axis=1, keepdims=True)
        return X
    X, X0 = rand_states(N), rand_states(N)
    G = model._h.errstate_jacobian(X)                                      # (N, 14, 16): column-major 16 x 14 per knot
    Gm = o.as_matrix(G)
    for k in range(0, N, 13):
        ref = np.zeros((16, 14))
        ref[0:3, 0:3] = np.eye(3); ref[3:7, 3:6] = _LH(X[k, 3:7]); ref[7:9, 6:8] = np.eye(2); ref[9:13, 8:11] = _LH(X[k, 9:13]); ref[13:16, 11:14] = np.eye(3)
        assert np.abs(Gm[k] - ref).max() < 1e-14
    d = model._h.state_diff(dev(torch_, X), dev(torch_, X0)).cpu().numpy()
    for k in range(0, N, 13):
        ref = np.zeros(14)
        ref[0:3] = X[k, 0:3] - X0[k, 0:3]; ref[6:8] = X[k, 7:9] - X0[k, 7:9]; ref[11:14] = X[k, 13:16] - X0[k, 13:16]
        for s0, e0 in ((3, 3), (9, 8)):
            q, q0 = X[k, s0:s0 + 4], X0[k, s0:s0 + 4]
            Lq0c = np.array([[q0[0], q0[1], q0[2], q0[3]], [-q0[1], q0[0], q0[3], -q0[2]], [-q0[2], -q0[3], q0[0], q0[1]], [-q0[3], q0[2], -q0[1], q0[0]]])
            e = Lq0c @ q                                                   # conj(q0) (x) q
            ref[e0:e0 + 3] = e[1:] / e[0]
        assert np.abs(d[k] - ref).max() < 1e-13 * max(1.0, np.abs(ref).max())       # Cayley errors of near-half-turn rotations are large
    H = o.as_matrix(model._h.grad_errstate_jacobian(X, X0))
    for k in range(0, N, 13):
        ref = np.zeros((14, 14))
        ref[3:6, 3:6] = -np.dot(X[k, 3:7], X0[k, 3:7]) * np.eye(3); ref[8:11, 8:11] = -np.dot(X[k, 9:13], X0[k, 9:13]) * np.eye(3)
        assert np.abs(H[k] - ref).max() < 1e-14
    # discrete Jacobian (plain forward mode: the kernels see a Euclidean user model) and its error-state projection
    Z = np.concatenate([X, rng.random((N, 2))], axis=1)
    h = 0.05
    xn = np.empty((N, 16))
    J = o.as_matrix(model._h.discrete_jacobian(o.RK4, Z, h, xn=xn))
    Jb = o.as_matrix(model._h.discrete_error_jacobian(o.RK4, dev(torch_, Z), h).cpu().numpy())
    Jbh = o.as_matrix(model._h.discrete_error_jacobian(o.RK4, Z, h))        # host pointers
    assert Jb.shape == (N, 14, 16) and np.abs(Jb - Jbh).max() < 1e-15
    for k in range(0, N, 13):
        Jr, xr = _rk4_cs(lambda zz: _two_body_f(zz, p), Z[k], 16, h)
        assert np.abs(J[k] - Jr).max() < 1e-10 and np.abs(xn[k] - xr).max() < 1e-12
        Gx, Gn = o.as_matrix(model._h.errstate_jacobian(Z[k:k + 1, :16].copy()))[0], o.as_matrix(model._h.errstate_jacobian(xn[k:k + 1].copy()))[0]
        ref = np.concatenate([Gn.T @ Jr[:, :16] @ Gx, Gn.T @ Jr[:, 16:]], axis=1)
        assert np.abs(Jb[k] - ref).max() < 1e-10
    # MRP partition with an empty leading block: LieState(MRP, (0, 4)) -> [p (3), v (4)], errstate_dim 7 == n (G is not the identity)
    m2 = rd.CustomLieModel(rd.LieState(rd.MRP, 0, 4), 1, "return vec(get<3>(x), get<4>(x), get<5>(x), -get<0>(x), -get<1>(x), -get<2>(x), get<0>(u));")
    Xm = rng.random((20, 7)) * 0.5
    Gm2 = o.as_matrix(m2._h.errstate_jacobian(Xm))
    for k in range(20):
        pp = Xm[k, :3]
        sk = np.array([[0, -pp[2], pp[1]], [pp[2], 0, -pp[0]], [-pp[1], pp[0], 0]])
        ref = np.eye(7); ref[:3, :3] = (1 - pp @ pp) * np.eye(3) + 2 * (sk + np.outer(pp, pp))
        assert np.abs(Gm2[k] - ref).max() < 1e-14


def test_c_abi_from_plain_c_on_the_gpu(tmp_path):
    """examples/c_abi_example.c — plain C99 against include/rdb200.h, no Python, no torch in the process — computes on the GPU."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "rdb_example")
    libdir = os.path.join(root, "robotdynamics.jl_b200")
    p = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(root, "include"), os.path.join(root, "examples", "c_abi_example.c"),
                        "-o", exe, "-L", libdir, "-lrdb200", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "[A B] of knot 0" in r.stdout, r.stdout + r.stderr


def test_registered_caller_arrays_take_the_host_path(rd, torch_):
    """rdb_host_register: arrays the caller already owns, page-locked in place, give the same results as pageable and pinned ones."""
    om, gm = zoo()["quad_quat_world"][0](), zoo()["quad_quat_world"][1](rd)
    N = 70001
    Z = rand_inputs(om.n, om.m, N, np.random.default_rng(77)).astype(np.float32)
    J_pageable = gm._h.discrete_jacobian(o.RK4, Z, 0.02)
    J = np.empty((N, om.n + om.m, om.n), dtype=np.float32)
    with rd.RegisteredArray(Z), rd.RegisteredArray(J):
        gm._h.discrete_jacobian(o.RK4, Z, 0.02, J=J)
    assert np.array_equal(J, J_pageable)
    assert np.abs(J - o.discrete_jacobian(om, o.RK4, Z.astype(np.float64), 0.02)).max() < TOL[np.float32]
    with pytest.raises(ValueError):
        rd.RegisteredArray(np.empty((0, 3)))


def test_plans_relaunched_and_concurrent(rd, torch_):
    """Plans over several waves of tiles, a ragged tail, per-knot dt arrays: relaunched, marked as sharing the GPU (rdb_plan_set_shared: wide
    tiles) or not, and in flight at once on two streams — every sampled knot against the checker, bit for bit against the direct call."""
    cases = [("cartpole", np.float64, 1 << 18), ("quad_quat_world", np.float32, 150000 + 77), ("body_mrp_body", np.float64, 40000 + 5)]
    streams = [torch_.cuda.Stream(), torch_.cuda.Stream()]
    plans, refs = [], []
    for i, (name, dtype, N) in enumerate(cases):
        om, gm = zoo()[name][0](), zoo()[name][1](rd)
        Z = rand_inputs(om.n, om.m, N, np.random.default_rng(300 + i)).astype(dtype)
        dt = np.random.default_rng(400 + i).uniform(0.005, 0.05, N)
        Zd, dtd = dev(torch_, Z), dev(torch_, dt)
        direct = gm._h.discrete_jacobian(o.RK4, Zd, dtd)
        idx = np.arange(0, N, 13)
        assert np.abs(direct.cpu().numpy()[idx] - o.discrete_jacobian(om, o.RK4, Z[idx].astype(np.float64), dt[idx])).max() < TOL[dtype]
        for shared in (False, True):
            plan = rd._abi.Plan(gm._h, rd._abi.OP_DISCRETE_JACOBIAN, o.RK4, Zd, dtd, shared_gpu=shared)
            for rep in range(2):
                plan.J.zero_()
                plan.launch()
                torch_.cuda.synchronize()
                assert torch_.equal(plan.J, direct), (name, shared, rep)
            plans.append(plan); refs.append(direct)
    for p in plans:
        p.J.zero_()
    torch_.cuda.synchronize()
    for rep in range(3):                                                       # concurrently: plans alternate between two streams
        for k, p in enumerate(plans):
            p.launch(streams[k % 2].cuda_stream)
    torch_.cuda.synchronize()
    for p, r in zip(plans, refs):
        assert torch_.equal(p.J, r)


@pytest.mark.parametrize("name", ["quad_quat_body", "body_quat_body", "quad_mrp_body", "body_rp_body"])
def test_body_frame_split_force_off_the_manifold(rd, torch_, name):
    """Body-frame models evaluate q \\ (q * Fb + Gw) as |q|^4 Fb + q \\ Gw (models.cuh, SPLIT) where the checker composes the two rotations like
    the reference (src/rigidbody.jl:229-230, test/quadrotor.jl:74): the same polynomial in q, so values and Jacobians must agree for state
    quaternions far from unit norm too (|q| in [0.5, 1.5]; the path never renormalises), for every rule, the continuous Jacobian and the
    ImplicitMidpoint rule."""
    om, gm = zoo()[name][0](), zoo()[name][1](rd)
    N = 2000
    rng = np.random.default_rng(500)
    Z = rand_inputs(om.n, om.m, N, rng)
    if om.n == 13:
        Z[:, 3:7] *= rng.uniform(0.5, 1.5, (N, 1))
    if om.m == 4:
        Z[::9, om.n] = 0.0; Z[::11, om.n + 2] = -0.2              # thrust clamp: ties and inactive rotors
    for dtype in (np.float64, np.float32):
        Zt = Z.astype(dtype); Z64 = Zt.astype(np.float64)
        Zd = dev(torch_, Zt)
        for Q in QS:
            J = gm._h.discrete_jacobian(Q, Zd, 0.05)
            torch_.cuda.synchronize()
            ref = o.discrete_jacobian(om, Q, Z64, 0.05)
            assert np.abs(J.cpu().numpy() - ref).max() < TOL[dtype] * max(1.0, np.abs(ref).max())
        Jc = gm._h.jacobian(Zd)
        torch_.cuda.synchronize()
        refc = o.jacobian(om, Z64)
        assert np.abs(Jc.cpu().numpy() - refc).max() < TOL[dtype] * max(1.0, np.abs(refc).max())
    Zd = dev(torch_, Z)
    Ji = gm._h.discrete_jacobian(rd._abi.IMPLICIT_MIDPOINT, Zd, 0.02)
    torch_.cuda.synchronize()
    refi = o.discrete_jacobian(om, o.IMPLICIT_MIDPOINT, Z, 0.02)
    assert np.abs(Ji.cpu().numpy() - refi).max() < 1e-9 * max(1.0, np.abs(refi).max())
