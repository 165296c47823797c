"""Small run for compute-sanitizer (memcheck / racecheck / synccheck): a few tiles of every kernel family, checked against the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import rdb200 as rd
from oracle import rd_oracle as o

rng = np.random.default_rng(0)
def rigid(n, m, N):
    Z = rng.random((N, n + m)); q = rng.standard_normal((N, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
    if n == 13: Z[:, 3:7] = q
    else: Z[:, 3:6] = q[:, 1:] / (1 + np.abs(q[:, :1]))
    return Z
cases = [("cartpole f64", rd.Cartpole(), o.cartpole(), rng.random((700, 5)), np.float64, 1e-10),
         ("quadrotor f32", rd.Quadrotor(), o.quadrotor(), rigid(13, 4, 700), np.float32, 1e-4),
         ("satellite mrp f64", rd.Satellite(rd.MRP), o.satellite(o.ROT_MRP), rigid(12, 6, 300), np.float64, 1e-10)]
for name, gm, om, Z, dt, tol in cases:
    Zd = torch.from_numpy(Z.astype(dt)).cuda()
    xn = torch.empty((Z.shape[0], om.n), dtype=Zd.dtype, device="cuda")
    for Q in (o.RK4, o.RK2):
        J = gm._h.discrete_jacobian(Q, Zd, 0.01, xn=xn)
        torch.cuda.synchronize()
        err = np.abs(J.cpu().numpy() - o.discrete_jacobian(om, Q, Z.astype(dt).astype(np.float64), 0.01)).max()
        assert err < tol, (name, Q, err)
    if om.n >= 12:
        Jb = gm._h.discrete_error_jacobian(o.RK4, Zd, 0.01); G = gm._h.errstate_jacobian(Zd); d = gm._h.state_diff(Zd, Zd.flip(0).contiguous())
        torch.cuda.synchronize()
    Jh = gm._h.discrete_jacobian(o.RK4, Z.astype(dt), 0.01)            # host path
    print("ok", name, flush=True)
print("SANITY DONE")

# ---- round 2 additions: component-major (tensor-map) kernels, ImplicitMidpoint (both kernels), dynamics_error, the device trajectory
#      (rollout in place, linearize, pipelined rollout + linearize on two streams), plans, a time-varying user model
cp, qd = rd.Cartpole(), rd.Quadrotor()
Zs = torch.from_numpy(np.ascontiguousarray(rng.random((640, 5)).T)).cuda()
cp._h.discrete_jacobian(o.RK4, Zs, 0.01, layout=rd.SOA)
for gm, Z in ((cp, rng.random((300, 5))), (qd, rigid(13, 4, 200))):
    Zd = torch.from_numpy(Z).cuda()
    gm._h.discrete_jacobian(rd._abi.IMPLICIT_MIDPOINT, Zd, 0.03)
    x2 = gm._h.discrete_dynamics(rd._abi.IMPLICIT_MIDPOINT, Zd, 0.03)
    gm._h.dynamics_error(rd._abi.IMPLICIT_MIDPOINT, Zd, x2, 0.03, jacobian=True)
    gm._h.dynamics_error(o.RK4, Zd, x2, 0.03, jacobian=True)
torch.cuda.synchronize()
for gm, dtn in ((cp, np.float64), (qd, np.float32)):
    dm = rd.DiscretizedDynamics(gm, rd.RK4)
    h = gm._h
    tb = rd.TrajectoryBatch(dm, 70, 19, dtn)
    tb.set_initial_state(rigid(13, 4, 70)[:, :13].astype(dtn) if h.n == 13 else rng.random((70, 4)))
    tb.set_controls((0.5 * rng.random((18, 70, h.m))).astype(dtn)); tb.set_timesteps(0.02)
    tb.rollout(); tb.linearize(); tb.rollout_linearize(chunks=4); tb.rollout_linearize(error_state=True, device=False); tb.states()
    torch.cuda.synchronize()
Zq = torch.from_numpy(rigid(13, 4, 500).astype(np.float32)).cuda()
plan = rd._abi.Plan(qd._h, rd._abi.OP_DISCRETE_ERROR_JACOBIAN, rd.RK4.code, Zq, 0.01)
plan.launch(); plan.launch()
tv = rd.CustomModel(2, 1, "return vec(get<1>(x), cos_(T(3) * t) * get<0>(u) - p[0] * sin_(get<0>(x)) + t * get<1>(x));", params=[1.7])
tv._h.discrete_jacobian(o.RK4, rng.random((333, 3)), 0.05, t=rng.random(333))
tv._h.discrete_jacobian(rd._abi.IMPLICIT_MIDPOINT, torch.from_numpy(rng.random((333, 3))).cuda(), 0.05, t=rng.random(333))
torch.cuda.synchronize()
print("SANITY ROUND-2 DONE")

# ---- late round 2: wide-tile kernels above the small-batch threshold (launch.cuh: RDB_SMALL_N), body-frame models with the split force,
#      a per-knot dt array riding on the tile's TMA transaction
for gm, om, n, m in ((rd.Quadrotor(), o.quadrotor(), 13, 4), (rd.Quadrotor(bodyframe=True), o.quadrotor(o.ROT_QUAT, o.BODYFRAME), 13, 4),
                     (rd.Body(rd.MRP, bodyframe=True), o.body(o.ROT_MRP, o.BODYFRAME), 12, 6)):
    Nb = 8192 + 411
    Z = rigid(n, m, Nb).astype(np.float32)
    dtv = rng.uniform(0.005, 0.05, Nb)
    J = gm._h.discrete_jacobian(o.RK4, torch.from_numpy(Z).cuda(), dtv)
    Jb = gm._h.discrete_error_jacobian(o.RK4, torch.from_numpy(Z).cuda(), dtv)
    torch.cuda.synchronize()
    idx = np.arange(0, Nb, 97)
    assert np.abs(J.cpu().numpy()[idx] - o.discrete_jacobian(om, o.RK4, Z[idx].astype(np.float64), dtv[idx])).max() < 1e-4
print("SANITY LATE ROUND-2 DONE")
