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
