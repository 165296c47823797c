"""Quick GPU sanity + timing sweep (development aid; not part of the test-suite)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import rdb200 as rd
from oracle import rd_oracle as o

def rand_inputs(model, N, rng, quat=True):
    n, m = model.n, model.m
    Z = rng.random((N, n + m))
    if n >= 12:
        np_ = n - 9
        q = rng.standard_normal((N, 4)); q /= np.linalg.norm(q, axis=1, keepdims=True)
        if np_ == 4: Z[:, 3:7] = q
        else: Z[:, 3:6] = q[:, 1:] / (1 + np.abs(q[:, :1]))
    return Z

def check(name, gm, om, Q, dtype, N=5000, dt=0.01):
    rng = np.random.default_rng(1)
    Z = rand_inputs(gm._h, N, rng).astype(dtype)
    Zd = torch.from_numpy(Z).cuda()
    xn = torch.empty((N, gm._h.n), dtype=Zd.dtype, device='cuda')
    J = gm._h.discrete_jacobian(Q, Zd, dt, xn=xn)
    torch.cuda.synchronize()
    Jo = o.discrete_jacobian(om, Q, Z.astype(np.float64), dt)
    xo = o.discrete_dynamics(om, Q, Z.astype(np.float64), dt)
    eJ = np.abs(J.cpu().numpy() - Jo).max(); ex = np.abs(xn.cpu().numpy() - xo).max()
    # host path
    Jh = gm._h.discrete_jacobian(Q, Z, dt)
    eh = np.abs(Jh - Jo).max()
    print(f"{name:28s} Q={Q} {np.dtype(dtype).name}: max|dJ|={eJ:.3e} max|dx+|={ex:.3e} host-path {eh:.3e}", flush=True)

def bench(name, gm, Q, dtype, N, dt=0.01, iters=20):
    rng = np.random.default_rng(2)
    Z = torch.from_numpy(rand_inputs(gm._h, N, rng).astype(dtype)).cuda()
    n, m = gm._h.n, gm._h.m
    J = torch.empty((N, n + m, n), dtype=Z.dtype, device='cuda')
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    for _ in range(3): gm._h.discrete_jacobian(Q, Z, dt, J=J)
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gm._h.discrete_jacobian(Q, Z, dt, J=J); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = np.median(ts) * 1e-3
    es = Z.element_size()
    gbs = N * es * ((n + m) + n * (n + m)) / t / 1e9
    print(f"BENCH {name:24s} Q={Q} {np.dtype(dtype).name} N={N}: {t*1e6:9.1f} us  {N/t:.3e} evals/s  {gbs:7.1f} GB/s algorithmic", flush=True)

def bench_err(name, gm, Q, dtype, N, dt=0.01, iters=20):
    rng = np.random.default_rng(2)
    Z = torch.from_numpy(rand_inputs(gm._h, N, rng).astype(dtype)).cuda()
    ne, m = gm._h.nerr, gm._h.m
    J = torch.empty((N, ne + m, ne), dtype=Z.dtype, device='cuda')
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    for _ in range(3): gm._h.discrete_error_jacobian(Q, Z, dt, J=J)
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gm._h.discrete_error_jacobian(Q, Z, dt, J=J); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = np.median(ts) * 1e-3
    es = Z.element_size()
    gbs = N * es * ((gm._h.n + m) + ne * (ne + m)) / t / 1e9
    print(f"BENCH-ERR {name:20s} Q={Q} {np.dtype(dtype).name} N={N}: {t*1e6:9.1f} us  {N/t:.3e} evals/s  {gbs:7.1f} GB/s algorithmic", flush=True)


def bench_value(name, gm, Q, dtype, N, dt=0.01, iters=20):
    """dynamics() (Q < 0), discrete_dynamics() (value-only kernels K1/K2) and the continuous jacobian!() (Q = None)."""
    rng = np.random.default_rng(2)
    Z = torch.from_numpy(rand_inputs(gm._h, N, rng).astype(dtype)).cuda()
    n, m = gm._h.n, gm._h.m
    if Q is None:
        J = torch.empty((N, n + m, n), dtype=Z.dtype, device='cuda')
        f = lambda: gm._h.jacobian(Z, J=J)
        nbytes, label = (n + m) + n * (n + m), "jacobian! (continuous)"
    else:
        out = torch.empty((N, n), dtype=Z.dtype, device='cuda')
        f = (lambda: gm._h.dynamics(Z, out=out)) if Q < 0 else (lambda: gm._h.discrete_dynamics(Q, Z, dt, out=out))
        nbytes, label = (n + m) + n, "dynamics" if Q < 0 else f"discrete_dynamics Q={Q}"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    for _ in range(3): f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = np.median(ts) * 1e-3
    gbs = N * Z.element_size() * nbytes / t / 1e9
    print(f"BENCH-VAL {name:20s} {label:24s} {np.dtype(dtype).name} N={N}: {t*1e6:9.1f} us  {N/t:.3e} evals/s  {gbs:7.1f} GB/s algorithmic", flush=True)


def bench_soa(name, gm, Q, dtype, N, dt=0.01, iters=10):
    rng = np.random.default_rng(2)
    n, m = gm._h.n, gm._h.m
    Z = torch.from_numpy(np.ascontiguousarray(rand_inputs(gm._h, N, rng).astype(dtype).T)).cuda()
    J = torch.empty((n * (n + m), N), dtype=Z.dtype, device='cuda')
    for _ in range(3): gm._h.discrete_jacobian(Q, Z, dt, J=J, layout=rd.SOA)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): gm._h.discrete_jacobian(Q, Z, dt, J=J, layout=rd.SOA)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / iters * 1e-3
    gbs = N * Z.element_size() * ((n + m) + n * (n + m)) / t / 1e9
    print(f"BENCH-SOA {name:20s} Q={Q} {np.dtype(dtype).name} N={N}: {t*1e6:9.1f} us  {N/t:.3e} evals/s  {gbs:7.1f} GB/s algorithmic", flush=True)


def bench_misc():
    """LieState kernels and rollout: device-resident timings."""
    qd = rd.Quadrotor()
    rng = np.random.default_rng(3)
    N = 1 << 20
    X = torch.from_numpy(rand_inputs(qd._h, N, rng).astype(np.float32)).cuda()
    G = torch.empty((N, 12, 13), dtype=torch.float32, device='cuda')
    d = torch.empty((N, 12), dtype=torch.float32, device='cuda')
    def timeit(f, iters=20):
        for _ in range(3): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters): f()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters * 1e-3
    t = timeit(lambda: qd._h.errstate_jacobian(X, G=G))
    print(f"BENCH-MISC errstate_jacobian fp32 N={N}: {t*1e6:.1f} us  {N*4*(4+156)/t/1e9:.0f} GB/s (read q 16 B + write G 624 B per knot)", flush=True)
    X0 = X.flip(0).contiguous()
    t = timeit(lambda: qd._h.state_diff(X, X0, dX=d))
    print(f"BENCH-MISC state_diff fp32 N={N}: {t*1e6:.1f} us  {N*4*(34+12)/t/1e9:.0f} GB/s", flush=True)
    ntraj, K = 4096, 256
    x0 = X[:ntraj, :13].contiguous(); U = torch.rand((ntraj, K - 1, 4), dtype=torch.float32, device='cuda')
    Xo = torch.empty((ntraj, K, 13), dtype=torch.float32, device='cuda')
    t = timeit(lambda: qd._h.rollout(3, x0, U, 0.01, X=Xo), iters=5)
    print(f"BENCH-MISC rollout quadrotor RK4 fp32 {ntraj} x {K}: {t*1e3:.2f} ms  {ntraj*(K-1)/t:.3e} steps/s", flush=True)
    cp = rd.Cartpole()
    x0c = torch.rand((ntraj, 4), dtype=torch.float64, device='cuda'); Uc = torch.rand((ntraj, K - 1, 1), dtype=torch.float64, device='cuda')
    t = timeit(lambda: cp._h.rollout(3, x0c, Uc, 0.01), iters=5)
    print(f"BENCH-MISC rollout cartpole RK4 fp64 {ntraj} x {K}: {t*1e3:.2f} ms  {ntraj*(K-1)/t:.3e} steps/s", flush=True)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    cp, cpo = rd.Cartpole(), o.cartpole()
    qd, qdo = rd.Quadrotor(), o.quadrotor()
    for Q in (0, 1, 2, 3):
        check("cartpole", cp, cpo, Q, np.float64)
    check("cartpole", cp, cpo, 3, np.float32)
    check("cartpole ragged", cp, cpo, 3, np.float64, N=1000 + 37)
    for Q in (0, 1, 2, 3):
        check("quadrotor quat world", qd, qdo, Q, np.float32)
    check("quadrotor quat world", qd, qdo, 3, np.float64)
    check("quadrotor mrp world", rd.Quadrotor(rd.MRP), o.quadrotor(o.ROT_MRP), 3, np.float64)
    check("quadrotor rp body", rd.Quadrotor(rd.RodriguesParam, bodyframe=True), o.quadrotor(o.ROT_RP, o.BODYFRAME), 3, np.float64)
    check("body quat body", rd.Body(bodyframe=True), o.body(o.ROT_QUAT, o.BODYFRAME), 3, np.float64)
    check("satellite mrp", rd.Satellite(rd.MRP), o.satellite(o.ROT_MRP), 1, np.float64, dt=0.1)
    check("double integrator 3", rd.DoubleIntegrator(3), o.double_integrator(3), 3, np.float64)
    bench_misc()
    for Q in (-1, 3, None):
        bench_value("cartpole", cp, Q, np.float64, 1 << 22)
        bench_value("quadrotor", qd, Q, np.float32, 1 << 21)
    bench("cartpole implicit-midpoint", cp, 4, np.float64, 1 << 20)
    bench("quadrotor implicit-midpoint", qd, 4, np.float32, 262144)
    bench("quadrotor implicit-midpoint", qd, 4, np.float64, 262144)
    bench_soa("cartpole", cp, 3, np.float64, 1 << 20)
    bench_soa("quadrotor", qd, 3, np.float32, 262144)
    bench_err("quadrotor", qd, 3, np.float32, 262144)
    bench_err("quadrotor", qd, 3, np.float64, 262144)
    bench_err("satellite mrp rk2", rd.Satellite(rd.MRP), 1, np.float64, 1 << 20, dt=0.1)
    bench("cartpole", cp, 3, np.float64, 1 << 20)
    bench("cartpole", cp, 3, np.float64, 1 << 23)
    bench("cartpole", cp, 3, np.float32, 1 << 22)
    bench("cartpole rk3", cp, 2, np.float64, 1 << 20)
    bench("quadrotor", qd, 3, np.float32, 262144)
    bench("quadrotor", qd, 3, np.float32, 1 << 21)
    bench("quadrotor", qd, 3, np.float64, 262144)
    bench("quadrotor rk3", qd, 2, np.float32, 262144)
    bench("quadrotor rk2", qd, 1, np.float32, 262144)
    bench("quadrotor mrp", rd.Quadrotor(rd.MRP), 3, np.float32, 262144)
    bench("quadrotor body-frame", rd.Quadrotor(bodyframe=True), 3, np.float32, 262144)
    bench("body quat", rd.Body(), 3, np.float32, 262144)
    bench("body quat", rd.Body(), 3, np.float64, 262144)
    bench("satellite mrp rk2", rd.Satellite(rd.MRP), 1, np.float64, 1 << 20, dt=0.1)
    bench("satellite mrp rk2", rd.Satellite(rd.MRP), 1, np.float32, 1 << 20, dt=0.1)
    bench("satellite mrp rk4", rd.Satellite(rd.MRP), 3, np.float64, 1 << 18, dt=0.1)
