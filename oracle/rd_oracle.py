"""ctypes front-end of the CPU oracle (oracle/rd_oracle.cpp).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product package never does.  Parity status: "parity unpinned" against Julia output
(see the header of rd_oracle.cpp and DESIGN.md §3).

All arrays are fp64, knot-major ("AoS"): Z is (N, n+m); J is (N, n+m, n) in C order, i.e. each knot holds
an n x (n+m) column-major matrix exactly like the reference's DynamicsJacobian (src/jacobian.jl:26-37).
`as_matrix(J)` returns the (N, n, n+m) view for math.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

CARTPOLE, QUADROTOR, BODY, DOUBLE_INTEGRATOR = 0, 1, 2, 3
ROT_NONE, ROT_QUAT, ROT_MRP, ROT_RP = 0, 1, 2, 3
WORLD, BODYFRAME = 0, 1
EULER, RK2, RK3, RK4, IMPLICIT_MIDPOINT = 0, 1, 2, 3, 4
AD, CHAIN = 0, 1


def build(force=False):
    so = os.path.join(_HERE, "librd_oracle.so")
    src = os.path.join(_HERE, "rd_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "librd_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = build()
        try:
            _LIB = ctypes.CDLL(so)
            _LIB.rdo_num_procs()
        except OSError:
            so = build(force=True)
            _LIB = ctypes.CDLL(so)
    return _LIB


class Model:
    """Model descriptor: (kind, rot, frame, params) with the parameter packing of include/rdb200.h."""

    def __init__(self, kind, rot=ROT_NONE, frame=WORLD, params=()):
        self.kind, self.rot, self.frame = int(kind), int(rot), int(frame)
        self.params = np.ascontiguousarray(params, dtype=np.float64)
        n, m, ne = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        rc = lib().rdo_dims(self.kind, self.rot, self.frame, _p(self.params), len(self.params),
                            ctypes.byref(n), ctypes.byref(m), ctypes.byref(ne))
        if rc:
            raise ValueError("bad model descriptor")
        self.n, self.m, self.nerr = n.value, m.value, ne.value

    def _head(self):
        return (self.kind, self.rot, self.frame, _p(self.params), len(self.params))


def cartpole(mc=1.0, mp=0.2, l=0.5, g=9.81):
    """test/cartpole_model.jl:9"""
    return Model(CARTPOLE, params=[mc, mp, l, g])


def quadrotor(rot=ROT_QUAT, frame=WORLD, mass=0.5, J=(0.0023, 0.0023, 0.004), gravity=(0, 0, -9.81),
              motor_dist=0.175, kf=1.0, km=0.0245):
    """test/quadrotor.jl:36-46"""
    J = np.diag(J) if np.ndim(J) == 1 else np.asarray(J, float)
    return Model(QUADROTOR, rot, frame, [mass, *J.reshape(-1), *gravity, motor_dist, kf, km])


def body(rot=ROT_QUAT, frame=WORLD, mass=2.0, J=(2.0, 3.0, 1.0)):
    """test/rigidbody_test.jl:23-56 (mass 2, J = diag(2,3,1)); Satellite = body(mass=1, J=(1,1,1)) (examples/single_satellite.jl:31-35)."""
    J = np.diag(J) if np.ndim(J) == 1 else np.asarray(J, float)
    return Model(BODY, rot, frame, [mass, *J.reshape(-1)])


def satellite(rot=ROT_QUAT, frame=WORLD):
    return body(rot, frame, mass=1.0, J=(1.0, 1.0, 1.0))


def double_integrator(D=1):
    return Model(DOUBLE_INTEGRATOR, params=[D])


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _arr(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def _times(N, t, dt):
    tt = None if t is None else _arr(np.broadcast_to(np.asarray(t, float), (N,)))
    if dt is None:
        return tt, None, 0.0
    if np.ndim(dt) == 0:
        return tt, None, float(dt)
    return tt, _arr(dt, (N,)), 0.0


def as_matrix(J):
    """(N, n+m, n) storage -> (N, n, n+m) math view."""
    return np.swapaxes(J, -1, -2)


def dynamics(model, Z, t=None, nthreads=1):
    Z = _arr(Z, (-1, model.n + model.m)); N = Z.shape[0]
    tt, _, _ = _times(N, t, None)
    out = np.empty((N, model.n))
    rc = lib().rdo_dynamics(*model._head(), ctypes.c_int64(N), _p(Z), _p(tt), _p(out), nthreads)
    assert rc == 0
    return out


def discrete_dynamics(model, Q, Z, dt, t=None, nthreads=1):
    Z = _arr(Z, (-1, model.n + model.m)); N = Z.shape[0]
    tt, dd, dt0 = _times(N, t, dt)
    out = np.empty((N, model.n))
    rc = lib().rdo_discrete_dynamics(*model._head(), Q, ctypes.c_int64(N), _p(Z), _p(tt), _p(dd),
                                     ctypes.c_double(dt0), _p(out), nthreads)
    assert rc == 0
    return out


def jacobian(model, Z, t=None, method=AD, nthreads=1):
    Z = _arr(Z, (-1, model.n + model.m)); N = Z.shape[0]
    tt, _, _ = _times(N, t, None)
    out = np.empty((N, model.n + model.m, model.n))
    rc = lib().rdo_jacobian(*model._head(), method, ctypes.c_int64(N), _p(Z), _p(tt), _p(out), nthreads)
    assert rc == 0
    return out


def discrete_jacobian(model, Q, Z, dt, t=None, method=AD, nthreads=1, out=None):
    Z = _arr(Z, (-1, model.n + model.m)); N = Z.shape[0]
    tt, dd, dt0 = _times(N, t, dt)
    if out is None:
        out = np.empty((N, model.n + model.m, model.n))
    rc = lib().rdo_discrete_jacobian(*model._head(), Q, method, ctypes.c_int64(N), _p(Z), _p(tt), _p(dd),
                                     ctypes.c_double(dt0), _p(out), nthreads)
    assert rc == 0
    return out


def errstate_jacobian(model, X, nthreads=1):
    X = _arr(X); X = X.reshape(-1, X.shape[-1]); N = X.shape[0]
    out = np.empty((N, model.nerr, model.n))
    rc = lib().rdo_errstate_jacobian(*model._head(), ctypes.c_int64(N), _p(X), X.shape[1], _p(out), nthreads)
    assert rc == 0
    return out


def grad_errstate_jacobian(model, X, B, nthreads=1):
    X = _arr(X); X = X.reshape(-1, X.shape[-1]); N = X.shape[0]
    B = _arr(B, (N, model.n))
    out = np.empty((N, model.nerr, model.nerr))
    rc = lib().rdo_grad_errstate_jacobian(*model._head(), ctypes.c_int64(N), _p(X), X.shape[1], _p(B), _p(out), nthreads)
    assert rc == 0
    return out


def state_diff(model, X, X0, nthreads=1):
    X = _arr(X); X = X.reshape(-1, X.shape[-1]); N = X.shape[0]
    X0 = _arr(X0); X0 = X0.reshape(-1, X0.shape[-1])
    out = np.empty((N, model.nerr))
    rc = lib().rdo_state_diff(*model._head(), ctypes.c_int64(N), _p(X), X.shape[1], _p(X0), X0.shape[1], _p(out), nthreads)
    assert rc == 0
    return out


def rollout(model, Q, x0, U, dt, t=None, nthreads=1):
    """x0: (ntraj, n); U: (ntraj, K-1, m); dt scalar or (ntraj, K). Returns X (ntraj, K, n)."""
    x0 = _arr(x0, (-1, model.n)); ntraj = x0.shape[0]
    U = _arr(U, (ntraj, -1, model.m)); K = U.shape[1] + 1
    tt = None if t is None else _arr(t, (ntraj, K))
    if np.ndim(dt) == 0:
        dd, dt0 = None, float(dt)
    else:
        dd, dt0 = _arr(dt, (ntraj, K)), 0.0
    X = np.empty((ntraj, K, model.n))
    rc = lib().rdo_rollout(*model._head(), Q, ctypes.c_int64(ntraj), K, _p(x0), _p(U), _p(tt), _p(dd),
                           ctypes.c_double(dt0), _p(X), nthreads)
    assert rc == 0
    return X


def num_procs():
    return lib().rdo_num_procs()
