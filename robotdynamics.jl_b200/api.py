"""Host-side mirror of the RobotDynamics.jl interface for the hot path (Python stands in for the Julia shim that
cannot run in this image; `julia/RobotDynamicsB200.jl` is the same thing spelled with `ccall`).

Names, argument order and error behaviour follow the reference; Julia's `f!` is spelled `f_`:

    jacobian_(sig, diff, model, J, y, z)          jacobian!(sig, diff, fun, J, y, z)     src/functionbase.jl:242
    discrete_jacobian_(Q, J, model, z)            v0.3 discrete_jacobian!(Q, ∇f, model, z)  README.md:81-82
    dynamics(model, z) / dynamics(model, x, u[, t])                                       src/dynamics.jl:81-83
    discrete_dynamics(dmodel, z) / (dmodel, x, u, t, dt) / (Q, model, z)                  src/discrete_dynamics.jl:80-81
    errstate_jacobian_(model, G, z) / ∇errstate_jacobian_ / state_diff                    src/statevectortype.jl:89-141
    rollout_(sig, dmodel, Z, x0)                                                          src/trajectories.jl:436-441

Every function takes ONE knot point (API fidelity) or a whole batch — a `SampledTrajectory`, or raw arrays /
CUDA tensors — and evaluates the batch in a single C-ABI call.  Nothing here computes: all arithmetic happens in
librdb200.so on the GPU.
"""
import numpy as np

from . import _abi
from ._abi import AOS, SOA, NotImplementedModelError  # noqa: F401


# ---------------------------------------------------------------------------------------------------
# traits (reference: src/functionbase.jl:78-120, src/integration.jl:69,109,258, src/statevectortype.jl:63-73)
# ---------------------------------------------------------------------------------------------------
class FunctionSignature: pass
class InPlace(FunctionSignature): pass
class StaticReturn(FunctionSignature): pass

class DiffMethod: pass
class ForwardAD(DiffMethod): pass
class FiniteDifference(DiffMethod): pass
class UserDefined(DiffMethod): pass
class B200(DiffMethod):
    """The extension point the reference documents ("Users are free to add more DiffMethod types",
    docs/src/autodiff.md:20-21): batched forward-mode evaluation on the GPU."""

class QuadratureRule:
    code = None
class Euler(QuadratureRule): code = _abi.EULER
class RK2(QuadratureRule): code = _abi.RK2
class RK3(QuadratureRule): code = _abi.RK3
class RK4(QuadratureRule): code = _abi.RK4
class ImplicitMidpoint(QuadratureRule): code = _abi.IMPLICIT_MIDPOINT

class QuatRotation: code = _abi.ROT_QUAT
class UnitQuaternion(QuatRotation): pass
class MRP: code = _abi.ROT_MRP
class RodriguesParam: code = _abi.ROT_RP

class EuclideanState: pass
class RotationState: pass


def _qcode(Q):
    if isinstance(Q, int):
        return Q
    if isinstance(Q, QuadratureRule) or (isinstance(Q, type) and issubclass(Q, QuadratureRule)):
        return Q.code
    raise TypeError(f"not a QuadratureRule: {Q!r}")


# ---------------------------------------------------------------------------------------------------
# models (reference: src/dynamics.jl:2,68; test/cartpole_model.jl; test/quadrotor.jl; test/rigidbody_test.jl:23-56;
#         examples/single_satellite.jl:7-35; test/double_integrator.jl:97-127)
# ---------------------------------------------------------------------------------------------------
class AbstractModel:
    """ContinuousDynamics model bound to a GPU through an rdb_model handle."""
    statevectortype = EuclideanState

    def __init__(self, kind, rot, frame, params, device=None):
        self._h = _abi.ModelHandle(kind, rot, frame, params, device)

    @property
    def n(self): return self._h.n
    @property
    def m(self): return self._h.m


class ContinuousDynamics(AbstractModel): pass


class Cartpole(ContinuousDynamics):
    def __init__(self, mc=1.0, mp=0.2, l=0.5, g=9.81, device=None):
        super().__init__(_abi.CARTPOLE, _abi.ROT_NONE, 0, [mc, mp, l, g], device)


class DoubleIntegrator(ContinuousDynamics):
    def __init__(self, D=1, device=None):
        super().__init__(_abi.DOUBLE_INTEGRATOR, _abi.ROT_NONE, 0, [D], device)


class RigidBody(ContinuousDynamics):
    """RigidBody{R} <: LieGroupModel; LieState(R, (3, 6))  (reference: src/rigidbody.jl:46-48)."""
    statevectortype = RotationState


def _inertia(J):
    J = np.asarray(J, dtype=np.float64)
    return np.diag(J) if J.ndim == 1 else J


class Quadrotor(RigidBody):
    def __init__(self, R=QuatRotation, mass=0.5, J=(0.0023, 0.0023, 0.004), gravity=(0.0, 0.0, -9.81), motor_dist=0.175,
                 kf=1.0, km=0.0245, bodyframe=False, device=None):
        super().__init__(_abi.QUADROTOR, R.code, int(bool(bodyframe)),
                         [mass, *_inertia(J).reshape(-1), *gravity, motor_dist, kf, km], device)


class Body(RigidBody):
    """test/rigidbody_test.jl:23-56: F_world = q*u[1:3], M_body = u[4:6], mass 2, J = diag(2,3,1)."""
    def __init__(self, R=QuatRotation, mass=2.0, J=(2.0, 3.0, 1.0), bodyframe=False, device=None):
        super().__init__(_abi.BODY, R.code, int(bool(bodyframe)), [mass, *_inertia(J).reshape(-1)], device)


class Satellite(Body):
    """examples/single_satellite.jl:7-35: the same wrench with mass 1, J = I."""
    def __init__(self, R=QuatRotation, bodyframe=False, device=None):
        super().__init__(R, 1.0, (1.0, 1.0, 1.0), bodyframe, device)


class CustomModel(ContinuousDynamics):
    """Any user `dynamics(model, x, u)`: `body` is the CUDA C++ body of `f(x, u)` written against csrc/sdual.cuh
    (get<i>(x), get<j>(u), p[k], T(c), sin_/cos_/exp_/sqrt_/relu_, `return vec(...)`), compiled with NVRTC and differentiated
    by the same forward-mode engine — the GPU counterpart of putting `@autodiff` on a model (src/jacobian_gen.jl:64-82)."""
    def __init__(self, n, m, body, params=(), device=None):
        self._h = _abi.ModelHandle(_abi.CUSTOM, _abi.ROT_NONE, 0, params, device, custom=(n, m, body))


class CustomRigidBody(RigidBody):
    """RigidBody{R} with user forces / moments (the reference's rigid-body extension interface, src/rigidbody.jl:244-257):
    `wrench_body` sees q, r, v, w, u, p[k], mass and returns vec(F_world (3), tau_body (3)); the library supplies the rest,
    including the LieState error maps and the error-state Jacobian."""
    def __init__(self, R, m, wrench_body, mass, J, params=(), bodyframe=False, device=None):
        self._h = _abi.ModelHandle(_abi.CUSTOM, R.code, int(bool(bodyframe)), params, device, custom=(m, wrench_body, mass, _inertia(J)))


class LieState:
    """LieState{R,P} (src/liestate.jl:75-132): vector blocks of lengths P with one rotation of type R between consecutive blocks."""
    def __init__(self, R, *P):
        self.R, self.P = R, tuple(int(p) for p in (P[0] if len(P) == 1 and not isinstance(P[0], int) else P))

    def __len__(self):
        return sum(self.P) + (len(self.P) - 1) * (4 if self.R.code == _abi.ROT_QUAT else 3)


def QuatState(n, Q):
    """QuatState(n, Q): LieState of QuatRotations whose first (1-based) indices are Q (src/liestate.jl:84-111)."""
    Q = list(Q)
    P = [Q[0] - 1] + [Q[i] - Q[i - 1] - 4 for i in range(1, len(Q))] + [n - (Q[-1] + 4) + 1]
    return LieState(QuatRotation, *P)


class CustomLieModel(ContinuousDynamics):
    """A user model (body of `f(x, u[, t])`, like CustomModel) whose state vector is a general LieState{R,P}: several rotations, any
    partition.  errstate_jacobian_, grad_errstate_jacobian_, state_diff and discrete_error_jacobian_ follow the partition."""
    statevectortype = RotationState

    def __init__(self, liestate, m, body, params=(), device=None):
        self.liestate = liestate
        self._h = _abi.ModelHandle(_abi.CUSTOM, liestate.R.code, 0, params, device, custom=(liestate.P, m, body, "lie"))


class DiscreteDynamics(AbstractModel): pass


class DiscretizedDynamics(DiscreteDynamics):
    """DiscretizedDynamics{L,Q}(model)  (reference: src/discretized_dynamics.jl:169-191)."""
    def __init__(self, model, Q=RK4):
        self.continuous_dynamics, self.integrator = model, Q
        self._h, self.statevectortype = model._h, model.statevectortype


def state_dim(f): return f._h.n if hasattr(f, "_h") else f.n
def control_dim(f): return f._h.m if hasattr(f, "_h") else f.m
def errstate_dim(f): return f._h.nerr
def output_dim(f): return state_dim(f)
def jacobian_width(f): return errstate_dim(f) + control_dim(f)          # src/functionbase.jl:135
def dims(f): return state_dim(f), control_dim(f), output_dim(f)
def default_diffmethod(f): return ForwardAD() if isinstance(f, DiscretizedDynamics) else UserDefined()
def default_signature(f): return StaticReturn()


# ---------------------------------------------------------------------------------------------------
# containers (reference: src/knotpoint.jl:146-219, src/trajectories.jl:40-50, src/jacobian.jl:26-43)
# ---------------------------------------------------------------------------------------------------
class KnotPoint:
    """z = [x;u], t, dt.  KnotPoint(x, u, t, dt) or KnotPoint(n, m, z, t, dt)."""

    def __init__(self, *args):
        if len(args) == 4:
            x, u, t, dt = args
            x, u = np.asarray(x), np.asarray(u)
            self.n, self.m, self.z = len(x), len(u), np.concatenate([x, u])
        elif len(args) == 5:
            n, m, z, t, dt = args
            assert n > 0 and m > 0 and n + m == len(z)
            self.n, self.m, self.z = int(n), int(m), np.array(z)
        else:
            raise TypeError("KnotPoint(x, u, t, dt) or KnotPoint(n, m, z, t, dt)")
        self.t, self.dt = float(t), float(dt)

StaticKnotPoint = KnotPoint


def state(z): return z.z[:z.n]
def control(z): return z.z[z.n:] * (0.0 if is_terminal(z) else 1.0)     # zeros at a terminal knot (src/knotpoint.jl:67)
def getdata(z): return z.z
def time(z): return z.t
def timestep(z): return z.dt
def is_terminal(z): return z.dt == 0.0


class SampledTrajectory:
    """A vector of knot points, stored as the batched arrays the GPU consumes: data (N, n+m), times (N,), dt (N,)."""

    def __init__(self, X, U=None, dt=None, t0=0.0, dtype=np.float64):
        if U is None:                       # list of KnotPoints
            kps = list(X)
            self.n, self.m = kps[0].n, kps[0].m
            self.data = np.ascontiguousarray(np.stack([k.z for k in kps]), dtype=dtype)
            self.times = np.array([k.t for k in kps], dtype=np.float64)
            self.dts = np.array([k.dt for k in kps], dtype=np.float64)
        else:
            X, U = np.asarray(X), np.asarray(U)
            N = X.shape[0]
            self.n, self.m = X.shape[1], U.shape[1]
            Uf = np.zeros((N, self.m))
            Uf[:U.shape[0]] = U
            self.data = np.ascontiguousarray(np.concatenate([X, Uf], axis=1), dtype=dtype)
            self.dts = np.broadcast_to(np.asarray(dt, dtype=np.float64), (N,)).copy()
            if U.shape[0] == N - 1:
                self.dts[-1] = 0.0          # no terminal control -> terminal dt = 0 (src/trajectories.jl:82-83,110)
            self.times = t0 + np.concatenate([[0.0], np.cumsum(self.dts[:-1])])

    def __len__(self): return self.data.shape[0]
    def __getitem__(self, k): return KnotPoint(self.n, self.m, self.data[k], self.times[k], self.dts[k])
    def __iter__(self): return (self[k] for k in range(len(self)))


class DeviceTrajectory:
    """Persistent device mirror of a SampledTrajectory — an `rdb_trajectory` of the C ABI (include/rdb200.h):
    data, times and steps live in HBM, so repeated linearisations never re-gather the host's Vector{KnotPoint} (an array of pointers
    to mutable structs, src/knotpoint.jl:213-217) and never move the inputs across PCIe.  setstates_/setcontrols_ are the reference's
    setstates!/setcontrols! (src/trajectories.jl:215-250) as H2D column-block updates; jacobian_ / discrete_error_jacobian_ /
    discrete_dynamics accept it wherever a SampledTrajectory is accepted; rollout_ rolls it out on the device.
    `data` (K, n+m), `dts`, `times_dev` are zero-copy torch views of the mirror."""

    def __init__(self, model, Z, dtype=None):
        h = model._h
        self.model, self.n, self.m = model, Z.n, Z.m
        if (h.n, h.m) != (Z.n, Z.m):
            raise ValueError(f"trajectory dimensions {(Z.n, Z.m)} do not match the model's {(h.n, h.m)}")
        dtype = Z.data.dtype if dtype is None else np.dtype(dtype)
        K = len(Z)
        self._t = _abi.Trajectory(h, 1, K, dtype)
        self._t.set_states(np.ascontiguousarray(Z.data[:, None, :Z.n], dtype=dtype))
        self._t.set_controls(np.ascontiguousarray(Z.data[:, None, Z.n:], dtype=dtype))
        self._t.set_timesteps(np.ascontiguousarray(Z.dts[:, None]), t0=float(Z.times[0]))
        Zv, tv, dv = self._t.views()
        self.data, self.times_dev, self.dts = Zv[:, 0, :], tv[:, 0], dv[:, 0]
        self.times = Z.times.copy()

    def __len__(self): return self.data.shape[0]

    def to_host(self):
        Z = SampledTrajectory.__new__(SampledTrajectory)
        Z.n, Z.m, Z.data, Z.dts, Z.times = self.n, self.m, self.data.cpu().numpy(), self.dts.cpu().numpy(), self.times_dev.cpu().numpy()
        return Z


def states(Z): return Z.data[:, :Z.n]
def controls(Z): return Z.data[:, Z.n:]
def gettimes(Z): return Z.times
def setstates_(Z, X):
    """setstates!(Z, X)  (src/trajectories.jl:215-230); on a DeviceTrajectory one H2D copy + a column-block update."""
    if isinstance(Z, DeviceTrajectory):
        X = X if _abi._is_torch(X) else np.asarray(X)
        Z._t.set_states(X.reshape(len(Z), 1, Z.n))
    else:
        Z.data[:, :Z.n] = X


def setcontrols_(Z, U):
    """setcontrols!(Z, U)  (src/trajectories.jl:232-250): K-1 controls (the terminal one stays zero) or K."""
    if isinstance(Z, DeviceTrajectory):
        U = U if _abi._is_torch(U) else np.asarray(U)
        Z._t.set_controls(U.reshape(U.shape[0], 1, Z.m))
    else:
        Z.data[:len(U), Z.n:] = U


class DynamicsJacobian:
    """n x (n+m) column-major [A B] with .A / .B views (reference: src/jacobian.jl:26-37).  `data` is stored as the
    (n+m, n) C-order array whose memory equals Julia's column-major Matrix{T}(n, n+m)."""

    def __init__(self, n, m=None, dtype=np.float64, data=None):
        if m is None and hasattr(n, "_h"):
            n, m = state_dim(n), control_dim(n)
        self.n, self.m = n, m
        self.data = np.zeros((n + m, n), dtype=dtype) if data is None else data

    @property
    def A(self): return self.data[:self.n].T
    @property
    def B(self): return self.data[self.n:].T
    def matrix(self): return self.data.T
    def __array__(self, dtype=None, copy=None): return np.asarray(self.data.T, dtype=dtype)


def get_data(D): return D.data.T


# ---------------------------------------------------------------------------------------------------
# argument plumbing
# ---------------------------------------------------------------------------------------------------
def _batch(z):
    """-> (Z (N, n+m) array/tensor, t, dt, single?)."""
    if isinstance(z, KnotPoint):
        return np.ascontiguousarray(z.z[None, :]), np.array([z.t]), np.array([z.dt]), True
    if isinstance(z, (SampledTrajectory, DeviceTrajectory)):
        return z.data, z.times, z.dts, False
    raise TypeError(f"expected a KnotPoint, SampledTrajectory or DeviceTrajectory, got {type(z).__name__}")


def _write(dst, src, single):
    if dst is None:
        return src[0] if single else src
    if isinstance(dst, DynamicsJacobian):
        if not single:
            raise TypeError("a single DynamicsJacobian cannot hold the Jacobians of a whole trajectory: pass an (N, n+m, n) array "
                            "(the memory of N DynamicsJacobians, src/jacobian.jl:26-37) or evaluate one KnotPoint")
        dst.data[...] = src[0]
    elif single:
        dst[...] = src[0].T if src[0].ndim == 2 else src[0]
    elif dst is not src:
        dst[...] = src
    return dst


# ---------------------------------------------------------------------------------------------------
# the hot path
# ---------------------------------------------------------------------------------------------------
def dynamics(model, *args, out=None):
    """dynamics(model, z) | dynamics(model, x, u[, t]) | dynamics(model, Z::SampledTrajectory) -> xdot."""
    if len(args) == 1:
        Z, t, _, single = _batch(args[0])
    else:
        x, u = np.asarray(args[0]), np.asarray(args[1])
        Z, single = np.ascontiguousarray(np.concatenate([x, u])[None, :]), True
        t = np.array([float(args[2])]) if len(args) > 2 else None
    r = model._h.dynamics(Z, t=t, out=None if single else out)
    return r[0] if single else r


def dynamics_(model, xdot, *args):
    xdot[...] = dynamics(model, *args)
    return None


def evaluate(fun, z):                                              # src/functionbase.jl:214-238
    return discrete_dynamics(fun, z) if isinstance(fun, DiscreteDynamics) else dynamics(fun, z)


def discrete_dynamics(*args, out=None):
    """discrete_dynamics(dmodel, z) | (dmodel, x, u, t, dt) | v0.3 (Q, model, z) -> x+."""
    if isinstance(args[0], (QuadratureRule, type)) and not isinstance(args[0], AbstractModel):
        args = (DiscretizedDynamics(args[1], args[0]),) + tuple(args[2:])
    dmodel = args[0]
    if len(args) == 2:
        Z, t, dt, single = _batch(args[1])
    else:
        x, u, t, h = args[1:5]
        Z, dt, single = np.ascontiguousarray(np.concatenate([np.asarray(x), np.asarray(u)])[None, :]), np.array([float(h)]), True
        t = np.array([float(t)])
    r = dmodel._h.discrete_dynamics(_qcode(dmodel.integrator), Z, dt, t=t, out=None if single else out)
    return r[0] if single else r


def discrete_dynamics_(dmodel, xn, *args):
    xn[...] = discrete_dynamics(dmodel, *args)
    return None


def jacobian_(sig, diff, fun, J, y, z):
    """jacobian!(sig, diff, fun, J, y, z): J <- d fun / d [x;u]; y <- fun(z) (the reference leaves y unspecified for
    StaticReturn; here it is always the true output when given).  FiniteDifference is refused:
    the GPU path is exact forward mode (== ForwardAD == UserDefined chain rule to rounding, test/integration_tests.jl:13-17)."""
    if isinstance(diff, FiniteDifference) or diff is FiniteDifference:
        raise NotImplementedModelError(_abi.ERR_NOT_IMPLEMENTED, "jacobian!(::FiniteDifference) on the B200 path")
    Z, t, dt, single = _batch(z)
    h = fun._h
    if isinstance(J, DynamicsJacobian) and not single:
        _write(J, None, single)                                   # raises: one DynamicsJacobian cannot hold a trajectory's Jacobians
    yb = None
    if y is not None:
        yb = _abi.empty_like_kind(Z, (Z.shape[0], h.n)) if single or not hasattr(y, "shape") else y
    if isinstance(fun, DiscretizedDynamics):
        Jb = h.discrete_jacobian(_qcode(fun.integrator), Z, dt, t=t, J=None if single or isinstance(J, DynamicsJacobian) else J, xn=yb)
    else:
        Jb = h.jacobian(Z, t=t, J=None if single or isinstance(J, DynamicsJacobian) else J, xdot=yb)
    _write(J, Jb, single)
    if y is not None and yb is not y:
        y[...] = yb[0] if single else yb
    return None


def discrete_jacobian_(Q, J, model, z):
    """v0.3 spelling: discrete_jacobian!(RK4, ∇f, model, z)  (README.md:81-82)."""
    return jacobian_(StaticReturn(), B200(), DiscretizedDynamics(model, Q), J, None, z)


def discrete_error_jacobian_(dmodel, Jbar, y, z):
    """Error-state expansion of the discrete dynamics for RotationState models:  Jbar <- G(x+)' [A B] blkdiag(G(x), I)
    (nerr x (nerr+m)), y <- x+.  The product Altro / TrajectoryOptimization build from jacobian! and errstate_jacobian!
    (src/liestate.jl:262-298, src/functionbase.jl:135), fused into the Jacobian kernel."""
    Z, t, dt, single = _batch(z)
    h = dmodel._h
    yb = None if y is None else (np.empty((Z.shape[0], h.n), dtype=Z.dtype) if single else y)
    Jb = h.discrete_error_jacobian(_qcode(dmodel.integrator), Z, dt, t=t, J=None if single else Jbar, xn=yb)
    if single:
        Jbar[...] = Jb[0].T
        if y is not None:
            y[...] = yb[0]
    return None


def dynamics_error(dmodel, z2, z1):
    """dynamics_error(dmodel, z2, z1)  (src/discrete_dynamics.jl:116-138; ImplicitMidpoint: src/integration.jl:640-654): for explicit
    rules discrete_dynamics(z1) - state(z2), for ImplicitMidpoint x1 + h f((x1+x2)/2, u1, t + h/2) - x2.  One knot pair, or two
    trajectories / arrays of equal length (pair k = (z1[k], z2[k]))."""
    Z1, t, dt, single = _batch(z1)
    Z2, _, _, _ = _batch(z2)
    e = dmodel._h.dynamics_error(_qcode(dmodel.integrator), Z1, Z2, dt, t=t)
    return e[0] if single else e


def dynamics_error_jacobian_(sig, diff, dmodel, J2, J1, y2, y1, z2, z1):
    """dynamics_error_jacobian!(sig, diff, dmodel, J2, J1, y2, y1, z2, z1)  (src/discrete_dynamics.jl:160-200): J1 <- d e / d z1,
    J2 <- d e / d z2 (n x (n+m) each, J2 = [-I 0] for explicit rules), y2 <- e."""
    Z1, t, dt, single = _batch(z1)
    Z2, _, _, _ = _batch(z2)
    j2, j1, e = dmodel._h.dynamics_error(_qcode(dmodel.integrator), Z1, Z2, dt, t=t, jacobian=True)
    _write(J2, j2, single); _write(J1, j1, single)
    if y2 is not None:
        y2[...] = e[0] if single else e
    return None


def _states_of(model, x):
    """x: KnotPoint | SampledTrajectory | (n,) | (N, >=n) -> (X (N, ld), single?)."""
    if isinstance(x, KnotPoint):
        return np.ascontiguousarray(x.z[None, :]), True
    if isinstance(x, (SampledTrajectory, DeviceTrajectory)):
        return x.data, False
    if _abi._is_torch(x):
        return (x[None, :].contiguous(), True) if x.dim() == 1 else (x, False)
    x = np.asarray(x)
    return (np.ascontiguousarray(x[None, :]), True) if x.ndim == 1 else (x, False)


def errstate_jacobian_(model, G, x):
    """errstate_jacobian!(model, G, x|z): G (n x nerr) fully written (reference writes only the non-zeros)."""
    X, single = _states_of(model, x)
    Gb = model._h.errstate_jacobian(X, G=None if single else G)
    if single:
        G[...] = Gb[0].T if not _abi._is_torch(Gb) else Gb[0].T
    return None


def grad_errstate_jacobian_(model, dG, x, xbar):
    """∇errstate_jacobian!(model, ∇G, x, x̄)."""
    X, single = _states_of(model, x)
    B, _ = _states_of(model, xbar)
    Hb = model._h.grad_errstate_jacobian(X, B, H=None if single else dG)
    if single:
        dG[...] = Hb[0].T
    return None


def state_diff(model, x, x0):
    X, single = _states_of(model, x)
    X0, _ = _states_of(model, x0)
    d = model._h.state_diff(X, X0)
    return d[0] if single else d


def rollout_(sig, dmodel, Z, x0=None):
    """rollout!(sig, dmodel, Z, x0): overwrite the states of Z with the simulated trajectory (src/trajectories.jl:436-441).  A
    DeviceTrajectory is rolled out in place on the device (rdb_trajectory_rollout)."""
    if isinstance(Z, DeviceTrajectory):
        if x0 is not None:
            Z._t.set_initial_state(np.asarray(x0)[None, :] if not _abi._is_torch(x0) else x0[None, :])
        Z._t.rollout(_qcode(dmodel.integrator))
        return None
    x0 = states(Z)[0] if x0 is None else np.asarray(x0)
    X = dmodel._h.rollout(_qcode(dmodel.integrator), np.ascontiguousarray(x0[None, :], dtype=Z.data.dtype),
                          np.ascontiguousarray(controls(Z)[None, :-1]), np.ascontiguousarray(Z.dts[None, :]),
                          t=np.ascontiguousarray(Z.times[None, :]))
    setstates_(Z, X[0])
    return None


class TrajectoryBatch:
    """`ntraj` independent trajectories of K knot points on the device (rdb_trajectory, knot-major across the batch): the container of
    the forward pass + linearisation of sampling-based / multi-shooting solvers.  Arrays are (K, ntraj, width)."""

    def __init__(self, dmodel, ntraj, K, dtype=np.float64):
        self.dmodel, self.Q = dmodel, _qcode(dmodel.integrator)
        self._t = _abi.Trajectory(dmodel._h, ntraj, K, dtype)
        self.ntraj, self.K = int(ntraj), int(K)

    def set_initial_state(self, x0): self._t.set_initial_state(x0)
    def set_states(self, X): self._t.set_states(X)
    def set_controls(self, U): self._t.set_controls(U)
    def set_timesteps(self, dt, t0=0.0): self._t.set_timesteps(dt, t0)
    def states(self, device=False): return self._t.states(device=device)
    def controls(self, device=False): return self._t.controls(device=device)
    def views(self): return self._t.views()
    def rollout(self): self._t.rollout(self.Q)
    def linearize(self, error_state=False, J=None, xn=None, device=True): return self._t.linearize(self.Q, error_state, J, xn, device)

    def rollout_linearize(self, error_state=False, J=None, chunks=0, device=True):
        """x_{k+1} for every knot AND the (error-state) Jacobians of every knot, as one pipelined unit of device work."""
        return self._t.rollout_linearize(self.Q, error_state, J, chunks, device)


def rollout_and_linearize(dmodel, x0, U, dt, error_state=False):
    """Forward pass + linearisation of many trajectories without leaving the device: x0 (ntraj, n), U (ntraj, K-1, m),
    scalar dt.  Returns X (ntraj, K, n) and J (ntraj, K-1, n+m, n) [or (ntraj, K-1, nerr+m, nerr)] at the non-terminal knots, host
    arrays for host inputs and CUDA tensors for CUDA inputs.  One call of rdb_trajectory_rollout_linearize: the rollout kernel writes
    z = [x;u] rows in place, the Jacobian kernel streams them chunk by chunk on a second stream."""
    ntraj, K = int(x0.shape[0]), int(U.shape[1]) + 1
    on_dev = _abi._is_torch(x0)
    dtype = np.dtype(str(x0.dtype).replace("torch.", ""))
    tb = TrajectoryBatch(dmodel, ntraj, K, dtype)
    tb.set_initial_state(x0)
    tb.set_controls(U.permute(1, 0, 2).contiguous() if on_dev else np.ascontiguousarray(np.transpose(U, (1, 0, 2))))
    tb.set_timesteps(float(dt))
    J = tb.rollout_linearize(error_state=error_state, device=on_dev)
    X = tb.states(device=on_dev)
    if on_dev:
        return X.permute(1, 0, 2).contiguous(), J[:-1].permute(1, 0, 2, 3).contiguous()
    return np.ascontiguousarray(np.transpose(X, (1, 0, 2))), np.ascontiguousarray(np.transpose(J[:-1], (1, 0, 2, 3)))


def rollout_batch(dmodel, x0, U, dt):
    """Many independent trajectories at once: x0 (ntraj, n), U (ntraj, K-1, m), dt scalar or (ntraj, K) -> X (ntraj, K, n)."""
    return dmodel._h.rollout(_qcode(dmodel.integrator), x0, U, dt)
