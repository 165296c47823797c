"""ctypes binding of librdb200.so (include/rdb200.h) — the only way this package computes anything.

There is deliberately no CPU fallback: if the CUDA library is missing, import fails loudly; if no GPU is
visible, `Context()` raises.  Arrays may be numpy arrays (host pointers: the library runs its pinned,
multi-stream H2D/compute/D2H pipeline and returns when the outputs are valid) or torch CUDA tensors (device
pointers: work is enqueued on the current torch stream and the call returns immediately).

Memory images (reference: src/jacobian.jl:26-37, src/knotpoint.jl:148-153):
  AOS  Z (N, n+m) C-order  == Julia Matrix{T}(n+m, N);   J (N, n+m, n) C-order == Julia Array{T,3}(n, n+m, N),
       i.e. every knot holds a column-major n x (n+m) [A B];  x+ (N, n).
  SOA  Z (n+m, N), J (n*(n+m), N) with row index i + n*j, x+ (n, N): one unit-stride stream per component.
"""
import ctypes
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RDB200_LIB") or os.path.join(_PKG, "librdb200.so")   # RDB200_LIB: tuning builds (scripts/tune.py)

F32, F64 = 0, 1
AOS, SOA = 0, 1
EULER, RK2, RK3, RK4, IMPLICIT_MIDPOINT = 0, 1, 2, 3, 4
CARTPOLE, QUADROTOR, BODY, DOUBLE_INTEGRATOR, CUSTOM = 0, 1, 2, 3, 4
ROT_NONE, ROT_QUAT, ROT_MRP, ROT_RP = 0, 1, 2, 3
FRAME_WORLD, FRAME_BODY = 0, 1

ERR_ARG, ERR_NOT_IMPLEMENTED, ERR_POINTER_MIX, ERR_NO_DEVICE, ERR_COMPILE = -1, -2, -3, -4, -5

# every symbol include/rdb200.h declares (tests check the library exports exactly these)
SYMBOLS = (
    "rdb_version", "rdb_strerror", "rdb_create", "rdb_destroy", "rdb_host_alloc", "rdb_host_free", "rdb_host_register", "rdb_host_unregister",
    "rdb_model_create", "rdb_model_create_custom", "rdb_model_create_custom_rigid", "rdb_model_create_custom_lie", "rdb_custom_check", "rdb_custom_rigid_check", "rdb_last_log", "rdb_model_destroy", "rdb_model_dims", "rdb_dynamics", "rdb_discrete_dynamics",
    "rdb_jacobian", "rdb_discrete_jacobian", "rdb_discrete_error_jacobian", "rdb_errstate_jacobian", "rdb_grad_errstate_jacobian",
    "rdb_state_diff", "rdb_rollout",
    "rdb_dynamics_error", "rdb_dynamics_error_jacobian",
    "rdb_plan_create", "rdb_plan_launch", "rdb_plan_set_shared", "rdb_plan_destroy",
    "rdb_trajectory_create", "rdb_trajectory_destroy", "rdb_trajectory_dims", "rdb_trajectory_data", "rdb_trajectory_set_states",
    "rdb_trajectory_set_initial_state", "rdb_trajectory_set_controls", "rdb_trajectory_set_timesteps", "rdb_trajectory_get_states",
    "rdb_trajectory_get_controls", "rdb_trajectory_rollout", "rdb_trajectory_linearize", "rdb_trajectory_rollout_linearize",
)
OP_DYNAMICS, OP_DISCRETE_DYNAMICS, OP_JACOBIAN, OP_DISCRETE_JACOBIAN, OP_DISCRETE_ERROR_JACOBIAN = 0, 1, 2, 3, 4


class RDBError(RuntimeError):
    def __init__(self, code, what):
        super().__init__(f"{what}: rdb200 status {code}: {strerror(code)}")
        self.code = code


class NotImplementedModelError(RDBError, NotImplementedError):
    """RobotDynamics.NotImplementedError (reference: src/utils.jl:1-8)."""


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built "
                "(run `python robotdynamics.jl_b200/csrc/build.py`); there is no CPU fallback")
        L = ctypes.CDLL(LIB_PATH)
        vp, i32, i64, dbl = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_double
        L.rdb_version.restype = i32
        L.rdb_strerror.restype = ctypes.c_char_p
        L.rdb_strerror.argtypes = [i32]
        L.rdb_create.argtypes = [i32, ctypes.POINTER(vp)]
        L.rdb_destroy.argtypes = [vp]
        L.rdb_host_alloc.restype = vp
        L.rdb_host_alloc.argtypes = [ctypes.c_size_t]
        L.rdb_host_free.argtypes = [vp]
        L.rdb_host_register.argtypes = [vp, ctypes.c_size_t]
        L.rdb_host_unregister.argtypes = [vp]
        L.rdb_model_create.argtypes = [vp, i32, i32, i32, ctypes.POINTER(dbl), i32, ctypes.POINTER(vp)]
        L.rdb_model_create_custom.argtypes = [vp, i32, i32, ctypes.c_char_p, ctypes.POINTER(dbl), i32, ctypes.POINTER(vp)]
        L.rdb_model_create_custom_rigid.argtypes = [vp, i32, i32, i32, ctypes.c_char_p, dbl, ctypes.POINTER(dbl), ctypes.POINTER(dbl), i32, ctypes.POINTER(vp)]
        L.rdb_model_create_custom_lie.argtypes = [vp, i32, i32, ctypes.POINTER(i32), i32, ctypes.c_char_p, ctypes.POINTER(dbl), i32, ctypes.POINTER(vp)]
        L.rdb_custom_check.argtypes = [i32, i32, ctypes.c_char_p, i32, i32]
        L.rdb_custom_rigid_check.argtypes = [i32, i32, i32, ctypes.c_char_p, i32, i32]
        L.rdb_last_log.restype = ctypes.c_char_p
        L.rdb_model_destroy.argtypes = [vp]
        L.rdb_model_dims.argtypes = [vp, ctypes.POINTER(i32), ctypes.POINTER(i32), ctypes.POINTER(i32)]
        L.rdb_dynamics.argtypes = [vp, i32, i32, i64, vp, vp, vp, vp]
        L.rdb_discrete_dynamics.argtypes = [vp, i32, i32, i32, i64, vp, vp, vp, dbl, vp, vp]
        L.rdb_jacobian.argtypes = [vp, i32, i32, i64, vp, vp, vp, vp, vp]
        L.rdb_discrete_jacobian.argtypes = [vp, i32, i32, i32, i64, vp, vp, vp, dbl, vp, vp, vp]
        L.rdb_discrete_error_jacobian.argtypes = [vp, i32, i32, i32, i64, vp, vp, vp, dbl, vp, vp, vp]
        L.rdb_errstate_jacobian.argtypes = [vp, i32, i64, vp, i32, vp, vp]
        L.rdb_grad_errstate_jacobian.argtypes = [vp, i32, i64, vp, i32, vp, i32, vp, vp]
        L.rdb_state_diff.argtypes = [vp, i32, i64, vp, i32, vp, i32, vp, vp]
        L.rdb_rollout.argtypes = [vp, i32, i32, i64, i32, vp, vp, vp, vp, dbl, vp, vp]
        L.rdb_dynamics_error.argtypes = [vp, i32, i32, i64, vp, vp, i32, vp, vp, dbl, vp, vp]
        L.rdb_dynamics_error_jacobian.argtypes = [vp, i32, i32, i64, vp, vp, i32, vp, vp, dbl, vp, vp, vp, vp]
        L.rdb_plan_create.argtypes = [vp, i32, i32, i32, i32, i64, vp, vp, vp, dbl, vp, vp, ctypes.POINTER(vp)]
        L.rdb_plan_launch.argtypes = [vp, vp]
        L.rdb_plan_destroy.argtypes = [vp]
        L.rdb_plan_set_shared.argtypes = [vp, i32]
        L.rdb_trajectory_create.argtypes = [vp, i32, i64, i32, ctypes.POINTER(vp)]
        L.rdb_trajectory_destroy.argtypes = [vp]
        L.rdb_trajectory_dims.argtypes = [vp, ctypes.POINTER(i64), ctypes.POINTER(i32), ctypes.POINTER(i32), ctypes.POINTER(i32), ctypes.POINTER(i32)]
        L.rdb_trajectory_data.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(vp), ctypes.POINTER(vp)]
        L.rdb_trajectory_set_states.argtypes = [vp, vp, vp]
        L.rdb_trajectory_set_initial_state.argtypes = [vp, vp, vp]
        L.rdb_trajectory_set_controls.argtypes = [vp, vp, i32, vp]
        L.rdb_trajectory_set_timesteps.argtypes = [vp, vp, dbl, dbl, vp]
        L.rdb_trajectory_get_states.argtypes = [vp, vp, vp]
        L.rdb_trajectory_get_controls.argtypes = [vp, vp, vp]
        L.rdb_trajectory_rollout.argtypes = [vp, i32, vp]
        L.rdb_trajectory_linearize.argtypes = [vp, i32, i32, vp, vp, vp]
        L.rdb_trajectory_rollout_linearize.argtypes = [vp, i32, i32, i32, vp, vp]
        _lib = L
    return _lib


def strerror(code):
    return lib().rdb_strerror(int(code)).decode()


def check(rc, what):
    if rc == ERR_COMPILE:
        raise RDBError(rc, what + "\n" + lib().rdb_last_log().decode(errors="replace")[-3000:])
    if rc != 0:
        raise (NotImplementedModelError if rc == ERR_NOT_IMPLEMENTED else RDBError)(rc, what)


def custom_rigid_check(rot, frame, m, body, nparams=0, dtype=F64):
    """Compile-only check of a user rigid-body wrench (no GPU needed).  Returns (ok, compiler log)."""
    rc = lib().rdb_custom_rigid_check(int(rot), int(frame), int(m), body.encode(), int(nparams), int(dtype))
    return rc == 0, lib().rdb_last_log().decode(errors="replace")


def custom_check(n, m, body, nparams=0, dtype=F64):
    """Compile-only check of a user model body (no GPU needed).  Returns (ok, compiler log)."""
    rc = lib().rdb_custom_check(int(n), int(m), body.encode(), int(nparams), int(dtype))
    return rc == 0, lib().rdb_last_log().decode(errors="replace")


# ---------------------------------------------------------------------------------------------------
# array plumbing: numpy (host) or torch CUDA tensors (device)
# ---------------------------------------------------------------------------------------------------
def _is_torch(a):
    return type(a).__module__.startswith("torch")


_DTYPE_CODES = {}          # dtype object -> F32 / F64 (filled lazily: a per-call string conversion costs more than the launch)


def dtype_code(a):
    dt = a.dtype
    code = _DTYPE_CODES.get(dt)
    if code is None:
        name = str(dt).replace("torch.", "")
        if name not in ("float32", "float64"):
            raise TypeError(f"rdb200 computes in float32 or float64, got {a.dtype}")
        code = _DTYPE_CODES[dt] = F32 if name == "float32" else F64
    return code


def ptr(a):
    """(address, keepalive) of a contiguous numpy array / torch tensor, or (None, None)."""
    if a is None:
        return None, None
    if _is_torch(a):
        if not a.is_contiguous():
            raise ValueError("rdb200 needs contiguous tensors")
        return a.data_ptr(), a
    if not isinstance(a, np.ndarray) or not a.flags.c_contiguous:
        raise ValueError("rdb200 needs C-contiguous numpy arrays")
    return a.ctypes.data, a


def current_stream(a):
    if a is not None and _is_torch(a) and a.is_cuda:
        import torch
        return torch.cuda.current_stream(a.device).cuda_stream
    return None


def empty_like_kind(ref, shape):
    """uninitialised array of `shape` living where `ref` lives, same dtype."""
    if _is_torch(ref):
        import torch
        return torch.empty(shape, dtype=ref.dtype, device=ref.device)
    return np.empty(shape, dtype=ref.dtype)


def check_buffer(name, a, shape, like):
    """Caller-supplied output / auxiliary arrays are handed to the C ABI as bare pointers: the library reads or writes exactly
    prod(shape) * sizeof(dtype) bytes there, so anything that does not match is refused HERE (ValueError / TypeError) instead of
    becoming an out-of-bounds access in host or device memory."""
    if a is None:
        return None
    if _is_torch(a) != _is_torch(like):
        raise RDBError(ERR_POINTER_MIX, f"{name}: host (numpy) and device (torch) arrays mixed in one call")
    if _is_torch(a):
        if a.device != like.device:
            raise ValueError(f"{name} lives on {a.device}, the inputs on {like.device}")
        if not a.is_contiguous():
            raise ValueError(f"{name} must be contiguous")
    elif not isinstance(a, np.ndarray) or not a.flags.c_contiguous:
        raise ValueError(f"{name} must be a C-contiguous numpy array")
    if a.dtype != like.dtype:
        raise TypeError(f"{name} has dtype {a.dtype}, the inputs {like.dtype}: the C ABI computes and stores in ONE type per call")
    if tuple(a.shape) != tuple(shape):
        raise ValueError(f"{name} must have shape {tuple(shape)}, got {tuple(a.shape)}")
    return a


def as_f64(a, like):
    """t / dt vectors are always Float64 (src/knotpoint.jl:148-153); put them where `like` lives."""
    if a is None:
        return None
    if _is_torch(like):
        import torch
        if _is_torch(a) and a.dtype == torch.float64 and a.device == like.device and a.is_contiguous():
            return a
        return torch.as_tensor(a, dtype=torch.float64, device=like.device).contiguous()
    return np.ascontiguousarray(a, dtype=np.float64)


class Context:
    """rdb_context: one per (process, GPU)."""

    def __init__(self, device=0):
        self._h = ctypes.c_void_p()
        check(lib().rdb_create(int(device), ctypes.byref(self._h)), f"rdb_create(device={device})")
        self.device = int(device)

    def close(self):
        if self._h:
            lib().rdb_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_contexts = {}


def context(device=None):
    """process-wide context for a device (default: torch's current CUDA device, else 0)."""
    if device is None:
        device = 0
        try:
            import torch
            if torch.cuda.is_available():
                device = torch.cuda.current_device()
        except ImportError:
            pass
    if device not in _contexts:
        _contexts[device] = Context(device)
    return _contexts[device]


class PinnedArray:
    """numpy view over cudaMallocHost memory (rdb_host_alloc); keeps the allocation alive."""

    def __init__(self, shape, dtype):
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self._p = lib().rdb_host_alloc(max(self.nbytes, 1))
        if not self._p:
            raise MemoryError("rdb_host_alloc failed")
        buf = (ctypes.c_char * max(self.nbytes, 1)).from_address(self._p)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def __del__(self):
        try:
            if self._p:
                lib().rdb_host_free(self._p)
                self._p = None
        except Exception:
            pass


class RegisteredArray:
    """Page-locks a numpy array the caller already owns (rdb_host_register) for as long as this object lives: the host-pointer
    path then moves it by DMA.  Use as a context manager or keep the object next to the array."""

    def __init__(self, array):
        if not isinstance(array, np.ndarray) or not array.flags.c_contiguous or array.nbytes == 0:
            raise ValueError("RegisteredArray needs a non-empty C-contiguous numpy array")
        self.array, self._p = array, array.ctypes.data
        rc = lib().rdb_host_register(self._p, array.nbytes)
        if rc != 0:
            self._p = None
            raise RDBError(f"rdb_host_register failed ({rc})")

    def close(self):
        if self._p:
            lib().rdb_host_unregister(self._p)
            self._p = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ModelHandle:
    """rdb_model: (kind, rot, frame, params) bound to a context."""

    def __init__(self, kind, rot, frame, params, device=None, custom=None):
        self.ctx = context(device)
        self.kind, self.rot, self.frame = int(kind), int(rot), int(frame)
        self.params = np.ascontiguousarray(params, dtype=np.float64)
        self.time_varying = custom is not None          # user models may define dynamics(model, x, u, t)
        self._h = ctypes.c_void_p()
        pp = self.params.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        if custom is not None and len(custom) == 4 and isinstance(custom[3], str) and custom[3] == "lie":      # (parts, m, body, "lie")
            parts, m, body, _ = custom
            arr = (ctypes.c_int * len(parts))(*[int(v) for v in parts])
            check(lib().rdb_model_create_custom_lie(self.ctx._h, self.rot, len(parts), arr, int(m), body.encode(), pp, len(self.params),
                                                    ctypes.byref(self._h)), "rdb_model_create_custom_lie")
        elif custom is not None and len(custom) == 3:    # (n, m, body of f): NVRTC-compiled user model
            n, m, body = custom
            check(lib().rdb_model_create_custom(self.ctx._h, int(n), int(m), body.encode(), pp, len(self.params),
                                                ctypes.byref(self._h)), "rdb_model_create_custom")
        elif custom is not None:                       # (m, wrench body, mass, J): RigidBody{R} with a user wrench
            m, body, mass, J = custom
            J = np.ascontiguousarray(J, dtype=np.float64).reshape(9)
            check(lib().rdb_model_create_custom_rigid(self.ctx._h, self.rot, self.frame, int(m), body.encode(), float(mass),
                                                      J.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), pp, len(self.params),
                                                      ctypes.byref(self._h)), "rdb_model_create_custom_rigid")
        else:
            check(lib().rdb_model_create(self.ctx._h, self.kind, self.rot, self.frame, pp, len(self.params),
                                         ctypes.byref(self._h)), "rdb_model_create")
        n, m, ne = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        check(lib().rdb_model_dims(self._h, ctypes.byref(n), ctypes.byref(m), ctypes.byref(ne)), "rdb_model_dims")
        self.n, self.m, self.nerr = n.value, m.value, ne.value

    def __del__(self):
        try:
            if self._h:
                lib().rdb_model_destroy(self._h)
                self._h = ctypes.c_void_p()
        except Exception:
            pass

    # ---- shape helpers -------------------------------------------------------------------------
    def _count(self, Z, layout):
        nz = self.n + self.m
        if layout == AOS:
            if Z.ndim != 2 or Z.shape[1] != nz:
                raise ValueError(f"AOS Z must be (N, {nz}), got {tuple(Z.shape)}")
            return int(Z.shape[0])
        if Z.ndim != 2 or Z.shape[0] != nz:
            raise ValueError(f"SOA Z must be ({nz}, N), got {tuple(Z.shape)}")
        return int(Z.shape[1])

    def _jshape(self, N, layout):
        nz = self.n + self.m
        return (N, nz, self.n) if layout == AOS else (self.n * nz, N)

    def _oshape(self, N, layout):
        return (N, self.n) if layout == AOS else (self.n, N)

    @staticmethod
    def _dt(dt, Z, count=None):
        """scalar -> (None, dt0); vector -> (Float64 array where Z lives, 0.0), which must hold one step per knot point."""
        if dt is None:
            raise ValueError("dt is required")
        if np.ndim(dt) == 0:
            return None, float(dt)
        v = as_f64(dt, Z)
        if count is not None and int(np.prod(tuple(v.shape))) != int(count):
            raise ValueError(f"dt must be a scalar or hold {count} steps (one per knot point), got shape {tuple(v.shape)}")
        return v, 0.0

    def _t(self, t, Z, count):
        """KnotPoint.t per knot point (Float64).  Only time-varying (user) models read it; for the shipped time-invariant families the
        vector is not even transferred (reference: dynamics(model, x, u, t) ignores t unless the model defines it, src/dynamics.jl:81-83)."""
        if t is None or not self.time_varying:
            return None
        v = as_f64(np.full(count, float(t)) if np.ndim(t) == 0 else t, Z)
        if int(np.prod(tuple(v.shape))) != int(count):
            raise ValueError(f"t must hold {count} times (one per knot point), got shape {tuple(v.shape)}")
        return v

    # ---- batch operations -------------------------------------------------------------------------
    def _zcheck(self, Z):
        if not (_is_torch(Z) or isinstance(Z, np.ndarray)):
            raise TypeError("Z must be a numpy array or a torch tensor")
        dtype_code(Z)

    def dynamics(self, Z, t=None, out=None, layout=AOS):
        self._zcheck(Z)
        N = self._count(Z, layout)
        out = empty_like_kind(Z, self._oshape(N, layout)) if out is None else check_buffer("out", out, self._oshape(N, layout), Z)
        tv = self._t(t, Z, N)
        pz, _ = ptr(Z); po, _ = ptr(out); pt, _ = ptr(tv)
        check(lib().rdb_dynamics(self._h, dtype_code(Z), layout, N, pz, pt, po, current_stream(Z)), "rdb_dynamics")
        return out

    def discrete_dynamics(self, Q, Z, dt, t=None, out=None, layout=AOS):
        self._zcheck(Z)
        N = self._count(Z, layout)
        out = empty_like_kind(Z, self._oshape(N, layout)) if out is None else check_buffer("out", out, self._oshape(N, layout), Z)
        dtv, dt0 = self._dt(dt, Z, N)
        tv = self._t(t, Z, N)
        pz, _ = ptr(Z); po, _ = ptr(out); pd, _ = ptr(dtv); pt, _ = ptr(tv)
        check(lib().rdb_discrete_dynamics(self._h, int(Q), dtype_code(Z), layout, N, pz, pt, pd, dt0, po,
                                          current_stream(Z)), "rdb_discrete_dynamics")
        return out

    def jacobian(self, Z, t=None, J=None, xdot=None, layout=AOS):
        self._zcheck(Z)
        N = self._count(Z, layout)
        J = empty_like_kind(Z, self._jshape(N, layout)) if J is None else check_buffer("J", J, self._jshape(N, layout), Z)
        check_buffer("xdot", xdot, self._oshape(N, layout), Z)
        tv = self._t(t, Z, N)
        pz, _ = ptr(Z); pj, _ = ptr(J); po, _ = ptr(xdot); pt, _ = ptr(tv)
        check(lib().rdb_jacobian(self._h, dtype_code(Z), layout, N, pz, pt, pj, po, current_stream(Z)), "rdb_jacobian")
        return J

    def discrete_jacobian(self, Q, Z, dt, t=None, J=None, xn=None, layout=AOS):
        self._zcheck(Z)
        N = self._count(Z, layout)
        J = empty_like_kind(Z, self._jshape(N, layout)) if J is None else check_buffer("J", J, self._jshape(N, layout), Z)
        check_buffer("xn", xn, self._oshape(N, layout), Z)
        dtv, dt0 = self._dt(dt, Z, N)
        tv = self._t(t, Z, N)
        pz, _ = ptr(Z); pj, _ = ptr(J); po, _ = ptr(xn); pd, _ = ptr(dtv); pt, _ = ptr(tv)
        check(lib().rdb_discrete_jacobian(self._h, int(Q), dtype_code(Z), layout, N, pz, pt, pd, dt0, pj, po,
                                          current_stream(Z)), "rdb_discrete_jacobian")
        return J

    def discrete_error_jacobian(self, Q, Z, dt, t=None, J=None, xn=None, layout=AOS):
        """Jbar = G(x+)' [A B] blkdiag(G(x), I): AOS (N, nerr+m, nerr) C-order, i.e. per knot a column-major nerr x (nerr+m)."""
        self._zcheck(Z)
        N = self._count(Z, layout)
        nc = self.nerr + self.m
        jshape = (N, nc, self.nerr) if layout == AOS else (self.nerr * nc, N)
        J = empty_like_kind(Z, jshape) if J is None else check_buffer("Jbar", J, jshape, Z)
        check_buffer("xn", xn, self._oshape(N, layout), Z)
        dtv, dt0 = self._dt(dt, Z, N)
        tv = self._t(t, Z, N)
        pz, _ = ptr(Z); pj, _ = ptr(J); po, _ = ptr(xn); pd, _ = ptr(dtv); pt, _ = ptr(tv)
        check(lib().rdb_discrete_error_jacobian(self._h, int(Q), dtype_code(Z), layout, N, pz, pt, pd, dt0, pj, po,
                                                current_stream(Z)), "rdb_discrete_error_jacobian")
        return J

    def _xcheck(self, name, X):
        if X.ndim != 2 or X.shape[1] < self.n:
            raise ValueError(f"{name} must be (N, ld) with ld >= {self.n}, got {tuple(X.shape)}")
        dtype_code(X)
        return int(X.shape[0]), int(X.shape[1])

    def errstate_jacobian(self, X, G=None):
        """X (N, ld) with ld >= n (pass Z itself to read the states in place).  G (N, nerr, n) C-order."""
        N, ld = self._xcheck("X", X)
        G = empty_like_kind(X, (N, self.nerr, self.n)) if G is None else check_buffer("G", G, (N, self.nerr, self.n), X)
        px, _ = ptr(X); pg, _ = ptr(G)
        check(lib().rdb_errstate_jacobian(self._h, dtype_code(X), N, px, ld, pg, current_stream(X)), "rdb_errstate_jacobian")
        return G

    def grad_errstate_jacobian(self, X, Xbar, H=None):
        N, ld = self._xcheck("X", X)
        Nb, ldb = self._xcheck("Xbar", Xbar)
        check_buffer("Xbar", Xbar, (N, ldb), X)
        H = empty_like_kind(X, (N, self.nerr, self.nerr)) if H is None else check_buffer("H", H, (N, self.nerr, self.nerr), X)
        px, _ = ptr(X); pb, _ = ptr(Xbar); ph, _ = ptr(H)
        check(lib().rdb_grad_errstate_jacobian(self._h, dtype_code(X), N, px, ld, pb, ldb, ph,
                                               current_stream(X)), "rdb_grad_errstate_jacobian")
        return H

    def state_diff(self, X, X0, dX=None):
        N, ld = self._xcheck("X", X)
        _, ld0 = self._xcheck("X0", X0)
        check_buffer("X0", X0, (N, ld0), X)
        dX = empty_like_kind(X, (N, self.nerr)) if dX is None else check_buffer("dX", dX, (N, self.nerr), X)
        px, _ = ptr(X); p0, _ = ptr(X0); pd, _ = ptr(dX)
        check(lib().rdb_state_diff(self._h, dtype_code(X), N, px, ld, p0, ld0, pd,
                                   current_stream(X)), "rdb_state_diff")
        return dX

    def dynamics_error(self, Q, Z1, Z2, dt, t=None, e=None, J1=None, J2=None, jacobian=False):
        """dynamics_error (and, jacobian=True, dynamics_error_jacobian!) of the knot pairs (Z1[k], Z2[k]): Z1 (N, n+m), Z2 (N, ld2 >= n).
        Returns e, or (J2, J1, e)."""
        self._zcheck(Z1)
        N = self._count(Z1, AOS)
        _, ld2 = self._xcheck("Z2", Z2)
        check_buffer("Z2", Z2, (N, ld2), Z1)
        e = empty_like_kind(Z1, (N, self.n)) if e is None else check_buffer("e", e, (N, self.n), Z1)
        dtv, dt0 = self._dt(dt, Z1, N)
        tv = self._t(t, Z1, N)
        p1, _ = ptr(Z1); p2, _ = ptr(Z2); pe, _ = ptr(e); pd, _ = ptr(dtv); pt, _ = ptr(tv)
        if not jacobian:
            check(lib().rdb_dynamics_error(self._h, int(Q), dtype_code(Z1), N, p1, p2, ld2, pt, pd, dt0, pe, current_stream(Z1)), "rdb_dynamics_error")
            return e
        J1 = empty_like_kind(Z1, self._jshape(N, AOS)) if J1 is None else check_buffer("J1", J1, self._jshape(N, AOS), Z1)
        J2 = empty_like_kind(Z1, self._jshape(N, AOS)) if J2 is None else check_buffer("J2", J2, self._jshape(N, AOS), Z1)
        check(lib().rdb_dynamics_error_jacobian(self._h, int(Q), dtype_code(Z1), N, p1, p2, ld2, pt, pd, dt0, ptr(J2)[0], ptr(J1)[0], pe,
                                                current_stream(Z1)), "rdb_dynamics_error_jacobian")
        return J2, J1, e

    def rollout(self, Q, x0, U, dt, X=None, t=None):
        """x0 (ntraj, n); U (ntraj, K-1, m); dt scalar or (ntraj, K); t None or (ntraj, K).  Returns X (ntraj, K, n)."""
        dtype_code(x0)
        if x0.ndim != 2 or x0.shape[1] != self.n:
            raise ValueError(f"x0 must be (ntraj, {self.n}), got {tuple(x0.shape)}")
        ntraj = int(x0.shape[0])
        if U.ndim != 3 or U.shape[0] != ntraj or U.shape[2] != self.m:
            raise ValueError(f"U must be ({ntraj}, K-1, {self.m}), got {tuple(U.shape)}")
        K = int(U.shape[1]) + 1
        check_buffer("U", U, (ntraj, K - 1, self.m), x0)
        X = empty_like_kind(x0, (ntraj, K, self.n)) if X is None else check_buffer("X", X, (ntraj, K, self.n), x0)
        dtv, dt0 = self._dt(dt, x0, ntraj * K)
        tv = self._t(t, x0, ntraj * K)
        p0, _ = ptr(x0); pu, _ = ptr(U); pd, _ = ptr(dtv); pX, _ = ptr(X); pt, _ = ptr(tv)
        check(lib().rdb_rollout(self._h, int(Q), dtype_code(x0), ntraj, K, p0, pu, pt, pd, dt0, pX,
                                current_stream(x0)), "rdb_rollout")
        return X


class Plan:
    """rdb_plan: one validated knot operation on DEVICE tensors; launch() is a single kernel launch on the current stream (or under
    CUDA-graph capture).  The tensors are kept alive by the plan; it evaluates whatever they hold at execution time."""

    def __init__(self, handle, op, Q, Z, dt, t=None, J=None, out=None, layout=AOS, shared_gpu=False):
        if not (_is_torch(Z) and Z.is_cuda):
            raise RDBError(ERR_POINTER_MIX, "a plan needs device tensors")
        handle._zcheck(Z)
        N = handle._count(Z, layout)
        with_j = op in (OP_JACOBIAN, OP_DISCRETE_JACOBIAN, OP_DISCRETE_ERROR_JACOBIAN)
        if op == OP_DISCRETE_ERROR_JACOBIAN:
            nc = handle.nerr + handle.m
            jshape = (N, nc, handle.nerr) if layout == AOS else (handle.nerr * nc, N)
        else:
            jshape = handle._jshape(N, layout)
        if with_j:
            J = empty_like_kind(Z, jshape) if J is None else check_buffer("J", J, jshape, Z)
        else:
            out = empty_like_kind(Z, handle._oshape(N, layout)) if out is None else out
        check_buffer("out", out, handle._oshape(N, layout), Z)
        dtv, dt0 = (None, 0.0) if op in (OP_DYNAMICS, OP_JACOBIAN) else handle._dt(dt, Z, N)
        tv = handle._t(t, Z, N)
        self.handle, self.Z, self.J, self.out, self._keep = handle, Z, J, out, (dtv, tv)
        self._p = ctypes.c_void_p()
        pz, _ = ptr(Z); pj, _ = ptr(J); po, _ = ptr(out); pd, _ = ptr(dtv); pt, _ = ptr(tv)
        check(lib().rdb_plan_create(handle._h, int(op), int(Q), dtype_code(Z), layout, N, pz, pt, pd, dt0, pj, po, ctypes.byref(self._p)),
              "rdb_plan_create")
        if shared_gpu:          # launches will overlap other kernels on the same GPU (rdb_plan_set_shared): keep SM-filling CTAs
            check(lib().rdb_plan_set_shared(self._p, 1), "rdb_plan_set_shared")
        self._launch = lib().rdb_plan_launch
        import torch
        self._stream = torch.cuda.current_stream

    def launch(self, stream=None):
        rc = self._launch(self._p, self._stream().cuda_stream if stream is None else stream)
        if rc:
            check(rc, "rdb_plan_launch")
        return self.J if self.J is not None else self.out

    def __del__(self):
        try:
            if self._p:
                lib().rdb_plan_destroy(self._p)
                self._p = ctypes.c_void_p()
        except Exception:
            pass


class Trajectory:
    """rdb_trajectory: persistent device mirror of `ntraj` trajectories x K knot points, knot-major across the batch (row k * ntraj + j).
    Arrays passed to the setters / getters are dense (K, ntraj, width) host (numpy) or device (torch) arrays."""

    def __init__(self, handle, ntraj, K, dtype=np.float64):
        self.handle, self.ntraj, self.K = handle, int(ntraj), int(K)
        self.dtype = np.dtype(dtype)
        self.code = F32 if self.dtype == np.float32 else F64
        if self.dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
            raise TypeError(f"rdb200 computes in float32 or float64, got {dtype}")
        self._p = ctypes.c_void_p()
        check(lib().rdb_trajectory_create(handle._h, self.code, self.ntraj, self.K, ctypes.byref(self._p)), "rdb_trajectory_create")

    def __del__(self):
        try:
            if self._p:
                lib().rdb_trajectory_destroy(self._p)
                self._p = ctypes.c_void_p()
        except Exception:
            pass

    def _arr(self, name, a, shape, f64=False):
        want = np.dtype(np.float64) if f64 else self.dtype
        if _is_torch(a):
            if str(a.dtype).replace("torch.", "") != want.name:
                raise TypeError(f"{name} has dtype {a.dtype}, the trajectory stores {want.name}")
            if not a.is_contiguous():
                raise ValueError(f"{name} must be contiguous")
        else:
            a = np.ascontiguousarray(a, dtype=want)
        if tuple(a.shape) != tuple(shape):
            raise ValueError(f"{name} must have shape {tuple(shape)}, got {tuple(a.shape)}")
        return a

    @staticmethod
    def _stream(a):
        s = current_stream(a)
        if s is None:
            try:
                import torch
                s = torch.cuda.current_stream().cuda_stream
            except ImportError:
                s = None
        return s

    def views(self):
        """Zero-copy torch views of the mirror itself: Z (K, ntraj, n+m), t and dt (K, ntraj) — the memory the kernels read."""
        import torch
        pz, pt, pd = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
        check(lib().rdb_trajectory_data(self._p, ctypes.byref(pz), ctypes.byref(pt), ctypes.byref(pd)), "rdb_trajectory_data")

        class _Mem:                      # __cuda_array_interface__ v2: lets torch wrap a raw device pointer without copying
            def __init__(self, p, shape, typestr, owner):
                self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (int(p), False), "version": 2}
                self._owner = owner
        ts = "<f4" if self.code == F32 else "<f8"
        nz = self.handle.n + self.handle.m
        dev = torch.device("cuda", self.handle.ctx.device)
        Z = torch.as_tensor(_Mem(pz.value, (self.K, self.ntraj, nz), ts, self), device=dev)
        t = torch.as_tensor(_Mem(pt.value, (self.K, self.ntraj), "<f8", self), device=dev)
        dt = torch.as_tensor(_Mem(pd.value, (self.K, self.ntraj), "<f8", self), device=dev)
        return Z, t, dt

    def set_states(self, X):
        X = self._arr("X", X, (self.K, self.ntraj, self.handle.n))
        check(lib().rdb_trajectory_set_states(self._p, ptr(X)[0], self._stream(X)), "rdb_trajectory_set_states")

    def set_initial_state(self, x0):
        x0 = self._arr("x0", x0, (self.ntraj, self.handle.n))
        check(lib().rdb_trajectory_set_initial_state(self._p, ptr(x0)[0], self._stream(x0)), "rdb_trajectory_set_initial_state")

    def set_controls(self, U):
        knots = int(U.shape[0])
        U = self._arr("U", U, (knots, self.ntraj, self.handle.m))
        check(lib().rdb_trajectory_set_controls(self._p, ptr(U)[0], knots, self._stream(U)), "rdb_trajectory_set_controls")

    def set_timesteps(self, dt, t0=0.0):
        if np.ndim(dt) == 0:
            check(lib().rdb_trajectory_set_timesteps(self._p, None, float(dt), float(t0), self._stream(None)), "rdb_trajectory_set_timesteps")
        else:
            dt = self._arr("dt", dt, (self.K, self.ntraj), f64=True)
            check(lib().rdb_trajectory_set_timesteps(self._p, ptr(dt)[0], 0.0, float(t0), self._stream(dt)), "rdb_trajectory_set_timesteps")

    def _get(self, fn, width, out, device):
        shape = (self.K, self.ntraj, width)
        if out is None:
            if device:
                import torch
                out = torch.empty(shape, dtype=getattr(torch, self.dtype.name), device="cuda")
            else:
                out = np.empty(shape, dtype=self.dtype)
        else:
            out = self._arr("out", out, shape)
        check(fn(self._p, ptr(out)[0], self._stream(out)), fn.__name__)
        return out

    def states(self, out=None, device=False):
        return self._get(lib().rdb_trajectory_get_states, self.handle.n, out, device)

    def controls(self, out=None, device=False):
        return self._get(lib().rdb_trajectory_get_controls, self.handle.m, out, device)

    def rollout(self, Q):
        check(lib().rdb_trajectory_rollout(self._p, int(Q), self._stream(None)), "rdb_trajectory_rollout")

    def _jout(self, J, error_state, device):
        h = self.handle
        rows, cols = (h.nerr, h.nerr + h.m) if (error_state and h.rot != ROT_NONE) else (h.n, h.n + h.m)
        shape = (self.K, self.ntraj, cols, rows)
        if J is None:
            if device:
                import torch
                return torch.empty(shape, dtype=getattr(torch, self.dtype.name), device="cuda")
            return np.empty(shape, dtype=self.dtype)
        return self._arr("J", J, shape)

    def linearize(self, Q, error_state=False, J=None, xn=None, device=True):
        J = self._jout(J, error_state, device)
        if xn is not None:
            xn = self._arr("xn", xn, (self.K, self.ntraj, self.handle.n))
            if _is_torch(xn) != _is_torch(J):
                raise RDBError(ERR_POINTER_MIX, "J and xn must both be host or both be device arrays")
        check(lib().rdb_trajectory_linearize(self._p, int(Q), int(bool(error_state)), ptr(J)[0], ptr(xn)[0], self._stream(J)),
              "rdb_trajectory_linearize")
        return J

    def rollout_linearize(self, Q, error_state=False, J=None, chunks=0, device=True):
        J = self._jout(J, error_state, device)
        check(lib().rdb_trajectory_rollout_linearize(self._p, int(Q), int(bool(error_state)), int(chunks), ptr(J)[0], self._stream(J)),
              "rdb_trajectory_rollout_linearize")
        return J
