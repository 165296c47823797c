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
    "rdb_version", "rdb_strerror", "rdb_create", "rdb_destroy", "rdb_host_alloc", "rdb_host_free",
    "rdb_model_create", "rdb_model_create_custom", "rdb_model_create_custom_rigid", "rdb_custom_check", "rdb_custom_rigid_check", "rdb_last_log", "rdb_model_destroy", "rdb_model_dims", "rdb_dynamics", "rdb_discrete_dynamics",
    "rdb_jacobian", "rdb_discrete_jacobian", "rdb_discrete_error_jacobian", "rdb_errstate_jacobian", "rdb_grad_errstate_jacobian",
    "rdb_state_diff", "rdb_rollout",
)


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
        L.rdb_model_create.argtypes = [vp, i32, i32, i32, ctypes.POINTER(dbl), i32, ctypes.POINTER(vp)]
        L.rdb_model_create_custom.argtypes = [vp, i32, i32, ctypes.c_char_p, ctypes.POINTER(dbl), i32, ctypes.POINTER(vp)]
        L.rdb_model_create_custom_rigid.argtypes = [vp, i32, i32, i32, ctypes.c_char_p, dbl, ctypes.POINTER(dbl), ctypes.POINTER(dbl), i32, ctypes.POINTER(vp)]
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


class ModelHandle:
    """rdb_model: (kind, rot, frame, params) bound to a context."""

    def __init__(self, kind, rot, frame, params, device=None, custom=None):
        self.ctx = context(device)
        self.kind, self.rot, self.frame = int(kind), int(rot), int(frame)
        self.params = np.ascontiguousarray(params, dtype=np.float64)
        self._h = ctypes.c_void_p()
        pp = self.params.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
        if custom is not None and len(custom) == 3:    # (n, m, body of f): NVRTC-compiled user model
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
    def _dt(dt, Z):
        if dt is None:
            raise ValueError("dt is required")
        if np.ndim(dt) == 0:
            return None, float(dt)
        return as_f64(dt, Z), 0.0

    # ---- batch operations -------------------------------------------------------------------------
    def dynamics(self, Z, t=None, out=None, layout=AOS):
        N = self._count(Z, layout)
        out = empty_like_kind(Z, self._oshape(N, layout)) if out is None else out
        pz, _ = ptr(Z); po, _ = ptr(out)
        check(lib().rdb_dynamics(self._h, dtype_code(Z), layout, N, pz, None, po, current_stream(Z)), "rdb_dynamics")
        return out

    def discrete_dynamics(self, Q, Z, dt, t=None, out=None, layout=AOS):
        N = self._count(Z, layout)
        out = empty_like_kind(Z, self._oshape(N, layout)) if out is None else out
        dtv, dt0 = self._dt(dt, Z)
        pz, _ = ptr(Z); po, _ = ptr(out); pd, _ = ptr(dtv)
        check(lib().rdb_discrete_dynamics(self._h, int(Q), dtype_code(Z), layout, N, pz, None, pd, dt0, po,
                                          current_stream(Z)), "rdb_discrete_dynamics")
        return out

    def jacobian(self, Z, t=None, J=None, xdot=None, layout=AOS):
        N = self._count(Z, layout)
        J = empty_like_kind(Z, self._jshape(N, layout)) if J is None else J
        pz, _ = ptr(Z); pj, _ = ptr(J); po, _ = ptr(xdot)
        check(lib().rdb_jacobian(self._h, dtype_code(Z), layout, N, pz, None, pj, po, current_stream(Z)), "rdb_jacobian")
        return J

    def discrete_jacobian(self, Q, Z, dt, t=None, J=None, xn=None, layout=AOS):
        N = self._count(Z, layout)
        J = empty_like_kind(Z, self._jshape(N, layout)) if J is None else J
        dtv, dt0 = self._dt(dt, Z)
        pz, _ = ptr(Z); pj, _ = ptr(J); po, _ = ptr(xn); pd, _ = ptr(dtv)
        check(lib().rdb_discrete_jacobian(self._h, int(Q), dtype_code(Z), layout, N, pz, None, pd, dt0, pj, po,
                                          current_stream(Z)), "rdb_discrete_jacobian")
        return J

    def discrete_error_jacobian(self, Q, Z, dt, t=None, J=None, xn=None, layout=AOS):
        """Jbar = G(x+)' [A B] blkdiag(G(x), I): AOS (N, nerr+m, nerr) C-order, i.e. per knot a column-major nerr x (nerr+m)."""
        N = self._count(Z, layout)
        nc = self.nerr + self.m
        J = empty_like_kind(Z, (N, nc, self.nerr) if layout == AOS else (self.nerr * nc, N)) if J is None else J
        dtv, dt0 = self._dt(dt, Z)
        pz, _ = ptr(Z); pj, _ = ptr(J); po, _ = ptr(xn); pd, _ = ptr(dtv)
        check(lib().rdb_discrete_error_jacobian(self._h, int(Q), dtype_code(Z), layout, N, pz, None, pd, dt0, pj, po,
                                                current_stream(Z)), "rdb_discrete_error_jacobian")
        return J

    def errstate_jacobian(self, X, G=None):
        """X (N, ld) with ld >= n (pass Z itself to read the states in place).  G (N, nerr, n) C-order."""
        N, ld = int(X.shape[0]), int(X.shape[1])
        G = empty_like_kind(X, (N, self.nerr, self.n)) if G is None else G
        px, _ = ptr(X); pg, _ = ptr(G)
        check(lib().rdb_errstate_jacobian(self._h, dtype_code(X), N, px, ld, pg, current_stream(X)), "rdb_errstate_jacobian")
        return G

    def grad_errstate_jacobian(self, X, Xbar, H=None):
        N = int(X.shape[0])
        H = empty_like_kind(X, (N, self.nerr, self.nerr)) if H is None else H
        px, _ = ptr(X); pb, _ = ptr(Xbar); ph, _ = ptr(H)
        check(lib().rdb_grad_errstate_jacobian(self._h, dtype_code(X), N, px, int(X.shape[1]), pb, int(Xbar.shape[1]), ph,
                                               current_stream(X)), "rdb_grad_errstate_jacobian")
        return H

    def state_diff(self, X, X0, dX=None):
        N = int(X.shape[0])
        dX = empty_like_kind(X, (N, self.nerr)) if dX is None else dX
        px, _ = ptr(X); p0, _ = ptr(X0); pd, _ = ptr(dX)
        check(lib().rdb_state_diff(self._h, dtype_code(X), N, px, int(X.shape[1]), p0, int(X0.shape[1]), pd,
                                   current_stream(X)), "rdb_state_diff")
        return dX

    def rollout(self, Q, x0, U, dt, X=None):
        """x0 (ntraj, n); U (ntraj, K-1, m); dt scalar or (ntraj, K).  Returns X (ntraj, K, n)."""
        ntraj, K = int(x0.shape[0]), int(U.shape[1]) + 1
        X = empty_like_kind(x0, (ntraj, K, self.n)) if X is None else X
        dtv, dt0 = self._dt(dt, x0)
        p0, _ = ptr(x0); pu, _ = ptr(U); pd, _ = ptr(dtv); pX, _ = ptr(X)
        check(lib().rdb_rollout(self._h, int(Q), dtype_code(x0), ntraj, K, p0, pu, None, pd, dt0, pX,
                                current_stream(x0)), "rdb_rollout")
        return X
