"""CPU-side tests: the C-ABI library loads and exports what include/rdb200.h declares, fails loudly without a GPU,
and the host-side mirror of the reference containers / multi-GPU partitioning behaves.  No compute calls."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    import rdb200
    return rdb200._abi.lib()


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "rdb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(rdb_[a-z_]+)\s*\(", hdr))
    import rdb200
    assert declared == set(rdb200._abi.SYMBOLS), declared ^ set(rdb200._abi.SYMBOLS)
    L = _lib()
    for name in sorted(declared):
        assert hasattr(L, name), f"librdb200.so does not export {name}"
    assert L.rdb_version() >= 100


def test_no_torch_types_in_the_abi():
    hdr = open(os.path.join(ROOT, "include", "rdb200.h")).read()
    assert "torch" not in hdr and "at::" not in hdr and "#include <cuda" not in hdr


def test_strerror_and_argument_errors():
    import rdb200
    L = _lib()
    assert b"not implemented" in L.rdb_strerror(-2)
    assert b"argument" in L.rdb_strerror(-1)
    assert L.rdb_create(0, None) == rdb200._abi.ERR_ARG
    assert L.rdb_model_dims(None, None, None, None) == rdb200._abi.ERR_ARG
    assert L.rdb_discrete_jacobian(None, 3, 1, 0, 8, None, None, None, 0.01, None, None, None) == rdb200._abi.ERR_ARG
    assert L.rdb_host_register(None, 16) == rdb200._abi.ERR_ARG and L.rdb_host_unregister(None) == rdb200._abi.ERR_ARG
    buf = np.zeros(8)
    assert L.rdb_host_register(buf.ctypes.data, 0) == rdb200._abi.ERR_ARG


def test_fails_loudly_without_a_gpu():
    """The product path never falls back to the CPU: with no device, creating a context raises."""
    import torch
    import rdb200
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible here")
    with pytest.raises(rdb200.RDBError) as e:
        rdb200.Cartpole()
    assert e.value.code == rdb200._abi.ERR_NO_DEVICE


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "robotdynamics.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert "oracle" not in open(os.path.join(dirpath, f)).read().lower().replace("# oracle", ""), f


def test_knotpoint_and_trajectory_containers():
    """test/knotpoints.jl, test/trajectories.jl: state/control slices, terminal knot, time grid."""
    import rdb200 as rd
    z = rd.KnotPoint(np.arange(4.0), np.array([9.0]), 0.5, 0.1)
    assert z.n == 4 and z.m == 1 and np.all(rd.state(z) == np.arange(4.0)) and rd.control(z)[0] == 9.0
    assert rd.time(z) == 0.5 and rd.timestep(z) == 0.1 and not rd.is_terminal(z)
    zt = rd.KnotPoint(4, 1, np.r_[np.arange(4.0), 9.0], 1.0, 0.0)
    assert rd.is_terminal(zt) and rd.control(zt)[0] == 0.0          # src/knotpoint.jl:57-67
    with pytest.raises(AssertionError):
        rd.KnotPoint(4, 2, np.zeros(5), 0.0, 0.1)
    X, U = np.random.rand(11, 4), np.random.rand(10, 1)
    Z = rd.SampledTrajectory(X, U, dt=0.1)
    assert len(Z) == 11 and Z.data.shape == (11, 5) and Z.dts[-1] == 0.0 and np.allclose(Z.times, 0.1 * np.arange(11))
    assert np.all(rd.states(Z) == X) and np.all(rd.controls(Z)[:10] == U)
    assert rd.is_terminal(Z[10]) and not rd.is_terminal(Z[0])
    Z2 = rd.SampledTrajectory([Z[k] for k in range(11)])
    assert np.all(Z2.data == Z.data) and np.all(Z2.dts == Z.dts)


def test_dynamics_jacobian_views():
    """test/jacobian_test.jl:5-41: A/B are views into the column-major n x (n+m) data."""
    import rdb200 as rd
    D = rd.DynamicsJacobian(3, 2)
    assert D.A.shape == (3, 3) and D.B.shape == (3, 2) and np.asarray(D).shape == (3, 5)
    D.A[:] = 1.0
    D.B[:] = 2.0
    assert np.all(np.asarray(D)[:, :3] == 1) and np.all(np.asarray(D)[:, 3:] == 2)
    # memory image == Julia column-major Matrix(n, n+m): column j is contiguous
    D.data[...] = np.arange(15.0).reshape(5, 3)
    assert np.all(np.asarray(D)[:, 1] == [3, 4, 5])
    assert rd.DynamicsJacobian(2, 1, dtype=np.float32).data.dtype == np.float32


def test_partition_covers_and_aligns():
    import rdb200 as rd
    from rdb200 import sharding as sh
    for N in (0, 1, 63, 64, 1000, 1 << 20, (1 << 20) + 17):
        for W in (1, 2, 3, 4, 8):
            for align in (1, 64, 256):
                spans = [sh.partition(N, W, r, align) for r in range(W)]
                assert spans[0][0] == 0 and spans[-1][1] == N
                assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
                assert all(lo % align == 0 for lo, _ in spans if lo < N)
                sizes = [hi - lo for lo, hi in spans]
                assert max(sizes) - min(sizes) < 2 * align
    seg = sh.partition_segments({"cartpole": 2048, "quadrotor": 2048}, 8, 3)
    assert seg == {"cartpole": (768, 1024), "quadrotor": (768, 1024)}
    with pytest.raises(ValueError):
        sh.partition(10, 2, 2)


_GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
import rdb200
from rdb200 import sharding as sh
dist.init_process_group("gloo")
rank, W, _ = sh.world()
N = 1000
lo, hi = sh.partition(N, W, rank, 64)
local = torch.arange(lo, hi, dtype=torch.float64).reshape(-1, 1).repeat(1, 3)
counts = [sh.partition(N, W, r, 64) for r in range(W)]
full = sh.all_gather_shards(local, [b - a for a, b in counts])
assert full.shape == (N, 3) and torch.equal(full[:, 0], torch.arange(N, dtype=torch.float64))
t = sh.barrier_max_ms(10.0 + rank)
assert t == 10.0 + W - 1
dist.barrier()
dist.destroy_process_group()
print("OK", rank)
"""


def test_two_rank_gloo_sharding(tmp_path):
    """N>1 host path on CPU: world_size 2, gloo — partition, optional all-gather epilogue, max-over-ranks timing."""
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", str(script), ROOT], capture_output=True, text=True, timeout=300, env=env)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert p.stdout.count("OK") == 2


PENDULUM = """
        auto th = get<0>(x); auto om = get<1>(x);
        return vec(om, -p[0] * sin_(th) - p[1] * om + get<0>(u));
"""


def test_user_model_compiles_without_a_gpu_and_reports_errors():
    """user-written dynamics are NVRTC-compiled against the embedded headers; syntax errors come back as a log."""
    import rdb200 as rd
    ok, log = rd._abi.custom_check(2, 1, PENDULUM, 2)
    assert ok, log
    ok, log = rd._abi.custom_check(2, 1, PENDULUM, 2, rd.F32)
    assert ok, log
    ok, log = rd._abi.custom_check(2, 1, "return vec(get<1>(x), no_such_symbol);", 0)
    assert not ok and "no_such_symbol" in log
    assert not rd._abi.custom_check(30, 5, PENDULUM, 2)[0]            # n + m > 32


def test_c_abi_from_plain_c(tmp_path):
    """include/rdb200.h is plain C99 and the library links from C; without a GPU the example fails loudly (exit status 3)."""
    import torch
    exe = str(tmp_path / "rdb_example")
    libdir = os.path.join(ROOT, "robotdynamics.jl_b200")
    p = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "c_abi_example.c"),
                        "-o", exe, "-L", libdir, "-lrdb200", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    if torch.cuda.is_available():
        assert r.returncode == 0 and "[A B] of knot 0" in r.stdout, r.stdout + r.stderr
    else:
        assert r.returncode == 3 and "no CUDA device" in r.stderr, r.stdout + r.stderr


QUAD_WRENCH = """
                // test/quadrotor.jl:56-96 as a user wrench: p = [kf, km, L, gx, gy, gz]
                auto F1 = relu_(p[0] * get<0>(u)); auto F2 = relu_(p[0] * get<1>(u));
                auto F3 = relu_(p[0] * get<2>(u)); auto F4 = relu_(p[0] * get<3>(u));
                auto qF = quat_rotate<T>(q, vec(Zero{}, Zero{}, F1 + F2 + F3 + F4));
                return vec(mass * p[3] + get<0>(qF), mass * p[4] + get<1>(qF), mass * p[5] + get<2>(qF),
                           p[2] * (F2 - F4), p[2] * (F3 - F1), p[1] * (get<0>(u) - get<1>(u) + get<2>(u) - get<3>(u)));
"""


def test_user_rigid_body_wrench_compiles_without_a_gpu():
    """The reference's rigid-body extension interface (forces / moments, src/rigidbody.jl:244-257) as a user wrench body."""
    import rdb200 as rd
    ok, log = rd._abi.custom_rigid_check(rd._abi.ROT_QUAT, 0, 4, QUAD_WRENCH, 6, rd.F32)
    assert ok, log
    ok, log = rd._abi.custom_rigid_check(rd._abi.ROT_MRP, 1, 4, "return vec(get<0>(u), oops);", 0)
    assert not ok and "oops" in log


# ---- round 2: host-side logic that needs no GPU --------------------------------------------------------------------------------------
def test_buffer_validation_refuses_what_would_be_out_of_bounds():
    """_abi.check_buffer: caller-supplied arrays are bare pointers for the C ABI, so kind / dtype / shape / contiguity mismatches are
    refused in the host mirror (ADVICE: a float32 or short J with a float64 Z would be an out-of-bounds write)."""
    import rdb200
    cb = rdb200._abi.check_buffer
    Z = np.zeros((8, 5))
    assert cb("J", None, (8, 5, 4), Z) is None
    assert cb("J", np.empty((8, 5, 4)), (8, 5, 4), Z).shape == (8, 5, 4)
    with pytest.raises(TypeError):
        cb("J", np.empty((8, 5, 4), dtype=np.float32), (8, 5, 4), Z)
    with pytest.raises(ValueError):
        cb("J", np.empty((7, 5, 4)), (8, 5, 4), Z)
    with pytest.raises(ValueError):
        cb("J", np.empty((8, 4, 5)).transpose(0, 2, 1), (8, 5, 4), Z)
    with pytest.raises(ValueError):
        cb("J", [[0.0]], (1, 1), Z)
    import torch
    with pytest.raises(rdb200.RDBError) as e:
        cb("J", torch.empty((8, 5, 4), dtype=torch.float64), (8, 5, 4), Z)          # torch tensor with numpy inputs
    assert e.value.code == rdb200._abi.ERR_POINTER_MIX


def test_liestate_index_algebra_matches_the_reference_examples():
    """LieState / QuatState (src/liestate.jl:75-132): the docstring example QuatState(16, (4, 10)) is [v3, q, v2, q, v3]."""
    import rdb200 as rd
    ls = rd.QuatState(16, (4, 10))
    assert ls.P == (3, 2, 3) and len(ls) == 16
    assert rd.QuatState(13, (4,)).P == (3, 6) and len(rd.LieState(rd.MRP, 3, 6)) == 12         # RigidBody's LieState(R, (3, 6))
    assert len(rd.LieState(rd.RodriguesParam, (0, 4))) == 7 and rd.LieState(rd.QuatRotation, 0, 0, 0).P == (0, 0, 0)


def test_plan_and_trajectory_entry_points_refuse_bad_arguments_without_a_gpu():
    import rdb200
    L = _lib()
    vp = ctypes.c_void_p()
    assert L.rdb_plan_create(None, 3, 3, 1, 0, 8, None, None, None, 0.01, None, None, ctypes.byref(vp)) == rdb200._abi.ERR_ARG
    assert L.rdb_plan_launch(None, None) == rdb200._abi.ERR_ARG and L.rdb_plan_destroy(None) == 0
    assert L.rdb_plan_set_shared(None, 1) == rdb200._abi.ERR_ARG
    assert L.rdb_trajectory_create(None, 1, 4, 16, ctypes.byref(vp)) == rdb200._abi.ERR_ARG
    assert L.rdb_trajectory_rollout(None, 3, None) == rdb200._abi.ERR_ARG and L.rdb_trajectory_destroy(None) == 0
    assert L.rdb_dynamics_error(None, 3, 1, 8, None, None, 5, None, None, 0.01, None, None) == rdb200._abi.ERR_ARG
    parts = (ctypes.c_int * 3)(3, 2, 3)
    assert L.rdb_model_create_custom_lie(None, 1, 3, parts, 2, b"return x;", None, 0, ctypes.byref(vp)) == rdb200._abi.ERR_ARG


def test_user_model_bodies_with_time_compile_without_a_gpu():
    """dynamics(model, x, u, t): a user body may read `t` (a plain scalar); checked by the compile-only entry point (NVRTC, sm_100a)."""
    import rdb200 as rd
    ok, log = rd._abi.custom_check(2, 1, "return vec(get<1>(x), cos_(T(3) * t) * get<0>(u) - p[0] * sin_(get<0>(x)) + t * get<1>(x));", 1)
    assert ok, log
    ok, log = rd._abi.custom_check(2, 1, "return vec(get<1>(x), tt * get<0>(u));", 0, rd.F32)
    assert not ok and "tt" in log


def test_julia_shim_binds_only_exported_symbols():
    """julia/RobotDynamicsB200.jl cannot be executed here (no Julia): at least every `ccall((:rdb_..., LIB), ...)` in it must name a symbol the
    header declares and the library exports, with as many argument types as the C prototype has parameters."""
    import rdb200
    src = open(os.path.join(ROOT, "julia", "RobotDynamicsB200.jl")).read()
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "rdb200.h")).read(), flags=re.S)
    protos = {m.group(1): m.group(2) for m in re.finditer(r"\b(rdb_[a-z_]+)\s*\(([^;{]*?)\)\s*;", hdr)}
    calls = re.findall(r"ccall\(\(:(rdb_[a-z_]+), LIB\),\s*\w+,\s*\(([^)]*)\)", src)
    assert len(calls) > 25
    for name, argtypes in calls:
        assert name in rdb200._abi.SYMBOLS, name
        nargs_c = 0 if protos[name].strip() in ("", "void") else protos[name].count(",") + 1
        nargs_jl = len([a for a in argtypes.split(",") if a.strip()])
        assert nargs_c == nargs_jl, (name, nargs_c, nargs_jl)
