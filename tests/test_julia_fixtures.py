"""The pin against the REAL reference (RobotDynamics.jl v0.4.8 run in Julia): oracle/ref_julia/gen_fixtures.jl turns the committed
seeded inputs tests/golden/julia_in/*.npy into the reference's own outputs tests/golden/julia_out/*.npy.  Whenever those outputs
exist, the CPU oracle (here) and the CUDA path (-m gpu) are compared with them at the north_star tolerance; until somebody with a
Julia installation has run the generator, the comparisons SKIP with a message that says so (parity status: unpinned against Julia
output) — the inputs, the recipe and the consumers are already in place and checked."""
import importlib.util
import os
import re

import numpy as np
import pytest

from oracle import rd_oracle as o
from common import GOLDEN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JIN, JOUT = os.path.join(GOLDEN, "julia_in"), os.path.join(GOLDEN, "julia_out")
GEN = os.path.join(ROOT, "oracle", "ref_julia", "gen_fixtures.jl")
TOL = 1e-10                                                  # north_star: 1e-10 (fp64) vs the Julia / ForwardDiff path

_spec = importlib.util.spec_from_file_location("make_inputs", os.path.join(ROOT, "oracle", "ref_julia", "make_inputs.py"))
make_inputs = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(make_inputs)

ROT = {"quat": o.ROT_QUAT, "mrp": o.ROT_MRP, "rp": o.ROT_RP}
FRAME = {"world": o.WORLD, "body": o.BODYFRAME}


def jin(name):
    return np.load(os.path.join(JIN, name + ".npy"))


def jout(name):
    p = os.path.join(JOUT, name + ".npy")
    if not os.path.exists(p):
        pytest.skip(f"{os.path.relpath(p, ROOT)} absent: Julia reference outputs not generated yet "
                    "(run oracle/ref_julia/gen_fixtures.jl; parity stays 'unpinned against Julia output' until then)")
    return np.load(p)


def have_outputs():
    return os.path.isdir(JOUT) and any(f.endswith(".npy") for f in os.listdir(JOUT))


# ---- the recipe itself (always runs) -------------------------------------------------------------------------------------------------
def test_julia_inputs_are_committed_and_reproducible():
    want = make_inputs.inputs()
    assert sorted(f[:-4] for f in os.listdir(JIN) if f.endswith(".npy")) == sorted(want)
    for name, a in want.items():
        assert np.array_equal(jin(name), a), name


def test_generator_reads_every_input_and_names_the_reference_api():
    src = open(GEN).read()
    for name in make_inputs.inputs():
        stem = re.sub(r"_(Z|X0)$", "", name)
        assert stem in src or re.sub(r"^(quad|body|lie)_(quat|mrp|rp)$", r"\1_$(rot)", stem) in src, name
    for api in ("RD.jacobian!(sig, diff, dmodel, Jk, y, z)", "RD.StaticReturn()", "RD.ForwardAD()", "RD.UserDefined()", "RD.errstate_jacobian!",
                "RD.∇errstate_jacobian!", "RD.state_diff", "RD.DiscretizedDynamics{RD.RK4}", "RD.ImplicitMidpoint", "ForwardDiff.jacobian"):
        assert api in src, api
    proj = open(os.path.join(ROOT, "oracle", "ref_julia", "Project.toml")).read()
    assert 'RobotDynamics = "=0.4.8"' in proj and 'Rotations = "1"' in proj and 'ForwardDiff = "0.10"' in proj


def test_npy_layout_matches_the_julia_reader():
    """C-order (N, n+m) on disk is what read_npy() reinterprets as the column-major Matrix (n+m, N); version-1 header, '<f8'."""
    with open(os.path.join(JIN, "c2_cartpole_Z.npy"), "rb") as f:
        head = f.read(128)
    assert head[:6] == b"\x93NUMPY" and head[6] == 1 and b"'descr': '<f8'" in head and b"'fortran_order': False" in head


# ---- consumers: every evaluator behind one interface --------------------------------------------------------------------------------
class OracleEval:
    def model(self, kind, rot="quat", frame="world"):
        return {"cartpole": o.cartpole, "quad": lambda: o.quadrotor(ROT[rot], FRAME[frame]), "body": lambda: o.body(ROT[rot], FRAME[frame]),
                "satellite": lambda: o.satellite(ROT[rot])}[kind]()

    discrete_jacobian = staticmethod(lambda m, Q, Z, dt: (o.discrete_jacobian(m, Q, Z, dt), o.discrete_dynamics(m, Q, Z, dt)))
    jacobian = staticmethod(lambda m, Z: (o.jacobian(m, Z), o.dynamics(m, Z)))
    errstate_jacobian = staticmethod(lambda m, X: o.errstate_jacobian(m, X))
    grad_errstate_jacobian = staticmethod(lambda m, X, B: o.grad_errstate_jacobian(m, X, B))
    state_diff = staticmethod(lambda m, X, X0: o.state_diff(m, X, X0))


class GpuEval:
    def __init__(self):
        import rdb200
        self.rd = rdb200

    def model(self, kind, rot="quat", frame="world"):
        rd = self.rd
        R = {"quat": rd.QuatRotation, "mrp": rd.MRP, "rp": rd.RodriguesParam}[rot]
        return {"cartpole": rd.Cartpole, "quad": lambda: rd.Quadrotor(R, bodyframe=frame == "body"), "body": lambda: rd.Body(R, bodyframe=frame == "body"),
                "satellite": lambda: rd.Satellite(R)}[kind]()

    def discrete_jacobian(self, m, Q, Z, dt):
        xn = np.empty((Z.shape[0], m._h.n))
        return m._h.discrete_jacobian(Q, Z, dt, xn=xn), xn

    def jacobian(self, m, Z):
        xd = np.empty((Z.shape[0], m._h.n))
        return m._h.jacobian(Z, xdot=xd), xd

    def errstate_jacobian(self, m, X):
        return m._h.errstate_jacobian(np.ascontiguousarray(X))

    def grad_errstate_jacobian(self, m, X, B):
        return m._h.grad_errstate_jacobian(np.ascontiguousarray(X), np.ascontiguousarray(B))

    def state_diff(self, m, X, X0):
        return m._h.state_diff(np.ascontiguousarray(X), np.ascontiguousarray(X0))


def close(a, ref, what, tol=TOL):
    scale = max(1.0, float(np.abs(ref).max()))
    err = float(np.abs(np.asarray(a) - ref).max())
    assert err < tol * scale, f"{what}: max|ours - Julia| = {err:.3e} (scale {scale:.1f})"


def run_all(ev):
    """Compare evaluator `ev` with every Julia output that exists; returns the number of arrays compared."""
    n = 0
    for name, kind, kw, Q, dt in (("c1_cartpole", "cartpole", {}, o.RK3, 0.01), ("c2_cartpole", "cartpole", {}, o.RK4, 0.01),
                                  ("c3_quadrotor", "quad", {}, o.RK4, 0.01), ("c3_quadrotor_dt01", "quad", {}, o.RK4, 0.1),
                                  ("c4_satellite_mrp", "satellite", {"rot": "mrp"}, o.RK2, 0.1)):
        J, xn = ev.discrete_jacobian(ev.model(kind, **kw), Q, jin(name + "_Z"), dt)
        close(J, jout(name + "_J"), name + " J"); close(xn, jout(name + "_xn"), name + " x+"); n += 2
    close(ev.discrete_jacobian(ev.model("cartpole"), o.RK4, jin("c2_cartpole_Z"), 0.01)[0], jout("c2_cartpole_J_userdefined"), "chain rule"); n += 1
    for name in ("quad_offmanifold", "quad_tie"):            # off-manifold q*r / q\\r and the thrust-clamp tie
        m = ev.model("quad")
        J, xd = ev.jacobian(m, jin(name + "_Z"))
        close(J, jout(name + "_J"), name + " continuous J"); close(xd, jout(name + "_xdot"), name + " xdot")
        Jd, xn = ev.discrete_jacobian(m, o.RK4, jin(name + "_Z"), 0.05)
        close(Jd, jout(name + "_Jd"), name + " RK4 J"); close(xn, jout(name + "_xn"), name + " x+"); n += 4
    for rot in ROT:
        for frame in FRAME:
            for kind in ("quad", "body"):
                J, xn = ev.discrete_jacobian(ev.model(kind, rot, frame), o.RK4, jin(f"{kind}_{rot}_Z"), 0.05)
                close(J, jout(f"{kind}_{rot}_{frame}_J"), f"{kind}_{rot}_{frame} J"); close(xn, jout(f"{kind}_{rot}_{frame}_xn"), "x+"); n += 2
        m = ev.model("quad", rot)
        X, X0 = jin(f"quad_{rot}_Z"), jin(f"lie_{rot}_X0")
        close(ev.errstate_jacobian(m, X), jout(f"lie_{rot}_G"), f"G {rot}")
        close(ev.grad_errstate_jacobian(m, X, X0), jout(f"lie_{rot}_dG"), f"dG {rot}")
        close(ev.state_diff(m, X, X0), jout(f"lie_{rot}_dx"), f"state_diff {rot}"); n += 3
    IM = 5 if isinstance(ev, OracleEval) else 4               # ImplicitMidpoint code: oracle / C ABI
    for name, kind in (("implicit_cartpole", "cartpole"), ("implicit_quadrotor", "quad")):
        J, xn = ev.discrete_jacobian(ev.model(kind), IM, jin(name + "_Z"), 0.05)
        close(J, jout(name + "_J"), name + " J", 1e-8); close(xn, jout(name + "_xn"), name + " x+"); n += 2
    return n


def test_oracle_matches_julia_outputs():
    if not have_outputs():
        pytest.skip("tests/golden/julia_out/ is empty: parity unpinned against Julia output (run oracle/ref_julia/gen_fixtures.jl)")
    assert o.IMPLICIT_MIDPOINT == 5
    assert run_all(OracleEval()) > 40


@pytest.mark.gpu
def test_cuda_path_matches_julia_outputs():
    if not have_outputs():
        pytest.skip("tests/golden/julia_out/ is empty: parity unpinned against Julia output (run oracle/ref_julia/gen_fixtures.jl)")
    assert run_all(GpuEval()) > 40


def test_oracle_and_recipe_agree_on_shapes_without_julia():
    """Dry run of the consumer against the oracle's own outputs: proves every comparison in run_all() is well-formed (shapes, model
    constructors, integrator codes), so a freshly generated julia_out/ is consumed without edits."""
    ev = OracleEval()
    Z = jin("body_mrp_Z")
    J, xn = ev.discrete_jacobian(ev.model("body", "mrp", "body"), o.RK4, Z, 0.05)
    assert J.shape == (16, 18, 12) and xn.shape == (16, 12)
    X, X0 = jin("quad_rp_Z"), jin("lie_rp_X0")
    m = ev.model("quad", "rp")
    assert ev.errstate_jacobian(m, X).shape == (16, 12, 12) and ev.grad_errstate_jacobian(m, X, X0).shape == (16, 12, 12)
    assert ev.state_diff(m, X, X0).shape == (16, 12)
    Jc, xd = ev.jacobian(ev.model("quad"), jin("quad_tie_Z"))
    assert Jc.shape == (16, 17, 13) and xd.shape == (16, 13)
    Ji, _ = ev.discrete_jacobian(ev.model("quad"), o.IMPLICIT_MIDPOINT, jin("implicit_quadrotor_Z"), 0.05)
    assert Ji.shape == (16, 17, 13) and np.isfinite(Ji).all()
