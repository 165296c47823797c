"""The arithmetic the CUDA kernels execute, checked WITHOUT a GPU: tests/host_harness/host_check.cu instantiates the product's own
device headers (csrc/sdual.cuh, models.cuh, integrators.cuh — `__host__ __device__` templates) in a host program; its discrete
Jacobians and next states are compared with the CPU checker.  Covers what round 2 changed in those headers — the elemental Cartpole
solve, the elemental rotations / kinematics of the rigid bodies, rolled and increment-form stage loops — for fp64.  (The kernels
proper, their TMA staging and every dtype / layout are the -m gpu suite's job.)"""
import os
import shutil
import subprocess
import tempfile

import numpy as np
import pytest

from oracle import rd_oracle as o
from common import rand_inputs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host_harness", "host_check.cu")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ROT = {1: o.ROT_QUAT, 2: o.ROT_MRP, 3: o.ROT_RP}
RULE = {0: o.EULER, 1: o.RK2, 2: o.RK3, 3: o.RK4}

# (kind, rot, frame, rule, roll): Cartpole all rules; the rigid-body variants that were furthest below the roofline in round 1
CASES = [(0, 0, 0, 3, 0), (0, 0, 0, 2, 0), (0, 0, 0, 1, 0),
         (1, 1, 1, 3, 1),      # quadrotor, body frame, RK4 rolled: two elemental rotations per stage
         (2, 2, 0, 1, 0),      # satellite-type body, MRP, RK2 increment form (BASELINE C4's kernel)
         (1, 2, 0, 3, 1),      # quadrotor{MRP}: elemental thrust rotation + elemental MRP kinematics
         (2, 3, 1, 3, 0),      # body, Rodrigues, body frame, RK4 unrolled increment form
         (1, 1, 1, 2, 0),      # body-frame models: q \ (q * Fb + Gw) evaluated as |q|^4 Fb + q \ Gw (models.cuh, SPLIT) — quadrotor, RK3
         (2, 1, 1, 3, 0),      # ... body with a body-frame force input, state quaternion (|q|^4 factor live off the manifold)
         (1, 2, 1, 3, 0),      # ... quadrotor{MRP}, body frame
         (2, 2, 1, 1, 1)]      # ... body{MRP}, body frame, RK2


@pytest.fixture(scope="module")
def bindir():
    if not os.path.exists(NVCC):
        pytest.skip("nvcc not available")
    d = tempfile.mkdtemp(prefix="rdb_host_")
    yield d
    shutil.rmtree(d, ignore_errors=True)


@pytest.mark.parametrize("kind,rot,frame,rule,roll", CASES)
def test_device_templates_on_host_match_the_checker(bindir, kind, rot, frame, rule, roll):
    exe = os.path.join(bindir, f"hc_{kind}_{rot}_{frame}_{rule}_{roll}")
    subprocess.check_call([NVCC, "-std=c++17", "-O1", "-w", "-I", os.path.join(ROOT, "robotdynamics.jl_b200", "csrc"), f"-DHK={kind}", f"-DHR={rot}",
                           f"-DHF={frame}", f"-DHQ={rule}", f"-DHROLL={roll}", "-o", exe, SRC])
    if kind == 0:
        om, pr = o.cartpole(), [1.0, 0.2, 0.5, 9.81] + [0.0] * 12
    elif kind == 1:
        om, pr = o.quadrotor(ROT[rot], frame), [0.5, 0.0023, 0, 0, 0, 0.0023, 0, 0, 0, 0.004, 0, 0, -9.81 * 0.5, 0.175, 1.0, 0.0245]
    else:
        om, pr = o.body(ROT[rot], frame), [2.0, 2, 0, 0, 0, 3, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0]
    N, h = 40, 0.05
    Z = rand_inputs(om.n, om.m, N, np.random.default_rng(5))
    if kind == 1:
        Z[::7, om.n] = 0.0; Z[::5, om.n + 1] = -0.3          # thrust clamp: exact ties and negative controls
    if om.n == 13:
        Z[::3, 3:7] *= np.linspace(0.8, 1.2, len(Z[::3]))[:, None]      # off-manifold state quaternions (never renormalised by the path)
    text = " ".join(map(repr, pr)) + f"\n{N} {h}\n" + "\n".join(" ".join(repr(float(v)) for v in row) for row in Z) + "\n"
    out = subprocess.run([exe], input=text, capture_output=True, text=True, check=True).stdout
    rows = np.array([[float(v) for v in ln.split()] for ln in out.strip().splitlines()])
    nz = om.n + om.m
    J, xn = rows[:, :om.n * nz].reshape(N, nz, om.n), rows[:, om.n * nz:]
    ref_J, ref_x = o.discrete_jacobian(om, RULE[rule], Z, h), o.discrete_dynamics(om, RULE[rule], Z, h)
    assert np.abs(J - ref_J).max() < 1e-12 * max(1.0, np.abs(ref_J).max()) and np.abs(xn - ref_x).max() < 1e-13 * max(1.0, np.abs(ref_x).max())
