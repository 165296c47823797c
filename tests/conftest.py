import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """The C-ABI library and the oracle are build products (git-ignored): on a fresh checkout build them once, here, with nvcc /
    g++ (no GPU needed) — the same thing __graft_entry__.build() does."""
    lib = os.path.join(ROOT, "robotdynamics.jl_b200", "librdb200.so")
    if not os.path.exists(lib):
        sys.stderr.write("[tests] librdb200.so missing: building it (nvcc, a few minutes)...\n")
        subprocess.check_call([sys.executable, os.path.join(ROOT, "robotdynamics.jl_b200", "csrc", "build.py")])
    if not os.path.exists(os.path.join(ROOT, "oracle", "librd_oracle.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
