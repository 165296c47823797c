"""Builds librdb200.so (the C-ABI library) for sm_100a with nvcc, in-tree.

    python robotdynamics.jl_b200/csrc/build.py [-j JOBS] [--force] [--verbose]

Every (model family, rotation, frame, dtype) is one nvcc job on unit.cu with different -D flags; objects land in
csrc/_obj/ (git-ignored), the library next to the package as robotdynamics.jl_b200/librdb200.so (git-ignored, but
it travels to the GPU box with the snapshot).  Incremental: an object is rebuilt when any source/header is newer.
"""
import argparse
import concurrent.futures as cf
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(PKG, "librdb200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC", "-I", HERE, "-I", os.path.join(ROOT, "include")]

HEADERS = ["sdual.cuh", "models.cuh", "integrators.cuh", "kernels.cuh", "launch.cuh", "units.h", "lie.h"]
ROTS = {"quat": 1, "mrp": 2, "rp": 3}
FRAMES = {"world": 0, "body": 1}


def units():
    """(symbol suffix, -D flags) for every unit.cu job; must match units.h."""
    out = []
    for dt, dn in ((0, "f32"), (1, "f64")):
        out.append((f"cartpole_{dn}", dict(RDB_KIND=0, RDB_DTYPE=dt)))
        for d in (1, 2, 3):
            out.append((f"di{d}_{dn}", dict(RDB_KIND=3, RDB_DI_D=d, RDB_DTYPE=dt)))
        for kind, kn in ((1, "quad"), (2, "body")):
            for rn, r in ROTS.items():
                for fn, f in FRAMES.items():
                    out.append((f"{kn}_{rn}_{fn}_{dn}", dict(RDB_KIND=kind, RDB_ROT=r, RDB_FRAME=f, RDB_DTYPE=dt)))
    return out


def stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def run(cmd, verbose):
    t0 = time.time()
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if p.returncode != 0:
        raise RuntimeError("command failed: " + " ".join(cmd) + "\n" + p.stdout[-4000:])
    if verbose:
        sys.stdout.write(p.stdout)
    return time.time() - t0


def build(jobs=None, force=False, verbose=False, ptxas_v=False):
    os.makedirs(OBJ, exist_ok=True)
    hdrs = [os.path.join(HERE, h) for h in HEADERS] + [os.path.join(ROOT, "include", "rdb200.h")]
    extra = ["-Xptxas", "-v"] if ptxas_v else []
    tasks = []
    for name, defs in units():
        obj = os.path.join(OBJ, f"unit_{name}.o")
        src = os.path.join(HERE, "unit.cu")
        if force or stale(obj, hdrs + [src]):
            flags = [f"-D{k}={v}" for k, v in defs.items()] + [f"-DRDB_UNIT_NAME={name}"]
            tasks.append((name, [NVCC] + ARCH + COMMON + extra + flags + ["-c", "-o", obj, src]))
    for srcname in ("abi.cu", "lie.cu"):
        obj = os.path.join(OBJ, srcname.replace(".cu", ".o"))
        src = os.path.join(HERE, srcname)
        if force or stale(obj, hdrs + [src]):
            tasks.append((srcname, [NVCC] + ARCH + COMMON + extra + ["-c", "-o", obj, src]))
    jobs = jobs or min(len(tasks), os.cpu_count() or 4) or 1
    t0 = time.time()
    if tasks:
        # longest jobs first (rigid-body fp64 units dominate)
        tasks.sort(key=lambda t: (("f64" in t[0]) * 2 + ("quad" in t[0] or "body" in t[0]) * 4), reverse=True)
        with cf.ThreadPoolExecutor(max_workers=jobs) as ex:
            futs = {ex.submit(run, cmd, verbose or ptxas_v): name for name, cmd in tasks}
            for f in cf.as_completed(futs):
                dt = f.result()
                if verbose:
                    print(f"[build] {futs[f]}: {dt:.1f}s", flush=True)
    objs = [os.path.join(OBJ, f"unit_{name}.o") for name, _ in units()] + [os.path.join(OBJ, "abi.o"), os.path.join(OBJ, "lie.o")]
    if tasks or force or stale(LIB, objs):
        run([NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-cudart", "static"], verbose)
    if verbose:
        print(f"[build] {len(tasks)} objects rebuilt in {time.time() - t0:.1f}s -> {LIB}")
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("-j", type=int, default=None)
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", "-v", action="store_true")
    ap.add_argument("--ptxas-v", action="store_true")
    a = ap.parse_args()
    build(a.j, a.force, True if a.verbose else False, a.ptxas_v)
    print(LIB)
