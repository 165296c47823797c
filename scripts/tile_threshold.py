"""Where should the rigid-body RK4 Jacobian kernels switch from 32-knot to wide tiles?  Times back-to-back plan launches (rotating buffers, CUDA
events) at several N with the threshold forced low (wide tiles everywhere) and high (32-knot tiles everywhere) through RDB200_SMALL_N — each setting
in its own process, because the library reads the variable once.

    python scripts/tile_threshold.py            # driver: runs both settings, prints the table
"""
import json, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NS = (8192, 16384, 32768, 65536, 98304, 131072, 196608, 262144)
MODELS = ("quad32", "quad64", "quadbody32", "sat64")


def child():
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import rdb200 as rd
    from common import rand_inputs
    mk = {"quad32": (lambda: rd.Quadrotor(), "float32"), "quad64": (lambda: rd.Quadrotor(), "float64"),
          "quadbody32": (lambda: rd.Quadrotor(bodyframe=True), "float32"), "sat64": (lambda: rd.Satellite(rd.MRP), "float64")}
    out = {}
    for name in MODELS:
        h = mk[name][0]()._h; dtn = mk[name][1]
        n, m = h.n, h.m
        for N in NS:
            per = np.dtype(dtn).itemsize * ((n + m) + n * (n + m))
            nsets = max(2, int(np.ceil(400e6 / (N * per))) + 1)
            nsets = min(nsets, 24)
            Zs = [torch.from_numpy(rand_inputs(n, m, N, np.random.default_rng(i)).astype(dtn)).cuda() for i in range(nsets)]
            Js = [torch.empty((N, n + m, n), dtype=Zs[0].dtype, device="cuda") for _ in range(nsets)]
            plans = [rd._abi.Plan(h, rd._abi.OP_DISCRETE_JACOBIAN, rd.RK4.code, Z, 0.01, J=J) for Z, J in zip(Zs, Js)]
            for i in range(5):
                plans[i % nsets].launch()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(2_000_000); e0.record()
            for i in range(60):
                plans[i % nsets].launch()
            e1.record(); torch.cuda.synchronize()
            out[f"{name}:{N}"] = round(e0.elapsed_time(e1) / 60 * 1e3, 2)
            del plans, Zs, Js
    print(json.dumps(out))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        res = {}
        for tag, thr in (("wide", "0"), ("t32", "100000000")):
            p = subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, RDB200_SMALL_N=thr), capture_output=True, text=True)
            res[tag] = json.loads(p.stdout.strip().splitlines()[-1]) if p.returncode == 0 else {"error": p.stderr[-500:]}
        print("| model | N | wide tiles us | 32-knot tiles us |"); print("|---|---|---|---|")
        for name in MODELS:
            for N in NS:
                k = f"{name}:{N}"
                print(f"| {name} | {N} | {res['wide'].get(k)} | {res['t32'].get(k)} |")
