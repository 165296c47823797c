"""Error-state Jacobian kernels (fused G-seeded form), wide vs 32-knot tiles at 262144 knots for several rigid variants (RDB200_SMALL_N forces either;
each setting in its own process).  Development aid behind launch.cuh: prefers_narrow_tiles."""
import json, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = ("quad_quat_world", "quad_quat_body", "quad_mrp_world", "quad_mrp_body", "body_quat_world", "body_mrp_world", "body_mrp_body", "body_rp_body")


def child():
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import rdb200 as rd
    from common import rand_inputs, zoo
    N, out = 262144, {}
    for name in NAMES:
        for dtn in ("float32", "float64"):
            h = zoo()[name][1](rd)._h
            n, m = h.n, h.m
            Zs = [torch.from_numpy(rand_inputs(n, m, N, np.random.default_rng(i)).astype(dtn)).cuda() for i in range(3)]
            Js = [torch.empty((N, h.nerr + m, h.nerr), dtype=Zs[0].dtype, device="cuda") for _ in range(3)]
            plans = [rd._abi.Plan(h, rd._abi.OP_DISCRETE_ERROR_JACOBIAN, rd.RK4.code, Z, 0.01, J=J) for Z, J in zip(Zs, Js)]
            for i in range(5):
                plans[i % 3].launch()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(2_000_000); e0.record()
            for i in range(40):
                plans[i % 3].launch()
            e1.record(); torch.cuda.synchronize()
            out[f"{name}:{dtn}"] = round(e0.elapsed_time(e1) / 40 * 1e3, 1)
            del plans, Zs, Js
    print(json.dumps(out))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        res = {}
        for tag, thr in (("wide", "0"), ("t32", "100000000"), ("default", None)):
            env = dict(os.environ)
            if thr is not None:
                env["RDB200_SMALL_N"] = thr
            p = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
            res[tag] = json.loads(p.stdout.strip().splitlines()[-1]) if p.returncode == 0 else {}
            if p.returncode: print(p.stderr[-400:])
        print("| error-state RK4, 262144 knots | wide us | 32-knot us | library default us |"); print("|---|---|---|---|")
        for k in res["wide"]:
            print(f"| {k} | {res['wide'].get(k)} | {res['t32'].get(k)} | {res['default'].get(k)} |")
