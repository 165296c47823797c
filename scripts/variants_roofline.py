"""Merge the un-profiled variants sweep (scripts/variants_sweep.py: time per launch) with the ncu pass over the same sweep
(RDB_SWEEP_STEPS=1 RDB_SWEEP_WARM=1, two launches per row, the second one is read) into one table that names, per kernel, the
roofline that binds it: HBM (algorithmic bytes / measured copy bandwidth), instruction issue (1 warp instruction per clock and SM
sub-partition) or the FP64 pipe (1 warp DFMA per 2 clocks and sub-partition; measured 60.6 DFMA/clk/SM, profiles/tuning_r01.md).

    python scripts/variants_roofline.py gpurun_out/variants.md gpurun_out/variants_ncu.csv > profiles/variants_rNN.md
"""
import csv
import sys

SMS, SMSP, CLK = 148, 4, 1.965e9
M = {"gpu__time_duration.sum": "t", "smsp__inst_executed.sum": "winst",
     "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fma",
     "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64", "launch__registers_per_thread": "regs",
     "sm__inst_executed_pipe_fp64.sum": "w64", "smsp__warps_active.avg.per_cycle_active": "warps"}


def ncu_rows(path):
    """per launch (in order): dict of the metrics above"""
    lines = open(path).read().splitlines()
    start = next(i for i, ln in enumerate(lines) if ln.startswith('"ID"'))
    launches = {}
    for r in csv.DictReader(lines[start:]):
        d = launches.setdefault(int(r["ID"]), {"name": r["Kernel Name"]})
        if r["Metric Name"] in M:
            try:
                d[M[r["Metric Name"]]] = float(r["Metric Value"].replace(",", ""))
            except ValueError:
                pass
    return [launches[k] for k in sorted(launches)]


def main(table, ncu):
    rows = [ln for ln in open(table).read().splitlines() if ln.startswith("| ") and not ln.startswith("| model") and not ln.startswith("|---")]
    prof = ncu_rows(ncu)
    assert len(prof) == 2 * len(rows), (len(prof), len(rows))
    print("| model | dtype | rule | N | us | evals/s | algorithmic GB/s | of HBM peak | warp inst / knot x32 | regs | issue slots busy | FMA pipe | FP64 pipe | most utilised resource (HBM = algorithmic fraction of the measured copy bandwidth; issue / pipes = ncu, single cold launch) | max err vs checker |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    for ln, p in zip(rows, prof[1::2]):
        c = [x.strip() for x in ln.strip("|").split("|")]
        name, dt, rule, N, us, evs, gbs, frac, err = c
        N = int(N)
        tinst = p["winst"] * 32 / N
        hbm = float(frac)
        cands = {"HBM": hbm, "issue": p.get("issue", 0) / 100, "FP64 pipe": p.get("fp64", 0) / 100, "FMA pipe": p.get("fma", 0) / 100}
        b = max(cands, key=cands.get)
        print(f"| {name} | {dt} | {rule} | {N} | {us} | {evs} | {gbs} | {frac} | {tinst:.0f} | {p.get('regs', 0):.0f} | {p.get('issue', 0):.0f} % | "
              f"{p.get('fma', 0):.0f} % | {p.get('fp64', 0):.0f} % | {b}: {cands[b]:.2f} | {err} |")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
