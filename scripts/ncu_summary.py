"""Summarise .ncu-rep captures (read here with `ncu -i ... --page raw --csv`) into profiles/*.md and profiles/traffic.json."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (registers), CTAs/SM"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem), CTAs/SM"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active %"),
    ("smsp__warps_active.avg.per_cycle_active", "warps active / scheduler"),
    ("smsp__warps_eligible.avg.per_cycle_active", "warps eligible / scheduler"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("sm__cycles_elapsed.max", "SM cycles elapsed"),
    ("smsp__cycles_active.avg", "SMSP cycles active (avg)"),
]
UNITS = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}


def load(rep):
    # a report, or its raw page already exported on the GPU box (`ncu -i rep --page raw --csv > rep.csv`: the reports are too large to bring back)
    out = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h = rows[0]
    return h, rows[1], rows[2:]


def summarise(rep, title):
    h, units, data = load(rep)
    r = data[-1]
    lines = [f"### {title}", "", f"`{os.path.basename(rep)}` — kernel `{r[h.index('Kernel Name')][:90]}`", "", "| metric | value |", "|---|---|"]
    vals = {}
    for key, label in WANT:
        if key in h:
            i = h.index(key)
            lines.append(f"| {label} (`{key}`) | {r[i]} {units[i]} |")
            try:
                vals[key] = float(r[i].replace(",", "")) * UNITS.get(units[i], 1.0)
            except ValueError:
                pass
    stalls = []
    for i, name in enumerate(h):
        if name.startswith("smsp__average_warps_issue_stalled") and name.endswith("per_issue_active.ratio"):
            try:
                stalls.append((float(r[i]), name[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
    stalls.sort(reverse=True)
    lines.append("| top stall reasons (warps stalled per issue) | " + ", ".join(f"{n} {v:.2f}" for v, n in stalls[:7]) + " |")
    traffic = vals.get("dram__bytes_read.sum", 0) + vals.get("dram__bytes_write.sum", 0)
    lines.append(f"| DRAM traffic per launch (read + write) | {traffic / 1e6:.1f} MB |")
    return "\n".join(lines) + "\n", traffic


if __name__ == "__main__":
    # usage: ncu_summary.py out.md name=path.ncu-rep[:title] ...
    out_md = sys.argv[1]
    traffic_path = os.path.join(ROOT, "profiles", "traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    parts = []
    for arg in sys.argv[2:]:
        name, rest = arg.split("=", 1)
        path, _, title = rest.partition(":")
        md, t = summarise(path, title or name)
        parts.append(md)
        traffic[name] = t
    with open(out_md, "a") as f:
        f.write("\n".join(parts) + "\n")
    json.dump(traffic, open(traffic_path, "w"), indent=1)
