"""Turns the ncu reports brought back in gpurun_out/ into the committed summaries under profiles/.
    python scripts/summarize_ncu.py <tag>      (tag e.g. r01c)
"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__block_size", "launch__grid_size", "launch__waves_per_multiprocessor", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for k in KEYS:
            if k in hdr:
                d[k] = {"value": r[hdr.index(k)], "unit": units[hdr.index(k)]}
        res.append(d)
    return res


def to_bytes(m):
    v = float(m["value"]); u = m["unit"].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]


summary = {}
for name in ("edge_step", "cell_step"):
    rep = os.path.join(ROOT, "gpurun_out", f"{name}_{tag}.ncu-rep")
    if not os.path.exists(rep):
        continue
    k = raw(rep)[0]
    k["dram_bytes_per_launch"] = to_bytes(k["dram__bytes_read.sum"]) + to_bytes(k["dram__bytes_write.sum"])
    summary[name] = k
    with open(os.path.join(ROOT, "profiles", f"{name}_{tag}_details.txt"), "w") as f:
        f.write(subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout)
with open(os.path.join(ROOT, "profiles", f"ncu_summary_{tag}.json"), "w") as f:
    json.dump(summary, f, indent=1)
if "edge_step" in summary:
    with open(os.path.join(ROOT, "profiles", "edge_step_summary.json"), "w") as f:
        json.dump({"source": f"profiles/ncu_summary_{tag}.json (ncu --set full, 655,362 cells)",
                   "dram_bytes_per_launch": summary["edge_step"]["dram_bytes_per_launch"]}, f, indent=1)
for name, k in summary.items():
    print(name, k["gpu__time_duration.sum"], "dram bytes/launch", k["dram_bytes_per_launch"] / 1e6, "MB",
          "regs", k["launch__registers_per_thread"]["value"], "dram%", k["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]["value"])
