"""Turns the round-2 ncu reports brought back under gpurun_out/r02*/ into the committed summaries under profiles/r02/:
per report `<name>_details.txt` (ncu --page details) and one `ncu_summary_r02.json` with the headline raw metrics of every captured
kernel; refreshes profiles/edge_step_summary.json (DRAM bytes per launch of the dominant kernel, read by bench.py for `roofline.traffic`)
and copies the launch lists.      python scripts/summarize_ncu_r02.py
"""
import csv, glob, io, json, os, shutil, subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "r02")
os.makedirs(OUT, exist_ok=True)
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__block_size", "launch__grid_size", "launch__waves_per_multiprocessor", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum",
        "sm__cycles_active.avg", "sm__cycles_elapsed.max",
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
import re
for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "r02*", "*.ncu-rep"))):
    name = os.path.splitext(os.path.basename(rep))[0]
    if name == "ens_kernels_r02a":            # caught only the set-up scatter launches; ens_kernels_r02c has the step kernels
        continue
    kernels = raw(rep)
    for k in kernels:
        if "dram__bytes_read.sum" in k:
            k["dram_bytes_per_launch"] = to_bytes(k["dram__bytes_read.sum"]) + to_bytes(k["dram__bytes_write.sum"])
    summary[name] = kernels
    with open(os.path.join(OUT, f"{name}_details.txt"), "w") as f:
        f.write(subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout)
    for k in kernels:
        print(f"{name:34s} {k['kernel'][:70]:70s} {k.get('gpu__time_duration.sum', {}).get('value', '?'):>8s} us  "
              f"{k.get('dram_bytes_per_launch', 0) / 1e6:8.1f} MB  regs {k.get('launch__registers_per_thread', {}).get('value', '?')}")
with open(os.path.join(OUT, "ncu_summary_r02.json"), "w") as f:
    json.dump(summary, f, indent=1)
for lst in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "r02*", "launches_*.csv"))):
    shutil.copy(lst, os.path.join(OUT, os.path.basename(lst)))
# DRAM bytes per launch of the dominant kernel (655,362 cells, one GPU), keyed by the launched kernel's name
edge = {}
for name, kernels in summary.items():
    for k in kernels:
        for key in ("edge_step_pipe16_kernel", "edge_step_pipe_kernel", "edge_step_kernel"):
            if re.search(r"(?<![A-Za-z0-9_])" + key + r"[(<]", k["kernel"]):
                edge[key] = {"dram_bytes_per_launch": k["dram_bytes_per_launch"], "source": f"profiles/r02/{name}_details.txt (ncu --set full, 655,362 cells)"}
if edge:
    with open(os.path.join(ROOT, "profiles", "edge_step_summary.json"), "w") as f:
        json.dump({"kernels": edge}, f, indent=1)
    print("edge_step_summary.json:", {k: v["dram_bytes_per_launch"] / 1e6 for k, v in edge.items()})
