"""Static evidence from the built objects (no GPU needed): per kernel, registers / shared memory (cuobjdump -res-usage) and the
counts of the SASS mnemonics that matter here — UBLKCP (cp.async.bulk, TMA engine), SYNCS (mbarrier), DMMA (FP64 tensor cores),
DFMA/DMUL/DADD, SHFL, ST/LD on .SYS scope (peer-memory flags).    python scripts/sass_evidence.py > profiles/sass_evidence_rNN.txt"""
import os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "geodesicodis_b200", "_build")
PATS = [("UBLKCP", r"\bUBLKCP"), ("SYNCS", r"\bSYNCS"), ("DMMA", r"\bDMMA"), ("DP", r"\bD(FMA|MUL|ADD)\b"), ("SHFL", r"\bSHFL"), ("LDG", r"\bLDG"),
        ("STG", r"\bSTG"), ("SYS-scope", r"\.SYS\b"), ("MUFU.RCP64H", r"MUFU\.RCP64H")]


def demangle(name):
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    except OSError:
        return name


rows = []
for obj in sorted(os.listdir(BUILD)):
    if not obj.endswith(".cu.o"):
        continue
    path = os.path.join(BUILD, obj)
    res = subprocess.run(["cuobjdump", "-res-usage", path], capture_output=True, text=True).stdout
    usage = {}
    for m in re.finditer(r"Function (\S+):\n\s+REG:(\d+) STACK:(\d+) SHARED:(\d+)", res):
        usage[m.group(1)] = (int(m.group(2)), int(m.group(3)), int(m.group(4)))
    sass = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    cur, counts = None, {}
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1); counts[cur] = {k: 0 for k, _ in PATS}; counts[cur]["instr"] = 0
            continue
        if cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            counts[cur]["instr"] += 1
            for k, pat in PATS:
                if re.search(pat, line):
                    counts[cur][k] += 1
    for fn, c in counts.items():
        full = demangle(fn).replace("odis::(anonymous namespace)::", "").replace("(anonymous namespace)::", "").replace("void ", "")
        short = re.sub(r"\(.*", "", full)
        reg, stack, shared = usage.get(fn, (0, 0, 0))
        rows.append((obj.replace(".cu.o", ""), short, reg, stack, shared, c))
print("# built with: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false   (geodesicodis_b200/build.py)")
print(f"{'object':18s} {'kernel':46s} {'regs':>4s} {'stack':>5s} {'smem':>6s} {'instr':>6s} " + " ".join(f"{k:>11s}" for k, _ in PATS))
for obj, short, reg, stack, shared, c in rows:
    print(f"{obj:18s} {short[:46]:46s} {reg:4d} {stack:5d} {shared:6d} {c['instr']:6d} " + " ".join(f"{c[k]:11d}" for k, _ in PATS))
