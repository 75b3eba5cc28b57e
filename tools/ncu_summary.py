"""Summarise an `ncu --csv` per-launch metrics log (tools/gpu_round.sh kernel_table) into one row per kernel of
libkosmosx_sm100.so: launches, median duration, DRAM bytes and achieved GB/s, tensor-pipe activity, registers, grid.
    python tools/ncu_summary.py gpurun_out/<tag>/kernel_table.csv > profiles/r1_kernel_table.md"""
import csv
import json
import os
import statistics
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [ln for ln in f if ln.startswith('"')]
rd = csv.DictReader(lines)
per = defaultdict(lambda: defaultdict(dict))          # (id) -> metric -> value
names = {}
for r in rd:
    k = r["ID"]
    names[k] = (r["Kernel Name"], r.get("Grid Size", ""), r.get("Block Size", ""))
    try:
        per[k][r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
    except ValueError:
        pass


def to_us(v, unit):
    return v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)


def to_bytes(v, unit):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)


groups = defaultdict(list)
for k, m in per.items():
    name = names[k][0]
    if "kx::" in name:
        short = name.split("kx::", 1)[1].split("(")[0]
    elif "_kernel" in name:                              # ncu's base-name demangling drops the namespace
        short = name.replace("void ", "").split("(")[0]
    else:
        continue
    groups[short].append((m, names[k]))
peaks = {}
try:
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")))
except Exception:
    pass
hbm = float(peaks.get("hbm_gbs", 6530.6))
print("Per-kernel ncu summary of one training step + forward at BASELINE shapes (B=8, T=2048; 2 decoder / 2 ViT layers), from `%s`." % os.path.basename(path))
print("DRAM GB/s = (dram__bytes_read + dram__bytes_write) / gpu__time_duration of the median launch; %% of the measured copy bandwidth (%.0f GB/s)." % hbm)
print("Tensor = sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed.  ncu serialises launches and runs them cold-cache: use shares, not absolutes.\n")
print("| kernel | launches | median us | max us | DRAM MB (rd+wr) | DRAM GB/s | % of HBM copy peak | tensor pipe % | SM throughput % | regs | grid x block |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
out = []
for short, items in groups.items():
    durs = sorted(to_us(*m["gpu__time_duration.sum"]) for m, _ in items if "gpu__time_duration.sum" in m)
    if not durs:
        continue
    med = statistics.median(durs)
    pick = min(items, key=lambda it: abs(to_us(*it[0]["gpu__time_duration.sum"]) - med))
    m, nm = pick
    d = to_us(*m["gpu__time_duration.sum"])
    by = to_bytes(*m.get("dram__bytes_read.sum", (0, "byte"))) + to_bytes(*m.get("dram__bytes_write.sum", (0, "byte")))
    gbs = by / (d * 1e-6) / 1e9 if d > 0 else 0
    tens = m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", (0, ""))[0]
    smt = m.get("sm__throughput.avg.pct_of_peak_sustained_elapsed", (0, ""))[0]
    regs = m.get("launch__registers_per_thread", (0, ""))[0]
    out.append((sum(durs), f"| `{short}` | {len(durs)} | {med:.1f} | {max(durs):.1f} | {by/1e6:.1f} | {gbs:.0f} | {100*gbs/hbm:.0f} | {tens:.1f} | {smt:.1f} | {int(regs)} | {nm[1]} x {nm[2]} |"))
for _, line in sorted(out, reverse=True):
    print(line)
