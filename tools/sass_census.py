"""Per-kernel SASS census of libkosmosx_sm100.so: `cuobjdump -sass`, tcgen05 / TMEM / TMA mnemonics
counted per kernel.  Usage: python tools/sass_census.py > profiles/r2_sass_census.md   (runs without a GPU)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "kosmos-x_b200", "lib", "libkosmosx_sm100.so")
txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
ops = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "UBLKCP", "SYNCS", "HMMA", "MUFU.EX2", "FFMA2",
       "FADD2", "FMUL2", "ELECT"]
rows = []
for f in funcs:
    name = f.split("\n", 1)[0].strip()
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    short = re.sub(r"\(.*", "", dem).replace("void ", "").replace("kx::", "")
    n_inst = len(re.findall(r"/\*[0-9a-f]{4}\*/", f))
    cnt = {o: (f.count(o) if "." in o else len(re.findall(r"\b" + re.escape(o) + r"\b", f))) for o in ops}
    cnt["UTCHMMA.2CTA"] = f.count("UTCHMMA.2CTA")
    rows.append((short, n_inst, cnt))
tot = collections.Counter()
for _, _, c in rows:
    tot.update(c)
print("# SASS census of `kosmos-x_b200/lib/libkosmosx_sm100.so`\n")
print("`cuobjdump -sass` of the sm_100a cubin, instruction mnemonics counted per kernel (`tools/sass_census.py`).  `UTCHMMA` = tcgen05.mma,")
print("`.2CTA` = cta_group::2, `LDTM` / `STTM` = tcgen05.ld / st (TMEM), `UTMALDG` / `UTMASTG` / `UTMAREDG` = TMA tensor load / store / reduce-add,")
print("`UBLKCP` = cp.async.bulk (1-D), `SYNCS` = mbarrier ops, `HMMA` = legacy mma.sync (only the weight-streaming decode kernels, where the batch")
print("of <= 32 rows is far below tcgen05's 128-row tile).  No CUTLASS / CuTe / Triton / cuBLAS symbol is linked (`nm -D` shows only kx_*).\n")
print("Totals over %d kernels: " % len(rows) + ", ".join(f"{k} {v}" for k, v in tot.items() if v) + "\n")
print("| kernel | SASS instr | UTCHMMA (.2CTA) | LDTM | STTM | UTMALDG | UTMASTG | UTMAREDG | UBLKCP | SYNCS | HMMA | MUFU.EX2 | FFMA2/FADD2/FMUL2 |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for short, n, c in sorted(rows, key=lambda r: -(r[2]["UTCHMMA"] * 100000 + r[2]["HMMA"] * 1000 + r[1])):
    if n < 64 and not c["UTCHMMA"]:
        continue
    print(f"| `{short[:120]}` | {n} | {c['UTCHMMA']} ({c['UTCHMMA.2CTA']}) | {c['LDTM']} | {c['STTM']} | {c['UTMALDG']} | {c['UTMASTG']} | "
          f"{c['UTMAREDG']} | {c['UBLKCP']} | {c['SYNCS']} | {c['HMMA']} | {c['MUFU.EX2']} | {c['FFMA2']}/{c['FADD2']}/{c['FMUL2']} |")
