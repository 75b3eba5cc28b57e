"""One markdown row per launch out of `ncu --page raw --csv` exports (tools/gpu_evidence.sh): duration, DRAM bytes, tensor pipe,
L2 hit rate, issue activity, XU pipe, registers.  Usage: python tools/ncu_raw_summary.py <raw.csv> [<raw.csv> ...]"""
import csv
import os
import sys

M = {"dur": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
     "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "l2": "lts__t_sector_hit_rate.pct",
     "issue": "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "xu": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
     "regs": "launch__registers_per_thread"}
SCALE = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "Tbyte": 1e6, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def main():
    print("| capture | kernel | grid x block | duration | DRAM read + write | tensor pipe | L2 hit | issue active | XU pipe | regs |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for path in sys.argv[1:]:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        col = {k: hdr.index(v) for k, v in M.items() if v in hdr}
        for r in rows[2:]:
            def val(k):
                if k not in col or r[col[k]] in ("", "n/a"):
                    return float("nan")
                return float(r[col[k]].replace(",", "")) * SCALE.get(units[col[k]], 1.0)
            name = r[hdr.index("Kernel Name")].replace("void ", "").split("(CUtensorMap")[0].split("(")[0]
            print(f"| `{os.path.basename(path)}` | `{name}` | {r[hdr.index('Grid Size')]} x {r[hdr.index('Block Size')]} | {val('dur'):.1f} us | "
                  f"{val('rd'):.0f} + {val('wr'):.0f} MB | {val('tensor'):.1f} % | {val('l2'):.1f} % | {val('issue'):.1f} % | {val('xu'):.1f} % | {val('regs'):.0f} |")


if __name__ == "__main__":
    main()
